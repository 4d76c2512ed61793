#!/bin/bash
# parity tests of the tensor-core kernels against tools/libB.so, then bench A/B of tools/libA.so vs tools/libB.so
cp tools/libB.so namdanalyzer_b200/lib/libnda_b200.so
timeout 600 python -m pytest tests/test_gpu_tensor_filter.py tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -2
bash tools/ab_lib.sh 2>&1 | tail -6
bash tools/ab_rdf.sh 2>&1 | tail -4
