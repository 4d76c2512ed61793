#!/bin/bash
# Pipeline-floor experiments of the tensor-core within kernel (results are WRONG by construction):
#   X1 = epilogue loads TMEM but skips the reduction, X2 = epilogue neither loads nor reduces.
# Build the two variants first (here or on the GPU box, nvcc is in the image):
#   NDA_TC_EXPERIMENT=1 python -m namdanalyzer_b200.build --force && cp namdanalyzer_b200/lib/libnda_b200.so tools/libX1.so
#   NDA_TC_EXPERIMENT=2 python -m namdanalyzer_b200.build --force && cp namdanalyzer_b200/lib/libnda_b200.so tools/libX2.so
#   python -m namdanalyzer_b200.build --force
# Measured on a B200 (round 1): production 0.670 ms, X1 0.512 ms, X2 0.510 ms for config 2, 1000 frames.
cp namdanalyzer_b200/lib/libnda_b200.so /tmp/libprod.so
for v in prod X1 X2; do
  if [ $v = prod ]; then cp /tmp/libprod.so namdanalyzer_b200/lib/libnda_b200.so; else cp tools/lib$v.so namdanalyzer_b200/lib/libnda_b200.so; fi
  echo -n "$v: "; python tools/profile_run.py within 1000 2>&1 | grep within
done
cp /tmp/libprod.so namdanalyzer_b200/lib/libnda_b200.so
