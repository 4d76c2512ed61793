#!/bin/bash
# A/B of the tensor-core filter against the packed-FP32 fold2 filter (CUDA-event kernel times)
for f in tc fold2a1; do
  echo "== NDA_FILTER=$f"
  NDA_FILTER=$f timeout 300 python tools/profile_run.py within 1000
  NDA_FILTER=$f timeout 300 python tools/profile_run.py rdf
  NDA_FILTER=$f timeout 300 python tools/profile_run.py rdf3
done
