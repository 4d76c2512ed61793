#!/bin/bash
# occupancy sweep of the two persistent pair kernels (CUDA-event kernel times, no profiler)
for c in 1 2 3 4 5; do
  echo "== NDA_CTAS_PER_SM=$c"
  NDA_CTAS_PER_SM=$c python tools/profile_run.py within
  NDA_CTAS_PER_SM=$c python tools/profile_run.py rdf
done
