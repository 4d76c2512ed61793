#!/bin/bash
# filter / occupancy sweep of the two persistent pair kernels (CUDA-event kernel times, no profiler)
for f in ${FILTERS:-rint fold fold2}; do
for c in ${CTAS:-4 5}; do
  echo "== NDA_FILTER=$f NDA_CTAS_PER_SM=$c"
  NDA_FILTER=$f NDA_CTAS_PER_SM=$c python tools/profile_run.py within
  NDA_FILTER=$f NDA_CTAS_PER_SM=$c python tools/profile_run.py rdf
done
done
