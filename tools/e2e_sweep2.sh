#!/bin/bash
for cs in 1 2; do for ch in 4 6 8; do
NDA_COPY_STREAMS=$cs NDA_WITHIN_CHUNKS=$ch PIN=1 python tools/e2e_sweep.py child 2>&1 | grep chunks | sed "s/^/copystreams=$cs /"
done; done
