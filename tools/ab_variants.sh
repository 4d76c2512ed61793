#!/bin/bash
# A/B of library builds inside ONE box (same GPU, same clocks): the production library against the variants
# given on the command line (tools/variants/libnda_b200_<tag>.so, built by tools/build_variants.py).
#   bash tools/ab_variants.sh within 1000 floor1 floor2
what=$1; frames=$2; shift 2
for rep in 1 2 3; do
  echo -n "prod  "; python tools/profile_run.py $what $frames 2>&1 | grep -E "^(within|rdf|rdf3|dist)"
  for v in "$@"; do
    echo -n "$v  "; python tools/profile_run.py $what $frames --lib tools/variants/libnda_b200_$v.so 2>&1 | grep -E "^(within|rdf|rdf3|dist)"
  done
done
