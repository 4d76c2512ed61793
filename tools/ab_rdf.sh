#!/bin/bash
# A/B of two library builds on the RDF cases (tools/libA.so vs tools/libB.so)
for rep in 1; do
for v in A B; do
  cp tools/lib$v.so namdanalyzer_b200/lib/libnda_b200.so
  echo -n "$v "; python tools/profile_run.py rdf 2>&1 | grep "rdf:"
  echo -n "$v "; NDA_TC_MAX_FRACTION=1 python tools/profile_run.py rdf3 2>&1 | grep rdf3
done
done
