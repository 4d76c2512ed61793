#!/bin/bash
# A/B of two library builds (tools/libA.so vs tools/libB.so): within step, sparse RDF, dense grouped RDF
for rep in 1 2; do
for v in A B; do
  cp tools/lib$v.so namdanalyzer_b200/lib/libnda_b200.so
  echo -n "$v "; python bench.py --steps 40 --warmup 5 --no-extras 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('step', round(d['ms_per_step'],4), 'kernel', round(d['roofline']['kernel_ms'],4), 'hits', d['check']['hits_resident'])"
  echo -n "$v "; python tools/profile_run.py rdf 2>&1 | grep "rdf:"
  echo -n "$v "; python tools/profile_run.py rdf3 2>&1 | grep rdf3
done
done
