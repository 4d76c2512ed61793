#!/bin/bash
# A/B of two builds of the library inside one box: tools/libA.so vs tools/libB.so, bench step time
for rep in 1 2 3; do
for v in A B; do
  cp tools/lib$v.so namdanalyzer_b200/lib/libnda_b200.so
  python bench.py --steps 40 --warmup 5 --no-extras 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$v', round(d['ms_per_step'],4), round(d['roofline']['kernel_ms'],4), '%.3e' % d['e2e']['value'])"
done
done
