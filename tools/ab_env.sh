#!/bin/bash
# A/B of an environment switch inside one box: bench step time with and without $1
for rep in 1 2 3; do
  echo -n "default      "; python bench.py --steps 40 --warmup 5 --no-extras 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('step', round(d['ms_per_step'],4), 'kernel', round(d['roofline']['kernel_ms'],4))"
  echo -n "$1 "; env $1 python bench.py --steps 40 --warmup 5 --no-extras 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('step', round(d['ms_per_step'],4), 'kernel', round(d['roofline']['kernel_ms'],4))"
done
