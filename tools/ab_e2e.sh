#!/bin/bash
# same-box A/B of an environment switch on the end-to-end within call: tools/ab_e2e.sh NDA_NO_TAPER=1
for rep in 1 2 3; do
  echo -n "default      "; python tools/e2e_one.py 1000 30
  echo -n "$1 "; env $1 python tools/e2e_one.py 1000 30
done
NDA_DEBUG=2 python tools/e2e_one.py 1000 3 2>&1 | tail -4
