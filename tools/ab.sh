#!/bin/bash
# A/B of the candidate-path variants inside one box
for v in "" "NDA_NO_BIN_TABLE=1" "NDA_STRICT_SLOW=1" "NDA_NO_BIN_TABLE=1 NDA_STRICT_SLOW=1"; do
  echo "== $v"
  env $v python tools/e2e_all.py 2>&1 | grep "cfg1\|cfg3 grouped\|cfg5"
done
