#!/bin/bash
# pipeline-floor experiments of the tensor-core within kernel (results are WRONG by construction):
#   X1 = epilogue loads TMEM but skips the reduction, X2 = epilogue neither loads nor reduces
cp namdanalyzer_b200/lib/libnda_b200.so /tmp/libprod.so
for v in prod X1 X2; do
  if [ $v = prod ]; then cp /tmp/libprod.so namdanalyzer_b200/lib/libnda_b200.so; else cp tools/lib$v.so namdanalyzer_b200/lib/libnda_b200.so; fi
  echo -n "$v: "; python tools/profile_run.py within 1000 2>&1 | grep within
done
cp /tmp/libprod.so namdanalyzer_b200/lib/libnda_b200.so
