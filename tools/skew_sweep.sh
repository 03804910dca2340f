#!/bin/bash
L=flemingo_b200/_lib/libflemingo_b200.so
for cfg in "2 16" "1 16" "1 8" "1 4"; do
  set -- $cfg
  echo -n "rows=$1 warps=$2: "
  FLEMINGO_DEBUG_LAUNCH=1 FLEMINGO_FAST_ROWS=$1 FLEMINGO_FAST_WARPS=$2 python tools/time_variant.py $L c2 2>&1 | sort | uniq -c | sort -rn | cut -c1-140 | head -2 | tr '\n' '|'; echo
done
