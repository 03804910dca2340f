#!/bin/bash
L=flemingo_b200/_lib/libflemingo_b200.so
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
for wl in c2; do
for w in 16 8 6 4; do
  echo "warps=$w: "
  FLEMINGO_DEBUG_LAUNCH=1 FLEMINGO_FAST_WARPS=$w python tools/time_variant.py $L $wl 2>&1 | sort | uniq -c | tail -3
done
done
python tools/time_variant.py $L c3 2>&1 | tail -1
