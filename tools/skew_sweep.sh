#!/bin/bash
L=flemingo_b200/_lib/libflemingo_b200.so
for wl in c2 c3; do
for w in 16 7 6; do
  echo "warps=$w: "
  FLEMINGO_DEBUG_LAUNCH=1 FLEMINGO_FAST_WARPS=$w python tools/time_variant.py $L $wl 2>&1 | sort | uniq -c | tail -4
done
done
