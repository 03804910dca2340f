#!/bin/bash
# one-row (in-place) vs two-row mode of place_fast_kernel: parity tests in the forced one-row mode, then timings
L=flemingo_b200/_lib/libflemingo_b200.so
echo "forced one-row mode:"; FLEMINGO_FAST_ROWS=1 timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
echo "automatic mode:"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
for wl in c2 c3 c4; do
  for rows in 2 1 auto; do
    echo -n "$wl rows=$rows: "
    if [ $rows = auto ]; then FLEMINGO_DEBUG_LAUNCH=1 python tools/time_variant.py $L $wl 2>&1 | sort | uniq -c | sort -rn | cut -c1-150 | head -3 | tr '\n' '|'; echo
    else FLEMINGO_FAST_ROWS=$rows python tools/time_variant.py $L $wl 2>&1 | tail -1; fi
  done
done
