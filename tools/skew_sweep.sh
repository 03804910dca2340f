#!/bin/bash
L=flemingo_b200/_lib/libflemingo_b200.so
for wl in c2 c3; do
for spc in 64 128 192 256 320 512; do
  echo -n "spc=$spc: "
  FLEMINGO_FAST_SPC=$spc python tools/time_variant.py $L $wl 2>&1 | tail -1
done
done
