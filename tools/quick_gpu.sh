#!/bin/bash
# tests + kernel timing of the production library (profiling aid)
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for wl in c2 c3; do FLEMINGO_DEBUG_LAUNCH=${DBG:-} python tools/time_variant.py flemingo_b200/_lib/libflemingo_b200.so $wl 2>&1 | tail -1; done
