"""Per-phase SM-clock breakdown of place_fast_kernel (profiling build with -DFL_PHASE_CLOCKS, see tools/phase_clocks.sh).
Usage: python tools/phase_clocks.py [c2|c3|tight|tight-mixed]   -- the library must have been built with -DFL_PHASE_CLOCKS."""
import sys
sys.path.insert(0, ".")
import torch  # noqa: F401  (CUDA runtime initialisation order as in bench.py)
from pathlib import Path
from flemingo_b200 import synthetic, _cabi
_cabi.LIB_PATH = Path("tools/_phase_lib/libflemingo_b200.so").resolve()   # the profiling copy, not the product
from flemingo_b200.engine import Engine

wl = sys.argv[1] if len(sys.argv) > 1 else "c2"
if wl == "tight":     # config 2 with connector sigmas from config.json's bounds [1, 3] (tools/bench_tight.py)
    _, pos, neg = synthetic.make_workload("c2")
    orgs = synthetic.pssm_organisms(256, 300, 3, sigma_range=(1.0, 3.0))
elif wl == "tight-mixed":
    _, pos, neg = synthetic.make_workload("c2")
    orgs = synthetic.mixed_organisms(256, 300, 3, sigma_range=(1.0, 3.0))
else:
    orgs, pos, neg = synthetic.make_workload(wl)
eng = Engine(0)
for _ in range(2):
    eng.evaluate(orgs, pos, neg, "Welch", 0.2)
eng.path_stats()           # resets the counters (prints the warm-up's)
print("---- steady state ----", file=sys.stderr, flush=True)
import os
for _ in range(5):
    eng.evaluate(orgs, pos, neg, "Welch", 0.2)
eng.path_stats()
# timeline of CTA 0 during ONE placement launch
os.environ["FLEMINGO_TIMELINE"] = f"gpurun_out/timeline_{wl}.txt"
sp = eng.upload_sequences(pos, 0)
eng.place(sp, trace=False, fetch=False)
eng.sync()
print(eng.path_stats())
