"""Times fl_place_batch (both sets, energies only) of a workload with an alternative build of the library.
Usage: python tools/time_variant.py <path/to/libflemingo_b200.so> [c2|c3]   (profiling aid, not part of the product)"""
import sys
sys.path.insert(0, ".")
from pathlib import Path
import torch
from flemingo_b200 import synthetic, _cabi
_cabi.LIB_PATH = Path(sys.argv[1]).resolve()
from flemingo_b200.engine import Engine
from flemingo_b200.population import pack_population

wl = sys.argv[2] if len(sys.argv) > 2 else "c2"
orgs, pos, neg = synthetic.make_workload(wl)
eng = Engine(0)
sp = eng.upload_sequences(pos, 0)
sn = eng.upload_sequences(neg, 1)
eng.upload_population(pack_population(orgs))
for _ in range(3):
    eng.place(sp, trace=False, fetch=False); eng.place(sn, trace=False, fetch=False)
torch.cuda.synchronize()
best = 1e9
import time
for _ in range(5):
    eng.sync(); t0 = time.perf_counter()
    eng.place(sp, trace=False, fetch=False); eng.place(sn, trace=False, fetch=False)
    eng.sync(); best = min(best, (time.perf_counter() - t0) * 1e3)
print(f"{sys.argv[1]} {wl}: place pos+neg {best:.3f} ms (host clock around sync, best of 5)")
