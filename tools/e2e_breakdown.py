"""Where the end-to-end time of one fitness evaluation goes (host packing, uploads, device work)."""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
from flemingo_b200 import synthetic
from flemingo_b200.engine import Engine
from flemingo_b200.population import pack_population

orgs, pos, neg = synthetic.make_workload(sys.argv[1] if len(sys.argv) > 1 else "c2")
eng = Engine(0)
eng.evaluate(orgs, pos, neg)
T = {}
def tick(name, t0):
    eng.sync(); T.setdefault(name, []).append((time.perf_counter() - t0) * 1e3)
for rep in range(10):
    t = time.perf_counter(); packed = pack_population(orgs); tick("pack_population", t)
    t = time.perf_counter(); eng.upload_population(packed); tick("upload_population", t)
    t = time.perf_counter(); sp = eng.upload_sequences(pos, 0); tick("upload_sequences(pos)", t)
    t = time.perf_counter(); sn = eng.upload_sequences(neg, 1); tick("upload_sequences(neg)", t)
    t = time.perf_counter(); eng.place(sp, fetch=False); tick("place(pos)", t)
    t = time.perf_counter(); eng.place(sn, fetch=False); tick("place(neg)", t)
    t = time.perf_counter(); eng.fitness(sp, sn); tick("fitness", t)
    t = time.perf_counter(); eng.evaluate(orgs, pos, neg, reuse_sets=False); tick("evaluate (all of the above, async)", t)
for k, v in T.items():
    print(f"{k:40s} {np.median(v):8.3f} ms")
