"""Stress: the filter+refine path against the exact-only path (both on the GPU) on adversarial inputs
(tests/helpers.adversarial_workload).  Any difference (energies, offsets, node scores) would be a bug in the error-bound
argument of place_fast_kernel.   usage: python tools/stress_compare.py [seed] [n_organisms]"""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
from flemingo_b200.engine import Engine
from tests.helpers import adversarial_workload

orgs, seqs = adversarial_workload(int(sys.argv[1]) if len(sys.argv) > 1 else 0, int(sys.argv[2]) if len(sys.argv) > 2 else 600)
eng = Engine(0)
eng.upload_population(orgs)
ss = eng.upload_sequences(seqs, 0)
eng.place(ss, trace=True)   # warm-up: lazy kernel loading, shape bins
eng.path_stats()
t = time.perf_counter(); eng.exact_only = False; a = eng.place(ss, trace=True); fast_pairs, back = eng.path_stats(); t1 = time.perf_counter() - t
t = time.perf_counter(); eng.exact_only = True; b = eng.place(ss, trace=True); t2 = time.perf_counter() - t
same = [np.array_equal(getattr(a, k), getattr(b, k)) for k in ("energies", "gaps", "rec_scores", "con_scores")]
print(f"{len(orgs)} organisms x {len(seqs)} sequences = {a.energies.size} pairs; filter+refine {t1:.3f} s (handed back {back} of {fast_pairs}), "
      f"exact-only {t2:.3f} s; identical: {same}")
sys.exit(0 if all(same) else 1)
