"""Per-iteration wall time of the public end-to-end call (set_population_fitness), cold sets, config 2."""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from flemingo_b200 import synthetic
from flemingo_b200.engine import Engine
from flemingo_b200.objects import set_population_fitness
orgs, pos, neg = synthetic.make_workload(sys.argv[1] if len(sys.argv) > 1 else "c2")
eng = Engine(0)
for reuse in (False, True):
    ts = []
    for i in range(12):
        t0 = time.perf_counter()
        set_population_fitness(orgs, pos, neg, "Welch", 0.2, engine=eng, reuse_sets=reuse)
        ts.append((time.perf_counter() - t0) * 1e3)
    print("reuse_sets=%s: " % reuse + " ".join("%.2f" % t for t in ts))
import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
for i in range(10):
    set_population_fitness(orgs, pos, neg, "Welch", 0.2, engine=eng, reuse_sets=False)
pr.disable()
pstats.Stats(pr).sort_stats("cumtime").print_stats(22)
