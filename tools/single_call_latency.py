"""Latency of the reference-facing single-placement call (OrganismObject.get_placement -> fl_calculate)."""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from flemingo_b200 import synthetic
from oracle import oracle

for name, orgs in (("3 PSSM x len 12", synthetic.pssm_organisms(4, 300, 3)), ("mixed", [o for o in synthetic.mixed_organisms(30, 300, 3) if len(o.recognizer_types) == 6][:4])):
    seqs = synthetic.random_sequences(50, 300, 1)
    orgs[0].get_placement(seqs[0])
    t = time.perf_counter(); n = 0
    for o in orgs:
        for s in seqs:
            o.get_placement(s); n += 1
    gpu = (time.perf_counter() - t) / n * 1e3
    t = time.perf_counter()
    for o in orgs:
        for s in seqs:
            oracle.place(o, s, "reference" if oracle.reference_module() else "restated")
    cpu = (time.perf_counter() - t) / n * 1e3
    print(f"{name:18s} get_placement via fl_calculate: {gpu:.3f} ms/call   reference C extension: {cpu:.3f} ms/call")
