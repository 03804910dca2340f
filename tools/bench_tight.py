"""Placement throughput on organisms with TIGHT connectors -- what fitted organisms look like (the reference's bundled
src/organism.json has sigma = 0.44 on every connector; its config.json bounds sigma to [1, 3]) -- with and without
the pruned DP stages (dp_stage_pruned), next to the factory's initial distribution sigma ~ Exp(S/12) of configs 2-5.

    python tools/bench_tight.py > profiles/r2_pruned_stages.txt          (on the GPU box; any argument: 300 bp cases only)

Device-timed with CUDA events on the engine's stream, best of 5; every case is first checked pair by pair against
the exact kernel on a sample.
"""
import os
import sys

import numpy as np

sys.path.insert(0, ".")
import torch  # noqa: E402

from flemingo_b200 import synthetic  # noqa: E402
from flemingo_b200.engine import Engine  # noqa: E402


def timed(eng, stream, sset, reps=5):
    eng.place(sset, fetch=False)
    eng.sync()
    best = 1e30
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(stream):
            e0.record(stream)
            eng.place(sset, fetch=False)
            e1.record(stream)
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


def check(eng, orgs, seqs):
    sub_o, sub_s = orgs[:: max(1, len(orgs) // 8)][:8], seqs[:: max(1, len(seqs) // 64)][:64]
    eng.upload_population(sub_o)
    sset = eng.upload_sequences(sub_s, 3)
    eng.exact_only = True
    want = eng.place(sset, trace=True)
    eng.exact_only = False
    for prune in ("0", "1", "2"):
        os.environ["FLEMINGO_PRUNE"] = prune
        got = eng.place(sset, trace=True)
        for k in ("energies", "gaps", "rec_scores", "con_scores"):
            assert np.array_equal(getattr(got, k), getattr(want, k), equal_nan=True), (prune, k)
    os.environ.pop("FLEMINGO_PRUNE", None)


def case(eng, stream, name, orgs, seqs):
    check(eng, orgs, seqs)
    eng.upload_population(orgs)
    sset = eng.upload_sequences(seqs, 0)
    pairs = len(orgs) * len(seqs)
    ms = {}
    for prune in ("0", "1", "2"):
        os.environ["FLEMINGO_PRUNE"] = prune
        eng.path_stats()
        ms[prune] = timed(eng, stream, sset)
        fast, exact, second = eng.path_stats(detail=True)
    os.environ.pop("FLEMINGO_PRUNE", None)
    print(f"{name:64s} {len(orgs):5d} org x {len(seqs):6d} seq   full stages {ms['0']:9.3f} ms ({pairs / ms['0'] / 1e3:7.1f} M/s)   "
          f"pruned where tight {ms['1']:9.3f} ms ({pairs / ms['1'] / 1e3:7.1f} M/s, {ms['0'] / ms['1']:4.2f}x)   "
          f"pruned kernels forced {ms['2']:9.3f} ms", flush=True)


def main():
    torch.cuda.set_device(0)
    eng = Engine(0)
    stream = torch.cuda.ExternalStream(eng.stream_handle)
    print("# tools/bench_tight.py -- one placement launch chain per line, best of 5, CUDA events; a sample of every case is checked "
          "against the exact kernel with FLEMINGO_PRUNE = 0, 1, 2 first")
    seqs = synthetic.random_sequences(2000, 300, 1)
    for label, sr in (("sigma = 0.44 (bundled organism.json)", (0.44, 0.44)), ("sigma ~ U[1, 3] (config.json bounds)", (1.0, 3.0)),
                      ("sigma ~ U[3, 8]", (3.0, 8.0)), ("sigma ~ Exp(25) (factory, = config 2)", None)):
        case(eng, stream, "3 PSSM x 12 bp, " + label, synthetic.pssm_organisms(256, 300, 3, sigma_range=sr), seqs)
    for label, sr in (("sigma ~ U[1, 3]", (1.0, 3.0)), ("sigma ~ Exp(25) (factory, = config 3 organisms)", None)):
        case(eng, stream, "mixed PSSM / shape, " + label, synthetic.mixed_organisms(256, 300, 3, sigma_range=sr), seqs)
    if len(sys.argv) > 1:     # only the first N cases
        eng.close()
        return
    long_seqs = synthetic.random_sequences(500, 2000, 1)
    for label, sr in (("sigma ~ U[1, 3]", (1.0, 3.0)), ("sigma ~ Exp(167) (factory, = config 4)", None)):
        case(eng, stream, "3 PSSM x 12 bp on 2 kb, mu ~ U[0, 1500], " + label,
             synthetic.pssm_organisms(64, 2000, 3, mu_range=(0, 1500), sigma_range=sr), long_seqs)
    eng.close()


if __name__ == "__main__":
    main()
