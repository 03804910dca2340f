"""Throughput of the placement path on the inputs the headline workloads do not exercise:

  * tie-heavy / low-complexity data (tests/helpers.adversarial_workload): how many pairs take the int32 second chance,
    how many reach the exact kernel, and what that costs;
  * homopolymer-rich and non-ACGT sequences (the exact kernel's territory);
  * a ragged dataset with >= 100 distinct sequence lengths (one launch chain per length class) against the same
    number of bases at one fixed length.

    python tools/bench_paths.py > profiles/r2_paths.txt          (on the GPU box)

Device-timed with CUDA events on the engine's stream (torch only for the events); every case is checked against the
exact-only path on a sample before it is timed.
"""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import torch  # noqa: E402

from flemingo_b200 import synthetic  # noqa: E402
from flemingo_b200.engine import Engine  # noqa: E402
from tests.helpers import adversarial_workload  # noqa: E402


def timed(eng, stream, sset, reps=5):
    eng.place(sset, fetch=False)
    eng.sync()
    eng.path_stats()
    best, host = 1e30, 1e30
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(stream):
            e0.record(stream)
            t0 = time.perf_counter()
            eng.place(sset, fetch=False)
            host = min(host, (time.perf_counter() - t0) * 1e3)   # time the host needs to plan and enqueue the launches
            e1.record(stream)
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    fast, exact, second = eng.path_stats(detail=True)
    return best, fast / reps, exact / reps, second / reps, host


def check(eng, orgs, seqs, n=6):
    sub_o, sub_s = orgs[:: max(1, len(orgs) // n)][:n], seqs[:: max(1, len(seqs) // 40)][:40]
    eng.upload_population(sub_o)
    sset = eng.upload_sequences(sub_s, 3)
    a = eng.place(sset, trace=True)
    eng.exact_only = True
    b = eng.place(sset, trace=True)
    eng.exact_only = False
    for k in ("energies", "gaps", "rec_scores", "con_scores"):
        assert np.array_equal(getattr(a, k), getattr(b, k), equal_nan=True), k


def case(eng, stream, name, orgs, seqs):
    check(eng, orgs, seqs)
    eng.upload_population(orgs)
    sset = eng.upload_sequences(seqs, 0)
    ms, fast, exact, second, host_ms = timed(eng, stream, sset)
    pairs = len(orgs) * len(seqs)
    eng.exact_only = True
    ms_exact = timed(eng, stream, sset, reps=2)[0]
    eng.exact_only = False
    print(f"{name:58s} {len(orgs):5d} org x {len(seqs):6d} seq  {ms:9.3f} ms  {pairs / ms / 1e3:8.2f} M placements/s   "
          f"filter pass {fast / pairs:6.1%}  second chance {second / pairs:7.3%}  exact kernel {exact / pairs:7.3%}   "
          f"host enqueue {host_ms:7.3f} ms   (everything on the exact kernel: {ms_exact:9.3f} ms)", flush=True)
    return ms


def main():
    torch.cuda.set_device(0)
    eng = Engine(0)
    stream = torch.cuda.ExternalStream(eng.stream_handle)
    rng = np.random.default_rng(5)
    print("# tools/bench_paths.py -- one placement launch chain per line, best of 5, CUDA events; sample checked against the exact-only path first")
    base_orgs = synthetic.pssm_organisms(256, 300, 3)
    base_seqs = synthetic.random_sequences(2000, 300, 1)
    t_base = case(eng, stream, "config 2 shape, random sequences (reference point)", base_orgs, base_seqs)

    orgs, seqs = adversarial_workload(11, 256, S=300, n_seq=2000)
    case(eng, stream, "adversarial organisms x low-complexity sequences", orgs, seqs)

    homo = []
    for _ in range(2000):
        s = list(rng.choice(list("acgt"), 300))
        a = int(rng.integers(0, 200))
        s[a:a + int(rng.integers(30, 100))] = rng.choice(list("acgt")) * 100
        homo.append("".join(s[:300]))
    case(eng, stream, "config 2 organisms x homopolymer-rich sequences", base_orgs, homo)

    nonacgt = []
    for k in range(2000):
        s = list(rng.choice(list("acgt"), 300))
        if k % 10 == 0:                          # every tenth sequence carries unknown bytes
            for p in rng.integers(0, 300, 3):
                s[p] = "n"
        nonacgt.append("".join(s))
    case(eng, stream, "config 2 organisms x 10 % sequences with N bytes", base_orgs, nonacgt)

    mixed = synthetic.mixed_organisms(256, 300, 3)
    t_fixed = case(eng, stream, "mixed organisms x fixed length 300", mixed, base_seqs)
    lengths = rng.integers(200, 301, 2000)
    ragged = ["".join(rng.choice(list("acgt"), int(L))) for L in lengths]
    t_rag = case(eng, stream, f"mixed organisms x ragged lengths ({len(set(lengths.tolist()))} distinct, 200..300)", mixed, ragged)
    same_len = int(np.mean(lengths))
    t_ref = case(eng, stream, f"mixed organisms x fixed length {same_len} (same number of bases)", mixed, synthetic.random_sequences(2000, same_len, 9))
    print(f"# ragged / fixed-length (same bases): {t_rag / t_ref:.2f}x")
    eng.close()


if __name__ == "__main__":
    main()
