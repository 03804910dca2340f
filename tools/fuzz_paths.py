"""Fuzz check of the integer passes against the exact kernel: random populations (PSSM-only of every shape the packed
16-bit pass takes, mixed, adversarial), random and low-complexity sequences, several lengths -- every pair, energies,
offsets and node scores, filter-and-refine (as planned, and with the packed pass forced on everything) == exact-only.

    python tools/fuzz_paths.py [n_rounds]        (on the GPU box; prints one line per round)
"""
import os
import sys

import numpy as np

sys.path.insert(0, ".")
from flemingo_b200 import objects, synthetic  # noqa: E402
from flemingo_b200.engine import Engine  # noqa: E402
from tests.helpers import adversarial_workload  # noqa: E402


# as planned / one arithmetic forced on every organism / the kernels with pruned stages forced on every organism
MODES = ("auto", "forced-16", "forced-32", "pruned-16", "pruned-32", "pruned-1row", "unpruned")
MODE_ENV = {"auto": {}, "forced-16": {"FLEMINGO_FAST_BITS": "16"}, "forced-32": {"FLEMINGO_FAST_BITS": "32"},
            "pruned-16": {"FLEMINGO_FAST_BITS": "16", "FLEMINGO_PRUNE": "2"}, "pruned-32": {"FLEMINGO_FAST_BITS": "32", "FLEMINGO_PRUNE": "2"},
            "pruned-1row": {"FLEMINGO_PRUNE": "2", "FLEMINGO_FAST_ROWS": "1"}, "unpruned": {"FLEMINGO_PRUNE": "0"}}


def random_population(rng, S, ragged):
    orgs = []
    for o in range(int(rng.integers(8, 40))):
        n = int(rng.integers(2, 7))
        kind = rng.integers(0, 3)
        recs = []
        for _ in range(n):
            if kind == 0 or (kind == 1 and rng.random() < 0.5):
                L = int(rng.integers(3, 13))
                if rng.random() < 0.2:      # PWM columns with zero probabilities: entries of -33.22
                    cols = []
                    for _ in range(L):
                        p = rng.dirichlet([0.4] * 4)
                        p[rng.integers(0, 4)] = 0.0
                        p = np.round(p / p.sum(), 2) if rng.random() < 0.5 else p / p.sum()
                        cols.append({"a": float(p[0]), "g": float(p[1]), "c": float(p[2]), "t": float(p[3])})
                    recs.append(objects.PssmObject(np.array(cols), synthetic.CONF_PSSM))
                else:
                    recs.append(synthetic.random_pssm(rng, L))
            else:
                recs.append(synthetic.random_shape(rng))
        cons = []
        for _ in range(n - 1):
            mu = float(rng.choice([rng.poisson(25), rng.uniform(0, S), 0.0]))
            # sigma == 0 with sequences shorter than the maximum is a ZeroDivisionError in the reference's rescale (and here)
            sigma = float(rng.choice([rng.exponential(S / 12), rng.uniform(0.3, 3), 1e-3 if ragged else 0.0, 1e-3]))
            cons.append(synthetic.random_connector(rng, S, mu, sigma))
        org = synthetic.assemble(o, recs, cons)
        if org.sum_recognizer_lengths <= S - 2:
            orgs.append(org)
    return orgs


def random_sequences(rng, S, n, ragged):
    out = []
    for _ in range(n):
        L = int(rng.choice([S, S, S, S - 1, max(S - 13, S // 2)])) if ragged else S
        kind = rng.integers(0, 6)
        if kind <= 2:
            s = "".join(rng.choice(list("acgt"), L))
        elif kind == 3:
            unit = "".join(rng.choice(list("acgt"), int(rng.integers(1, 9))))
            s = (unit * (L // len(unit) + 1))[:L]
        elif kind == 4:
            s = "".join(rng.choice(list("ac"), L))
        else:
            s = list(rng.choice(list("acgt"), L))
            s[L // 4: L // 4 + 30] = ["g"] * 30
            s = "".join(s[:L])
        out.append(s)
    return out


def main():
    rounds = int(sys.argv[1]) if len(sys.argv) > 1 else 30
    synthetic.ensure_null_models()
    eng = Engine(0)
    total = 0
    for r in range(rounds):
        rng = np.random.default_rng(1000 + r)
        S = int(rng.choice([40, 64, 97, 150, 200, 300, 333, 500]))
        if r % 7 == 6:
            orgs, seqs = adversarial_workload(r, 40, S=max(S, 200), n_seq=60)
        else:
            ragged = bool(r % 2)
            orgs = random_population(rng, S, ragged)
            seqs = random_sequences(rng, S, int(rng.integers(20, 80)), ragged)
            longest = max(o.sum_recognizer_lengths for o in orgs)
            seqs = [s for s in seqs if len(s) >= longest]
        eng.upload_population(orgs)
        sset = eng.upload_sequences(seqs, 0)
        eng.exact_only = True
        try:
            want = eng.place(sset, trace=True)
        except (ZeroDivisionError, OverflowError) as exc:
            # a connector the reference's adjust_scores cannot rescale either (sigma 0 / denormal, mu beyond a sequence)
            print(f"round {r:3d} skipped: {type(exc).__name__} in the rescale, as in the reference", flush=True)
            eng.exact_only = False
            continue
        eng.exact_only = False
        stats = []
        for mode in MODES:
            env = MODE_ENV[mode]
            os.environ.update(env)
            eng.path_stats()
            got = eng.place(sset, trace=True)
            stats.append(eng.path_stats(detail=True))
            for key in env:
                os.environ.pop(key, None)
            for k in ("energies", "gaps", "rec_scores", "con_scores"):
                a, b = getattr(got, k), getattr(want, k)
                if not np.array_equal(a, b, equal_nan=True):
                    bad = np.argwhere(a != b)[:3]
                    raise SystemExit(f"round {r} S={S} mode {mode}: {k} differs at {bad.tolist()}")
        total += want.energies.size
        print(f"round {r:3d} S={S:3d} {len(orgs):3d} organisms x {len(seqs):3d} sequences: equal in all modes; "
              f"(filter pairs, exact kernel, second chance) " + " ".join(f"{m} {st}" for m, st in zip(MODES, stats)), flush=True)
    print(f"fuzz ok: {total} pairs x {len(MODES)} modes bit-identical to the exact kernel")


if __name__ == "__main__":
    main()
