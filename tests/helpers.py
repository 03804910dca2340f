"""Shared test helpers: rebuild golden organisms with the mirror objects, compare placements."""
import hashlib

import numpy as np

from flemingo_b200 import objects
from flemingo_b200.synthetic import CONF_CON, CONF_ORG, CONF_PSSM, CONF_SHAPE


def build_case_organism(case):
    org = objects.OrganismObject(case["name"], CONF_ORG)
    M = case["max_seq_length"]
    for node in case["nodes"]:
        if node["kind"] == "pssm":
            org.recognizers.append(objects.PssmObject(np.array(node["pwm"]), CONF_PSSM))
        elif node["kind"] == "shape":
            org.recognizers.append(objects.ShapeObject(node["type"], node["length"], CONF_SHAPE, node["mu"], node["sigma"]))
        else:
            org.connectors.append(objects.ConnectorObject(CONF_CON, M, node["mu"], node["sigma"]))
    org.flatten()
    return org


def sha(arr):
    return hashlib.sha256(np.ascontiguousarray(arr).tobytes()).hexdigest()


def positions_from_gaps(gaps, lengths):
    """recognizers_positions / connectors_positions as organism_object.py:1344-1355 derives them."""
    rec, con = [], []
    cur = int(gaps[0])
    n = len(lengths)
    for i in range(n):
        rec.append([cur, cur + int(lengths[i])])
        cur += int(lengths[i])
        if i < n - 1:
            con.append([cur, cur + int(gaps[i + 1])])
            cur += int(gaps[i + 1])
    return rec, con


def assert_placement_equals_golden(ref, energy, rec_scores, con_scores, gaps, lengths, where=""):
    """Bit-exact comparison against one golden placement record."""
    n = len(lengths)
    rec_pos, con_pos = positions_from_gaps(gaps, lengths)
    assert rec_pos == ref["recognizers_positions"], f"{where}: recognizer positions {rec_pos} != {ref['recognizers_positions']}"
    assert con_pos == ref["connectors_positions"], f"{where}: connector positions"
    assert float(energy) == ref["energy"], f"{where}: energy {float(energy)!r} != {ref['energy']!r}"
    assert [float(x) for x in rec_scores[:n]] == ref["recognizers_scores"], f"{where}: recognizer scores"
    assert [float(x) for x in con_scores[:n - 1]] == ref["connectors_scores"], f"{where}: connector scores"


def adversarial_workload(seed: int, n_org: int, S: int = 240, n_seq: int = 160):
    """Organisms and sequences chosen to stress the error-bound argument of the integer filter pass: zero-probability
    and near-uniform PWM columns, 2-decimal PWMs (exact score ties), tiny / huge connector sigmas and mus beyond the
    sequence, shape recognizers, low-complexity and repetitive sequences, ragged lengths."""
    from flemingo_b200 import synthetic
    rng = np.random.default_rng(seed)
    synthetic.ensure_null_models()

    def pssm(L):
        cols = []
        for _ in range(L):
            mode = rng.integers(0, 4)
            if mode == 0:
                p = rng.dirichlet([0.2] * 4)
            elif mode == 1:
                p = rng.dirichlet([5] * 4)
            elif mode == 2:
                p = rng.dirichlet([0.5] * 4)
                p[rng.integers(0, 4)] = 0
                p /= p.sum()
            else:
                c = np.maximum(rng.multinomial(100, rng.dirichlet([1] * 4)), 1)
                p = c / c.sum()
            p = np.round(p, 2) if rng.random() < 0.5 else p
            cols.append({"a": float(p[0]), "g": float(p[1]), "c": float(p[2]), "t": float(p[3])})
        return objects.PssmObject(np.array(cols), CONF_PSSM)

    orgs = []
    for o in range(n_org):
        n = int(rng.integers(2, 7))
        recs = [pssm(int(rng.integers(3, 13))) if rng.random() < 0.6 else synthetic.random_shape(rng) for _ in range(n)]
        cons = []
        for _ in range(n - 1):
            mu = float(rng.choice([0.0, rng.integers(0, 60), rng.uniform(0, 400)]))
            sigma = float(rng.choice([1e-3, rng.uniform(0.3, 3), rng.exponential(20), 500.0]))
            cons.append(synthetic.random_connector(rng, S, mu, sigma))
        orgs.append(synthetic.assemble(o, recs, cons))
    orgs = [o for o in orgs if o.sum_recognizer_lengths <= 150]

    def low_complexity(L):
        kind = rng.integers(0, 5)
        if kind == 0:
            return "".join(rng.choice(list("acgt"), L))
        if kind == 1:
            unit = "".join(rng.choice(list("acgt"), int(rng.integers(1, 7))))
            return (unit * (L // len(unit) + 1))[:L]
        if kind == 2:
            return "".join(rng.choice(list("ac"), L))
        if kind == 3:
            half = "".join(rng.choice(list("acgt"), L // 2))
            return (half + half + "a")[:L]
        s = list(rng.choice(list("acgt"), L))
        s[L // 3: L // 3 + 40] = ["a"] * 40
        return "".join(s)

    seqs = [low_complexity(int(rng.choice([S, S, S - 7, 160]))) for _ in range(n_seq)]
    return orgs, seqs
