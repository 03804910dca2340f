"""CPU: the batched table builders behind the C ABI (fl_build_connector_tables, fl_round_decimals,
fl_build_shape_masses + numpy logs) produce the same doubles as the reference's per-object Python loops.

The loops restated below follow connector_object.py:175-206, pssm_object.py:368-400 and shape_object.py:156-180
operation by operation (scalar math.erf / np.log2 / np.log / round); where /root/reference exists the reference's own
classes are used as well.  No CUDA call is made: the builders are host code."""
import json
import math
import sys
from pathlib import Path

import numpy as np
import pytest

from flemingo_b200 import builders

REF = Path("/root/reference")


def _norm_cdf(x, mu, sigma):
    if sigma == 0:
        return 0 if x < mu else (1 if x > mu else 0.5)
    return (1.0 + math.erf(((x - mu) / abs(sigma)) / math.sqrt(2.0))) / 2.0


def _connector_loop(mu, sigma, M, pc):
    pdfs, cdfs = [], []
    prev = first = _norm_cdf(-0.5, mu, sigma)
    with np.errstate(divide="ignore", invalid="ignore"):
        for j in range(1, M + 1):
            cur = _norm_cdf(j - 0.5, mu, sigma)
            cdfs.append(np.double(np.log2(prev - first + (pc * j))))
            pdfs.append(np.double(np.log2(cur - prev + pc)))
            prev = cur
    return np.array(pdfs), np.array(cdfs)


def test_connector_tables_bit_identical_to_the_scalar_loop():
    rng = np.random.default_rng(5)
    cases = [(25.0, 25.0), (0.0, 0.0), (7.0, 0.0), (3.5, 1e-3), (298.0, 0.4), (1e4, 3.0), (12.0, 1e-120), (40.0, 500.0)]
    cases += [(float(rng.poisson(25)), float(rng.exponential(25))) for _ in range(40)]
    cases += [(float(rng.uniform(0, 1500)), float(rng.exponential(160))) for _ in range(10)]
    for M in (1, 7, 300):
        mu = np.array([c[0] for c in cases])
        sigma = np.array([c[1] for c in cases])
        pdfs, cdfs = builders.connector_tables(mu, sigma, M, 1e-15)
        assert pdfs.shape == (len(cases), M)
        for i, (m, s) in enumerate(cases):
            want_p, want_c = _connector_loop(m, s, M, 1e-15)
            assert pdfs[i].tobytes() == want_p.tobytes(), (M, m, s)
            assert cdfs[i].tobytes() == want_c.tobytes(), (M, m, s)
    # threaded and single-threaded runs agree
    a = builders.connector_tables(mu, sigma, 300, 1e-15, n_threads=1)
    b = builders.connector_tables(mu, sigma, 300, 1e-15, n_threads=0)
    assert a[0].tobytes() == b[0].tobytes() and a[1].tobytes() == b[1].tobytes()


def test_round_decimals_is_python_round():
    rng = np.random.default_rng(6)
    vals = np.concatenate([rng.uniform(-40, 3, 20000), np.log2(4.0 * rng.integers(0, 101, 20000) / 100 + 1e-10),
                           np.array([0.125, 0.375, 2.675, -2.675, 1.005, -0.005, 0.0, -0.0, 1e-9, 123456.785, np.inf, -np.inf, np.nan]),
                           (rng.integers(-4000, 300, 5000) + 0.5) / 100.0])
    got = builders.round_decimals(vals.copy(), 2)
    want = np.array([round(float(v), 2) if math.isfinite(v) else v for v in vals])
    assert got.tobytes() == want.tobytes()


def test_pssm_values_match_recalculate_pssm_loop():
    rng = np.random.default_rng(8)
    counts = rng.multinomial(100, [0.25] * 4, size=500)
    counts[::7, 0] = 0
    p = counts / counts.sum(1, keepdims=True)
    got = builders.pssm_values(p, 1e-10)
    want = np.array([[round(float(np.log2(4.0 * x + 1e-10)), 2) for x in row] for row in p])
    assert got.tobytes() == want.tobytes()


def test_shape_alt_models_match_the_scalar_loop(golden_null_models):
    rng = np.random.default_rng(9)
    mus, sigmas, edges_list, want = [], [], [], []
    for kind in ("mgw", "prot", "roll", "helt"):
        for L in (5, 7, 9):
            edges = golden_null_models[kind][L]["bins"]
            for sigma in (0.0, float(rng.exponential((edges[-1] - edges[0]) / 12)), 1e-6):
                mu = float(edges[3]) if sigma == 0.0 else float(rng.uniform(edges[0], edges[-1]))
                nb = len(edges) - 1
                auc = _norm_cdf(edges[-1], mu, sigma) - _norm_cdf(edges[0], mu, sigma)
                auc += nb * 1e-100
                with np.errstate(divide="ignore"):
                    log_auc = np.log(auc)
                    row = []
                    for i in range(nb):
                        if sigma != 0:
                            mass = _norm_cdf(edges[i + 1], mu, sigma) - _norm_cdf(edges[i], mu, sigma)
                        else:
                            mass = 1 if edges[i] == mu else 0
                        row.append(np.log(mass + 1e-100) - log_auc)
                mus.append(mu)
                sigmas.append(sigma)
                edges_list.append(edges)
                want.append(np.array(row))
    got = builders.shape_alt_models(mus, sigmas, edges_list, 1e-100)
    for g, w in zip(got, want):
        assert g.tobytes() == w.tobytes()


@pytest.mark.skipif(not (REF / "src" / "objects").exists(), reason="needs /root/reference")
def test_builders_match_the_reference_classes(golden_null_models):
    sys.path.insert(0, str(REF / "src"))
    try:
        from objects import shape_object
        from objects.connector_object import ConnectorObject
        from objects.pssm_object import PssmObject
        from objects.shape_object import ShapeObject
    finally:
        sys.path.remove(str(REF / "src"))
    conf = json.load(open(REF / "src" / "config.json"))
    shape_object.null_models = golden_null_models
    rng = np.random.default_rng(10)
    for _ in range(25):
        M = int(rng.integers(5, 320))
        mu, sigma = float(rng.poisson(25)), float(rng.choice([0.0, rng.exponential(M / 12)]))
        ref = ConnectorObject(conf["connector"], M, mu, sigma)
        pdfs, cdfs = builders.connector_tables([mu], [sigma], M, ref.pseudo_count)
        assert np.array(ref.stored_pdfs).tobytes() == pdfs[0].tobytes()
        assert np.array(ref.stored_cdfs).tobytes() == cdfs[0].tobytes()
    for _ in range(25):
        pwm = []
        for _ in range(int(rng.integers(3, 13))):
            c = np.maximum(rng.multinomial(100, rng.dirichlet([1] * 4)), rng.integers(0, 2))
            q = c / c.sum()
            pwm.append({"a": float(q[0]), "g": float(q[1]), "c": float(q[2]), "t": float(q[3])})
        ref = PssmObject(np.array(pwm), conf["pssm"])
        got = builders.pssm_values(np.array([[col[b] for b in "agct"] for col in pwm]), ref.pseudo_count)
        want = np.array([[col[b] for b in "agct"] for col in ref.pssm])
        assert got.tobytes() == want.tobytes()
    for kind in ("mgw", "prot", "roll", "helt"):
        for L in (5, 8):
            edges = golden_null_models[kind][L]["bins"]
            mu, sigma = float(rng.uniform(edges[0], edges[-1])), float(rng.exponential((edges[-1] - edges[0]) / 12))
            ref = ShapeObject(kind, L, conf["shape"], mu, sigma)
            got = builders.shape_alt_models([ref._mu], [ref._sigma], [ref.edges], ref.pseudo_count)[0]
            assert np.array(ref.alt_model).tobytes() == got.tobytes()


def test_flatten_on_array_twins_equals_flatten_on_lists(golden_null_models):
    """flatten_organism concatenates the array twins the builders leave next to every table; an organism whose tables
    are plain lists (built elsewhere) goes through the conversion path.  Both give the same flat arrays, and a deep copy
    -- the reference's clone_organism -- keeps tables, twins and their identity."""
    import copy

    from flemingo_b200 import builders, synthetic
    from flemingo_b200.flatten import flatten_organism
    orgs = synthetic.mixed_organisms(30, 200, 17) + synthetic.pssm_organisms(5, 200, 18)
    names = ("recognizers_flat", "recognizer_lengths", "recognizer_models", "recognizer_bin_edges", "recognizer_bin_nums",
             "connectors_scores_flat")
    for org in orgs:
        want = {k: getattr(org, k).copy() for k in names}
        clone = copy.deepcopy(org)
        for con in clone.connectors:
            assert isinstance(con.stored_pdfs, builders.TableList) and con.stored_pdfs.np is not None
            assert con.stored_pdfs == list(con.stored_pdfs.np)
        # strip every twin: plain lists, no cached PSSM rows
        for con in clone.connectors:
            con.stored_pdfs, con.stored_cdfs = list(con.stored_pdfs), list(con.stored_cdfs)
        for rec in clone.recognizers:
            if rec.get_type() == 'p':
                rec._pssm_flat = None
            else:
                rec.alt_model = list(rec.alt_model)
        flatten_organism(clone)
        for k in names:
            got = getattr(clone, k)
            assert got.dtype == want[k].dtype and got.tobytes() == want[k].tobytes(), k
        assert clone.recognizer_types == org.recognizer_types and clone.sum_recognizer_lengths == org.sum_recognizer_lengths
    # a PSSM twin is only trusted while `pssm` is the array it was derived from
    org = synthetic.pssm_organisms(1, 100, 19)[0]
    rec = org.recognizers[0]
    rec.pssm = np.array([dict(col, a=col["a"] + 1.0) for col in rec.pssm])
    org.flatten()
    assert org.recognizers_flat[0] == rec.pssm[0]["a"]


def test_fitness_finalize_is_the_reference_expression_on_numpy_scalars():
    """fl_fitness_finalize recomputes (M_p - M_n) / (SE_p ** 2 + SE_n ** 2) ** (1/2) as OrganismObject.set_fitness does
    with numpy scalars (organism_object.py:954-956) -- bit for bit."""
    rng = np.random.default_rng(12)
    stats = np.empty((20000, 5))
    stats[:, 1] = rng.normal(-20, 15, 20000)
    stats[:, 3] = rng.normal(-25, 15, 20000)
    stats[:, 2] = np.maximum(1.0, rng.gamma(2.0, 2.0, 20000)) / np.sqrt(rng.integers(1, 20000, 20000))
    stats[:, 4] = np.maximum(1.0, rng.gamma(2.0, 2.0, 20000)) / np.sqrt(rng.integers(1, 20000, 20000))
    stats[:, 0] = 0.0
    want = np.array([(np.float64(r[1]) - np.float64(r[3])) / (np.float64(r[2]) ** 2 + np.float64(r[4]) ** 2) ** (1 / 2) for r in stats])
    got = builders.fitness_finalize(stats.copy())[:, 0]
    assert got.tobytes() == want.tobytes()
