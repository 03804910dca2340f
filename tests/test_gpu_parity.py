"""GPU parity tests: the CUDA path, called through the C ABI, against (a) the committed golden fixtures that
the reference itself produced and (b) the CPU oracle on seeded synthetic inputs.

Bar: energies, node scores, placement offsets AND fitness statistics BIT-EXACT (==).  The kernels do all arithmetic in
IEEE double in the reference's operand order; the fitness reduction follows numpy's pairwise summation order and the
final quotient is formed on the host with the same libm calls the reference's numpy scalars make, so no tolerance is
needed anywhere (north_star allows 1e-5).
"""
import numpy as np
import pytest

from oracle import oracle
from tests.helpers import assert_placement_equals_golden

pytestmark = pytest.mark.gpu



def same_doubles(got, want):
    """bit-identical float64 arrays (NaNs in the same places)"""
    return np.array_equal(np.asarray(got, dtype=np.float64), np.asarray(want, dtype=np.float64), equal_nan=True)



def test_golden_cases_batched_trace(engine, golden_cases):
    """Every golden organism on its own sequences through fl_place_batch with traceback."""
    n = 0
    for case, org in golden_cases:
        engine.upload_population([org])
        sset = engine.upload_sequences(case["sequences"], 0)
        res = engine.place(sset, trace=True)
        for s, ref in enumerate(case["placements"]):
            assert_placement_equals_golden(ref, res.energies[0, s], res.rec_scores[0, s], res.con_scores[0, s], res.gaps[0, s],
                                           org.recognizer_lengths, f"{case['name']}[{s}]")
            n += 1
        # the energy-only kernel variant must agree with the trace variant
        assert np.array_equal(engine.place(sset, trace=False).energies, res.energies), case["name"]
    assert n > 400


def test_golden_cases_through_get_placement(engine, golden_cases):
    """The reference-facing call: OrganismObject.get_placement -> multiplacement.calculate -> fl_calculate."""
    for case, org in golden_cases[::3]:
        for seq, ref in list(zip(case["sequences"], case["placements"]))[:3]:
            p = org.get_placement(seq)
            assert p.energy == ref["energy"], case["name"]
            assert p.recognizers_scores == ref["recognizers_scores"]
            assert p.connectors_scores == ref["connectors_scores"]
            assert p.recognizers_positions == ref["recognizers_positions"]
            assert p.connectors_positions == ref["connectors_positions"]
            assert p.recognizer_types == org.recognizer_types


def test_golden_population_in_one_batch(engine, golden_cases):
    """All golden organisms that share a table length as ONE population on one ragged sequence set."""
    by_M = {}
    for case, org in golden_cases:
        by_M.setdefault(case["max_seq_length"], []).append((case, org))
    checked = 0
    for M, group in by_M.items():
        if M > 300:
            continue
        orgs = [o for _, o in group]
        longest = max(o.sum_recognizer_lengths for o in orgs)
        seqs = [s for c, _ in group for s in c["sequences"] if longest <= len(s) <= M]
        if not seqs:
            continue
        engine.upload_population(orgs)
        res = engine.place(engine.upload_sequences(seqs, 0), trace=True)
        E, G, R, Cs = oracle.place_all(orgs, seqs, "restated")
        assert np.array_equal(res.energies, E)
        assert np.array_equal(res.gaps, G)
        assert np.array_equal(res.rec_scores, R)
        assert np.array_equal(res.con_scores, Cs)
        checked += E.size
    assert checked > 100


def test_golden_fitness_all_methods(engine, golden_cases):
    n = 0
    for case, org in golden_cases:
        if not case["fitness"]:
            continue
        half = case["fitness"][0]["n_pos"]
        pos, neg = case["sequences"][:half], case["sequences"][half:]
        engine.upload_population([org])
        sp, sn = engine.upload_sequences(pos, 0), engine.upload_sequences(neg, 1)
        engine.place(sp, fetch=False)
        engine.place(sn, fetch=False)
        for f in case["fitness"]:
            got = engine.fitness(sp, sn, f["method"], f["gamma"])[0]
            want = np.array([f["fitness"], f["M_p"], f["SE_p"], f["M_n"], f["SE_n"]])
            assert same_doubles(got, want), (case["name"], f, got, want)
            n += 1
    assert n > 300


def test_object_api_matches_golden_fitness(engine, golden_cases):
    """set_fitness / get_M_and_SE_on_DNA_set / get_binding_energies keep the reference's return values."""
    case, org = next((c, o) for c, o in golden_cases if c["name"] == "c2_shape_0")
    half = case["fitness"][0]["n_pos"]
    pos, neg = case["sequences"][:half], case["sequences"][half:]
    energies = org.get_binding_energies(case["sequences"])
    assert energies == [p["energy"] for p in case["placements"]]
    for f in case["fitness"]:
        org.set_fitness(pos, neg, f["method"], f["gamma"])
        assert same_doubles(org.fitness, f["fitness"])
        m, se = org.get_M_and_SE_on_DNA_set(pos, f["method"], f["gamma"])
        assert same_doubles(m, f["M_p"]) and same_doubles(se, f["SE_p"])
    with pytest.raises(ValueError):
        org.set_fitness(pos, neg, "NoSuchMethod", 0.2)


@pytest.mark.parametrize("kind,n_org,S,n_seq", [("pssm", 24, 300, 40), ("mixed", 48, 300, 24), ("mixed", 40, 97, 33)])
def test_synthetic_population_vs_oracle(engine, kind, n_org, S, n_seq):
    from flemingo_b200 import synthetic
    orgs = synthetic.pssm_organisms(n_org, S, 3) if kind == "pssm" else synthetic.mixed_organisms(n_org, S, 3)
    seqs = synthetic.random_sequences(n_seq, S, 1)
    engine.upload_population(orgs)
    res = engine.place(engine.upload_sequences(seqs, 0), trace=True)
    E, G, R, Cs = oracle.place_all(orgs, seqs, "restated")
    assert np.array_equal(res.energies, E)
    assert np.array_equal(res.gaps, G)
    assert np.array_equal(res.rec_scores, R)
    assert np.array_equal(res.con_scores, Cs)


def test_long_promoter_vs_oracle(engine):
    """BASELINE config-4 shape: 2 kb sequences, wide gap windows (P ~ 1965)."""
    from flemingo_b200 import synthetic
    orgs = synthetic.pssm_organisms(3, 2000, 5, mu_range=(0, 1500))
    seqs = synthetic.random_sequences(5, 2000, 6)
    engine.upload_population(orgs)
    res = engine.place(engine.upload_sequences(seqs, 0), trace=True)
    E, G, R, Cs = oracle.place_all(orgs, seqs, "restated")
    assert np.array_equal(res.energies, E) and np.array_equal(res.gaps, G)
    assert np.array_equal(res.rec_scores, R) and np.array_equal(res.con_scores, Cs)


def test_full_size_config2_properties(engine):
    """BASELINE config 2 at full size (256 organisms x 2000 sequences of 300 bp): size-independent properties.

    (1) a random sample of pairs equals the oracle bit for bit; (2) energy-only and trace kernels agree;
    (3) energy == sum of the reported node scores in the reference's association order; (4) positions are
    a valid chain inside the sequence; (5) duplicating a sequence duplicates its column (no cross-talk).
    """
    from flemingo_b200 import synthetic
    orgs, pos, neg = synthetic.make_workload("c2")
    seqs = pos + neg
    seqs[1234] = seqs[7]
    engine.upload_population(orgs)
    sset = engine.upload_sequences(seqs, 0)
    res = engine.place(sset, trace=True)
    assert np.array_equal(engine.place(sset, trace=False).energies, res.energies)
    assert np.array_equal(res.energies[:, 1234], res.energies[:, 7])
    rng = np.random.default_rng(0)
    for o, s in zip(rng.integers(0, len(orgs), 300), rng.integers(0, len(seqs), 300)):
        e, rs, cs, gaps, _ = oracle.place(orgs[o], seqs[s])
        assert res.energies[o, s] == e and np.array_equal(res.gaps[o, s, :3], gaps)
        assert np.array_equal(res.rec_scores[o, s, :3], rs) and np.array_equal(res.con_scores[o, s, :2], cs)
    # chain validity and energy recomposition for every pair
    L = 12
    g = res.gaps.astype(np.int64)
    end = g[:, :, 0] + g[:, :, 1] + g[:, :, 2] + 3 * L
    assert (g >= 0).all() and (end <= 300).all()
    r, c = res.rec_scores, res.con_scores
    recomposed = ((r[:, :, 0] + c[:, :, 0]) + r[:, :, 1] + c[:, :, 1]) + r[:, :, 2]
    assert np.array_equal(recomposed, res.energies)


def test_error_behaviour(engine, golden_cases):
    from flemingo_b200 import _cabi, synthetic
    case, org = next((c, o) for c, o in golden_cases if c["name"] == "c2_shape_0")
    # organism longer than the sequence: ValueError, as organism_object.py:1279
    with pytest.raises(ValueError):
        org.get_placement("acgt" * 5)
    engine.upload_population([org])
    with pytest.raises(ValueError):
        engine.place(engine.upload_sequences(["acgt" * 5], 0))
    # sequence longer than the precomputed gap tables: legacy branch, rejected
    with pytest.raises(_cabi.FlemingoError) as ei:
        engine.place(engine.upload_sequences(["acgt" * 100], 0))
    assert ei.value.code == _cabi.FL_E_UNSUPPORTED
    # empty set and empty population
    empty = engine.upload_sequences([], 0)
    assert engine.place(empty).energies.shape == (1, 0)
    assert org.get_binding_energies([]) == []
    # NaN null bin: the reference exits the process; we raise
    shp = synthetic.random_shape(np.random.default_rng(1), 5, "mgw")
    shp.null_model = np.full_like(np.asarray(shp.null_model, dtype=np.float64), np.nan)
    bad = synthetic.assemble("nan", [shp], [])
    engine.upload_population([bad])
    with pytest.raises(_cabi.FlemingoError) as ei:
        engine.place(engine.upload_sequences(["acgtacgtacgt"], 0))
    assert ei.value.code == _cabi.FL_E_NAN_NULL
    # the context stays usable afterwards
    engine.upload_population([org])
    ok = engine.place(engine.upload_sequences(case["sequences"][:2], 0))
    assert ok.energies[0, 0] == case["placements"][0]["energy"]


def test_single_recognizer_and_tight_fits(engine):
    """n = 1 (no connector) and P = 1 / P = 2, where -1e10 enters the energy (connector.c:53-54)."""
    from flemingo_b200 import synthetic
    rng = np.random.default_rng(9)
    one = synthetic.assemble("one", [synthetic.random_pssm(rng, 7)], [])
    three = synthetic.pssm_organisms(1, 40, 10, n_rec=3, rec_len=5)[0]
    for org, lens in ((one, (7, 8, 30)), (three, (15, 16, 17, 40))):
        seqs = [s for L in lens for s in synthetic.random_sequences(3, L, L)]
        engine.upload_population([org])
        res = engine.place(engine.upload_sequences(seqs, 0), trace=True)
        E, G, R, Cs = oracle.place_all([org], seqs, "restated")
        assert np.array_equal(res.energies, E) and np.array_equal(res.gaps, G)
        assert np.array_equal(res.rec_scores, R) and np.array_equal(res.con_scores, Cs)


# ---------------------------------------------------------------------------------------------------------------
# the integer filter pass: every tile configuration, its hand-back paths, and its fixed-point range
# ---------------------------------------------------------------------------------------------------------------
def _assert_equals_oracle(engine, orgs, seqs):
    engine.upload_population(orgs)
    res = engine.place(engine.upload_sequences(seqs, 0), trace=True)
    E, G, R, Cs = oracle.place_all(orgs, seqs, "restated")
    assert np.array_equal(res.energies, E)
    assert np.array_equal(res.gaps, G)
    assert np.array_equal(res.rec_scores, R)
    assert np.array_equal(res.con_scores, Cs)
    return res


@pytest.mark.parametrize("S", [24, 40, 66, 97, 130, 175, 210, 260, 300, 340, 450, 700, 1100])
def test_every_tile_configuration_vs_oracle(engine, S):
    """Sequence lengths chosen so that P = S - sum(L) + 1 walks through the (J, TEAM) table of the fast path and
    the tile sizes of the exact kernel, including multi-round folds."""
    from flemingo_b200 import synthetic
    n_org = 10 if S <= 450 else 4
    orgs = synthetic.mixed_organisms(n_org, S, 100 + S, max_rec=4) + synthetic.pssm_organisms(4, S, 200 + S, n_rec=2, rec_len=6)
    orgs = [o for o in orgs if o.sum_recognizer_lengths <= S - 3]
    seqs = synthetic.random_sequences(9 if S <= 450 else 3, S, S) + synthetic.random_sequences(2, S - 3, S + 1)
    _assert_equals_oracle(engine, orgs, seqs)


def test_massive_ties_are_handed_to_the_exact_kernel(engine):
    """Homopolymers and short repeats make every placement tie: the candidate sets overflow, the pairs go to the
    list-mode exact kernel, and the reference's first-k / first-j tie rule still decides the offsets."""
    from flemingo_b200 import synthetic
    orgs = synthetic.pssm_organisms(6, 200, 31, n_rec=3, rec_len=7)
    seqs = ["a" * 200, "c" * 200, "ac" * 100, "acg" * 66 + "ac", "t" * 100 + "g" * 100] + synthetic.random_sequences(3, 200, 32)
    engine.path_stats()
    _assert_equals_oracle(engine, orgs, seqs)
    fast, handed_back = engine.path_stats()
    if not engine.exact_only:
        assert fast == len(orgs) * len(seqs)
        assert handed_back >= len(orgs) * 2               # at least the homopolymers
        assert handed_back < fast                         # the random sequences stay on the integer path


def test_unknown_bytes_take_the_exact_path(engine):
    from flemingo_b200 import synthetic
    orgs = synthetic.mixed_organisms(8, 120, 41, max_rec=4)
    rng = np.random.default_rng(5)
    seqs = ["".join(rng.choice(list("acgtnNx-ACGT"), 120)) for _ in range(5)] + synthetic.random_sequences(4, 120, 42)
    engine.path_stats()
    _assert_equals_oracle(engine, orgs, seqs)
    fast, handed_back = engine.path_stats()
    if not engine.exact_only and fast:
        assert handed_back >= 5 * sum(1 for o in orgs if len(o.recognizer_types) >= 2)


def test_extreme_pssm_values_shrink_the_fixed_point_scale_but_stay_exact(engine):
    """PWM columns with zero probabilities give PSSM entries of log2(1e-10) = -33.22: the value range grows, the
    fixed-point scale shrinks, thresholds widen -- results must not change."""
    from flemingo_b200 import objects, synthetic
    rng = np.random.default_rng(77)
    orgs = []
    for o in range(6):
        recs = []
        for _ in range(3):
            cols = []
            for _ in range(int(rng.integers(4, 11))):
                p = rng.dirichlet([0.3] * 4)
                p[rng.integers(0, 4)] = 0.0
                p = p / p.sum()
                cols.append({"a": float(p[0]), "g": float(p[1]), "c": float(p[2]), "t": float(p[3])})
            recs.append(objects.PssmObject(np.array(cols), synthetic.CONF_PSSM))
        cons = [synthetic.random_connector(rng, 150) for _ in range(2)]
        orgs.append(synthetic.assemble(f"x{o}", recs, cons))
    _assert_equals_oracle(engine, orgs, synthetic.random_sequences(12, 150, 78))


def test_full_size_config3_sample_vs_oracle(engine):
    """BASELINE config 3 shapes (mixed PSSM + shape organisms, up to 6 recognizers, 300 bp): a slice big enough to
    fill several CTAs per bucket, every pair against the oracle."""
    from flemingo_b200 import synthetic
    orgs = synthetic.mixed_organisms(120, 300, 3)
    seqs = synthetic.random_sequences(40, 300, 1)
    _assert_equals_oracle(engine, orgs, seqs)


def test_adversarial_population_fast_path_equals_exact_path_and_oracle(engine):
    """Adversarial organisms x low-complexity sequences (tests/helpers.adversarial_workload): the two GPU paths agree on
    every pair and a sample equals the CPU oracle."""
    from tests.helpers import adversarial_workload
    orgs, seqs = adversarial_workload(11, 150, n_seq=60)
    engine.upload_population(orgs)
    sset = engine.upload_sequences(seqs, 0)
    mine = engine.place(sset, trace=True)
    was = engine.exact_only
    engine.exact_only = not was
    other = engine.place(sset, trace=True)
    engine.exact_only = was
    for k in ("energies", "gaps", "rec_scores", "con_scores"):
        assert np.array_equal(getattr(mine, k), getattr(other, k)), k
    rng = np.random.default_rng(3)
    for o, s in zip(rng.integers(0, len(orgs), 150), rng.integers(0, len(seqs), 150)):
        e, rs, cs, gaps, _ = oracle.place(orgs[o], seqs[s])
        n = len(orgs[o].recognizer_types)
        assert mine.energies[o, s] == e and np.array_equal(mine.gaps[o, s, :n], gaps)
        assert np.array_equal(mine.rec_scores[o, s, :n], rs) and np.array_equal(mine.con_scores[o, s, :n - 1], cs)


# ---------------------------------------------------------------------------------------------------------------
# the BASELINE configurations at FULL size: sampled pairs against the oracle, size-independent properties, and no
# hand-backs on the workloads the bench reports
# ---------------------------------------------------------------------------------------------------------------
import functools


@functools.lru_cache(maxsize=None)
def _workload(name):
    from flemingo_b200 import synthetic
    return synthetic.make_workload(name)


def _check_sampled_pairs(engine, orgs, seqs, energies, n_pairs, seed):
    rng = np.random.default_rng(seed)
    for o, s in zip(rng.integers(0, len(orgs), n_pairs), rng.integers(0, len(seqs), n_pairs)):
        e = oracle.place(orgs[o], seqs[s])[0]
        assert energies[o, s] == e, (o, s, energies[o, s], e)


def _check_traced_slice(engine, orgs, seqs, pick_orgs, n_pairs, seed):
    """Offsets and node scores of a slice of the population on the whole set; sampled pairs against the oracle."""
    sub = [orgs[i] for i in pick_orgs]
    engine.upload_population(sub)
    res = engine.place(engine.upload_sequences(seqs, 0), trace=True)
    rng = np.random.default_rng(seed)
    for o, s in zip(rng.integers(0, len(sub), n_pairs), rng.integers(0, len(seqs), n_pairs)):
        e, rs, cs, gaps, _ = oracle.place(sub[o], seqs[s])
        n = len(sub[o].recognizer_types)
        assert res.energies[o, s] == e and np.array_equal(res.gaps[o, s, :n], gaps)
        assert np.array_equal(res.rec_scores[o, s, :n], rs) and np.array_equal(res.con_scores[o, s, :n - 1], cs)
    # chain validity for every pair of the slice
    lens = np.zeros((len(sub), res.gaps.shape[2]), dtype=np.int64)
    for i, o in enumerate(sub):
        lens[i, :len(o.recognizer_lengths)] = o.recognizer_lengths
    end = res.gaps.astype(np.int64).sum(axis=2) + lens.sum(axis=1)[:, None]
    assert (res.gaps >= 0).all() and (end <= len(seqs[0])).all()
    return res


def test_full_size_config3_vs_oracle(engine):
    """BASELINE config 3 at full size: 2000 mixed organisms x 10 000 sequences of 300 bp (20 M placements)."""
    orgs, pos, neg = _workload("c3")
    seqs = pos + neg
    engine.upload_population(orgs)
    sset = engine.upload_sequences(seqs, 0)
    engine.path_stats()
    res = engine.place(sset)
    fast, handed_back = engine.path_stats()
    assert res.energies.shape == (2000, 10000)
    if not engine.exact_only:
        assert fast > 0.8 * res.energies.size          # organisms with >= 2 recognizers take the integer path ...
        assert handed_back < 1e-3 * fast               # ... and hands back a few pairs in 10 000 (exact ties of shape bins)
    _check_sampled_pairs(engine, orgs, seqs, res.energies, 320, 1)
    # sharded == unsharded, bit for bit: the same organisms in another batch composition give the same doubles
    pick = np.arange(0, 2000, 37)
    traced = _check_traced_slice(engine, orgs, seqs, pick, 120, 2)
    assert np.array_equal(traced.energies, res.energies[pick])
    # fitness statistics of the full batch against numpy on the device energies
    sp, sn = engine.upload_sequences(pos, 0), engine.upload_sequences(neg, 1)
    engine.upload_population(orgs)
    engine.place(sp, fetch=False)
    engine.place(sn, fetch=False)
    stats = engine.fitness(sp, sn, "Welch", 0.2)
    for o in range(0, 2000, 97):
        want = oracle.fitness_stats(res.energies[o, :5000], res.energies[o, 5000:], "Welch", 0.2)
        assert same_doubles(stats[o], want)


def test_full_size_config4_vs_oracle(engine):
    """BASELINE config 4 at full size: 512 organisms x 5000 sequences of 2 kb, wide gap windows (P ~ 1965)."""
    orgs, pos, neg = _workload("c4")
    seqs = pos + neg
    engine.upload_population(orgs)
    sset = engine.upload_sequences(seqs, 0)
    engine.path_stats()
    res = engine.place(sset)
    fast, handed_back = engine.path_stats()
    assert res.energies.shape == (512, 5000)
    if not engine.exact_only:
        assert fast == res.energies.size and handed_back == 0
    _check_sampled_pairs(engine, orgs, seqs, res.energies, 300, 3)
    traced = _check_traced_slice(engine, orgs, seqs[:500], np.arange(0, 512, 16), 40, 4)
    assert np.array_equal(traced.energies, res.energies[::16, :500])


def test_full_size_config5_generation_vs_oracle(engine):
    """BASELINE config 5, one generation's evaluation at full size: 4096 mixed organisms x 40 000 sequences."""
    if engine.exact_only or "FLEMINGO_FAST_ROWS" in __import__("os").environ:
        pytest.skip("full config 5 is checked once, in the default kernel mode")
    orgs, pos, neg = _workload("c5")
    engine.upload_population(orgs)
    sp, sn = engine.upload_sequences(pos, 0), engine.upload_sequences(neg, 1)
    engine.path_stats()
    ep = engine.place(sp).energies
    engine.place(sn, fetch=False)
    fast, handed_back = engine.path_stats()
    assert ep.shape == (4096, 20000) and handed_back < 1e-3 * fast and fast > 0.8 * 2 * ep.size
    _check_sampled_pairs(engine, orgs, pos, ep, 300, 5)
    stats = engine.fitness(sp, sn, "Welch", 0.2)
    en = engine.place(sn).energies
    for o in range(0, 4096, 211):
        assert same_doubles(stats[o], oracle.fitness_stats(ep[o], en[o], "Welch", 0.2))


def test_minus_infinity_pssm_entries(engine):
    """A -inf PSSM entry (not producible by recalculate_pssm, but legal input to calculate): where every chain through a
    cell is -inf the reference's strict '<' never fires and its calloc'ed zeros stay in the outputs (organism.c:365)."""
    from flemingo_b200 import synthetic
    orgs = synthetic.pssm_organisms(3, 60, 21, n_rec=3, rec_len=4)
    for org in orgs:
        org.recognizers_flat = org.recognizers_flat.copy()
        org.recognizers_flat[4 * 4 + 0] = -np.inf      # recognizer 1, column 0, base A
        org.recognizers_flat[4 * 4 + 1] = -np.inf      # ... and G
    seqs = ["a" * 60, "ag" * 30, "aaaagggg" * 7 + "aaaa"] + synthetic.random_sequences(4, 60, 22)
    res = _assert_equals_oracle(engine, orgs, seqs)
    if oracle.reference_module() is not None:      # pin the restatement on this unusual input to the reference itself
        for o, org in enumerate(orgs):
            for s, seq in enumerate(seqs):
                e, rs, cs, gaps, _ = oracle.place(org, seq, "reference")
                assert (res.energies[o, s] == e or (np.isnan(e) and np.isnan(res.energies[o, s])))
                assert np.array_equal(res.gaps[o, s, :3], gaps) and np.array_equal(res.rec_scores[o, s, :3], rs, equal_nan=True)
                assert np.array_equal(res.con_scores[o, s, :2], cs)


def test_single_recognizer_on_a_long_sequence(engine):
    """One recognizer, no connector: DEN_EXPANSION is never read, so sequences beyond 10 kb are placed."""
    from flemingo_b200 import synthetic
    rng = np.random.default_rng(23)
    one = synthetic.assemble("one", [synthetic.random_pssm(rng, 9)], [])
    seqs = synthetic.random_sequences(2, 10500, 24)
    _assert_equals_oracle(engine, [one], seqs)


def test_strongly_negative_scores_like_mle_fitted_shapes(engine):
    """Shape recognizers as the MLE drive produces them (organism_factory.py:1172-1279): sigma = 0 or ~1e-15, so every
    bin but one scores ln(1e-100) = -230 and whole rows lie far below any gap energy.  (Round 1 started the DPX
    accumulators at the sentinel, which clamped such rows; found by running the reference's own driver.)"""
    from flemingo_b200 import objects, synthetic
    synthetic.ensure_null_models()
    rng = np.random.default_rng(31)
    orgs = []
    for o in range(10):
        n = int(rng.integers(2, 5))
        recs = []
        for _ in range(n):
            kind = synthetic.SHAPES[int(rng.integers(0, 4))]
            L = int(rng.integers(5, 8))
            edges = objects.null_models[kind][L]["bins"]
            sigma = float(rng.choice([0.0, 3.57e-15, 1e-6, 0.05]))
            recs.append(objects.ShapeObject(kind, L, synthetic.CONF_SHAPE, float(rng.uniform(edges[0], edges[-1])), sigma))
        if o % 3 == 0:
            recs[0] = synthetic.random_pssm(rng, 5)
        cons = [synthetic.random_connector(rng, 200, float(rng.uniform(5, 40)), float(rng.choice([0.0, 0.6, 5.0]))) for _ in range(n - 1)]
        orgs.append(synthetic.assemble(f"mle{o}", recs, cons))
    seqs = synthetic.random_sequences(40, 200, 32) + synthetic.random_sequences(8, 150, 33)
    _assert_equals_oracle(engine, orgs, seqs)


def test_forced_int16_pass_rejects_what_does_not_fit(engine):
    """FLEMINGO_FAST_BITS=16 routes EVERY organism to the packed 16-bit pass; the device-side plan must turn down the
    ones whose value range does not fit (their pairs go to the exact kernel) and stay exact on the rest."""
    import os
    from flemingo_b200 import synthetic
    from tests.helpers import adversarial_workload
    if engine.exact_only or os.environ.get("FLEMINGO_FAST_BITS"):
        pytest.skip("needs the automatic filter pass")
    os.environ["FLEMINGO_FAST_BITS"] = "16"
    try:
        orgs, seqs = adversarial_workload(13, 60, n_seq=40)
        _assert_equals_oracle(engine, orgs, seqs)
        mixed = synthetic.mixed_organisms(60, 300, 5)
        _assert_equals_oracle(engine, mixed, synthetic.random_sequences(24, 300, 6))
    finally:
        os.environ.pop("FLEMINGO_FAST_BITS", None)
