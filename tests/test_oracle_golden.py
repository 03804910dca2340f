"""CPU: pins the oracle restatement (oracle/placement_oracle.c) and the host-side producers to the reference.

Golden values come from the reference's own objects + C extension (tests/golden/make_golden.py).  Comparison
is bit-exact (==) on energies, node scores and positions.
"""
import numpy as np
import pytest

from oracle import oracle
from tests.helpers import assert_placement_equals_golden, sha


def test_host_producers_match_reference_flat_arrays(golden_cases):
    """PssmObject.recalculate_pssm, ShapeObject.set_alt_model, ConnectorObject tables, OrganismObject.flatten."""
    for case, org in golden_cases:
        flat = case["flat"]
        assert org.recognizer_types == flat["recognizer_types"], case["name"]
        for key in ("recognizers_flat", "recognizer_lengths", "recognizer_models", "recognizer_bin_edges",
                    "recognizer_bin_nums", "connectors_scores_flat"):
            mine = getattr(org, key)
            ref = flat[key]
            if isinstance(ref, dict):
                assert mine.size == ref["size"] and sha(mine) == ref["sha256"], (case["name"], key)
            else:
                ref = np.array(ref, dtype=mine.dtype)
                assert mine.shape == ref.shape, (case["name"], key)
                assert mine.tobytes() == ref.tobytes(), (case["name"], key)


def test_restated_oracle_reproduces_golden_placements(golden_cases):
    n = 0
    for case, org in golden_cases:
        for seq, ref in zip(case["sequences"], case["placements"]):
            energy, rs, cs, gaps, status = oracle.place(org, seq, "restated")
            assert status == 0
            assert_placement_equals_golden(ref, energy, rs, cs, gaps, org.recognizer_lengths, f"{case['name']}")
            n += 1
    assert n > 400


def test_rescale_cases_really_trigger_adjust_scores(golden_cases):
    """The 'rescale_*' cases must exercise connector_object.py:454-487 (otherwise the fixture is vacuous)."""
    from flemingo_b200.population import pack_population
    hit = 0
    for case, org in golden_cases:
        if not case["name"].startswith("rescale_"):
            continue
        packed = pack_population([org])
        lengths = sorted({len(s) for s in case["sequences"]})
        packed.con_begin_for_lengths(lengths)
        hit += len(packed.variants)
    assert hit >= 6


def test_reference_extension_agrees_with_restatement_on_random_cases():
    """oracle/_ref (unmodified reference C) vs the restatement on fresh random organisms, incl. shapes."""
    if oracle.reference_module() is None:
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    from flemingo_b200 import synthetic
    orgs = synthetic.mixed_organisms(60, 120, 11) + synthetic.pssm_organisms(10, 120, 12, n_rec=4, rec_len=9)
    seqs = synthetic.random_sequences(12, 120, 13) + synthetic.random_sequences(6, 77, 14)
    a = oracle.place_all(orgs, seqs, "restated")
    b = oracle.place_all(orgs, seqs, "reference")
    for x, y in zip(a, b):
        assert np.array_equal(x, y)


def test_fitness_restatement_matches_golden(golden_cases):
    n = 0
    for case, org in golden_cases:
        if not case["fitness"]:
            continue
        energies = [p["energy"] for p in case["placements"]]
        for f in case["fitness"]:
            half = f["n_pos"]
            got = oracle.fitness_stats(energies[:half], energies[half:], f["method"], f["gamma"])
            want = np.array([f["fitness"], f["M_p"], f["SE_p"], f["M_n"], f["SE_n"]])
            assert np.allclose(got, want, rtol=1e-12, atol=0, equal_nan=True), (case["name"], f)
            n += 1
    assert n > 300


def test_null_models_match_reference(golden_null_models):
    """Our vectorised null-model builder vs models/null.py (run under this Python).  Frequencies are equal;
    edges may differ in the last ulp because Python >= 3.12 `sum()` is compensated (null.py:44) while the
    reference's own environment (ENV.yml, Python 3.9) adds left to right, as we do."""
    from flemingo_b200.null_models import build_null_models
    mine = build_null_models(25, 5, 9)
    for t, d in golden_null_models.items():
        for L, ref in d.items():
            assert np.array_equal(mine[t][L]["frequencies"], ref["frequencies"], equal_nan=True), (t, L)
            assert np.allclose(mine[t][L]["bins"], ref["bins"], rtol=4e-16, atol=0), (t, L)
