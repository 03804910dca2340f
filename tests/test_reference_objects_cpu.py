"""CPU, build container only: the packer accepts the REFERENCE's own OrganismObject instances (duck typing) and
produces the same device arrays as for our mirror objects -- which is what makes the GPU path a drop-in.
Skipped where /root/reference does not exist (the GPU box)."""
import json
import os
import sys
from pathlib import Path

import numpy as np
import pytest

REF = Path("/root/reference")
REPO = Path(__file__).resolve().parents[1]

pytestmark = pytest.mark.skipif(not (REF / "src" / "objects").exists() or not list((REPO / "oracle" / "_ref").glob("_multiplacement*.so")),
                                reason="needs /root/reference and oracle/_ref")


def test_packer_on_reference_objects_matches_mirror(golden, golden_cases):
    sys.path.insert(0, str(REPO / "oracle" / "_ref"))
    sys.path.insert(0, str(REF / "src"))
    try:
        from objects import shape_object
        from objects.connector_object import ConnectorObject
        from objects.organism_object import OrganismObject
        from objects.pssm_object import PssmObject
        from objects.shape_object import ShapeObject
    finally:
        sys.path.remove(str(REF / "src"))
        sys.path.remove(str(REPO / "oracle" / "_ref"))
    conf = json.load(open(REF / "src" / "config.json"))
    shape_object.null_models = {t: {int(L): {"frequencies": np.array(v["frequencies"]), "bins": np.array(v["bins"])}
                                    for L, v in d.items()} for t, d in golden["null_models"].items()}
    from flemingo_b200.population import pack_population
    ref_orgs, mirror_orgs = [], []
    for case, org in golden_cases:
        if case["max_seq_length"] > 300:
            continue
        ro = OrganismObject(case["name"], conf["organism"])
        for node in case["nodes"]:
            if node["kind"] == "pssm":
                ro.recognizers.append(PssmObject(np.array(node["pwm"]), conf["pssm"]))
            elif node["kind"] == "shape":
                ro.recognizers.append(ShapeObject(node["type"], node["length"], conf["shape"], node["mu"], node["sigma"]))
            else:
                ro.connectors.append(ConnectorObject(conf["connector"], case["max_seq_length"], node["mu"], node["sigma"]))
        ro.flatten()
        ref_orgs.append(ro)
        mirror_orgs.append(org)
    a, b = pack_population(ref_orgs), pack_population(mirror_orgs)
    for name in ("rec_begin", "rec_type", "rec_len", "rec_nb", "rec_data", "pool_begin", "rec_pool", "con_M", "con_pool",
                 "con_base", "sum_len", "pseudo_count", "rec_class", "cls_kind", "cls_len", "cls_nb", "cls_edges", "edges_pool"):
        x, y = getattr(a, name), getattr(b, name)
        assert x.dtype == y.dtype and x.shape == y.shape, name
        assert x.tobytes() == y.tobytes(), name
    # the per-class connector rescale sees the same triggers
    lengths = sorted({len(s) for c, _ in golden_cases for s in c["sequences"] if len(s) <= 300})
    longest = int(max(a.sum_len))
    lengths = [L for L in lengths if L >= longest]
    assert np.array_equal(a.con_begin_for_lengths(lengths), b.con_begin_for_lengths(lengths))
    assert a.con_pool.tobytes() == b.con_pool.tobytes()


def test_closed_form_rescale_equals_reference_adjust_scores():
    """population.rescale_connector_block (closed form) writes the same doubles as the reference's
    ConnectorObject.adjust_scores loop (connector_object.py:454-487) wherever get_placement can call it."""
    sys.path.insert(0, str(REF / "src"))
    try:
        from objects.connector_object import ConnectorObject
    finally:
        sys.path.remove(str(REF / "src"))
    conf = json.load(open(REF / "src" / "config.json"))["connector"]
    from flemingo_b200.population import rescale_connector_block
    rng = np.random.default_rng(11)
    checked = 0
    for _ in range(200):
        M = int(rng.integers(20, 320))
        S = int(rng.integers(10, M + 1))
        sum_len = int(rng.integers(1, S + 1))
        mu = float(S + 0.5 + rng.choice([1e-9, 0.5, 3.0, 40.0, 1e4]) * rng.uniform(1, 2))
        sigma = float(rng.choice([1e-300, 1e-120, 1e-100, 1e-5, 0.3, 2.0, 50.0]))
        con = ConnectorObject(conf, M, mu, sigma)
        want = np.full(2 * M, -7.0)
        got = want.copy()
        try:
            con.adjust_scores(want, S, sum_len)
        except OverflowError:      # a denormal sigma overflows the reference's float power: same exception here
            with pytest.raises(OverflowError):
                rescale_connector_block(got, mu, sigma, con.pseudo_count, M, S, sum_len)
            continue
        rescale_connector_block(got, mu, sigma, con.pseudo_count, M, S, sum_len)
        assert got.tobytes() == want.tobytes(), (M, S, sum_len, mu, sigma)
        checked += 1
    assert checked > 150
