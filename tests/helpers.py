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
