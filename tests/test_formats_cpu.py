"""CPU: FASTA ingest, organism JSON round trip and the Gini statistic."""
import json
import random
from pathlib import Path

import numpy as np
import pytest

from flemingo_b200 import formats

REF = Path("/root/reference")


def naive_gini(values):
    # src/search_organisms.py:1095-1142, the O(N^2) definition
    n = len(values)
    num = sum(abs(i - j) for i in values for j in values)
    tot = sum(abs(x) for x in values)
    mu = (n - 1) * tot / n ** 2
    return 0 if mu == 0 else num / (2 * n ** 2 * mu)


def test_gini_matches_quadratic_definition():
    rng = random.Random(4)
    for n in (1, 2, 3, 10, 57, 200):
        vals = [rng.uniform(-50, 120) for _ in range(n)]
        assert np.isclose(formats.gini_RSV(vals), naive_gini(vals), rtol=1e-11, atol=1e-14)
    assert formats.gini_RSV([0.0, 0.0, 0.0]) == 0 and formats.gini_RSV([7.0]) == 0


def test_fasta_reader(tmp_path):
    p = tmp_path / "x.fas"
    p.write_text(">one desc\nACGT\nacgt\n\n>two\nTTTT\n>three\n")
    assert formats.read_fasta_file(str(p)) == ["ACGTacgt", "TTTT", ""]


@pytest.mark.skipif(not (REF / "src" / "datasets").exists(), reason="needs the bundled FASTA files")
def test_fasta_reader_on_bundled_datasets(golden):
    pos = formats.read_fasta_file(str(REF / "src/datasets/positive_dataset_7_100.fas"))
    ctrl = formats.read_fasta_file(str(REF / "src/datasets/positive_dataset_7_100pc5.fas"))
    assert len(pos) == 100 and len(ctrl) == 5000 and {len(s) for s in pos + ctrl} == {200}
    case = next(c for c in golden["cases"] if c["name"] == "bundled_organism")
    assert case["sequences"] == pos[:10] + ctrl[:10]


def test_organism_json_round_trip(tmp_path, golden_cases):
    """export -> import reproduces the flat arrays; the bundled organism.json layout imports to the golden arrays."""
    orgs = [o for c, o in golden_cases if c["max_seq_length"] == 90][:6]
    path = tmp_path / "orgs.json"
    formats.export_organisms(orgs, str(path))
    back = formats.import_organisms(str(path), 90)
    assert len(back) == len(orgs)
    for a, b in zip(orgs, back):
        assert a.recognizer_types == b.recognizer_types
        for key in ("recognizers_flat", "recognizer_lengths", "recognizer_models", "recognizer_bin_edges", "connectors_scores_flat"):
            assert getattr(a, key).tobytes() == getattr(b, key).tobytes(), key
    case, org = next((c, o) for c, o in golden_cases if c["name"] == "bundled_organism")
    nodes = []
    for nd in case["nodes"]:
        if nd["kind"] == "pssm":
            nodes.append({"objectType": "pssm", "pwm": nd["pwm"]})
        elif nd["kind"] == "shape":
            nodes.append({"objectType": "shape", "recType": nd["type"], "length": nd["length"], "mu": nd["mu"], "sigma": nd["sigma"]})
        else:
            nodes.append({"objectType": "connector", "mu": nd["mu"], "sigma": nd["sigma"]})
    path2 = tmp_path / "bundled.json"
    path2.write_text(json.dumps([nodes]))
    imported = formats.import_organisms(str(path2), case["max_seq_length"])[0]
    assert imported.connectors_scores_flat.tobytes() == org.connectors_scores_flat.tobytes()
    assert imported.recognizers_flat.tobytes() == org.recognizers_flat.tobytes()
