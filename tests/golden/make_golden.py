#!/usr/bin/env python3
"""Generates tests/golden/golden_v1.json from the REFERENCE itself, run in the build container.

What runs: the reference's own Python objects (/root/reference/src/objects/*.py, models/null.py) on top of the
unmodified reference C extension compiled into oracle/_ref by `make -C oracle ref`.  Nothing from this repo's
product code is involved, so the fixture pins the oracle and the CUDA path to reference outputs.
The GPU box has no /root/reference: tests there read only the committed JSON.

    python tests/golden/make_golden.py          # needs /root/reference and oracle/_ref

Cases: n = 1..6 recognizers, PSSM and all four shape types, S in {sum L (P=1), sum L + 1, 60..100,
200, 300, 2000}, connectors with mu > S + 0.5 (adjust_scores rescale), tiny sigma, sigma = 0, mu = 0, repeated
sequences (exact ties), mixed case, non-ACGT bytes, the bundled organism.json on bundled FASTA sequences, and
the four fitness variants with gamma in {0, 0.2, 0.5}.
"""
import hashlib
import json
import os
import random
import sys
import tempfile
from pathlib import Path

import numpy as np

REPO = Path(__file__).resolve().parents[2]
REF = Path("/root/reference")
OUT = Path(__file__).resolve().parent / "golden_v1.json"

sys.path.insert(0, str(REPO / "oracle" / "_ref"))   # the compiled reference `_multiplacement`
sys.path.insert(0, str(REF / "src"))

# null.generate_range pickles into ./models/models -> run from a scratch cwd
_scratch = tempfile.mkdtemp()
os.makedirs(os.path.join(_scratch, "models"))
os.chdir(_scratch)

from models import null                                    # noqa: E402
from objects import shape_object                           # noqa: E402
from objects.connector_object import ConnectorObject      # noqa: E402
from objects.organism_object import OrganismObject        # noqa: E402
from objects.pssm_object import PssmObject                 # noqa: E402
from objects.shape_object import ShapeObject               # noqa: E402

CONF = json.load(open(REF / "src" / "config.json"))
METHODS = ["Welch", "Yuen", "Trim-left-M", "Trim-left-M-SD"]
GAMMAS = [0.0, 0.2, 0.5]
INLINE_LIMIT = 4000   # flat arrays longer than this are stored as a sha256 only


def read_fasta(path):
    """Records of a (possibly multi-line) FASTA file; sequences upper-cased like search_organisms.py:931-949."""
    seqs, cur = [], []
    for ln in open(path):
        ln = ln.strip()
        if ln.startswith(">"):
            if cur:
                seqs.append("".join(cur))
            cur = []
        elif ln:
            cur.append(ln.upper())
    if cur:
        seqs.append("".join(cur))
    return seqs


def sha(arr):
    return hashlib.sha256(np.ascontiguousarray(arr).tobytes()).hexdigest()


def pwm_column(rng):
    left, counts = 100, []
    for item in (3, 2, 1):
        c = rng.randint(1, left - item) if left > item else 1
        left -= c
        counts.append(c)
    counts.append(left)
    rng.shuffle(counts)
    return {"a": counts[0] / 100, "g": counts[1] / 100, "c": counts[2] / 100, "t": counts[3] / 100}


def build(nodes, M, org_id):
    org = OrganismObject(org_id, CONF["organism"])
    for node in nodes:
        if node["kind"] == "pssm":
            org.recognizers.append(PssmObject(np.array(node["pwm"]), CONF["pssm"]))
        elif node["kind"] == "shape":
            org.recognizers.append(ShapeObject(node["type"], node["length"], CONF["shape"], node["mu"], node["sigma"]))
        else:
            org.connectors.append(ConnectorObject(CONF["connector"], M, node["mu"], node["sigma"]))
    org.flatten()
    return org


def random_nodes(rng, n, shape_prob, M, len_lo=3, len_hi=12, mu=None, sigma=None):
    nodes = []
    for i in range(n):
        if rng.random() < shape_prob:
            kind = rng.choice(["mgw", "prot", "roll", "helt"])
            L = rng.randint(5, 9)
            edges = shape_object.null_models[kind][L]["bins"]
            nodes.append({"kind": "shape", "type": kind, "length": L, "mu": rng.uniform(edges[0], edges[-1]),
                          "sigma": rng.expovariate(12 / (edges[-1] - edges[0]))})
        else:
            nodes.append({"kind": "pssm", "pwm": [pwm_column(rng) for _ in range(rng.randint(len_lo, len_hi))]})
        if i < n - 1:
            nodes.append({"kind": "connector",
                          "mu": float(np.random.poisson(25)) if mu is None else mu(rng),
                          "sigma": rng.expovariate(12 / M) if sigma is None else sigma(rng)})
    return nodes


def seqs_random(rng, n, S, alphabet="acgt"):
    return ["".join(rng.choice(alphabet) for _ in range(S)) for _ in range(n)]


def record(name, nodes, M, sequences, fitness=True):
    org = build(nodes, M, name)
    flat = {}
    for key in ("recognizers_flat", "recognizer_lengths", "recognizer_models", "recognizer_bin_edges",
                "recognizer_bin_nums", "connectors_scores_flat"):
        arr = getattr(org, key)
        flat[key] = arr.tolist() if arr.size <= INLINE_LIMIT else {"sha256": sha(arr), "size": int(arr.size)}
    flat["recognizer_types"] = org.recognizer_types
    placements = []
    for s in sequences:
        p = org.get_placement(s)
        placements.append({"energy": p.energy, "recognizers_scores": p.recognizers_scores,
                           "connectors_scores": p.connectors_scores, "recognizers_positions": p.recognizers_positions,
                           "connectors_positions": p.connectors_positions})
    fits = []
    if fitness and len(sequences) >= 4:
        half = len(sequences) // 2
        pos, neg = sequences[:half], sequences[half:]
        for method in METHODS:
            for gamma in GAMMAS:
                m_p, se_p = org.get_M_and_SE_on_DNA_set(pos, method, gamma)
                m_n, se_n = org.get_M_and_SE_on_DNA_set(neg, method, gamma)
                org.set_fitness(pos, neg, method, gamma)
                fits.append({"method": method, "gamma": gamma, "n_pos": half, "fitness": float(org.fitness),
                             "M_p": float(m_p), "SE_p": float(se_p), "M_n": float(m_n), "SE_n": float(se_n)})
    return {"name": name, "max_seq_length": M, "nodes": nodes, "flat": flat, "sequences": sequences,
            "placements": placements, "fitness": fits}


def main():
    rng = random.Random(20261019)
    np.random.seed(20261019)
    models = {}
    null.generate_range(5, CONF["shape"]["MAX_LENGTH"] + 1, models, CONF["shape"]["NUM_BINS"])
    shape_object.null_models = models[CONF["shape"]["NUM_BINS"]]
    nm = {t: {str(L): {"frequencies": v["frequencies"].tolist(), "bins": v["bins"].tolist()} for L, v in d.items()}
          for t, d in shape_object.null_models.items()}
    cases = []

    # 1. random organisms, n = 1..6, PSSM + shapes, S 60..100
    for n in range(1, 7):
        for rep in range(4):
            M = rng.randint(60, 100)
            nodes = random_nodes(rng, n, 0.4, M, len_hi=8)
            cases.append(record(f"rand_n{n}_{rep}", nodes, M, seqs_random(rng, 8, M)))
    # 2. PSSM-only, the BASELINE config-2 shape (3 x len 12 on 300 bp) and a 200 bp variant
    for rep in range(3):
        nodes = random_nodes(rng, 3, 0.0, 300, len_lo=12, len_hi=12)
        cases.append(record(f"c2_shape_{rep}", nodes, 300, seqs_random(rng, 10, 300)))
    cases.append(record("pssm_200", random_nodes(rng, 4, 0.0, 200, len_hi=10), 200, seqs_random(rng, 8, 200)))
    # 3. every shape type alone and chained, every length 5..9
    for kind in ("mgw", "prot", "roll", "helt"):
        for L in (5, 7, 9):
            edges = shape_object.null_models[kind][L]["bins"]
            node = {"kind": "shape", "type": kind, "length": L, "mu": rng.uniform(edges[0], edges[-1]),
                    "sigma": rng.expovariate(12 / (edges[-1] - edges[0]))}
            cases.append(record(f"shape_{kind}_{L}", [node], 80, seqs_random(rng, 6, 80), fitness=False))
    cases.append(record("shape_chain", random_nodes(rng, 6, 1.0, 120), 120, seqs_random(rng, 8, 120)))
    # 4. P = 1 (sum L == S) and P = 2, where BIG_NEGATIVE enters the energy (connector.c:53-54)
    for n in (1, 2, 3, 5):
        nodes = random_nodes(rng, n, 0.3, 90, len_lo=5, len_hi=9)
        sum_len = sum(len(x["pwm"]) if x["kind"] == "pssm" else x.get("length", 0) for x in nodes if x["kind"] != "connector")
        cases.append(record(f"tight_n{n}", nodes, 90, seqs_random(rng, 4, sum_len) + seqs_random(rng, 4, sum_len + 1),
                            fitness=False))
    # 5. connector corner cases: rescale trigger (mu > S + 0.5), tiny sigma, sigma = 0, mu = 0
    for rep in range(6):
        M = rng.randint(70, 100)
        nodes = random_nodes(rng, rng.randint(2, 4), 0.25, M, len_hi=7,
                             mu=lambda r, M=M: r.uniform(M + 1, 3 * M), sigma=lambda r: r.uniform(0.3, 3.0))
        seqs = seqs_random(rng, 4, M) + seqs_random(rng, 4, M - rng.randint(1, 20))
        cases.append(record(f"rescale_{rep}", nodes, M, seqs))
    nodes = random_nodes(rng, 3, 0.0, 80, mu=lambda r: float(r.randint(5, 30)), sigma=lambda r: 1e-3)
    cases.append(record("sigma_tiny", nodes, 80, seqs_random(rng, 8, 80)))
    nodes = random_nodes(rng, 3, 0.0, 80, mu=lambda r: float(r.randint(5, 30)), sigma=lambda r: 0)
    cases.append(record("sigma_zero", nodes, 80, seqs_random(rng, 8, 80)))
    nodes = random_nodes(rng, 4, 0.2, 80, mu=lambda r: 0.0, sigma=lambda r: r.uniform(0.5, 4))
    cases.append(record("mu_zero", nodes, 80, seqs_random(rng, 8, 80)))
    # 6. exact ties: homopolymers and short repeats; mixed case; non-ACGT bytes
    nodes = random_nodes(rng, 3, 0.3, 90, len_hi=6)
    cases.append(record("ties", nodes, 90, ["a" * 90, "gtacaacatg" * 9, "ac" * 45, "t" * 60, "acgt" * 20, "ggc" * 30,
                                            "a" * 45 + "c" * 45, "tgca" * 22]))
    nodes = random_nodes(rng, 4, 0.5, 80, len_hi=6)
    mixed = ["".join(rng.choice("acgtACGT") for _ in range(80)) for _ in range(4)]
    unknown = ["".join(rng.choice("acgtnNxACGT-") for _ in range(80)) for _ in range(4)]
    cases.append(record("case_and_unknown", nodes, 80, mixed + unknown))
    # 7. ragged set: one organism, sequences of many lengths (tables sized for the longest)
    nodes = random_nodes(rng, 3, 0.3, 150, len_hi=8)
    ragged = [s for L in (150, 149, 120, 97, 64, 40, 33, 150) for s in seqs_random(rng, 1, L)]
    cases.append(record("ragged", nodes, 150, ragged))
    # 8. long promoter (BASELINE config 4 shape): 2 kb, wide gap windows
    nodes = random_nodes(rng, 3, 0.0, 2000, len_lo=12, len_hi=12, mu=lambda r: r.uniform(0, 1500),
                         sigma=lambda r: r.expovariate(12 / 2000))
    cases.append(record("long_2kb", nodes, 2000, seqs_random(rng, 4, 2000)))
    # 9. the bundled example organism on bundled FASTA sequences (BASELINE config 1 inputs)
    raw = json.load(open(REF / "src" / "organism.json"))[0]
    nodes = []
    for el in raw:
        if el["objectType"] == "pssm":
            nodes.append({"kind": "pssm", "pwm": el["pwm"]})
        elif el["objectType"] == "connector":
            nodes.append({"kind": "connector", "mu": el["mu"], "sigma": el["sigma"]})
        else:
            nodes.append({"kind": "shape", "type": el["recType"], "length": el["length"], "mu": el["mu"], "sigma": el["sigma"]})
    fasta = read_fasta(REF / "src" / "datasets" / "positive_dataset_7_100.fas")
    ctrl = read_fasta(REF / "src" / "datasets" / "positive_dataset_7_100pc5.fas")
    cases.append(record("bundled_organism", nodes, 200, fasta[:10] + ctrl[:10]))

    with open(OUT, "w") as fh:
        json.dump({"generator": "tests/golden/make_golden.py", "reference": "ErillLab/FLEMINGO @ /root/reference",
                   "python": sys.version.split()[0], "numpy": np.__version__, "null_models": nm, "cases": cases}, fh)
    n_pl = sum(len(c["placements"]) for c in cases)
    print(f"wrote {OUT}: {len(cases)} cases, {n_pl} placements, {OUT.stat().st_size/1e6:.2f} MB")


if __name__ == "__main__":
    main()
