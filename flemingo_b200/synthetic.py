"""Synthetic workloads of BASELINE.json: random-nucleotide datasets and organisms drawn with
the reference factory's distributions (organism_factory.py:312-350, connector_object.py:150,169-173,
shape_object.py:116,134-138), from seeded numpy generators so GPU, oracle and reference see the same inputs.
"""
from __future__ import annotations

import numpy as np

from . import objects
from .null_models import build_null_models

CONF_ORG = {"PRECOMPUTE": True, "MIN_NUM_OF_RECOGNIZERS": 1, "MAX_NUM_OF_RECOGNIZERS": 6}
CONF_PSSM = {"PSEUDO_COUNT": 1e-10}
CONF_CON = {"AVG_MU": 25, "PSEUDO_COUNT": 1e-15}
CONF_SHAPE = {"PSEUDO_COUNT": 1e-100, "NUM_BINS": 25, "MIN_LENGTH": 5, "MAX_LENGTH": 9}
SHAPES = ("mgw", "prot", "roll", "helt")


def random_sequences(n: int, length: int, seed: int, alphabet: str = "acgt") -> list:
    """n iid-uniform sequences of the given length (lower case like generate_negative_set, search_organisms.py:585)."""
    rng = np.random.default_rng(seed)
    letters = np.frombuffer(alphabet.encode(), dtype=np.uint8)
    codes = rng.integers(0, len(letters), size=(n, length))
    return [letters[row].tobytes().decode() for row in codes]


def ensure_null_models(num_bins: int = 25, max_length: int = 9) -> dict:
    if not objects.null_models or max(objects.null_models["mgw"].keys()) < max_length:
        objects.set_null_models(build_null_models(num_bins, 5, max_length))
    return objects.null_models


def pwm_column(rng: np.random.Generator, n_sites: int = 100) -> dict:
    """Stick-breaking of n_sites counts over the four bases, every base >= 1 (organism_factory.py:312-350)."""
    left = n_sites
    counts = []
    for item in (3, 2, 1):
        c = int(rng.integers(1, left - item + 1)) if left > item else 1
        left -= c
        counts.append(c)
    counts.append(left)
    rng.shuffle(counts)
    p = (np.array(counts) / n_sites).tolist()
    return {"a": p[0], "g": p[1], "c": p[2], "t": p[3]}


def random_pssm(rng, length: int) -> objects.PssmObject:
    return objects.PssmObject(np.array([pwm_column(rng) for _ in range(length)]), CONF_PSSM)


def random_shape(rng, length: int | None = None, kind: str | None = None) -> objects.ShapeObject:
    ensure_null_models()
    kind = kind or SHAPES[int(rng.integers(0, 4))]
    length = length or int(rng.integers(5, 10))
    edges = objects.null_models[kind][length]["bins"]
    mu = float(rng.uniform(edges[0], edges[-1]))
    sigma = float(rng.exponential((edges[-1] - edges[0]) / 12))
    return objects.ShapeObject(kind, length, CONF_SHAPE, mu, sigma)


def random_connector(rng, max_seq_length: int, mu=None, sigma=None) -> objects.ConnectorObject:
    mu = float(rng.poisson(CONF_CON["AVG_MU"])) if mu is None else mu
    sigma = float(rng.exponential(max_seq_length / 12)) if sigma is None else sigma
    return objects.ConnectorObject(CONF_CON, max_seq_length, mu, sigma)


def assemble(org_id, recognizers, connectors) -> objects.OrganismObject:
    org = objects.OrganismObject(str(org_id), CONF_ORG)
    org.recognizers = list(recognizers)
    org.connectors = list(connectors)
    org.flatten()
    return org


def pssm_organisms(n_org: int, max_seq_length: int, seed: int, n_rec: int = 3, rec_len: int = 12,
                   mu_range=None, sigma_range=None) -> list:
    """Config 2 / config 4 organisms: n_rec PSSM recognizers of rec_len columns.  sigma_range: connector sigmas drawn
    uniformly from it instead of the factory's exponential (tight connectors, as fitted organisms have)."""
    rng = np.random.default_rng(seed)
    out = []
    for o in range(n_org):
        recs = [random_pssm(rng, rec_len) for _ in range(n_rec)]
        cons = []
        for _ in range(n_rec - 1):
            mu = None if mu_range is None else float(rng.uniform(*mu_range))
            sigma = None if sigma_range is None else float(rng.uniform(*sigma_range))
            cons.append(random_connector(rng, max_seq_length, mu, sigma))
        out.append(assemble(o, recs, cons))
    return out


def mixed_organisms(n_org: int, max_seq_length: int, seed: int, max_rec: int = 6, p_pssm: float = 0.5, sigma_range=None) -> list:
    """Config 3 organisms: n ~ U{1..max_rec}; each recognizer a PSSM (L ~ U{4..12}) with prob p_pssm, else a
    shape recognizer (uniform type, L ~ U{5..9})."""
    ensure_null_models()
    rng = np.random.default_rng(seed)
    out = []
    for o in range(n_org):
        n = int(rng.integers(1, max_rec + 1))
        recs = []
        for _ in range(n):
            if rng.random() < p_pssm:
                recs.append(random_pssm(rng, int(rng.integers(4, 13))))
            else:
                recs.append(random_shape(rng))
        if sigma_range is None:
            cons = [random_connector(rng, max_seq_length) for _ in range(n - 1)]
        else:
            cons = [random_connector(rng, max_seq_length, None, float(rng.uniform(*sigma_range))) for _ in range(n - 1)]
        out.append(assemble(o, recs, cons))
    return out


WORKLOADS = {
    # name: (organism builder kwargs, n_pos, n_neg, seq_len)
    "c2": dict(kind="pssm", n_org=256, n_pos=1000, n_neg=1000, seq_len=300),
    "c3": dict(kind="mixed", n_org=2000, n_pos=5000, n_neg=5000, seq_len=300),
    "c4": dict(kind="pssm", n_org=512, n_pos=2500, n_neg=2500, seq_len=2000, mu_range=(0, 1500)),
    # config 5: one generation of the 4k-population GA run (every step re-uploads the population, as a generation
    # of new children would); the mutation / crowding operators themselves are host logic outside the hot path
    "c5": dict(kind="mixed", n_org=4096, n_pos=20000, n_neg=20000, seq_len=300),
}


def make_workload(name: str, n_org: int | None = None, n_pos: int | None = None, n_neg: int | None = None):
    """(organisms, pos_sequences, neg_sequences) of a named BASELINE config (seeds: pos 1, neg 2, organisms 3)."""
    w = dict(WORKLOADS[name])
    n_org = n_org or w["n_org"]
    pos = random_sequences(n_pos or w["n_pos"], w["seq_len"], 1)
    neg = random_sequences(n_neg or w["n_neg"], w["seq_len"], 2)
    if w["kind"] == "pssm":
        orgs = pssm_organisms(n_org, w["seq_len"], 3, mu_range=w.get("mu_range"))
    else:
        orgs = mixed_organisms(n_org, w["seq_len"], 3)
    return orgs, pos, neg
