"""Data formats on either side of the hot path: FASTA ingest without Biopython,
organism JSON import / export round trip, and the population inequality statistic without the O(N^2) loop.

Reference behaviour mirrored:
  read_fasta_file     src/search_organisms.py:931-949 (Bio.SeqIO records -> list of sequence strings)
  import_organisms    src/objects/organism_factory.py:352-391, :444-476 (JSON -> flattened organisms)
  export_organisms    src/objects/organism_factory.py:478-559
  gini_RSV            src/search_organisms.py:1095-1142
"""
from __future__ import annotations

import json

import numpy as np

from . import objects
from .synthetic import CONF_CON, CONF_ORG, CONF_PSSM, CONF_SHAPE


def read_fasta_file(filename: str) -> list:
    """Sequences of a FASTA file as strings (multi-line records joined, headers dropped, case preserved)."""
    dataset, current, seen_header = [], [], False
    with open(filename) as handle:
        for line in handle:
            line = line.strip()
            if not line:
                continue
            if line.startswith(">"):
                if seen_header:
                    dataset.append("".join(current))
                current, seen_header = [], True
            elif seen_header:
                current.append(line)
    if seen_header:
        dataset.append("".join(current))
    return dataset


def import_organisms(file_name: str, max_seq_length: int, conf_org=None, conf_pssm=None, conf_con=None, conf_shape=None,
                     first_id: int = 1) -> list:
    """Organisms of a FLEMINGO JSON file, built with the mirror objects and flattened (ready for the GPU path).

    The shape null models must have been installed (objects.set_null_models / synthetic.ensure_null_models).
    """
    conf_org, conf_pssm = conf_org or CONF_ORG, conf_pssm or CONF_PSSM
    conf_con, conf_shape = conf_con or CONF_CON, conf_shape or CONF_SHAPE
    with open(file_name) as json_file:
        organism_json = json.load(json_file)
    out = []
    for k, organism in enumerate(organism_json):
        org = objects.OrganismObject(str(first_id + k), conf_org)
        for element in organism:
            kind = element["objectType"]
            if kind == "pssm":
                org.recognizers.append(objects.PssmObject(np.array(element["pwm"]), conf_pssm))
            elif kind == "shape":
                org.recognizers.append(objects.ShapeObject(element["recType"], element["length"], conf_shape, element["mu"], element["sigma"]))
            elif kind == "connector":
                org.connectors.append(objects.ConnectorObject(conf_con, max_seq_length, element["mu"], element["sigma"]))
        org.flatten()
        out.append(org)
    return out


def export_organisms(a_organisms: list, filename: str) -> None:
    """Writes organisms in the reference's JSON layout (recognizer, connector, recognizer, ...)."""
    def rec(r):
        if r.get_type() == 'p':
            pwm = r.pwm.tolist() if hasattr(r.pwm, "tolist") else list(r.pwm)
            return {"objectType": "pssm", "pwm": pwm}
        return {"objectType": "shape", "recType": r.type, "length": r.length, "mu": r._mu, "sigma": r._sigma}

    listing = []
    for org in a_organisms:
        items = []
        for i in range(len(org.recognizers) - 1):
            items.append(rec(org.recognizers[i]))
            items.append({"objectType": "connector", "mu": org.connectors[i]._mu, "sigma": org.connectors[i]._sigma})
        items.append(rec(org.recognizers[-1]))
        listing.append(items)
    with open(filename, "w+") as json_file:
        json.dump(listing, json_file, indent=2)


def gini_RSV(values_for_each_class) -> float:
    """Gini coefficient modified for negative values (Raffinetti, Siletti, Vernizzi), as the reference computes it,
    in O(N log N): sum_i sum_j |x_i - x_j| = 2 * sum_k (2k - N + 1) x_(k) over the ascending order statistics."""
    x = np.sort(np.asarray(values_for_each_class, dtype=np.float64))
    n = x.size
    if n == 0:
        return 0
    numerator = 2.0 * float(np.dot(2.0 * np.arange(n) - n + 1.0, x))
    mu = (n - 1) * float(np.abs(x).sum()) / n ** 2
    if mu == 0:
        return 0
    return numerator / (2 * n ** 2 * mu)
