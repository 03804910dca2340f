"""OrganismObject.flatten on arrays -- shared by the host mirror (objects.py) and by the hook that replaces the method of
the reference's own class (dropin.accelerate_reference)."""
from __future__ import annotations

import numpy as np

from . import builders


def flatten_organism(self) -> None:
    """OrganismObject.flatten (organism_object.py:1362-1504) on arrays: the same flat arrays, dtypes and attribute
    names, built by concatenating the array twins the accelerated builders leave next to every table instead of
    appending ~1500 Python floats one by one and converting the lists (150 us -> 25 us per organism)."""
    types, lengths, pssm_parts, model_parts, edge_parts, bin_nums = [], [], [], [], [], []
    for rec in self.recognizers:
        kind = rec.get_type()
        if kind == 'p':
            flat = getattr(rec, "_pssm_flat", None)
            if flat is None or getattr(rec, "_pssm_src", None) is not rec.pssm:
                flat = np.array([[col[b] for b in ('a', 'g', 'c', 't')] for col in rec.pssm], dtype=np.float64).ravel()
            pssm_parts.append(flat)
        else:
            model_parts.append(np.asarray(rec.null_model, dtype=np.float64))
            model_parts.append(builders.as_array(rec.alt_model))
            edge_parts.append(np.asarray(rec.edges, dtype=np.float64))
            bin_nums.append(len(rec.edges))
        lengths.append(rec.length)
        types.append(kind)
    empty = np.zeros(0, dtype=np.dtype('d'))
    self.recognizer_types = "".join(types)
    self.recognizers_flat = np.concatenate(pssm_parts) if pssm_parts else empty
    self.recognizer_lengths = np.array(lengths, dtype=np.dtype('i'))
    self.recognizer_models = np.concatenate(model_parts) if model_parts else empty
    self.recognizer_bin_nums = np.array(bin_nums, dtype=np.dtype('i'))
    self.recognizer_bin_edges = np.concatenate(edge_parts) if edge_parts else empty
    self.set_sum_recognizer_lengths()
    con_parts = []
    for con in self.connectors:
        con_parts.append(np.array([con._mu, con._sigma], dtype=np.float64))
        if self.is_precomputed == True:  # noqa: E712  (the reference's own test, organism_object.py:1487)
            con_parts.append(builders.as_array(con.stored_pdfs))
            con_parts.append(builders.as_array(con.stored_cdfs))
    self.connectors_scores_flat = np.concatenate(con_parts) if con_parts else empty
