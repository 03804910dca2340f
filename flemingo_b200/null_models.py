"""Shape null models: the binned distribution of a shape score over ALL k-mers of a given length.

Input producer of the hot path (not part of it): these arrays are the `null[]` / `edges[]` every shape
recognizer carries.  Restates /root/reference/src/models/null.py:19-134 (scores) and :137-279 (histograms)
with numpy over all 4^k k-mers at once; same dictionary layout as the reference keeps in
shape_object.null_models: {type: {length: {"frequencies": ln(freq) or NaN, "bins": edges}}}.

Summation order follows null.py, which differs from the C kernel's for Roll/HelT: null.py interleaves the two
step values per pentamer (`:97-103`), the C code lists all first steps and then all second steps
(recognizer.c:447-450).  The null model is defined by null.py, the placement score by the C code.
"""
from __future__ import annotations

import numpy as np

from .tables import table

SHAPES = ("mgw", "prot", "roll", "helt")


def kmer_shape_scores(shape: str, k: int) -> np.ndarray:
    """Score of every k-mer (itertools.product('AGCT', repeat=k) order; A=0 G=1 C=2 T=3)."""
    n = 4 ** k
    codes = (np.arange(n, dtype=np.int64)[:, None] >> (2 * np.arange(k - 1, -1, -1))[None, :]) & 3
    n_pent = k - 4
    weights = 4 ** np.arange(4, -1, -1)
    idx = [(codes[:, i:i + 5] * weights).sum(axis=1) for i in range(n_pent)]
    tab = table(shape)
    if shape in ("mgw", "prot"):
        total = np.zeros(n)
        for i in range(n_pent):
            total = total + tab[idx[i]]
        return total / n_pent
    total = np.zeros(n)
    for i in range(n_pent):
        total = total + tab[idx[i]]
        total = total + tab[1024 + idx[i]]
    return (tab[idx[0]] + total + tab[1024 + idx[n_pent - 1]]) / ((2 * n_pent) + 2)


def build_null_models(num_bins: int = 25, min_length: int = 5, max_length: int = 9) -> dict:
    models = {s: {} for s in SHAPES}
    for k in range(min_length, max_length + 1):
        for s in SHAPES:
            counts, edges = np.histogram(kmer_shape_scores(s, k), bins=num_bins)
            freqs = counts / sum(counts)
            with np.errstate(divide="ignore"):
                logf = np.where(freqs != 0.0, np.log(np.where(freqs != 0.0, freqs, 1.0)), np.nan)
            models[s][k] = {"frequencies": logf, "bins": edges}
    return models
