"""Batched table builders: the input producers of the placement path, off the per-object Python loops.

    connector_tables(mu, sigma, M, pc)   <->  ConnectorObject.set_precomputed_pdfs_cdfs   connector_object.py:175-206
    pssm_values(pwm_probabilities, pc)   <->  PssmObject.recalculate_pssm                 pssm_object.py:368-400
    shape_alt_models(mu, sigma, edges, pc)  <->  ShapeObject.set_alt_model                shape_object.py:156-180

The arithmetic up to the logarithm runs in the library (fl_build_* in include/flemingo_b200.h: threaded C++, libm erf
like math.erf); the logarithm itself is numpy's, applied to the whole batch at once, because that is the function the
reference calls and its vector kernels are not bit-identical to libm's.  Results are bit-identical to the reference's
objects (tests/test_builders_cpu.py, golden flat arrays).  These run on the host: no CUDA context is needed.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _cabi


class TableList(list):
    """A list of immutable numpy scalars (stored_pdfs / stored_cdfs / alt_model) that copy.deepcopy duplicates in one
    step.  The reference clones every organism with deepcopy (organism_factory.py:893 clone_organism), which otherwise
    walks the ~2 * max_seq_length float objects of each connector one by one -- 4 ms per clone, the largest host cost of
    a generation once placement runs on the GPU.  Behaves as a plain list everywhere else (`a + b` gives a list).
    `np` is the same table as a float64 array (never written to), which lets flatten() concatenate arrays instead of
    converting lists."""

    np = None

    @classmethod
    def of(cls, array):
        out = cls(array)
        out.np = array
        return out

    def __deepcopy__(self, memo):
        out = TableList(self)
        out.np = self.np
        return out


def as_array(table) -> "np.ndarray":
    """float64 array of a table attribute: its array twin if it is a TableList that has one, else a conversion."""
    twin = getattr(table, "np", None)
    if twin is not None and len(twin) == len(table):
        return twin
    return np.array(table, dtype=np.float64)


def _ptr(a):
    return C.c_void_p(a.__array_interface__["data"][0])


def _check(rc: int, what: str) -> None:
    if rc != 0:
        raise ValueError(f"{what}: bad arguments (code {rc})")


def connector_tables(mu, sigma, M: int, pseudo_count: float, n_threads: int = 0):
    """(stored_pdfs[n, M], stored_cdfs[n, M]) of n connectors, log2 like the reference (-inf where its log2 is)."""
    lib = _cabi.load_library()
    mu = np.array(mu, dtype=np.float64, ndmin=1)
    sigma = np.array(sigma, dtype=np.float64, ndmin=1)
    n = int(mu.size)
    args = np.empty((2, n, int(M)), dtype=np.float64)
    _check(lib.fl_build_connector_tables(_ptr(mu), _ptr(sigma), n, int(M), float(pseudo_count), _ptr(args[0]), _ptr(args[1]),
                                         int(n_threads)), "fl_build_connector_tables")
    with np.errstate(divide="ignore", invalid="ignore"):
        np.log2(args, out=args)
    return args[0], args[1]


def round_decimals(values: np.ndarray, ndigits: int = 2) -> np.ndarray:
    """Python's round(x, ndigits) on every element of a float64 array (in place when contiguous)."""
    lib = _cabi.load_library()
    out = np.ascontiguousarray(values, dtype=np.float64)
    _check(lib.fl_round_decimals(_ptr(out), int(out.size), int(ndigits)), "fl_round_decimals")
    return out


def pssm_values(probabilities: np.ndarray, pseudo_count: float) -> np.ndarray:
    """round(log2(4 p + pc), 2) for an array of PWM probabilities (any shape)."""
    with np.errstate(divide="ignore"):
        vals = np.log2(4.0 * np.asarray(probabilities, dtype=np.float64) + pseudo_count)
    return round_decimals(vals, 2)


def shape_alt_models(mu, sigma, edges_list, pseudo_count: float) -> list:
    """alt models of several shape recognizers: alt[b] = ln(mass_b + pc) - ln(auc); mu already clamped to the edges."""
    lib = _cabi.load_library()
    n = len(edges_list)
    mu = np.ascontiguousarray(np.atleast_1d(mu), dtype=np.float64)
    sigma = np.ascontiguousarray(np.atleast_1d(sigma), dtype=np.float64)
    n_edges = np.fromiter((len(e) for e in edges_list), dtype=np.int32, count=n)
    begin = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(n_edges, out=begin[1:])
    edges = np.ascontiguousarray(np.concatenate([np.asarray(e, dtype=np.float64) for e in edges_list])) if n else np.zeros(0)
    mass = np.empty(int(begin[-1]) - n, dtype=np.float64)
    auc = np.empty(n, dtype=np.float64)
    _check(lib.fl_build_shape_masses(_ptr(mu), _ptr(sigma), _ptr(edges), _ptr(begin), _ptr(n_edges), n, float(pseudo_count),
                                     _ptr(mass), _ptr(auc)), "fl_build_shape_masses")
    with np.errstate(divide="ignore", invalid="ignore"):
        log_mass, log_auc = np.log(mass), np.log(auc)
    out = []
    for s in range(n):
        lo = int(begin[s]) - s
        out.append(log_mass[lo:lo + int(n_edges[s]) - 1] - log_auc[s])
    return out


def fitness_finalize(stats: np.ndarray) -> np.ndarray:
    """Recomputes column 0 of [n, 5] fitness statistics from M_pos, SE_pos, M_neg, SE_neg with the reference's own
    arithmetic (fl_fitness_finalize); in place.  For rows that came off the device without passing through fl_fitness."""
    lib = _cabi.load_library()
    if not (stats.dtype == np.float64 and stats.flags.c_contiguous and stats.ndim == 2 and stats.shape[1] == 5):
        raise ValueError("stats must be a C-contiguous float64 [n, 5] array")
    _check(lib.fl_fitness_finalize(_ptr(stats), int(stats.shape[0])), "fl_fitness_finalize")
    return stats
