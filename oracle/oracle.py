"""Python face of the CPU oracles -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this
module; nothing under flemingo_b200/ does.  Parity status: PINNED (see placement_oracle.c header).

Two checkers:
  * `restated_*`  : oracle/_build/libplacement_oracle.so, our plain-C restatement (placement_oracle.c);
  * `reference_*` : oracle/_ref/_multiplacement*.so, the UNMODIFIED reference extension compiled from
                    /root/reference by `make -C oracle ref` (present only if it was built where the reference
                    tree exists; it travels to the GPU box as a built file).
The numpy fitness restatement follows src/objects/organism_object.py:848-956 line by line.
"""
from __future__ import annotations

import ctypes as C
import importlib.util
import subprocess
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
LIB = HERE / "_build" / "libplacement_oracle.so"
REF_DIR = HERE / "_ref"
TABLES = HERE.parent / "flemingo_b200" / "data" / "tables.bin"

FO_OK, FO_TOO_LARGE, FO_NAN_NULL, FO_UNSUPPORTED, FO_BAD_INPUT = 0, -1, -2, -3, -4


def build(ref: bool = True) -> None:
    """Compiles the restatement and, when /root/reference is present, the reference extension."""
    subprocess.run(["make", "-s", "-C", str(HERE), "oracle"], check=True)
    if ref and Path("/root/reference/extensions/multiplacement/src/_interface.c").exists():
        subprocess.run(["make", "-s", "-C", str(HERE), "ref"], check=True)
        # the reference's Python driver and objects for the end-to-end GA tests (git-ignored copy, see the Makefile)
        subprocess.run(["make", "-s", "-C", str(HERE), "refsrc"], check=True)


_lib = None
_tables = None


def _load():
    global _lib, _tables
    if _lib is None:
        if not LIB.exists():
            build(ref=False)
        _lib = C.CDLL(str(LIB))
        _lib.fo_place.restype = C.c_int
        _lib.fo_place_batch.restype = C.c_int
        _tables = np.fromfile(TABLES, dtype="<f8")
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def restated_calculate(sequence: bytes, rec_types: bytes, rec_matrices, rec_lengths, con_matrices, rec_scores,
                       con_scores, con_lengths, max_length, bin_freqs, bin_edges, num_bins) -> int:
    """Same argument list as _multiplacement.calculate; returns the oracle's status code."""
    lib = _load()
    f = lambda a: np.ascontiguousarray(a, dtype=np.float64)
    i = lambda a: np.ascontiguousarray(a, dtype=np.int32)
    rm, rl, cm, bf, be, nb = f(rec_matrices), i(rec_lengths), f(con_matrices), f(bin_freqs), f(bin_edges), i(num_bins)
    return lib.fo_place(_p(_tables), C.c_char_p(bytes(sequence)), C.c_int(len(sequence)), C.c_char_p(bytes(rec_types)),
                        C.c_int(len(rec_types)), _p(rm), _p(rl), _p(cm), C.c_int(int(max_length)), _p(bf), _p(be), _p(nb),
                        _p(rec_scores), _p(con_scores), _p(con_lengths))


def reference_module():
    """The compiled reference extension, or None if oracle/_ref was not built."""
    hits = sorted(REF_DIR.glob("_multiplacement*.so"))
    if not hits:
        return None
    name = "_flemingo_reference_multiplacement"
    if name in sys.modules:
        return sys.modules[name]
    spec = importlib.util.spec_from_file_location("_multiplacement", hits[0])
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    sys.modules[name] = mod
    return mod


def _marshal(org, sequence: str):
    """The per-call marshalling of OrganismObject.get_placement (organism_object.py:1298-1329)."""
    n = len(org.recognizer_types)
    gaps = np.empty(n, dtype=np.int32)
    gap_scores = np.empty(max(n - 1, 0), dtype=np.float64)
    rec_scores = np.empty(n + 1, dtype=np.float64)
    c_scores = np.copy(org.connectors_scores_flat)
    max_length = -1
    if n > 1:
        max_length = org.connectors[0].max_seq_length - 1
        for i in range(len(org.connectors)):
            if float(len(sequence)) + 0.5 < org.connectors[i]._mu:
                idx = len(sequence) - org.sum_recognizer_lengths
                if org.connectors[i].stored_pdfs[idx] <= np.log(1E-15):
                    offset = sum([(org.connectors[i].max_seq_length * 2 + 2) for i in range(0, i)]) + 2
                    org.connectors[i].adjust_scores(c_scores[offset:], len(sequence), org.sum_recognizer_lengths)
    return c_scores, max_length, rec_scores, gap_scores, gaps


def place(org, sequence: str, impl: str = "restated"):
    """One placement of a flattened organism: (energy, rec_scores[n], con_scores[n-1], gaps[n], status)."""
    if org.sum_recognizer_lengths > len(sequence):
        raise ValueError("organism longer than sequence")
    c_scores, max_length, rs, cs, gaps = _marshal(org, sequence)
    args = (bytes(sequence, "ASCII"), bytes(org.recognizer_types, "ASCII"), org.recognizers_flat, org.recognizer_lengths,
            c_scores, rs, cs, gaps, max_length, org.recognizer_models, org.recognizer_bin_edges, org.recognizer_bin_nums)
    status = 0
    if impl == "reference":
        reference_module().calculate(*args)
    else:
        status = restated_calculate(*args)
    return float(rs[-1]), rs[:-1].copy(), cs.copy(), gaps.copy(), status


def place_all(organisms, sequences, impl: str = "restated", trace: bool = True):
    """energies[o, s] (+ gaps / rec_scores / con_scores [o, s, max_rec]) by looping `place`."""
    n_org, n_seq = len(organisms), len(sequences)
    mr = max(len(o.recognizer_types) for o in organisms) if n_org else 0
    E = np.zeros((n_org, n_seq))
    G = np.zeros((n_org, n_seq, mr), dtype=np.int32)
    R = np.zeros((n_org, n_seq, mr))
    Cs = np.zeros((n_org, n_seq, mr))
    for o, org in enumerate(organisms):
        n = len(org.recognizer_types)
        for s, seq in enumerate(sequences):
            e, rs, cs, gaps, _ = place(org, seq, impl)
            E[o, s] = e
            G[o, s, :n] = gaps
            R[o, s, :n] = rs
            Cs[o, s, :n - 1] = cs
    return E, G, R, Cs


def mean_and_se(energy_scores, method: str, gamma: float):
    """organism_object.py:891-936 on a list of energies."""
    energy_scores = list(energy_scores)
    n_to_trim = round(gamma * len(energy_scores))
    if n_to_trim == 0 or method == "Welch":
        mean = np.mean(energy_scores)
        stdev = max(1, np.std(energy_scores, ddof=1))
        sterr = stdev / np.sqrt(len(energy_scores))
    elif method == "Yuen":
        n_to_trim = min(n_to_trim, int((len(energy_scores) - 1) / 2))
        energy_scores.sort()
        mean = np.mean(energy_scores[n_to_trim:-n_to_trim])
        stdev = max(1, np.std(energy_scores, ddof=1))
        sterr = stdev / np.sqrt(len(energy_scores))
    elif method == "Trim-left-M":
        energy_scores.sort()
        mean = np.mean(energy_scores[n_to_trim:])
        stdev = max(1, np.std(energy_scores, ddof=1))
        sterr = stdev / np.sqrt(len(energy_scores))
    elif method == "Trim-left-M-SD":
        energy_scores.sort()
        trimmed = energy_scores[n_to_trim:]
        mean = np.mean(trimmed)
        stdev = max(1, np.std(trimmed, ddof=1))
        sterr = stdev / np.sqrt(len(trimmed))
    else:
        raise ValueError('Unknown method "' + method + '".')
    return mean, sterr


def fitness_stats(e_pos, e_neg, method: str = "Welch", gamma: float = 0.2):
    """[fitness, M_p, SE_p, M_n, SE_n] (organism_object.py:938-956)."""
    with np.errstate(all="ignore"):
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            m_p, se_p = mean_and_se(e_pos, method, gamma)
            m_n, se_n = mean_and_se(e_neg, method, gamma)
            fit = (m_p - m_n) / (se_p ** 2 + se_n ** 2) ** (1 / 2)
    return np.array([fit, m_p, se_p, m_n, se_n], dtype=np.float64)
