"""Engine: the Python face of the C ABI -- one CUDA context, resident sequence sets, batched placement.

Batched equivalents of the reference's per-organism Python loops:

    Engine.place(...)     <->  for org: for seq: org.get_placement(seq)        (organism_object.py:958-966,1263)
    Engine.fitness(...)   <->  for org: org.set_fitness(pos, neg, method, g)   (organism_object.py:938-956)
    Engine.calculate(...) <->  _multiplacement.calculate(...)                  (src/_interface.c:45-185)
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass

import numpy as np

from . import _cabi
from .population import PackedPopulation, pack_population
from .tables import load_tables


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


@dataclass
class SequenceSet:
    set_id: int
    n_seq: int
    lengths: np.ndarray      # int64 [n_seq]
    class_len: list          # distinct lengths, ascending (the library's class order)
    sequences: list          # the caller's strings (kept for PlacementObject construction)
    blob: bytes = b""        # the uploaded bytes (to recognise an unchanged dataset)


@dataclass
class PlacementBatch:
    """Result of Engine.place: energies[o, s]; with trace also gaps/rec_scores/con_scores[o, s, i]
    (gaps[...,0] = start of the first recognizer, gaps[...,i] = gap before recognizer i; organism.c:398-427)."""
    energies: np.ndarray
    gaps: np.ndarray | None = None
    rec_scores: np.ndarray | None = None
    con_scores: np.ndarray | None = None


class DeviceStats:
    """A [rows, 5] float64 array in device memory owned by the library (valid until the next fitness call)."""

    def __init__(self, ptr: int, rows: int, n_org: int, stream: int):
        self.ptr, self.rows, self.n_org, self.stream = ptr, rows, n_org, stream

    @property
    def __cuda_array_interface__(self):
        # no "stream" entry: consumers would synchronise with it on import; callers order their work by running on
        # Engine.stream_handle (self.stream) instead
        return {"shape": (self.rows, 5), "typestr": "<f8", "data": (self.ptr, False), "version": 2, "strides": None}


class Engine:
    def __init__(self, device: int | None = None):
        self._lib = _cabi.load_library()
        if device is None:
            device = int(os.environ.get("LOCAL_RANK", "0"))
        self.device = device
        tables = load_tables()
        handle = C.c_void_p()
        _cabi.check(self._lib.fl_ctx_create(device, _ptr(tables), tables.size, C.byref(handle)))
        self._ctx = handle
        self._packed: PackedPopulation | None = None
        self._uploaded_con_size = -1
        self._sets: dict[int, SequenceSet] = {}
        # True: skip the integer filter pass and place every pair with the exact fp64 kernel (same results)
        self.exact_only = os.environ.get("FLEMINGO_EXACT_ONLY", "0") == "1"

    # -- lifetime -------------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_ctx", None):
            self._lib.fl_ctx_destroy(self._ctx)
            self._ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def stream_handle(self) -> int:
        """cudaStream_t all kernels of this engine run on (wrap with torch.cuda.ExternalStream to time them)."""
        return int(self._lib.fl_ctx_stream(self._ctx) or 0)

    @property
    def launch_count(self) -> int:
        return int(self._lib.fl_launch_count(self._ctx))

    def path_stats(self, detail: bool = False):
        """(pairs sent to the integer filter pass, pairs it handed to the exact kernel in the end) since the last call;
        detail=True adds the pairs that took the int32 second chance after the packed 16-bit pass."""
        a, b, c = C.c_longlong(0), C.c_longlong(0), C.c_longlong(0)
        _cabi.check(self._lib.fl_path_stats(self._ctx, C.byref(a), C.byref(b), C.byref(c)))
        if detail:
            return int(a.value), int(b.value), int(c.value)
        return int(a.value), int(b.value)

    def measure_dpx_rate(self) -> dict:
        """DP cells per second the DPX instruction of the DP stage can issue on this device (fl_measure_dpx_rate)."""
        a, b, n, k = C.c_double(0), C.c_double(0), C.c_int(0), C.c_int(0)
        _cabi.check(self._lib.fl_measure_dpx_rate(self._ctx, C.byref(a), C.byref(b), C.byref(n), C.byref(k)))
        return {"cells_per_s_i32": a.value, "cells_per_s_i16x2": b.value, "n_sm": n.value, "sm_clock_khz": k.value}

    def sync(self):
        _cabi.check(self._lib.fl_sync(self._ctx))

    # -- uploads --------------------------------------------------------------------------------
    @staticmethod
    def join_sequences(sequences):
        """(bytes of all sequences back to back, int64 lengths): what upload_sequences sends; callers that feed several
        engines with one dataset join it once."""
        lengths = np.fromiter(map(len, sequences), dtype=np.int64, count=len(sequences))
        try:
            joined = "".join(sequences).encode("ascii")          # the common case: a list of str
        except TypeError:
            joined = b"".join(s if isinstance(s, (bytes, bytearray)) else bytes(s, "ASCII") for s in sequences)
        if int(lengths.sum()) != len(joined):
            raise ValueError("sequences must be ASCII")
        return joined, lengths

    def upload_sequences(self, sequences, set_id: int = 0, reuse: bool = False, joined=None) -> SequenceSet:
        """Uploads a DNA set (list of str/bytes); bases are 2-bit packed on the way (fl_upload_sequences).

        reuse=True skips the upload when the set resident under set_id holds byte-for-byte the same sequences: the
        GA evaluates thousands of generations on one dataset, so the packed copy stays in HBM across calls.
        joined: the result of join_sequences(sequences), if the caller already has it.
        """
        old = self._sets.get(set_id)
        seqs = sequences
        joined, lengths = joined if joined is not None else self.join_sequences(seqs)
        offsets = np.zeros(len(seqs) + 1, dtype=np.int64)
        np.cumsum(lengths, out=offsets[1:])
        if reuse and old is not None and old.n_seq == len(seqs) and old.blob == joined and np.array_equal(old.lengths, lengths):
            return old
        blob = np.frombuffer(joined + b"\0", dtype=np.uint8)
        _cabi.check(self._lib.fl_upload_sequences(self._ctx, set_id, _ptr(blob), _ptr(offsets), len(seqs)))
        sset = SequenceSet(set_id=set_id, n_seq=len(seqs), lengths=lengths,
                           class_len=[int(x) for x in np.unique(lengths)], sequences=list(sequences), blob=joined)
        self._sets[set_id] = sset
        return sset

    def upload_population(self, population) -> PackedPopulation:
        """Uploads a population: a PackedPopulation or a list of flattened organisms."""
        packed = population if isinstance(population, PackedPopulation) else pack_population(population)
        self._upload_packed(packed)
        return packed

    def _upload_packed(self, packed: PackedPopulation):
        pop = _cabi.FlPopulation()
        pop.n_org = packed.n_org
        pop.n_rec_total = int(packed.rec_type.size)
        pop.rec_begin = _ptr(packed.rec_begin)
        pop.rec_type = _ptr(packed.rec_type)
        pop.rec_len = _ptr(packed.rec_len)
        pop.rec_nb = _ptr(packed.rec_nb)
        pop.rec_data = _ptr(packed.rec_data)
        pop.pool_begin = _ptr(packed.pool_begin)
        pop.rec_pool = _ptr(packed.rec_pool)
        pop.n_pool = int(packed.rec_pool.size)
        pop.con_M = _ptr(packed.con_M)
        pop.con_pool = _ptr(packed.con_pool)
        pop.n_con = int(packed.con_pool.size)
        pop.rec_class = _ptr(packed.rec_class)
        pop.n_classes = int(packed.cls_kind.size)
        pop.cls_kind = _ptr(packed.cls_kind)
        pop.cls_len = _ptr(packed.cls_len)
        pop.cls_nb = _ptr(packed.cls_nb)
        pop.cls_edges = _ptr(packed.cls_edges)
        pop.edges_pool = _ptr(packed.edges_pool)
        pop.n_edges = int(packed.edges_pool.size)
        _cabi.check(self._lib.fl_upload_population(self._ctx, C.byref(pop)))
        self._uploaded_con_size = int(packed.con_pool.size)   # per engine: one packed population may feed several GPUs
        self._packed = packed

    # -- batched placement ----------------------------------------------------------------------
    def place(self, seqset: SequenceSet | int, trace: bool = False, fetch: bool = True) -> PlacementBatch | None:
        """Places every organism of the uploaded population on every sequence of the set.

        fetch=False leaves the energies on the device (for Engine.fitness) and returns None without
        synchronising.
        """
        if self._packed is None:
            raise _cabi.FlemingoError(_cabi.FL_E_STATE, "no population uploaded")
        sset = self._sets[seqset] if isinstance(seqset, int) else seqset
        packed = self._packed
        con_begin = packed.con_begin_for_lengths(sset.class_len)
        if packed.con_pool.size != self._uploaded_con_size:
            # a rescaled connector variant was appended (rare): grow the device pool, keep resident energies
            _cabi.check(self._lib.fl_update_connectors(self._ctx, _ptr(packed.con_pool), int(packed.con_pool.size)))
            self._uploaded_con_size = int(packed.con_pool.size)
        n_org, n_seq, mr = packed.n_org, sset.n_seq, packed.max_rec
        energies = gaps = rsc = csc = None
        if fetch:
            energies = np.empty((n_org, n_seq), dtype=np.float64)
            if trace:
                gaps = np.empty((n_org, n_seq, mr), dtype=np.int32)
                rsc = np.empty((n_org, n_seq, mr), dtype=np.float64)
                csc = np.empty((n_org, n_seq, mr), dtype=np.float64)
        flags = (_cabi.FL_WANT_TRACE if trace else 0) | (_cabi.FL_EXACT_ONLY if self.exact_only else 0)
        _cabi.check(self._lib.fl_place_batch(self._ctx, sset.set_id, _ptr(con_begin), flags, _ptr(energies), _ptr(gaps),
                                             _ptr(rsc), _ptr(csc)))
        if not fetch:
            return None
        return PlacementBatch(energies, gaps, rsc, csc)

    def fitness(self, pos: SequenceSet | int, neg: SequenceSet | int, method: str = "Welch", gamma: float = 0.2) -> np.ndarray:
        """[n_org, 5] = fitness, M_pos, SE_pos, M_neg, SE_neg from the resident energies of both sets."""
        if method not in _cabi.FIT_METHODS:
            # organism_object.py:933-935
            raise ValueError('Unknown method "' + str(method) + '". Please check ' +
                             'the FITNESS_FUNCTION paramter in the config file.')
        ps = pos if isinstance(pos, int) else pos.set_id
        ns = neg if isinstance(neg, int) else neg.set_id
        out = np.empty((self._packed.n_org, 5), dtype=np.float64)
        _cabi.check(self._lib.fl_fitness(self._ctx, ps, ns, _cabi.FIT_METHODS[method], float(gamma), _ptr(out)))
        return out

    def fitness_dev(self, pos: SequenceSet | int, neg: SequenceSet | int, method: str = "Welch", gamma: float = 0.2,
                    pad_rows: int = 0) -> "DeviceStats":
        """Same reduction, but the [max(n_org, pad_rows), 5] statistics stay in HBM (fl_fitness_dev): no copy, no
        synchronisation.  The result exposes __cuda_array_interface__, so `torch.as_tensor(stats, device=...)` is a
        zero-copy view -- the send buffer of the multi-GPU all-gather (flemingo_b200.distributed)."""
        if method not in _cabi.FIT_METHODS:
            raise ValueError('Unknown method "' + str(method) + '". Please check ' +
                             'the FITNESS_FUNCTION paramter in the config file.')
        ps = pos if isinstance(pos, int) else pos.set_id
        ns = neg if isinstance(neg, int) else neg.set_id
        ptr = C.c_void_p()
        _cabi.check(self._lib.fl_fitness_dev(self._ctx, ps, ns, _cabi.FIT_METHODS[method], float(gamma), int(pad_rows), C.byref(ptr)))
        return DeviceStats(int(ptr.value or 0), max(self._packed.n_org, int(pad_rows)), self._packed.n_org, self.stream_handle)

    def evaluate(self, organisms, pos_sequences, neg_sequences, method: str = "Welch", gamma: float = 0.2,
                 reuse_sets: bool = True) -> np.ndarray:
        """One generation's fitness evaluation through host buffers: upload, place on both sets, reduce."""
        # sequences first: their copies and the packing kernel run while the host packs the population
        pos = self.upload_sequences(pos_sequences, 0, reuse=reuse_sets)
        neg = self.upload_sequences(neg_sequences, 1, reuse=reuse_sets)
        self.upload_population(organisms)
        self.place(pos, fetch=False)
        self.place(neg, fetch=False)
        return self.fitness(pos, neg, method, gamma)

    # -- single placement with the reference's argument list ------------------------------------------
    def calculate(self, sequence, rec_types, rec_matrices, rec_lengths, con_matrices, rec_scores, con_scores,
                  con_lengths, max_length, bin_freqs, bin_edges, num_bins) -> None:
        n_rec = len(rec_types)
        seq = np.frombuffer(bytes(sequence) + b"\0", dtype=np.uint8)
        types = np.frombuffer(bytes(rec_types) + b"\0", dtype=np.uint8)

        def f64(a):
            return np.ascontiguousarray(a, dtype=np.float64)

        def i32(a):
            return np.ascontiguousarray(a, dtype=np.int32)

        rm, rl, cm = f64(rec_matrices), i32(rec_lengths), f64(con_matrices)
        bf, be, nb = f64(bin_freqs), f64(bin_edges), i32(num_bins)
        # outputs are written in place, like the reference does through the buffer protocol
        for name, arr, dt, need in (("rec_scores", rec_scores, np.float64, n_rec + 1),
                                    ("con_scores", con_scores, np.float64, n_rec - 1),
                                    ("con_lengths", con_lengths, np.int32, n_rec)):
            if not isinstance(arr, np.ndarray) or arr.dtype != dt or not arr.flags.c_contiguous or arr.size < need:
                raise TypeError(f"{name} must be a C-contiguous {np.dtype(dt).name} array with at least {need} elements")
        _cabi.check(self._lib.fl_calculate(self._ctx, _ptr(seq), len(sequence), _ptr(types), n_rec, _ptr(rm), _ptr(rl), _ptr(cm),
                                           int(cm.size), _ptr(rec_scores), _ptr(con_scores) if n_rec > 1 else None,
                                           _ptr(con_lengths), int(max_length), _ptr(bf), _ptr(be), _ptr(nb)))


_default_engine: Engine | None = None


def get_engine() -> Engine:
    """Process-wide engine on cuda:LOCAL_RANK (created on first use)."""
    global _default_engine
    if _default_engine is None:
        _default_engine = Engine()
    return _default_engine


def reset_engine():
    global _default_engine
    if _default_engine is not None:
        _default_engine.close()
    _default_engine = None
