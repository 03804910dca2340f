"""Host-side packer: a list of organisms -> the struct-of-arrays `fl_population` of include/flemingo_b200.h.

An "organism" here is anything that carries the flat arrays `OrganismObject.flatten()` produces
(reference: src/objects/organism_object.py:1362-1504): `recognizer_types`, `recognizers_flat`,
`recognizer_lengths`, `recognizer_models`, `recognizer_bin_edges`, `recognizer_bin_nums`,
`connectors_scores_flat`, `sum_recognizer_lengths`, and `connectors` (objects with `_mu`, `_sigma`,
`pseudo_count`, `max_seq_length`).  Both flemingo_b200.objects.OrganismObject and the reference's own class
qualify, which is what makes the GPU path a drop-in.

The packer also owns the one piece of per-call host arithmetic of `get_placement`: the connector rescale
(organism_object.py:1312-1329 -> ConnectorObject.adjust_scores, connector_object.py:454-487).  It depends on
(organism, connector, sequence length) only, so it is evaluated once per length class and appended to the
connector pool as a variant table.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np

SHAPE_TYPES = "mtrh"


def rescale_connector_block(block: np.ndarray, mu: float, sigma: float, pseudo_count: float, M: int,
                            sequence_length: int, sum_recognizer_lengths: int) -> None:
    """What ConnectorObject.adjust_scores (connector_object.py:454-487) writes, in closed form.

    `block` is the connector's [pdf[0..M) ++ cdf[0..M)] slice (the reference passes `c_scores[offset:]`).
    The reference re-normalises the Gaussian over the gaps 0..top (top = S - sum L) with sigma' = min(sigma, 1e-100).
    It is only called when mu > S + 0.5, so every gap is at least 1.5 below mu and, with sigma' <= 1e-100, the ratio
    of neighbouring weights is exp(-(2 (mu - x) - 1) / (2 sigma'^2)) <= exp(-1e200) = 0.0: all the mass sits on the
    largest gap.  What is left is the pseudo count:

        weight(top) = 1 + pc,  weight(x < top) = 0 + pc,  total = sequential sum in that order,
        pdf[x] = ln(weight(x) / total),  cdf[x] = ln(running sum of the probabilities from gap 0 up)

    (natural logs on purpose, unlike the log2 tables of :175-206).  The sums are accumulated in the reference's order
    so the doubles are bit-identical (golden `rescale_*` cases).  sigma == 0 raises ZeroDivisionError as the reference
    does; a sigma' so small that the exponent overflows makes the reference produce NaN everywhere, and so do we.
    """
    sigma = min(sigma, 1E-100)
    top = sequence_length - sum_recognizer_lengths
    if sigma == 0:
        raise ZeroDivisionError("float division by zero")
    n = top + 1
    if not (float(sequence_length) + 0.5 < mu):
        raise ValueError("adjust_scores is only defined for mu beyond the sequence (organism_object.py:1317)")
    # the exponent of the gap farthest from mu, evaluated like the reference does (connector_object.py:15-16): a float
    # power that overflows raises OverflowError there and here; an infinite quotient turns its weights into NaN
    if not math.isfinite(-(1 / 2) * ((0 - mu) / sigma) ** 2):
        block[:n] = np.nan
        block[M:M + n] = np.nan
        return
    weights = np.full(n, 0.0 + pseudo_count)      # gap order 0..top
    weights[top] = 1 + pseudo_count
    total = np.add.accumulate(weights[::-1])[-1]  # the reference adds from the largest gap down
    probs = weights / total
    block[:n] = np.log(probs)
    block[M:M + n] = np.log(np.add.accumulate(probs))


@dataclass
class PackedPopulation:
    """Struct-of-arrays population (host numpy arrays, C-contiguous, dtypes as in fl_population)."""
    n_org: int
    rec_begin: np.ndarray     # int32 [n_org+1]
    rec_type: np.ndarray      # uint8 [n_rec_total]
    rec_len: np.ndarray       # int32 [n_rec_total]
    rec_nb: np.ndarray        # int32 [n_rec_total]
    rec_data: np.ndarray      # int64 [n_rec_total]
    pool_begin: np.ndarray    # int64 [n_org+1]
    rec_pool: np.ndarray      # float64
    con_M: np.ndarray         # int32 [n_org]
    con_pool: np.ndarray      # float64, base tables followed by rescaled variants
    con_base: np.ndarray      # int64 [n_org] offset of each organism's own tables in con_pool
    sum_len: np.ndarray       # int32 [n_org]
    pseudo_count: np.ndarray  # float64 [n_org] connector pseudo count
    # shape classes: distinct (kind, length, edges) of the shape recognizers (fl_population.rec_class ...)
    rec_class: np.ndarray = None      # int32 [n_rec_total], -1 for PSSM
    cls_kind: np.ndarray = None       # uint8 [n_classes]
    cls_len: np.ndarray = None        # int32 [n_classes]
    cls_nb: np.ndarray = None         # int32 [n_classes]
    cls_edges: np.ndarray = None      # int64 [n_classes]
    edges_pool: np.ndarray = None     # float64
    ids: list = field(default_factory=list)
    # (organism index, sequence length) -> offset of the rescaled variant in con_pool
    variants: dict = field(default_factory=dict)
    n_base_con: int = 0
    _con_org: np.ndarray = None
    _con_mu_pos: np.ndarray = None

    @property
    def n_rec(self) -> np.ndarray:
        return np.diff(self.rec_begin)

    @property
    def max_rec(self) -> int:
        return int(self.n_rec.max()) if self.n_org else 0

    def con_begin_for_lengths(self, lengths) -> np.ndarray:
        """con_begin[class][organism] for the given class lengths; creates rescaled variants on demand.

        Trigger, per connector (organism_object.py:1317-1329): float(S) + 0.5 < mu and
        pdf[S - sum_len] <= ln(1e-15).
        """
        lengths = [int(x) for x in lengths]
        out = np.empty((len(lengths), self.n_org), dtype=np.int64)
        n_rec = self.n_rec
        threshold = np.log(1E-15)
        extra = []
        extra_off = self.con_pool.size
        if self._con_org is None:
            # per-connector views for the vectorised trigger test: owner organism, position of mu in con_pool
            nc = np.maximum(n_rec - 1, 0)
            self._con_org = np.repeat(np.arange(self.n_org), nc)
            first = np.cumsum(nc) - nc
            within = np.arange(int(nc.sum())) - np.repeat(first, nc)
            self._con_mu_pos = self.con_base[self._con_org] + within * (2 * self.con_M[self._con_org].astype(np.int64) + 2)
        mu_all = self.con_pool[self._con_mu_pos] if self._con_mu_pos.size else np.zeros(0)
        for ci, S in enumerate(lengths):
            out[ci, :] = self.con_base
            hot = np.unique(self._con_org[float(S) + 0.5 < mu_all]) if mu_all.size else ()
            for o in hot:
                key = (int(o), S)
                if key in self.variants:
                    out[ci, o] = self.variants[key]
                    continue
                M = int(self.con_M[o])
                stride = 2 * M + 2
                base = int(self.con_base[o])
                nc = int(n_rec[o]) - 1
                idx = S - int(self.sum_len[o])
                if idx < 0 or idx >= M:
                    continue  # inadmissible pair; the library reports it
                block = None
                for c in range(nc):
                    mu = float(self.con_pool[base + c * stride])
                    if float(S) + 0.5 < mu and self.con_pool[base + c * stride + 2 + idx] <= threshold:
                        if block is None:
                            block = self.con_pool[base:base + nc * stride].copy()
                        sigma = float(block[c * stride + 1])
                        rescale_connector_block(block[c * stride + 2:(c + 1) * stride], mu, sigma,
                                                float(self.pseudo_count[o]), M, S, int(self.sum_len[o]))
                if block is not None:
                    self.variants[key] = extra_off
                    out[ci, o] = extra_off
                    extra.append(block)
                    extra_off += block.size
        if extra:
            self.con_pool = np.concatenate([self.con_pool] + extra)
        return out


def _f64(a) -> np.ndarray:
    if isinstance(a, np.ndarray) and a.dtype == np.float64 and a.flags.c_contiguous:
        return a
    return np.ascontiguousarray(a, dtype=np.float64)


def _segment_exclusive_cumsum(values: np.ndarray, seg_first: np.ndarray, seg_id: np.ndarray) -> np.ndarray:
    """Exclusive prefix sum of `values` restarting at every segment (seg_first[s] = first index of segment s)."""
    if values.size == 0:
        return values.astype(np.int64)
    csum = np.cumsum(values, dtype=np.int64) - values
    return csum - csum[seg_first][seg_id]


def pack_population(organisms) -> PackedPopulation:
    """Packs organisms (flattened, see module docstring) into one PackedPopulation.

    Vectorised: a handful of list comprehensions and numpy concatenations, no per-recognizer Python work.
    Pool layout per organism: recognizers_flat (all its PSSMs, [col][A,G,C,T]) followed by recognizer_models
    (null[nb-1] ++ alt[nb-1] per shape recognizer, organism.c:159-172).  Bin edges live in the shape classes.
    """
    n_org = len(organisms)
    type_strs = [org.recognizer_types for org in organisms]
    n_rec = np.fromiter((len(t) for t in type_strs), dtype=np.int64, count=n_org)
    if n_org and int(n_rec.min()) < 1:
        bad = int(np.argmin(n_rec))
        raise ValueError(f"organism {getattr(organisms[bad], '_id', bad)} has no recognizers")
    rec_begin = np.zeros(n_org + 1, dtype=np.int32)
    np.cumsum(n_rec, out=rec_begin[1:])
    n_total = int(rec_begin[-1])
    rec_type = np.frombuffer("".join(type_strs).encode("ascii"), dtype=np.uint8) if n_total else np.zeros(0, np.uint8)
    len_parts = [org.recognizer_lengths for org in organisms]
    rec_len = np.concatenate(len_parts).astype(np.int32, copy=False) if n_org else np.zeros(0, np.int32)
    if rec_len.size != n_total:
        raise ValueError("an organism is not flattened: recognizer_lengths does not match recognizer_types")
    is_p = (rec_type == 112) | (rec_type == 80)
    org_of = np.repeat(np.arange(n_org), n_rec)
    seg_first = rec_begin[:-1].astype(np.int64)

    # parts are concatenated as they are (numpy converts lists / other dtypes on the way); sizes come from len()
    pssm_parts = [org.recognizers_flat for org in organisms]
    pssm_sizes = np.fromiter(map(len, pssm_parts), dtype=np.int64, count=n_org)
    size_p = np.where(is_p, 4 * rec_len.astype(np.int64), 0)
    if not np.array_equal(np.bincount(org_of, weights=size_p, minlength=n_org).astype(np.int64), pssm_sizes):
        raise ValueError("recognizers_flat does not match the PSSM lengths of an organism")
    any_shape = not bool(is_p.all())
    rec_nb = np.zeros(n_total, dtype=np.int32)
    rec_class = np.full(n_total, -1, dtype=np.int32)
    cls_kind = np.zeros(0, np.uint8)
    cls_len = np.zeros(0, np.int32)
    cls_nb = np.zeros(0, np.int32)
    cls_edges = np.zeros(0, np.int64)
    edges_pool = np.zeros(0, np.float64)
    if any_shape:
        model_parts = [org.recognizer_models for org in organisms]
        model_sizes = np.fromiter(map(len, model_parts), dtype=np.int64, count=n_org)
        nb_shapes = np.concatenate([np.asarray(org.recognizer_bin_nums, dtype=np.int32) for org in organisms])
        shape_idx = np.nonzero(~is_p)[0]
        if nb_shapes.size != shape_idx.size:
            raise ValueError("recognizer_bin_nums does not match the number of shape recognizers")
        rec_nb[shape_idx] = nb_shapes
        size_m = np.where(is_p, 0, 2 * (rec_nb.astype(np.int64) - 1))
        if not np.array_equal(np.bincount(org_of, weights=size_m, minlength=n_org).astype(np.int64), model_sizes):
            raise ValueError("recognizer_models does not match the shape recognizers of an organism")
        pool_parts = [x for pair in zip(pssm_parts, model_parts) for x in pair]
        org_pool = pssm_sizes + model_sizes
        # shape classes: distinct (kind, length, edges) -- grouped by the number of edges so rows can be compared
        all_edges = np.concatenate([org.recognizer_bin_edges for org in organisms]).astype(np.float64, copy=False)
        if all_edges.size != int(nb_shapes.sum()):
            raise ValueError("recognizer_bin_edges does not match recognizer_bin_nums")
        e_start = np.cumsum(nb_shapes, dtype=np.int64) - nb_shapes
        kinds = rec_type[shape_idx] | 32   # lower case
        lens_s = rec_len[shape_idx]
        ck, cl, cn, ce, pools = [], [], [], [], []
        e_off = 0
        for nb in np.unique(nb_shapes):
            sel = np.nonzero(nb_shapes == nb)[0]
            rows = all_edges[(e_start[sel][:, None] + np.arange(nb)[None, :])]
            key = np.concatenate([kinds[sel][:, None].astype(np.float64), lens_s[sel][:, None].astype(np.float64), rows], axis=1)
            # NaN-free keys: np.unique on the byte view keeps bit patterns distinct
            view = np.ascontiguousarray(key).view(np.dtype((np.void, key.dtype.itemsize * key.shape[1]))).ravel()
            _, first, inverse = np.unique(view, return_index=True, return_inverse=True)
            rec_class[shape_idx[sel]] = len(ck) + inverse.ravel()
            for f in first:
                ck.append(int(kinds[sel][f]))
                cl.append(int(lens_s[sel][f]))
                cn.append(int(nb))
                ce.append(e_off)
                pools.append(rows[f])
                e_off += int(nb)
        cls_kind = np.asarray(ck, dtype=np.uint8)
        cls_len = np.asarray(cl, dtype=np.int32)
        cls_nb = np.asarray(cn, dtype=np.int32)
        cls_edges = np.asarray(ce, dtype=np.int64)
        edges_pool = np.ascontiguousarray(np.concatenate(pools), dtype=np.float64)
    else:
        size_m = np.zeros(n_total, dtype=np.int64)
        pool_parts = pssm_parts
        org_pool = pssm_sizes
    pool_begin = np.zeros(n_org + 1, dtype=np.int64)
    np.cumsum(org_pool, out=pool_begin[1:])
    rec_pool = np.concatenate(pool_parts).astype(np.float64, copy=False) if n_org else np.zeros(0)
    # recognizer data offsets: PSSMs in order inside the organism's PSSM segment, shape models in order inside its model segment
    off_p = _segment_exclusive_cumsum(size_p, seg_first, org_of)
    off_m = _segment_exclusive_cumsum(size_m, seg_first, org_of)
    rec_data = pool_begin[:-1][org_of] + np.where(is_p, off_p, pssm_sizes[org_of] + off_m)

    sum_len = np.bincount(org_of, weights=rec_len, minlength=n_org).astype(np.int32) if n_total else np.zeros(n_org, np.int32)
    con_parts = [org.connectors_scores_flat for org in organisms]
    con_sizes = np.fromiter(map(len, con_parts), dtype=np.int64, count=n_org)
    multi = n_rec > 1
    if bool(np.any(multi & (con_sizes == 2 * (n_rec - 1)))):
        raise NotImplementedError(
            "organism.PRECOMPUTE=false (connectors without pdf/cdf tables) uses the reference's legacy "
            "float maths (connector.c:70-72, _aux.c:47-139); the GPU path requires precomputed tables")
    firsts = [org.connectors[0] if m else None for org, m in zip(organisms, multi.tolist())]
    con_M = np.fromiter((f.max_seq_length if f is not None else 0 for f in firsts), dtype=np.int32, count=n_org)
    pcs = np.fromiter((f.pseudo_count if f is not None else 0.0 for f in firsts), dtype=np.float64, count=n_org)
    # one table length and one pseudo count per organism: the reference takes both from each connector
    # (organism_object.py:1303-1329), the packed layout from the first one -- refuse organisms where they differ
    for org, f in zip(organisms, firsts):
        if f is not None and len(org.connectors) > 1:
            for c in org.connectors[1:]:
                if c.max_seq_length != f.max_seq_length or c.pseudo_count != f.pseudo_count:
                    raise ValueError(f"organism {getattr(org, '_id', '?')}: connectors with different max_seq_length / "
                                     "pseudo_count in one organism are not supported")
    if bool(np.any(multi & (con_sizes != (n_rec - 1) * (2 * con_M.astype(np.int64) + 2)))):
        raise ValueError("connectors_scores_flat does not match (n-1) * (2*max_seq_length + 2)")
    con_sizes = np.where(multi, con_sizes, 0)
    con_base = np.cumsum(con_sizes) - con_sizes
    if bool(multi.all()):
        con_pool = np.concatenate(con_parts).astype(np.float64, copy=False)
    elif bool(multi.any()):
        con_pool = np.concatenate([c for c, m in zip(con_parts, multi) if m]).astype(np.float64, copy=False)
    else:
        con_pool = np.zeros(0)
    return PackedPopulation(
        n_org=n_org, rec_begin=rec_begin, rec_type=np.ascontiguousarray(rec_type), rec_len=np.ascontiguousarray(rec_len),
        rec_nb=rec_nb, rec_data=np.ascontiguousarray(rec_data, dtype=np.int64), pool_begin=pool_begin,
        rec_pool=np.ascontiguousarray(rec_pool, dtype=np.float64), con_M=con_M,
        con_pool=np.ascontiguousarray(con_pool, dtype=np.float64), con_base=con_base.astype(np.int64), sum_len=sum_len,
        pseudo_count=pcs, ids=[getattr(org, "_id", str(i)) for i, org in enumerate(organisms)],
        n_base_con=int(con_pool.size), rec_class=rec_class, cls_kind=cls_kind, cls_len=cls_len, cls_nb=cls_nb,
        cls_edges=cls_edges, edges_pool=edges_pool)


def placement_cost(packed: PackedPopulation, seq_len: int) -> np.ndarray:
    """Relative cost of placing each organism on one sequence of length seq_len: (n-1) * P^2 DP cells
    plus the score rows.  Used to balance organisms across GPUs."""
    P = np.maximum(seq_len - packed.sum_len.astype(np.int64) + 1, 1)
    n = packed.n_rec.astype(np.int64)
    return (n - 1) * P * (P + 1) // 2 + P * packed.sum_len
