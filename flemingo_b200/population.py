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


def _g(x, mu, sigma):
    # connector_object.py:15-16
    return -(1 / 2) * ((x - mu) / sigma) ** 2


def rescale_connector_block(block: np.ndarray, mu: float, sigma: float, pseudo_count: float, M: int,
                            sequence_length: int, sum_recognizer_lengths: int) -> None:
    """In-place restatement of ConnectorObject.adjust_scores (connector_object.py:454-487).

    `block` is the connector's [pdf[0..M) ++ cdf[0..M)] slice (the reference passes `c_scores[offset:]`).
    Natural logs on purpose: that is what the reference writes (unlike the log2 tables of :175-206).
    sigma == 0 raises ZeroDivisionError exactly as the reference does.
    """
    sigma = min(sigma, 1E-100)
    top = sequence_length - sum_recognizer_lengths
    g_curr = _g(top, mu, sigma)
    h_curr = 1
    weights = [1 + pseudo_count]
    total = weights[0]
    for i in range(top - 1, -1, -1):
        g_next = _g(i, mu, sigma)
        h_next = h_curr * math.exp(g_next - g_curr)
        weights.append(h_next + pseudo_count)
        total += h_next + pseudo_count
        g_curr = g_next
        h_curr = h_next
    probs = [h / total for h in reversed(weights)]
    running = 0.0
    for i, p in enumerate(probs):
        running += p
        block[i] = np.log(p)
        block[M + i] = np.log(running)


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
    uploaded_con_size: int = -1
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


def pack_population(organisms) -> PackedPopulation:
    """Packs organisms (flattened, see module docstring) into one PackedPopulation."""
    n_org = len(organisms)
    rec_begin = np.zeros(n_org + 1, dtype=np.int32)
    pool_begin = np.zeros(n_org + 1, dtype=np.int64)
    con_base = np.zeros(n_org, dtype=np.int64)
    con_M = np.zeros(n_org, dtype=np.int32)
    sum_len = np.zeros(n_org, dtype=np.int32)
    pcs = np.zeros(n_org, dtype=np.float64)
    types, lens, nbs, datas, pool_parts, con_parts, ids = [], [], [], [], [], [], []
    classes, class_of, cls_kind, cls_len, cls_nb, cls_edges, edge_parts = {}, [], [], [], [], [], []
    edge_off = 0
    pool_off = 0
    con_off = 0
    for o, org in enumerate(organisms):
        rtypes = org.recognizer_types
        n = len(rtypes)
        if n < 1:
            raise ValueError(f"organism {getattr(org, '_id', o)} has no recognizers")
        rlens = org.recognizer_lengths
        rl = rlens.tolist() if isinstance(rlens, np.ndarray) else [int(x) for x in rlens]
        if len(rl) != n:
            raise ValueError("organism is not flattened: recognizer_lengths does not match recognizer_types")
        total_len = sum(rl)
        pssm = _f64(org.recognizers_flat)
        all_pssm = rtypes == 'p' * n
        if not all_pssm:
            models = _f64(org.recognizer_models)
            edges = _f64(org.recognizer_bin_edges)
            bin_nums = np.asarray(org.recognizer_bin_nums, dtype=np.int32)
        m_off = a_off = e_off = shape_i = 0
        org_parts = []
        org_pool = 0
        if all_pssm:
            # common case, no per-recognizer python work: the PSSMs are already back to back in recognizers_flat
            if pssm.size != 4 * total_len:
                raise ValueError("recognizers_flat does not match the PSSM lengths")
            at = pool_off
            for L in rl:
                datas.append(at)
                at += 4 * L
            types.extend([112] * n)
            lens.extend(rl)
            nbs.extend([0] * n)
            class_of.extend([-1] * n)
            org_parts.append(pssm)
            org_pool = 4 * total_len
        for i, t in enumerate(rtypes if not all_pssm else ()):
            L = rl[i]
            datas.append(pool_off + org_pool)
            if t in "pP":
                org_parts.append(pssm[m_off:m_off + 4 * L])
                if org_parts[-1].size != 4 * L:
                    raise ValueError("recognizers_flat is shorter than the PSSM lengths require")
                m_off += 4 * L
                org_pool += 4 * L
                nbs.append(0)
                class_of.append(-1)
            else:
                nb = int(bin_nums[shape_i])
                # parse_org (organism.c:159-185): null = models+a, alt = models+a+(nb-1), edges = edges+e
                org_parts.append(models[a_off:a_off + 2 * (nb - 1)])
                org_parts.append(edges[e_off:e_off + nb])
                key = (t.lower(), L, edges[e_off:e_off + nb].tobytes())
                if key not in classes:
                    classes[key] = len(classes)
                    cls_kind.append(ord(t))
                    cls_len.append(L)
                    cls_nb.append(nb)
                    cls_edges.append(edge_off)
                    edge_parts.append(edges[e_off:e_off + nb])
                    edge_off += nb
                class_of.append(classes[key])
                a_off += 2 * (nb - 1)
                e_off += nb
                org_pool += 3 * nb - 2
                shape_i += 1
                nbs.append(nb)
            types.append(ord(t))
            lens.append(L)
        pool_parts.extend(org_parts)
        pool_off += org_pool
        rec_begin[o + 1] = rec_begin[o] + n
        pool_begin[o + 1] = pool_off
        sum_len[o] = total_len
        con = _f64(org.connectors_scores_flat)
        con_base[o] = con_off
        if n > 1:
            if con.size == 2 * (n - 1):
                raise NotImplementedError(
                    "organism.PRECOMPUTE=false (connectors without pdf/cdf tables) uses the reference's legacy "
                    "float maths (connector.c:70-72, _aux.c:47-139); the GPU path requires precomputed tables")
            M = int(org.connectors[0].max_seq_length)
            if con.size != (n - 1) * (2 * M + 2):
                raise ValueError("connectors_scores_flat does not match (n-1) * (2*max_seq_length + 2)")
            con_M[o] = M
            pcs[o] = float(org.connectors[0].pseudo_count)
            con_parts.append(con)
            con_off += con.size
        ids.append(getattr(org, "_id", str(o)))
    rec_pool = np.concatenate(pool_parts) if pool_parts else np.zeros(0)
    con_pool = np.concatenate(con_parts) if con_parts else np.zeros(0)
    return PackedPopulation(
        n_org=n_org, rec_begin=rec_begin, rec_type=np.asarray(types, dtype=np.uint8),
        rec_len=np.asarray(lens, dtype=np.int32), rec_nb=np.asarray(nbs, dtype=np.int32),
        rec_data=np.asarray(datas, dtype=np.int64), pool_begin=pool_begin,
        rec_pool=np.ascontiguousarray(rec_pool, dtype=np.float64), con_M=con_M,
        con_pool=np.ascontiguousarray(con_pool, dtype=np.float64), con_base=con_base, sum_len=sum_len,
        pseudo_count=pcs, ids=ids, n_base_con=int(con_pool.size),
        rec_class=np.asarray(class_of, dtype=np.int32), cls_kind=np.asarray(cls_kind, dtype=np.uint8),
        cls_len=np.asarray(cls_len, dtype=np.int32), cls_nb=np.asarray(cls_nb, dtype=np.int32),
        cls_edges=np.asarray(cls_edges, dtype=np.int64),
        edges_pool=np.ascontiguousarray(np.concatenate(edge_parts) if edge_parts else np.zeros(0), dtype=np.float64))


def placement_cost(packed: PackedPopulation, seq_len: int) -> np.ndarray:
    """Relative cost of placing each organism on one sequence of length seq_len: (n-1) * P^2 DP cells
    plus the score rows.  Used to balance organisms across GPUs."""
    P = np.maximum(seq_len - packed.sum_len.astype(np.int64) + 1, 1)
    n = packed.n_rec.astype(np.int64)
    return (n - 1) * P * (P + 1) // 2 + P * packed.sum_len
