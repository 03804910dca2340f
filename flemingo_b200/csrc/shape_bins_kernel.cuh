// shape_bins_kernel.cuh -- batched shape scoring, the organism-independent half of K1.
//
// A shape recognizer's score at an offset is alt[b] - null[b], where the bin b depends only on
// (shape kind, recognizer length, bin edges, sequence, offset): the averaged pentamer score is binned against the
// edges (recognizer.c:187-589, _aux.c:184-217).  Edges come from the null model of (kind, length), so a whole
// population shares a handful of "shape classes".  This kernel evaluates every class on every sequence ONCE per
// dataset, in exact double arithmetic and the reference's summation order (the binning is discontinuous, so the
// average must be bit-identical), and stores the bin index as one byte per offset.  The placement kernels then
// score a shape recognizer with a 1-byte load and a table lookup.
#pragma once

struct ShapeBinParams {
  const unsigned *packed;    // 2-bit bases; non-ACGT bytes are stored as code 0, which is what the reference's
  const int *word_off;       // pentamer index does with them (recognizer.c:219-241 has no default case)
  const int *seq_len;
  int n_seq;
  const char *cls_kind;      // 'm','t','r','h'
  const int *cls_len;
  const int *cls_nb;         // number of bin edges
  const long long *cls_edges;  // offset into edges_pool
  const double *edges_pool;
  const double *tables;
  unsigned char *bins;       // [class][seq][row_stride]
  long long row_stride;
};

__global__ void __launch_bounds__(128) shape_bins_kernel(const ShapeBinParams q) {
  __shared__ double s_edges[256];
  const int seq = blockIdx.x, cls = blockIdx.y;
  const char kind = q.cls_kind[cls];
  const int L = q.cls_len[cls], nb = q.cls_nb[cls], np = L - 4;
  const int S = q.seq_len[seq];
  const double *edges = q.edges_pool + q.cls_edges[cls];
  const bool in_smem = nb <= 256;
  if (in_smem)
    for (int i = threadIdx.x; i < nb; i += blockDim.x) s_edges[i] = edges[i];
  __syncthreads();
  const double *e = in_smem ? s_edges : edges;
  const unsigned *w = q.packed + q.word_off[seq];
  unsigned char *out = q.bins + ((size_t)cls * q.n_seq + seq) * q.row_stride;
  const bool two_step = (kind == 'r' || kind == 'R' || kind == 'h' || kind == 'H');
  const double *tab = q.tables + ((kind == 'm' || kind == 'M') ? T_MGW : (kind == 't' || kind == 'T') ? T_PROT : (kind == 'r' || kind == 'R') ? T_ROLL : T_HELT);
  for (int p = threadIdx.x; p + L <= S; p += blockDim.x) {
    auto code = [&](int pos) { return (int)((__ldg(w + (pos >> 4)) >> (2 * (pos & 15))) & 3u); };
    int idx = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) idx = idx * 4 + code(p + k);
    double score;
    if (!two_step) {
      // _aux.c:211-217
      double sum = 0.0;
      for (int j = 0; j < np; ++j) {
        idx = ((idx << 2) | code(p + j + 4)) & 1023;
        sum += __ldg(tab + idx);
      }
      score = sum / (double)np;
    } else {
      // _aux.c:203-209 on pent = [TAB[idx_0..]] ++ [TAB[1024+idx_0..]] (recognizer.c:447-450)
      const int idx0 = ((idx << 2) | code(p + 4)) & 1023;
      int idxl = 0;
#pragma unroll
      for (int k = 0; k < 5; ++k) idxl = idxl * 4 + code(p + np - 1 + k);
      double sum = __ldg(tab + idx0) + __ldg(tab + 1024 + idxl);
      int id = idx;
      for (int j = 0; j < np; ++j) {
        id = ((id << 2) | code(p + j + 4)) & 1023;
        sum += __ldg(tab + id);
      }
      id = idx;
      for (int j = 0; j < np; ++j) {
        id = ((id << 2) | code(p + j + 4)) & 1023;
        sum += __ldg(tab + 1024 + id);
      }
      score = sum / (double)(2 * np + 2);
    }
    // _aux.c:184-190: first bin with edges[b-1] <= score < edges[b], otherwise the last bin
    int b = 1;
    for (; b < nb; ++b)
      if (score >= e[b - 1] && score < e[b]) break;
    if (b == nb) b = nb - 1;
    out[p] = (unsigned char)(b - 1);
  }
}

// ------------------------------------------------------------------------------------------------
// 2-bit packing of a DNA set on the device: the host uploads the raw ASCII bytes once, one CTA per sequence packs
// 16 bases per 32-bit word (A=0, G=1, C=2, T=3, case-insensitive, recognizer.c:143-163) plus a 1-bit "unknown byte"
// mask, and flags sequences that contain a non-ACGT byte.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned base_code_dev(unsigned char c) {
  switch (c | 0x20) {   // lower case
    case 'a': return 0u;
    case 'g': return 1u;
    case 'c': return 2u;
    case 't': return 3u;
    default: return 4u;
  }
}

constexpr int kK8Pad = 48;  // bytes of 4-mer codes kept past the end of a sequence (>= kMaxJ + longest group overshoot)

__global__ void __launch_bounds__(64) pack_sequences_kernel(const unsigned char *__restrict__ raw, const long long *__restrict__ offsets,
                                                            const int *__restrict__ word_off, const int *__restrict__ unk_off,
                                                            unsigned *__restrict__ packed, unsigned *__restrict__ unk,
                                                            unsigned char *__restrict__ has_unk, const int *__restrict__ k8_off,
                                                            unsigned char *__restrict__ k8) {
  const int seq = blockIdx.x;
  const unsigned char *b = raw + offsets[seq];
  const int len = (int)(offsets[seq + 1] - offsets[seq]);
  unsigned *pw = packed + word_off[seq];
  unsigned *uw = unk + unk_off[seq];
  int any = 0;
  // one thread per unknown-mask word = 32 bases = two packed words
  for (int u = threadIdx.x; u * 32 < len; u += blockDim.x) {
    unsigned lo = 0u, hi = 0u, mask = 0u;
    const int base = u * 32;
    const int cnt = min(32, len - base);
    for (int i = 0; i < cnt; ++i) {
      const unsigned c = base_code_dev(b[base + i]);
      if (c > 3u) mask |= 1u << i;
      else if (i < 16) lo |= c << (2 * i);
      else hi |= c << (2 * (i - 16));
    }
    uw[u] = mask;
    pw[2 * u] = lo;
    if (base + 16 < len) pw[2 * u + 1] = hi;
    any |= (mask != 0u);
  }
  any = __syncthreads_or(any);
  if (threadIdx.x == 0) has_unk[seq] = any ? 1 : 0;
  // 4-mer codes for the integer pass of place_fast_kernel: k8[q] = sum_t code(b[q+t]) << 2t (0 past the end and for
  // unknown bytes: sequences with unknown bytes never take that path), padded so that tiles may overshoot
  unsigned char *kq = k8 + 16ull * (unsigned long long)k8_off[seq];
  const int klen = (len + kK8Pad + 15) & ~15;
  for (int q = threadIdx.x; q < klen; q += blockDim.x) {
    unsigned code = 0u;
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const unsigned c = (q + t < len) ? base_code_dev(b[q + t]) : 0u;
      code |= (c & 3u) * (c < 4u) << (2 * t);
    }
    kq[q] = (unsigned char)code;
  }
}
