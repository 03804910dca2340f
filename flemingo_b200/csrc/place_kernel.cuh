// place_kernel.cuh -- K1 + K2 fused: recognizer score rows, gap-energy tables, chain-placement DP, argmax and
// traceback, exact in IEEE double (included by flemingo_b200.cu).
//
// Reference semantics (paths relative to /root/reference/extensions/multiplacement/src):
//   score rows   organism/recognizer.c:130-168 (PSSM), :187-589 (shape), _aux.c:184-217
//   gap energy   organism/connector.c:31-66, auc pick organism/organism.c:336-340
//   DP           organism/organism.c:321-393       argmax + traceback  _aux.c:149-158, organism.c:398-427
//
// Work decomposition
//   CTA  = one organism x a chunk of sequences of one length class: PSSM / shape tables and the gap-energy tables
//          G[c][g] (sequence independent; the reference recomputes them per DP cell) live in shared memory.
//   warp = one (organism, sequence) pair at a time; its cumulative rows C_i[0..P) live in shared memory.
//   lane = blocks of J consecutive offsets j.  For a block the lane walks k upward in steps of J keeping
//          J accumulators and a 2J-wide sliding window of G in REGISTERS, so a J x J tile of DP cells costs
//          2J shared-memory loads (J of prev, J of G) instead of J*J.
//   The k <= j triangle is balanced by FOLDING: a lane owns block b and block NB-1-b of a round of <= 64 blocks,
//   so every lane executes the same number of tiles; both jobs run through ONE loop so the warp never diverges
//   in the hot code.
#pragma once
#include <type_traits>

#define T_MGW 0
#define T_PROT 1024
#define T_ROLL 2048
#define T_HELT 4096
#define T_DEN 6144

#define BIG_NEG (-1e10)
#define BIG_POS (1e10)

constexpr int kGPad = 32;          // -inf entries on both sides of every G table (>= largest tile size J)
constexpr int kMaxSmem = 232448;   // 227 KB opt-in limit per CTA on sm_100
constexpr int kMaxJ = 21;

#define ERR_NAN_NULL 1

struct RecDesc {
  int kind;  // 0 PSSM, 1 MGW, 2 ProT, 3 Roll, 4 HelT
  int len;
  int nb;    // number of bin edges (shape only)
  int data;  // offset into the organism's slice of the recognizer pool (doubles)
  int fwd;   // sum of the lengths of the recognizers before this one (organism.c:273,392)
  int cls;   // shape class (index into the precomputed bin rows), -1 for a PSSM
};

struct PlaceParams {
  // population (device pointers)
  const int *rec_begin;
  const char *rec_type;
  const int *rec_len;
  const int *rec_nb;
  const long long *rec_data;
  const long long *pool_begin;
  const double *rec_pool;
  const int *con_M;
  const double *con_pool;
  const long long *con_begin;  // [n_org] for this class
  const int *sum_len;
  const int *org_list;         // organisms handled by this launch (all share the tile size J)
  // sequence set
  const unsigned *packed;
  const unsigned *unk;
  const int *word_off;
  const int *unk_off;
  const int *order;
  const unsigned char *k8g;  // 4-mer codes of every sequence (pack_sequences_kernel); sequence s starts at 16 * k8_off[s]
  const int *k8_off;
  int class_begin, class_count, S, n_seq_total;
  // data tables and precomputed shape bins (shape_bins_kernel.cuh)
  const double *tables;
  const int *rec_class;           // [n_rec_total] shape class of a recognizer, -1 for PSSM
  const unsigned char *bins;      // [class][seq][bins_row_stride] bin index per offset
  long long bins_row_stride;
  // outputs
  double *E;
  int *gaps;
  double *rsc;
  double *csc;
  int max_rec;
  int *err_flag;
  // plan
  int seqs_per_cta, chunks;
  int pool_stride, g_stride, row_stride, code_stride, nrows;
};

__device__ __forceinline__ double neg_inf() { return __longlong_as_double(0xfff0000000000000LL); }

__device__ __forceinline__ int code0(unsigned char c) { return (c & 3) * (c < 4); }

// Score of a shape recognizer from its precomputed bin (recognizer.c:246-262): alt[b] - null[b], floored at -1e10;
// selecting a NaN null bin is the error the reference exits on.
__device__ __forceinline__ double score_shape_bin(const RecDesc &r, const double *s_pool, int b, int &err) {
  const double *nullf = s_pool + r.data;
  const double n0 = nullf[b];
  double v = nullf[(r.nb - 1) + b] - n0;
  if (v < BIG_NEG) v = BIG_NEG;
  if (n0 != n0) err |= ERR_NAN_NULL;
  return v;
}

// Score of one recognizer at offset `pos` (relative to its forward offset, organism.c:328).
// codes: the sequence as byte codes (0..3 = A,G,C,T; 4 = any other byte); bins_seq: this sequence's bin rows
// (class 0), cls_stride: distance between classes.
__device__ __forceinline__ double score_rec(const RecDesc &r, const double *s_pool, const unsigned char *codes, int pos,
                                            const unsigned char *__restrict__ bins_seq, size_t cls_stride, int &err) {
  if (r.kind == 0) {
    // recognizer.c:130-168: left-to-right sum from 0.0; an unknown byte adds nothing
    const unsigned char *c = codes + r.fwd + pos;
    const double *m = s_pool + r.data;
    double s = 0.0;
    for (int col = 0; col < r.len; ++col) {
      const int b = c[col];
      if (b < 4) s += m[4 * col + b];
    }
    return s;
  }
  return score_shape_bin(r, s_pool, bins_seq[(size_t)r.cls * cls_stride + r.fwd + pos], err);
}

// connector.c:35-66 for one gap length; D = DEN_EXPANSION
__device__ __forceinline__ double gap_energy(const double *__restrict__ pdf, double auc, int g, int n, int eff,
                                             const double *__restrict__ D) {
  if (g + n + 1 <= eff) {
    double num = __ldg(pdf + g);
    if (num < BIG_NEG) num = BIG_NEG;
    num -= auc;
    const double den = (__ldg(D + (eff - (g + 1)) - 1) - __ldg(D + (n - 1) - 1) - __ldg(D + eff - (g + 1) - (n - 1) - 1)) -
                       (__ldg(D + eff - 1) - __ldg(D + n - 1) - __ldg(D + eff - n - 1));
    double res = num - den;
    if (res < BIG_NEG) return BIG_NEG;
    if (res > BIG_POS) return BIG_POS;
    return res;
  }
  return BIG_NEG;
}

// ------------------------------------------------------------------------------------------------
// One DP stage for one pair, executed by a TEAM of lanes:  cur[j] = max_{k<=j} (prev[k] + G[j-k]),  j < P.
// (R_i[j] is added afterwards by the caller; fl(max + r) == max fl(. + r) because x -> fl(x+r) is monotone.)
// Op selects the arithmetic: exact double (DADD + compare/select) or int32 fixed point (one DPX VIADDMNMX).
// G points at gap 0 and is readable on [-kGPad, P-1+kGPad] with `lowest` outside [0, P).
// prev must be readable and cur writable up to ceil(P/J)*J - 1; for doubles the values beyond P-1 are never selected because
// their G is -inf, for ints the caller keeps them at 0 so that 0 + lowest cannot wrap.
// ------------------------------------------------------------------------------------------------
struct OpF64 {
  typedef double T;
  static __device__ __forceinline__ T cell(T acc, T p, T g) {
    const T c = p + g;
    return (acc < c) ? c : acc;
  }
};
struct OpI32 {
  typedef int T;
  static __device__ __forceinline__ T cell(T acc, T p, T g) { return __viaddmax_s32(p, g, acc); }
};

struct TileState {
  int b;      // row block of the current job (rows b*J .. b*J+J-1)
  int kb;     // next k block
  int kend;   // number of k blocks of the current job (b+1); -1 when idle
  int kstep;  // 1 while working, 0 when idle
  int next_b; // row block of the second job, or -1
};

// Job bookkeeping shared by tile_step and tile_last: when the current job has walked all its k blocks, flush its J
// rows, then start the folded partner block (or go idle) and load the first window of the new job into `A`.
template <int J, class Op>
__device__ __forceinline__ void tile_switch(typename Op::T (&A)[J], typename Op::T (&acc)[J], TileState &st,
                                            typename Op::T *cur, const typename Op::T *__restrict__ G,
                                            typename Op::T lowest) {
  typedef typename Op::T T;
  if (st.kb == st.kend) {
    // rows beyond P-1 of the last block land in the row's padding (the caller owns >= ceil(P/J)*J entries)
    typename Op::T *dst = cur + st.b * J;
#pragma unroll
    for (int t = 0; t < J; ++t) dst[t] = acc[t];
    if (st.next_b >= 0) {
      st.b = st.next_b;
      st.next_b = -1;
      st.kend = st.b + 1;
    } else {
      st.b = 0;
      st.kend = -1;
      st.kstep = 0;
    }
    st.kb = 0;
    const T *ga = G + st.b * J + (J - 1);
#pragma unroll
    for (int t = 0; t < J; ++t) {
      acc[t] = lowest;
      A[t] = ga[-t];
    }
  }
}

template <int J, class Op>
__device__ __forceinline__ void tile_step(typename Op::T (&A)[J], typename Op::T (&B)[J], typename Op::T (&acc)[J],
                                          TileState &st, const typename Op::T *prev,
                                          typename Op::T *cur, const typename Op::T *__restrict__ G, int P,
                                          typename Op::T lowest) {
  typedef typename Op::T T;
  // in-place stages: the loads every lane did in the previous step are ordered before the flush a lane may do now
  __syncwarp();
  tile_switch<J, Op>(A, acc, st, cur, G, lowest);
  // A[i] = G[c - i], B[i] = G[c - J - i] with c = (b - kb)*J + J-1: row t at step u needs G[(b-kb)*J + t - u]
  const T *gb = G + (st.b - st.kb) * J - 1;
  const T *pk = prev + st.kb * J;
  T pv[J];
#pragma unroll
  for (int u = 0; u < J; ++u) {
    B[u] = gb[-u];
    pv[u] = pk[u];
  }
#pragma unroll
  for (int u = 0; u < J; ++u) {
#pragma unroll
    for (int t = 0; t < J; ++t) {
      const T g = (u <= t) ? A[J - 1 - t + u] : B[u - t - 1];
      acc[t] = Op::cell(acc[t], pv[u], g);
    }
  }
  st.kb += st.kstep;
}

// The LAST step of a round: every lane that still works is on the diagonal tile of its second job (kb == b), where
// only the cells k <= j exist -- J(J+1)/2 of the J*J -- and only the `A` window is needed.
template <int J, class Op>
__device__ __forceinline__ void tile_last(typename Op::T (&A)[J], typename Op::T (&acc)[J], TileState &st,
                                          const typename Op::T *prev, typename Op::T *cur,
                                          const typename Op::T *__restrict__ G, typename Op::T lowest) {
  typedef typename Op::T T;
  __syncwarp();
  tile_switch<J, Op>(A, acc, st, cur, G, lowest);
  const T *pk = prev + st.kb * J;
  T pv[J];
#pragma unroll
  for (int u = 0; u < J; ++u) pv[u] = pk[u];
#pragma unroll
  for (int u = 0; u < J; ++u) {
#pragma unroll
    for (int t = u; t < J; ++t) acc[t] = Op::cell(acc[t], pv[u], A[J - 1 - t + u]);
  }
  st.kb += st.kstep;
}

// `tl` = lane index inside the team (0..TEAM-1).  Rounds of 2*TEAM row blocks; inside a round lane tl owns block
// r0+tl and the mirrored block r0+nbr-1-tl, so every lane of the team runs the same number of tiles.
//
// `cur` may be the SAME buffer as `prev` (in-place stage, one row per pair in shared memory).  That is safe because
// of the order of the work: rounds run from the highest blocks down, and inside a round a lane does its UPPER block
// first.  A job on block b starts at k block 0 and is flushed after b+1 steps; every other job of the round that
// still needs k block b started at the same step 0 and read it at its step b, i.e. before.  The LOWER blocks are all
// flushed together after the last step, and a lower block only ever reads k blocks below the upper half.  Lower
// rounds read lower k blocks than anything flushed so far.  (The lanes of a team run these steps in lockstep.)
template <int J, int TEAM, class Op>
__device__ __forceinline__ void dp_stage(const typename Op::T *prev, typename Op::T *cur,
                                         const typename Op::T *__restrict__ G, int P, int tl, typename Op::T lowest) {
  typedef typename Op::T T;
  const int NB = (P + J - 1) / J;
  for (int r0 = ((NB - 1) / (2 * TEAM)) * (2 * TEAM); r0 >= 0; r0 -= 2 * TEAM) {
    const int nbr = min(2 * TEAM, NB - r0);
    const int half = (nbr + 1) >> 1;
    const int bA = r0 + tl, bB = r0 + nbr - 1 - tl;
    TileState st;
    if (tl < half) {
      st.b = bB;  // upper block first (equal to bA for the middle lane of an odd round: a single job)
      st.kend = bB + 1;
      st.kstep = 1;
      st.next_b = (bB > bA) ? bA : -1;
    } else {
      st.b = 0;
      st.kend = -1;
      st.kstep = 0;
      st.next_b = -1;
    }
    st.kb = 0;
    T acc[J], A[J], B[J];
    {
      const T *ga = G + st.b * J + (J - 1);
#pragma unroll
      for (int t = 0; t < J; ++t) {
        acc[t] = lowest;
        A[t] = ga[-t];
      }
    }
    // (bB+1) + (bA+1) steps for every folded lane; exactly `iters` of them run, the last one -- the diagonal tile of
    // the lower block for every lane -- in its triangular form
    const int iters = 2 * r0 + nbr + 1;
    int it = 0;
    for (; it + 2 < iters; it += 2) {
      tile_step<J, Op>(A, B, acc, st, prev, cur, G, P, lowest);
      tile_step<J, Op>(B, A, acc, st, prev, cur, G, P, lowest);
    }
    if (iters - it == 2) {
      tile_step<J, Op>(A, B, acc, st, prev, cur, G, P, lowest);  // leaves the current window in B
    } else {
#pragma unroll
      for (int t = 0; t < J; ++t) B[t] = A[t];
    }
    tile_last<J, Op>(B, acc, st, prev, cur, G, lowest);
    __syncwarp();
    if (st.kstep) {  // the last job of this lane has not been flushed by a later step
      T *dst = cur + st.b * J;
#pragma unroll
      for (int t = 0; t < J; ++t) dst[t] = acc[t];
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Packed 16-bit variant of the stage: TWO cells per DPX instruction (VIADDMNMX.16x2).
//   rows   int16, one entry per offset; a 32-bit load fetches (prev[k], prev[k+1]) for even k
//   H      one 32-bit word per gap d:  lo half = G[d], hi half = G[d-1]  (any d, so no parity cases), holding the
//          stage's sentinel for d outside [0, P-2] -- H[0] = (G[0], sentinel) masks the k = j+1 half of a diagonal pair
//   acc[t] = (max over even k, max over odd k) of row t; the two halves are merged when the row block is flushed
// A J x J tile (J even) is J * J/2 instructions plus J loads of H and J/2 loads of prev.  Lane stride of H is J words:
// J = 2 * odd keeps the lanes of a team on distinct banks.  Same folding, same step order (so the same in-place
// guarantee) as dp_stage.  The caller keeps every row value and every prev + H sum inside int16 (FastPlan16).
// ------------------------------------------------------------------------------------------------
constexpr unsigned kLowest16x2 = 0x80008000u;

template <int J>
__device__ __forceinline__ void flush16(const unsigned (&acc)[J], short *cur, int b) {
  unsigned *dst = reinterpret_cast<unsigned *>(cur + b * J);  // J even and rows 4-byte aligned
#pragma unroll
  for (int t = 0; t < J; t += 2) {
    const unsigned lo = __byte_perm(acc[t], acc[t + 1], 0x5410);  // (even-k max of row t, even-k max of row t+1)
    const unsigned hi = __byte_perm(acc[t], acc[t + 1], 0x7632);  // (odd-k max of row t,  odd-k max of row t+1)
    dst[t >> 1] = __vmaxs2(lo, hi);
  }
}

template <int J>
__device__ __forceinline__ void tile_switch16(unsigned (&A)[J], unsigned (&acc)[J], TileState &st, short *cur,
                                              const unsigned *__restrict__ H) {
  if (st.kb == st.kend) {
    flush16<J>(acc, cur, st.b);
    if (st.next_b >= 0) {
      st.b = st.next_b;
      st.next_b = -1;
      st.kend = st.b + 1;
    } else {
      st.b = 0;
      st.kend = -1;
      st.kstep = 0;
    }
    st.kb = 0;
    const unsigned *ha = H + st.b * J + (J - 1);
#pragma unroll
    for (int t = 0; t < J; ++t) {
      acc[t] = kLowest16x2;
      A[t] = ha[-t];
    }
  }
}

// A[i] = H[m + J-1 - i], B[i] = H[m - 1 - i] with m = (b - kb) * J: row t and the k pair u need H[m + t - 2u].
template <int J>
__device__ __forceinline__ void tile_step16(unsigned (&A)[J], unsigned (&B)[J], unsigned (&acc)[J], TileState &st,
                                            const short *prev, short *cur, const unsigned *__restrict__ H) {
  __syncwarp();
  tile_switch16<J>(A, acc, st, cur, H);
  const unsigned *hb = H + (st.b - st.kb) * J - 1;
  const unsigned *pk = reinterpret_cast<const unsigned *>(prev + st.kb * J);
  unsigned pv[J / 2];
#pragma unroll
  for (int u = 0; u < J; ++u) B[u] = hb[-u];
#pragma unroll
  for (int u = 0; u < J / 2; ++u) pv[u] = pk[u];
#pragma unroll
  for (int u = 0; u < J / 2; ++u) {
#pragma unroll
    for (int t = 0; t < J; ++t) {
      const unsigned h = (t >= 2 * u) ? A[J - 1 - t + 2 * u] : B[2 * u - t - 1];
      acc[t] = __viaddmax_s16x2(pv[u], h, acc[t]);
    }
  }
  st.kb += st.kstep;
}

// the diagonal tile (kb == b, m = 0): only the pairs with k <= j; H[0]'s hi half masks k = j + 1
template <int J>
__device__ __forceinline__ void tile_last16(unsigned (&A)[J], unsigned (&acc)[J], TileState &st, const short *prev,
                                            short *cur, const unsigned *__restrict__ H) {
  __syncwarp();
  tile_switch16<J>(A, acc, st, cur, H);
  const unsigned *pk = reinterpret_cast<const unsigned *>(prev + st.kb * J);
  unsigned pv[J / 2];
#pragma unroll
  for (int u = 0; u < J / 2; ++u) pv[u] = pk[u];
#pragma unroll
  for (int u = 0; u < J / 2; ++u) {
#pragma unroll
    for (int t = 2 * u; t < J; ++t) acc[t] = __viaddmax_s16x2(pv[u], A[J - 1 - t + 2 * u], acc[t]);
  }
  st.kb += st.kstep;
}

template <int J, int TEAM>
__device__ __forceinline__ void dp_stage16(const short *prev, short *cur, const unsigned *__restrict__ H, int P, int tl) {
  static_assert(J % 2 == 0, "packed stage needs an even tile size");
  const int NB = (P + J - 1) / J;
  for (int r0 = ((NB - 1) / (2 * TEAM)) * (2 * TEAM); r0 >= 0; r0 -= 2 * TEAM) {
    const int nbr = min(2 * TEAM, NB - r0);
    const int half = (nbr + 1) >> 1;
    const int bA = r0 + tl, bB = r0 + nbr - 1 - tl;
    TileState st;
    if (tl < half) {
      st.b = bB;
      st.kend = bB + 1;
      st.kstep = 1;
      st.next_b = (bB > bA) ? bA : -1;
    } else {
      st.b = 0;
      st.kend = -1;
      st.kstep = 0;
      st.next_b = -1;
    }
    st.kb = 0;
    unsigned acc[J], A[J], B[J];
    {
      const unsigned *ha = H + st.b * J + (J - 1);
#pragma unroll
      for (int t = 0; t < J; ++t) {
        acc[t] = kLowest16x2;
        A[t] = ha[-t];
      }
    }
    const int iters = 2 * r0 + nbr + 1;
    int it = 0;
    for (; it + 2 < iters; it += 2) {
      tile_step16<J>(A, B, acc, st, prev, cur, H);
      tile_step16<J>(B, A, acc, st, prev, cur, H);
    }
    if (iters - it == 2) {
      tile_step16<J>(A, B, acc, st, prev, cur, H);
    } else {
#pragma unroll
      for (int t = 0; t < J; ++t) B[t] = A[t];
    }
    tile_last16<J>(B, acc, st, prev, cur, H);
    __syncwarp();
    if (st.kstep) flush16<J>(acc, cur, st.b);
  }
}

// ------------------------------------------------------------------------------------------------
// Pruned stage: the same cur[j] = max_{k<=j} (prev[k] + G[j-k]) with the tiles that cannot matter left out -- exactly.
//
// A tile (row block b, k block kb, delta = b - kb) only holds gaps in [(delta-1)J+1, (delta+1)J-1], so every cell of it
// is at most  bmax[kb] + gt[delta]  (bmax = maximum of prev over the k block, gt = maximum of G over that gap range).
// Once the rows of block b hold some values, a tile whose bound does not exceed the SMALLEST of them cannot raise any
// of the J maxima and is skipped; the rows come out bit-identical to dp_stage.  What makes this pay is a connector whose
// gap energies fall off quickly around their peak (small sigma): the tiles around the peak [dlo, dhi] are run first and
// unconditionally, which lifts the J maxima to their final level or close to it, and then
//   near side (delta < dlo): one bound test per tile;
//   far side  (delta > dhi): first ONE test for the whole rest, pmax[kb] + gtail[delta] (prefix maximum of prev, suffix
//                            maximum of gt), then -- for rows of 32 blocks or more -- one test per aligned group of four
//                            k blocks (bmax of the group + gtail), then tile by tile.
// The per-connector tables gt / gtail / dlo / dhi come from the organism's plan; bmax / pmax (and the group maxima) are
// built here, per pair and stage.  Row blocks are not folded (the work per block no longer grows with b): a round gives one block to each
// lane of the team, top blocks first, and flushes after the round -- lower rounds only read lower k blocks, so the
// stage may run in place like dp_stage.  The lanes of a warp leave the tile loop together (ballot), whatever their
// number of tiles.
// ------------------------------------------------------------------------------------------------
template <int J, class Op>
__device__ __forceinline__ void tile_any(typename Op::T (&acc)[J], const typename Op::T *prev,
                                         const typename Op::T *__restrict__ G, int b, int kb) {
  typedef typename Op::T T;
  // A[i] = G[m + J-1 - i], B[i] = G[m - 1 - i], m = (b - kb) * J; on the diagonal tile (m = 0) B is the sentinel pad
  const T *ga = G + (b - kb) * J + (J - 1);
  const T *pk = prev + kb * J;
  T A[J], B[J], pv[J];
#pragma unroll
  for (int u = 0; u < J; ++u) {
    A[u] = ga[-u];
    B[u] = ga[-J - u];
    pv[u] = pk[u];
  }
#pragma unroll
  for (int u = 0; u < J; ++u) {
#pragma unroll
    for (int t = 0; t < J; ++t) {
      const T g = (u <= t) ? A[J - 1 - t + u] : B[u - t - 1];
      acc[t] = Op::cell(acc[t], pv[u], g);
    }
  }
}

template <int J>
__device__ __forceinline__ void tile_any16(unsigned (&acc)[J], const short *prev, const unsigned *__restrict__ H, int b, int kb) {
  const unsigned *ha = H + (b - kb) * J + (J - 1);
  const unsigned *pk = reinterpret_cast<const unsigned *>(prev + kb * J);
  unsigned A[J], B[J], pv[J / 2];
#pragma unroll
  for (int u = 0; u < J; ++u) {
    A[u] = ha[-u];
    B[u] = ha[-J - u];
  }
#pragma unroll
  for (int u = 0; u < J / 2; ++u) pv[u] = pk[u];
#pragma unroll
  for (int u = 0; u < J / 2; ++u) {
#pragma unroll
    for (int t = 0; t < J; ++t) {
      const unsigned h = (t >= 2 * u) ? A[J - 1 - t + 2 * u] : B[2 * u - t - 1];
      acc[t] = __viaddmax_s16x2(pv[u], h, acc[t]);
    }
  }
}

struct PruneTabs {
  const int *gt;     // [NB] bound of G over the gap range of tile offset delta
  const int *gtail;  // [NB] max of gt[delta..]
  int dlo, dhi;      // tile offsets run unconditionally (around the peak of G)
};

// The next k block at or below kb (walking down) whose tile can still raise one of the maxima of row block b, given
// the smallest of them so far; -1 when none is left.  Tiles dlo..dhi have been run already.  On the far side an aligned
// group of four k blocks is skipped with one test (bm4 = maximum of the row over the four blocks; gtail[d] bounds the gap
// energies of every tile at or beyond d): a single good predecessor far to the left keeps the prefix maximum high for
// every block above it, and walking down to it block by block was 20 % of the pruned kernel's time.
__device__ __forceinline__ int prune_next_tile(int b, int kb, int minacc, const PruneTabs &tabs, const int *bm, const int *pm,
                                               const int *bm4) {
  while (kb >= 0) {
    const int d = b - kb;
    if (d >= tabs.dlo && d <= tabs.dhi) {
      kb = b - tabs.dhi - 1;
      continue;
    }
    if (d > tabs.dhi) {
      const int tail = tabs.gtail[d];
      if (pm[kb] + tail <= minacc) return -1;  // nothing on the far side can matter any more
      if (bm4 != nullptr && (kb & 3) == 3 && bm4[kb >> 2] + tail <= minacc) {
        kb -= 4;
        continue;
      }
    }
    // two tiles per trip (their four loads in flight together: the walk is a chain of load -> add -> compare -> branch)
    const int kb2 = kb > 0 ? kb - 1 : 0, d2 = b - kb2;
    const int v1 = bm[kb] + tabs.gt[d], v2 = bm[kb2] + tabs.gt[d2];
    if (v1 > minacc) break;
    --kb;
    if (kb >= 0 && !(d2 >= tabs.dlo && d2 <= tabs.dhi)) {
      if (v2 > minacc) break;
      --kb;
    }
  }
  return kb;
}

// smallest maximum among the rows of block b that exist (j < P)
template <int J>
__device__ __forceinline__ int prune_row_min(const int (&acc)[J], int b, int P) {
  int mn = 0x7fffffff;
#pragma unroll
  for (int t = 0; t < J; ++t) mn = min(mn, (b * J + t < P) ? acc[t] : 0x7fffffff);
  return mn;
}
template <int J>
__device__ __forceinline__ int prune_row_min(const unsigned (&acc)[J], int b, int P) {
  int mn = 0x7fffffff;
#pragma unroll
  for (int t = 0; t < J; ++t) {
    const int lo = (int)(short)(acc[t] & 0xffffu), hi = (int)(short)(acc[t] >> 16);
    mn = min(mn, (b * J + t < P) ? (lo > hi ? lo : hi) : 0x7fffffff);
  }
  return mn;
}
template <int J>
__device__ __forceinline__ void prune_run_tile(int (&acc)[J], const int *prev, const int *__restrict__ G, int b, int kb) {
  tile_any<J, OpI32>(acc, prev, G, b, kb);
}
template <int J>
__device__ __forceinline__ void prune_run_tile(unsigned (&acc)[J], const short *prev, const int *__restrict__ G, int b, int kb) {
  tile_any16<J>(acc, prev, reinterpret_cast<const unsigned *>(G), b, kb);
}
template <int J>
__device__ __forceinline__ void prune_flush(const int (&acc)[J], int *cur, int b) {
  int *dst = cur + b * J;
#pragma unroll
  for (int t = 0; t < J; ++t) dst[t] = acc[t];
}
template <int J>
__device__ __forceinline__ void prune_flush(const unsigned (&acc)[J], short *cur, int b) {
  flush16<J>(acc, cur, b);
}

#ifdef __CUDACC__
// RowT = int (G = int table) or short (G = packed words H).  bm / pm: NB ints each, per pair.
// Returns tiles this lane ran + 65536 * tile steps of the warp (statistics of the profiling build; dead code otherwise).
template <int J, int TEAM, class RowT>
__device__ __forceinline__ int dp_stage_pruned(const RowT *prev, RowT *cur, const int *__restrict__ G, const PruneTabs tabs,
                                               int *bm, int *pm, int *bm4, int P, int tl) {
  constexpr bool k16 = sizeof(RowT) == 2;
  constexpr int kLow = -0x7fffffff - 1;
  typedef typename std::conditional<k16, unsigned, int>::type AccT;
  const int NB = (P + J - 1) / J;
  int tiles_run = 0;
  // block maxima and their prefix maxima (entries k >= P of the last block are padding, not values)
  for (int r0 = 0, carry = kLow; r0 < NB; r0 += TEAM) {
    const int kb = r0 + tl;
    int v = kLow;
    if (kb < NB) {
      const RowT *pk = prev + kb * J;
#pragma unroll
      for (int u = 0; u < J; ++u) v = max(v, (kb * J + u < P) ? (int)pk[u] : kLow);
      bm[kb] = v;
    }
#pragma unroll
    for (int o = 1; o < TEAM; o <<= 1) {
      const int w = __shfl_up_sync(0xffffffffu, v, o, TEAM);
      if (tl >= o) v = max(v, w);
    }
    v = max(v, carry);
    if (kb < NB) pm[kb] = v;
    carry = __shfl_sync(0xffffffffu, v, TEAM - 1, TEAM);
  }
  __syncwarp();
  // maxima over aligned groups of four blocks: pays when there are many blocks (2 kb promoters: config 4 259 -> 251 ms),
  // costs 2-3 % at 300 bp
  const bool groups = NB >= 32;
  if (groups) {
    for (int q = tl; 4 * q < NB; q += TEAM) {
      int v = bm[4 * q];
#pragma unroll
      for (int e = 1; e < 4; ++e) v = max(v, (4 * q + e < NB) ? bm[4 * q + e] : kLow);
      bm4[q] = v;
    }
    __syncwarp();
  }
  for (int hi = NB - 1; hi >= 0; hi -= TEAM) {
    const int b = hi - tl;   // < 0: idle lane of the last round
    AccT acc[J];
#pragma unroll
    for (int t = 0; t < J; ++t) acc[t] = k16 ? (AccT)kLowest16x2 : (AccT)kLow;
    // the tiles around the peak
    for (int d = tabs.dlo; d <= tabs.dhi; ++d) {
      tiles_run += 65536;
      if (b >= 0 && d <= b) {
        prune_run_tile<J>(acc, prev, G, b, b - d);
        ++tiles_run;
      }
    }
    int minacc = (b >= 0 && tabs.dlo <= b) ? prune_row_min<J>(acc, b, P) : kLow;
    // everything else, by bound
    int kb = b;
    for (;;) {
      kb = prune_next_tile(b, kb, minacc, tabs, bm, pm, groups ? bm4 : nullptr);
      if (!__any_sync(0xffffffffu, kb >= 0)) break;
      tiles_run += 65536;
      if (kb >= 0) {
        prune_run_tile<J>(acc, prev, G, b, kb);
        ++tiles_run;
        minacc = prune_row_min<J>(acc, b, P);
        --kb;
      }
    }
    __syncwarp();
    if (b >= 0) prune_flush<J>(acc, cur, b);
    __syncwarp();
  }
  return tiles_run;
}
#endif

// ------------------------------------------------------------------------------------------------
// Shared pieces of the placement kernels
// ------------------------------------------------------------------------------------------------
struct OrgView {
  int o, n, sumL, P, eff;
};

// Organism prologue (all threads of the CTA): recognizer tables and descriptors into shared memory and the
// exact gap-energy tables G[c][g] (connector.c:35-66 hoisted; -inf outside [0, P)).  Caller syncs afterwards.
struct ConView {
  const double *pdf;  // pdf block of the connector (gap 0 first)
  double auc;         // cdf[S - sumL] through the reference's shifted cdf pointer
};
__device__ __forceinline__ ConView connector_view(const PlaceParams &p, const OrgView &v, int c) {
  const int M = p.con_M[v.o];
  ConView cv;
  cv.pdf = p.con_pool + p.con_begin[v.o] + (long long)c * (2 * M + 2) + 2;
  cv.auc = __ldg(cv.pdf + (M - 1) + (p.S - v.sumL));  // cdf pointer = pdf + max_len, max_len = M-1 (connector.c:3-4)
  return cv;
}
// exact gap energy of connector view cv for gap g (what the G tables hold)
__device__ __forceinline__ double exact_gap(const PlaceParams &p, const OrgView &v, const ConView &cv, int g) {
  return gap_energy(cv.pdf, cv.auc, g, v.n, v.eff, p.tables + T_DEN);
}

__device__ __forceinline__ OrgView organism_prologue(const PlaceParams &p, int o, double *s_pool, double *s_G, RecDesc *s_rec) {
  OrgView v;
  const int tid = threadIdx.x;
  const int rb = p.rec_begin[o];
  v.o = o;
  v.n = p.rec_begin[o + 1] - rb;
  v.sumL = p.sum_len[o];
  v.P = p.S - v.sumL + 1;
  v.eff = p.S - v.sumL + v.n;
  const long long pb = p.pool_begin[o];
  const int pn = (int)(p.pool_begin[o + 1] - pb);
  for (int i = tid; i < pn; i += blockDim.x) s_pool[i] = __ldg(p.rec_pool + pb + i);
  if (tid == 0) {
    int fwd = 0;
    for (int i = 0; i < v.n; ++i) {
      RecDesc d;
      const char t = p.rec_type[rb + i];
      d.kind = (t == 'p' || t == 'P') ? 0 : (t == 'm' || t == 'M') ? 1 : (t == 't' || t == 'T') ? 2 : (t == 'r' || t == 'R') ? 3 : 4;
      d.len = p.rec_len[rb + i];
      d.nb = p.rec_nb[rb + i];
      d.data = (int)(p.rec_data[rb + i] - pb);
      d.fwd = fwd;
      d.cls = p.rec_class[rb + i];
      fwd += d.len;
      s_rec[i] = d;
    }
  }
  if (s_G != nullptr) {
    for (int c = 0; c < v.n - 1; ++c) {
      const ConView cv = connector_view(p, v, c);
      double *Gc = s_G + (size_t)c * p.g_stride;
      for (int i = tid; i < p.g_stride; i += blockDim.x) {
        const int g = i - kGPad;
        Gc[i] = (g < 0 || g >= v.P) ? neg_inf() : exact_gap(p, v, cv, g);
      }
    }
  }
  return v;
}

// 2-bit bases (+ unknown mask) of one sequence into byte codes (0..3, 4 = unknown), by one warp.
__device__ __forceinline__ void unpack_codes(const PlaceParams &p, int seq, unsigned char *codes, int lane) {
  const unsigned *pw = p.packed + p.word_off[seq];
  const unsigned *uw = p.unk + p.unk_off[seq];
  for (int b = lane; b < p.S; b += 32) {
    const unsigned w = __ldg(pw + (b >> 4));
    const unsigned u = __ldg(uw + (b >> 5));
    const unsigned code = (w >> (2 * (b & 15))) & 3u;
    codes[b] = ((u >> (b & 31)) & 1u) ? 4 : (unsigned char)code;
  }
}

// One (organism, sequence) pair by one warp, exact fp64: score rows, chain DP, first maximum, optional traceback.
template <int J, bool TRACE>
__device__ __forceinline__ void place_pair_exact(const PlaceParams &p, const OrgView &v, int seq, const double *s_pool,
                                                 const double *s_G, const RecDesc *s_rec, double *rows,
                                                 unsigned char *codes, int lane, int &err) {
  const int n = v.n, P = v.P;
  unpack_codes(p, seq, codes, lane);
  __syncwarp();

  // ---- stage 0: C_0 = R_0 (organism.c:381-385) --------------------------------------------------
  const unsigned char *bins_seq = p.bins + (size_t)seq * p.bins_row_stride;
  const size_t cls_stride = (size_t)p.n_seq_total * p.bins_row_stride;
  for (int j = lane; j < P; j += 32) rows[j] = score_rec(s_rec[0], s_pool, codes, j, bins_seq, cls_stride, err);
  __syncwarp();

  // ---- stages 1..n-1: C_i[j] = max_k (C_{i-1}[k] + G[j-k]) + R_i[j] (organism.c:347-369) ----------
  for (int i = 1; i < n; ++i) {
    const double *prev = rows + (size_t)(TRACE ? (i - 1) : ((i - 1) & 1)) * p.row_stride;
    double *cur = rows + (size_t)(TRACE ? i : (i & 1)) * p.row_stride;
    const double *Gc = s_G + (size_t)(i - 1) * p.g_stride + kGPad;
    dp_stage<J, 32, OpF64>(prev, cur, Gc, P, lane, neg_inf());
    __syncwarp();
    const RecDesc rd = s_rec[i];
    for (int j = lane; j < P; j += 32) cur[j] = cur[j] + score_rec(rd, s_pool, codes, j, bins_seq, cls_stride, err);
    __syncwarp();
  }

  // ---- first maximum of the last row (_aux.c:149-158) --------------------------------------------
  const double *lastrow = rows + (size_t)(TRACE ? (n - 1) : ((n - 1) & 1)) * p.row_stride;
  double bv = neg_inf();
  int bi = 0x7fffffff;
  for (int j = lane; j < P; j += 32) {
    const double x = lastrow[j];
    if (bi == 0x7fffffff || x > bv) { bv = x; bi = j; }
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    const double ov = __shfl_xor_sync(0xffffffffu, bv, off);
    const int ob = __shfl_xor_sync(0xffffffffu, bi, off);
    if (ob != 0x7fffffff && (bi == 0x7fffffff || ov > bv || (ov == bv && ob < bi))) { bv = ov; bi = ob; }
  }
  const size_t pair = (size_t)v.o * p.n_seq_total + seq;
  if (lane == 0) p.E[pair] = bv;

  // ---- traceback (organism.c:398-427): re-scan one column per stage for the first k that reproduces
  //      the stored value exactly -- identical to the reference's strict '<' update rule ---------------
  if (TRACE) {
    int m = bi;
    for (int i = n - 1; i >= 1; --i) {
      const double *prev = rows + (size_t)(i - 1) * p.row_stride;
      const double *Gc = s_G + (size_t)(i - 1) * p.g_stride + kGPad;
      const RecDesc rd = s_rec[i];
      const double ri = score_rec(rd, s_pool, codes, m, bins_seq, cls_stride, err);
      const double target = rows[(size_t)i * p.row_stride + m];
      int found = -1;
      for (int kb = 0; kb <= m && found < 0; kb += 32) {
        const int k = kb + lane;
        // target == -inf (e.g. a -inf PSSM entry on the path): the reference's strict '<' never fires from its -inf
        // start value, so nothing "won" and its calloc'ed zeros stay (organism.c:365)
        const bool hit = (k <= m) && (target > neg_inf()) && ((prev[k] + Gc[m - k]) + ri == target);
        const unsigned bal = __ballot_sync(0xffffffffu, hit);
        if (bal) found = kb + __ffs(bal) - 1;
      }
      const int gap = found < 0 ? 0 : m - found;  // nothing ever won: the reference keeps its calloc'ed zeros
      if (lane == 0) {
        p.rsc[pair * p.max_rec + i] = ri;
        p.csc[pair * p.max_rec + i - 1] = found < 0 ? 0.0 : Gc[gap];
        p.gaps[pair * p.max_rec + i] = gap;
      }
      m -= gap;
    }
    if (lane == 0) {
      p.rsc[pair * p.max_rec] = rows[m];
      p.gaps[pair * p.max_rec] = m;
    }
  }
  __syncwarp();
}

// Shared-memory carve-up of the exact kernels.
struct ExactSmem {
  double *pool, *G, *rows;
  RecDesc *rec;
  unsigned char *codes;
};
__device__ __forceinline__ ExactSmem carve_exact(const PlaceParams &p, unsigned char *raw, int nwarps) {
  ExactSmem s;
  s.pool = reinterpret_cast<double *>(raw);
  s.G = s.pool + p.pool_stride;
  s.rows = s.G + (size_t)(p.max_rec - 1) * p.g_stride;
  s.rec = reinterpret_cast<RecDesc *>(s.rows + (size_t)nwarps * p.nrows * p.row_stride);
  s.codes = reinterpret_cast<unsigned char *>(s.rec + p.max_rec);
  return s;
}

// ------------------------------------------------------------------------------------------------
// The exact fused kernel: CTA = one organism x a chunk of sequences; J = tile size chosen by the host from P.
// ------------------------------------------------------------------------------------------------
template <int J, bool TRACE>
__global__ void __launch_bounds__(256) place_tiled_kernel(const PlaceParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int oi = blockIdx.x / p.chunks;
  const int chunk = blockIdx.x - oi * p.chunks;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const ExactSmem sm = carve_exact(p, smem_raw, nwarps);
  const OrgView v = organism_prologue(p, p.org_list[oi], sm.pool, sm.G, sm.rec);
  __syncthreads();
  double *rows = sm.rows + (size_t)warp * p.nrows * p.row_stride;
  unsigned char *codes = sm.codes + (size_t)warp * p.code_stride;
  int err = 0;
  const int first = chunk * p.seqs_per_cta;
  const int last = min(p.class_count, first + p.seqs_per_cta);
  for (int si = first + warp; si < last; si += nwarps)
    place_pair_exact<J, TRACE>(p, v, p.order[p.class_begin + si], sm.pool, sm.G, sm.rec, rows, codes, lane, err);
  if (err) atomicOr(p.err_flag, err);
}

// ------------------------------------------------------------------------------------------------
// List mode: the pairs the integer fast path declined (too many near-ties, unknown bytes, tables that do
// not fit its fixed-point range).  The fast kernel files them per work item (organism x chunk of sequences).
// list_units_kernel turns the per-item counts into work UNITS of up to kListUnitPairs pairs (exclusive prefix sum over the
// items); place_list_kernel gives every CTA a CONTIGUOUS range of units, so the work is balanced whatever the
// distribution of the hand-backs (a few organisms that tie on every sequence, or a pair here and there), and a CTA
// re-runs the organism prologue only when the organism changes between consecutive units.
// ------------------------------------------------------------------------------------------------
constexpr int kMaxListBuckets = 24;
constexpr int kListUnitPairs = 8;
struct ListBuckets {
  int n;                                  // fast-path launches (tile buckets) of this length class
  int item_begin[kMaxListBuckets + 1];    // first global item of bucket b; [n] = total
  int chunks[kMaxListBuckets];            // chunks per organism
  int spc[kMaxListBuckets];               // sequences per chunk = slots per item in fb_item_seqs
  int org_begin[kMaxListBuckets];         // first entry of the bucket in p.org_list
  long long pair_base[kMaxListBuckets];   // first slot of the bucket in fb_item_seqs
};

// unit_begin[g] = number of units of the items before g, unit_begin[total_items] = all units; one CTA.
__global__ void __launch_bounds__(1024) list_units_kernel(const int *__restrict__ fb_item_count, int total_items, int *__restrict__ unit_begin,
                                                          unsigned long long *fb_total) {
  __shared__ int s_part[1024];
  __shared__ long long s_pairs[32];
  const int tid = threadIdx.x, nthr = blockDim.x;
  const int per = (total_items + nthr - 1) / nthr;
  const int lo = min(total_items, tid * per), hi = min(total_items, lo + per);
  int units = 0;
  long long pairs = 0;
  for (int g = lo; g < hi; ++g) {
    const int c = fb_item_count[g];
    units += (c + kListUnitPairs - 1) / kListUnitPairs;
    pairs += c;
  }
  s_part[tid] = units;
  __syncthreads();
  for (int off = 1; off < nthr; off <<= 1) {  // inclusive scan of the per-thread totals
    const int add = (tid >= off) ? s_part[tid - off] : 0;
    __syncthreads();
    s_part[tid] += add;
    __syncthreads();
  }
  int run = s_part[tid] - units;
  for (int g = lo; g < hi; ++g) {
    unit_begin[g] = run;
    run += (fb_item_count[g] + kListUnitPairs - 1) / kListUnitPairs;
  }
  if (tid == nthr - 1) unit_begin[total_items] = s_part[tid];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) pairs += __shfl_xor_sync(0xffffffffu, pairs, o);
  if ((tid & 31) == 0) s_pairs[tid >> 5] = pairs;
  __syncthreads();
  if (tid == 0) {
    long long all = 0;
    for (int w = 0; w < (nthr + 31) / 32; ++w) all += s_pairs[w];
    atomicAdd(fb_total, (unsigned long long)all);
  }
}

template <int J, bool TRACE>
__global__ void __launch_bounds__(128) place_list_kernel(const PlaceParams p, const ListBuckets lb, const int *__restrict__ fb_item_count,
                                                         const int *__restrict__ fb_item_seqs, const int *__restrict__ unit_begin) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const ExactSmem sm = carve_exact(p, smem_raw, nwarps);
  double *rows = sm.rows + (size_t)warp * p.nrows * p.row_stride;
  unsigned char *codes = sm.codes + (size_t)warp * p.code_stride;
  const int total_items = lb.item_begin[lb.n];
  const int total_units = __ldg(unit_begin + total_items);
  // this CTA's contiguous share of the units
  const int u_lo = (int)((long long)total_units * blockIdx.x / gridDim.x);
  const int u_hi = (int)((long long)total_units * (blockIdx.x + 1) / gridDim.x);
  int err = 0, cur_org = -1;
  OrgView v;
  v.o = -1;
  // item of the first unit: the last g with unit_begin[g] <= u_lo (binary search; later units walk forward)
  int g = 0;
  if (u_lo < u_hi) {
    int a = 0, b = total_items;  // invariant: unit_begin[a] <= u_lo < unit_begin[b]
    while (b - a > 1) {
      const int mid = (a + b) >> 1;
      if (__ldg(unit_begin + mid) <= u_lo) a = mid; else b = mid;
    }
    g = a;
  }
  for (int u = u_lo; u < u_hi; ++u) {
    while (__ldg(unit_begin + g + 1) <= u) ++g;  // items without hand-backs have no units
    const int part = u - __ldg(unit_begin + g);
    const int cnt = __ldg(fb_item_count + g);
    int b = 0;
    while (b + 1 < lb.n && lb.item_begin[b + 1] <= g) ++b;
    const int item = g - lb.item_begin[b];
    const int o = p.org_list[lb.org_begin[b] + item / lb.chunks[b]];
    const int *seqs = fb_item_seqs + lb.pair_base[b] + (long long)item * lb.spc[b];
    if (o != cur_org) {   // uniform over the CTA
      __syncthreads();    // the previous organism's tables are dead
      v = organism_prologue(p, o, sm.pool, sm.G, sm.rec);
      __syncthreads();
      cur_org = o;
    }
    const int first = part * kListUnitPairs, last = min(cnt, first + kListUnitPairs);
    for (int i = first + warp; i < last; i += nwarps)
      place_pair_exact<J, TRACE>(p, v, __ldg(seqs + i), sm.pool, sm.G, sm.rec, rows, codes, lane, err);
  }
  if (err) atomicOr(p.err_flag, err);
}
