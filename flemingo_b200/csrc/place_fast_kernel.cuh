// place_fast_kernel.cuh -- filter-and-refine placement: an integer fixed-point DPX pass (int32, or packed 16-bit:
// two DP cells per instruction) finds, with a rigorous error bound, every DP cell that can lie on the optimal path;
// those few cells are then re-evaluated in exact IEEE double in the reference's operand order.  Results are
// bit-identical to the reference; pairs for which the candidate sets grow too large (massive ties) or that the
// fixed-point range cannot hold are filed in device-side lists -- from the packed pass for a second chance on the int32
// pass, from the int32 pass for the exact kernel (place_list_kernel).
//
// Why it is exact.  Let C be the reference's doubles (organism.c:347-369), Chat the same sums in real
// arithmetic (|C - Chat| < 2^-30 units after scaling, see `slack`), and Cq the integer pass with every table
// entry rounded to the nearest integer in units of 1/s (s a power of two, so scaling itself is exact; the packed pass
// may also use 100 * 2^e, under which 2-decimal PSSM entries are integers up to ~1e-13 units, covered by the slack):
//     |Cq_i[j] - s*Chat_i[j]| <= err_i = 0.5 * (number of rounded table entries on a path up to stage i).
// Therefore every j that maximises C_{n-1} has Cq_{n-1}[j] >= max Cq_{n-1} - 2 err_{n-1} - slack, and every k
// that can win cell (i, j) has Cq_{i-1}[k] + Gq[j-k] >= (Cq_i[j] - Rq_i[j]) - 2 (err_{i-1} + 0.5) - slack.
// Walking these inequalities down from the last row yields per-stage candidate sets F_i; evaluating
//     C_i[j] = max_{k in F_{i-1}, k <= j} fl(fl(C_{i-1}[k] + G[j-k]) + R_i[j])   (ascending k, strict <)
// on them reproduces the reference values AND its first-k / first-j tie rule.  The sentinel that stands for the
// reference's -1e10 (gap P-1) and for out-of-triangle cells is chosen below (min path - max path - threshold),
// so a path through it can never enter a candidate set nor overflow int32.
//
// PSSM score rows in the integer pass use 4-mer lookup tables (256 entries per group of 4 columns, rounded ONCE
// per entry), so a row costs ceil(L/4) lookups per offset instead of L adds; integer addition is associative,
// which makes this legal here and illegal in the exact pass.
#pragma once

// Two arithmetics share this file (template parameter BITS):
//   32  int32 rows, one VIADDMNMX per DP cell; scale ~2^20 units per energy unit, thresholds of ~1e-5 energy units:
//       the candidate sets hold exact ties only.  Fits every organism whose tables are finite.
//   16  int16 rows, VIADDMNMX.S16x2 = TWO DP cells per instruction (dp_stage16).  The whole value range of a path has to
//       fit 16 bits, so the scale is 16..128 units per energy unit and the thresholds are 0.05..1 energy units: the
//       candidate sets hold a few near-optimal cells per stage, which the same exact refinement sorts out.  To make
//       room the sentinel is chosen per stage, and when every PSSM entry is a multiple of 0.01 -- what
//       recalculate_pssm produces -- the scale 100 * 2^e makes the recognizer tables EXACT (no rounding error, tighter
//       thresholds).  The host routes an organism here when its estimated scale is at least kMinScale16.
// In both arithmetics every quantised table is shifted so that its maximum is 0 (a constant per stage changes no arg-max
// and no threshold): every row value is then <= 0, which is what makes the zero padding behind a row and the sentinel
// (chosen relative to the lowest legit gap energy) safe whatever the absolute level of the scores -- recognizers whose
// scores are all strongly negative (MLE-fitted shape recognizers with sigma = 0: ln(1e-100) in every bin but one) put
// whole rows far below any gap energy.
//
// Candidate cells per stage: 16 shared by the pairs of a warp (4 per pair for TEAM = 8).  Eight per pair for the packed
// pass were measured and rejected: the longer lists cost more in every refinement than the second chance they save.
template <int BITS, int NT> struct CandCap { static constexpr int kPerPair = 16 / NT; };
constexpr int cand_per_warp(int bits, int nt) { return 16; }
constexpr int kNanBinMarker = (int)0x80000000;  // quantised score of a bin whose null frequency is NaN
constexpr int kSlackUnits = 2; // covers fp64-vs-real rounding of the reference sums (<< 1 unit) with margin
constexpr int kSlackUnits16 = 1;
constexpr double kMinScale16 = 24.0;   // below this the 16-bit thresholds let too many cells through: use int32
constexpr int kBudget16 = 32000;       // usable magnitude of an int16 value (a little below 32767)

template <int BITS> struct RowOf { typedef int T; };
template <> struct RowOf<16> { typedef short T; };

// quantised gap energy at gap d: a plain int table (BITS = 32) or the low half of the packed word H[d] (BITS = 16)
template <int BITS>
__device__ __forceinline__ int gq_at(const int *__restrict__ tab, int d) {
  if (BITS == 32) return tab[d];
  return (int)reinterpret_cast<const short *>(tab)[2 * d];
}

struct FastParams {
  PlaceParams base;
  const unsigned char *has_unk;  // [n_seq] sequence holds a non-ACGT byte -> exact kernel
  // hand-back lists, one per work item (organism x chunk of sequences), so that the exact list kernel meets the pairs
  // already grouped by organism: item g = item_base + item owns fb_item_seqs[pair_base + item * seqs_per_cta ...]
  int *fb_item_count;
  int *fb_item_seqs;
  int item_base;
  long long pair_base;
  // second-chance launch (int32 pass over what the packed 16-bit pass handed back): the sequences of item i are
  // retry_seqs[i * seqs_per_cta ...], retry_count[i] of them; null for a first launch
  const int *retry_count;
  const int *retry_seqs;
  unsigned long long *retried_total;  // statistics: pairs that took the second chance
  int gq_stride;     // ints per quantised connector table (kGPad + Pmax + kGPad)
  int rowq_stride;   // row ELEMENTS (int32 or int16) per quantised row
  int slotq_stride;  // row elements per pair slot (1 or 2 rows); in 32-bit words == TEAM (mod 32): teams hit disjoint banks
  int single_row;    // 1: ONE quantised row per pair in shared memory -- the DP stage runs in place (dp_stage) and every
                     // older row is streamed back from global scratch when the refinement needs it; chosen by the host
                     // when halving the row storage lets more warps (a higher multiple of four) live on an SM
  int cand_w;        // candidate cells per stage and WARP in the lists (cand_per_warp of the launched kernel)
  int k8_stride;     // bytes per 4-mer code array (S + kK8Pad rounded up to 16)
  int kmer_cap;      // ints available for the 4-mer tables of one organism
  int want_trace;
  int *scratch;      // global rows 0..n-3 of every resident pair slot (only the last two rows stay in shared memory)
  long long scratch_slot_stride;  // row elements per pair slot in `scratch` (max_rec * rowq_stride)
  int *work_counter;  // device counter the persistent CTAs claim work items from (zeroed before the launch)
  int n_items;        // organisms of the bucket x chunks
  // pruned stages (kernels instantiated with PRUNE, see dp_stage_pruned): per connector, behind its gap table, the tile
  // bounds gt[nb_max] and their suffix maxima gtail[nb_max]; per pair slot, block maxima and prefix maxima of the row
  int gt_off;         // offset of gt inside a connector's table (ints); gq_stride covers gt_off + 2 * nb_max
  int nb_max;         // row blocks of the largest organism of the launch: ceil(Pmax / J)
  int bm_stride;      // ints per pair slot for bmax + pmax + bmax of groups of four (>= 2 * nb_max + nb_max / 4 + 1; 0 in a launch without pruning)
  int prune_drop;     // a tile belongs to the peak while its bound is within this many energy units of the maximum
  int prune_frac;     // a stage is pruned when (peak tiles + 2) * prune_frac <= row blocks
};

struct FastPlan {
  double scale;                          // s: quantised = rint(value * s)
  int sent;                              // sentinel ("minus infinity" of the integer pass)
  int ok;                                // 0: this organism cannot use the integer pass
  int item;                              // work item the CTA is processing (claimed by thread 0)
  int next_round;                        // next round (NT sequences) of the item; the warps claim rounds as they get free
#ifdef FL_PHASE_CLOCKS
  int trace_base;                        // profiling build: first timeline row of this CTA, or -1
#endif
  int sent_stage[FL_MAX_RECOGNIZERS];    // BITS = 16: sentinel of connector c (per stage, as small in magnitude as the bound allows)
  double rrange[FL_MAX_RECOGNIZERS];     // BITS = 16: value range of recognizer r / of connector c in energy units (scale choice)
  double grange[FL_MAX_RECOGNIZERS];
  int thr_final;                         // slack on the last row
  int thr_pred[FL_MAX_RECOGNIZERS];      // slack on predecessor scans of stage i
  int ngroups[FL_MAX_RECOGNIZERS];       // 4-mer groups of PSSM recognizer i (0 for a shape recognizer)
  int kmer_off[FL_MAX_RECOGNIZERS];      // offset of its tables in s_kmer (4-mer tables, or nb-1 quantised bin scores)
  int err_half[FL_MAX_RECOGNIZERS];      // rounded table entries on a path through recognizer i (half units of error)
  int unit_begin[FL_MAX_RECOGNIZERS];    // first quantisation work unit (table) of recognizer i
  int units;                             // tables of all recognizers (the connectors follow)
  int prune[FL_MAX_RECOGNIZERS];         // PRUNE kernels: stage of connector c runs dp_stage_pruned, with the peak tiles
  int dlo[FL_MAX_RECOGNIZERS];           //   dlo[c] .. dhi[c]
  int dhi[FL_MAX_RECOGNIZERS];
};

template <class RowT>
struct FastSmemT {
  double *pool, *G, *cv, *cr;
  int *Gq, *kmer, *cj, *ca, *cn, *bm;
  RowT *rowsq;
  unsigned char *k8;
  RecDesc *rec;
  FastPlan *plan;
};

// strides of FastParams that concern rows (rowq_stride, slotq_stride, scratch_slot_stride) count ROW ELEMENTS
template <class RowT>
__device__ __forceinline__ FastSmemT<RowT> carve_fast(const FastParams &f, unsigned char *raw, int nwarps, int nt) {
  const PlaceParams &p = f.base;
  FastSmemT<RowT> s;
  s.pool = reinterpret_cast<double *>(raw);
  s.G = s.pool + p.pool_stride;                       // exact gap energies [max_rec-1][g_stride], -inf outside [0, P)
  s.cv = s.G + (size_t)(p.max_rec - 1) * p.g_stride;
  s.cr = s.cv + (size_t)nwarps * p.max_rec * f.cand_w;
  s.plan = reinterpret_cast<FastPlan *>(s.cr + (size_t)nwarps * p.max_rec * f.cand_w);
  s.rec = reinterpret_cast<RecDesc *>(s.plan + 1);
  s.Gq = reinterpret_cast<int *>(s.rec + p.max_rec);
  s.kmer = s.Gq + (size_t)(p.max_rec - 1) * f.gq_stride;
  // 16-byte aligned rows (cp.async / int4 targets).  Pointer + offset, not an integer round trip: the compiler must keep
  // seeing a shared-memory pointer, or every row access of the DP loop turns into a generic load.
  int *rows_unaligned = s.kmer + f.kmer_cap;
  int *rows_aligned = rows_unaligned + (((16u - ((unsigned)__cvta_generic_to_shared(rows_unaligned) & 15u)) & 15u) >> 2);
  s.rowsq = reinterpret_cast<RowT *>(rows_aligned);
  s.cj = reinterpret_cast<int *>(s.rowsq + (size_t)nwarps * nt * f.slotq_stride);  // slot strides are even: 4-byte aligned
  s.ca = s.cj + (size_t)nwarps * p.max_rec * f.cand_w;
  s.cn = s.ca + (size_t)nwarps * p.max_rec * f.cand_w;
  s.bm = s.cn + (size_t)nwarps * nt * p.max_rec;
  unsigned char *k8_unaligned = reinterpret_cast<unsigned char *>(s.bm + (size_t)nwarps * nt * f.bm_stride);
  s.k8 = k8_unaligned + ((16u - ((unsigned)__cvta_generic_to_shared(k8_unaligned) & 15u)) & 15u);
  return s;
}

// ---- warp reductions -------------------------------------------------------------------------------
__device__ __forceinline__ double warp_min(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ int warp_min_i(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = min(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ int warp_max_i(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// quantised PSSM score at `pos` (already including the recognizer's forward offset): one 4-mer lookup per group.
// k8[q] = code of the 4-mer starting at q, built once per pair (the same codes serve every recognizer and group).
__device__ __forceinline__ int score_q(const int *__restrict__ tab, int ngroups, const unsigned char *__restrict__ k8, int pos) {
  int v = 0;
  const unsigned char *k = k8 + pos;
  for (int g = 0; g < ngroups; ++g) v += tab[256 * g + k[4 * g]];
  return v;
}

// exact PSSM score (recognizer.c:130-168); the base at position q is the low 2 bits of its 4-mer code, and the fast
// path only sees ACGT sequences
__device__ __forceinline__ double score_exact_k8(const RecDesc &r, const double *s_pool, const unsigned char *k8, int pos) {
  const double *m = s_pool + r.data;
  double s = 0.0;
  for (int col = 0; col < r.len; ++col) s += m[4 * col + (k8[pos + col] & 3)];
  return s;
}

// quantised score of recognizer r at relative offset j: 4-mer tables for a PSSM, bin table for a shape recognizer
__device__ __forceinline__ int score_q_any(const FastPlan &plan, int r, const RecDesc &rd, const int *__restrict__ kmer,
                                           const unsigned char *__restrict__ k8, const unsigned char *__restrict__ bins_seq,
                                           size_t cls_stride, int j) {
  if (rd.kind == 0) return score_q(kmer + plan.kmer_off[r], plan.ngroups[r], k8, rd.fwd + j);
  return kmer[plan.kmer_off[r] + bins_seq[(size_t)rd.cls * cls_stride + rd.fwd + j]];
}
// exact score of recognizer rd at relative offset j
__device__ __forceinline__ double score_exact_any(const RecDesc &rd, const double *s_pool, const unsigned char *k8,
                                                  const unsigned char *__restrict__ bins_seq, size_t cls_stride, int j) {
  if (rd.kind == 0) return score_exact_k8(rd, s_pool, k8, rd.fwd + j);
  int dummy = 0;  // a NaN null model keeps the organism off this path (build_fast_plan), so nothing to flag here
  return score_shape_bin(rd, s_pool, bins_seq[(size_t)rd.cls * cls_stride + rd.fwd + j], dummy);
}

// ---- quantisation plan of one organism (all threads of the CTA, after the exact tables are in shared memory) ----
// Step A (warp 0) finds the value ranges and the scale; step B builds the quantised tables, one warp per table
// (a 4-mer group of a PSSM, the bin table of a shape recognizer, a connector); step C (warp 0) adds up the integer
// ranges, places the sentinel(s) and the thresholds; step D fills the sentinel pads (BITS = 16: and the high halves of
// the packed gap words).  Contains block barriers.
template <int BITS, class RowT, int J, bool PRUNE>
__device__ void build_fast_plan(const FastParams &f, const OrgView &v, const FastSmemT<RowT> &sm, int warp, int lane, int nwarps) {
  const PlaceParams &p = f.base;
  FastPlan *plan = sm.plan;
  const int n = v.n, P = v.P;
  const double inf = __longlong_as_double(0x7ff0000000000000LL);
  // per-table results of step B: red[2u], red[2u+1] = lo, hi of unit u (BITS = 16: lo after the shift, and whether any
  // entry was rounded); the candidate lists are idle now
  int *red = sm.cj;
  const int red_cap = 2 * nwarps * p.max_rec * f.cand_w;  // cj and ca are contiguous
  if (warp == 0) {
    // (A1) value ranges in doubles
    double lo_sum = 0.0, hi_sum = 0.0, range = 0.0, g_lo_min = 0.0;
    int kmer_need = 0, units = 0;
    bool bad = false;      // NaN anywhere in the tables
    bool cents = true;     // every recognizer entry is a multiple of 0.01 (PSSMs made by recalculate_pssm)
    for (int r = 0; r < n; ++r) {
      const RecDesc rd = sm.rec[r];
      const double *m = sm.pool + rd.data;
      double lo = 0.0, hi = 0.0;
      if (rd.kind == 0) {
        for (int c = lane; c < rd.len; c += 32) {
          lo += fmin(fmin(m[4 * c], m[4 * c + 1]), fmin(m[4 * c + 2], m[4 * c + 3]));
          hi += fmax(fmax(m[4 * c], m[4 * c + 1]), fmax(m[4 * c + 2], m[4 * c + 3]));
          bad |= (m[4 * c] != m[4 * c]) | (m[4 * c + 1] != m[4 * c + 1]) | (m[4 * c + 2] != m[4 * c + 2]) | (m[4 * c + 3] != m[4 * c + 3]);
          if (BITS == 16) {
#pragma unroll
            for (int b = 0; b < 4; ++b) {
              const double x = m[4 * c + b] * 100.0;
              cents = cents && (fabs(x - rint(x)) < 1e-7);
            }
          }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          lo += __shfl_xor_sync(0xffffffffu, lo, o);
          hi += __shfl_xor_sync(0xffffffffu, hi, o);
        }
        const int ng = (rd.len + 3) >> 2;
        if (lane == 0) {
          plan->ngroups[r] = ng;
          plan->kmer_off[r] = kmer_need;
          plan->err_half[r] = ng;
          plan->unit_begin[r] = units;
        }
        kmer_need += 256 * ng;
        units += ng;
      } else {
        // shape: alt[b] - null[b] over all bins (the bin actually selected varies per offset)
        lo = inf;
        hi = -inf;
        cents = false;
        // A NaN null bin (an empty bin of the null histogram) is never selected by a real k-mer; if it ever is, the
        // reference exits.  Such bins are left out of the range and carry a marker that sends the pair to the exact
        // kernel, which reports the error.
        for (int b = lane; b < rd.nb - 1; b += 32) {
          double x = m[(rd.nb - 1) + b] - m[b];
          bad |= (m[(rd.nb - 1) + b] != m[(rd.nb - 1) + b]);
          if (x != x) continue;
          if (x < BIG_NEG) x = BIG_NEG;
          lo = fmin(lo, x);
          hi = fmax(hi, x);
        }
        lo = warp_min(lo);
        hi = warp_max(hi);
        if (!(lo <= hi)) bad = true;  // every bin is NaN
        if (lane == 0) {
          plan->ngroups[r] = 0;
          plan->kmer_off[r] = kmer_need;
          plan->err_half[r] = 1;
          plan->unit_begin[r] = units;
        }
        kmer_need += rd.nb - 1;
        units += 1;
      }
      lo_sum += lo;
      hi_sum += hi;
      range += hi - lo;
      if (lane == 0) plan->rrange[r] = hi - lo;
    }
    for (int c = 0; c < n - 1; ++c) {
      const double *Gc = sm.G + (size_t)c * p.g_stride + kGPad;
      double lo = inf, hi = -inf;
      for (int g = lane; g <= P - 2; g += 32) {  // legit gaps; g = P-1 is the reference's -1e10 sentinel
        const double x = Gc[g];
        lo = fmin(lo, x);
        hi = fmax(hi, x);
        bad |= (x != x);
      }
      lo = warp_min(lo);
      hi = warp_max(hi);
      lo_sum += lo;
      hi_sum += hi;
      range += hi - lo;
      g_lo_min = fmin(g_lo_min, lo);
      if (lane == 0) plan->grange[c] = hi - lo;
    }
    bad = __any_sync(0xffffffffu, bad);
    cents = __all_sync(0xffffffffu, cents);
    __syncwarp();
    int ok = !bad && (P >= 2) && (kmer_need <= f.kmer_cap) && (2 * (units + n - 1) <= red_cap);
    double scale = 1.0;
    if (BITS == 32) {
      // (A2) scale: everything, including (n-1) sentinels of size ~range, must stay inside +-2^30
      // the sentinel sits below the lowest legit gap energy by more than the whole value range (see below), and up
      // to n-1 of them can add up on a (never selected) path
      const double need = fabs(lo_sum) + fabs(hi_sum) + (double)n * (range + fabs(g_lo_min) + 2.0) + 2.0;
      ok = ok && (need == need) && (need < 1.0e9);
      int e = 0;
      if (ok) {
        e = ilogb(1073741824.0 / need) - 1;
        if (e > 40) e = 40;
        if (e < 0) ok = 0;
      }
      scale = ok ? scalbn(1.0, e) : 1.0;
    } else {
      // (A2) every table is shifted to a maximum of 0, so row i lives in [-cum_i, 0] with cum_i the summed ranges up
      // to recognizer i; stage c+1 forms prev + H >= -cum_c + sentinel_c with sentinel_c ~ -(cum_c + grange_c):
      // 2 cum_c + grange_c (every stage) and cum_{n-1} must fit the int16 budget.
      double cum = plan->rrange[0], need = 0.0;
      for (int c = 0; c < n - 1; ++c) {
        need = fmax(need, 2.0 * cum + plan->grange[c]);
        cum += plan->grange[c] + plan->rrange[c + 1];
      }
      need = fmax(need, cum);
      // room for the thresholds inside the sentinel and for the rounding of every table's range
      const double budget = (double)kBudget16 - 4.0 * (units + n) - 16.0;
      ok = ok && (need == need) && (need > 0.0) && (need < 1.0e9);
      if (ok) {
        const double ratio = budget / need;
        if (cents && ratio >= 100.0) {
          int e = ilogb(ratio / 100.0);
          if (e > 6) e = 6;
          scale = scalbn(100.0, e);   // recognizer tables become exact integers
        } else {
          int e = ilogb(ratio);
          if (e > 12) e = 12;
          scale = scalbn(1.0, e);
        }
        if (!(scale >= kMinScale16)) ok = 0;
      }
      if (!ok) scale = 1.0;
    }
    if (lane == 0) {
      plan->scale = scale;
      plan->ok = ok;
      plan->units = units;
    }
  }
  __syncthreads();
  // (B) quantised tables and their integer ranges, one warp per table
  const double s = plan->scale;
  const int units = plan->units;
  if (plan->ok) {
    for (int u = warp; u < units + (n - 1); u += nwarps) {
      int lo = 0x7fffffff, hi = -0x7fffffff;
      bool rounded = false;   // BITS = 16: some entry of this table is not an integer after scaling
      if (u < units) {
        int r = 0;
        while (r + 1 < n && plan->unit_begin[r + 1] <= u) ++r;
        const RecDesc rd = sm.rec[r];
        const double *m = sm.pool + rd.data;
        const int off = plan->kmer_off[r];
        if (rd.kind != 0) {
          // shape recognizer: <= 256 bins, eight per lane
          int q[8];
          bool nanbin[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const int b = lane + 32 * e;
            q[e] = 0;
            nanbin[e] = true;
            if (b < rd.nb - 1) {
              double x = m[(rd.nb - 1) + b] - m[b];
              if (x == x) {
                if (x < BIG_NEG) x = BIG_NEG;
                q[e] = __double2int_rn(x * s);
                rounded = rounded || (fabs(x * s - (double)q[e]) > 1e-6);
                nanbin[e] = false;
                lo = min(lo, q[e]);
                hi = max(hi, q[e]);
              }
            }
          }
          const int shift = warp_max_i(hi);
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const int b = lane + 32 * e;
            if (b < rd.nb - 1) sm.kmer[off + b] = nanbin[e] ? kNanBinMarker : q[e] - shift;
          }
          if (lo <= hi) {   // lanes without a bin keep their neutral elements (shifting those would wrap)
            lo -= shift;
            hi -= shift;
          }
        } else {
          const int g = u - plan->unit_begin[r];
          int q[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const int kmer = lane + 32 * e;
            double acc = 0.0;
#pragma unroll
            for (int t = 0; t < 4; ++t) {
              const int col = 4 * g + t;
              if (col < rd.len) acc += m[4 * col + ((kmer >> (2 * t)) & 3)];
            }
            q[e] = __double2int_rn(acc * s);
            rounded = rounded || (fabs(acc * s - (double)q[e]) > 1e-6);
            lo = min(lo, q[e]);
            hi = max(hi, q[e]);
          }
          const int shift = warp_max_i(hi);
#pragma unroll
          for (int e = 0; e < 8; ++e) sm.kmer[off + 256 * g + lane + 32 * e] = q[e] - shift;
          lo -= shift;
          hi -= shift;
        }
      } else {
        const int c = u - units;
        const double *Gc = sm.G + (size_t)c * p.g_stride + kGPad;
        int *Gq = sm.Gq + (size_t)c * f.gq_stride + kGPad;
        for (int g = lane; g <= P - 2; g += 32) hi = max(hi, __double2int_rn(Gc[g] * s));
        const int shift = warp_max_i(hi);
        if (BITS == 32) {
          for (int g = lane; g <= P - 2; g += 32) {
            const int q = __double2int_rn(Gc[g] * s) - shift;
            Gq[g] = q;
            lo = min(lo, q);
          }
          hi = 0;
        } else {
          short *Hs = reinterpret_cast<short *>(Gq);   // low half of word g = G[g]; the high halves follow in step D
          for (int g = lane; g <= P - 2; g += 32) {
            const double x = Gc[g] * s;
            const int q = __double2int_rn(x) - shift;
            rounded = rounded || (fabs(x - rint(x)) > 1e-6);
            Hs[2 * g] = (short)max(q, -32768);   // a range beyond int16 is caught in step C (plan not ok)
            lo = min(lo, q);
          }
          hi = 0;
        }
      }
      lo = warp_min_i(lo);
      hi = warp_max_i(hi);
      rounded = __any_sync(0xffffffffu, rounded);
      if (lane == 0) {
        red[2 * u] = lo;
        red[2 * u + 1] = (BITS == 16) ? (rounded ? 1 : 0) : hi;
      }
    }
  }
  __syncthreads();
  // (C) integer ranges, sentinel(s), thresholds
  if (warp == 0) {
    int ok = plan->ok;
    if (BITS == 32) {
      long long lo_q = 0, hi_q = 0, lo_q_rec = 0;
      int g_lo_q = 0;
      if (ok) {
        for (int u = 0; u < units + (n - 1); ++u) {
          if (u == units) lo_q_rec = lo_q;
          lo_q += red[2 * u];
          hi_q += red[2 * u + 1];
          if (u >= units) g_lo_q = min(g_lo_q, red[2 * u]);
        }
        if (n == 1) lo_q_rec = lo_q;
      }
      const long long range_q = hi_q - lo_q;
      const int thr_max = units + n + 2 * kSlackUnits + 2;
      // sentinel: below the lowest legit gap energy by more than (range of any path) + (largest threshold), so a path
      // through it is below EVERY legit path of the same stage by more than any threshold
      const long long sent = (long long)g_lo_q - (range_q + thr_max + 4);
      if (ok && (lo_q_rec + (long long)(n - 1) * sent < -2000000000LL || hi_q > 2000000000LL || lo_q < -2000000000LL)) ok = 0;
      if (lane == 0) {
        plan->sent = (int)sent;
        plan->ok = ok;
        // err_i (in half units) = groups of recognizers 0..i + i connectors
        int half_units = 0;
        for (int i = 0; i < n; ++i) {
          // predecessors of stage i: values Cq_{i-1}[k] + Gq  ->  err_{i-1} + 0.5
          plan->thr_pred[i] = half_units + 1 + 2 * kSlackUnits;   // 2 * (err_{i-1} + 0.5) = half_units + 1   (i >= 1)
          half_units += plan->err_half[i] + (i > 0 ? 1 : 0);
        }
        plan->thr_final = half_units + 2 * kSlackUnits;          // 2 * err_{n-1}
      }
    } else if (lane == 0) {
      // thresholds first (a table whose entries are all integers after scaling contributes no error), then, per stage,
      // the sentinel and the int16 bound:  row c in [-cum_c, 0],  prev + H >= -cum_c + sent_c >= -32768
      int half_units = 0, thr_max = 0;
      if (ok) {
        for (int i = 0; i < n; ++i) {
          int e_rec = 0;
          const int u0 = plan->unit_begin[i], u1 = (i + 1 < n) ? plan->unit_begin[i + 1] : units;
          for (int u = u0; u < u1; ++u) e_rec += red[2 * u + 1];
          const int e_con = (i > 0) ? red[2 * (units + i - 1) + 1] : 0;
          plan->thr_pred[i] = half_units + e_con + 2 * kSlackUnits16;   // 2 * (err_{i-1} + rounding of the gap entry)
          half_units += e_rec + e_con;
          plan->err_half[i] = e_rec;
        }
        plan->thr_final = half_units + 2 * kSlackUnits16;
        thr_max = half_units + 2 * kSlackUnits16 + 2;
        long long cum = 0;
        for (int u = plan->unit_begin[0]; u < ((n > 1) ? plan->unit_begin[1] : units); ++u) cum -= red[2 * u];
        for (int c = 0; c < n - 1 && ok; ++c) {
          const long long gr = -(long long)red[2 * (units + c)];
          const long long sent = -(cum + gr + thr_max + 2);
          if (cum - sent > 32767 || gr > 32767) ok = 0;
          plan->sent_stage[c] = (int)sent;
          cum += gr;
          const int u0 = plan->unit_begin[c + 1], u1 = (c + 2 < n) ? plan->unit_begin[c + 2] : units;
          for (int u = u0; u < u1; ++u) cum -= red[2 * u];
        }
        if (cum > 32767) ok = 0;
      }
      plan->sent = -32768;
      plan->ok = ok;
    }
  }
  __syncthreads();
  // (D) sentinel pads of the connector tables
  if (plan->ok) {
    if (BITS == 32) {
      const int sent = plan->sent;
      for (int i = threadIdx.x; i < (n - 1) * f.gq_stride; i += blockDim.x) {
        const int g = i % f.gq_stride - kGPad;
        if (!(g >= 0 && g <= P - 2)) sm.Gq[i] = sent;
      }
    } else {
      // word d of connector c: lo half = G[d] (written in step B where legit), hi half = G[d-1]; the stage's sentinel
      // wherever the gap is not legit.  Only 16-bit stores: the low half a neighbour reads is never rewritten here.
      short *Hs = reinterpret_cast<short *>(sm.Gq);
      for (int i = threadIdx.x; i < (n - 1) * f.gq_stride; i += blockDim.x) {
        const int c = i / f.gq_stride;
        const int d = i - c * f.gq_stride - kGPad;
        const short sent = (short)plan->sent_stage[c];
        if (!(d >= 0 && d <= P - 2)) Hs[2 * i] = sent;
        Hs[2 * i + 1] = (d - 1 >= 0 && d - 1 <= P - 2) ? Hs[2 * (i - 1)] : sent;
      }
    }
  }
  // (E) tile bounds of every connector for the pruned stages (dp_stage_pruned), one warp per connector
  if constexpr (PRUNE) {
    __syncthreads();
    if (plan->ok) {
      const int NB = (P + J - 1) / J;
      for (int c = warp; c < n - 1; c += nwarps) {
        const int *Gq = sm.Gq + (size_t)c * f.gq_stride + kGPad;
        int *gt = sm.Gq + (size_t)c * f.gq_stride + f.gt_off;
        int *gtail = gt + f.nb_max;
        const int sent = gq_at<BITS>(Gq, P - 1);   // what every gap outside [0, P-2] holds
        for (int d = lane; d < NB; d += 32) {
          const int lo = max(0, (d - 1) * J + 1), hi = min(P - 2, (d + 1) * J - 1);
          int m = sent;
          for (int g = lo; g <= hi; ++g) m = max(m, gq_at<BITS>(Gq, g));
          gt[d] = m;
        }
        __syncwarp();
        if (lane == 0) {
          int best = gt[NB - 1], dstar = NB - 1;
          for (int d = NB - 1, run = -0x7fffffff - 1; d >= 0; --d) {
            run = max(run, gt[d]);
            gtail[d] = run;
            if (gt[d] >= best) {   // the smallest offset among equals
              best = gt[d];
              dstar = d;
            }
          }
          // peak tiles: within prune_drop energy units of the maximum -- but at most six of them: for a wide connector
          // the window is halved (down to 2 units) and the bound tests decide about the rest, which is what pays there
          int dlo = dstar, dhi = dstar;
          for (int drop_units = f.prune_drop;; drop_units >>= 1) {
            const double dropd = (double)drop_units * plan->scale;
            const int drop = dropd < 1.0e9 ? (int)dropd : 1000000000;
            dlo = dhi = dstar;
            while (dlo > 0 && gt[dlo - 1] >= best - drop) --dlo;
            while (dhi + 1 < NB && gt[dhi + 1] >= best - drop) ++dhi;
            if (dhi - dlo + 1 <= 6 || drop_units <= 2) break;
          }
          plan->dlo[c] = dlo;
          plan->dhi[c] = dhi;
          plan->prune[c] = ((dhi - dlo + 1 + 2) * f.prune_frac <= NB) ? 1 : 0;
        }
      }
    }
  }
}

// ---- exact refinement of the warp's NT pairs on their candidate sets ---------------------------------------------
// The NT pairs a warp has just pushed through the integer pass are refined TOGETHER.  Steps (1) and (2) -- the
// candidate cells of the last row and, top down, of every earlier stage -- run team-parallel: the TEAM lanes of a
// pair scan that pair's rows, all NT pairs in the same instructions; a lane collects a bitmask of the cells that pass
// the threshold and a min-reduction loop emits them in ascending order.  Steps (3)-(5) -- exact scores of the
// candidates, exact forward recurrence, first maximum and traceback -- run with one lane per (pair, stage) or
// (pair, candidate).  Candidate lists hold CandCap<BITS, NT>::kPerPair cells per stage and pair; a pair that needs more is handed
// to the exact kernel.
// Quantised rows.  Two-row mode: row i lives in the ping-pong buffer (i & 1); rows n-1 and n-2 are still in shared
// memory after the forward pass, older rows are streamed back from global scratch (cp.async) one stage ahead of the
// scan that needs them.  One-row mode (FastParams::single_row): only row n-1 is there, every older row is fetched when
// its scan starts.
__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gsrc) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

#ifdef FL_PHASE_CLOCKS
// profiling build only (tools/phase_clocks.sh): SM-clock ticks per phase, summed over the lane-0 threads of all warps
__device__ unsigned long long g_phase_clk[16];
// timeline of CTA 0: [warp][event] = (clock << 3) | phase that just ended
constexpr int kTimelineEvents = 8192;
__device__ unsigned long long g_timeline[16 * kTimelineEvents];
__device__ int g_timeline_n[16];
__device__ int g_timeline_slots;       // CTAs that arrived on the traced SM
__device__ int g_timeline_sm = 0;
#define PHASE_MARK(idx)                                                   \
  do {                                                                    \
    const long long now_ = clock64();                                     \
    if (lane == 0) {                                                      \
      atomicAdd(&g_phase_clk[idx], (unsigned long long)(now_ - phase_t0)); \
      const int row_ = tl_base + warp;                                    \
      if (tl_base >= 0 && row_ < 16 && g_timeline_n[row_] < kTimelineEvents) \
        g_timeline[row_ * kTimelineEvents + g_timeline_n[row_]++] = ((unsigned long long)now_ << 4) | (idx); \
    }                                                                     \
    phase_t0 = clock64();                                                 \
  } while (0)
#define PHASE_ARG , long long &phase_t0, int tl_base
#define PHASE_PASS , phase_t0, tl_base
#else
#define PHASE_MARK(idx) do { } while (0)
#define PHASE_ARG
#define PHASE_PASS
#endif

template <int NT, int BITS, class RowT>
__device__ __forceinline__ void refine_round(const FastParams &f, const OrgView &v, const FastSmemT<RowT> &sm, int warp, int lane,
                                             int nwarps, const int (&seq)[NT], const bool (&active)[NT], bool (&failed)[NT], int part_max PHASE_ARG) {
  constexpr int EPC = 16 / (int)sizeof(RowT);   // row elements per 16-byte cp.async chunk
  constexpr int TEAM = 32 / NT;
  constexpr int CAP = CandCap<BITS, NT>::kPerPair;
  constexpr int kCandW = NT * CAP;   // == f.cand_w
  constexpr int kLowest = -0x7fffffff - 1;
  constexpr int kNone = 0x7fffffff;
  const PlaceParams &p = f.base;
  const FastPlan &plan = *sm.plan;
  const int n = v.n, P = v.P;
  const int tl = lane % TEAM, team = lane / TEAM;
  // per-warp candidate storage, indexed [t][stage][c]
  int *cj = sm.cj + (size_t)warp * p.max_rec * kCandW;
  int *ca = sm.ca + (size_t)warp * p.max_rec * kCandW;
  int *cn = sm.cn + (size_t)warp * NT * p.max_rec;
  double *cv = sm.cv + (size_t)warp * p.max_rec * kCandW;
  double *cr = sm.cr + (size_t)warp * p.max_rec * kCandW;
  const size_t cls_stride = (size_t)p.n_seq_total * p.bins_row_stride;
  auto at = [&](int t, int i, int c) { return (t * p.max_rec + i) * CAP + c; };
  // value of a small per-pair register array at a lane-dependent pair index (select chain, no local memory)
  auto pick = [&](const int (&a)[NT], int t) {
    int r = a[0];
#pragma unroll
    for (int tt = 1; tt < NT; ++tt) r = (t == tt) ? a[tt] : r;
    return r;
  };

  const unsigned char *k8s[NT];
  const unsigned char *bins[NT];
#pragma unroll
  for (int t = 0; t < NT; ++t) {
    k8s[t] = sm.k8 + (size_t)(warp * NT + t) * f.k8_stride;
    bins[t] = p.bins + (size_t)seq[t] * p.bins_row_stride;
  }
  // Stages (1) and (2) run team-parallel: the TEAM lanes of a pair scan that pair's rows, the NT pairs of the warp
  // side by side in the same instructions.
  const int slot = warp * NT + team;
  RowT *my_rows = sm.rowsq + (size_t)slot * f.slotq_stride;
  const RowT *my_rowg = reinterpret_cast<const RowT *>(f.scratch) + ((size_t)blockIdx.x * nwarps * NT + slot) * f.scratch_slot_stride;
  const unsigned char *my_k8 = sm.k8 + (size_t)slot * f.k8_stride;
  int my_seq = seq[0];
  bool my_active = active[0];
#pragma unroll
  for (int t = 1; t < NT; ++t) {
    my_seq = (team == t) ? seq[t] : my_seq;
    my_active = (team == t) ? active[t] : my_active;
  }
  const unsigned char *my_bins = p.bins + (size_t)my_seq * p.bins_row_stride;
  bool my_failed = false;
  int cmax_all = 1;       // longest candidate list of any pair of the warp (warp-uniform)
  const bool single = f.single_row != 0;
  auto row_s = [&](int i) { return my_rows + (single ? (size_t)0 : (size_t)(i & 1) * f.rowq_stride); };
  // streams row i (<= n-3) of the team's pair from global scratch into its ping-pong buffer; the caller makes sure
  // the buffer's previous content (row i+2) is dead and waits (cp_async_wait_all + __syncwarp) before reading
  const int row_chunks = (P + EPC - 1) / EPC;
  auto prefetch_row = [&](int i) {
    RowT *dst = row_s(i);
    const RowT *src = my_rowg + (size_t)i * f.rowq_stride;
    for (int c = tl; c < row_chunks; c += TEAM) cp_async16(dst + EPC * c, src + EPC * c);
  };
  // Appends the cells flagged in `m` (bit b of lane tl = cell k0 + TEAM*b + tl) to the candidate list of stage i in
  // ascending order; `count` is the team's running total (lists are capped, the total keeps counting).
  auto emit = [&](unsigned m, int k0, int i, int &count) {
    for (;;) {
      const int mine = m ? k0 + TEAM * (__ffs(m) - 1) + tl : kNone;
      int best = mine;
#pragma unroll
      for (int o = TEAM >> 1; o > 0; o >>= 1) best = min(best, __shfl_xor_sync(0xffffffffu, best, o));
      if (!__any_sync(0xffffffffu, best != kNone)) break;
      if (best != kNone) {
        if (mine == best) {
          m &= m - 1;
          if (count < CAP) cj[at(team, i, count)] = best;
        }
        ++count;
        if (count > CAP) m = 0;  // the pair is handed back anyway: stop enumerating a massive tie
      }
    }
  };

  // (1) candidates on the last row: every offset within thr_final of the row maximum
  {
    const RowT *row = row_s(n - 1);
    int mx = part_max;  // the last score pass kept every lane's share of the row maximum
#pragma unroll
    for (int o = TEAM >> 1; o > 0; o >>= 1) mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    const int thr = mx - plan.thr_final;
    int count = 0;
    for (int k0 = 0; k0 < P; k0 += 32 * TEAM) {
      // predicated and unrolled in chunks of 8 independent loads
      unsigned m = 0;
      for (int b0 = 0; b0 < 32 && k0 + TEAM * b0 < P; b0 += 8) {
        int x[8];
#pragma unroll
        for (int b = 0; b < 8; ++b) {
          const int k = k0 + TEAM * (b0 + b) + tl;
          x[b] = (k < P) ? row[k] : kLowest;
        }
        unsigned mm = 0;
#pragma unroll
        for (int b = 0; b < 8; ++b) mm |= (x[b] >= thr ? 1u : 0u) << b;
        m |= mm << b0;
      }
      emit(my_active ? m : 0u, k0, n - 1, count);
    }
    if (count > CAP) my_failed = true;
    {
      int cm = min(count, CAP);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) cm = max(cm, __shfl_xor_sync(0xffffffffu, cm, o));
      cmax_all = max(cmax_all, cm);
    }
    if (tl == 0) cn[team * p.max_rec + n - 1] = min(count, CAP);
  }
  __syncwarp();
  PHASE_MARK(8);

  // (2) candidate sets of the earlier stages, top down: union over the stage's candidates j of the predecessors k
  //     with Cq_{i-1}[k] + Gq[j-k] within thr_pred of (Cq_i[j] - Rq_i[j])
  for (int i = n - 1; i >= 1; --i) {
    const int *Gq = sm.Gq + (size_t)(i - 1) * f.gq_stride + kGPad;
    const RecDesc rdi = sm.rec[i];
    if (!single && i - 1 <= n - 3) {  // row i-1 was requested during the previous stage
      cp_async_wait_all();
      __syncwarp();
    }
    const RowT *prev = row_s(i - 1);
    const RowT *cur = row_s(i);  // one-row mode: the same buffer; row i-1 replaces row i once the thresholds are known
    const int cnt = (my_active && !my_failed) ? cn[team * p.max_rec + i] : 0;
    int cmax = cnt, jj[CAP], thr[CAP], jmax = -1;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) cmax = max(cmax, __shfl_xor_sync(0xffffffffu, cmax, o));
#pragma unroll
    for (int c = 0; c < CAP; ++c) {
      jj[c] = -1;
      thr[c] = 0;
      if (c < cmax && c < cnt) {
        jj[c] = cj[at(team, i, c)];
        thr[c] = (cur[jj[c]] - score_q_any(plan, i, rdi, sm.kmer, my_k8, my_bins, cls_stride, jj[c])) - plan.thr_pred[i];
        jmax = max(jmax, jj[c]);
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) jmax = max(jmax, __shfl_xor_sync(0xffffffffu, jmax, o));
    if (single) {  // row i is dead from here on: row i-1 takes its place (the scan has to wait for it)
      __syncwarp();
      prefetch_row(i - 1);
      cp_async_wait_all();
      __syncwarp();
    } else if (i >= 2) {  // row i is dead from here on: its buffer receives row i-2 while row i-1 is scanned
      __syncwarp();
      prefetch_row(i - 2);
    }
    int count = 0;
    cmax_all = max(cmax_all, cmax);
    for (int k0 = 0; k0 <= jmax; k0 += 32 * TEAM) {
      // one pass per candidate rank (usually one; the packed 16-bit arithmetic lets two or three cells through now and
      // then): predicated and unrolled in chunks of 8 independent load pairs; a pair without a c-th candidate has
      // jj[c] = -1 and tests nothing.  The flags of all ranks are OR-ed: the stage's candidate set is the union.
      unsigned m = 0;
#pragma unroll
      for (int c = 0; c < CAP; ++c) {
        if (c < cmax) {  // warp-uniform
          const int j0 = jj[c], th = thr[c];
          for (int b0 = 0; b0 < 32 && k0 + TEAM * b0 <= jmax; b0 += 8) {
            int x[8];
#pragma unroll
            for (int b = 0; b < 8; ++b) {
              const int k = k0 + TEAM * (b0 + b) + tl;
              x[b] = (k <= j0) ? prev[k] + gq_at<BITS>(Gq, j0 - k) : kLowest;
            }
            unsigned mm = 0;
#pragma unroll
            for (int b = 0; b < 8; ++b) mm |= (x[b] >= th ? 1u : 0u) << b;
            m |= mm << b0;
          }
        }
      }
      emit(m, k0, i - 1, count);
    }
    if (count > CAP) my_failed = true;
    if (tl == 0) cn[team * p.max_rec + i - 1] = min(count, CAP);
    __syncwarp();
  }
#pragma unroll
  for (int t = 0; t < NT; ++t) failed[t] = __shfl_sync(0xffffffffu, (int)my_failed, t * TEAM) != 0;

  int okp[NT];
#pragma unroll
  for (int t = 0; t < NT; ++t) okp[t] = (active[t] && !failed[t]) ? 1 : 0;
  PHASE_MARK(9);

  // (3) exact scores of every candidate cell (they do not depend on the chain): one lane per (candidate rank, pair,
  //     stage) slot, rank-major, so the usual case -- one cell per list -- is a single pass over NT * n lanes
  {
    int cm = cn[team * p.max_rec];   // the stage-0 list was counted after the last cmax update
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) cm = max(cm, __shfl_xor_sync(0xffffffffu, cm, o));
    cmax_all = max(cmax_all, cm);
    const int per_rank = NT * n;
    for (int idx = lane; idx < per_rank * cmax_all; idx += 32) {
      const int c = idx / per_rank, rest = idx - c * per_rank;
      const int t = rest / n, i = rest - t * n;
      if (pick(okp, t) && c < cn[t * p.max_rec + i]) {
        const int slot_t = warp * NT + t;
        cr[at(t, i, c)] = score_exact_any(sm.rec[i], sm.pool, sm.k8 + (size_t)slot_t * f.k8_stride,
                                          p.bins + (size_t)pick(seq, t) * p.bins_row_stride, cls_stride, cj[at(t, i, c)]);
      }
    }
  }
  __syncwarp();
  PHASE_MARK(10);

  // (4) exact forward recurrence on the candidate sets: lane = (pair, candidate), all pairs side by side
  {
    const int t = lane / CAP, c = lane - t * CAP;
    const bool mine = (t < NT) && pick(okp, t);
    if (mine && c < cn[t * p.max_rec]) cv[at(t, 0, c)] = cr[at(t, 0, c)];
    __syncwarp();
    for (int i = 1; i < n; ++i) {
      if (mine && c < cn[t * p.max_rec + i]) {
        const int j = cj[at(t, i, c)];
        const double rj = cr[at(t, i, c)];
        const double *G = sm.G + (size_t)(i - 1) * p.g_stride + kGPad;
        double best = neg_inf();
        int arg = -1;
        const int np = cn[t * p.max_rec + i - 1];
        for (int d = 0; d < np; ++d) {  // ascending k: the first k that attains the maximum wins (organism.c:365)
          const int k = cj[at(t, i - 1, d)];
          if (k <= j) {
            const double cand = (cv[at(t, i - 1, d)] + G[j - k]) + rj;
            if (best < cand) {
              best = cand;
              arg = d;
            }
          }
        }
        cv[at(t, i, c)] = best;
        ca[at(t, i, c)] = arg;
      }
      __syncwarp();
    }
  }

  PHASE_MARK(11);
  // (5) first maximum of the last row (_aux.c:149-158) and traceback (organism.c:398-427): lane t serves pair t
  if (lane < NT && pick(okp, lane)) {
    const int t = lane;
    int c = 0;
    const int cnt = cn[t * p.max_rec + n - 1];
    for (int d = 1; d < cnt; ++d)
      if (cv[at(t, n - 1, d)] > cv[at(t, n - 1, c)]) c = d;
    const size_t pair = (size_t)v.o * p.n_seq_total + pick(seq, t);
    p.E[pair] = cv[at(t, n - 1, c)];
    if (f.want_trace) {
      for (int i = n - 1; i >= 1; --i) {
        const int d = ca[at(t, i, c)];
        const int j = cj[at(t, i, c)];
        const int k = cj[at(t, i - 1, d < 0 ? 0 : d)];
        const int gap = d < 0 ? 0 : j - k;
        p.rsc[pair * p.max_rec + i] = cr[at(t, i, c)];
        p.csc[pair * p.max_rec + i - 1] = d < 0 ? 0.0 : sm.G[(size_t)(i - 1) * p.g_stride + kGPad + gap];
        p.gaps[pair * p.max_rec + i] = gap;
        c = d < 0 ? 0 : d;
      }
      p.rsc[pair * p.max_rec] = cr[at(t, 0, c)];
      p.gaps[pair * p.max_rec] = cj[at(t, 0, c)];
    }
  }
  __syncwarp();
}

// Adds the quantised scores of recognizer r to a row (or, for stage 0, writes them): row[j] (+)= Rq_r[j], j < P,
// zeroes the J padding entries behind the row, and mirrors the finished row to global scratch when a later
// refinement needs it.  The recognizer kind and the number of 4-mer groups are hoisted out of the offset loop.
// Returns the largest value this lane wrote (the refinement of the last stage starts from the row maximum).
// A lane handles four consecutive offsets at a time: the 4-mer codes arrive as aligned words (funnel-shifted by the
// recognizer's byte offset), the row moves as int4.  Entries P..P+3 of the row's padding are written as zeros.
template <int NG, int TEAM, bool FIRST>
__device__ __forceinline__ int add_pssm_scores_q(const int *__restrict__ tab, int ng, const unsigned char *__restrict__ k8, int fwd,
                                                 int *__restrict__ row, int *__restrict__ glob, int P, int tl) {
  constexpr int kLow = -0x7fffffff - 1;
  int mx = kLow;
  const unsigned *kw = reinterpret_cast<const unsigned *>(k8 + (fwd & ~3));
  const int sh = 8 * (fwd & 3);
  const int groups = NG > 0 ? NG : ng;
#pragma unroll 2
  for (int j0 = 4 * tl; j0 < P; j0 += 4 * TEAM) {
    int4 v = FIRST ? make_int4(0, 0, 0, 0) : *reinterpret_cast<const int4 *>(row + j0);
    const unsigned *kp = kw + (j0 >> 2);
    unsigned w0 = kp[0];
#pragma unroll
    for (int g = 0; g < groups; ++g) {
      const unsigned w1 = kp[g + 1];
      const unsigned c = __funnelshift_r(w0, w1, sh);  // codes of the 4-mers at fwd + j0 + 4g + {0,1,2,3}
      const int *tg = tab + 256 * g;
      v.x += tg[c & 0xffu];
      v.y += tg[(c >> 8) & 0xffu];
      v.z += tg[(c >> 16) & 0xffu];
      v.w += tg[c >> 24];
      w0 = w1;
    }
    if (j0 + 3 >= P) {
      if (j0 + 1 >= P) v.y = 0;
      if (j0 + 2 >= P) v.z = 0;
      v.w = 0;
    }
    *reinterpret_cast<int4 *>(row + j0) = v;
    if (glob) *reinterpret_cast<int4 *>(glob + j0) = v;
    mx = max(mx, v.x);
    mx = max(mx, (j0 + 1 < P) ? v.y : kLow);
    mx = max(mx, (j0 + 2 < P) ? v.z : kLow);
    mx = max(mx, (j0 + 3 < P) ? v.w : kLow);
  }
  return mx;
}

// ---- the same passes on int16 rows: four offsets = two packed words (uint2) per lane and iteration ----
__device__ __forceinline__ int lo16(unsigned w) { return (int)(short)(w & 0xffffu); }
__device__ __forceinline__ int hi16(unsigned w) { return (int)w >> 16; }
__device__ __forceinline__ unsigned pack16(int lo, int hi) { return __byte_perm((unsigned)lo, (unsigned)hi, 0x5410); }

// adds the four quantised scores s[0..3] of offsets j0..j0+3 to the row, zeroes what lies beyond P-1, tracks the maximum
template <bool FIRST>
__device__ __forceinline__ void add4_row16(short *__restrict__ row, short *__restrict__ glob, int j0, int P, const int (&sc)[4], int &mx) {
  constexpr int kLow = -0x7fffffff - 1;
  uint2 v = FIRST ? make_uint2(0u, 0u) : *reinterpret_cast<const uint2 *>(row + j0);
  int r0 = lo16(v.x) + sc[0], r1 = hi16(v.x) + sc[1], r2 = lo16(v.y) + sc[2], r3 = hi16(v.y) + sc[3];
  if (j0 + 3 >= P) {
    if (j0 + 1 >= P) r1 = 0;
    if (j0 + 2 >= P) r2 = 0;
    r3 = 0;
  }
  v.x = pack16(r0, r1);
  v.y = pack16(r2, r3);
  *reinterpret_cast<uint2 *>(row + j0) = v;
  if (glob) *reinterpret_cast<uint2 *>(glob + j0) = v;
  mx = max(mx, r0);
  mx = max(mx, (j0 + 1 < P) ? r1 : kLow);
  mx = max(mx, (j0 + 2 < P) ? r2 : kLow);
  mx = max(mx, (j0 + 3 < P) ? r3 : kLow);
}

template <int NG, int TEAM, bool FIRST>
__device__ __forceinline__ int add_pssm_scores_q16(const int *__restrict__ tab, int ng, const unsigned char *__restrict__ k8, int fwd,
                                                   short *__restrict__ row, short *__restrict__ glob, int P, int tl) {
  int mx = -0x7fffffff - 1;
  const unsigned *kw = reinterpret_cast<const unsigned *>(k8 + (fwd & ~3));
  const int sh = 8 * (fwd & 3);
  const int groups = NG > 0 ? NG : ng;
#pragma unroll 2
  for (int j0 = 4 * tl; j0 < P; j0 += 4 * TEAM) {
    int sc[4] = {0, 0, 0, 0};
    const unsigned *kp = kw + (j0 >> 2);
    unsigned w0 = kp[0];
#pragma unroll
    for (int g = 0; g < groups; ++g) {
      const unsigned w1 = kp[g + 1];
      const unsigned c = __funnelshift_r(w0, w1, sh);  // codes of the 4-mers at fwd + j0 + 4g + {0,1,2,3}
      const int *tg = tab + 256 * g;
      sc[0] += tg[c & 0xffu];
      sc[1] += tg[(c >> 8) & 0xffu];
      sc[2] += tg[(c >> 16) & 0xffu];
      sc[3] += tg[c >> 24];
      w0 = w1;
    }
    add4_row16<FIRST>(row, glob, j0, P, sc, mx);
  }
  return mx;
}

template <int J, int TEAM, bool FIRST, class RowT>
__device__ __forceinline__ void add_scores_q(const FastPlan &plan, int r, const RecDesc &rd, const int *__restrict__ kmer,
                                             const unsigned char *__restrict__ k8, const unsigned char *__restrict__ bins_seq,
                                             size_t cls_stride, RowT *__restrict__ row, RowT *__restrict__ glob, int P, int tl,
                                             bool &nan_hit, int &row_max) {
  constexpr bool k16 = sizeof(RowT) == 2;
  if (rd.kind == 0) {
    const int *tab = kmer + plan.kmer_off[r];
    const int ng = plan.ngroups[r];
    if constexpr (k16) {
      switch (ng) {
        case 1: row_max = add_pssm_scores_q16<1, TEAM, FIRST>(tab, ng, k8, rd.fwd, row, glob, P, tl); break;
        case 2: row_max = add_pssm_scores_q16<2, TEAM, FIRST>(tab, ng, k8, rd.fwd, row, glob, P, tl); break;
        case 3: row_max = add_pssm_scores_q16<3, TEAM, FIRST>(tab, ng, k8, rd.fwd, row, glob, P, tl); break;
        case 4: row_max = add_pssm_scores_q16<4, TEAM, FIRST>(tab, ng, k8, rd.fwd, row, glob, P, tl); break;
        default: row_max = add_pssm_scores_q16<0, TEAM, FIRST>(tab, ng, k8, rd.fwd, row, glob, P, tl); break;
      }
    } else {
      switch (ng) {
        case 1: row_max = add_pssm_scores_q<1, TEAM, FIRST>(tab, ng, k8, rd.fwd, row, glob, P, tl); break;
        case 2: row_max = add_pssm_scores_q<2, TEAM, FIRST>(tab, ng, k8, rd.fwd, row, glob, P, tl); break;
        case 3: row_max = add_pssm_scores_q<3, TEAM, FIRST>(tab, ng, k8, rd.fwd, row, glob, P, tl); break;
        case 4: row_max = add_pssm_scores_q<4, TEAM, FIRST>(tab, ng, k8, rd.fwd, row, glob, P, tl); break;
        default: row_max = add_pssm_scores_q<0, TEAM, FIRST>(tab, ng, k8, rd.fwd, row, glob, P, tl); break;
      }
    }
  } else {
    // shape recognizer: the precomputed bin of each offset (global, one byte per offset) indexes the quantised bin
    // scores; four offsets per lane and iteration, bins fetched as aligned words like the 4-mer codes
    const int *tab = kmer + plan.kmer_off[r];
    const unsigned char *b = bins_seq + (size_t)rd.cls * cls_stride;
    const unsigned *bw = reinterpret_cast<const unsigned *>(b + (rd.fwd & ~3));
    const int sh = 8 * (rd.fwd & 3);
    constexpr int kLow = -0x7fffffff - 1;
    int mx = kLow;
#pragma unroll 2
    for (int j0 = 4 * tl; j0 < P; j0 += 4 * TEAM) {
      const unsigned w0 = __ldg(bw + (j0 >> 2)), w1 = __ldg(bw + (j0 >> 2) + 1);
      const unsigned c = __funnelshift_r(w0, w1, sh);
      int x[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        x[e] = (j0 + e < P) ? tab[(c >> (8 * e)) & 0xffu] : 0;
        if (x[e] == kNanBinMarker) {
          nan_hit = true;
          x[e] = 0;
        }
      }
      if constexpr (k16) {
        add4_row16<FIRST>(row, glob, j0, P, x, mx);
      } else {
        int4 v = FIRST ? make_int4(0, 0, 0, 0) : *reinterpret_cast<const int4 *>(row + j0);
        v.x += x[0];
        v.y += x[1];
        v.z += x[2];
        v.w += x[3];
        if (j0 + 3 >= P) {
          if (j0 + 1 >= P) v.y = 0;
          if (j0 + 2 >= P) v.z = 0;
          v.w = 0;
        }
        *reinterpret_cast<int4 *>(row + j0) = v;
        if (glob) *reinterpret_cast<int4 *>(glob + j0) = v;
        mx = max(mx, v.x);
        mx = max(mx, (j0 + 1 < P) ? v.y : kLow);
        mx = max(mx, (j0 + 2 < P) ? v.z : kLow);
        mx = max(mx, (j0 + 3 < P) ? v.w : kLow);
      }
    }
    row_max = mx;
  }
  for (int t = tl; t < J; t += TEAM) row[P + t] = 0;  // padding behind the row: 0 + sentinel cannot wrap
}

// ------------------------------------------------------------------------------------------------
// The fast kernel.  CTA = one organism x a chunk of sequences; a warp holds 32/TEAM pairs at a time, each
// placed by a team of TEAM lanes with J x J register tiles of DPX cells.
// ------------------------------------------------------------------------------------------------

// (20 warps per CTA for the packed kernel -- its rows are half as large -- were measured: launch bound 640 squeezes it to
// 86 registers and config 2 takes 2.98 ms instead of 2.85-2.90 ms with 16 warps.)
template <int J, int TEAM, int BITS, bool PRUNE = false>
__global__ void __launch_bounds__(512, 1) place_fast_kernel(const FastParams f) {
  constexpr int NT = 32 / TEAM;
  typedef typename RowOf<BITS>::T RowT;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const PlaceParams &p = f.base;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const int tl = lane % TEAM, team = lane / TEAM;
  const FastSmemT<RowT> sm = carve_fast<RowT>(f, smem_raw, nwarps, NT);
  // persistent CTAs: work items (organism x chunk of sequences) are claimed from a device-side counter, so the
  // global row scratch is sized by the resident CTAs, not by the number of items, and the tail balances itself
#ifdef FL_PHASE_CLOCKS
  // the CTAs resident on one SM share the 16 timeline rows (rows = arrival order x warps per CTA)
  if (threadIdx.x == 0) {
    unsigned smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    sm.plan->trace_base = ((int)smid == g_timeline_sm) ? atomicAdd(&g_timeline_slots, 1) * nwarps : -1;
  }
  __syncthreads();
  const int tl_base = sm.plan->trace_base;
  long long phase_t0 = clock64();
#endif
  for (;;) {
  __syncthreads();
  PHASE_MARK(7);
  if (threadIdx.x == 0) {
    sm.plan->item = atomicAdd(f.work_counter, 1);
    sm.plan->next_round = 0;
  }
  __syncthreads();
  const int item = sm.plan->item;
  if (item >= f.n_items) break;
  const int oi = item / p.chunks;
  const int chunk = item - oi * p.chunks;
  const int first = chunk * p.seqs_per_cta;
  int n_in = min(p.class_count, first + p.seqs_per_cta) - first;   // sequences of this item
  const int *in_seqs = nullptr;
  if (f.retry_count != nullptr) {
    n_in = __ldg(f.retry_count + item);
    if (n_in == 0) continue;   // uniform: nothing of this item was handed back
    in_seqs = f.retry_seqs + (long long)item * p.seqs_per_cta;
    if (threadIdx.x == 0) atomicAdd(f.retried_total, (unsigned long long)n_in);
  }
  const OrgView v = organism_prologue(p, p.org_list[oi], sm.pool, sm.G, sm.rec);
  __syncthreads();
  PHASE_MARK(0);
  build_fast_plan<BITS, RowT, J, PRUNE>(f, v, sm, warp, lane, nwarps);
  __syncthreads();
  PHASE_MARK(1);

  const FastPlan &plan = *sm.plan;
  const int n = v.n, P = v.P;
  const int slot = warp * NT + team;
  RowT *rowsq = sm.rowsq + (size_t)slot * f.slotq_stride;
  RowT *rowsg = reinterpret_cast<RowT *>(f.scratch) + ((size_t)blockIdx.x * nwarps * NT + slot) * f.scratch_slot_stride;
  unsigned char *k8 = sm.k8 + (size_t)slot * f.k8_stride;

  // rounds are claimed dynamically: a warp that met a pair with several candidates (a longer refinement) takes fewer
  // rounds, and the warps of the CTA drift apart instead of marching through the phases in lockstep
  // (static assignment -- warp w takes rounds w, w + nwarps, ... -- measures the same: 162.3 vs 162.8 ms on config 3)
  for (;;) {
    int round = 0;
    if (lane == 0) round = atomicAdd(&sm.plan->next_round, 1);
    round = __shfl_sync(0xffffffffu, round, 0);
    const int sbase = round * NT;
    if (sbase >= n_in) break;
    // the warp's NT pairs of this round (idle teams replay the last valid sequence and write nothing)
    const int si = min(sbase + team, n_in - 1);
    const bool live = (sbase + team) < n_in;
    const int seq = in_seqs ? __ldg(in_seqs + si) : p.order[p.class_begin + first + si];
    bool decline = !plan.ok || f.has_unk[seq];
    bool nan_hit = false;
    int row_max = -0x7fffffff - 1;  // this lane's share of the maximum of the row written last
    // ---- phase 1: integer pass ---------------------------------------------------------------------
    // (skipped when every pair of the warp's round is declined up front -- sequences with unknown bytes; a second-chance
    // launch over such data would otherwise repeat the work the first launch already wasted)
    if (plan.ok && !__all_sync(0xffffffffu, decline)) {
      {  // the sequence's 4-mer codes, precomputed at upload: one 16-byte cp.async per chunk
        const unsigned char *src = p.k8g + 16ull * (unsigned long long)p.k8_off[seq];
        for (int c = tl; c < (f.k8_stride >> 4); c += TEAM) cp_async16(k8 + 16 * c, src + 16 * c);
        cp_async_wait_all();
      }
      __syncwarp();
      PHASE_MARK(2);
      const unsigned char *bins_seq = p.bins + (size_t)seq * p.bins_row_stride;
      const size_t cls_stride = (size_t)p.n_seq_total * p.bins_row_stride;
      // rows the refinement will have to fetch from global scratch: all but the last two (two-row mode) or the last one
      const int last_mirrored = n - 2 - (f.single_row ? 0 : 1);
      add_scores_q<J, TEAM, true, RowT>(plan, 0, sm.rec[0], sm.kmer, k8, bins_seq, cls_stride, rowsq,
                                        (0 <= last_mirrored) ? rowsg : static_cast<RowT *>(nullptr), P, tl, nan_hit, row_max);
      __syncwarp();
      PHASE_MARK(3);
      for (int i = 1; i < n; ++i) {
        const RowT *prev = f.single_row ? rowsq : rowsq + (size_t)((i - 1) & 1) * f.rowq_stride;
        RowT *cur = f.single_row ? rowsq : rowsq + (size_t)(i & 1) * f.rowq_stride;
        const int *Gq = sm.Gq + (size_t)(i - 1) * f.gq_stride + kGPad;
        bool pruned = false;
        if constexpr (PRUNE) {
          if (plan.prune[i - 1]) {   // uniform over the CTA: a property of the organism's connector
            pruned = true;
            PruneTabs tabs;
            tabs.gt = sm.Gq + (size_t)(i - 1) * f.gq_stride + f.gt_off;
            tabs.gtail = tabs.gt + f.nb_max;
            tabs.dlo = plan.dlo[i - 1];
            tabs.dhi = plan.dhi[i - 1];
            int *bm = sm.bm + (size_t)slot * f.bm_stride;
            const int tiles_run = dp_stage_pruned<J, TEAM, RowT>(prev, cur, Gq, tabs, bm, bm + f.nb_max, bm + 2 * f.nb_max, P, tl);
#ifdef FL_PHASE_CLOCKS
            {   // tiles run / tiles of the full triangle / pruned pair-stages
              int sum = live ? (tiles_run & 0xffff) : 0;
#pragma unroll
              for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
              if (lane == 0) atomicAdd(&g_phase_clk[15], 32ull * (unsigned long long)(tiles_run >> 16));
              const unsigned live_teams = __popc(__ballot_sync(0xffffffffu, live && tl == 0));
              if (lane == 0) {
                const unsigned long long nb = (unsigned long long)((P + J - 1) / J);
                atomicAdd(&g_phase_clk[12], (unsigned long long)sum);
                atomicAdd(&g_phase_clk[13], live_teams * (nb * (nb + 1) / 2));
                atomicAdd(&g_phase_clk[14], (unsigned long long)live_teams);
              }
            }
#else
            (void)tiles_run;
#endif
          }
        }
        if (pruned) {
        } else if constexpr (BITS == 16)
          dp_stage16<J, TEAM>(prev, cur, reinterpret_cast<const unsigned *>(Gq), P, tl);
        else
          // accumulators start at the lowest int, NOT at the sentinel: the sentinel only has to lie below the legit gap
          // energies of its stage, while the row values themselves may lie far below it (recognizers whose scores are all
          // strongly negative, e.g. MLE-fitted shape recognizers with sigma = 0: ln(1e-100) per recognizer)
          dp_stage<J, TEAM, OpI32>(prev, cur, Gq, P, tl, -0x7fffffff - 1);
        __syncwarp();
        PHASE_MARK(4);
        add_scores_q<J, TEAM, false, RowT>(plan, i, sm.rec[i], sm.kmer, k8, bins_seq, cls_stride, cur,
                                           (i <= last_mirrored) ? rowsg + (size_t)i * f.rowq_stride : static_cast<RowT *>(nullptr), P, tl,
                                           nan_hit, row_max);
        __syncwarp();
        PHASE_MARK(3);
      }
    }
    // a selected NaN-null bin: the pair goes to the exact kernel, which raises the reference's error
    {
      const unsigned hit = __ballot_sync(0xffffffffu, nan_hit);
      if ((hit >> (team * TEAM)) & ((TEAM == 32) ? 0xffffffffu : ((1u << TEAM) - 1u))) decline = true;
    }
    // ---- phase 2: exact refinement of the warp's NT pairs, side by side ------------------------------------------
    {
      int r_seq[NT];
      bool r_active[NT], r_failed[NT], r_live[NT];
#pragma unroll
      for (int t = 0; t < NT; ++t) {
        r_seq[t] = __shfl_sync(0xffffffffu, seq, t * TEAM);
        r_live[t] = __shfl_sync(0xffffffffu, (int)live, t * TEAM) != 0;
        r_active[t] = r_live[t] && !__shfl_sync(0xffffffffu, (int)decline, t * TEAM);
      }
      PHASE_MARK(6);
      refine_round<NT, BITS, RowT>(f, v, sm, warp, lane, nwarps, r_seq, r_active, r_failed, row_max PHASE_PASS);
      PHASE_MARK(5);
      if (lane == 0) {
#pragma unroll
        for (int t = 0; t < NT; ++t) {
          if (r_live[t] && (!r_active[t] || r_failed[t])) {
            const int idx = atomicAdd(f.fb_item_count + f.item_base + item, 1);  // < seqs_per_cta: a pair is handed back once
            f.fb_item_seqs[f.pair_base + (long long)item * p.seqs_per_cta + idx] = r_seq[t];
          }
        }
      }
    }
    __syncwarp();
  }
  }  // work items
}
