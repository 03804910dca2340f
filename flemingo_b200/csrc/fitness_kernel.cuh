// fitness_kernel.cuh -- K3: per-organism fitness statistic (included by flemingo_b200.cu)
#pragma once
// ------------------------------------------------------------------------------------------------
// K3: per-organism fitness statistic (organism_object.py:848-956).  One CTA per organism.
// ------------------------------------------------------------------------------------------------
//
// The statistics are bit-identical to the reference's numpy calls, which fixes the summation order: np.mean / np.std
// reduce a contiguous float64 array with numpy's PAIRWISE summation (numpy/_core/src/umath/loops_utils.h.src,
// DOUBLE_pairwise_sum): fewer than 8 elements are added left to right starting from -0.0; up to 128 elements go through
// eight interleaved accumulators r[j] += a[8i + j], combined as ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)), the remainder
// added one by one; longer arrays are split at n/2 rounded down to a multiple of 8 and the halves added.  The leaves of
// that recursion are independent, so the CTA computes them in parallel (one thread per leaf) and one thread folds the
// leaf sums in the recursion's order.  np.std(ddof=1) is sqrt(pairwise_sum((x - mean)^2) / (n - 1)) with
// mean = pairwise_sum(x) / n.  (Verified against numpy on every length 1..300 and on 1000..40001.)
constexpr int kPwLeafCap = 1024;   // leaves of one reduction: n / 64 at most, i.e. sets of up to 65 536 sequences
struct PairwisePlan {
  int n_leaves, n_tokens;
  int leaf_lo[kPwLeafCap];
  short leaf_n[kPwLeafCap];
  unsigned char token[2 * kPwLeafCap];  // post-order walk of the recursion: 0 = next leaf, 1 = add the two on top
  double leaf_sum[kPwLeafCap];
};

// the recursion tree for n elements (thread 0): leaves in address order, and the order in which their sums are folded
__device__ void pairwise_plan(PairwisePlan &pl, int n) {
  int lo_s[40], n_s[40], ph_s[40];
  int top = 0, leaves = 0, tokens = 0;
  lo_s[0] = 0;
  n_s[0] = n;
  ph_s[0] = 0;
  while (top >= 0) {
    const int lo = lo_s[top], cnt = n_s[top];
    if (cnt <= 128) {
      pl.leaf_lo[leaves] = lo;
      pl.leaf_n[leaves] = (short)cnt;
      ++leaves;
      pl.token[tokens++] = 0;
      --top;
      continue;
    }
    int n2 = cnt / 2;
    n2 -= n2 % 8;
    if (ph_s[top] == 0) {
      ph_s[top] = 1;
      ++top;
      lo_s[top] = lo;
      n_s[top] = n2;
      ph_s[top] = 0;
    } else if (ph_s[top] == 1) {
      ph_s[top] = 2;
      ++top;
      lo_s[top] = lo + n2;
      n_s[top] = cnt - n2;
      ph_s[top] = 0;
    } else {
      pl.token[tokens++] = 1;
      --top;
    }
  }
  pl.n_leaves = leaves;
  pl.n_tokens = tokens;
}

// one leaf (n <= 128) of  sum_i f(a[i]),  f(x) = x  or  (x - centre)^2, in numpy's order
template <bool SQUARED>
__device__ __forceinline__ double pairwise_leaf(const double *__restrict__ a, int n, double centre) {
  auto f = [&](int i) {
    const double x = a[i];
    if (!SQUARED) return x;
    const double d = x - centre;
    return d * d;
  };
  if (n < 8) {
    double res = -0.0;
    for (int i = 0; i < n; ++i) res += f(i);
    return res;
  }
  double r[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) r[j] = f(j);
  int i = 8;
  for (; i < n - (n % 8); i += 8) {
#pragma unroll
    for (int j = 0; j < 8; ++j) r[j] += f(i + j);
  }
  double res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
  for (; i < n; ++i) res += f(i);
  return res;
}

// numpy's pairwise sum of f(a[0..n)) by the whole CTA; the plan must have been made for this n.  Returns on all threads.
template <bool SQUARED>
__device__ double pairwise_sum_cta(const double *__restrict__ a, int n, double centre, PairwisePlan &pl, double *s_result) {
  for (int l = threadIdx.x; l < pl.n_leaves; l += blockDim.x)
    pl.leaf_sum[l] = pairwise_leaf<SQUARED>(a + pl.leaf_lo[l], (int)pl.leaf_n[l], centre);
  __syncthreads();
  if (threadIdx.x == 0) {
    double st[40];
    int top = 0, next = 0;
    for (int t = 0; t < pl.n_tokens; ++t) {
      if (pl.token[t] == 0) {
        st[top++] = pl.leaf_sum[next++];
      } else {
        st[top - 2] = st[top - 2] + st[top - 1];
        --top;
      }
    }
    *s_result = (n > 0) ? st[0] : 0.0;   // np.sum of an empty array is 0.0
  }
  __syncthreads();
  const double out = *s_result;
  __syncthreads();
  return out;
}

// in-place bitonic sort of n2 (power of two) doubles by one CTA
__device__ void block_bitonic_sort(double *a, int n2) {
  for (int k = 2; k <= n2; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = threadIdx.x; i < n2; i += blockDim.x) {
        const int l = i ^ j;
        if (l > i) {
          const double x = a[i], y = a[l];
          const bool up = ((i & k) == 0);
          if ((x > y) == up) { a[i] = y; a[l] = x; }
        }
      }
      __syncthreads();
    }
  }
}

struct SetStat {
  double mean, se;
};

// mean / standard error of one organism's energies on one set; `scratch` (n2 doubles) is used by the
// trimmed variants.  lo/hi = number of elements dropped from the sorted ends for the mean.
__device__ SetStat set_stat(const double *__restrict__ e, int n, int method, int n_trim, double *scratch, int n2,
                            PairwisePlan &pl, double *s_result) {
  SetStat r;
  const double *src = e;
  int lo = 0, hi = n;          // mean over sorted[lo:hi]
  int sd_lo = 0, sd_hi = n;    // std over [sd_lo:sd_hi]
  bool empty_mean = false;
  if (n_trim != 0 && method != FL_FIT_WELCH) {
    for (int i = threadIdx.x; i < n2; i += blockDim.x) scratch[i] = i < n ? e[i] : __longlong_as_double(0x7ff0000000000000LL);
    __syncthreads();
    block_bitonic_sort(scratch, n2);
    src = scratch;
    if (method == FL_FIT_YUEN) {
      const int t = min(n_trim, (n - 1) / 2);
      lo = t;
      hi = n - t;
      if (t == 0) empty_mean = true;  // python: energy_scores[0:-0] is the empty list
    } else {
      lo = min(n_trim, n);
      if (method == FL_FIT_TRIM_LEFT_M_SD) sd_lo = lo;
    }
  }
  // np.mean(window)
  const int cnt = hi - lo;
  if (threadIdx.x == 0) pairwise_plan(pl, max(cnt, 0));
  __syncthreads();
  const double total = pairwise_sum_cta<false>(src + lo, max(cnt, 0), 0.0, pl, s_result);
  r.mean = (cnt <= 0 || empty_mean) ? __longlong_as_double(0x7ff8000000000000LL) : total / (double)cnt;
  // np.std(window, ddof=1): its own mean, then the squared deviations / (N - 1)
  const int scnt = sd_hi - sd_lo;
  double smean = total / (double)cnt;
  if (sd_lo != lo || sd_hi != hi) {
    if (threadIdx.x == 0) pairwise_plan(pl, scnt);
    __syncthreads();
    smean = pairwise_sum_cta<false>(src + sd_lo, scnt, 0.0, pl, s_result) / (double)scnt;
  }
  const double ss = pairwise_sum_cta<true>(src + sd_lo, scnt, smean, pl, s_result);
  double sd = sqrt(ss / (double)max(scnt - 1, 0));
  if (!(sd > 1.0)) sd = 1.0;  // max(1, stdev): python's max keeps 1 when stdev is nan (N == 1)
  r.se = sd / sqrt((double)scnt);
  return r;
}

__global__ void __launch_bounds__(256) fitness_kernel(const double *__restrict__ Epos, int n_pos, const double *__restrict__ Eneg,
                                                      int n_neg, int method, int trim_pos, int trim_neg, double *scratch,
                                                      int n2, double *__restrict__ out) {
  __shared__ PairwisePlan pl;
  __shared__ double s_result;
  const int o = blockIdx.x;
  double *sc = scratch + (size_t)o * n2;
  const SetStat sp = set_stat(Epos + (size_t)o * n_pos, n_pos, method, trim_pos, sc, n2, pl, &s_result);
  __syncthreads();
  const SetStat sn = set_stat(Eneg + (size_t)o * n_neg, n_neg, method, trim_neg, sc, n2, pl, &s_result);
  if (threadIdx.x == 0) {
    // The reference forms (M_p - M_n) / (SE_p ** 2 + SE_n ** 2) ** (1/2) with numpy scalars, i.e. libm pow(); the host
    // entry points recompute this column with the same libm calls (fl_fitness_finalize); this is the device's value
    // for consumers that stay on the device.
    out[o * 5 + 0] = (sp.mean - sn.mean) / sqrt(sp.se * sp.se + sn.se * sn.se);
    out[o * 5 + 1] = sp.mean;
    out[o * 5 + 2] = sp.se;
    out[o * 5 + 3] = sn.mean;
    out[o * 5 + 4] = sn.se;
  }
}

