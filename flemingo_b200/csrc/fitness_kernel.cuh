// fitness_kernel.cuh -- K3: per-organism fitness statistic (included by flemingo_b200.cu)
#pragma once
// ------------------------------------------------------------------------------------------------
// K3: per-organism fitness statistic (organism_object.py:848-956).  One CTA per organism.
// ------------------------------------------------------------------------------------------------
__device__ double block_sum(double v, double *s_red) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
  __syncthreads();
  if (lane == 0) s_red[warp] = v;
  __syncthreads();
  double t = 0.0;
  for (int w = 0; w < nw; ++w) t += s_red[w];
  return t;
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
                            double *s_red) {
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
  const int cnt = hi - lo;
  double acc = 0.0;
  for (int i = lo + threadIdx.x; i < hi; i += blockDim.x) acc += src[i];
  const double total = block_sum(acc, s_red);
  r.mean = (cnt <= 0 || empty_mean) ? __longlong_as_double(0x7ff8000000000000LL) : total / (double)cnt;
  // np.std(ddof=1): mean of the std window, then sum of squared deviations / (N-1)
  const int scnt = sd_hi - sd_lo;
  double a2 = 0.0;
  for (int i = sd_lo + threadIdx.x; i < sd_hi; i += blockDim.x) a2 += src[i];
  const double smean = block_sum(a2, s_red) / (double)scnt;
  double q = 0.0;
  for (int i = sd_lo + threadIdx.x; i < sd_hi; i += blockDim.x) {
    const double d = src[i] - smean;
    q += d * d;
  }
  const double ss = block_sum(q, s_red);
  double sd = sqrt(ss / (double)(scnt - 1));
  if (!(sd > 1.0)) sd = 1.0;  // max(1, stdev): python's max keeps 1 when stdev is nan (N == 1)
  r.se = sd / sqrt((double)scnt);
  return r;
}

__global__ void __launch_bounds__(256) fitness_kernel(const double *__restrict__ Epos, int n_pos, const double *__restrict__ Eneg,
                                                      int n_neg, int method, int trim_pos, int trim_neg, double *scratch,
                                                      int n2, double *__restrict__ out) {
  __shared__ double s_red[32];
  const int o = blockIdx.x;
  double *sc = scratch + (size_t)o * n2;
  const SetStat sp = set_stat(Epos + (size_t)o * n_pos, n_pos, method, trim_pos, sc, n2, s_red);
  __syncthreads();
  const SetStat sn = set_stat(Eneg + (size_t)o * n_neg, n_neg, method, trim_neg, sc, n2, s_red);
  if (threadIdx.x == 0) {
    out[o * 5 + 0] = (sp.mean - sn.mean) / sqrt(sp.se * sp.se + sn.se * sn.se);
    out[o * 5 + 1] = sp.mean;
    out[o * 5 + 2] = sp.se;
    out[o * 5 + 3] = sn.mean;
    out[o * 5 + 4] = sn.se;
  }
}

