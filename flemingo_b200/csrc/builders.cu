// builders.cu -- host-side table builders behind the C ABI (include/flemingo_b200.h, "table builders").
//
// These are the input producers of the placement path: the per-connector pdf/cdf tables, the
// 2-decimal PSSM entries and the shape alt models.  In the reference they are per-object Python loops
// (connector_object.py:175-206, pssm_object.py:368-400, shape_object.py:156-180) that dominate the host time of a GA
// generation once the placements run on the GPU.  Here they are plain C++ loops over whole batches, fanned out over
// host threads.
//
// Bit-identity with the reference is the contract, which fixes the split of the work:
//   * Phi(x) is evaluated with libm's erf(), the function CPython's math.erf calls, in the reference's operand order;
//   * the logarithms are NOT taken here: the reference calls numpy's log2 / log, whose SIMD kernels differ from
//     libm's in the last bit for a fraction of the arguments, so the builders return the ARGUMENTS of the logarithm
//     and the Python layer applies numpy's function to the whole batch in one vectorised call;
//   * rounding to two decimals follows Python's round(x, 2): the correctly rounded decimal expansion of the double
//     (printf "%.2f" is exact in glibc), parsed back with strtod.
// No device code in this file; it is compiled into libflemingo_b200.so by the same nvcc command.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <thread>
#include <vector>

#include "flemingo_b200.h"

namespace {

// norm_cdf of connector_object.py:19-30 / shape_object.py:8-19
inline double norm_cdf(double x, double mu, double sigma) {
  if (sigma == 0.0) return x < mu ? 0.0 : (x > mu ? 1.0 : 0.5);
  const double z = (x - mu) / std::fabs(sigma);
  return (1.0 + std::erf(z / std::sqrt(2.0))) / 2.0;
}

template <class F>
void parallel_for(long long n, long long grain, int max_threads, F body) {
  int hw = (int)std::thread::hardware_concurrency();
  if (hw < 1) hw = 1;
  if (max_threads > 0) hw = std::min(hw, max_threads);
  const long long want = (n + grain - 1) / grain;
  const int nt = (int)std::max<long long>(1, std::min<long long>(hw, want));
  if (nt == 1) {
    body(0, n);
    return;
  }
  std::vector<std::thread> pool;
  pool.reserve(nt);
  const long long step = (n + nt - 1) / nt;
  for (int t = 0; t < nt; ++t) {
    const long long lo = t * step, hi = std::min(n, lo + step);
    if (lo >= hi) break;
    pool.emplace_back([=] { body(lo, hi); });
  }
  for (auto &th : pool) th.join();
}

}  // namespace

extern "C" int fl_build_connector_tables(const double *mu, const double *sigma, int n_con, int M, double pseudo_count,
                                         double *pdf_arg, double *cdf_arg, int n_threads) {
  if (n_con < 0 || M < 1 || (n_con > 0 && (!mu || !sigma || !pdf_arg || !cdf_arg))) return FL_E_ARG;
  parallel_for(n_con, std::max(1, 4096 / M), n_threads, [=](long long lo, long long hi) {
    for (long long c = lo; c < hi; ++c) {
      double *pa = pdf_arg + c * (long long)M, *ca = cdf_arg + c * (long long)M;
      const double m = mu[c], s = sigma[c];
      double prev = norm_cdf(-0.5, m, s);
      const double first = prev;
      for (int j = 1; j <= M; ++j) {
        const double cur = norm_cdf((double)j - 0.5, m, s);
        ca[j - 1] = prev - first + (pseudo_count * (double)j);
        pa[j - 1] = cur - prev + pseudo_count;
        prev = cur;
      }
    }
  });
  return FL_OK;
}

extern "C" int fl_round_decimals(double *values, long long n, int ndigits) {
  if (n < 0 || (n > 0 && !values) || ndigits < 0 || ndigits > 15) return FL_E_ARG;
  parallel_for(n, 4096, 0, [=](long long lo, long long hi) {
    char buf[400];
    for (long long i = lo; i < hi; ++i) {
      const double x = values[i];
      if (!std::isfinite(x)) continue;  // round() returns inf / nan unchanged
      std::snprintf(buf, sizeof buf, "%.*f", ndigits, x);
      values[i] = std::strtod(buf, nullptr);
    }
  });
  return FL_OK;
}

extern "C" int fl_build_shape_masses(const double *mu, const double *sigma, const double *edges, const long long *edge_begin,
                                     const int *n_edges, int n_shapes, double pseudo_count, double *mass_arg, double *auc_arg) {
  if (n_shapes < 0 || (n_shapes > 0 && (!mu || !sigma || !edges || !edge_begin || !n_edges || !mass_arg || !auc_arg))) return FL_E_ARG;
  for (int s = 0; s < n_shapes; ++s)
    if (n_edges[s] < 2) return FL_E_ARG;
  parallel_for(n_shapes, 64, 0, [=](long long lo, long long hi) {
    for (long long s = lo; s < hi; ++s) {
      const double *e = edges + edge_begin[s];
      const int nb = n_edges[s] - 1;
      // mass_arg rows are packed like the edges, one entry fewer per shape: row s starts at edge_begin[s] - s
      double *out = mass_arg + (edge_begin[s] - s);
      const double m = mu[s], sg = sigma[s];
      double auc = norm_cdf(e[nb], m, sg) - norm_cdf(e[0], m, sg);
      auc += (double)nb * pseudo_count;
      auc_arg[s] = auc;
      for (int b = 0; b < nb; ++b) {
        double mass;
        if (sg != 0.0)
          mass = norm_cdf(e[b + 1], m, sg) - norm_cdf(e[b], m, sg);
        else
          mass = (e[b] == m) ? 1.0 : 0.0;  // shape_object.py:21-36
        out[b] = mass + pseudo_count;
      }
    }
  });
  return FL_OK;
}

extern "C" int fl_fitness_finalize(double *stats, long long n_org) {
  if (n_org < 0 || (n_org > 0 && !stats)) return FL_E_ARG;
  // organism_object.py:954-956 with numpy scalars: `**` is libm pow(), also for the squares.  The exponents are read
  // through volatiles so that the compiler cannot fold pow(x, 2.0) into x * x or pow(x, 0.5) into sqrt(x): libm's pow is
  // accurate to ~0.52 ulp, not correctly rounded, and the reference gets whatever it returns.
  volatile double two_v = 2.0, half_v = 1.0 / 2.0;
  const double two = two_v, half = half_v;
  for (long long o = 0; o < n_org; ++o) {
    double *r = stats + 5 * o;
    r[0] = (r[1] - r[3]) / std::pow(std::pow(r[2], two) + std::pow(r[4], two), half);
  }
  return FL_OK;
}
