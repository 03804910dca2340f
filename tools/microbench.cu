// Pipe-rate microbenchmarks that steer the DP kernel design (run on the B200 via gpurun).
// Each kernel runs ACC independent accumulator chains per thread of the candidate "cell update"
//   best = max(best, prev + g)
// in a given arithmetic, and reports cell updates per second for the whole chip.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

constexpr int ACC = 8;
constexpr int ITERS = 2048;

template <class Op>
__global__ void __launch_bounds__(256) bench_kernel(typename Op::T *out, const typename Op::T *in, int iters) {
  using T = typename Op::T;
  T acc[ACC], g[ACC];
#pragma unroll
  for (int a = 0; a < ACC; ++a) { acc[a] = in[threadIdx.x + a * 256]; g[a] = in[threadIdx.x + 2048 + a * 256]; }
  T p = in[4096 + (threadIdx.x & 31)];
  T dp = in[4200];
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
#pragma unroll
      for (int a = 0; a < ACC; ++a) acc[a] = Op::cell(acc[a], p, g[a]);
      p = Op::next(p, dp);
    }
  }
  T r = acc[0];
#pragma unroll
  for (int a = 1; a < ACC; ++a) r = Op::combine(r, acc[a]);
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

struct OpF64 {  // DADD + DSETP + 2 FSEL
  using T = double;
  static __device__ __forceinline__ T cell(T best, T p, T g) { T c = p + g; return best < c ? c : best; }
  static __device__ __forceinline__ T next(T p, T d) { return p + d; }
  static __device__ __forceinline__ T combine(T a, T b) { return a + b; }
};
struct OpF64AddOnly {  // DADD only (fp64 pipe rate)
  using T = double;
  static __device__ __forceinline__ T cell(T best, T p, T g) { return best + (p + g); }
  static __device__ __forceinline__ T next(T p, T d) { return p + d; }
  static __device__ __forceinline__ T combine(T a, T b) { return a + b; }
};
struct OpF64HiMax {  // DADD + FMNMX on the high word (filter idea)
  using T = double;
  static __device__ __forceinline__ T cell(T best, T p, T g) {
    T c = p + g;
    float h = fmaxf(__int_as_float(__double2hiint(best)), __int_as_float(__double2hiint(c)));
    return __hiloint2double(__float_as_int(h), __double2loint(best));
  }
  static __device__ __forceinline__ T next(T p, T d) { return p + d; }
  static __device__ __forceinline__ T combine(T a, T b) { return a + b; }
};
struct OpF32 {  // FADD + FMNMX
  using T = float;
  static __device__ __forceinline__ T cell(T best, T p, T g) { return fmaxf(best, p + g); }
  static __device__ __forceinline__ T next(T p, T d) { return p + d; }
  static __device__ __forceinline__ T combine(T a, T b) { return a + b; }
};
struct OpI32 {  // IADD + IMNMX (compiler may fuse into VIADDMNMX)
  using T = int;
  static __device__ __forceinline__ T cell(T best, T p, T g) { return max(best, p + g); }
  static __device__ __forceinline__ T next(T p, T d) { return p + d; }
  static __device__ __forceinline__ T combine(T a, T b) { return a + b; }
};
struct OpDpx32 {  // DPX: max(a + b, c) in one instruction
  using T = int;
  static __device__ __forceinline__ T cell(T best, T p, T g) { return __viaddmax_s32(p, g, best); }
  static __device__ __forceinline__ T next(T p, T d) { return p + d; }
  static __device__ __forceinline__ T combine(T a, T b) { return a + b; }
};
struct OpDpx16x2 {  // DPX packed: two 16-bit cells per instruction
  using T = unsigned;
  static __device__ __forceinline__ T cell(T best, T p, T g) { return __viaddmax_s16x2(p, g, best); }
  static __device__ __forceinline__ T next(T p, T d) { return __vadd2(p, d); }
  static __device__ __forceinline__ T combine(T a, T b) { return a + b; }
};
struct OpI64 {  // 64-bit integer add + max (fixed point alternative)
  using T = long long;
  static __device__ __forceinline__ T cell(T best, T p, T g) { return max(best, p + g); }
  static __device__ __forceinline__ T next(T p, T d) { return p + d; }
  static __device__ __forceinline__ T combine(T a, T b) { return a + b; }
};

// fp32 two cells per step: packed add + 3-input max (sm_100 PTX), if the assembler takes them
struct OpF32x2 {
  using T = float;
  static __device__ __forceinline__ T cell(T best, T p, T g) { return best; }
  static __device__ __forceinline__ T next(T p, T d) { return p + d; }
  static __device__ __forceinline__ T combine(T a, T b) { return a + b; }
};
__global__ void __launch_bounds__(256) bench_f32x2(float *out, const float *in, int iters) {
  float acc[ACC], g0[ACC], g1[ACC];
#pragma unroll
  for (int a = 0; a < ACC; ++a) { acc[a] = in[threadIdx.x + a * 256]; g0[a] = in[threadIdx.x + 2048 + a * 256]; g1[a] = in[threadIdx.x + 2049 + a * 256]; }
  float p0 = in[4096 + (threadIdx.x & 31)], p1 = in[4097 + (threadIdx.x & 31)];
  float dp = in[4200];
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 2; ++u) {
#pragma unroll
      for (int a = 0; a < ACC; ++a) {
        float c0, c1;
        // two cells: (p0+g0, p1+g1) with one packed add, then one 3-input max
        asm volatile("{ .reg .b64 pp, gg, cc; mov.b64 pp, {%2, %3}; mov.b64 gg, {%4, %5}; add.rn.f32x2 cc, pp, gg; mov.b64 {%0, %1}, cc; }"
                     : "=f"(c0), "=f"(c1) : "f"(p0), "f"(p1), "f"(g0[a]), "f"(g1[a]));
        asm volatile("max.f32 %0, %1, %2, %3;" : "=f"(acc[a]) : "f"(acc[a]), "f"(c0), "f"(c1));
      }
      p0 += dp; p1 += dp;
    }
  }
  float r = 0;
#pragma unroll
  for (int a = 0; a < ACC; ++a) r += acc[a];
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

// shared-memory bound variant of the fp64 cell: one LDS.64 of G per cell, prev broadcast (the v1 kernel's loop)
__global__ void __launch_bounds__(256) bench_f64_lds(double *out, const double *in, int iters) {
  __shared__ double G[4096 + 512];
  __shared__ double prev[1024];
  for (int i = threadIdx.x; i < 4096 + 512; i += 256) G[i] = in[i % 4300];
  for (int i = threadIdx.x; i < 1024; i += 256) prev[i] = in[(i * 7) % 4300];
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double acc[4] = {-1e300, -1e300, -1e300, -1e300};
  const double *gp = G + 2048 + lane + warp * 128;
  for (int it = 0; it < iters / 256; ++it) {
#pragma unroll 2
    for (int k = 0; k < 2048; ++k) {
      const double pv = prev[k & 1023];
#pragma unroll
      for (int t = 0; t < 4; ++t) { double c = pv + gp[32 * t - k]; acc[t] = acc[t] < c ? c : acc[t]; }
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc[0] + acc[1] + acc[2] + acc[3];
}

template <class K, class T>
static void run(const char *name, K kernel, T *out, const T *in, int blocks, double cells_per_thread_iter, int iters) {
  cudaEvent_t a, b;
  CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
  kernel<<<blocks, 256>>>(out, in, iters);
  CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int rep = 0; rep < 5; ++rep) {
    CK(cudaEventRecord(a));
    kernel<<<blocks, 256>>>(out, in, iters);
    CK(cudaEventRecord(b));
    CK(cudaEventSynchronize(b));
    float ms; CK(cudaEventElapsedTime(&ms, a, b));
    if (ms < best) best = ms;
  }
  double cells = (double)blocks * 256 * cells_per_thread_iter * iters;
  printf("%-28s %8.3f ms  %9.3f Gcell/s  (%.1f cells/clk/SM at 1.90 GHz, 148 SMs)\n", name, best, cells / best * 1e-6,
         cells / (best * 1e-3) / 148.0 / 1.9e9);
}

int main() {
  int dev = 0; cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, dev));
  printf("device %s, %d SMs, clock %d kHz\n", prop.name, prop.multiProcessorCount, prop.clockRate);
  const int blocks = prop.multiProcessorCount * 8;
  double *din, *dout; CK(cudaMalloc(&din, 8192 * 8)); CK(cudaMalloc(&dout, (size_t)blocks * 256 * 8));
  double h[8192]; for (int i = 0; i < 8192; ++i) h[i] = -(double)(i % 97) * 0.37 - 1.0;
  CK(cudaMemcpy(din, h, sizeof h, cudaMemcpyHostToDevice));
  float *fin, *fout; CK(cudaMalloc(&fin, 8192 * 4)); CK(cudaMalloc(&fout, (size_t)blocks * 256 * 4));
  float hf[8192]; for (int i = 0; i < 8192; ++i) hf[i] = -(float)(i % 97) * 0.37f - 1.0f;
  CK(cudaMemcpy(fin, hf, sizeof hf, cudaMemcpyHostToDevice));
  int *iin, *iout; CK(cudaMalloc(&iin, 8192 * 4)); CK(cudaMalloc(&iout, (size_t)blocks * 256 * 4));
  int hi[8192]; for (int i = 0; i < 8192; ++i) hi[i] = -(i % 97) * 37 - 1;
  CK(cudaMemcpy(iin, hi, sizeof hi, cudaMemcpyHostToDevice));
  long long *lin, *lout; CK(cudaMalloc(&lin, 8192 * 8)); CK(cudaMalloc(&lout, (size_t)blocks * 256 * 8));
  long long hl[8192]; for (int i = 0; i < 8192; ++i) hl[i] = -(long long)(i % 97) * 37000000000LL - 1;
  CK(cudaMemcpy(lin, hl, sizeof hl, cudaMemcpyHostToDevice));

  const double cpi = 4.0 * ACC;
  run("f64 add+max (DADD,DSETP,2SEL)", bench_kernel<OpF64>, dout, din, blocks, cpi, ITERS);
  run("f64 add only (2 DADD/cell)", bench_kernel<OpF64AddOnly>, dout, din, blocks, cpi, ITERS);
  run("f64 add + hi-word FMNMX", bench_kernel<OpF64HiMax>, dout, din, blocks, cpi, ITERS);
  run("f32 add+max (FADD,FMNMX)", bench_kernel<OpF32>, fout, fin, blocks, cpi, ITERS);
  run("f32x2 add + 3-input max", bench_f32x2, fout, fin, blocks, 2.0 * 2 * ACC, ITERS);
  run("i32 add+max", bench_kernel<OpI32>, iout, iin, blocks, cpi, ITERS);
  run("i32 DPX viaddmax", bench_kernel<OpDpx32>, iout, iin, blocks, cpi, ITERS);
  run("i16x2 DPX viaddmax (2 cells)", bench_kernel<OpDpx16x2>, (unsigned *)iout, (const unsigned *)iin, blocks, 2 * cpi, ITERS);
  run("i64 add+max", bench_kernel<OpI64>, lout, lin, blocks, cpi, ITERS);
  run("f64 cell, 1 LDS.64 per cell", bench_f64_lds, dout, din, blocks, 2048.0 * 4 / 256, ITERS);
  return 0;
}
