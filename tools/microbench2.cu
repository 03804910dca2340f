// Can the DP cell rate exceed the DPX (ALU pipe) rate by moving the ADD to the FMA pipe (IMAD) and using a
// 3-input max, so that two cells cost one ALU-pipe op?  (run on the B200 via gpurun)
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)
constexpr int ACC = 8;

// variant 0: DPX only. variant 1: IMAD adds + VIMNMX3 (2 cells per max). variant 2: mixed (per 3 cells: 1 DPX, 2 via IMAD+max3)
template <int V>
__global__ void __launch_bounds__(256) k(int *out, const int *in, int iters, int one) {
  int acc[ACC], g0[ACC], g1[ACC], g2[ACC];
#pragma unroll
  for (int a = 0; a < ACC; ++a) { acc[a] = in[threadIdx.x + a * 256]; g0[a] = in[threadIdx.x + 2048 + a * 256]; g1[a] = in[threadIdx.x + 2049 + a * 256]; g2[a] = in[threadIdx.x + 2050 + a * 256]; }
  int p0 = in[4096 + (threadIdx.x & 31)], p1 = in[4097 + (threadIdx.x & 31)], p2 = in[4098 + (threadIdx.x & 31)];
  const int dp = in[4200];
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 2; ++u) {
#pragma unroll
      for (int a = 0; a < ACC; ++a) {
        if (V == 0) {
          acc[a] = __viaddmax_s32(p0, g0[a], acc[a]);
          acc[a] = __viaddmax_s32(p1, g1[a], acc[a]);
          acc[a] = __viaddmax_s32(p2, g2[a], acc[a]);
        } else if (V == 1) {
          const int c0 = p0 * one + g0[a], c1 = p1 * one + g1[a], c2 = p2 * one + g2[a];
          acc[a] = __vimax3_s32(acc[a], c0, c1);
          acc[a] = max(acc[a], c2);
        } else {
          const int c0 = p0 * one + g0[a], c1 = p1 * one + g1[a];
          acc[a] = __vimax3_s32(acc[a], c0, c1);
          acc[a] = __viaddmax_s32(p2, g2[a], acc[a]);
        }
      }
      p0 += dp; p1 += dp; p2 += dp;
    }
  }
  int r = 0;
#pragma unroll
  for (int a = 0; a < ACC; ++a) r += acc[a];
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

template <class K>
static void run(const char *name, K kernel, int *out, const int *in, int blocks, int iters) {
  cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
  kernel<<<blocks, 256>>>(out, in, iters, 1); CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int rep = 0; rep < 5; ++rep) {
    CK(cudaEventRecord(a)); kernel<<<blocks, 256>>>(out, in, iters, 1); CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b));
    float ms; CK(cudaEventElapsedTime(&ms, a, b)); if (ms < best) best = ms;
  }
  double cells = (double)blocks * 256 * 2.0 * ACC * 3 * iters;
  printf("%-36s %8.3f ms %9.3f Gcell/s\n", name, best, cells / best * 1e-6);
}
int main() {
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  const int blocks = prop.multiProcessorCount * 8;
  int *in, *out; CK(cudaMalloc(&in, 8192 * 4)); CK(cudaMalloc(&out, (size_t)blocks * 256 * 4));
  int h[8192]; for (int i = 0; i < 8192; ++i) h[i] = -(i % 97) * 37 - 1;
  CK(cudaMemcpy(in, h, sizeof h, cudaMemcpyHostToDevice));
  run("DPX viaddmax only", k<0>, out, in, blocks, 2048);
  run("IMAD adds + vimax3 (+max)", k<1>, out, in, blocks, 2048);
  run("mixed: 2 cells IMAD+vimax3, 1 DPX", k<2>, out, in, blocks, 2048);
  return 0;
}
