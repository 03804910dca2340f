// How fast is the register-tiled, folded DP stage by itself (no score passes, no refinement, no prologue)?
// Runs dp_stage<J, TEAM, OpI32> of csrc/place_kernel.cuh back to back on shared-memory rows, 16 warps per SM.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include "../include/flemingo_b200.h"
#include "../flemingo_b200/csrc/place_kernel.cuh"
#define CKK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

template <int J, int TEAM>
__global__ void __launch_bounds__(512) tile_bench(int P, int reps, int row_stride, int slot_stride, int *out) {
  extern __shared__ __align__(16) unsigned char raw[];
  constexpr int NT = 32 / TEAM;
  int *Gq = reinterpret_cast<int *>(raw);                 // kGPad + P + kGPad
  int *rows = Gq + (kGPad + P + kGPad);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int tl = lane % TEAM, team = lane / TEAM;
  const int sent = -40000000;
  for (int i = threadIdx.x; i < kGPad + P + kGPad; i += blockDim.x) {
    const int g = i - kGPad;
    Gq[i] = (g >= 0 && g <= P - 2) ? -((g * 37) % 1000) * 100 : sent;
  }
  int *my = rows + (size_t)(warp * NT + team) * slot_stride;
  for (int j = tl; j < row_stride; j += TEAM) { my[j] = (j < P) ? -((j * 91) % 777) * 50 : 0; my[row_stride + j] = 0; }
  __syncthreads();
  for (int r = 0; r < reps; ++r) {
    const int *prev = my + (r & 1) * row_stride;
    int *cur = my + ((r + 1) & 1) * row_stride;
    dp_stage<J, TEAM, OpI32>(prev, cur, Gq + kGPad, P, tl, sent);
    __syncwarp();
    for (int t = tl; t < J; t += TEAM) cur[P + t] = 0;
    __syncwarp();
  }
  if (tl == 0) out[(blockIdx.x * (blockDim.x >> 5) + warp) * NT + team] = my[(reps & 1) * row_stride + P - 1];
}

template <int J, int TEAM>
void run(int P, int nwarps) {
  const int NT = 32 / TEAM;
  const int row_stride = (P + J + 3) & ~3;
  const int slot_stride = ((2 * row_stride + 31) & ~31) + (TEAM & 31);
  const size_t smem = 4ull * (kGPad + P + kGPad + (size_t)nwarps * NT * slot_stride) + 64;
  CKK(cudaFuncSetAttribute(tile_bench<J, TEAM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int *out; CKK(cudaMalloc(&out, 148 * nwarps * NT * 4));
  const int reps = 400;
  cudaEvent_t a, b; CKK(cudaEventCreate(&a)); CKK(cudaEventCreate(&b));
  tile_bench<J, TEAM><<<148, nwarps * 32, smem>>>(P, 20, row_stride, slot_stride, out); CKK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int i = 0; i < 3; ++i) {
    CKK(cudaEventRecord(a)); tile_bench<J, TEAM><<<148, nwarps * 32, smem>>>(P, reps, row_stride, slot_stride, out); CKK(cudaEventRecord(b));
    CKK(cudaEventSynchronize(b)); float ms; CKK(cudaEventElapsedTime(&ms, a, b)); if (ms < best) best = ms;
  }
  const double cells = 148.0 * nwarps * NT * reps * ((double)P * (P + 1) / 2);
  printf("J=%2d TEAM=%2d P=%d warps/SM=%2d smem=%6zu B : %7.3f ms  %6.2f Tcell/s (useful cells)\n", J, TEAM, P, nwarps, smem, best, cells / best * 1e-9);
  cudaFree(out);
}

// the packed 16-bit stage (dp_stage16): int16 rows, H[d] = (G[d], G[d-1])
template <int J, int TEAM>
__global__ void __launch_bounds__(512) tile_bench16(int P, int reps, int row_stride, int slot_stride, int *out) {
  extern __shared__ __align__(16) unsigned char raw[];
  constexpr int NT = 32 / TEAM;
  unsigned *H = reinterpret_cast<unsigned *>(raw);          // kGPad + P + kGPad words
  short *rows = reinterpret_cast<short *>(H + (kGPad + P + kGPad));
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int tl = lane % TEAM, team = lane / TEAM;
  const int sent = -20000;
  for (int i = threadIdx.x; i < kGPad + P + kGPad; i += blockDim.x) {
    const int g = i - kGPad;
    const int lo = (g >= 0 && g <= P - 2) ? -((g * 37) % 1000) * 4 : sent;
    const int hi = (g - 1 >= 0 && g - 1 <= P - 2) ? -(((g - 1) * 37) % 1000) * 4 : sent;
    H[i] = ((unsigned)lo & 0xffffu) | ((unsigned)hi << 16);
  }
  short *my = rows + (size_t)(warp * NT + team) * slot_stride;
  for (int j = tl; j < row_stride; j += TEAM) { my[j] = (j < P) ? (short)(-((j * 91) % 777) * 4) : 0; my[row_stride + j] = 0; }
  __syncthreads();
  for (int r = 0; r < reps; ++r) {
    const short *prev = my + (r & 1) * row_stride;
    short *cur = my + ((r + 1) & 1) * row_stride;
    dp_stage16<J, TEAM>(prev, cur, H + kGPad, P, tl);
    __syncwarp();
    // keep the values bounded over many repetitions (a real stage is followed by the score pass) and zero the padding
    for (int j = tl; j < P; j += TEAM) cur[j] = (short)(cur[j] / 2);
    for (int t = tl; t < J; t += TEAM) cur[P + t] = 0;
    __syncwarp();
  }
  if (tl == 0) out[(blockIdx.x * (blockDim.x >> 5) + warp) * NT + team] = my[(reps & 1) * row_stride + P - 1];
}

template <int J, int TEAM>
void run16(int P, int nwarps) {
  const int NT = 32 / TEAM;
  const int row_stride = (P + J + 7) & ~7;                                   // shorts
  const int slot_words = ((2 * row_stride / 2 + 31) & ~31) + (TEAM & 31);    // 32-bit words per pair slot (two rows)
  const int slot_stride = 2 * slot_words;
  const size_t smem = 4ull * (kGPad + P + kGPad) + 2ull * (size_t)nwarps * NT * slot_stride + 64;
  CKK(cudaFuncSetAttribute(tile_bench16<J, TEAM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int *out; CKK(cudaMalloc(&out, 148 * nwarps * NT * 4));
  const int reps = 400;
  cudaEvent_t a, b; CKK(cudaEventCreate(&a)); CKK(cudaEventCreate(&b));
  tile_bench16<J, TEAM><<<148, nwarps * 32, smem>>>(P, 20, row_stride, slot_stride, out); CKK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int i = 0; i < 3; ++i) {
    CKK(cudaEventRecord(a)); tile_bench16<J, TEAM><<<148, nwarps * 32, smem>>>(P, reps, row_stride, slot_stride, out); CKK(cudaEventRecord(b));
    CKK(cudaEventSynchronize(b)); float ms; CKK(cudaEventElapsedTime(&ms, a, b)); if (ms < best) best = ms;
  }
  const double cells = 148.0 * nwarps * NT * reps * ((double)P * (P + 1) / 2);
  printf("16x2 J=%2d TEAM=%2d P=%d warps/SM=%2d smem=%6zu B : %7.3f ms  %6.2f Tcell/s (useful cells)\n", J, TEAM, P, nwarps, smem, best, cells / best * 1e-9);
  cudaFree(out);
}

int main() {
  run16<18, 8>(265, 16);
  run16<18, 8>(265, 20);
  run16<18, 8>(265, 24);
  run16<18, 8>(265, 8);
  run16<14, 8>(200, 16);
  run16<22, 8>(340, 16);
  run16<10, 16>(265, 16);
  run16<22, 32>(1965, 8);
  run16<22, 32>(1965, 16);
  run<17, 8>(265, 16);
  run<17, 8>(265, 8);
  run<17, 8>(265, 4);
  run<19, 8>(297, 16);
  run<9, 16>(265, 16);
  run<21, 32>(1965, 8);
  return 0;
}
