#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <algorithm>
#include <limits>
#define __device__
#define __forceinline__ inline
#define __restrict__
using std::min;
static inline double neg_inf() { return -std::numeric_limits<double>::infinity(); }
constexpr int kGPad = 32;
static inline int __viaddmax_s32(int a, int b, int c) { return std::max(a + b, c); }
#include TILE_EXTRACT
template <int J, int TEAM> int run_i32(int P, unsigned seed) {
  srand(seed);
  const int SENT = -50000000;
  int NBJ = ((P + J - 1) / J) * J;
  std::vector<int> prev(NBJ + 32, 0), G(P + 2 * kGPad, SENT), cur(P + 32, -12345), ref(P);
  for (int i = 0; i < P; i++) { prev[i] = -(rand() % 1000000); G[kGPad + i] = -(rand() % 1000000); }
  for (int j = 0; j < P; j++) { int b = SENT; for (int k = 0; k <= j; k++) b = std::max(b, prev[k] + G[kGPad + j - k]); ref[j] = b; }
  for (int tl = 0; tl < TEAM; tl++) dp_stage<J, TEAM, OpI32>(prev.data(), cur.data(), G.data() + kGPad, P, tl, SENT);
  int bad = 0;
  for (int j = 0; j < P; j++) if (cur[j] != ref[j]) bad++;
  return bad;
}
template <int J> int run(int P, unsigned seed) {
  srand(seed);
  int NBJ = ((P + J - 1) / J) * J;
  std::vector<double> prev(NBJ + 16, std::nan("")), G(P + 2 * kGPad, neg_inf()), cur(P + 16, -12345.0), ref(P);
  for (int i = 0; i < P; i++) { prev[i] = -(rand() % 100000) / 997.0; G[kGPad + i] = -(rand() % 100000) / 1013.0; }
  for (int j = 0; j < P; j++) { double b = neg_inf(); for (int k = 0; k <= j; k++) { double c = prev[k] + G[kGPad + j - k]; if (b < c) b = c; } ref[j] = b; }
  // emulate 32 lanes sequentially (each lane independent)
  for (int lane = 0; lane < 32; lane++) dp_stage<J, 32, OpF64>(prev.data(), cur.data(), G.data() + kGPad, P, lane, neg_inf());
  int bad = 0;
  for (int j = 0; j < P; j++) if (cur[j] != ref[j]) bad++;
  return bad;
}
int main() {
  int total = 0;
  for (int P : {1, 2, 3, 5, 31, 32, 33, 64, 65, 100, 191, 192, 193, 265, 289, 319, 320, 321, 447, 449, 576, 577, 700, 1153, 1965}) {
    int b = run<1>(P, P) + run<3>(P, P + 1) + run<5>(P, P + 2) + run<7>(P, P + 3) + run<9>(P, P + 4);
    b += run_i32<17, 8>(P, P + 5) + run_i32<19, 8>(P, P + 6) + run_i32<9, 16>(P, P + 7) + run_i32<13, 32>(P, P + 8) +
         run_i32<3, 8>(P, P + 9) + run_i32<21, 8>(P, P + 10) + run_i32<11, 16>(P, P + 11);
    if (b) printf("P=%d bad=%d\n", P, b);
    total += b;
  }
  printf("total bad %d\n", total);
  return total != 0;
}
