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
constexpr int kGPad = 16;
#include TILE_EXTRACT
template <int J> int run(int P, unsigned seed) {
  srand(seed);
  int NBJ = ((P + J - 1) / J) * J;
  std::vector<double> prev(NBJ + 16, std::nan("")), G(P + 2 * kGPad, neg_inf()), cur(P + 16, -12345.0), ref(P);
  for (int i = 0; i < P; i++) { prev[i] = -(rand() % 100000) / 997.0; G[kGPad + i] = -(rand() % 100000) / 1013.0; }
  for (int j = 0; j < P; j++) { double b = neg_inf(); for (int k = 0; k <= j; k++) { double c = prev[k] + G[kGPad + j - k]; if (b < c) b = c; } ref[j] = b; }
  // emulate 32 lanes sequentially (each lane independent)
  for (int lane = 0; lane < 32; lane++) dp_stage<J>(prev.data(), cur.data(), G.data() + kGPad, P, lane);
  int bad = 0;
  for (int j = 0; j < P; j++) if (cur[j] != ref[j]) bad++;
  return bad;
}
int main() {
  int total = 0;
  for (int P : {1, 2, 3, 5, 31, 32, 33, 64, 65, 100, 191, 192, 193, 265, 289, 319, 320, 321, 447, 449, 576, 577, 700, 1153, 1965}) {
    int b = run<1>(P, P) + run<3>(P, P + 1) + run<5>(P, P + 2) + run<7>(P, P + 3) + run<9>(P, P + 4);
    if (b) printf("P=%d bad=%d\n", P, b);
    total += b;
  }
  printf("total bad %d\n", total);
  return total != 0;
}
