#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <algorithm>
#include <limits>
#define __device__
#define __forceinline__ inline
#define __restrict__
static inline void __syncwarp() {}
using std::min;
static inline double neg_inf() { return -std::numeric_limits<double>::infinity(); }
constexpr int kGPad = 32;
static inline int __viaddmax_s32(int a, int b, int c) { return std::max(a + b, c); }
// host stand-ins for the packed 16-bit intrinsics of dp_stage16
static inline short lo16(unsigned x) { return (short)(x & 0xffffu); }
static inline short hi16(unsigned x) { return (short)(x >> 16); }
static inline unsigned pack16(int lo, int hi) { return ((unsigned)lo & 0xffffu) | ((unsigned)hi << 16); }
static inline unsigned __viaddmax_s16x2(unsigned a, unsigned b, unsigned c) {
  const short l = (short)(lo16(a) + lo16(b)), h = (short)(hi16(a) + hi16(b));  // wrapping add, like the instruction
  return pack16(std::max<short>(l, lo16(c)), std::max<short>(h, hi16(c)));
}
static inline unsigned __vmaxs2(unsigned a, unsigned b) { return pack16(std::max(lo16(a), lo16(b)), std::max(hi16(a), hi16(b))); }
static inline unsigned __byte_perm(unsigned a, unsigned b, unsigned sel) {
  unsigned long long v = ((unsigned long long)b << 32) | a;
  unsigned r = 0;
  for (int i = 0; i < 4; i++) r |= (unsigned)((v >> (8 * ((sel >> (4 * i)) & 7))) & 0xff) << (8 * i);
  return r;
}
#include <type_traits>
static long long g_tiles_run = 0, g_tiles_all = 0;
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

// In-place stage (cur == prev), with the TEAM lanes advanced in LOCKSTEP the way a warp executes them: within one
// step every lane's job switch (flush) comes before every lane's loads.  Mirrors the control flow of dp_stage.
template <int J, int TEAM> int run_i32_inplace(int P, unsigned seed) {
  srand(seed);
  const int SENT = -50000000;
  int NBJ = ((P + J - 1) / J) * J;
  std::vector<int> row(NBJ + 32, 0), G(P + 2 * kGPad, SENT), ref(P);
  for (int i = 0; i < P; i++) { row[i] = -(rand() % 1000000); G[kGPad + i] = -(rand() % 1000000); }
  for (int j = 0; j < P; j++) { int b = SENT; for (int k = 0; k <= j; k++) b = std::max(b, row[k] + G[kGPad + j - k]); ref[j] = b; }
  const int *Gp = G.data() + kGPad;
  int *buf = row.data();
  const int NB = (P + J - 1) / J;
  for (int r0 = ((NB - 1) / (2 * TEAM)) * (2 * TEAM); r0 >= 0; r0 -= 2 * TEAM) {
    const int nbr = std::min(2 * TEAM, NB - r0), half = (nbr + 1) >> 1;
    TileState st[TEAM];
    int acc[TEAM][J], A[TEAM][J], B[TEAM][J];
    for (int tl = 0; tl < TEAM; tl++) {
      const int bA = r0 + tl, bB = r0 + nbr - 1 - tl;
      if (tl < half) { st[tl].b = bB; st[tl].kend = bB + 1; st[tl].kstep = 1; st[tl].next_b = (bB > bA) ? bA : -1; }
      else { st[tl].b = 0; st[tl].kend = -1; st[tl].kstep = 0; st[tl].next_b = -1; }
      st[tl].kb = 0;
      const int *ga = Gp + st[tl].b * J + (J - 1);
      for (int t = 0; t < J; t++) { acc[tl][t] = SENT; A[tl][t] = ga[-t]; }
    }
    const int iters = 2 * r0 + nbr + 1;
    bool a_role = true;  // which array currently plays the A role
    for (int it = 0; it < iters; it++) {
      const bool last = (it == iters - 1);
      for (int tl = 0; tl < TEAM; tl++) {  // phase 1: job switches of every lane
        if (a_role) tile_switch<J, OpI32>(A[tl], acc[tl], st[tl], buf, Gp, SENT);
        else tile_switch<J, OpI32>(B[tl], acc[tl], st[tl], buf, Gp, SENT);
      }
      for (int tl = 0; tl < TEAM; tl++) {  // phase 2: loads and cells (the switch inside is a no-op now)
        if (last) {
          if (a_role) tile_last<J, OpI32>(A[tl], acc[tl], st[tl], buf, buf, Gp, SENT);
          else tile_last<J, OpI32>(B[tl], acc[tl], st[tl], buf, buf, Gp, SENT);
        } else if (a_role) tile_step<J, OpI32>(A[tl], B[tl], acc[tl], st[tl], buf, buf, Gp, P, SENT);
        else tile_step<J, OpI32>(B[tl], A[tl], acc[tl], st[tl], buf, buf, Gp, P, SENT);
      }
      a_role = !a_role;
    }
    for (int tl = 0; tl < TEAM; tl++)
      if (st[tl].kstep) for (int t = 0; t < J; t++) buf[st[tl].b * J + t] = acc[tl][t];
  }
  int bad = 0;
  for (int j = 0; j < P; j++) if (buf[j] != ref[j]) bad++;
  return bad;
}
// Packed 16-bit stage: H[d] = (G[d], G[d-1]), sentinel outside [0, P-2]; values sized so that nothing wraps.
template <int J, int TEAM> int run_i16(int P, unsigned seed, bool inplace) {
  srand(seed);
  const int SENT = -20000;
  int NBJ = ((P + J - 1) / J) * J;
  std::vector<short> prev(NBJ + 64, 0), cur(NBJ + 64, -12345);
  std::vector<int> G(P + 2 * kGPad + 2, SENT), ref(P);
  for (int i = 0; i < P; i++) prev[i] = (short)(-(rand() % 6000));
  for (int i = 0; i + 1 < P; i++) G[kGPad + i] = -(rand() % 6000);   // gap P-1 keeps the sentinel
  for (int j = 0; j < P; j++) { int b = -32768; for (int k = 0; k <= j; k++) b = std::max(b, prev[k] + G[kGPad + j - k]); ref[j] = b; }
  std::vector<unsigned> H(P + 2 * kGPad + 2);
  for (int i = 1; i < (int)H.size(); i++) H[i] = pack16(G[i], G[i - 1]);
  H[0] = pack16(G[0], SENT);
  const unsigned *Hp = H.data() + kGPad;
  if (!inplace) {
    for (int tl = 0; tl < TEAM; tl++) dp_stage16<J, TEAM>(prev.data(), cur.data(), Hp, P, tl);
  } else {
    // lanes in lockstep on ONE buffer, mirroring dp_stage16's control flow (see run_i32_inplace)
    short *buf = prev.data();
    const int NB = (P + J - 1) / J;
    for (int r0 = ((NB - 1) / (2 * TEAM)) * (2 * TEAM); r0 >= 0; r0 -= 2 * TEAM) {
      const int nbr = std::min(2 * TEAM, NB - r0), half = (nbr + 1) >> 1;
      TileState st[TEAM];
      unsigned acc[TEAM][J], A[TEAM][J], B[TEAM][J];
      for (int tl = 0; tl < TEAM; tl++) {
        const int bA = r0 + tl, bB = r0 + nbr - 1 - tl;
        if (tl < half) { st[tl].b = bB; st[tl].kend = bB + 1; st[tl].kstep = 1; st[tl].next_b = (bB > bA) ? bA : -1; }
        else { st[tl].b = 0; st[tl].kend = -1; st[tl].kstep = 0; st[tl].next_b = -1; }
        st[tl].kb = 0;
        const unsigned *ha = Hp + st[tl].b * J + (J - 1);
        for (int t = 0; t < J; t++) { acc[tl][t] = kLowest16x2; A[tl][t] = ha[-t]; }
      }
      const int iters = 2 * r0 + nbr + 1;
      bool a_role = true;
      for (int it = 0; it < iters; it++) {
        const bool last = (it == iters - 1);
        for (int tl = 0; tl < TEAM; tl++) {
          if (a_role) tile_switch16<J>(A[tl], acc[tl], st[tl], buf, Hp);
          else tile_switch16<J>(B[tl], acc[tl], st[tl], buf, Hp);
        }
        for (int tl = 0; tl < TEAM; tl++) {
          if (last) {
            if (a_role) tile_last16<J>(A[tl], acc[tl], st[tl], buf, buf, Hp);
            else tile_last16<J>(B[tl], acc[tl], st[tl], buf, buf, Hp);
          } else if (a_role) tile_step16<J>(A[tl], B[tl], acc[tl], st[tl], buf, buf, Hp);
          else tile_step16<J>(B[tl], A[tl], acc[tl], st[tl], buf, buf, Hp);
        }
        a_role = !a_role;
      }
      for (int tl = 0; tl < TEAM; tl++)
        if (st[tl].kstep) flush16<J>(acc[tl], buf, st[tl].b);
    }
    cur = prev;
  }
  int bad = 0;
  for (int j = 0; j < P; j++) if (cur[j] != ref[j]) bad++;
  return bad;
}

// Pruned stage (dp_stage_pruned): same rows as the naive recurrence whatever the gap profile -- peaked (where the
// pruning bites), flat (where every tile runs) or random -- in place, the lanes of a round flushing together after the
// round the way the kernel does.  The tables gt / gtail / dlo / dhi are built like build_fast_plan step E builds them.
template <int J, int TEAM, bool K16> int run_pruned(int P, unsigned seed, int profile) {
  srand(seed);
  typedef typename std::conditional<K16, short, int>::type RowT;
  typedef typename std::conditional<K16, unsigned, int>::type AccT;
  const int SENT = K16 ? -20000 : -50000000, AMP = K16 ? 6000 : 1000000;
  const int NB = (P + J - 1) / J, NBJ = NB * J;
  std::vector<RowT> row(NBJ + 64, 0);
  std::vector<int> G(P + 2 * kGPad + 2, SENT), ref(P);
  for (int i = 0; i < P; i++) row[i] = (RowT)(-(rand() % AMP));
  const int mu = rand() % (P + 20), width = 1 + rand() % 12;
  for (int i = 0; i + 1 < P; i++) {
    int g;
    if (profile == 0) g = -(int)std::min((double)(AMP - 1), (double)(i - mu) * (i - mu) * AMP / (8.0 * width * width)) - rand() % 3;  // peaked, floor at -AMP
    else if (profile == 1) g = -(rand() % 4);                                                                                // flat
    else g = -(rand() % AMP);                                                                                                // random
    G[kGPad + i] = g;
  }
  for (int j = 0; j < P; j++) { int b = -0x7fffffff; for (int k = 0; k <= j; k++) b = std::max(b, (int)row[k] + G[kGPad + j - k]); ref[j] = b; }
  std::vector<unsigned> H(G.size());
  for (int i = 1; i < (int)H.size(); i++) H[i] = pack16(G[i], G[i - 1]);
  H[0] = pack16(G[0], SENT);
  const int *Gp = K16 ? reinterpret_cast<const int *>(H.data()) + kGPad : G.data() + kGPad;
  // step E of the plan
  std::vector<int> gt(NB), gtail(NB), bm(NB), pm(NB), bm4(NB / 4 + 1, -0x7fffffff - 1);
  for (int d = 0; d < NB; d++) {
    const int lo = std::max(0, (d - 1) * J + 1), hi = std::min(P - 2, (d + 1) * J - 1);
    int m = SENT;
    for (int g = lo; g <= hi; g++) m = std::max(m, G[kGPad + g]);
    gt[d] = m;
  }
  int best = gt[NB - 1], dstar = NB - 1;
  for (int d = NB - 1, run = -0x7fffffff - 1; d >= 0; d--) {
    run = std::max(run, gt[d]);
    gtail[d] = run;
    if (gt[d] >= best) { best = gt[d]; dstar = d; }
  }
  const int drop = AMP / 8;
  PruneTabs tabs;
  tabs.gt = gt.data();
  tabs.gtail = gtail.data();
  tabs.dlo = tabs.dhi = dstar;
  while (tabs.dlo > 0 && gt[tabs.dlo - 1] >= best - drop) --tabs.dlo;
  while (tabs.dhi + 1 < NB && gt[tabs.dhi + 1] >= best - drop) ++tabs.dhi;
  for (int kb = 0, run = -0x7fffffff - 1; kb < NB; kb++) {
    int v = -0x7fffffff - 1;
    for (int u = 0; u < J; u++) if (kb * J + u < P) v = std::max(v, (int)row[kb * J + u]);
    bm[kb] = v;
    run = std::max(run, v);
    pm[kb] = run;
    bm4[kb / 4] = std::max(bm4[kb / 4], v);
  }
  long long tiles = 0;
  RowT *buf = row.data();
  for (int hi = NB - 1; hi >= 0; hi -= TEAM) {
    AccT acc[TEAM][J];
    for (int tl = 0; tl < TEAM; tl++) {
      const int b = hi - tl;
      if (b < 0) continue;
      for (int t = 0; t < J; t++) acc[tl][t] = K16 ? (AccT)kLowest16x2 : (AccT)(-0x7fffffff - 1);
      for (int d = tabs.dlo; d <= tabs.dhi; d++)
        if (d <= b) { prune_run_tile<J>(acc[tl], buf, Gp, b, b - d); tiles++; }
      int minacc = (tabs.dlo <= b) ? prune_row_min<J>(acc[tl], b, P) : -0x7fffffff - 1;
      int kb = b;
      for (;;) {
        kb = prune_next_tile(b, kb, minacc, tabs, bm.data(), pm.data(), (seed & 1) ? bm4.data() : nullptr);
        if (kb < 0) break;
        prune_run_tile<J>(acc[tl], buf, Gp, b, kb);
        tiles++;
        minacc = prune_row_min<J>(acc[tl], b, P);
        --kb;
      }
    }
    for (int tl = 0; tl < TEAM; tl++)
      if (hi - tl >= 0) prune_flush<J>(acc[tl], buf, hi - tl);
  }
  int bad = 0;
  for (int j = 0; j < P; j++) if ((int)buf[j] != ref[j]) bad++;
  g_tiles_run += tiles;
  g_tiles_all += (long long)NB * (NB + 1) / 2;
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
    b += run_i32_inplace<17, 8>(P, P + 12) + run_i32_inplace<19, 8>(P, P + 13) + run_i32_inplace<9, 16>(P, P + 14) +
         run_i32_inplace<13, 32>(P, P + 15) + run_i32_inplace<3, 8>(P, P + 16) + run_i32_inplace<21, 16>(P, P + 17) +
         run_i32_inplace<5, 8>(P, P + 18);
    for (int ip = 0; ip < 2 && P >= 2; ip++)   // P = 1 has no legit gap: the integer pass never sees it (FastPlan.ok)
      b += run_i16<18, 8>(P, P + 20, ip) + run_i16<14, 8>(P, P + 21, ip) + run_i16<10, 16>(P, P + 22, ip) + run_i16<6, 8>(P, P + 23, ip) +
           run_i16<22, 32>(P, P + 24, ip) + run_i16<18, 16>(P, P + 25, ip) + run_i16<2, 8>(P, P + 26, ip);
    for (int prof = 0; prof < 3 && P >= 2; prof++)
      b += run_pruned<9, 8, false>(P, P + 30 + prof, prof) + run_pruned<13, 16, false>(P, P + 33 + prof, prof) + run_pruned<17, 32, false>(P, P + 36 + prof, prof) +
           run_pruned<10, 8, true>(P, P + 39 + prof, prof) + run_pruned<14, 16, true>(P, P + 42 + prof, prof) + run_pruned<18, 32, true>(P, P + 45 + prof, prof) +
           run_pruned<3, 8, false>(P, P + 48 + prof, prof) + run_pruned<2, 8, true>(P, P + 51 + prof, prof);
    if (b) printf("P=%d bad=%d\n", P, b);
    total += b;
  }
  printf("pruned stages ran %lld of %lld tiles\n", g_tiles_run, g_tiles_all);
  printf("total bad %d\n", total);
  return total != 0;
}
