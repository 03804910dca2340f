// flemingo_b200.cu -- sm_100a kernels + C ABI (include/flemingo_b200.h) for FLEMINGO's placement hot path.
//
// What the kernels compute is specified by the reference C extension (paths relative to
// /root/reference/extensions/multiplacement/src):
//   K1  recognizer score rows      organism/recognizer.c:130-168 (PSSM), :187-589 (shape), _aux.c:184-217
//   K2  gap-energy table + chain DP organism/connector.c:31-66, organism/organism.c:321-393
//       argmax + traceback          _aux.c:149-158, organism/organism.c:398-427
//   K3  fitness statistic           src/objects/organism_object.py:848-956 (reference does this in numpy)
//
// Arithmetic contract: every value is an IEEE double produced by the same operations in the same operand
// order as the reference ((prev[k] + G[j-k]) + R[j], left-to-right score sums, IEEE division), compiled with
// -fmad=false, so energies, node scores and offsets are bit-identical, including the reference's tie rule
// (smallest k per stage, smallest j at the end).  Two identities make that cheap:
//   (1) x -> fl(x + r) is monotone, so  max_k fl(fl(prev[k]+G[j-k]) + R[j]) == fl(max_k fl(prev[k]+G[j-k]) + R[j]):
//       the inner loop needs one add and one max per cell;
//   (2) the winning k is only needed on the optimal path, so the forward pass keeps no argmax; the traceback
//       re-scans one column per stage for the FIRST k that reproduces the stored value bit for bit.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "flemingo_b200.h"

// ------------------------------------------------------------------------------------------------
// error plumbing
// ------------------------------------------------------------------------------------------------
static thread_local std::string g_last_error;

static int fail(int code, const char *fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_last_error = buf;
  return code;
}

#define CK(call)                                                                               \
  do {                                                                                         \
    cudaError_t e_ = (call);                                                                   \
    if (e_ != cudaSuccess) return fail(FL_E_CUDA, "%s failed: %s", #call, cudaGetErrorString(e_)); \
  } while (0)

#include "place_kernel.cuh"
#include "place_fast_kernel.cuh"
#include "shape_bins_kernel.cuh"
#include "fitness_kernel.cuh"

// ------------------------------------------------------------------------------------------------
// host side: context, uploads, launch planning
// ------------------------------------------------------------------------------------------------
typedef void (*place_kernel_fn)(const PlaceParams);
constexpr int kNumTileSizes = 5;
static const int kTileSizes[kNumTileSizes] = {1, 3, 5, 7, 9};  // odd: lane stride J doubles is bank-conflict free
static const place_kernel_fn kPlaceKernels[kNumTileSizes][2] = {
    {place_tiled_kernel<1, false>, place_tiled_kernel<1, true>}, {place_tiled_kernel<3, false>, place_tiled_kernel<3, true>},
    {place_tiled_kernel<5, false>, place_tiled_kernel<5, true>}, {place_tiled_kernel<7, false>, place_tiled_kernel<7, true>},
    {place_tiled_kernel<9, false>, place_tiled_kernel<9, true>}};

typedef void (*list_kernel_fn)(const PlaceParams, const ListBuckets, const int *, const int *, const int *);
static const list_kernel_fn kListKernels[kNumTileSizes][2] = {
    {place_list_kernel<1, false>, place_list_kernel<1, true>}, {place_list_kernel<3, false>, place_list_kernel<3, true>},
    {place_list_kernel<5, false>, place_list_kernel<5, true>}, {place_list_kernel<7, false>, place_list_kernel<7, true>},
    {place_list_kernel<9, false>, place_list_kernel<9, true>}};

// Tile size for P offsets: the smallest J whose single folded round (64 blocks) covers P, so that most lanes are busy.
static int pick_tile(int P) {
  for (int j = 0; j < kNumTileSizes; ++j)
    if (64 * kTileSizes[j] >= P) return j;
  return kNumTileSizes - 1;
}

// Integer fast path: (tile size J, lanes per pair TEAM, arithmetic).  A team covers 2*TEAM*J offsets per folded round.
// bits = 32: one DPX instruction per cell, fits every organism; bits = 16: packed rows, two cells per instruction, for
// organisms whose whole value range fits int16 at a useful scale (place_fast_kernel.cuh); J = 2 * odd there (bank layout).
typedef void (*fast_kernel_fn)(const FastParams);
struct FastConfig {
  int J, team, bits, prune;
  fast_kernel_fn fn;
};
#define FAST_CFG(J, T) {J, T, 32, 0, place_fast_kernel<J, T, 32>}
#define FAST_CFG16(J, T) {J, T, 16, 0, place_fast_kernel<J, T, 16>}
// kernels with pruned stages (dp_stage_pruned) for organisms whose connectors are tight: small tiles, so that the tiles
// around the peak of the gap energies are few cells; one configuration per range of P (4 rounds of TEAM row blocks)
#define FAST_CFGP(J, T) {J, T, 32, 1, place_fast_kernel<J, T, 32, true>}
#define FAST_CFGP16(J, T) {J, T, 16, 1, place_fast_kernel<J, T, 16, true>}
static const FastConfig kFastConfigs[] = {
    FAST_CFG(3, 8),  FAST_CFG(5, 8),  FAST_CFG(7, 8),  FAST_CFG(9, 8),  FAST_CFG(11, 8), FAST_CFG(13, 8), FAST_CFG(15, 8),
    FAST_CFG(17, 8), FAST_CFG(19, 8), FAST_CFG(21, 8), FAST_CFG(9, 16), FAST_CFG(11, 16), FAST_CFG(13, 16), FAST_CFG(15, 16), FAST_CFG(17, 16), FAST_CFG(19, 16),
    FAST_CFG(21, 16), FAST_CFG(13, 32), FAST_CFG(17, 32), FAST_CFG(21, 32),
    FAST_CFG16(6, 8), FAST_CFG16(10, 8), FAST_CFG16(14, 8), FAST_CFG16(18, 8), FAST_CFG16(22, 8), FAST_CFG16(10, 16), FAST_CFG16(14, 16),
    FAST_CFG16(18, 16), FAST_CFG16(22, 16), FAST_CFG16(14, 32), FAST_CFG16(18, 32), FAST_CFG16(22, 32),
    FAST_CFGP(9, 8), FAST_CFGP(13, 16), FAST_CFGP(17, 32), FAST_CFGP16(10, 8), FAST_CFGP16(14, 16), FAST_CFGP16(18, 32)};
constexpr int kNumFastConfigs = (int)(sizeof(kFastConfigs) / sizeof(kFastConfigs[0]));

// Estimated issue slots per pair: tiles x (cells + loads and bookkeeping), scaled by the share of the warp a team uses.
static double fast_cost(int P, int J, int team, int bits) {
  const int NB = (P + J - 1) / J;
  const double full = bits == 16 ? J * J / 2.0 + 2.5 * J + 24.0 : J * J + 3.0 * J + 24.0;
  const double diag = bits == 16 ? J * (J + 2) / 4.0 + 1.5 * J + 24.0 : J * (J + 1) / 2 + 2.0 * J + 24.0;
  double slots = 0;
  for (int r0 = 0; r0 < NB; r0 += 2 * team) {
    const int nbr = std::min(2 * team, NB - r0);
    const int iters = 2 * r0 + nbr + 1;  // dp_stage: iters - 1 full tiles, then one diagonal tile
    slots += (iters - 1) * full + diag;
  }
  return slots * team / 32.0;
}
static int pick_fast(int P, int bits, int prune = 0) {
  int best = -1;
  double best_cost = 1e300;
  if (prune) {  // the first configuration whose four rounds of TEAM row blocks cover P (the list is sorted by TEAM)
    for (int i = 0; i < kNumFastConfigs; ++i)
      if (kFastConfigs[i].prune && kFastConfigs[i].bits == bits) {
        best = i;
        if (4 * kFastConfigs[i].team * kFastConfigs[i].J >= P) break;
      }
    return best;
  }
  // experiments: FLEMINGO_FAST_CFG="J,TEAM" forces one configuration (of the arithmetic that has it)
  static const char *force = getenv("FLEMINGO_FAST_CFG");
  if (force) {
    int fj = 0, ft = 0;
    if (sscanf(force, "%d,%d", &fj, &ft) == 2)
      for (int i = 0; i < kNumFastConfigs; ++i)
        if (kFastConfigs[i].J == fj && kFastConfigs[i].team == ft && kFastConfigs[i].bits == bits && !kFastConfigs[i].prune) return i;
  }
  // the choice depends on (P, bits) only: a ragged dataset asks for every organism of every length class
  static std::vector<int> cache[2];
  std::vector<int> &memo = cache[bits == 16 ? 1 : 0];
  if (P >= 0 && P < (int)memo.size() && memo[P] >= 0) return memo[P];
  for (int i = 0; i < kNumFastConfigs; ++i) {
    if (kFastConfigs[i].bits != bits || kFastConfigs[i].prune) continue;
    const double c = fast_cost(P, kFastConfigs[i].J, kFastConfigs[i].team, bits);
    if (c < best_cost) {
      best_cost = c;
      best = i;
    }
  }
  if (P >= 0 && P < (1 << 20)) {
    if (P >= (int)memo.size()) memo.resize(P + 1, -1);
    memo[P] = best;
  }
  return best;
}
struct DevBuf {
  void *p = nullptr;
  size_t cap = 0;
  // `cap` is what callers may ask for; every allocation carries kSlack more bytes that are never handed out, so
  // the word-sized over-reads of the kernels (up to 16 bytes past the last element) stay inside the allocation
  // even when a later request fills the capacity exactly.
  static constexpr size_t kSlack = 256;
  int ensure(size_t bytes) {
    if (bytes <= cap && p) return FL_OK;
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
    const size_t want = bytes + bytes / 4;
    CK(cudaMalloc(&p, want + kSlack));
    cap = want;
    return FL_OK;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
  }
  template <class T>
  T *as() const { return reinterpret_cast<T *>(p); }
};

struct SeqClass {
  int len, begin, count;
  float low_share = 0.f;  // share of low-complexity sequences in a sample of the class (see low_complexity)
};

struct SeqSet {
  bool valid = false;
  int n_seq = 0;
  std::vector<SeqClass> classes;
  DevBuf packed, unk, word_off, unk_off, order, has_unk, d_len, bins, raw, d_offsets, k8, k8_off;
  int max_len = 0;
  unsigned long long bins_signature = 0;  // shape-class signature the bins were computed for (0 = none)
  long long bins_row_stride = 0;
  DevBuf E;  // energies [n_org][n_seq]
  int E_n_org = 0;
  bool E_valid = false;
};

struct Population {
  bool valid = false;
  int n_org = 0, n_rec_total = 0, max_rec = 0;
  std::vector<int> rec_begin, sum_len, con_M, pool_len;
  int n_classes = 0, max_class_len = 0;
  unsigned long long class_signature = 0;
  std::vector<int> kmer_need;  // ints of the integer pass's tables if the organism has >= 2 recognizers, else -1
  // upper bounds on the value ranges of the recognizer tables (energy units), for the host's choice between the 16-bit
  // and the 32-bit integer pass
  std::vector<double> rec_range;
  std::vector<int> n_units;    // rounded tables per organism (4-mer groups + shape recognizers): error budget
  std::vector<char> pssm_only; // every recognizer of the organism is a PSSM
  // routing to the kernels with pruned stages: sigma of connector c of organism o at [rec_begin[o] + c], read from the
  // block headers at the canonical pool layout (base tables in organism order, con_base_guess[o]); NaN if unknown
  std::vector<float> con_sigma;
  std::vector<long long> con_base_guess;
  long long n_con = 0;
  int max_pool = 0, min_sum_len = 0;
  DevBuf d_rec_begin, d_rec_type, d_rec_len, d_rec_nb, d_rec_data, d_pool_begin, d_rec_pool, d_con_M, d_con_pool, d_sum_len, d_rec_class, d_cls_kind, d_cls_len, d_cls_nb, d_cls_edges, d_edges_pool;
};

struct fl_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  DevBuf tables, err_flag, con_begin, org_list, gaps, rsc, csc, fit_out, fit_scratch, fb_total, fb_retried;
  // A dataset with several sequence lengths is placed class by class (one launch chain per length); the chains of
  // different classes are independent, so they go round-robin onto kLanes streams -- the main stream and kLanes - 1
  // helpers -- each with its own hand-back lists, row scratch and work counter.  A single-length dataset uses lane 0 only.
  static constexpr int kLanes = 4;
  struct Lane {
    cudaStream_t stream = nullptr;
    cudaEvent_t done = nullptr;
    DevBuf fb_item_count, fb_item_seqs, fb_unit_begin, fast_scratch, work_counter;
  } lanes[kLanes];
  cudaEvent_t ready = nullptr;
  DevBuf single_arena;             // fl_calculate: device arena and pinned staging buffer of the one-placement path
  void *single_stage = nullptr;
  size_t single_stage_cap = 0;
  int fast_reg_warps[64] = {0};
  std::vector<double> h_den;  // host copy of DEN_EXPANSION (fits_int16)
  long long fast_pairs = 0, fallback_pairs = 0;  // statistics of the last fl_place_batch
  Population pop, scratch_pop;
  SeqSet sets[FL_MAX_SETS + 1];  // the extra slot serves fl_calculate
  long long launches = 0;
};

extern "C" const char *fl_last_error(void) { return g_last_error.c_str(); }
extern "C" int fl_abi_version(void) { return 2; }

extern "C" int fl_device_count(void) {
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return count;
}

extern "C" int fl_ctx_create(int device, const double *tables, long long n_tables, fl_ctx **out) {
  if (!out || !tables) return fail(FL_E_ARG, "fl_ctx_create: null argument");
  if (n_tables != FL_N_TABLE_DOUBLES) return fail(FL_E_ARG, "fl_ctx_create: expected %d table doubles, got %lld", FL_N_TABLE_DOUBLES, n_tables);
  int count = 0;
  CK(cudaGetDeviceCount(&count));
  if (device < 0 || device >= count) return fail(FL_E_ARG, "fl_ctx_create: device %d not present (%d devices)", device, count);
  CK(cudaSetDevice(device));
  fl_ctx *ctx = new fl_ctx();
  ctx->device = device;
  ctx->h_den.assign(tables + T_DEN, tables + T_DEN + 10000);
  cudaError_t e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking);
  if (e != cudaSuccess) {
    delete ctx;
    return fail(FL_E_CUDA, "cudaStreamCreate failed: %s", cudaGetErrorString(e));
  }
  int rc = ctx->tables.ensure(sizeof(double) * FL_N_TABLE_DOUBLES);
  if (rc == FL_OK) rc = ctx->err_flag.ensure(sizeof(int));
  if (rc == FL_OK) rc = ctx->fb_total.ensure(sizeof(unsigned long long));
  if (rc == FL_OK) rc = ctx->fb_retried.ensure(sizeof(unsigned long long));
  if (rc == FL_OK && cudaMemsetAsync(ctx->fb_retried.p, 0, sizeof(unsigned long long), ctx->stream) != cudaSuccess) rc = fail(FL_E_CUDA, "memset failed");
  for (int l = 0; l < fl_ctx::kLanes && rc == FL_OK; ++l) {
    fl_ctx::Lane &ln = ctx->lanes[l];
    if (l == 0)
      ln.stream = ctx->stream;
    else if (cudaStreamCreateWithFlags(&ln.stream, cudaStreamNonBlocking) != cudaSuccess)
      rc = fail(FL_E_CUDA, "cudaStreamCreate failed");
    if (rc == FL_OK && cudaEventCreateWithFlags(&ln.done, cudaEventDisableTiming) != cudaSuccess) rc = fail(FL_E_CUDA, "cudaEventCreate failed");
    if (rc == FL_OK) rc = ln.work_counter.ensure(sizeof(int));
  }
  if (rc == FL_OK && cudaEventCreateWithFlags(&ctx->ready, cudaEventDisableTiming) != cudaSuccess) rc = fail(FL_E_CUDA, "cudaEventCreate failed");
  if (rc == FL_OK && cudaMemsetAsync(ctx->fb_total.p, 0, sizeof(unsigned long long), ctx->stream) != cudaSuccess) rc = fail(FL_E_CUDA, "memset failed");
  if (rc != FL_OK) {
    delete ctx;
    return rc;
  }
  CK(cudaMemcpyAsync(ctx->tables.p, tables, sizeof(double) * FL_N_TABLE_DOUBLES, cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemsetAsync(ctx->err_flag.p, 0, sizeof(int), ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  for (int j = 0; j < kNumTileSizes; ++j) {
    CK(cudaFuncSetAttribute(kPlaceKernels[j][0], cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem));
    CK(cudaFuncSetAttribute(kPlaceKernels[j][1], cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem));
  }
  for (int j = 0; j < kNumFastConfigs; ++j) {
    CK(cudaFuncSetAttribute(kFastConfigs[j].fn, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem));
    // resident warps per SM the register file allows (64 K registers, allocated per warp in units of 8 per thread)
    cudaFuncAttributes fa;
    CK(cudaFuncGetAttributes(&fa, kFastConfigs[j].fn));
    const int regs = std::max(8, (fa.numRegs + 7) & ~7);
    ctx->fast_reg_warps[j] = std::max(1, 65536 / (regs * 32));
  }
  for (int j = 0; j < kNumTileSizes; ++j) {
    CK(cudaFuncSetAttribute(kListKernels[j][0], cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem));
    CK(cudaFuncSetAttribute(kListKernels[j][1], cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem));
  }
  *out = ctx;
  return FL_OK;
}

static void release_pop(Population &p) {
  DevBuf *bufs[] = {&p.d_rec_begin, &p.d_rec_type, &p.d_rec_len, &p.d_rec_nb, &p.d_rec_data,
                    &p.d_pool_begin, &p.d_rec_pool, &p.d_con_M, &p.d_con_pool, &p.d_sum_len,
                    &p.d_rec_class, &p.d_cls_kind, &p.d_cls_len, &p.d_cls_nb, &p.d_cls_edges, &p.d_edges_pool};
  for (DevBuf *b : bufs) b->release();
}

extern "C" int fl_ctx_destroy(fl_ctx *ctx) {
  if (!ctx) return FL_OK;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  for (SeqSet &s : ctx->sets) {
    s.packed.release(); s.unk.release(); s.word_off.release(); s.unk_off.release(); s.order.release(); s.has_unk.release(); s.d_len.release(); s.bins.release(); s.raw.release(); s.d_offsets.release(); s.k8.release(); s.k8_off.release(); s.E.release();
  }
  release_pop(ctx->pop);
  release_pop(ctx->scratch_pop);
  DevBuf *bufs[] = {&ctx->tables, &ctx->err_flag, &ctx->con_begin, &ctx->org_list, &ctx->gaps, &ctx->rsc, &ctx->csc, &ctx->fit_out, &ctx->fit_scratch, &ctx->fb_total, &ctx->fb_retried};
  for (DevBuf *b : bufs) b->release();
  for (fl_ctx::Lane &ln : ctx->lanes) {
    cudaStreamSynchronize(ln.stream);
    ln.fb_item_count.release(); ln.fb_item_seqs.release(); ln.fb_unit_begin.release(); ln.fast_scratch.release(); ln.work_counter.release();
    if (ln.done) cudaEventDestroy(ln.done);
    if (ln.stream && ln.stream != ctx->stream) cudaStreamDestroy(ln.stream);
  }
  if (ctx->ready) cudaEventDestroy(ctx->ready);
  ctx->single_arena.release();
  if (ctx->single_stage) cudaFreeHost(ctx->single_stage);
  cudaStreamDestroy(ctx->stream);
  delete ctx;
  return FL_OK;
}

extern "C" void *fl_ctx_stream(fl_ctx *ctx) { return ctx ? (void *)ctx->stream : nullptr; }
extern "C" long long fl_launch_count(fl_ctx *ctx) { return ctx ? ctx->launches : 0; }

static int check_device_errors(fl_ctx *ctx) {
  int flag = 0;
  CK(cudaMemcpyAsync(&flag, ctx->err_flag.p, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  if (flag) {
    CK(cudaMemsetAsync(ctx->err_flag.p, 0, sizeof(int), ctx->stream));
    if (flag & ERR_NAN_NULL)
      return fail(FL_E_NAN_NULL, "a shape score fell into a null-model bin whose frequency is NaN (the reference exits here, recognizer.c:258-262)");
  }
  return FL_OK;
}

extern "C" int fl_sync(fl_ctx *ctx) {
  if (!ctx) return fail(FL_E_ARG, "fl_sync: null context");
  CK(cudaSetDevice(ctx->device));
  CK(cudaStreamSynchronize(ctx->stream));
  return check_device_errors(ctx);
}

// An asynchronous copy FROM PAGEABLE host memory returns only after the bytes have been staged for DMA, so the source
// may be reused as soon as the call returns; a copy from pinned memory really is asynchronous.  Uploads therefore
// synchronise only when a caller hands over pinned buffers.
static bool is_pinned(const void *p) {
  if (!p) return false;
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
    cudaGetLastError();
    return true;  // unknown: be conservative
  }
  return a.type != cudaMemoryTypeUnregistered;
}

// Low-complexity sequence (2-letter alphabets, short-period repeats): fewer than 0.4 of the distinct 4-mers a random
// sequence of the same length would show.  On such sequences the recognizer scores are periodic, every k block holds a
// good predecessor and the pruned stages (dp_stage_pruned) prune nothing while paying for their small tiles, so a length
// class with many of them keeps the regular kernels.  A routing heuristic only: results do not depend on it.
static bool low_complexity(const char *s, int L) {
  unsigned long long seen[4] = {0, 0, 0, 0};
  int code = 0, valid = 0, distinct = 0, total = 0;
  for (int i = 0; i < L; ++i) {
    int b;
    switch (s[i]) {
      case 'a': case 'A': b = 0; break;
      case 'g': case 'G': b = 1; break;
      case 'c': case 'C': b = 2; break;
      case 't': case 'T': b = 3; break;
      default: b = -1;
    }
    if (b < 0) {
      valid = 0;
      continue;
    }
    code = ((code << 2) | b) & 255;
    if (++valid >= 4) {
      ++total;
      const unsigned long long bit = 1ULL << (code & 63);
      if (!(seen[code >> 6] & bit)) {
        seen[code >> 6] |= bit;
        ++distinct;
      }
    }
  }
  if (total < 12) return false;
  return distinct < 0.4 * 256.0 * (1.0 - std::exp(-total / 256.0));
}

static int upload_sequences(fl_ctx *ctx, SeqSet &set, const char *bytes, const long long *offsets, int n_seq) {
  if (n_seq < 0 || (n_seq > 0 && (!bytes || !offsets))) return fail(FL_E_ARG, "fl_upload_sequences: bad arguments");
  std::vector<int> len(n_seq), word_off(n_seq), unk_off(n_seq), order(n_seq), k8_off(n_seq);
  long long words = 0, uwords = 0, k8_units = 0;
  bool one_length = true;
  for (int s = 0; s < n_seq; ++s) {
    const long long l = offsets[s + 1] - offsets[s];
    if (l < 0 || l > 100000000) return fail(FL_E_ARG, "fl_upload_sequences: sequence %d has length %lld", s, l);
    len[s] = (int)l;
    word_off[s] = (int)words;
    unk_off[s] = (int)uwords;
    k8_off[s] = (int)k8_units;
    words += (l + 15) / 16;
    uwords += (l + 31) / 32;
    k8_units += (l + kK8Pad + 15) / 16;
    if (words > 0x7fffffffLL || k8_units > 0x7fffffffLL) return fail(FL_E_ARG, "fl_upload_sequences: set too large");
    order[s] = s;
    one_length = one_length && (len[s] == len[0]);
  }
  if (!one_length) std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return len[a] < len[b]; });
  set.classes.clear();
  for (int i = 0; i < n_seq; ++i) {
    const int l = len[order[i]];
    if (set.classes.empty() || set.classes.back().len != l) set.classes.push_back({l, i, 0});
    set.classes.back().count++;
  }
  for (SeqClass &cl : set.classes) {  // up to 64 sequences per class, evenly spread
    const int n_sample = std::min(cl.count, 64);
    int low = 0;
    for (int k = 0; k < n_sample; ++k) {
      const int sq = order[cl.begin + (int)((long long)k * cl.count / n_sample)];
      if (low_complexity(bytes + offsets[sq], len[sq])) ++low;
    }
    cl.low_share = (float)low / (float)n_sample;
  }
  set.n_seq = n_seq;
  const size_t raw_bytes = n_seq ? (size_t)(offsets[n_seq] - offsets[0]) : 0;
  int rc;
  if ((rc = set.packed.ensure(((size_t)words + 2) * 4)) || (rc = set.unk.ensure(((size_t)uwords + 2) * 4)) ||
      (rc = set.word_off.ensure((size_t)n_seq * 4 + 4)) || (rc = set.unk_off.ensure((size_t)n_seq * 4 + 4)) ||
      (rc = set.order.ensure((size_t)n_seq * 4 + 4)) || (rc = set.has_unk.ensure((size_t)n_seq + 4)) ||
      (rc = set.d_len.ensure((size_t)n_seq * 4 + 4)) || (rc = set.raw.ensure(raw_bytes + 16)) ||
      (rc = set.d_offsets.ensure(((size_t)n_seq + 1) * 8)) || (rc = set.k8.ensure((size_t)k8_units * 16 + 16)) ||
      (rc = set.k8_off.ensure((size_t)n_seq * 4 + 4)))
    return rc;
  if (n_seq) {
    // raw bytes go up as they are; the 2-bit packing runs on the device (pack_sequences_kernel)
    std::vector<long long> rel((size_t)n_seq + 1);
    for (int s = 0; s <= n_seq; ++s) rel[s] = offsets[s] - offsets[0];
    CK(cudaMemcpyAsync(set.raw.p, bytes + offsets[0], raw_bytes, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(set.d_offsets.p, rel.data(), ((size_t)n_seq + 1) * 8, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(set.word_off.p, word_off.data(), (size_t)n_seq * 4, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(set.unk_off.p, unk_off.data(), (size_t)n_seq * 4, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(set.order.p, order.data(), (size_t)n_seq * 4, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(set.d_len.p, len.data(), (size_t)n_seq * 4, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(set.k8_off.p, k8_off.data(), (size_t)n_seq * 4, cudaMemcpyHostToDevice, ctx->stream));
    pack_sequences_kernel<<<(unsigned)n_seq, 64, 0, ctx->stream>>>(set.raw.as<unsigned char>(), set.d_offsets.as<long long>(),
                                                                    set.word_off.as<int>(), set.unk_off.as<int>(),
                                                                    set.packed.as<unsigned>(), set.unk.as<unsigned>(),
                                                                    set.has_unk.as<unsigned char>(), set.k8_off.as<int>(),
                                                                    set.k8.as<unsigned char>());
    CK(cudaGetLastError());
    ctx->launches++;
    if (is_pinned(bytes)) CK(cudaStreamSynchronize(ctx->stream));  // the staging vectors are pageable (see is_pinned)
  }
  set.max_len = 0;
  for (int s2 = 0; s2 < n_seq; ++s2) set.max_len = std::max(set.max_len, len[s2]);
  set.bins_signature = 0;
  set.valid = true;
  set.E_valid = false;
  return FL_OK;
}

extern "C" int fl_upload_sequences(fl_ctx *ctx, int set_id, const char *bytes, const long long *offsets, int n_seq) {
  if (!ctx) return fail(FL_E_ARG, "fl_upload_sequences: null context");
  if (set_id < 0 || set_id >= FL_MAX_SETS) return fail(FL_E_ARG, "fl_upload_sequences: set_id %d out of range [0,%d)", set_id, FL_MAX_SETS);
  CK(cudaSetDevice(ctx->device));
  return upload_sequences(ctx, ctx->sets[set_id], bytes, offsets, n_seq);
}

extern "C" int fl_sequence_classes(fl_ctx *ctx, int set_id, int *n_classes, int *class_len) {
  if (!ctx || set_id < 0 || set_id >= FL_MAX_SETS || !ctx->sets[set_id].valid) return fail(FL_E_STATE, "fl_sequence_classes: no such set");
  const SeqSet &s = ctx->sets[set_id];
  if (n_classes) *n_classes = (int)s.classes.size();
  if (class_len)
    for (size_t i = 0; i < s.classes.size(); ++i) class_len[i] = s.classes[i].len;
  return FL_OK;
}

static int upload_population(fl_ctx *ctx, Population &P, const fl_population *pop) {
  if (!pop || pop->n_org < 0) return fail(FL_E_ARG, "fl_upload_population: bad population");
  const int n_org = pop->n_org;
  P.valid = false;
  P.n_org = n_org;
  P.n_rec_total = pop->n_rec_total;
  P.rec_begin.assign(pop->rec_begin, pop->rec_begin + n_org + 1);
  P.sum_len.assign(n_org, 0);
  P.con_M.assign(pop->con_M, pop->con_M + n_org);
  P.pool_len.assign(n_org, 0);
  P.kmer_need.assign(n_org, -1);
  P.rec_range.assign((size_t)pop->n_rec_total, 0.0);
  P.n_units.assign(n_org, 0);
  P.pssm_only.assign(n_org, 1);
  P.con_sigma.assign((size_t)pop->n_rec_total, NAN);
  P.con_base_guess.assign(n_org, -1);
  long long canon = 0;
  P.max_rec = 0;
  P.max_pool = 0;
  P.min_sum_len = 0x7fffffff;
  P.n_con = pop->n_con;
  if (n_org && (P.rec_begin[0] != 0 || P.rec_begin[n_org] != pop->n_rec_total)) return fail(FL_E_ARG, "fl_upload_population: rec_begin does not cover n_rec_total");
  for (int o = 0; o < n_org; ++o) {
    const int rb = P.rec_begin[o], n = P.rec_begin[o + 1] - rb;
    if (n < 1 || n > FL_MAX_RECOGNIZERS) return fail(FL_E_ARG, "organism %d has %d recognizers (supported: 1..%d)", o, n, FL_MAX_RECOGNIZERS);
    int sl = 0;
    for (int i = 0; i < n; ++i) {
      const char t = pop->rec_type[rb + i];
      const int L = pop->rec_len[rb + i];
      const bool is_p = (t == 'p' || t == 'P');
      const bool is_s = (t == 'm' || t == 'M' || t == 't' || t == 'T' || t == 'r' || t == 'R' || t == 'h' || t == 'H');
      if (!is_p && !is_s) return fail(FL_E_ARG, "organism %d recognizer %d: unknown type '%c'", o, i, t);
      if (L < 1 || (is_s && L < 5)) return fail(FL_E_ARG, "organism %d recognizer %d: length %d not allowed for type '%c'", o, i, L, t);
      if (is_s && pop->rec_nb[rb + i] < 2) return fail(FL_E_ARG, "organism %d recognizer %d: shape model needs >= 2 bin edges", o, i);
      const long long need = is_p ? 4LL * L : 2LL * pop->rec_nb[rb + i] - 2;
      const long long off = pop->rec_data[rb + i];
      if (off < pop->pool_begin[o] || off + need > pop->pool_begin[o + 1]) return fail(FL_E_ARG, "organism %d recognizer %d: data outside the organism's pool slice", o, i);
      sl += L;
    }
    P.sum_len[o] = sl;
    P.pool_len[o] = (int)(pop->pool_begin[o + 1] - pop->pool_begin[o]);
    if (n > 1 && P.con_M[o] >= 1) {
      const long long stride = 2LL * P.con_M[o] + 2;
      if (pop->con_pool && canon + (n - 1) * stride <= pop->n_con) {
        P.con_base_guess[o] = canon;
        for (int c = 0; c < n - 1; ++c) P.con_sigma[rb + c] = (float)pop->con_pool[canon + c * stride + 1];
      }
      canon += (n - 1) * stride;
    }
    {
      // ints the integer fast path needs for this organism: 4-mer tables per PSSM, one bin table per shape recognizer
      int need = 0, units = 0;
      for (int i = 0; i < n; ++i) {
        const char t = pop->rec_type[rb + i];
        const bool is_p = (t == 'p' || t == 'P');
        const int L = pop->rec_len[rb + i];
        need += is_p ? 256 * ((L + 3) / 4) : pop->rec_nb[rb + i] - 1;
        units += is_p ? (L + 3) / 4 : 1;
        if (!is_p) P.pssm_only[o] = 0;
        if (n >= 2) {
          // value range of the recognizer's scores (what build_fast_plan computes on the device)
          const double *m = pop->rec_pool + pop->rec_data[rb + i];
          double range = 0.0;
          if (is_p) {
            for (int c = 0; c < L; ++c) {
              const double lo = std::min(std::min(m[4 * c], m[4 * c + 1]), std::min(m[4 * c + 2], m[4 * c + 3]));
              const double hi = std::max(std::max(m[4 * c], m[4 * c + 1]), std::max(m[4 * c + 2], m[4 * c + 3]));
              range += hi - lo;
            }
          } else {
            const int nbins = pop->rec_nb[rb + i] - 1;
            double lo = INFINITY, hi = -INFINITY;
            for (int b = 0; b < nbins; ++b) {
              double x = m[nbins + b] - m[b];
              if (x != x) continue;
              if (x < -1e10) x = -1e10;
              lo = std::min(lo, x);
              hi = std::max(hi, x);
            }
            range = hi - lo;
          }
          P.rec_range[rb + i] = (range == range) ? range : INFINITY;
        }
      }
      // a single recognizer stays on the exact kernel: without a connector the energy takes few distinct values (one
      // of 25 bin scores for a shape recognizer, the score of a short k-mer for a PSSM), offsets tie exactly all the
      // time and the candidate lists overflow (measured: 3.5 % of config 3's pairs handed back, 165 ms instead of 162)
      P.kmer_need[o] = n >= 2 ? need : -1;
      P.n_units[o] = units;
    }
    P.max_rec = std::max(P.max_rec, n);
    P.max_pool = std::max(P.max_pool, P.pool_len[o]);
    P.min_sum_len = std::min(P.min_sum_len, sl);
    if (n > 1 && P.con_M[o] < 1) return fail(FL_E_UNSUPPORTED, "organism %d: connectors without precomputed tables (legacy branch connector.c:70-72) are not supported", o);
  }
  // shape classes
  P.n_classes = pop->n_classes;
  P.max_class_len = 0;
  unsigned long long sig = 1469598103934665603ULL;
  auto mix = [&sig](const void *data, size_t bytes) {
    const unsigned char *b = static_cast<const unsigned char *>(data);
    for (size_t i = 0; i < bytes; ++i) sig = (sig ^ b[i]) * 1099511628211ULL;
  };
  if (pop->n_classes < 0) return fail(FL_E_ARG, "fl_upload_population: negative n_classes");
  for (int c = 0; c < pop->n_classes; ++c) {
    const int nb = pop->cls_nb[c], L = pop->cls_len[c];
    if (nb < 2 || nb > 256) return fail(FL_E_UNSUPPORTED, "shape class %d has %d bin edges (supported: 2..256)", c, nb);
    if (L < 5) return fail(FL_E_ARG, "shape class %d has length %d", c, L);
    if (pop->cls_edges[c] < 0 || pop->cls_edges[c] + nb > pop->n_edges) return fail(FL_E_ARG, "shape class %d: edges outside edges_pool", c);
    P.max_class_len = std::max(P.max_class_len, L);
    mix(&pop->cls_kind[c], 1);
    mix(&L, sizeof L);
    mix(&nb, sizeof nb);
    mix(pop->edges_pool + pop->cls_edges[c], sizeof(double) * nb);
  }
  P.class_signature = pop->n_classes ? (sig | 1ULL) : 0ULL;
  for (int r = 0; r < pop->n_rec_total; ++r) {
    const char t = pop->rec_type[r];
    const bool is_p = (t == 'p' || t == 'P');
    const int cls = pop->rec_class ? pop->rec_class[r] : -1;
    if (is_p) continue;
    if (cls < 0 || cls >= pop->n_classes) return fail(FL_E_ARG, "recognizer %d: shape class %d out of range", r, cls);
    if (pop->cls_len[cls] != pop->rec_len[r] || pop->cls_nb[cls] != pop->rec_nb[r])
      return fail(FL_E_ARG, "recognizer %d does not match its shape class (length / number of edges)", r);
  }
  int rc;
  const size_t nr = (size_t)pop->n_rec_total;
  const size_t ncl = (size_t)pop->n_classes;
  if ((rc = P.d_rec_class.ensure(nr * 4 + 4)) || (rc = P.d_cls_kind.ensure(ncl + 4)) || (rc = P.d_cls_len.ensure(ncl * 4 + 4)) ||
      (rc = P.d_cls_nb.ensure(ncl * 4 + 4)) || (rc = P.d_cls_edges.ensure(ncl * 8 + 8)) || (rc = P.d_edges_pool.ensure((size_t)pop->n_edges * 8 + 8)))
    return rc;
  {
    std::vector<int> cls_of(nr, -1);
    if (pop->rec_class) cls_of.assign(pop->rec_class, pop->rec_class + nr);
    if (nr) CK(cudaMemcpyAsync(P.d_rec_class.p, cls_of.data(), nr * 4, cudaMemcpyHostToDevice, ctx->stream));
    if (ncl) {
      CK(cudaMemcpyAsync(P.d_cls_kind.p, pop->cls_kind, ncl, cudaMemcpyHostToDevice, ctx->stream));
      CK(cudaMemcpyAsync(P.d_cls_len.p, pop->cls_len, ncl * 4, cudaMemcpyHostToDevice, ctx->stream));
      CK(cudaMemcpyAsync(P.d_cls_nb.p, pop->cls_nb, ncl * 4, cudaMemcpyHostToDevice, ctx->stream));
      CK(cudaMemcpyAsync(P.d_cls_edges.p, pop->cls_edges, ncl * 8, cudaMemcpyHostToDevice, ctx->stream));
      CK(cudaMemcpyAsync(P.d_edges_pool.p, pop->edges_pool, (size_t)pop->n_edges * 8, cudaMemcpyHostToDevice, ctx->stream));
    }
  }
  if ((rc = P.d_rec_begin.ensure((n_org + 1) * 4)) || (rc = P.d_rec_type.ensure(nr + 4)) || (rc = P.d_rec_len.ensure(nr * 4 + 4)) ||
      (rc = P.d_rec_nb.ensure(nr * 4 + 4)) || (rc = P.d_rec_data.ensure(nr * 8 + 8)) || (rc = P.d_pool_begin.ensure((n_org + 1) * 8)) ||
      (rc = P.d_rec_pool.ensure((size_t)pop->n_pool * 8 + 8)) || (rc = P.d_con_M.ensure((size_t)n_org * 4 + 4)) ||
      (rc = P.d_con_pool.ensure((size_t)pop->n_con * 8 + 8)) || (rc = P.d_sum_len.ensure((size_t)n_org * 4 + 4)))
    return rc;
  cudaStream_t st = ctx->stream;
  CK(cudaMemcpyAsync(P.d_rec_begin.p, pop->rec_begin, (n_org + 1) * 4, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(P.d_rec_type.p, pop->rec_type, nr, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(P.d_rec_len.p, pop->rec_len, nr * 4, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(P.d_rec_nb.p, pop->rec_nb, nr * 4, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(P.d_rec_data.p, pop->rec_data, nr * 8, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(P.d_pool_begin.p, pop->pool_begin, (n_org + 1) * 8, cudaMemcpyHostToDevice, st));
  if (pop->n_pool) CK(cudaMemcpyAsync(P.d_rec_pool.p, pop->rec_pool, (size_t)pop->n_pool * 8, cudaMemcpyHostToDevice, st));
  if (n_org) CK(cudaMemcpyAsync(P.d_con_M.p, pop->con_M, (size_t)n_org * 4, cudaMemcpyHostToDevice, st));
  if (pop->n_con) CK(cudaMemcpyAsync(P.d_con_pool.p, pop->con_pool, (size_t)pop->n_con * 8, cudaMemcpyHostToDevice, st));
  if (n_org) CK(cudaMemcpyAsync(P.d_sum_len.p, P.sum_len.data(), (size_t)n_org * 4, cudaMemcpyHostToDevice, st));
  // the caller may free its arrays as soon as we return: copies from pageable memory are staged before
  // cudaMemcpyAsync returns, copies from pinned memory are not -- rather than classify sixteen pointers, wait for
  // the (small) uploads
  CK(cudaStreamSynchronize(st));
  P.valid = true;
  return FL_OK;
}

extern "C" int fl_upload_population(fl_ctx *ctx, const fl_population *pop) {
  if (!ctx) return fail(FL_E_ARG, "fl_upload_population: null context");
  CK(cudaSetDevice(ctx->device));
  for (int s = 0; s < FL_MAX_SETS; ++s) ctx->sets[s].E_valid = false;
  return upload_population(ctx, ctx->pop, pop);
}

extern "C" int fl_update_connectors(fl_ctx *ctx, const double *con_pool, long long n_con) {
  if (!ctx || (n_con > 0 && !con_pool)) return fail(FL_E_ARG, "fl_update_connectors: null argument");
  if (!ctx->pop.valid) return fail(FL_E_STATE, "fl_update_connectors: no population uploaded");
  if (n_con < ctx->pop.n_con) return fail(FL_E_ARG, "fl_update_connectors: the pool may only grow (%lld -> %lld doubles)", ctx->pop.n_con, n_con);
  CK(cudaSetDevice(ctx->device));
  CK(cudaStreamSynchronize(ctx->stream));  // kernels in flight still read the old pool
  int rc = ctx->pop.d_con_pool.ensure((size_t)n_con * 8 + 8);
  if (rc) return rc;
  if (n_con) CK(cudaMemcpyAsync(ctx->pop.d_con_pool.p, con_pool, (size_t)n_con * 8, cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  ctx->pop.n_con = n_con;
  return FL_OK;
}

// Which organisms take the packed 16-bit integer pass.  PSSM-only organisms whose value range fits: their scores are
// diverse (a 2-decimal PSSM entry per column), so even at 32-128 units per energy unit a stage keeps one or two
// candidate cells.  Organisms with shape recognizers stay on the int32 pass: a shape score takes one of 25 bin values
// and neighbouring offsets share bins, so near-ties are the rule, the 16-bit thresholds let 2-18 % of their pairs
// overflow the candidate lists (measured on config 3 per organism class), and their refinement costs more than the DP
// saves (config 3: 32 ms per launch on the packed pass against ~16 ms on int32 for the same items).
// Host-side estimate of build_fast_plan<16>'s scale for organism o on sequences of length S, from UPPER bounds of the
// table ranges: recognizers from the uploaded tables, a connector's gap energies from  pdf in [log2(1e-15), 0]  plus
// the spread of the combinatorial term den(g) (connector.c:51-56).  True when the packed 16-bit integer pass will
// accept the organism at a scale of at least 32 units per energy unit (the device re-derives the plan from the real
// ranges and hands the organism's pairs to the exact kernel should it disagree).
static bool fits_int16(fl_ctx *ctx, const Population &P, int o, int S, const char *force) {
  if (force) return atoi(force) == 16;
  const int rb = P.rec_begin[o], n = P.rec_begin[o + 1] - rb;
  static const bool shapes_too = getenv("FLEMINGO_16_SHAPES") != nullptr;   // experiment knob (see the comment above)
  if (n < 2 || (!P.pssm_only[o] && !shapes_too)) return false;
  const int Pn = S - P.sum_len[o] + 1, eff = S - P.sum_len[o] + n;
  if (Pn < 2 || eff > 10000) return false;
  // spread of the combinatorial term: den(g) = log C(eff - g - 1, n - 1) - log C(eff, n) falls monotonically with g, so its
  // range over the legit gaps 0 .. Pn-2 is den(0) - den(Pn - 2)
  const double *D = ctx->h_den.data();
  auto den_at = [&](int g) {
    return (D[(eff - (g + 1)) - 1] - D[(n - 1) - 1] - D[eff - (g + 1) - (n - 1) - 1]) - (D[eff - 1] - D[n - 1] - D[eff - n - 1]);
  };
  const double den_range = std::fabs(den_at(0) - den_at(Pn - 2));
  const double gr = 50.0 + den_range + 0.02;   // DEN_EXPANSION is rounded to 0.01: monotone up to that
  double cum = P.rec_range[rb], need = 0.0;
  for (int c = 0; c < n - 1; ++c) {
    need = std::max(need, 2.0 * cum + gr);
    cum += gr + P.rec_range[rb + c + 1];
  }
  need = std::max(need, cum);
  const double budget = (double)kBudget16 - 4.0 * (P.n_units[o] + n) - 16.0;
  return need == need && need > 0.0 && budget / need >= 32.0;
}

// Which organisms take a kernel with pruned stages (dp_stage_pruned): those whose estimated stage costs there -- a tight
// connector's stage pruned, the others in full at the pruned kernel's smaller tile -- stay below 0.97 of the full stages
// of the regular configuration.  The estimate uses sigma only (a Gaussian is within `drop` bits of its peak for
// |g - mu| <= sigma sqrt(2 ln2 drop)) and fast_cost's unit, issue slots per pair: the peak tiles the plan would choose
// times the cost of a fresh-loaded tile, plus the per-block bookkeeping (thresholds 0.9 / 0.97 / 1.05 measured: config 3
// 150.1 ms each, config 4 279 / 259 / 259 ms).  The kernel decides stage by stage from the real tables, so a wrong guess
// costs time, never correctness.
// mode: 0 never, 1 by estimate, 2 every organism (tests).
static bool wants_prune(const Population &P, int o, int Po, int bits, long long cb, int mode, int drop, int frac) {
  if (mode == 0) return false;
  const int rb = P.rec_begin[o], n = P.rec_begin[o + 1] - rb;
  if (n < 2) return false;
  const int j = pick_fast(Po, bits, 1), j0 = pick_fast(Po, bits, 0);
  if (j < 0 || j0 < 0) return false;
  if (mode == 2) return true;
  if (P.con_base_guess[o] < 0 || cb != P.con_base_guess[o]) return false;   // not the tables whose headers were read
  const int J = kFastConfigs[j].J, team = kFastConfigs[j].team, NB = (Po + J - 1) / J;
  const double tile = bits == 16 ? J * J / 2.0 + 2.5 * J + 24.0 : J * J + 3.0 * J + 24.0;
  const double full_here = fast_cost(Po, J, team, bits);
  const double full_regular = fast_cost(Po, kFastConfigs[j0].J, kFastConfigs[j0].team, bits);
  const int rounds = (NB + team - 1) / team;
  double cost = 0.0;
  for (int c = 0; c < n - 1; ++c) {
    const double sigma = P.con_sigma[rb + c];
    double stage = full_here;
    if (sigma >= 0.0) {
      // the kernel's peak window: `drop` units, halved while it spans more than six tiles (build_fast_plan step E)
      int tiles = 0;
      for (int d = drop;; d >>= 1) {
        const double width = 2.0 * sigma * std::sqrt(1.3863 * d) + 1.0;
        tiles = (int)(width / J) + 2;
        if (tiles - 1 <= 6 || d <= 2) break;
      }
      if ((long long)(tiles + 2) * frac <= NB)
        stage = std::min(stage, rounds * (tiles * tile + 150.0 + 2.0 * J) * team / 32.0);
    }
    cost += stage;
  }
  return cost <= 0.97 * (n - 1) * full_regular;
}

static int place_batch(fl_ctx *ctx, Population &P, SeqSet &set, const long long *con_begin, int flags, double *energies,
                       int *gaps, double *rec_scores, double *con_scores) {
  if (!P.valid) return fail(FL_E_STATE, "fl_place_batch: no population uploaded");
  if (!set.valid) return fail(FL_E_STATE, "fl_place_batch: sequence set not uploaded");
  const bool trace = (flags & FL_WANT_TRACE) != 0;
  if (!trace && (gaps || rec_scores || con_scores)) return fail(FL_E_ARG, "fl_place_batch: gaps/scores need FL_WANT_TRACE");
  const int n_org = P.n_org, n_seq = set.n_seq, n_cls = (int)set.classes.size();
  set.E_valid = false;
  if (n_org == 0 || n_seq == 0) {
    set.E_n_org = n_org;
    set.E_valid = true;
    return FL_OK;
  }
  if (!con_begin) return fail(FL_E_ARG, "fl_place_batch: con_begin is required");
  // ---- admissibility (organism_object.py:1279, organism.c:262,336, _constants.h:11) ---------------
  for (int c = 0; c < n_cls; ++c) {
    const int S = set.classes[c].len;
    for (int o = 0; o < n_org; ++o) {
      const int n = P.rec_begin[o + 1] - P.rec_begin[o];
      if (P.sum_len[o] > S)
        return fail(FL_E_TOO_LARGE, "organism %d (sum of recognizer lengths %d) is too large to be placed on a sequence of length %d", o, P.sum_len[o], S);
      if (n > 1) {
        if (S - P.sum_len[o] > P.con_M[o] - 1)
          return fail(FL_E_UNSUPPORTED, "organism %d: gap tables cover %d lengths but a sequence of length %d needs %d (legacy non-precomputed branch, organism.c:336-338)",
                      o, P.con_M[o], S, S - P.sum_len[o] + 1);
        const long long cb = con_begin[(size_t)c * n_org + o];
        if (cb < 0 || cb + (long long)(n - 1) * (2LL * P.con_M[o] + 2) > P.n_con)
          return fail(FL_E_ARG, "con_begin[class %d][organism %d] points outside con_pool", c, o);
        // only organisms with a connector read DEN_EXPANSION (connector.c:51-56); a single recognizer is placed on
        // any sequence length, as in the reference
        if (S - P.sum_len[o] + n > 10000) return fail(FL_E_UNSUPPORTED, "effective length %d exceeds DEN_EXPANSION (10000 entries, _constants.h:11)", S - P.sum_len[o] + n);
      }
    }
  }
  int rc;
  if ((rc = set.E.ensure((size_t)n_org * n_seq * 8))) return rc;
  if ((rc = ctx->con_begin.ensure((size_t)n_cls * n_org * 8))) return rc;
  if ((rc = ctx->org_list.ensure((size_t)n_cls * n_org * 4))) return rc;
  if (trace) {
    const size_t cells = (size_t)n_org * n_seq * P.max_rec;
    if ((rc = ctx->gaps.ensure(cells * 4)) || (rc = ctx->rsc.ensure(cells * 8)) || (rc = ctx->csc.ensure(cells * 8))) return rc;
    CK(cudaMemsetAsync(ctx->gaps.p, 0, cells * 4, ctx->stream));
    CK(cudaMemsetAsync(ctx->rsc.p, 0, cells * 8, ctx->stream));
    CK(cudaMemsetAsync(ctx->csc.p, 0, cells * 8, ctx->stream));
  }
  CK(cudaMemcpyAsync(ctx->con_begin.p, con_begin, (size_t)n_cls * n_org * 8, cudaMemcpyHostToDevice, ctx->stream));
  if (is_pinned(con_begin)) CK(cudaStreamSynchronize(ctx->stream));  // a pinned source is read asynchronously

  int n_sm = 148;
  cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, ctx->device);

  // ---- shape bins of this dataset for the population's shape classes (once per dataset / class set) ----------
  if (P.n_classes > 0 && set.bins_signature != P.class_signature) {
    set.bins_row_stride = (set.max_len + 15) & ~15LL;
    if ((rc = set.bins.ensure((size_t)P.n_classes * n_seq * (size_t)set.bins_row_stride))) return rc;
    ShapeBinParams q;
    q.packed = set.packed.as<unsigned>();
    q.word_off = set.word_off.as<int>();
    q.seq_len = set.d_len.as<int>();
    q.n_seq = n_seq;
    q.cls_kind = P.d_cls_kind.as<char>();
    q.cls_len = P.d_cls_len.as<int>();
    q.cls_nb = P.d_cls_nb.as<int>();
    q.cls_edges = P.d_cls_edges.as<long long>();
    q.edges_pool = P.d_edges_pool.as<double>();
    q.tables = ctx->tables.as<double>();
    q.bins = set.bins.as<unsigned char>();
    q.row_stride = set.bins_row_stride;
    for (int c0 = 0; c0 < P.n_classes; c0 += 65535) {  // gridDim.y limit
      ShapeBinParams qq = q;
      qq.cls_kind += c0; qq.cls_len += c0; qq.cls_nb += c0; qq.cls_edges += c0;
      qq.bins += (size_t)c0 * n_seq * (size_t)set.bins_row_stride;
      const int ny = std::min(65535, P.n_classes - c0);
      shape_bins_kernel<<<dim3((unsigned)n_seq, (unsigned)ny), 128, 0, ctx->stream>>>(qq);
      CK(cudaGetLastError());
      ctx->launches++;
    }
    set.bins_signature = P.class_signature;
  }

  const char *env_bits = getenv("FLEMINGO_FAST_BITS");   // experiments / tests: force one arithmetic (read at every call)
  // pruned stages: FLEMINGO_PRUNE = 0 (never) / 1 (organisms with tight connectors, the default) / 2 (every organism);
  // FLEMINGO_PRUNE_DROP, FLEMINGO_PRUNE_FRAC: see FastParams
  const char *env_prune = getenv("FLEMINGO_PRUNE");
  const int prune_mode = env_prune ? atoi(env_prune) : 1;
  const int prune_drop = getenv("FLEMINGO_PRUNE_DROP") ? std::max(1, atoi(getenv("FLEMINGO_PRUNE_DROP"))) : 8;
  const int prune_frac = getenv("FLEMINGO_PRUNE_FRAC") ? std::max(1, atoi(getenv("FLEMINGO_PRUNE_FRAC"))) : 2;
  // several length classes: their launch chains are independent and go round-robin onto the lanes
  const int n_lanes = n_cls > 1 ? fl_ctx::kLanes : 1;
  if (n_lanes > 1) {
    CK(cudaEventRecord(ctx->ready, ctx->stream));
    for (int l = 1; l < n_lanes; ++l) CK(cudaStreamWaitEvent(ctx->lanes[l].stream, ctx->ready, 0));
  }
  for (int c = 0; c < n_cls; ++c) {
    const SeqClass &cl = set.classes[c];
    fl_ctx::Lane &lane = ctx->lanes[c % n_lanes];
    const cudaStream_t st = lane.stream;
    PlaceParams pp;
    pp.rec_begin = P.d_rec_begin.as<int>();
    pp.rec_type = P.d_rec_type.as<char>();
    pp.rec_len = P.d_rec_len.as<int>();
    pp.rec_nb = P.d_rec_nb.as<int>();
    pp.rec_data = P.d_rec_data.as<long long>();
    pp.pool_begin = P.d_pool_begin.as<long long>();
    pp.rec_pool = P.d_rec_pool.as<double>();
    pp.con_M = P.d_con_M.as<int>();
    pp.con_pool = P.d_con_pool.as<double>();
    pp.con_begin = ctx->con_begin.as<long long>() + (size_t)c * n_org;
    pp.sum_len = P.d_sum_len.as<int>();
    pp.packed = set.packed.as<unsigned>();
    pp.unk = set.unk.as<unsigned>();
    pp.word_off = set.word_off.as<int>();
    pp.k8g = set.k8.as<unsigned char>();
    pp.k8_off = set.k8_off.as<int>();
    pp.unk_off = set.unk_off.as<int>();
    pp.order = set.order.as<int>();
    pp.class_begin = cl.begin;
    pp.class_count = cl.count;
    pp.S = cl.len;
    pp.n_seq_total = n_seq;
    pp.tables = ctx->tables.as<double>();
    pp.rec_class = P.d_rec_class.as<int>();
    pp.bins = set.bins.as<unsigned char>();
    pp.bins_row_stride = set.bins_row_stride;
    pp.E = set.E.as<double>();
    pp.gaps = ctx->gaps.as<int>();
    pp.rsc = ctx->rsc.as<double>();
    pp.csc = ctx->csc.as<double>();
    pp.max_rec = P.max_rec;
    pp.err_flag = ctx->err_flag.as<int>();

    const int Pmax = cl.len - P.min_sum_len + 1;
    pp.pool_stride = (P.max_pool + 1) & ~1;
    pp.g_stride = (kGPad + Pmax + kGPad + 1) & ~1;
    pp.row_stride = (Pmax + kMaxJ + 3) & ~3;
    pp.code_stride = (cl.len + 15) & ~15;
    pp.nrows = trace ? P.max_rec : std::min(2, P.max_rec);
    int nwarps = 8;
    size_t smem = 0;
    for (;; nwarps >>= 1) {
      smem = 8ull * ((size_t)pp.pool_stride + (size_t)(P.max_rec - 1) * pp.g_stride + (size_t)nwarps * pp.nrows * pp.row_stride) +
             sizeof(RecDesc) * P.max_rec + (size_t)nwarps * pp.code_stride;
      // prefer configurations that leave room for >= 2 resident CTAs unless that costs warps
      if (smem <= (size_t)kMaxSmem && (nwarps <= 4 || smem <= (size_t)kMaxSmem / 2)) break;
      if (nwarps == 1) break;
    }
    if (smem > (size_t)kMaxSmem)
      return fail(FL_E_UNSUPPORTED, "placement of %d offsets x %d recognizers needs %zu bytes of shared memory (limit %d)", Pmax, P.max_rec, smem, kMaxSmem);
    // ---- buckets: integer fast path (PSSM-only organisms, P >= 2) by (J, TEAM); everything else exact by J ----
    const bool allow_fast = !(flags & FL_EXACT_ONLY);
    FastParams fp;
    fp.gq_stride = kGPad + Pmax + kGPad;   // per configuration below (the pruned kernels append their tile bounds)
    fp.gt_off = fp.nb_max = fp.bm_stride = 0;
    fp.prune_drop = prune_drop;
    fp.prune_frac = prune_frac;
    fp.k8_stride = (cl.len + kK8Pad + 15) & ~15;  // the sequence's block of 4-mer codes, as laid out by pack_sequences_kernel
    fp.want_trace = trace ? 1 : 0;
    fp.has_unk = set.has_unk.as<unsigned char>();
    int kmer_cap = 0;
    std::vector<int> exact_bucket[kNumTileSizes], fast_bucket[kNumFastConfigs];
    long long n_fast_org = 0;
    // by estimate: not on a length class with many low-complexity sequences (see low_complexity)
    const int prune_mode_c = (prune_mode == 1 && cl.low_share > 0.15f) ? 0 : prune_mode;
    const bool small_class = n_cls > 1 && cl.count < 128 && prune_mode != 2 && !getenv("FLEMINGO_FAST_CFG");
    for (int o = 0; o < n_org; ++o) {
      const int Po = cl.len - P.sum_len[o] + 1;
      if (allow_fast && P.kmer_need[o] > 0 && P.kmer_need[o] <= 16384 && Po >= 2) {
        const int bits = fits_int16(ctx, P, o, cl.len, env_bits) ? 16 : 32;
        if (small_class) {
          // a class of a few sequences is bound by its launches, not by its DP cells: ONE configuration per arithmetic
          // (the one of the largest P; a smaller P only wastes part of a tile) instead of one launch per (J, TEAM)
          fast_bucket[pick_fast(Pmax, bits, 0)].push_back(o);
        } else {
          const int prune = wants_prune(P, o, Po, bits, con_begin[(size_t)c * n_org + o], prune_mode_c, prune_drop, prune_frac) ? 1 : 0;
          fast_bucket[pick_fast(Po, bits, prune)].push_back(o);
        }
        kmer_cap = std::max(kmer_cap, P.kmer_need[o]);
        ++n_fast_org;
      } else {
        exact_bucket[pick_tile(Po)].push_back(o);
      }
    }
    // Splitting a length class's organisms of one arithmetic over a regular and a pruned launch pays only when both are
    // large enough to fill the machine eight times over; otherwise the launch tails cost more than the pruned stages
    // save and the smaller bucket joins the larger one (the pruned kernels run any organism: they decide per stage).
    if (prune_mode_c == 1) {
      const long long min_items = 8LL * n_sm;
      auto items_of = [&](int j) { return (long long)fast_bucket[j].size() * ((cl.count + 255) / 256); };
      for (int j = 0; j < kNumFastConfigs; ++j) {
        if (!kFastConfigs[j].prune || fast_bucket[j].empty()) continue;
        // the regular buckets of this arithmetic (one per (J, TEAM) picked for the organisms' P)
        long long regular = 0;
        for (int k = 0; k < kNumFastConfigs; ++k)
          if (!kFastConfigs[k].prune && kFastConfigs[k].bits == kFastConfigs[j].bits) regular += items_of(k);
        if (regular == 0) continue;
        if (items_of(j) < min_items && items_of(j) <= regular) {
          for (int o : fast_bucket[j]) fast_bucket[pick_fast(cl.len - P.sum_len[o] + 1, kFastConfigs[j].bits, 0)].push_back(o);
          fast_bucket[j].clear();
        } else if (regular < min_items) {
          for (int k = 0; k < kNumFastConfigs; ++k) {
            if (kFastConfigs[k].prune || kFastConfigs[k].bits != kFastConfigs[j].bits) continue;
            for (int o : fast_bucket[k]) fast_bucket[pick_fast(cl.len - P.sum_len[o] + 1, kFastConfigs[k].bits, 1)].push_back(o);
            fast_bucket[k].clear();
          }
        }
      }
    }
    fp.kmer_cap = kmer_cap;
    // fast-path geometry: warps per CTA chosen for the most resident warps per SM; if even one warp does not fit,
    // those organisms go to the exact kernel
    int fast_warps[kNumFastConfigs], fast_rowq[kNumFastConfigs], fast_slot[kNumFastConfigs], fast_ctas_per_sm[kNumFastConfigs];
    int fast_gq[kNumFastConfigs], fast_nb[kNumFastConfigs], fast_bm[kNumFastConfigs];
    bool fast_single[kNumFastConfigs] = {false}, fast_planned[kNumFastConfigs] = {false};
    size_t fast_smem[kNumFastConfigs];
    for (int j = 0; j < kNumFastConfigs; ++j) {
      fast_warps[j] = 0;
      fast_smem[j] = 0;
    }
    // launch geometry of configuration j for this class (once); false if not even one warp fits
    auto plan_config = [&](int j) -> bool {
      if (fast_planned[j]) return fast_warps[j] > 0;
      fast_planned[j] = true;
      const int team = kFastConfigs[j].team, nt = 32 / team;
      const int elem = kFastConfigs[j].bits / 8;                  // bytes per row element
      const int align = 16 / elem;                                // row elements per 16 bytes (cp.async / vector stores)
      fast_rowq[j] = (Pmax + kFastConfigs[j].J + align - 1) & ~(align - 1);
      // pruned kernels: gt + gtail behind every gap table, bmax + pmax per pair slot (odd stride: teams on distinct banks)
      fast_nb[j] = kFastConfigs[j].prune ? (Pmax + kFastConfigs[j].J - 1) / kFastConfigs[j].J : 0;
      fast_gq[j] = kGPad + Pmax + kGPad + 2 * fast_nb[j];
      fast_bm[j] = kFastConfigs[j].prune ? ((2 * fast_nb[j] + fast_nb[j] / 4 + 2) | 1) : 0;
      int best_occ = 0;
      // multiples of four only: the DP stages are bound by the ALU pipe of each of the SM's four schedulers, so a
      // scheduler that holds one warp more than the others sets the pace (13 warps: 194 ms on config 3, 12 warps: 169 ms)
      static const int kWarpChoices[] = {16, 12, 8, 4, 2, 1};
      const int max_warps = 16;
      const char *force_w = getenv("FLEMINGO_FAST_WARPS");
      // two quantised rows per pair (ping-pong; the refinement prefetches older rows behind its scans) unless ONE row
      // (in-place stages, older rows fetched on demand) lets more warps live on the SM
      const char *force_rows = getenv("FLEMINGO_FAST_ROWS");
      for (int rows = 2; rows >= 1; --rows) {
        if (force_rows && atoi(force_rows) != rows) continue;
        // slot stride in 32-bit words = TEAM (mod 32), so that the teams of a warp work on disjoint banks
        const int slot_words = ((rows * fast_rowq[j] * elem / 4 + 31) & ~31) + (team & 31);
        const int slot = slot_words * 4 / elem;                   // in row elements
        // a small length class (a ragged dataset has many) cannot fill a 16-warp CTA: size the CTA for one round of the
        // class and let several organisms' CTAs share the SM instead
        const int need_w = std::max(4, (((cl.count + nt - 1) / nt) + 3) & ~3);
        for (int nw : kWarpChoices) {
          if (nw > max_warps) continue;
          if (force_w && atoi(force_w) != nw) continue;
          if (!force_w && nw > need_w && nw > 4) continue;
          const int cand_w = cand_per_warp(kFastConfigs[j].bits, nt);
          const size_t d = (size_t)pp.pool_stride + (size_t)(P.max_rec - 1) * pp.g_stride + 2ull * nw * P.max_rec * cand_w;
          const size_t i = (size_t)(P.max_rec - 1) * fast_gq[j] + fp.kmer_cap + (size_t)nw * nt * slot_words +
                           2ull * nw * P.max_rec * cand_w + (size_t)nw * nt * P.max_rec + (size_t)nw * nt * fast_bm[j];
          const size_t bytes = d * 8 + i * 4 + sizeof(RecDesc) * P.max_rec + sizeof(FastPlan) + 48 + (size_t)nw * nt * fp.k8_stride;
          if (bytes > (size_t)kMaxSmem) continue;
          const int ctas = (int)std::min<size_t>(233472 / (bytes + 1024), 32);
          const int reg_warps = ctx->fast_reg_warps[j];  // 16-19 for the 101-124 registers ptxas settles at
          // whole CTAs only; one large CTA beats several small ones at equal occupancy (the plan and the prologue are
          // paid per CTA and the hardware scheduler favours one CTA's warps), and more than 16 resident warps bought
          // with small CTAs measured slower (packed kernel on config 2: 5 CTAs of 4 warps 1.82 ms per launch, 1 CTA of 16 warps 1.39 ms)
          const int occ = std::min(max_warps, nw * std::min(ctas, reg_warps / nw)) * 64 + nw;
          if (occ > best_occ) {
            best_occ = occ;
            fast_warps[j] = nw;
            fast_smem[j] = bytes;
            fast_slot[j] = slot;
            fast_single[j] = rows == 1;
            fast_ctas_per_sm[j] = std::max(1, std::min(ctas, reg_warps / nw));
          }
        }
      }
      return fast_warps[j] > 0;
    };
    for (int j = 0; j < kNumFastConfigs; ++j) {
      if (fast_bucket[j].empty()) continue;
      if (!plan_config(j)) {
        for (int o : fast_bucket[j]) exact_bucket[pick_tile(cl.len - P.sum_len[o] + 1)].push_back(o);
        n_fast_org -= (long long)fast_bucket[j].size();
        fast_bucket[j].clear();
      }
    }
    std::vector<int> list;
    list.reserve(n_org);
    int exact_begin[kNumTileSizes], fast_begin[kNumFastConfigs];
    for (int j = 0; j < kNumTileSizes; ++j) {
      exact_begin[j] = (int)list.size();
      list.insert(list.end(), exact_bucket[j].begin(), exact_bucket[j].end());
    }
    for (int j = 0; j < kNumFastConfigs; ++j) {
      fast_begin[j] = (int)list.size();
      list.insert(list.end(), fast_bucket[j].begin(), fast_bucket[j].end());
    }
    int *d_list = ctx->org_list.as<int>() + (size_t)c * n_org;
    CK(cudaMemcpyAsync(d_list, list.data(), (size_t)n_org * 4, cudaMemcpyHostToDevice, st));
    const long long target_ctas = (long long)n_sm * 16;
    for (int j = 0; j < kNumTileSizes; ++j) {
      const int n_b = (int)exact_bucket[j].size();
      if (!n_b) continue;
      pp.org_list = d_list + exact_begin[j];
      // chunking: enough CTAs to fill the machine a few times over, at least one sequence per warp
      const long long pairs = (long long)n_b * cl.count;
      int spc = (int)std::max<long long>(nwarps, (pairs + target_ctas - 1) / target_ctas);
      spc = std::min(spc, std::max(nwarps, 64));
      spc = ((spc + nwarps - 1) / nwarps) * nwarps;
      pp.seqs_per_cta = spc;
      pp.chunks = (cl.count + spc - 1) / spc;
      const long long grid = (long long)n_b * pp.chunks;
      if (grid > 0x7fffffffLL) return fail(FL_E_UNSUPPORTED, "grid too large");
      kPlaceKernels[j][trace ? 1 : 0]<<<(unsigned)grid, nwarps * 32, smem, st>>>(pp);
      CK(cudaGetLastError());
      ctx->launches++;
    }
    if (n_fast_org > 0) {
      ctx->fast_pairs += n_fast_org * cl.count;
      // Geometry of every launch first, so that the hand-back lists can be sized before anything is launched.
      // A bucket of the packed 16-bit pass gets a SECOND-CHANCE launch of the int32 pass over the pairs it hands back
      // (near-ties beyond its candidate lists are exact ties only at the int32 scale); what that launch hands back, and
      // what the int32 buckets hand back, goes to the exact kernel.  Hand-back lists, in this order:
      //   [16-bit buckets with a second chance] [16-bit buckets without] [int32 buckets] [second-chance launches]
      // -- the exact list kernel reads everything from the second group on.
      struct Geo {
        int j, spc, chunks, per_round, n_b, org_begin, retry_of;
        long long items, grid, item_base, pair_base;
      };
      std::vector<Geo> geo;
      auto make_geo = [&](int j, int n_b, int org_begin) {
        const int team = kFastConfigs[j].team, nt = 32 / team, nw = fast_warps[j];
        Geo g;
        g.j = j;
        g.n_b = n_b;
        g.org_begin = org_begin;
        g.retry_of = -1;
        g.per_round = nw * nt;
        const long long pairs = (long long)n_b * cl.count;
        const long long fast_target = (long long)n_sm * fast_ctas_per_sm[j] * 6;
        int spc = (int)std::max<long long>(g.per_round, (pairs + fast_target - 1) / fast_target);
        spc = std::min(spc, std::max(g.per_round, getenv("FLEMINGO_FAST_SPC") ? atoi(getenv("FLEMINGO_FAST_SPC")) : 256));
        spc = ((spc + g.per_round - 1) / g.per_round) * g.per_round;
        g.spc = spc;
        g.chunks = (cl.count + spc - 1) / spc;
        g.items = (long long)n_b * g.chunks;
        g.grid = std::min<long long>(g.items, (long long)n_sm * fast_ctas_per_sm[j]);
        return g;
      };
      const bool allow_retry = !getenv("FLEMINGO_NO_RETRY");
      std::vector<Geo> first16r, first16, first32, retries;
      for (int j = 0; j < kNumFastConfigs; ++j) {
        const int n_b = (int)fast_bucket[j].size();
        if (!n_b) continue;
        Geo g = make_geo(j, n_b, fast_begin[j]);
        if (kFastConfigs[j].bits == 16) {
          const int j32 = pick_fast(cl.len - P.sum_len[fast_bucket[j][0]] + 1, 32, kFastConfigs[j].prune);
          if (allow_retry && j32 >= 0 && plan_config(j32)) {
            Geo r = g;   // same items, same per-item lists; another kernel
            r.j = j32;
            r.per_round = fast_warps[j32] * (32 / kFastConfigs[j32].team);
            r.grid = std::min<long long>(r.items, (long long)n_sm * fast_ctas_per_sm[j32]);
            r.retry_of = (int)first16r.size();
            first16r.push_back(g);
            retries.push_back(r);
          } else {
            first16.push_back(g);
          }
        } else {
          first32.push_back(g);
        }
      }
      const int n_first16r = (int)first16r.size();
      for (const std::vector<Geo> *grp : {&first16r, &first16, &first32, &retries}) geo.insert(geo.end(), grp->begin(), grp->end());
      if ((int)geo.size() - n_first16r > kMaxListBuckets) return fail(FL_E_UNSUPPORTED, "too many tile buckets");
      long long total_items = 0, pair_slots = 0, scratch_ints = 0;
      for (Geo &g : geo) {
        g.item_base = total_items;
        g.pair_base = pair_slots;
        total_items += g.items;
        pair_slots += g.items * g.spc;
        scratch_ints = std::max(scratch_ints, g.grid * g.per_round * (long long)P.max_rec * fast_rowq[g.j]);  // upper bound for both element sizes
        if (total_items > 0x7fffffffLL) return fail(FL_E_UNSUPPORTED, "too many work items");
      }
      // what the exact list kernel reads: every launch from the second group on
      ListBuckets lb;
      lb.n = 0;
      lb.item_begin[0] = 0;
      const long long list_item_base = n_first16r ? geo[n_first16r].item_base : 0;
      for (size_t b = n_first16r; b < geo.size(); ++b) {
        const Geo &g = geo[b];
        lb.chunks[lb.n] = g.chunks;
        lb.spc[lb.n] = g.spc;
        lb.org_begin[lb.n] = g.org_begin;
        lb.pair_base[lb.n] = g.pair_base;
        lb.item_begin[lb.n + 1] = lb.item_begin[lb.n] + (int)g.items;
        ++lb.n;
      }
      const int list_items = lb.item_begin[lb.n];
      if ((rc = lane.fb_item_count.ensure((size_t)total_items * 4 + 4)) || (rc = lane.fb_item_seqs.ensure((size_t)pair_slots * 4 + 4)) ||
          (rc = lane.fast_scratch.ensure((size_t)scratch_ints * 4)))
        return rc;
      CK(cudaMemsetAsync(lane.fb_item_count.p, 0, (size_t)total_items * 4, st));
      fp.fb_item_count = lane.fb_item_count.as<int>();
      fp.fb_item_seqs = lane.fb_item_seqs.as<int>();
      fp.scratch = lane.fast_scratch.as<int>();
      fp.retried_total = ctx->fb_retried.as<unsigned long long>();
      for (size_t b = 0; b < geo.size(); ++b) {
        const Geo &g = geo[b];
        const int j = g.j, nw = fast_warps[j];
        fp.rowq_stride = fast_rowq[j];
        fp.gq_stride = fast_gq[j];
        fp.gt_off = kGPad + Pmax + kGPad;
        fp.nb_max = fast_nb[j];
        fp.bm_stride = fast_bm[j];
        fp.cand_w = cand_per_warp(kFastConfigs[j].bits, 32 / kFastConfigs[j].team);
        fp.slotq_stride = fast_slot[j];
        fp.single_row = fast_single[j] ? 1 : 0;
        fp.scratch_slot_stride = (long long)P.max_rec * fast_rowq[j];
        fp.base = pp;
        fp.base.org_list = d_list + g.org_begin;
        fp.base.seqs_per_cta = g.spc;
        fp.base.chunks = g.chunks;
        fp.n_items = (int)g.items;
        fp.item_base = (int)g.item_base;
        fp.pair_base = g.pair_base;
        fp.retry_count = nullptr;
        fp.retry_seqs = nullptr;
        if (g.retry_of >= 0) {  // second chance: the item's sequences are the ones its first launch handed back
          fp.retry_count = fp.fb_item_count + geo[g.retry_of].item_base;
          fp.retry_seqs = fp.fb_item_seqs + geo[g.retry_of].pair_base;
        }
        fp.work_counter = lane.work_counter.as<int>();   // one per lane: launches of a lane run in stream order
        CK(cudaMemsetAsync(fp.work_counter, 0, sizeof(int), st));
        if (getenv("FLEMINGO_DEBUG_LAUNCH"))
          fprintf(stderr, "fast launch%s: cfg bits=%d%s J=%d team=%d warps=%d ctas/sm=%d grid=%lld items=%lld spc=%d smem=%zu rows=%d\n",
                  g.retry_of >= 0 ? " (second chance)" : "", kFastConfigs[j].bits, kFastConfigs[j].prune ? " pruned" : "", kFastConfigs[j].J, kFastConfigs[j].team, nw,
                  fast_ctas_per_sm[j], g.grid, g.items, g.spc, fast_smem[j], fast_single[j] ? 1 : 2);
        kFastConfigs[j].fn<<<(unsigned)g.grid, nw * 32, fast_smem[j], st>>>(fp);
        CK(cudaGetLastError());
        ctx->launches++;
      }
      // the pairs the integer pass declined, by the exact kernel in list mode: one organism prologue per work item that
      // has hand-backs, the item's pairs spread over the warps of the CTA
      PlaceParams lp = pp;
      lp.org_list = d_list;
      int lw = 4;
      size_t lsmem = 0;
      for (;; lw >>= 1) {
        lsmem = 8ull * ((size_t)lp.pool_stride + (size_t)(P.max_rec - 1) * lp.g_stride + (size_t)lw * lp.nrows * lp.row_stride) +
                sizeof(RecDesc) * P.max_rec + (size_t)lw * lp.code_stride;
        if (lsmem <= (size_t)kMaxSmem / 2 || lw == 1) break;
      }
      if (lsmem > (size_t)kMaxSmem) return fail(FL_E_UNSUPPORTED, "list-mode placement needs %zu bytes of shared memory (limit %d)", lsmem, kMaxSmem);
      if ((rc = lane.fb_unit_begin.ensure(((size_t)list_items + 1) * 4))) return rc;
      const int *list_count = fp.fb_item_count + list_item_base;
      list_units_kernel<<<1, 1024, 0, st>>>(list_count, list_items, lane.fb_unit_begin.as<int>(),
                                                     ctx->fb_total.as<unsigned long long>());
      CK(cudaGetLastError());
      const int list_ctas_per_sm = std::max(1, std::min(4, (int)(233472 / (lsmem + 1024))));
      const unsigned lgrid = (unsigned)((long long)n_sm * list_ctas_per_sm);
      kListKernels[pick_tile(Pmax)][trace ? 1 : 0]<<<lgrid, lw * 32, lsmem, st>>>(lp, lb, list_count, fp.fb_item_seqs,
                                                                                          lane.fb_unit_begin.as<int>());
      CK(cudaGetLastError());
      ctx->launches += 2;
    }
  }
  if (n_lanes > 1) {  // the main stream continues after every lane has finished
    for (int l = 1; l < n_lanes; ++l) {
      CK(cudaEventRecord(ctx->lanes[l].done, ctx->lanes[l].stream));
      CK(cudaStreamWaitEvent(ctx->stream, ctx->lanes[l].done, 0));
    }
  }
  set.E_n_org = n_org;
  set.E_valid = true;
  if (energies || gaps || rec_scores || con_scores) {
    const size_t pairs = (size_t)n_org * n_seq;
    if (energies) CK(cudaMemcpyAsync(energies, set.E.p, pairs * 8, cudaMemcpyDeviceToHost, ctx->stream));
    if (gaps) CK(cudaMemcpyAsync(gaps, ctx->gaps.p, pairs * P.max_rec * 4, cudaMemcpyDeviceToHost, ctx->stream));
    if (rec_scores) CK(cudaMemcpyAsync(rec_scores, ctx->rsc.p, pairs * P.max_rec * 8, cudaMemcpyDeviceToHost, ctx->stream));
    if (con_scores) CK(cudaMemcpyAsync(con_scores, ctx->csc.p, pairs * P.max_rec * 8, cudaMemcpyDeviceToHost, ctx->stream));
    return check_device_errors(ctx);
  }
  return FL_OK;
}

extern "C" int fl_place_batch(fl_ctx *ctx, int set_id, const long long *con_begin, int flags, double *energies, int *gaps,
                              double *rec_scores, double *con_scores) {
  if (!ctx) return fail(FL_E_ARG, "fl_place_batch: null context");
  if (set_id < 0 || set_id >= FL_MAX_SETS) return fail(FL_E_ARG, "fl_place_batch: set_id %d out of range", set_id);
  CK(cudaSetDevice(ctx->device));
  return place_batch(ctx, ctx->pop, ctx->sets[set_id], con_begin, flags, energies, gaps, rec_scores, con_scores);
}

extern "C" int fl_path_stats(fl_ctx *ctx, long long *fast_pairs, long long *fallback_pairs, long long *second_chance_pairs) {
  if (!ctx) return fail(FL_E_ARG, "fl_path_stats: null context");
  CK(cudaSetDevice(ctx->device));
  unsigned long long fb = 0, retried = 0;
  CK(cudaMemcpyAsync(&fb, ctx->fb_total.p, sizeof fb, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaMemcpyAsync(&retried, ctx->fb_retried.p, sizeof retried, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
#ifdef FL_PHASE_CLOCKS
  {
    unsigned long long clk[16] = {0};
    CK(cudaMemcpyFromSymbol(clk, g_phase_clk, sizeof clk));
    static const char *names[16] = {"prologue", "plan", "unpack", "scores", "dp_stage", "refine(5)", "misc", "claim/tail", "refine(1)", "refine(2)", "refine(3)", "refine(4)", "", "", "", ""};
    double tot = 0;
    for (int i = 0; i < 12; ++i) tot += (double)clk[i];
    fprintf(stderr, "pruned stages: %llu pair-stages ran %llu of %llu tiles (%.1f %%) in %llu lane slots (%.1f %% of the slots ran a tile)\n", clk[14], clk[12], clk[13],
            clk[13] ? 100.0 * (double)clk[12] / (double)clk[13] : 0.0, clk[15], clk[15] ? 100.0 * (double)clk[12] / (double)clk[15] : 0.0);
    for (int i = 0; i < 12; ++i) fprintf(stderr, "phase %-10s %14llu ticks %5.1f %%\n", names[i], clk[i], tot > 0 ? 100.0 * clk[i] / tot : 0.0);
    unsigned long long zero[16] = {0};
    CK(cudaMemcpyToSymbol(g_phase_clk, zero, sizeof zero));
    if (!getenv("FLEMINGO_TIMELINE")) {
      int zc[16] = {0};
      CK(cudaMemcpyToSymbol(g_timeline_n, zc, sizeof zc));
      CK(cudaMemcpyToSymbol(g_timeline_slots, zc, sizeof(int)));
    }
    if (const char *path = getenv("FLEMINGO_TIMELINE")) {
      std::vector<unsigned long long> tl(16 * kTimelineEvents);
      int cnt[16];
      CK(cudaMemcpyFromSymbol(tl.data(), g_timeline, tl.size() * 8));
      CK(cudaMemcpyFromSymbol(cnt, g_timeline_n, sizeof cnt));

      if (FILE *fh = fopen(path, "w")) {
        for (int w = 0; w < 16; ++w)
          for (int e = 0; e < cnt[w]; ++e) fprintf(fh, "%d %llu %d\n", w, tl[(size_t)w * kTimelineEvents + e] >> 4, (int)(tl[(size_t)w * kTimelineEvents + e] & 15));
        fclose(fh);
      }
    }
  }
#endif
  if (fast_pairs) *fast_pairs = ctx->fast_pairs;
  if (fallback_pairs) *fallback_pairs = (long long)fb;
  if (second_chance_pairs) *second_chance_pairs = (long long)retried;
  ctx->fast_pairs = 0;
  CK(cudaMemsetAsync(ctx->fb_total.p, 0, sizeof fb, ctx->stream));
  CK(cudaMemsetAsync(ctx->fb_retried.p, 0, sizeof retried, ctx->stream));
  return FL_OK;
}

// Launches the fitness reduction of the resident energies into ctx->fit_out ([max(n_org, pad_rows)][5], rows beyond
// n_org zeroed).  Asynchronous.
static int fitness_launch(fl_ctx *ctx, int pos_set, int neg_set, int method, double gamma, int pad_rows, int *n_org_out) {
  if (pos_set < 0 || pos_set >= FL_MAX_SETS || neg_set < 0 || neg_set >= FL_MAX_SETS) return fail(FL_E_ARG, "fl_fitness: set id out of range");
  if (method < FL_FIT_WELCH || method > FL_FIT_TRIM_LEFT_M_SD) return fail(FL_E_ARG, "fl_fitness: unknown method %d", method);
  CK(cudaSetDevice(ctx->device));
  SeqSet &sp = ctx->sets[pos_set], &sn = ctx->sets[neg_set];
  if (!sp.E_valid || !sn.E_valid || sp.E_n_org != sn.E_n_org) return fail(FL_E_STATE, "fl_fitness: energies of both sets must be computed (fl_place_batch) for the current population");
  const int n_org = sp.E_n_org;
  *n_org_out = n_org;
  const int rows = std::max(n_org, pad_rows);
  int rc;
  if ((rc = ctx->fit_out.ensure((size_t)std::max(rows, 1) * 5 * 8))) return rc;
  if (rows > n_org) CK(cudaMemsetAsync(ctx->fit_out.as<double>() + (size_t)n_org * 5, 0, (size_t)(rows - n_org) * 5 * 8, ctx->stream));
  if (n_org == 0) return FL_OK;
  if (sp.n_seq < 1 || sn.n_seq < 1) return fail(FL_E_ARG, "fl_fitness: empty sequence set");
  if (std::max(sp.n_seq, sn.n_seq) > 60 * kPwLeafCap)
    return fail(FL_E_UNSUPPORTED, "fl_fitness: sets of more than %d sequences exceed the reduction plan of the fitness kernel", 60 * kPwLeafCap);
  // n_to_trim = round(gamma * N): python rounds half to even (organism_object.py:892), as nearbyint does
  const int trim_p = (int)std::nearbyint(gamma * (double)sp.n_seq);
  const int trim_n = (int)std::nearbyint(gamma * (double)sn.n_seq);
  int n2 = 1;
  while (n2 < std::max(sp.n_seq, sn.n_seq)) n2 <<= 1;
  const bool need_sort = method != FL_FIT_WELCH && (trim_p != 0 || trim_n != 0);
  if (need_sort && (rc = ctx->fit_scratch.ensure((size_t)n_org * n2 * 8))) return rc;
  fitness_kernel<<<n_org, 256, 0, ctx->stream>>>(sp.E.as<double>(), sp.n_seq, sn.E.as<double>(), sn.n_seq, method, trim_p, trim_n,
                                                 ctx->fit_scratch.as<double>(), n2, ctx->fit_out.as<double>());
  CK(cudaGetLastError());
  ctx->launches++;
  return FL_OK;
}

extern "C" int fl_fitness(fl_ctx *ctx, int pos_set, int neg_set, int method, double gamma, double *out) {
  if (!ctx || !out) return fail(FL_E_ARG, "fl_fitness: null argument");
  int n_org = 0;
  const int rc = fitness_launch(ctx, pos_set, neg_set, method, gamma, 0, &n_org);
  if (rc) return rc;
  if (n_org == 0) return FL_OK;
  CK(cudaMemcpyAsync(out, ctx->fit_out.p, (size_t)n_org * 5 * 8, cudaMemcpyDeviceToHost, ctx->stream));
  const int rc2 = check_device_errors(ctx);   // synchronises
  if (rc2) return rc2;
  return fl_fitness_finalize(out, n_org);
}

extern "C" int fl_fitness_dev(fl_ctx *ctx, int pos_set, int neg_set, int method, double gamma, int pad_rows, void **out_dev) {
  if (!ctx || !out_dev) return fail(FL_E_ARG, "fl_fitness_dev: null argument");
  if (pad_rows < 0) return fail(FL_E_ARG, "fl_fitness_dev: negative pad_rows");
  int n_org = 0;
  const int rc = fitness_launch(ctx, pos_set, neg_set, method, gamma, pad_rows, &n_org);
  if (rc) return rc;
  *out_dev = ctx->fit_out.p;
  return FL_OK;
}

extern "C" int fl_calculate(fl_ctx *ctx, const char *sequence, int seq_len, const char *rec_types, int n_rec,
                            const double *rec_matrices, const int *rec_lengths, const double *con_matrices,
                            long long n_con_doubles, double *rec_scores, double *con_scores, int *con_lengths, int max_length,
                            const double *bin_freqs, const double *bin_edges, const int *num_bins) {
  if (!ctx || !sequence || !rec_types || !rec_lengths || !rec_scores || !con_lengths) return fail(FL_E_ARG, "fl_calculate: null argument");
  if (n_rec < 1 || n_rec > FL_MAX_RECOGNIZERS) return fail(FL_E_ARG, "fl_calculate: %d recognizers (supported 1..%d)", n_rec, FL_MAX_RECOGNIZERS);
  if (n_rec > 1 && !con_scores) return fail(FL_E_ARG, "fl_calculate: con_scores is null");
  CK(cudaSetDevice(ctx->device));
  // _interface.c:131-136: 2 doubles per connector means "not precomputed"
  if (n_rec > 1 && n_con_doubles == 2LL * (n_rec - 1))
    return fail(FL_E_UNSUPPORTED, "connectors without precomputed pdf/cdf tables (PRECOMPUTE=false, connector.c:70-72) are not supported");
  const int M = max_length + 1;
  if (n_rec > 1 && (M < 1 || n_con_doubles != (long long)(n_rec - 1) * (2LL * M + 2)))
    return fail(FL_E_ARG, "fl_calculate: con_matrices has %lld doubles, expected %lld for max_length %d", n_con_doubles, (long long)(n_rec - 1) * (2LL * M + 2), max_length);
  // build a one-organism population
  std::vector<double> pool, edges_pool;
  std::vector<long long> rec_data(n_rec), cls_edges;
  std::vector<int> rec_nb(n_rec, 0), rec_class(n_rec, -1), cls_len, cls_nb;
  std::vector<char> cls_kind;
  long long m_off = 0, a_off = 0, e_off = 0;
  int shape_i = 0;
  for (int i = 0; i < n_rec; ++i) {
    rec_data[i] = (long long)pool.size();
    const int L = rec_lengths[i];
    if (L < 1) return fail(FL_E_ARG, "fl_calculate: recognizer %d has length %d", i, L);
    if (rec_types[i] == 'p' || rec_types[i] == 'P') {
      if (!rec_matrices) return fail(FL_E_ARG, "fl_calculate: rec_matrices is null");
      pool.insert(pool.end(), rec_matrices + m_off, rec_matrices + m_off + 4LL * L);
      m_off += 4LL * L;
    } else {
      if (!bin_freqs || !bin_edges || !num_bins) return fail(FL_E_ARG, "fl_calculate: shape model arrays are null");
      const int nb = num_bins[shape_i];
      if (nb < 2) return fail(FL_E_ARG, "fl_calculate: shape recognizer %d has %d bin edges", i, nb);
      rec_nb[i] = nb;
      pool.insert(pool.end(), bin_freqs + a_off, bin_freqs + a_off + 2LL * (nb - 1));  // null ++ alt (organism.c:159-172)
      rec_class[i] = (int)cls_kind.size();
      cls_kind.push_back(rec_types[i]);
      cls_len.push_back(L);
      cls_nb.push_back(nb);
      cls_edges.push_back((long long)edges_pool.size());
      edges_pool.insert(edges_pool.end(), bin_edges + e_off, bin_edges + e_off + nb);
      a_off += 2LL * (nb - 1);
      e_off += nb;
      ++shape_i;
    }
  }
  // ---- one placement = ONE upload, three small kernels, ONE download -------------------------------------------
  // Everything the kernels read goes into one blob (pinned staging buffer -> device arena, a single copy): the
  // one-organism population, the shape classes, the raw sequence; the arena also holds what pack_sequences_kernel and
  // shape_bins_kernel produce and the result block.  The reference's callers make one such call per (organism,
  // sequence) -- 100 + 100 per fitness evaluation -- so the ~35 separate copies of the batched entry points
  // (fl_upload_population + fl_upload_sequences + fl_place_batch, 0.19 ms) are worth avoiding here.
  int sum_len = 0;
  for (int i = 0; i < n_rec; ++i) sum_len += rec_lengths[i];
  if (sum_len > seq_len)
    return fail(FL_E_TOO_LARGE, "organism 0 (sum of recognizer lengths %d) is too large to be placed on a sequence of length %d", sum_len, seq_len);
  const int Pn = seq_len - sum_len + 1;
  if (n_rec > 1) {
    if (seq_len - sum_len > M - 1)
      return fail(FL_E_UNSUPPORTED, "organism 0: gap tables cover %d lengths but a sequence of length %d needs %d (legacy non-precomputed branch, organism.c:336-338)",
                  M, seq_len, seq_len - sum_len + 1);
    if (seq_len - sum_len + n_rec > 10000)
      return fail(FL_E_UNSUPPORTED, "effective length %d exceeds DEN_EXPANSION (10000 entries, _constants.h:11)", seq_len - sum_len + n_rec);
  }
  const int n_cls = (int)cls_kind.size();
  for (int c = 0; c < n_cls; ++c)
    if (cls_nb[c] > 256) return fail(FL_E_UNSUPPORTED, "shape class %d has %d bin edges (supported: 2..256)", c, cls_nb[c]);
  struct Blob {
    std::vector<unsigned char> bytes;
    size_t add(const void *src, size_t n, size_t align = 16) {
      const size_t off = (bytes.size() + align - 1) / align * align;
      bytes.resize(off + n, 0);
      if (src && n) memcpy(bytes.data() + off, src, n);
      return off;
    }
  } blob;
  const int i_rec_begin[2] = {0, n_rec}, i_zero = 0, i_M = n_rec > 1 ? M : 0;
  const long long l_pool_begin[2] = {0, (long long)pool.size()}, l_zero = 0, l_offs[2] = {0, seq_len};
  const long long bins_row_stride = (seq_len + 15) & ~15LL;
  const size_t o_result = blob.add(nullptr, 16 + 8 + (size_t)n_rec * (4 + 8 + 8) + 32);   // err flag | E | gaps | rsc | csc (zeroed)
  const size_t o_rec_begin = blob.add(i_rec_begin, sizeof i_rec_begin), o_rec_len = blob.add(rec_lengths, (size_t)n_rec * 4);
  const size_t o_rec_nb = blob.add(rec_nb.data(), (size_t)n_rec * 4), o_rec_class = blob.add(rec_class.data(), (size_t)n_rec * 4);
  const size_t o_con_M = blob.add(&i_M, 4), o_sum_len = blob.add(&sum_len, 4), o_zero_i = blob.add(&i_zero, 4), o_seq_len = blob.add(&seq_len, 4);
  const size_t o_cls_len = blob.add(cls_len.data(), (size_t)n_cls * 4), o_cls_nb = blob.add(cls_nb.data(), (size_t)n_cls * 4);
  const size_t o_rec_type = blob.add(rec_types, (size_t)n_rec), o_cls_kind = blob.add(cls_kind.data(), (size_t)n_cls);
  const size_t o_raw = blob.add(sequence, (size_t)seq_len);
  blob.add(nullptr, 16);
  const size_t o_rec_data = blob.add(rec_data.data(), (size_t)n_rec * 8), o_pool_begin = blob.add(l_pool_begin, sizeof l_pool_begin);
  const size_t o_zero_l = blob.add(&l_zero, 8), o_offs = blob.add(l_offs, sizeof l_offs), o_cls_edges = blob.add(cls_edges.data(), (size_t)n_cls * 8);
  const size_t o_pool = blob.add(pool.data(), pool.size() * 8), o_con = blob.add(n_rec > 1 ? con_matrices : nullptr, n_rec > 1 ? (size_t)n_con_doubles * 8 : 0);
  const size_t o_edges = blob.add(edges_pool.data(), edges_pool.size() * 8);
  const size_t upload_bytes = blob.add(nullptr, 16);
  // device-only part of the arena
  const size_t o_packed = blob.add(nullptr, ((size_t)(seq_len + 15) / 16 + 2) * 4), o_unk = blob.add(nullptr, ((size_t)(seq_len + 31) / 32 + 2) * 4);
  const size_t o_has_unk = blob.add(nullptr, 16), o_k8 = blob.add(nullptr, (((size_t)seq_len + kK8Pad + 15) & ~(size_t)15) + 16);
  const size_t o_bins = blob.add(nullptr, (size_t)std::max(1, n_cls) * (size_t)bins_row_stride + 16);
  const size_t arena_bytes = blob.add(nullptr, 256);
  int rc;
  if ((rc = ctx->single_arena.ensure(arena_bytes))) return rc;
  if (ctx->single_stage_cap < upload_bytes) {
    if (ctx->single_stage) cudaFreeHost(ctx->single_stage);
    ctx->single_stage = nullptr;
    ctx->single_stage_cap = 0;
    CK(cudaMallocHost(&ctx->single_stage, upload_bytes * 2 + 4096));
    ctx->single_stage_cap = upload_bytes * 2 + 4096;
  }
  memcpy(ctx->single_stage, blob.bytes.data(), upload_bytes);
  unsigned char *A = ctx->single_arena.as<unsigned char>();
  cudaStream_t st = ctx->stream;
  CK(cudaMemcpyAsync(A, ctx->single_stage, upload_bytes, cudaMemcpyHostToDevice, st));
  auto at = [&](size_t off) { return A + off; };
  pack_sequences_kernel<<<1, 64, 0, st>>>(at(o_raw), (const long long *)at(o_offs), (const int *)at(o_zero_i), (const int *)at(o_zero_i),
                                        (unsigned *)at(o_packed), (unsigned *)at(o_unk), at(o_has_unk), (const int *)at(o_zero_i), at(o_k8));
  if (n_cls > 0) {
    ShapeBinParams q;
    q.packed = (const unsigned *)at(o_packed);
    q.word_off = (const int *)at(o_zero_i);
    q.seq_len = (const int *)at(o_seq_len);
    q.n_seq = 1;
    q.cls_kind = (const char *)at(o_cls_kind);
    q.cls_len = (const int *)at(o_cls_len);
    q.cls_nb = (const int *)at(o_cls_nb);
    q.cls_edges = (const long long *)at(o_cls_edges);
    q.edges_pool = (const double *)at(o_edges);
    q.tables = ctx->tables.as<double>();
    q.bins = at(o_bins);
    q.row_stride = bins_row_stride;
    shape_bins_kernel<<<dim3(1u, (unsigned)n_cls), 128, 0, st>>>(q);
  }
  PlaceParams pp;
  pp.rec_begin = (const int *)at(o_rec_begin);
  pp.rec_type = (const char *)at(o_rec_type);
  pp.rec_len = (const int *)at(o_rec_len);
  pp.rec_nb = (const int *)at(o_rec_nb);
  pp.rec_data = (const long long *)at(o_rec_data);
  pp.pool_begin = (const long long *)at(o_pool_begin);
  pp.rec_pool = (const double *)at(o_pool);
  pp.con_M = (const int *)at(o_con_M);
  pp.con_pool = (const double *)at(o_con);
  pp.con_begin = (const long long *)at(o_zero_l);
  pp.sum_len = (const int *)at(o_sum_len);
  pp.org_list = (const int *)at(o_zero_i);
  pp.packed = (const unsigned *)at(o_packed);
  pp.unk = (const unsigned *)at(o_unk);
  pp.word_off = (const int *)at(o_zero_i);
  pp.k8g = at(o_k8);
  pp.k8_off = (const int *)at(o_zero_i);
  pp.unk_off = (const int *)at(o_zero_i);
  pp.order = (const int *)at(o_zero_i);
  pp.class_begin = 0;
  pp.class_count = 1;
  pp.S = seq_len;
  pp.n_seq_total = 1;
  pp.tables = ctx->tables.as<double>();
  pp.rec_class = (const int *)at(o_rec_class);
  pp.bins = at(o_bins);
  pp.bins_row_stride = bins_row_stride;
  // result block: [err flag, pad] [E] [rsc n] [csc n] [gaps n]
  pp.err_flag = (int *)at(o_result);
  pp.E = (double *)at(o_result + 16);
  pp.rsc = pp.E + 1;
  pp.csc = pp.rsc + n_rec;
  pp.gaps = (int *)(pp.csc + n_rec);
  pp.max_rec = n_rec;
  pp.pool_stride = ((int)pool.size() + 1) & ~1;
  pp.g_stride = (kGPad + Pn + kGPad + 1) & ~1;
  pp.row_stride = (Pn + kMaxJ + 3) & ~3;
  pp.code_stride = (seq_len + 15) & ~15;
  pp.nrows = n_rec;
  pp.seqs_per_cta = 1;
  pp.chunks = 1;
  int nwarps = 4;   // warps 1-3 only help with the prologue (gap tables); the pair itself is one warp's work
  size_t smem = 0;
  for (;; nwarps >>= 1) {
    smem = 8ull * ((size_t)pp.pool_stride + (size_t)(n_rec - 1) * pp.g_stride + (size_t)nwarps * pp.nrows * pp.row_stride) +
           sizeof(RecDesc) * n_rec + (size_t)nwarps * pp.code_stride;
    if (smem <= (size_t)kMaxSmem || nwarps == 1) break;
  }
  if (smem > (size_t)kMaxSmem)
    return fail(FL_E_UNSUPPORTED, "placement of %d offsets x %d recognizers needs %zu bytes of shared memory (limit %d)", Pn, n_rec, smem, kMaxSmem);
  kPlaceKernels[pick_tile(Pn)][1]<<<1, nwarps * 32, smem, st>>>(pp);
  CK(cudaGetLastError());
  ctx->launches += n_cls > 0 ? 3 : 2;
  const size_t result_bytes = 16 + 8 + (size_t)n_rec * (8 + 8 + 4);
  unsigned char *host_result = (unsigned char *)ctx->single_stage + upload_bytes;   // second half of the pinned buffer
  CK(cudaMemcpyAsync(host_result, at(o_result), result_bytes, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  const int flag = *(const int *)host_result;
  if (flag & ERR_NAN_NULL)
    return fail(FL_E_NAN_NULL, "a shape score fell into a null-model bin whose frequency is NaN (the reference exits here, recognizer.c:258-262)");
  const double *res = (const double *)(host_result + 16);
  const int *res_gaps = (const int *)(res + 1 + 2 * n_rec);
  for (int i = 0; i < n_rec; ++i) {
    rec_scores[i] = res[1 + i];
    con_lengths[i] = res_gaps[i];
    if (i + 1 < n_rec) con_scores[i] = res[1 + n_rec + i];
  }
  rec_scores[n_rec] = res[0];
  return FL_OK;
}

// ------------------------------------------------------------------------------------------------
// Roofline denominator, measured where the bench runs: the issue rate of the instruction the DP stage is made of
// (VIADDMNMX, max(a + b, c) in one instruction; 8 independent chains per thread, no memory traffic).
// ------------------------------------------------------------------------------------------------
template <bool PACKED16>
__global__ void __launch_bounds__(256) dpx_rate_kernel(int *out, const int *in, int iters) {
  constexpr int ACC = 8;
  int acc[ACC], g[ACC];
#pragma unroll
  for (int a = 0; a < ACC; ++a) {
    acc[a] = in[threadIdx.x + a * 256];
    g[a] = in[threadIdx.x + 2048 + a * 256];
  }
  int p = in[4096 + (threadIdx.x & 31)];
  const int dp = in[4200];
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
#pragma unroll
      for (int a = 0; a < ACC; ++a)
        acc[a] = PACKED16 ? (int)__viaddmax_s16x2((unsigned)p, (unsigned)g[a], (unsigned)acc[a]) : __viaddmax_s32(p, g[a], acc[a]);
      p += dp;
    }
  }
  int r = acc[0];
#pragma unroll
  for (int a = 1; a < ACC; ++a) r += acc[a];
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

extern "C" int fl_measure_dpx_rate(fl_ctx *ctx, double *cells_per_s_i32, double *cells_per_s_i16x2, int *n_sm, int *sm_clock_khz) {
  if (!ctx) return fail(FL_E_ARG, "fl_measure_dpx_rate: null context");
  CK(cudaSetDevice(ctx->device));
  int sms = 0, khz = 0;
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ctx->device));
  CK(cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, ctx->device));
  const int blocks = sms * 8, iters = 2048;
  DevBuf in, out;
  int rc;
  if ((rc = in.ensure(8192 * 4)) || (rc = out.ensure((size_t)blocks * 256 * 4))) return rc;
  std::vector<int> h(8192);
  for (int i = 0; i < 8192; ++i) h[i] = -(i % 97) * 37 - 1;
  CK(cudaMemcpyAsync(in.p, h.data(), 8192 * 4, cudaMemcpyHostToDevice, ctx->stream));
  cudaEvent_t a, b;
  CK(cudaEventCreate(&a));
  CK(cudaEventCreate(&b));
  double rate[2] = {0, 0};
  for (int v = 0; v < 2; ++v) {
    float best = 1e30f;
    for (int rep = 0; rep < 6; ++rep) {  // the first repetition warms up
      CK(cudaEventRecord(a, ctx->stream));
      if (v == 0)
        dpx_rate_kernel<false><<<blocks, 256, 0, ctx->stream>>>(out.as<int>(), in.as<int>(), iters);
      else
        dpx_rate_kernel<true><<<blocks, 256, 0, ctx->stream>>>(out.as<int>(), in.as<int>(), iters);
      CK(cudaEventRecord(b, ctx->stream));
      CK(cudaEventSynchronize(b));
      float ms = 0;
      CK(cudaEventElapsedTime(&ms, a, b));
      if (rep > 0 && ms < best) best = ms;
    }
    const double cells = (double)blocks * 256 * 4.0 * 8 * iters * (v == 1 ? 2.0 : 1.0);
    rate[v] = cells / (best * 1e-3);
  }
  cudaEventDestroy(a);
  cudaEventDestroy(b);
  in.release();
  out.release();
  if (cells_per_s_i32) *cells_per_s_i32 = rate[0];
  if (cells_per_s_i16x2) *cells_per_s_i16x2 = rate[1];
  if (n_sm) *n_sm = sms;
  if (sm_clock_khz) *sm_clock_khz = khz;
  return FL_OK;
}
