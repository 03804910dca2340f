/*
 * flemingo_b200.h -- C ABI of libflemingo_b200.so, the B200 (sm_100a) replacement for FLEMINGO's
 * placement/fitness hot path.
 *
 * Every entry point cites the reference interface it replaces (paths relative to the FLEMINGO tree).
 * Conventions:
 *   - plain C types only; every pointer argument is HOST memory owned by the caller unless a name ends in
 *     "_dev"; device memory is owned by the context;
 *   - every function returns 0 (FL_OK) or a negative FL_E* code; the message is in fl_last_error()
 *     (thread-local).  Nothing calls exit() and no C++ exception crosses the boundary (the reference calls
 *     exit(1) on a NaN null-model bin, extensions/multiplacement/src/organism/recognizer.c:258-262);
 *   - all device work of a context is issued on ONE stream (fl_ctx_stream); calls that return host data
 *     synchronise that stream, the others are asynchronous until fl_sync();
 *   - input buffers may be reused or freed as soon as a call returns: copies from pageable memory are staged
 *     before the call returns, and an upload whose source is pinned (page-locked) memory synchronises the stream;
 *   - one context per process and device; contexts are not thread-safe (the reference holds the GIL for the
 *     whole call, src/_interface.c:45-185).
 */
#ifndef FLEMINGO_B200_H
#define FLEMINGO_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define FL_OK 0
#define FL_E_CUDA (-1)        /* a CUDA runtime call failed */
#define FL_E_ARG (-2)         /* malformed argument */
#define FL_E_TOO_LARGE (-3)   /* sum of recognizer lengths > sequence length:
                                 organism.c:262 returns -1 -> RuntimeError at _interface.c:162-166;
                                 organism_object.py:1279 raises ValueError first */
#define FL_E_NAN_NULL (-4)    /* a NaN null-model bin was selected (reference: exit(1), recognizer.c:258-262) */
#define FL_E_UNSUPPORTED (-5) /* input needs the legacy non-precomputed connector branch (connector.c:70-72,
                                 organism.c:336-338), gap tables shorter than the sequence, eff_len > 10000
                                 (reads past DEN_EXPANSION, _constants.h:11), or exceeds the shared-memory plan */
#define FL_E_STATE (-6)       /* call order: missing population / sequence set / energies */

#define FL_N_TABLE_DOUBLES 16144 /* MGW[1024] PROT[1024] ROLL[2048] HELT[2048] DEN_EXPANSION[10000], _constants.h:7-11 */
#define FL_MAX_RECOGNIZERS 16    /* per organism; reference config default MAX_NUM_OF_RECOGNIZERS is 5 */
#define FL_MAX_SETS 8

/* flags of fl_place_batch */
#define FL_WANT_TRACE 1   /* also produce start/gaps and per-node scores (organism.c:398-427) */
#define FL_EXACT_ONLY 2   /* skip the integer filter pass: every pair goes through the exact fp64 kernel */

/* fitness methods, src/objects/organism_object.py:848-936 */
#define FL_FIT_WELCH 0
#define FL_FIT_YUEN 1
#define FL_FIT_TRIM_LEFT_M 2
#define FL_FIT_TRIM_LEFT_M_SD 3

typedef struct fl_ctx fl_ctx;

/*
 * A packed population: struct-of-arrays over organisms and their recognizers.  Built by the host from the
 * per-organism flat arrays that OrganismObject.flatten() produces (organism_object.py:1362-1504).
 */
typedef struct fl_population {
  int n_org;
  int n_rec_total;
  const int *rec_begin;        /* [n_org+1] first recognizer of organism o in the per-recognizer arrays */
  const char *rec_type;        /* [n_rec_total] 'p','m','t','r','h' (shape_object.py:407-418) */
  const int *rec_len;          /* [n_rec_total] recognizer length in bp */
  const int *rec_nb;           /* [n_rec_total] number of bin EDGES of a shape recognizer (0 for 'p') */
  const long long *rec_data;   /* [n_rec_total] offset into rec_pool: PSSM = 4*len doubles, column major
                                  [col][A,G,C,T] (organism_object.py:1455-1458); shape = null[nb-1] ++
                                  alt[nb-1] (organism.c:159-172); the bin edges live in the shape classes */
  const long long *pool_begin; /* [n_org+1] slice of rec_pool used by organism o */
  const double *rec_pool;
  long long n_pool;
  const int *con_M;            /* [n_org] length M of each pdf / cdf block = ConnectorObject.max_seq_length
                                  (the reference passes max_length = M-1, organism_object.py:1307) */
  const double *con_pool;      /* connector tables: per organism (n_rec-1) blocks [mu, sigma, pdf[M], cdf[M]]
                                  (organism_object.py:1490-1495); variants rescaled by adjust_scores
                                  (connector_object.py:454-487) are appended by the host */
  long long n_con;
  /* shape classes: the distinct (kind, length, bin edges) among the population's shape recognizers.  The bin
     an averaged shape score falls into depends only on the class, the sequence and the offset, so it is
     computed once per dataset (recognizer.c:187-589, _aux.c:184-217) instead of once per organism. */
  const int *rec_class;        /* [n_rec_total] class of a shape recognizer, -1 for 'p' */
  int n_classes;
  const char *cls_kind;        /* [n_classes] 'm','t','r','h' */
  const int *cls_len;          /* [n_classes] recognizer length */
  const int *cls_nb;           /* [n_classes] number of bin edges (<= 256) */
  const long long *cls_edges;  /* [n_classes] offset of the class's edges in edges_pool */
  const double *edges_pool;
  long long n_edges;
} fl_population;

const char *fl_last_error(void);
int fl_abi_version(void);

/* CUDA devices visible to this process (0 if there is none or the driver is missing); one context per device lets a
   single GA driver process feed all GPUs of a box (flemingo_b200.distributed.MultiDeviceEvaluator). */
int fl_device_count(void);

/* Creates the context on `device` and uploads the data tables (_constants.h:7-11; tables.bin order). */
int fl_ctx_create(int device, const double *tables, long long n_tables, fl_ctx **out);
int fl_ctx_destroy(fl_ctx *ctx);
/* The cudaStream_t all work of this context is issued on (for event timing by the caller). */
void *fl_ctx_stream(fl_ctx *ctx);
int fl_sync(fl_ctx *ctx);

/*
 * Uploads a DNA set (replaces the per-call `bytes(sequence, "ASCII")` of organism_object.py:1331).
 * bytes = the sequences back to back, offsets[n_seq+1].  Bases are packed 2 bits each (A=0,G=1,C=2,T=3,
 * case-insensitive, recognizer.c:143-163) plus a 1-bit "unknown byte" mask so that non-ACGT bytes keep
 * the reference behaviour (no PSSM contribution; code 0 in a pentamer index).
 * Sequences are grouped by length into classes (ascending length); n_classes/class_len report them.
 */
int fl_upload_sequences(fl_ctx *ctx, int set_id, const char *bytes, const long long *offsets, int n_seq);
int fl_sequence_classes(fl_ctx *ctx, int set_id, int *n_classes, int *class_len /* [n_classes] or NULL */);

/* Uploads the population (replaces parse_org, organism.c:112-211). */
int fl_upload_population(fl_ctx *ctx, const fl_population *pop);

/*
 * Replaces the connector pool of the uploaded population by a longer one (same base tables, more
 * adjust_scores variants appended).  Energies already computed stay valid.  The reference rescales a private
 * copy of the tables inside every get_placement call (organism_object.py:1303,1312-1329).
 */
int fl_update_connectors(fl_ctx *ctx, const double *con_pool, long long n_con);

/*
 * Places every organism on every sequence of the set (replaces the loop of
 * OrganismObject.get_binding_energies over get_placement, organism_object.py:958-966,1263-1357, i.e.
 * n_org*n_seq calls of _multiplacement.calculate, _interface.c:45).
 *   con_begin[class*n_org + o] = offset into con_pool of the connector tables organism o uses for
 *                                sequences of that class (base tables, or an adjust_scores variant).
 * Energies stay resident on the device for fl_fitness.  Host outputs may each be NULL:
 *   energies[o*n_seq + s]                       (rec_scores[n] of calculate)
 *   gaps[(o*n_seq+s)*max_rec + i]               (con_lengths of calculate: start, then gaps)
 *   rec_scores / con_scores likewise            (max_rec = largest n_rec in the population)
 * gaps/rec_scores/con_scores require FL_WANT_TRACE.
 */
int fl_place_batch(fl_ctx *ctx, int set_id, const long long *con_begin, int flags, double *energies, int *gaps,
                   double *rec_scores, double *con_scores);

/*
 * Per-organism fitness statistic from the resident energies of two sets (replaces
 * OrganismObject.get_M_and_SE_on_DNA_set + set_fitness, organism_object.py:848-956).
 * out[o*5 + {0..4}] = fitness, M_pos, SE_pos, M_neg, SE_neg.
 */
int fl_fitness(fl_ctx *ctx, int pos_set, int neg_set, int method, double gamma, double *out);

/*
 * The same reduction, but the statistics stay on the device: *out_dev is a context-owned buffer of
 * max(n_org, pad_rows) x 5 doubles (rows beyond n_org are zero), valid until the next fl_fitness* call.
 * Asynchronous on fl_ctx_stream(); no device-to-host copy.  This is the send buffer of the one collective of the
 * multi-GPU path: every rank all-gathers its equally padded block in place (NCCL), replacing the reference's
 * comm.gather of pickled organisms and fitness lists (src/search_organisms.py:268-277).  Device errors (NaN null
 * bin) surface at the next synchronising call (fl_sync).
 */
int fl_fitness_dev(fl_ctx *ctx, int pos_set, int neg_set, int method, double gamma, int pad_rows, void **out_dev);

/*
 * Host code: recomputes column 0 of n_org rows [fitness, M_pos, SE_pos, M_neg, SE_neg] from the other four exactly as
 * OrganismObject.set_fitness does (organism_object.py:954-956: `(M_p - M_n) / (SE_p ** 2 + SE_n ** 2) ** (1/2)` on numpy
 * scalars, i.e. with libm pow()).  fl_fitness applies it before returning; callers that take the statistics from
 * fl_fitness_dev (the multi-GPU gather) apply it to the gathered rows.  M and SE come off the device bit-identical to
 * numpy's np.mean / np.std (pairwise summation order, csrc/fitness_kernel.cuh).
 */
int fl_fitness_finalize(double *stats, long long n_org);

/*
 * One placement with the exact argument list of _multiplacement.calculate (src/_interface.c:45-185,
 * kwlist :49-52): sequence (+length), rec_types (+n_rec), rec_matrices, rec_lengths, con_matrices,
 * rec_scores (out, n+1), con_scores (out, n-1), con_lengths (out, n), max_length, bin_freqs, bin_edges,
 * num_bins.  n_con_doubles is the length of con_matrices (the reference reads it from the buffer shape,
 * _interface.c:131).
 */
int fl_calculate(fl_ctx *ctx, const char *sequence, int seq_len, const char *rec_types, int n_rec,
                 const double *rec_matrices, const int *rec_lengths, const double *con_matrices,
                 long long n_con_doubles, double *rec_scores, double *con_scores, int *con_lengths, int max_length,
                 const double *bin_freqs, const double *bin_edges, const int *num_bins);

/*
 * Since the last call: pairs routed to the integer filter pass; how many of those it handed to the exact kernel in the
 * end (too many ties, unknown bytes, tables outside its fixed-point range); and how many took the int32 second chance
 * after the packed 16-bit pass had let too many near-ties through (may be NULL).  Synchronises.
 */
int fl_path_stats(fl_ctx *ctx, long long *fast_pairs, long long *fallback_pairs, long long *second_chance_pairs);

/*
 * ---- table builders (host code, no device work, no context) -------------------------------------------------------
 * The input producers of the path, batched and threaded; in the reference they are per-object Python loops.  They
 * return the ARGUMENTS of the logarithms: the reference takes those with numpy's log2 / log, whose vector kernels are
 * not bit-identical to libm's, so the Python layer applies numpy's function to the whole batch in one call.  Phi is
 * evaluated with libm erf() (what math.erf calls) in the reference's operand order.
 *
 * fl_build_connector_tables replaces ConnectorObject.set_precomputed_pdfs_cdfs (connector_object.py:175-206) for n_con
 * connectors of table length M:  pdf_arg[c*M + d] = Phi(d+.5) - Phi(d-.5) + pc,
 * cdf_arg[c*M + d] = Phi(d-.5) - Phi(-.5) + pc*(d+1);  stored_pdfs / stored_cdfs = log2 of these.
 * n_threads <= 0: all host cores.
 */
int fl_build_connector_tables(const double *mu, const double *sigma, int n_con, int M, double pseudo_count,
                              double *pdf_arg, double *cdf_arg, int n_threads);
/*
 * In-place Python round(x, ndigits) (correctly rounded decimal, ties on the exact binary value): the rounding of
 * PssmObject.recalculate_pssm (pssm_object.py:368-400), pssm = round(log2(4 p + pc), 2).
 */
int fl_round_decimals(double *values, long long n, int ndigits);
/*
 * Replaces the arithmetic of ShapeObject.set_alt_model (shape_object.py:156-180) for n_shapes recognizers whose bin
 * edges are edges[edge_begin[s] .. +n_edges[s]):  auc_arg[s] = Phi(e_last) - Phi(e_0) + (n_edges-1)*pc and, packed one
 * entry per bin (row s starts at edge_begin[s] - s),  mass_arg = Phi(e[b+1]) - Phi(e[b]) + pc;
 * alt[b] = ln(mass_arg) - ln(auc_arg).  mu must already be clamped to [e_0, e_last] as the reference does first.
 */
int fl_build_shape_masses(const double *mu, const double *sigma, const double *edges, const long long *edge_begin,
                          const int *n_edges, int n_shapes, double pseudo_count, double *mass_arg, double *auc_arg);

/*
 * Measures, on this device and now, the issue rate of the DPX instruction the DP stage consists of (VIADDMNMX:
 * max(a + b, c), one per DP cell; as 32-bit and as packed 2 x 16-bit cells) in DP cells per second, and reports the
 * SM count and the maximum SM clock.  bench.py divides its achieved cell rate by this (the ALU-bound roofline of
 * SURVEY.md 8d); the reference has no counterpart (its DP is the scalar loop organism.c:347-371).
 */
int fl_measure_dpx_rate(fl_ctx *ctx, double *cells_per_s_i32, double *cells_per_s_i16x2, int *n_sm, int *sm_clock_khz);

/* Counters for bench.py: kernels launched by this context since creation, and the last launch geometry. */
long long fl_launch_count(fl_ctx *ctx);

#ifdef __cplusplus
}
#endif
#endif /* FLEMINGO_B200_H */
