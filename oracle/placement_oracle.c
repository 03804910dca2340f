/*
 * placement_oracle.c -- TEST INFRASTRUCTURE ONLY.
 *
 * A plain-C, single-threaded CPU restatement of FLEMINGO's placement path, used by tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline leg as the CHECKER for the CUDA kernels.  Nothing
 * under flemingo_b200/ may include, link or call this file.
 *
 * Parity status: PINNED.  oracle/_ref (the unmodified reference extension, compiled from
 * /root/reference by oracle/Makefile) is run against this restatement by tests/test_oracle_golden.py and
 * by tests/golden/make_golden.py; the committed fixtures under tests/golden/ are reference outputs.
 *
 * What it restates (reference file:line, relative to /root/reference/extensions/multiplacement/src):
 *   score rows      organism/recognizer.c:130-168 (PSSM), :187-589 (MGW/ProT/Roll/HelT), _aux.c:184-217
 *   gap energies    organism/connector.c:31-66 (precomputed branch), auc pick organism/organism.c:336-340,
 *                   the cdf pointer off-by-one of connector.c:3-4 / organism.c:191-206
 *   chain DP        organism/organism.c:321-393
 *   traceback       _aux.c:149-158, organism/organism.c:398-427
 *
 * It is NOT a transliteration: the gap energy is hoisted into a per-connector table G[g] (the reference
 * recomputes it for every (j,k) cell), rows are laid out per stage, and errors are return codes.  The
 * floating-point expressions keep the reference's operand order so every double is bit-identical.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define FO_OK            0
#define FO_TOO_LARGE    -1   /* sum of recognizer lengths > sequence length (organism.c:262) */
#define FO_NAN_NULL     -2   /* a NaN null-model bin was selected (reference calls exit(1), recognizer.c:258) */
#define FO_UNSUPPORTED  -3   /* would need the legacy non-precomputed connector maths (connector.c:70-72) */
#define FO_BAD_INPUT    -4

#define BIG_NEG (-1E10)
#define BIG_POS (1E10)

/* layout of the table blob (flemingo_b200/data/tables.bin) */
#define T_MGW  0
#define T_PROT 1024
#define T_ROLL 2048
#define T_HELT 4096
#define T_DEN  6144
#define T_TOTAL 16144

static int base_code(char c) {
  switch (c) {
    case 'a': case 'A': return 0;
    case 'g': case 'G': return 1;
    case 'c': case 'C': return 2;
    case 't': case 'T': return 3;
    default: return -1;   /* contributes nothing to a PSSM column, and 0 to a pentamer index */
  }
}

/* recognizer.c:130-168 */
static void row_pssm(const double *mat, int len, const char *seq, int n_pos, double *row) {
  for (int p = 0; p < n_pos; p++) {
    double s = 0.0;
    for (int c = 0; c < len; c++) {
      int b = base_code(seq[p + c]);
      if (b >= 0) s += mat[4 * c + b];
    }
    row[p] = s;
  }
}

static int pentamer_index(const char *seq) {
  int idx = 0;
  for (int k = 0; k < 5; k++) {
    int b = base_code(seq[k]);
    idx = idx * 4 + (b < 0 ? 0 : b);
  }
  return idx;
}

/* _aux.c:184-190: first bin whose [lo, hi) holds the score, otherwise the LAST bin */
static double bin_lookup(double score, const double *freq, const double *edges, int n_edges) {
  for (int b = 1; b < n_edges; b++)
    if (score >= edges[b - 1] && score < edges[b]) return freq[b - 1];
  return freq[n_edges - 2];
}

/* recognizer.c:187-589 with the averages of _aux.c:203-217 */
static int row_shape(char feat, int len, const double *null_f, const double *alt_f, const double *edges,
                     int n_edges, const double *tables, const char *seq, int n_pos, double *row) {
  int n_pent = len - 4;
  const double *tab;
  int two_step;
  switch (feat) {
    case 'm': case 'M': tab = tables + T_MGW;  two_step = 0; break;
    case 't': case 'T': tab = tables + T_PROT; two_step = 0; break;
    case 'r': case 'R': tab = tables + T_ROLL; two_step = 1; break;
    case 'h': case 'H': tab = tables + T_HELT; two_step = 1; break;
    default: return FO_BAD_INPUT;
  }
  if (n_pent < 1) return FO_BAD_INPUT;
  double *pent = (double *)malloc(sizeof(double) * 2 * n_pent);
  int rc = FO_OK;
  for (int p = 0; p < n_pos; p++) {
    double score;
    for (int j = 0; j < n_pent; j++) {
      int idx = pentamer_index(seq + p + j);
      pent[j] = tab[idx];
      if (two_step) pent[j + n_pent] = tab[idx + 1024];
    }
    if (two_step) {
      int m = 2 * n_pent;
      double sum = pent[0] + pent[m - 1];
      for (int i = 0; i < m; i++) sum += pent[i];
      score = sum / ((double)(m + 2));
    } else {
      double sum = 0.0;
      for (int i = 0; i < n_pent; i++) sum += pent[i];
      score = sum / ((double)n_pent);
    }
    double a = bin_lookup(score, alt_f, edges, n_edges);
    double n0 = bin_lookup(score, null_f, edges, n_edges);
    double r = a - n0;
    if (r < BIG_NEG) r = BIG_NEG;
    if (isnan(n0)) rc = FO_NAN_NULL;
    row[p] = r;
  }
  free(pent);
  return rc;
}

/* connector.c:35-66 hoisted to a table over the gap length */
static void gap_table(const double *pdf, double auc, int n_rec, int eff_len, int n_pos, const double *den_tab,
                      double *G) {
  for (int g = 0; g < n_pos; g++) {
    if (g + n_rec + 1 <= eff_len) {
      double num = pdf[g];
      if (num < BIG_NEG) num = BIG_NEG;
      num -= auc;
      double den = (den_tab[(eff_len - (g + 1)) - 1] - den_tab[(n_rec - 1) - 1] -
                    den_tab[eff_len - (g + 1) - (n_rec - 1) - 1]) -
                   (den_tab[eff_len - 1] - den_tab[n_rec - 1] - den_tab[eff_len - n_rec - 1]);
      double res = num - den;
      if (res < BIG_NEG) res = BIG_NEG;
      else if (res > BIG_POS) res = BIG_POS;
      G[g] = res;
    } else {
      G[g] = BIG_NEG;
    }
  }
}

/*
 * One placement.  Arguments have the meaning of _multiplacement.calculate (_interface.c:45-185):
 * rec_matrices = concatenated PSSMs [col][A,G,C,T]; con_matrices = per connector [mu, sigma, pdf[M], cdf[M]]
 * with max_length = M-1; bin_freqs = per shape recognizer null[nb-1] ++ alt[nb-1]; bin_edges = per shape
 * recognizer nb edges; num_bins = nb per shape recognizer.  Outputs: rec_scores[n+1] (last = energy),
 * con_scores[n-1], con_lengths[n] (first = start of recognizer 0, then the gap before recognizer i).
 */
int fo_place(const double *tables, const char *seq, int s_len, const char *rec_types, int n_rec,
             const double *rec_matrices, const int *rec_lengths, const double *con_matrices, int max_length,
             const double *bin_freqs, const double *bin_edges, const int *num_bins, double *rec_scores,
             double *con_scores, int *con_lengths) {
  if (n_rec < 1) return FO_BAD_INPUT;
  int sum_len = 0;
  for (int i = 0; i < n_rec; i++) sum_len += rec_lengths[i];
  if (sum_len > s_len) return FO_TOO_LARGE;
  int n_pos = s_len - sum_len + 1;
  int eff_len = s_len - sum_len + n_rec;
  if (n_rec > 1 && (max_length < 0 || s_len - sum_len > max_length)) return FO_UNSUPPORTED;
  if (n_rec > 1 && eff_len > 10000) return FO_UNSUPPORTED;  /* DEN_EXPANSION is only read through a connector */

  double *R = (double *)malloc(sizeof(double) * (size_t)n_rec * n_pos);       /* score rows */
  double *C = (double *)malloc(sizeof(double) * (size_t)n_rec * n_pos);       /* cumulative rows, one per stage */
  double *G = (double *)malloc(sizeof(double) * (size_t)(n_rec > 1 ? n_rec - 1 : 1) * n_pos);
  int *T = (int *)malloc(sizeof(int) * (size_t)n_rec * n_pos);                /* chosen gap per (stage, j) */
  int rc = FO_OK;

  int m_off = 0, a_off = 0, e_off = 0, shape_i = 0, fwd = 0;
  for (int i = 0; i < n_rec; i++) {
    double *row = R + (size_t)i * n_pos;
    if (rec_types[i] == 'p' || rec_types[i] == 'P') {
      row_pssm(rec_matrices + m_off, rec_lengths[i], seq + fwd, n_pos, row);
      m_off += 4 * rec_lengths[i];
    } else {
      int nb = num_bins[shape_i];
      int r = row_shape(rec_types[i], rec_lengths[i], bin_freqs + a_off, bin_freqs + a_off + (nb - 1),
                        bin_edges + e_off, nb, tables, seq + fwd, n_pos, row);
      if (r != FO_OK) rc = r;
      a_off += 2 * (nb - 1);
      e_off += nb;
      shape_i++;
    }
    fwd += rec_lengths[i];
  }
  if (rc == FO_BAD_INPUT) goto done;

  memcpy(C, R, sizeof(double) * n_pos);
  for (int i = 1; i < n_rec; i++) {
    const double *con = con_matrices + (size_t)(i - 1) * (2 * (max_length + 1) + 2);
    const double *pdf = con + 2;
    const double *cdf = pdf + max_length;       /* the reference's off-by-one (connector.c:3-4) */
    double auc = cdf[s_len - sum_len];
    double *Gi = G + (size_t)(i - 1) * n_pos;
    gap_table(pdf, auc, n_rec, eff_len, n_pos, tables + T_DEN, Gi);
    const double *prev = C + (size_t)(i - 1) * n_pos;
    double *cur = C + (size_t)i * n_pos;
    const double *Ri = R + (size_t)i * n_pos;
    int *Ti = T + (size_t)i * n_pos;
    for (int j = 0; j < n_pos; j++) {
      double best = -INFINITY;
      int best_gap = 0;                           /* calloc'ed in the reference: stays 0 if nothing wins */
      for (int k = 0; k <= j; k++) {
        double cand = prev[k] + Gi[j - k] + Ri[j];
        if (best < cand) { best = cand; best_gap = j - k; }
      }
      cur[j] = best;
      Ti[j] = best_gap;
    }
  }

  {
    const double *last = C + (size_t)(n_rec - 1) * n_pos;
    int m = 0;
    for (int j = 0; j < n_pos; j++) if (last[j] > last[m]) m = j;
    rec_scores[n_rec] = last[m];
    for (int i = n_rec - 1; i >= 1; i--) {
      int g = T[(size_t)i * n_pos + m];
      rec_scores[i] = R[(size_t)i * n_pos + m];
      /* when no candidate ever won (all NaN) the reference reports the calloc'ed 0.0 */
      con_scores[i - 1] = (C[(size_t)i * n_pos + m] == -INFINITY) ? 0.0 : G[(size_t)(i - 1) * n_pos + g];
      con_lengths[i] = g;
      m -= g;
    }
    rec_scores[0] = R[m];
    con_lengths[0] = m;
  }

done:
  free(R); free(C); free(G); free(T);
  return rc;
}

/*
 * Energies (and optionally full placements) of n_org organisms on n_seq sequences; plain loops.
 * Organism o's arrays start at the given offsets (prefix sums, n_org+1 entries each).
 * seq_off has n_seq+1 entries into seq_bytes.  con_stride_M[o] = M (table length) of organism o.
 * energies[o*n_seq+s]; if gaps != NULL: gaps[(o*n_seq+s)*max_rec + i], rscores likewise (+ energy not repeated),
 * cscores[(o*n_seq+s)*max_rec + i].  status[o*n_seq+s] = return code of the single placement.
 */
int fo_place_batch(const double *tables, int n_org, int n_seq, int max_rec, const char *seq_bytes,
                   const long long *seq_off, const char *types, const int *rec_off, const int *lengths,
                   const double *pssm, const long long *pssm_off, const double *con, const long long *con_off,
                   const int *max_length, const double *bin_freqs, const long long *freq_off,
                   const double *bin_edges, const long long *edge_off, const int *num_bins,
                   const int *shape_off, double *energies, int *gaps, double *rscores, double *cscores,
                   int *status) {
  double rs[64], cs[64];
  int cl[64];
  if (max_rec > 63) return FO_BAD_INPUT;
  for (int o = 0; o < n_org; o++) {
    int n = rec_off[o + 1] - rec_off[o];
    for (int s = 0; s < n_seq; s++) {
      size_t pair = (size_t)o * n_seq + s;
      int rc = fo_place(tables, seq_bytes + seq_off[s], (int)(seq_off[s + 1] - seq_off[s]), types + rec_off[o], n,
                        pssm + pssm_off[o], lengths + rec_off[o], con + con_off[o], max_length[o],
                        bin_freqs + freq_off[o], bin_edges + edge_off[o], num_bins + shape_off[o], rs, cs, cl);
      status[pair] = rc;
      if (rc == FO_OK || rc == FO_NAN_NULL) {
        energies[pair] = rs[n];
        if (gaps) {
          for (int i = 0; i < n; i++) { gaps[pair * max_rec + i] = cl[i]; rscores[pair * max_rec + i] = rs[i]; }
          for (int i = 0; i + 1 < n; i++) cscores[pair * max_rec + i] = cs[i];
        }
      } else {
        energies[pair] = NAN;
      }
    }
  }
  return FO_OK;
}
