/* oracle/design_oracle.c -- CPU ORACLE (score function + Monte Carlo).  TEST INFRASTRUCTURE ONLY.
 *
 * Restates the reference's design loop on top of fold_oracle.c:
 *   ScoreFunction::evaluate / evaluate_terms   reference src/scoring.cc:114-158
 *   MacrostateProbTerm::evaluate               reference src/scoring.cc:233-259
 *   MonteCarlo::apply                          reference src/sampling.cc:22-107
 *   can_be_(freely_)mutated, mutate_recursively, UnbiasedMutationMove   src/sampling.cc:170-303
 *   Fixed / Annealing thermostats              reference src/sampling.cc:305-338
 * The reference draws from three copies of one std::mt19937 through libstdc++ distributions
 * (sampling.cc:36-37), which no other implementation can reproduce; trajectories here use the
 * engine's documented counter-based convention instead (Philox4x32-10, DESIGN.md "RNG"), so the
 * kernel can be checked step for step.
 */
#include "fold_oracle.h"
#include <ctype.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

static void up(char *dst, const char *src, int n) {
  for (int i = 0; i < n; i++) { char c = (char)toupper((unsigned char)src[i]); dst[i] = (c == 'T') ? 'U' : c; }
  dst[n] = 0;
}

double orc_score(const orc_model *M, const orc_problem *P, const char *seq, int float_dg, double *terms) {
  int nctx = P->n_ctx > 0 ? P->n_ctx : 1, nT = P->n_terms + P->n_bp_terms;
  double kT = orc_kT_kcal(M), score = 0.;
  for (int x = 0; x < nctx; x++) {
    const char *bef = P->n_ctx > 0 ? P->ctx_before[x] : "", *aft = P->n_ctx > 0 ? P->ctx_after[x] : "";
    int nb = (int)strlen(bef), na = (int)strlen(aft), n = nb + P->n + na;
    char *full = (char *)malloc(n + 1), *cst = (char *)malloc(n + 1);
    up(full, bef, nb); up(full + nb, seq, P->n); up(full + nb + P->n, aft, na);
    double g_tot[2] = {NAN, NAN};
    for (int t = 0; t < P->n_terms; t++) {
      int cond = P->term_cond[t];
      const orc_motif *mo = (cond == 1 && P->motif.seq) ? &P->motif : NULL;
      if (isnan(g_tot[cond])) orc_pf(M, n, full, NULL, mo, &g_tot[cond], NULL);
      memset(cst, '.', n); cst[n] = 0;
      memcpy(cst + nb, P->macro_cst[P->term_macro[t]], P->n);
      double g_act, g0 = g_tot[cond];
      orc_pf(M, n, full, cst, mo, &g_act, NULL);
      if (float_dg) { g0 = (double)(float)g0; g_act = (double)(float)g_act; }
      double p = exp((g0 - g_act) / kT);
      if (!P->term_favorable[t]) p = 1. - p;
      double v = log(p);
      if (terms) terms[x * nT + t] = v;
      score += P->term_weight[t] * v;
    }
    /* terms over base-pair probabilities: full bppm of the unconstrained ensemble of the term's condition, read the way
     * ViennaRnaFold::base_pair_prob reads it (reference src/scoring.cc:37-51) */
    double *bp[2] = {NULL, NULL};
    for (int t = 0; t < P->n_bp_terms; t++) {
      const orc_bp_term *b = &P->bp_terms[t];
      int cond = b->cond;
      if (!bp[cond]) {
        double g;
        bp[cond] = (double *)calloc((size_t)n * (n + 1) / 2, sizeof(double));
        orc_pf(M, n, full, NULL, (cond == 1 && P->motif.seq) ? &P->motif : NULL, &g, bp[cond]);
      }
#define ORC_TRI(a_, b_) ((size_t)(a_) * n - (size_t)(a_) * ((a_) - 1) / 2 + ((b_) - (a_)))
      double xv = 0.;
      if (b->kind == 0) {
        int a = b->i + nb, c = b->j + nb;
        if (a > c) { int q = a; a = c; c = q; }
        xv = a == c ? 0. : bp[cond][ORC_TRI(a, c)];
      } else {
        for (int q = 0; q < b->n_pos; q++) {
          int p = b->pos[q] + nb;
          double s = 0.;
          for (int k = 0; k < p; k++) s += bp[cond][ORC_TRI(k, p)];
          for (int k = p + 1; k < n; k++) s += bp[cond][ORC_TRI(p, k)];
          xv += s;
        }
        xv /= b->n_pos;
      }
      if (!b->favorable) xv = 1. - xv;
      double v = log(xv);
      if (terms) terms[x * nT + P->n_terms + t] = v;
      score += b->weight * v;
    }
    free(bp[0]); free(bp[1]);
    free(full); free(cst);
  }
  return score;
}

/* ---------------------------------------------------------------- Philox4x32-10 ------------- */
void orc_philox4x32(uint32_t k0, uint32_t k1, const uint32_t ctr[4], uint32_t out[4]) {
  uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
  for (int r = 0; r < 10; r++) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1, n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

/* ---------------------------------------------------------------- moves --------------------- */
static char comp(char c) { return c == 'A' ? 'U' : c == 'U' ? 'A' : c == 'G' ? 'C' : c == 'C' ? 'G' : 0; }

static int mutate_rec(const orc_problem *P, char *seq, int pos, char mut, unsigned char *done) {
  seq[pos] = mut; done[pos] = 1;
  for (int m = 0; m < P->n_macro; m++) {
    const char *ms = P->macro_cst[m];
    char open, close; int step;
    if (ms[pos] == '(') { open = '('; close = ')'; step = 1; }
    else if (ms[pos] == ')') { open = ')'; close = '('; step = -1; }
    else continue;
    int level = 1, partner = pos;
    while (level) {
      partner += step;
      if (partner < 0 || partner >= P->n) return -1;            /* mismatched base-pair */
      level += (ms[partner] == open); level -= (ms[partner] == close);
    }
    if (!isupper((unsigned char)seq[partner])) return -2;       /* partner not mutable */
    char c = comp(mut);
    if (!done[partner]) { int rc = mutate_rec(P, seq, partner, c, done); if (rc) return rc; }
    else if (seq[partner] != c) return -3;                      /* unsatisfiable */
  }
  return 0;
}

int orc_mutate(const orc_problem *P, char *seq, int pos, char nucleotide) {
  unsigned char *done = (unsigned char *)calloc(P->n, 1);
  int rc = mutate_rec(P, seq, pos, nucleotide, done);
  free(done);
  return rc;
}

static int mc_run_impl(const orc_model *M, const orc_problem *P, uint64_t seed, uint64_t traj, int n_steps, int tk, const double *ta, int float_dg,
                       char *final_seq, double *final_score, char *best_seq, double *best_score, int32_t *best_step, uint32_t *outcomes, double *trace,
                       int interval, double *rec_vals, char *rec_seqs);

int orc_mc_run(const orc_model *M, const orc_problem *P, uint64_t seed, uint64_t traj, int n_steps, int tk, const double *ta, int float_dg,
               char *final_seq, double *final_score, char *best_seq, double *best_score, int32_t *best_step, uint32_t *outcomes, double *trace) {
  return mc_run_impl(M, P, seed, traj, n_steps, tk, ta, float_dg, final_seq, final_score, best_seq, best_score, best_step, outcomes, trace, 0, NULL, NULL);
}

int orc_mc_run_trace(const orc_model *M, const orc_problem *P, uint64_t seed, uint64_t traj, int n_steps, int tk, const double *ta, int float_dg,
                     char *final_seq, double *final_score, char *best_seq, double *best_score, int32_t *best_step, uint32_t *outcomes,
                     int interval, double *rec_vals, char *rec_seqs) {
  return mc_run_impl(M, P, seed, traj, n_steps, tk, ta, float_dg, final_seq, final_score, best_seq, best_score, best_step, outcomes, NULL, interval,
                     rec_vals, rec_seqs);
}

static int mc_run_impl(const orc_model *M, const orc_problem *P, uint64_t seed, uint64_t traj, int n_steps, int tk, const double *ta, int float_dg,
                       char *final_seq, double *final_score, char *best_seq, double *best_score, int32_t *best_step, uint32_t *outcomes, double *trace,
                       int interval, double *rec_vals, char *rec_seqs) {
  int n = P->n, nfree = 0;
  int nterms = (P->n_ctx > 0 ? P->n_ctx : 1) * (P->n_terms + P->n_bp_terms);
  double *table = (double *)calloc(nterms > 0 ? nterms : 1, sizeof(double));      /* the reference's step.score_table */
  double last_prop, last_diff_rec = 0., last_crit = 0., last_u = 0.;
  int *freep = (int *)malloc(sizeof(int) * n);
  for (int i = 0; i < n; i++) {
    if (!isupper((unsigned char)P->init_seq[i])) continue;
    int ok = 1;
    for (int m = 0; m < P->n_macro; m++) if (P->macro_cst[m][i] == ')') ok = 0;
    if (ok) freep[nfree++] = i;
  }
  char *cur = (char *)malloc(n + 1), *prop = (char *)malloc(n + 1), *best = (char *)malloc(n + 1);
  unsigned char *done = (unsigned char *)malloc(n);
  memcpy(cur, P->init_seq, n); cur[n] = 0;
  double cur_s = orc_score(M, P, cur, float_dg, table), bs = cur_s;
  last_prop = cur_s;
  memcpy(best, cur, n + 1);
  int bstep = -1, rc = 0;
  uint32_t oc[4] = {0, 0, 0, 0};
  /* auto-scaling thermostat (reference src/sampling.cc:371-401): every step appends the most recent score difference to a
   * training set; once it holds `period` values the temperature becomes max(median / ln(rate), 0) and the set is cleared */
  int period = tk == 2 ? (int)ta[1] : 0, ntrain = 0;
  double *train = tk == 2 ? (double *)malloc(sizeof(double) * (period > 0 ? period : 1)) : NULL;
  double T_auto = tk == 2 ? ta[2] : 0., last_diff = 0.;
  for (int i = 0; i < n_steps && nfree > 0; i++) {
    double T = (tk == 1) ? ((ta[1] - ta[0]) / (int)ta[2]) * (i % (int)ta[2]) + ta[0] : ta[0];
    if (tk == 2) {
      train[ntrain++] = last_diff;
      if (ntrain >= period) {
        int mid = ntrain / 2;
        /* the value nth_element leaves at position mid = the mid-th smallest */
        for (int a = 0; a < ntrain; a++) {
          int rank = 0;
          for (int b = 0; b < ntrain; b++) rank += (train[b] < train[a]) || (train[b] == train[a] && b < a);
          if (rank == mid) { double t = train[a] / log(ta[0]); T_auto = t > 0. ? t : 0.; break; }
        }
        ntrain = 0;
      }
      T = T_auto;
    }
    uint32_t ctr[4] = {(uint32_t)i, 0u, (uint32_t)traj, (uint32_t)(traj >> 32)}, r[4];
    orc_philox4x32((uint32_t)seed, (uint32_t)(seed >> 32), ctr, r);
    int pos = freep[(uint32_t)(((uint64_t)r[0] * (uint64_t)nfree) >> 32)];
    char nt = "ACGU"[r[1] >> 30];
    double u = ((double)(r[2] >> 5) * 67108864. + (double)(r[3] >> 6)) / 9007199254740992.;
    memcpy(prop, cur, n + 1);
    memset(done, 0, n);
    rc = mutate_rec(P, prop, pos, nt, done);
    if (rc) break;
    int outcome;
    if (!memcmp(prop, cur, n)) outcome = 2;
    else {
      double ps = orc_score(M, P, prop, float_dg, table), d = ps - cur_s, crit = exp(d / T);
      last_diff = d;
      last_prop = ps; last_diff_rec = d; last_crit = crit; last_u = u;
      if (crit < u) outcome = 0;
      else {
        outcome = (d > 0) ? 3 : 1;
        memcpy(cur, prop, n + 1); cur_s = ps;
        if (cur_s > bs) { bs = cur_s; bstep = i; memcpy(best, cur, n + 1); }
      }
    }
    oc[outcome]++;
    if (trace) trace[i] = cur_s;
    if (interval > 0 && i % interval == 0) {
      double *v = rec_vals + (size_t)(i / interval) * (8 + nterms);
      char *q = rec_seqs + (size_t)(i / interval) * 2 * n;
      v[0] = i; v[1] = cur_s; v[2] = last_prop; v[3] = last_diff_rec; v[4] = T; v[5] = last_crit; v[6] = last_u; v[7] = outcome;
      memcpy(v + 8, table, sizeof(double) * nterms);
      memcpy(q, cur, n); memcpy(q + n, prop, n);
    }
  }
  free(table);
  if (final_seq) memcpy(final_seq, cur, n);
  if (final_score) *final_score = cur_s;
  if (best_seq) memcpy(best_seq, best, n);
  if (best_score) *best_score = bs;
  if (best_step) *best_step = bstep;
  if (outcomes) memcpy(outcomes, oc, sizeof oc);
  free(freep); free(cur); free(prop); free(best); free(done); free(train);
  return rc;
}
