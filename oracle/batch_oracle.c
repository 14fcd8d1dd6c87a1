/* oracle/batch_oracle.c -- CPU ORACLE, threaded batch drivers.  TEST INFRASTRUCTURE ONLY.
 * One sequence per thread, all host cores: this is the "reference CPU path" timed beside the GPU
 * numbers -- the serial per-sequence calls the reference makes
 * (src/scoring.cc:43,58,65), farmed out over sequences like its production_run.sh farms out seeds.
 * pthreads (not OpenMP) so it builds with any gcc in the image. */
#include "fold_oracle.h"
#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

int orc_num_threads(void) { long n = sysconf(_SC_NPROCESSORS_ONLN); return n > 0 ? (int)n : 1; }

typedef struct {
  int kind; /* 0 pf, 1 mfe, 2 score */
  const orc_model *M; const orc_problem *P;
  int B, n; const char *seqs; const char *cst; const orc_motif *mo; int float_dg;
  double *dG, *bppm, *score, *terms; int32_t *e; char *structs;
  int next; int rc; pthread_mutex_t mu;
} job_t;

static void *worker(void *arg) {
  job_t *J = (job_t *)arg;
  int n = J->n;
  size_t tsz = (size_t)n * (n + 1) / 2;
  char *s = (char *)malloc(n + 1), *st = (char *)malloc(n + 1);
  for (;;) {
    pthread_mutex_lock(&J->mu);
    int b0 = J->next; J->next += 4;
    pthread_mutex_unlock(&J->mu);
    if (b0 >= J->B) break;
    for (int b = b0; b < b0 + 4 && b < J->B; b++) {
      memcpy(s, J->seqs + (size_t)b * n, n); s[n] = 0;
      int rc = 0;
      if (J->kind == 0) rc = orc_pf(J->M, n, s, J->cst, J->mo, &J->dG[b], J->bppm ? J->bppm + (size_t)b * tsz : NULL);
      else if (J->kind == 1) { rc = orc_mfe(J->M, n, s, J->cst, J->mo, &J->e[b], st); if (J->structs) memcpy(J->structs + (size_t)b * n, st, n); }
      else {
        int nt = (J->P->n_terms + J->P->n_bp_terms) * (J->P->n_ctx > 0 ? J->P->n_ctx : 1);
        J->score[b] = orc_score(J->M, J->P, s, J->float_dg, J->terms ? J->terms + (size_t)b * nt : NULL);
      }
      if (rc) J->rc = rc;
    }
  }
  free(s); free(st);
  return NULL;
}

static int run(job_t *J, int threads) {
  if (threads <= 0) threads = orc_num_threads();
  if (threads > 256) threads = 256;
  pthread_t th[256];
  pthread_mutex_init(&J->mu, NULL);
  J->next = 0; J->rc = 0;
  for (int t = 0; t < threads; t++) pthread_create(&th[t], NULL, worker, J);
  for (int t = 0; t < threads; t++) pthread_join(th[t], NULL);
  pthread_mutex_destroy(&J->mu);
  return J->rc;
}

/* seqs: B x n upper-case chars (no terminators) */
int orc_pf_batch(const orc_model *M, int B, int n, const char *seqs, const char *cst, const orc_motif *mo,
                 double *dG, double *bppm, int threads) {
  job_t J; memset(&J, 0, sizeof J);
  J.kind = 0; J.M = M; J.B = B; J.n = n; J.seqs = seqs; J.cst = cst; J.mo = mo; J.dG = dG; J.bppm = bppm;
  return run(&J, threads);
}

int orc_mfe_batch(const orc_model *M, int B, int n, const char *seqs, const char *cst, const orc_motif *mo,
                  int32_t *e, char *structs, int threads) {
  job_t J; memset(&J, 0, sizeof J);
  J.kind = 1; J.M = M; J.B = B; J.n = n; J.seqs = seqs; J.cst = cst; J.mo = mo; J.e = e; J.structs = structs;
  return run(&J, threads);
}

int orc_score_batch(const orc_model *M, const orc_problem *P, int B, const char *seqs, int float_dg,
                    double *score, double *terms, int threads) {
  job_t J; memset(&J, 0, sizeof J);
  J.kind = 2; J.M = M; J.P = P; J.B = B; J.n = P->n; J.seqs = seqs; J.float_dg = float_dg; J.score = score; J.terms = terms;
  return run(&J, threads);
}
