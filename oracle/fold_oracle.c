/* oracle/fold_oracle.c -- CPU ORACLE.  TEST INFRASTRUCTURE, NOT PRODUCT CODE.  See fold_oracle.h.
 *
 * Restates, in plain C, what the reference's ViennaRNA calls compute:
 *   - energy model: Turner-2004 nearest-neighbour rules, dangles=2, special hairpins, MAXLOOP=30,
 *     TURN=3, 37 C (vrna_md_set_default, reference src/scoring.cc:80-83)
 *   - McCaskill inside (vrna_pf, compute_bpp=0; src/scoring.cc:58,65) -- loop order j outer,
 *     i inner descending with rolling qq/qqm rows, like libRNA's pf_linear
 *   - outside / bppm (vrna_pf, compute_bpp=1; src/scoring.cc:43) -- l descending with the
 *     prml / prm_l / prm_MLb rolling arrays, like libRNA's pf_create_bppm
 *   - dot-bracket hard constraints with enforced pairs (vrna_constraints_add, src/scoring.cc:61-62)
 *   - hairpin ligand motif soft constraint + bppm correction (vrna_sc_add_hi_motif, :94-99)
 *   - Zuker MFE + traceback (no call site in the reference; BASELINE.json asks for it)
 * libRNA's source is not on this machine; this is a restatement of its published algorithm.
 * PARITY UNPINNED for exact values; thresholds of reference tests/test_scoring.cc are pinned.
 */
#include "fold_oracle.h"
#include <ctype.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define INF ORC_INF
#define MAXLOOP 30
#define TURN 3
#define K0 273.15
#define GASCONST 1.98717
#define MIN2(a, b) ((a) < (b) ? (a) : (b))
#define MAX2(a, b) ((a) > (b) ? (a) : (b))

static char g_err[512];
const char *orc_error(void) { return g_err; }

typedef struct {
  int stack[8][8];
  int hairpin[31], bulge[31], interior[31];
  int mmH[8][5][5], mmI[8][5][5], mm1nI[8][5][5], mm23I[8][5][5], mmM[8][5][5], mmExt[8][5][5];
  int dangle5[8][5], dangle3[8][5];
  int int11[8][8][5][5];
  int int21[8][8][5][5][5];
  int int22[8][8][5][5][5][5];
  int ML_base, ML_closing, ML_intern, ninio, max_ninio, TerminalAU, DuplexInit;
  double lxc;
  int n_tetra, n_tri, n_hex;
  char tetra[64][8], tri[64][8], hex[64][12];
  int tetraE[64], triE[64], hexE[64];
} raw_par;

typedef struct {
  double kT; /* cal/mol */
  double pf_scale;
  double stack[8][8];
  double hairpin[31], bulge[31], interior[31];
  double mmH[8][5][5], mmI[8][5][5], mm1nI[8][5][5], mm23I[8][5][5], mmM[8][5][5], mmExt[8][5][5];
  double dangle5[8][5], dangle3[8][5];
  double int11[8][8][5][5];
  double int21[8][8][5][5][5];
  double int22[8][8][5][5][5][5];
  double ninio[MAXLOOP + 1];
  double MLbase, MLclosing, MLintern, TermAU;
  double tetra[64], tri[64], hex[64];
} pf_par;

struct orc_model {
  raw_par r;    /* file values                                   */
  raw_par m;    /* MFE view: dangles / mmM / mmExt clamped to <= 0 */
  pf_par p;     /* Boltzmann factors                              */
};

static const int rtype[8] = {0, 2, 1, 4, 3, 6, 5, 7};
/* pair[a][b], codes A1 C2 G3 U4 */
static const int PAIR[5][5] = {{0, 0, 0, 0, 0}, {0, 0, 0, 0, 5}, {0, 0, 0, 1, 0}, {0, 0, 2, 0, 3}, {0, 6, 0, 4, 0}};

static int enc(char c) {
  switch (toupper((unsigned char)c)) {
    case 'A': return 1;
    case 'C': return 2;
    case 'G': return 3;
    case 'U': case 'T': return 4;
    default: return 0;
  }
}

/* ------------------------------------------------------------------ .par parsing ------------ */
typedef struct { double *v; int n, cap; } numvec;
static void nv_push(numvec *a, double x) {
  if (a->n == a->cap) { a->cap = a->cap ? 2 * a->cap : 1024; a->v = (double *)realloc(a->v, sizeof(double) * a->cap); }
  a->v[a->n++] = x;
}

static int fill_pairs_nuc(int *dst, const numvec *nv, int dims_after, const char *name) {
  /* tables indexed [type 1..6|7][...5^k] : accept 6 or 7 leading blocks */
  int blk = 1;
  for (int i = 0; i < dims_after; i++) blk *= 5;
  int nb = nv->n / blk;
  if (nv->n % blk || (nb != 6 && nb != 7)) { snprintf(g_err, sizeof g_err, "section %s: %d numbers", name, nv->n); return -1; }
  for (int t = 0; t < nb; t++)
    for (int k = 0; k < blk; k++) dst[(t + 1) * blk + k] = (int)nv->v[t * blk + k];
  if (nb == 6) for (int k = 0; k < blk; k++) dst[7 * blk + k] = 0;
  return 0;
}

static int fill_pairs2_nuc(int *dst, const numvec *nv, int dims_after, const char *name) {
  int blk = 1;
  for (int i = 0; i < dims_after; i++) blk *= 5;
  int nb2 = nv->n / blk, nb = 0;
  if (nv->n % blk == 0) { if (nb2 == 36) nb = 6; else if (nb2 == 49) nb = 7; }
  if (!nb) { snprintf(g_err, sizeof g_err, "section %s: %d numbers", name, nv->n); return -1; }
  for (int a = 0; a < nb; a++)
    for (int b = 0; b < nb; b++)
      for (int k = 0; k < blk; k++) dst[((a + 1) * 8 + (b + 1)) * blk + k] = (int)nv->v[(a * nb + b) * blk + k];
  return 0;
}

static int parse_par(const char *path, raw_par *R) {
  FILE *f = fopen(path, "r");
  if (!f) { snprintf(g_err, sizeof g_err, "cannot open %s", path); return -1; }
  fseek(f, 0, SEEK_END);
  long sz = ftell(f);
  fseek(f, 0, SEEK_SET);
  char *buf = (char *)malloc(sz + 2);
  if (fread(buf, 1, sz, f) != (size_t)sz) { fclose(f); free(buf); snprintf(g_err, sizeof g_err, "read error"); return -1; }
  fclose(f);
  buf[sz] = '\n'; buf[sz + 1] = 0;
  /* blank out C comments */
  for (char *p = buf; (p = strstr(p, "/*"));) {
    char *q = strstr(p, "*/");
    if (!q) break;
    for (char *c = p; c < q + 2; c++) if (*c != '\n') *c = ' ';
    p = q + 2;
  }
  memset(R, 0, sizeof *R);
  int rc = 0, seen = 0;
  char *line = buf;
  char sec[64] = "";
  numvec nv = {0, 0, 0};
  char loops[64][16]; int loopE[64]; int nloops = 0;
  for (;;) {
    char *eol = line ? strchr(line, '\n') : NULL;
    int at_end = (eol == NULL);
    if (!at_end) *eol = 0;
    int new_sec = at_end || (line[0] == '#' && line[1] != '#');
    if (new_sec && sec[0]) { /* flush */
      seen++;
      if (!strcmp(sec, "stack")) {
        int nb = nv.n == 49 ? 7 : nv.n == 36 ? 6 : 0;
        if (!nb) { snprintf(g_err, sizeof g_err, "stack: %d numbers", nv.n); rc = -1; }
        for (int a = 0; a < nb; a++) for (int b = 0; b < nb; b++) R->stack[a + 1][b + 1] = (int)nv.v[a * nb + b];
      } else if (!strcmp(sec, "mismatch_hairpin")) rc |= fill_pairs_nuc(&R->mmH[0][0][0], &nv, 2, sec);
      else if (!strcmp(sec, "mismatch_interior")) rc |= fill_pairs_nuc(&R->mmI[0][0][0], &nv, 2, sec);
      else if (!strcmp(sec, "mismatch_interior_1n")) rc |= fill_pairs_nuc(&R->mm1nI[0][0][0], &nv, 2, sec);
      else if (!strcmp(sec, "mismatch_interior_23")) rc |= fill_pairs_nuc(&R->mm23I[0][0][0], &nv, 2, sec);
      else if (!strcmp(sec, "mismatch_multi")) rc |= fill_pairs_nuc(&R->mmM[0][0][0], &nv, 2, sec);
      else if (!strcmp(sec, "mismatch_exterior")) rc |= fill_pairs_nuc(&R->mmExt[0][0][0], &nv, 2, sec);
      else if (!strcmp(sec, "dangle5")) rc |= fill_pairs_nuc(&R->dangle5[0][0], &nv, 1, sec);
      else if (!strcmp(sec, "dangle3")) rc |= fill_pairs_nuc(&R->dangle3[0][0], &nv, 1, sec);
      else if (!strcmp(sec, "int11")) rc |= fill_pairs2_nuc(&R->int11[0][0][0][0], &nv, 2, sec);
      else if (!strcmp(sec, "int21")) rc |= fill_pairs2_nuc(&R->int21[0][0][0][0][0], &nv, 3, sec);
      else if (!strcmp(sec, "int22")) {
        if (nv.n != 9216) { snprintf(g_err, sizeof g_err, "int22: %d numbers (want 9216)", nv.n); rc = -1; }
        else {
          int q = 0;
          for (int a = 1; a <= 6; a++) for (int b = 1; b <= 6; b++)
            for (int i = 1; i <= 4; i++) for (int j = 1; j <= 4; j++) for (int k = 1; k <= 4; k++) for (int l = 1; l <= 4; l++)
              R->int22[a][b][i][j][k][l] = (int)nv.v[q++];
        }
      } else if (!strcmp(sec, "hairpin") || !strcmp(sec, "bulge") || !strcmp(sec, "interior")) {
        int *dst = !strcmp(sec, "hairpin") ? R->hairpin : !strcmp(sec, "bulge") ? R->bulge : R->interior;
        if (nv.n != 31) { snprintf(g_err, sizeof g_err, "%s: %d numbers", sec, nv.n); rc = -1; }
        else for (int i = 0; i < 31; i++) dst[i] = (int)nv.v[i];
      } else if (!strcmp(sec, "ML_params")) {
        if (nv.n < 6) { snprintf(g_err, sizeof g_err, "ML_params"); rc = -1; }
        else { R->ML_base = (int)nv.v[0]; R->ML_closing = (int)nv.v[2]; R->ML_intern = (int)nv.v[4]; }
      } else if (!strcmp(sec, "NINIO")) {
        if (nv.n < 3) { snprintf(g_err, sizeof g_err, "NINIO"); rc = -1; }
        else { R->ninio = (int)nv.v[0]; R->max_ninio = (int)nv.v[2]; }
      } else if (!strcmp(sec, "Misc")) {
        if (nv.n < 5) { snprintf(g_err, sizeof g_err, "Misc"); rc = -1; }
        else { R->DuplexInit = (int)nv.v[0]; R->TerminalAU = (int)nv.v[2]; R->lxc = nv.v[4]; }
      } else if (!strcmp(sec, "Tetraloops")) {
        R->n_tetra = nloops; for (int i = 0; i < nloops; i++) { strncpy(R->tetra[i], loops[i], 7); R->tetraE[i] = loopE[i]; }
      } else if (!strcmp(sec, "Triloops")) {
        R->n_tri = nloops; for (int i = 0; i < nloops; i++) { strncpy(R->tri[i], loops[i], 7); R->triE[i] = loopE[i]; }
      } else if (!strcmp(sec, "Hexaloops")) {
        R->n_hex = nloops; for (int i = 0; i < nloops; i++) { strncpy(R->hex[i], loops[i], 11); R->hexE[i] = loopE[i]; }
      } else seen--; /* enthalpy sections etc. are ignored (37 C only) */
      nv.n = 0; nloops = 0; sec[0] = 0;
    }
    if (at_end || rc) break;
    if (line[0] == '#' && line[1] != '#') {
      char *s = line + 1;
      while (*s == ' ') s++;
      snprintf(sec, sizeof sec, "%s", s);
      for (char *e = sec + strlen(sec) - 1; e >= sec && isspace((unsigned char)*e); e--) *e = 0;
      if (!strcmp(sec, "END")) sec[0] = 0;
    } else if (sec[0]) {
      int is_loop = !strcmp(sec, "Tetraloops") || !strcmp(sec, "Triloops") || !strcmp(sec, "Hexaloops");
      char *s = line;
      if (is_loop) {
        char w[32]; int e, h;
        if (sscanf(s, "%31s %d %d", w, &e, &h) >= 2 && nloops < 64) { snprintf(loops[nloops], 16, "%s", w); loopE[nloops++] = e; }
      } else {
        for (char *tok = strtok(s, " \t\r"); tok; tok = strtok(NULL, " \t\r")) {
          if (!strcmp(tok, "INF")) nv_push(&nv, INF);
          else if (!strcmp(tok, "DEF")) nv_push(&nv, -50);
          else { char *end; double x = strtod(tok, &end); if (end != tok) nv_push(&nv, x); }
        }
      }
    }
    line = eol + 1;
  }
  free(nv.v); free(buf);
  if (!rc && seen < 18) { snprintf(g_err, sizeof g_err, "%s: only %d known sections", path, seen); rc = -1; }
  return rc;
}

/* SMOOTH() of libRNA's params.c, used for dangles / multi / exterior mismatches in the PF only */
static double smooth(double X) {
  const double SCALE = 10.;
  if (X / SCALE < -1.2283697) return 0.;
  if (X / SCALE > 0.8660254) return X;
  double s = sin(X / SCALE - 0.34242663) + 1.;
  return SCALE * 0.38490018 * s * s;
}

static void scale_params(orc_model *M) {
  raw_par *r = &M->r, *m = &M->m;
  pf_par *p = &M->p;
  *m = *r;
  for (int t = 0; t < 8; t++)
    for (int a = 0; a < 5; a++) {
      m->dangle5[t][a] = MIN2(0, r->dangle5[t][a]);
      m->dangle3[t][a] = MIN2(0, r->dangle3[t][a]);
      for (int b = 0; b < 5; b++) {
        m->mmM[t][a][b] = MIN2(0, r->mmM[t][a][b]);
        m->mmExt[t][a][b] = MIN2(0, r->mmExt[t][a][b]);
      }
    }
  double kT = (37. + K0) * GASCONST; /* cal/mol */
  p->kT = kT;
  p->pf_scale = exp(-(1.07 * (-185. + (37. - 37.) * 7.27)) / kT); /* no MFE estimate: SURVEY A.4 */
  if (p->pf_scale < 1.) p->pf_scale = 1.;
#define BF(E) exp(-(double)(E)*10. / kT)
  for (int i = 0; i <= 30; i++) { p->hairpin[i] = BF(r->hairpin[i]); p->bulge[i] = BF(r->bulge[i]); p->interior[i] = BF(r->interior[i]); }
  for (int i = 0; i <= MAXLOOP; i++) p->ninio[i] = BF(MIN2(r->max_ninio, i * r->ninio));
  p->MLbase = BF(r->ML_base); p->MLclosing = BF(r->ML_closing); p->MLintern = BF(r->ML_intern); p->TermAU = BF(r->TerminalAU);
  for (int i = 0; i < r->n_tetra; i++) p->tetra[i] = BF(r->tetraE[i]);
  for (int i = 0; i < r->n_tri; i++) p->tri[i] = BF(r->triE[i]);
  for (int i = 0; i < r->n_hex; i++) p->hex[i] = BF(r->hexE[i]);
  for (int a = 0; a < 8; a++)
    for (int b = 0; b < 8; b++) {
      p->stack[a][b] = BF(r->stack[a][b]);
      for (int i = 0; i < 5; i++) for (int j = 0; j < 5; j++) {
        p->int11[a][b][i][j] = BF(r->int11[a][b][i][j]);
        for (int k = 0; k < 5; k++) {
          p->int21[a][b][i][j][k] = BF(r->int21[a][b][i][j][k]);
          for (int l = 0; l < 5; l++) p->int22[a][b][i][j][k][l] = BF(r->int22[a][b][i][j][k][l]);
        }
      }
    }
  for (int t = 0; t < 8; t++)
    for (int a = 0; a < 5; a++) {
      p->dangle5[t][a] = exp(smooth(-(double)r->dangle5[t][a]) * 10. / kT);
      p->dangle3[t][a] = exp(smooth(-(double)r->dangle3[t][a]) * 10. / kT);
      for (int b = 0; b < 5; b++) {
        p->mmH[t][a][b] = BF(r->mmH[t][a][b]);
        p->mmI[t][a][b] = BF(r->mmI[t][a][b]);
        p->mm1nI[t][a][b] = BF(r->mm1nI[t][a][b]);
        p->mm23I[t][a][b] = BF(r->mm23I[t][a][b]);
        p->mmM[t][a][b] = exp(smooth(-(double)r->mmM[t][a][b]) * 10. / kT);
        p->mmExt[t][a][b] = exp(smooth(-(double)r->mmExt[t][a][b]) * 10. / kT);
      }
    }
#undef BF
}

orc_model *orc_model_load(const char *path) {
  orc_model *M = (orc_model *)calloc(1, sizeof *M);
  if (parse_par(path, &M->r)) { free(M); return NULL; }
  scale_params(M);
  return M;
}
void orc_model_free(orc_model *M) { free(M); }
double orc_kT_kcal(const orc_model *M) { return M->p.kT / 1000.; }

/* ------------------------------------------------------------------ loop energies (int) ----- */
static int special_hp(const raw_par *P, const char *s /* 5' closing base */, int u, int *e) {
  if (u == 4) { for (int i = 0; i < P->n_tetra; i++) if (!strncmp(P->tetra[i], s, 6)) { *e = P->tetraE[i]; return 1; } }
  else if (u == 6) { for (int i = 0; i < P->n_hex; i++) if (!strncmp(P->hex[i], s, 8)) { *e = P->hexE[i]; return 1; } }
  else if (u == 3) { for (int i = 0; i < P->n_tri; i++) if (!strncmp(P->tri[i], s, 5)) { *e = P->triE[i]; return 1; } }
  return 0;
}
static int special_hp_idx(const raw_par *P, const char *s, int u) {
  if (u == 4) { for (int i = 0; i < P->n_tetra; i++) if (!strncmp(P->tetra[i], s, 6)) return i; }
  else if (u == 6) { for (int i = 0; i < P->n_hex; i++) if (!strncmp(P->hex[i], s, 8)) return i; }
  else if (u == 3) { for (int i = 0; i < P->n_tri; i++) if (!strncmp(P->tri[i], s, 5)) return i; }
  return -1;
}

static int E_hairpin(const raw_par *P, int u, int type, int si1, int sj1, const char *s) {
  int e = (u <= 30) ? P->hairpin[u] : P->hairpin[30] + (int)(P->lxc * log(u / 30.));
  if (u < 3) return e;
  int se;
  if (special_hp(P, s, u, &se)) return se;
  if (u == 3) return e + (type > 2 ? P->TerminalAU : 0);
  return e + P->mmH[type][si1][sj1];
}

static int E_intloop(const raw_par *P, int u1, int u2, int type, int type2, int si1, int sj1, int sp1, int sq1) {
  int nl = MAX2(u1, u2), ns = MIN2(u1, u2);
  if (nl == 0) return P->stack[type][type2];
  if (ns == 0) {
    int e = P->bulge[nl];
    if (nl == 1) e += P->stack[type][type2];
    else { if (type > 2) e += P->TerminalAU; if (type2 > 2) e += P->TerminalAU; }
    return e;
  }
  if (ns == 1) {
    if (nl == 1) return P->int11[type][type2][si1][sj1];
    if (nl == 2) return (u1 == 1) ? P->int21[type][type2][si1][sq1][sj1] : P->int21[type2][type][sq1][si1][sp1];
    return P->interior[nl + 1] + MIN2(P->max_ninio, (nl - ns) * P->ninio) + P->mm1nI[type][si1][sj1] + P->mm1nI[type2][sq1][sp1];
  }
  if (ns == 2) {
    if (nl == 2) return P->int22[type][type2][si1][sp1][sq1][sj1];
    if (nl == 3) return P->interior[5] + P->ninio + P->mm23I[type][si1][sj1] + P->mm23I[type2][sq1][sp1];
  }
  return P->interior[u1 + u2] + MIN2(P->max_ninio, (nl - ns) * P->ninio) + P->mmI[type][si1][sj1] + P->mmI[type2][sq1][sp1];
}

/* stem in a multiloop (ext=0) or the exterior loop (ext=1); s5 / s3 = -1 when the neighbour is missing */
static int E_stem(const raw_par *P, int type, int s5, int s3, int ext) {
  int e = 0;
  if (s5 >= 0 && s3 >= 0) e = ext ? P->mmExt[type][s5][s3] : P->mmM[type][s5][s3];
  else if (s5 >= 0) e = P->dangle5[type][s5];
  else if (s3 >= 0) e = P->dangle3[type][s3];
  if (type > 2) e += P->TerminalAU;
  if (!ext) e += P->ML_intern;
  return e;
}

/* ------------------------------------------------------------------ loop weights (PF) ------- */
static double X_hairpin(const orc_model *M, int u, int type, int si1, int sj1, const char *s) {
  const pf_par *p = &M->p;
  double q = (u <= 30) ? p->hairpin[u] : p->hairpin[30] * exp(-(M->r.lxc * log(u / 30.)) * 10. / p->kT);
  if (u < 3) return q;
  int k = special_hp_idx(&M->r, s, u);
  if (k >= 0) return u == 4 ? p->tetra[k] : u == 6 ? p->hex[k] : p->tri[k];
  if (u == 3) return (type > 2) ? q * p->TermAU : q;
  return q * p->mmH[type][si1][sj1];
}

static double X_intloop(const pf_par *p, int u1, int u2, int type, int type2, int si1, int sj1, int sp1, int sq1) {
  int ul = MAX2(u1, u2), us = MIN2(u1, u2);
  if (ul == 0) return p->stack[type][type2];
  if (us == 0) {
    double z = p->bulge[ul];
    if (ul == 1) z *= p->stack[type][type2];
    else { if (type > 2) z *= p->TermAU; if (type2 > 2) z *= p->TermAU; }
    return z;
  }
  if (us == 1) {
    if (ul == 1) return p->int11[type][type2][si1][sj1];
    if (ul == 2) return (u1 == 1) ? p->int21[type][type2][si1][sq1][sj1] : p->int21[type2][type][sq1][si1][sp1];
    return p->interior[ul + us] * p->mm1nI[type][si1][sj1] * p->mm1nI[type2][sq1][sp1] * p->ninio[ul - us];
  }
  if (us == 2) {
    if (ul == 2) return p->int22[type][type2][si1][sp1][sq1][sj1];
    if (ul == 3) return p->interior[5] * p->mm23I[type][si1][sj1] * p->mm23I[type2][sq1][sp1] * p->ninio[1];
  }
  return p->interior[ul + us] * p->mmI[type][si1][sj1] * p->mmI[type2][sq1][sp1] * p->ninio[ul - us];
}

static double X_stem(const pf_par *p, int type, int s5, int s3, int ext) {
  double e = 1.;
  if (s5 >= 0 && s3 >= 0) e = ext ? p->mmExt[type][s5][s3] : p->mmM[type][s5][s3];
  else if (s5 >= 0) e = p->dangle5[type][s5];
  else if (s3 >= 0) e = p->dangle3[type][s3];
  if (type > 2) e *= p->TermAU;
  if (!ext) e *= p->MLintern;
  return e;
}

/* ------------------------------------------------------------------ per-sequence context ---- */
typedef struct {
  int n;
  int *S;               /* 1..n, S[0]=S[n+1]=0 */
  char *str;            /* upper-cased, 1-based: str[1..n] */
  unsigned char *pt;    /* (n+2)^2 pair type incl. hard constraints, 0 = may not pair */
  unsigned char *cu;    /* can be unpaired, 1..n */
  int *upr;             /* run of cu starting at i; upr[n+1]=0 */
  int n_occ; int occ[64]; /* motif occurrences (1-based i) */
  int m_len; int m_c; double m_x; /* motif length, contribution dcal, Boltzmann factor */
  int n_mpairs; int mp_i[64], mp_j[64]; /* inner motif pairs as offsets from occurrence start */
} fctx;

#define IX(c, i, j) ((size_t)(i) * ((c)->n + 2) + (j))

static void ctx_free(fctx *c) { free(c->S); free(c->str); free(c->pt); free(c->cu); free(c->upr); }

static int pair_table(int n, const char *db, int *pt /* 1-based, 0 = unpaired */) {
  int *stk = (int *)malloc(sizeof(int) * (n + 1)), sp = 0, rc = 0;
  for (int i = 1; i <= n; i++) pt[i] = 0;
  for (int i = 1; i <= n && !rc; i++) {
    if (db[i - 1] == '(') stk[sp++] = i;
    else if (db[i - 1] == ')') { if (!sp) rc = -1; else { int k = stk[--sp]; pt[k] = i; pt[i] = k; } }
  }
  if (sp) rc = -1;
  free(stk);
  return rc;
}

int orc_motif_contribution(const orc_model *M, const orc_motif *mo) {
  int m = (int)strlen(mo->seq);
  char *alt = (char *)malloc(m + 1);
  memset(alt, '.', m); alt[0] = '('; alt[m - 1] = ')'; alt[m] = 0;
  /* vrna_eval_structure returns float kcal/mol; the correction is computed in float and truncated */
  float e_alt = (float)((double)orc_eval_structure(M, m, mo->seq, alt) / 100.);
  float e_mot = (float)((double)orc_eval_structure(M, m, mo->seq, mo->fold) / 100.);
  float diff = e_mot - e_alt;
  free(alt);
  return (int)(mo->energy_kcal * 100.) + (int)(diff * 100.);
}

static int ctx_init(fctx *c, const orc_model *M, int n, const char *seq, const char *cst, const orc_motif *mo) {
  memset(c, 0, sizeof *c);
  c->n = n;
  c->S = (int *)calloc(n + 2, sizeof(int));
  c->str = (char *)calloc(n + 3, 1);
  c->pt = (unsigned char *)calloc((size_t)(n + 2) * (n + 2), 1);
  c->cu = (unsigned char *)calloc(n + 2, 1);
  c->upr = (int *)calloc(n + 3, sizeof(int));
  for (int i = 1; i <= n; i++) { c->S[i] = enc(seq[i - 1]); c->str[i] = (char)toupper((unsigned char)seq[i - 1]); if (c->str[i] == 'T') c->str[i] = 'U'; c->cu[i] = 1; }
  for (int i = 1; i <= n; i++)
    for (int j = i + TURN + 1; j <= n; j++) c->pt[IX(c, i, j)] = (unsigned char)PAIR[c->S[i]][c->S[j]];
  if (cst) {
    if ((int)strlen(cst) != n) { snprintf(g_err, sizeof g_err, "constraint length %d != %d", (int)strlen(cst), n); return -1; }
    int *ptab = (int *)malloc(sizeof(int) * (n + 2));
    if (pair_table(n, cst, ptab)) { free(ptab); snprintf(g_err, sizeof g_err, "unbalanced brackets in constraint"); return -1; }
    for (int p = 1; p <= n; p++) {
      char ch = cst[p - 1];
      if (ch == 'x') { /* p may not pair */
        for (int q = 1; q <= n; q++) { c->pt[IX(c, MIN2(p, q), MAX2(p, q))] = 0; }
      } else if (ch == '(') {
        int i = p, j = ptab[p];
        unsigned char keep = c->pt[IX(c, i, j)]; /* enforced pairs must be canonical here (libRNA would use type 7) */
        for (int q = 1; q <= n; q++) { c->pt[IX(c, MIN2(i, q), MAX2(i, q))] = 0; c->pt[IX(c, MIN2(j, q), MAX2(j, q))] = 0; }
        for (int a = 1; a <= n; a++)
          for (int b = a + 1; b <= n; b++)
            if ((a < i && i < b && b < j) || (i < a && a < j && j < b)) c->pt[IX(c, a, b)] = 0;
        c->pt[IX(c, i, j)] = keep;
        c->cu[i] = c->cu[j] = 0; /* ENFORCE_BP: neither end may stay unpaired */
      } else if (ch != ')' && ch != '.') { free(ptab); snprintf(g_err, sizeof g_err, "unsupported constraint symbol '%c'", ch); return -1; }
    }
    free(ptab);
  }
  for (int i = n; i >= 1; i--) c->upr[i] = c->cu[i] ? c->upr[i + 1] + 1 : 0;
  if (mo && mo->seq) {
    int m = (int)strlen(mo->seq);
    c->m_len = m;
    c->m_c = orc_motif_contribution(M, mo);
    c->m_x = exp(-(double)c->m_c * 10. / M->p.kT);
    for (int i = 1; i + m - 1 <= n && c->n_occ < 64; i++)
      if (!strncmp(c->str + i, mo->seq, m)) c->occ[c->n_occ++] = i;
    int *ptab = (int *)malloc(sizeof(int) * (m + 2));
    if (pair_table(m, mo->fold, ptab)) { free(ptab); snprintf(g_err, sizeof g_err, "bad motif fold"); return -1; }
    for (int a = 2; a < m; a++) if (ptab[a] > a && c->n_mpairs < 64) { c->mp_i[c->n_mpairs] = a - 1; c->mp_j[c->n_mpairs++] = ptab[a] - 1; }
    free(ptab);
  }
  return 0;
}

static int is_motif(const fctx *c, int i, int j) {
  if (!c->n_occ || j - i + 1 != c->m_len) return 0;
  for (int k = 0; k < c->n_occ; k++) if (c->occ[k] == i) return 1;
  return 0;
}

/* ------------------------------------------------------------------ eval_structure ---------- */
static int eval_pair(const orc_model *M, const int *S, const char *str, const int *pt, int n, int i, int j) {
  const raw_par *P = &M->m;
  int type = PAIR[S[i]][S[j]]; if (!type) type = 7;
  int p = i + 1, nstems = 0, k1 = 0, l1 = 0;
  while (p < j) { if (pt[p] > p) { if (!nstems) { k1 = p; l1 = pt[p]; } nstems++; p = pt[p] + 1; } else p++; }
  if (nstems == 0) return E_hairpin(P, j - i - 1, type, S[i + 1], S[j - 1], str + i);
  if (nstems == 1) {
    int t2 = PAIR[S[k1]][S[l1]]; if (!t2) t2 = 7;
    return E_intloop(P, k1 - i - 1, j - l1 - 1, type, rtype[t2], S[i + 1], S[j - 1], S[k1 - 1], S[l1 + 1]) + eval_pair(M, S, str, pt, n, k1, l1);
  }
  int e = P->ML_closing + E_stem(P, rtype[type], S[j - 1], S[i + 1], 0), unp = 0;
  p = i + 1;
  while (p < j) {
    if (pt[p] > p) {
      int q = pt[p], t2 = PAIR[S[p]][S[q]]; if (!t2) t2 = 7;
      e += E_stem(P, t2, S[p - 1], S[q + 1], 0) + eval_pair(M, S, str, pt, n, p, q);
      p = q + 1;
    } else { unp++; p++; }
  }
  return e + unp * P->ML_base;
}

int orc_eval_structure(const orc_model *M, int n, const char *seq, const char *structure) {
  int *S = (int *)calloc(n + 2, sizeof(int)), *pt = (int *)calloc(n + 2, sizeof(int));
  char *str = (char *)calloc(n + 3, 1);
  for (int i = 1; i <= n; i++) { S[i] = enc(seq[i - 1]); str[i] = (char)toupper((unsigned char)seq[i - 1]); }
  int e = 0;
  if (pair_table(n, structure, pt)) e = INF;
  else {
    for (int p = 1; p <= n;) {
      if (pt[p] > p) {
        int q = pt[p], t = PAIR[S[p]][S[q]]; if (!t) t = 7;
        e += E_stem(&M->m, t, p > 1 ? S[p - 1] : -1, q < n ? S[q + 1] : -1, 1) + eval_pair(M, S, str, pt, n, p, q);
        p = q + 1;
      } else p++;
    }
  }
  free(S); free(pt); free(str);
  return e;
}

/* ------------------------------------------------------------------ MFE --------------------- */
static int hairpin_e(const orc_model *M, const fctx *c, int i, int j, int type) {
  int u = j - i - 1;
  if (c->upr[i + 1] < u) return INF;
  int e = E_hairpin(&M->m, u, type, c->S[i + 1], c->S[j - 1], c->str + i);
  if (is_motif(c, i, j)) e += c->m_c;
  return e;
}

typedef struct { int i, j, ml; } sect;

int orc_mfe(const orc_model *M, int n, const char *seq, const char *cst, const orc_motif *mo, int32_t *e_out, char *structure) {
  fctx cx, *c = &cx;
  if (ctx_init(c, M, n, seq, cst, mo)) { ctx_free(c); return -1; }
  const raw_par *P = &M->m;
  const int *S = c->S;
  size_t N2 = (size_t)(n + 2) * (n + 2);
  int *C = (int *)malloc(sizeof(int) * N2), *F = (int *)malloc(sizeof(int) * N2), *f5 = (int *)malloc(sizeof(int) * (n + 2));
  for (size_t k = 0; k < N2; k++) C[k] = F[k] = INF;
  for (int j = TURN + 2; j <= n; j++)
    for (int i = j - TURN - 1; i >= 1; i--) {
      int type = c->pt[IX(c, i, j)], best = INF;
      if (type) {
        best = hairpin_e(M, c, i, j, type);
        for (int k = i + 1; k <= MIN2(j - 2 - TURN, i + MAXLOOP + 1); k++) {
          int u1 = k - i - 1;
          if (c->upr[i + 1] < u1) break;
          int minl = MAX2(k + TURN + 1, j - 1 - MAXLOOP + u1);
          for (int l = j - 1; l >= minl; l--) {
            int u2 = j - l - 1;
            if (c->upr[l + 1] < u2) break;
            int t2 = c->pt[IX(c, k, l)];
            if (!t2 || C[IX(c, k, l)] >= INF) continue;
            int e = E_intloop(P, u1, u2, type, rtype[t2], S[i + 1], S[j - 1], S[k - 1], S[l + 1]) + C[IX(c, k, l)];
            if (e < best) best = e;
          }
        }
        int dec = INF;
        for (int k = i + 2; k <= j - 1; k++) {
          int a = F[IX(c, i + 1, k - 1)], b = F[IX(c, k, j - 1)];
          if (a < INF && b < INF && a + b < dec) dec = a + b;
        }
        if (dec < INF) { int e = dec + P->ML_closing + E_stem(P, rtype[type], S[j - 1], S[i + 1], 0); if (e < best) best = e; }
      }
      C[IX(c, i, j)] = best;
      int fm = INF;
      if (c->cu[i] && F[IX(c, i + 1, j)] < INF) fm = F[IX(c, i + 1, j)] + P->ML_base;
      if (c->cu[j] && F[IX(c, i, j - 1)] < INF) fm = MIN2(fm, F[IX(c, i, j - 1)] + P->ML_base);
      if (best < INF) fm = MIN2(fm, best + E_stem(P, type, i > 1 ? S[i - 1] : -1, j < n ? S[j + 1] : -1, 0));
      for (int k = i + 1; k <= j; k++) {
        int a = F[IX(c, i, k - 1)], b = F[IX(c, k, j)];
        if (a < INF && b < INF && a + b < fm) fm = a + b;
      }
      F[IX(c, i, j)] = fm;
    }
  f5[0] = 0;
  for (int j = 1; j <= n; j++) {
    int b = (c->cu[j] && f5[j - 1] < INF) ? f5[j - 1] : INF;
    for (int i = j - TURN - 1; i >= 1; i--) {
      int type = c->pt[IX(c, i, j)];
      if (!type || C[IX(c, i, j)] >= INF || f5[i - 1] >= INF) continue;
      int e = f5[i - 1] + C[IX(c, i, j)] + E_stem(P, type, i > 1 ? S[i - 1] : -1, j < n ? S[j + 1] : -1, 1);
      if (e < b) b = e;
    }
    f5[j] = b;
  }
  *e_out = f5[n];
  if (structure) {
    memset(structure, '.', n); structure[n] = 0;
    if (f5[n] < INF) {
      /* Traceback.  Tie-break order (the engine reproduces it exactly):
       *  f5(j):  "j unpaired" first; else i from j-TURN-1 DOWN to 1, first match.
       *  pair:   hairpin; interior k ascending / l descending; multiloop split k ascending.
       *  fML:    "j unpaired"; "i unpaired"; stem (i,j); split k ascending. */
      sect *stk = (sect *)malloc(sizeof(sect) * (4 * n + 8));
      int sp = 0;
      stk[sp++] = (sect){1, n, 0};
      while (sp) {
        sect s = stk[--sp];
        int i = s.i, j = s.j;
        if (s.ml == 0) { /* exterior prefix f5[j] */
          while (j > 0) {
            if (c->cu[j] && f5[j - 1] < INF && f5[j] == f5[j - 1]) { j--; continue; }
            int found = 0;
            for (int k = j - TURN - 1; k >= 1; k--) {
              int type = c->pt[IX(c, k, j)];
              if (!type || C[IX(c, k, j)] >= INF || f5[k - 1] >= INF) continue;
              if (f5[j] == f5[k - 1] + C[IX(c, k, j)] + E_stem(P, type, k > 1 ? S[k - 1] : -1, j < n ? S[j + 1] : -1, 1)) {
                stk[sp++] = (sect){1, k - 1, 0};
                stk[sp++] = (sect){k, j, 2};
                found = 1; break;
              }
            }
            if (!found) { snprintf(g_err, sizeof g_err, "f5 traceback failed at %d", j); }
            break;
          }
        } else if (s.ml == 2) { /* pair (i,j) */
          structure[i - 1] = '('; structure[j - 1] = ')';
          int type = c->pt[IX(c, i, j)], e = C[IX(c, i, j)], done = 0;
          if (e == hairpin_e(M, c, i, j, type)) {
            if (is_motif(c, i, j)) for (int q = 0; q < c->n_mpairs; q++) { structure[i - 1 + c->mp_i[q]] = '('; structure[i - 1 + c->mp_j[q]] = ')'; }
            continue;
          }
          for (int k = i + 1; k <= MIN2(j - 2 - TURN, i + MAXLOOP + 1) && !done; k++) {
            int u1 = k - i - 1;
            if (c->upr[i + 1] < u1) break;
            int minl = MAX2(k + TURN + 1, j - 1 - MAXLOOP + u1);
            for (int l = j - 1; l >= minl; l--) {
              int u2 = j - l - 1;
              if (c->upr[l + 1] < u2) break;
              int t2 = c->pt[IX(c, k, l)];
              if (!t2 || C[IX(c, k, l)] >= INF) continue;
              if (e == E_intloop(P, u1, u2, type, rtype[t2], S[i + 1], S[j - 1], S[k - 1], S[l + 1]) + C[IX(c, k, l)]) { stk[sp++] = (sect){k, l, 2}; done = 1; break; }
            }
          }
          if (done) continue;
          int cl = P->ML_closing + E_stem(P, rtype[type], S[j - 1], S[i + 1], 0);
          for (int k = i + 2; k <= j - 1; k++) {
            int a = F[IX(c, i + 1, k - 1)], b = F[IX(c, k, j - 1)];
            if (a < INF && b < INF && e == a + b + cl) { stk[sp++] = (sect){i + 1, k - 1, 1}; stk[sp++] = (sect){k, j - 1, 1}; done = 1; break; }
          }
          if (!done) snprintf(g_err, sizeof g_err, "pair traceback failed at %d,%d", i, j);
        } else { /* fML[i,j] */
          int e = F[IX(c, i, j)];
          if (c->cu[j] && F[IX(c, i, j - 1)] < INF && e == F[IX(c, i, j - 1)] + P->ML_base) { stk[sp++] = (sect){i, j - 1, 1}; continue; }
          if (c->cu[i] && F[IX(c, i + 1, j)] < INF && e == F[IX(c, i + 1, j)] + P->ML_base) { stk[sp++] = (sect){i + 1, j, 1}; continue; }
          int type = c->pt[IX(c, i, j)];
          if (type && C[IX(c, i, j)] < INF && e == C[IX(c, i, j)] + E_stem(P, type, i > 1 ? S[i - 1] : -1, j < n ? S[j + 1] : -1, 0)) { stk[sp++] = (sect){i, j, 2}; continue; }
          int done = 0;
          for (int k = i + 1; k <= j; k++) {
            int a = F[IX(c, i, k - 1)], b = F[IX(c, k, j)];
            if (a < INF && b < INF && e == a + b) { stk[sp++] = (sect){i, k - 1, 1}; stk[sp++] = (sect){k, j, 1}; done = 1; break; }
          }
          if (!done) snprintf(g_err, sizeof g_err, "fML traceback failed at %d,%d", i, j);
        }
      }
      free(stk);
    }
  }
  free(C); free(F); free(f5); ctx_free(c);
  return 0;
}

/* ------------------------------------------------------------------ partition function ------ */
static double hairpin_x(const orc_model *M, const fctx *c, const double *scale, int i, int j, int type) {
  int u = j - i - 1;
  if (c->upr[i + 1] < u) return 0.;
  double q = X_hairpin(M, u, type, c->S[i + 1], c->S[j - 1], c->str + i) * scale[u + 2];
  if (is_motif(c, i, j)) q *= c->m_x; /* bound motif REPLACES the empty loop */
  return q;
}

static size_t tri(int n, int i, int j) { return (size_t)i * n - (size_t)i * (i - 1) / 2 + (j - i); } /* 0-based */

int orc_pf(const orc_model *M, int n, const char *seq, const char *cst, const orc_motif *mo, double *dG, double *bppm) {
  fctx cx, *c = &cx;
  if (ctx_init(c, M, n, seq, cst, mo)) { ctx_free(c); return -1; }
  const pf_par *p = &M->p;
  const int *S = c->S;
  size_t N2 = (size_t)(n + 2) * (n + 2);
  double *q = (double *)calloc(N2, sizeof(double)), *qb = (double *)calloc(N2, sizeof(double)), *qm = (double *)calloc(N2, sizeof(double));
  double *scale = (double *)malloc(sizeof(double) * (n + 3)), *eMLb = (double *)malloc(sizeof(double) * (n + 3));
  double *rows = (double *)calloc(4 * (size_t)(n + 2), sizeof(double));
  double *qq = rows, *qq1 = rows + (n + 2), *qqm = rows + 2 * (n + 2), *qqm1 = rows + 3 * (n + 2);
  scale[0] = 1.; eMLb[0] = 1.;
  for (int k = 1; k <= n + 2; k++) { scale[k] = scale[k - 1] / p->pf_scale; eMLb[k] = pow(p->MLbase, (double)k) * scale[k]; }
  for (int d = 0; d <= TURN; d++)
    for (int i = 1; i + d <= n; i++) q[IX(c, i, i + d)] = (c->upr[i] > d) ? scale[d + 1] : 0.;
  /* q[i,i-1] = 1 convention for empty segments */
  for (int i = 1; i <= n + 1; i++) q[IX(c, i, i - 1)] = 1.;
  for (int j = TURN + 2; j <= n; j++) {
    for (int i = j - TURN - 1; i >= 1; i--) {
      int type = c->pt[IX(c, i, j)];
      double qbt = 0.;
      if (type) {
        qbt += hairpin_x(M, c, scale, i, j, type);
        for (int k = i + 1; k <= MIN2(j - 2 - TURN, i + MAXLOOP + 1); k++) {
          int u1 = k - i - 1;
          if (c->upr[i + 1] < u1) break;
          int minl = MAX2(k + TURN + 1, j - 1 - MAXLOOP + u1);
          for (int l = j - 1; l >= minl; l--) {
            int u2 = j - l - 1;
            if (c->upr[l + 1] < u2) break;
            int t2 = c->pt[IX(c, k, l)];
            if (!t2) continue;
            qbt += qb[IX(c, k, l)] * scale[u1 + u2 + 2] * X_intloop(p, u1, u2, type, rtype[t2], S[i + 1], S[j - 1], S[k - 1], S[l + 1]);
          }
        }
        double t = 0.;
        for (int k = i + 2; k <= j - 1; k++) t += qm[IX(c, i + 1, k - 1)] * qqm1[k];
        qbt += t * p->MLclosing * X_stem(p, rtype[type], S[j - 1], S[i + 1], 0) * scale[2];
      }
      qb[IX(c, i, j)] = qbt;
      /* qqm[i] = Qm1[i,j]: exactly one stem starting at i, unpaired tail up to j */
      double t = c->cu[j] ? qqm1[i] * eMLb[1] : 0.;
      if (type) t += qbt * X_stem(p, type, i > 1 ? S[i - 1] : -1, j < n ? S[j + 1] : -1, 0);
      qqm[i] = t;
      t = 0.;
      for (int k = j; k > i; k--) t += qm[IX(c, i, k - 1)] * qqm[k];
      int maxk = MIN2(i + c->upr[i], j);
      for (int k = maxk; k > i; k--) t += eMLb[k - i] * qqm[k];
      qm[IX(c, i, j)] = t + qqm[i];
      /* exterior */
      t = c->cu[j] ? qq1[i] * scale[1] : 0.;
      if (type) t += qbt * X_stem(p, type, i > 1 ? S[i - 1] : -1, j < n ? S[j + 1] : -1, 1);
      qq[i] = t;
      if (c->upr[i] >= j - i + 1) t += scale[j - i + 1];
      for (int k = i; k < j; k++) t += q[IX(c, i, k)] * qq[k + 1];
      q[IX(c, i, j)] = t;
    }
    double *tmp = qq1; qq1 = qq; qq = tmp;
    tmp = qqm1; qqm1 = qqm; qqm = tmp;
    /* entries of the new qq/qqm rows for i > j-TURN-1 must read as "no stem" */
    for (int i = MAX2(1, j - TURN); i <= n; i++) { qq1[i] = 0.; qqm1[i] = 0.; }
  }
  double Z = q[IX(c, 1, n)];
  *dG = (Z > 0.) ? (-log(Z) - n * log(p->pf_scale)) * p->kT / 1000. : INFINITY;

  if (bppm) {
    for (size_t k = 0; k < (size_t)n * (n + 1) / 2; k++) bppm[k] = 0.;
    if (Z > 0.) {
      double *pr = (double *)calloc(N2, sizeof(double));
      double *q1k = (double *)calloc(n + 3, sizeof(double)), *qln = (double *)calloc(n + 3, sizeof(double));
      double *prml = (double *)calloc(n + 3, sizeof(double)), *prm_l = (double *)calloc(n + 3, sizeof(double)), *prm_l1 = (double *)calloc(n + 3, sizeof(double));
      for (int k = 1; k <= n; k++) { q1k[k] = q[IX(c, 1, k)]; qln[k] = q[IX(c, k, n)]; }
      q1k[0] = 1.; qln[n + 1] = 1.;
      for (int i = 1; i <= n; i++)
        for (int j = i + TURN + 1; j <= n; j++) {
          int type = c->pt[IX(c, i, j)];
          if (type && qb[IX(c, i, j)] > 0.)
            pr[IX(c, i, j)] = q1k[i - 1] * qln[j + 1] / q1k[n] * X_stem(p, type, i > 1 ? S[i - 1] : -1, j < n ? S[j + 1] : -1, 1);
        }
      for (int l = n; l > TURN + 1; l--) {
        /* (k,l) enclosed by (i,j) in an interior loop */
        for (int k = 1; k < l - TURN; k++) {
          int t2 = c->pt[IX(c, k, l)];
          if (!t2 || qb[IX(c, k, l)] == 0.) continue;
          t2 = rtype[t2];
          double acc = 0.;
          for (int i = MAX2(1, k - MAXLOOP - 1); i <= k - 1; i++) {
            int u1 = k - i - 1;
            if (c->upr[i + 1] < u1) continue;
            for (int j = l + 1; j <= MIN2(l + MAXLOOP - u1 + 1, n); j++) {
              int u2 = j - l - 1;
              if (c->upr[l + 1] < u2) break;
              int type = c->pt[IX(c, i, j)];
              if (type && pr[IX(c, i, j)] > 0.)
                acc += pr[IX(c, i, j)] * scale[u1 + u2 + 2] * X_intloop(p, u1, u2, type, t2, S[i + 1], S[j - 1], S[k - 1], S[l + 1]);
            }
          }
          pr[IX(c, k, l)] += acc;
        }
        /* (k,l) as a stem of a multiloop closed by (i,j) */
        double prm_MLb = 0.;
        if (l < n)
          for (int k = 2; k < l - TURN; k++) {
            int i = k - 1;
            double prmt = 0., prmt1 = 0.;
            int tt = c->pt[IX(c, i, l + 1)];
            if (tt) prmt1 = pr[IX(c, i, l + 1)] * p->MLclosing * X_stem(p, rtype[tt], S[l], S[i + 1], 0);
            for (int j = l + 2; j <= n; j++) {
              tt = c->pt[IX(c, i, j)];
              if (tt) prmt += pr[IX(c, i, j)] * X_stem(p, rtype[tt], S[j - 1], S[i + 1], 0) * qm[IX(c, l + 1, j - 1)];
            }
            prmt *= p->MLclosing;
            prml[i] = prmt;
            prm_l[i] = (c->cu[l + 1] ? prm_l1[i] * eMLb[1] : 0.) + prmt1;
            prm_MLb = (c->cu[i] ? prm_MLb * eMLb[1] : 0.) + prml[i];
            prml[i] = prml[i] + prm_l[i];
            int tkl = c->pt[IX(c, k, l)];
            if (!tkl || qb[IX(c, k, l)] == 0.) continue;
            double t = prm_MLb;
            for (int i2 = 1; i2 <= k - 2; i2++) t += prml[i2] * qm[IX(c, i2 + 1, k - 1)];
            t *= X_stem(p, tkl, S[k - 1], l < n ? S[l + 1] : -1, 0) * scale[2];
            pr[IX(c, k, l)] += t;
          }
        double *tmp = prm_l1; prm_l1 = prm_l; prm_l = tmp;
      }
      for (int i = 1; i <= n; i++)
        for (int j = i + TURN + 1; j <= n; j++) pr[IX(c, i, j)] *= qb[IX(c, i, j)];
      /* ligand motif: inner pairs inherit P(motif hairpin closed by (i,j)) -- "bp_correction" */
      for (int o = 0; o < c->n_occ; o++) {
        int i = c->occ[o], j = i + c->m_len - 1, type = c->pt[IX(c, i, j)];
        if (!type || qb[IX(c, i, j)] == 0.) continue;
        double pm = pr[IX(c, i, j)] / qb[IX(c, i, j)] * hairpin_x(M, c, scale, i, j, type);
        for (int m = 0; m < c->n_mpairs; m++) pr[IX(c, i + c->mp_i[m], i + c->mp_j[m])] += pm;
      }
      for (int i = 1; i <= n; i++)
        for (int j = i + 1; j <= n; j++) bppm[tri(n, i - 1, j - 1)] = pr[IX(c, i, j)];
      free(pr); free(q1k); free(qln); free(prml); free(prm_l); free(prm_l1);
    }
  }
  free(q); free(qb); free(qm); free(scale); free(eMLb); free(rows); ctx_free(c);
  return 0;
}

double orc_macrostate_prob(const orc_model *M, int n, const char *seq, const char *cst, const orc_motif *mo, int float_dg) {
  double g_tot, g_act;
  if (orc_pf(M, n, seq, NULL, mo, &g_tot, NULL) || orc_pf(M, n, seq, cst, mo, &g_act, NULL)) return NAN;
  if (float_dg) { g_tot = (double)(float)g_tot; g_act = (double)(float)g_act; }
  return exp((g_tot - g_act) / (M->p.kT / 1000.));
}

void orc_count_terms(int n, const char *seq, double *w_in, double *w_out) {
  int *S = (int *)calloc(n + 2, sizeof(int));
  for (int i = 1; i <= n; i++) S[i] = enc(seq[i - 1]);
  double win = 0, wout = 0, nil = 0;
#define PT(i, j) ((j) - (i) > TURN ? PAIR[S[i]][S[j]] : 0)
  for (int i = 1; i <= n; i++)
    for (int j = i + TURN + 1; j <= n; j++) {
      win += 3. * (j - i);
      if (!PT(i, j)) continue;
      win += 1. + (j - i - 2);
      for (int k = i + 1; k <= MIN2(j - 2 - TURN, i + MAXLOOP + 1); k++) {
        int u1 = k - i - 1, minl = MAX2(k + TURN + 1, j - 1 - MAXLOOP + u1);
        for (int l = j - 1; l >= minl; l--) if (PT(k, l)) nil += 1.;
      }
      if (i >= 2) wout += i - 2;
    }
  win += nil; wout += nil;
  for (int l = 5; l <= n - 1; l++) for (int k = 2; k <= l - 4; k++) wout += n - l - 1;
#undef PT
  *w_in = win; *w_out = wout;
  free(S);
}
