/* oracle/fold_oracle.h -- CPU ORACLE.  TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * A plain-C restatement of the RNA folding calls that kalekundert/addapt's hot path makes into
 * ViennaRNA libRNA (third-party, un-vendored, un-pinned 2.2.x; absent from this image):
 *   reference src/scoring.cc:53-71   macrostate_prob  = vrna_pf / vrna_constraints_add / vrna_pf
 *   reference src/scoring.cc:37-51   base_pair_prob   = vrna_pf with compute_bpp
 *   reference src/scoring.cc:73-103  make_fold_compound (default model details, vrna_sc_add_hi_motif)
 * plus Zuker MFE (asked for by BASELINE.json; the reference never calls vrna_mfe).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * call into this file.  The product (addapt_b200/) never links or loads it.
 *
 * PARITY STATUS: "parity unpinned" for exact values -- the reference holds no golden numbers for
 * this path and libRNA cannot be run here.  What IS pinned: every threshold of the reference's
 * tests/test_scoring.cc:83-259 and the energies quoted in its comments (tests/golden/).
 */
#ifndef FOLD_ORACLE_H
#define FOLD_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct orc_model orc_model;

/* Load a ViennaRNA "parameter file v2.0"; returns NULL on failure (message in orc_error()). */
orc_model *orc_model_load(const char *par_path);
void orc_model_free(orc_model *);
const char *orc_error(void);
double orc_kT_kcal(const orc_model *);            /* 0.61632... kcal/mol at 37 C */

/* A ligand hairpin motif, reference src/scoring.cc:92-100 (vrna_sc_add_hi_motif). */
typedef struct {
  const char *seq;      /* e.g. GAUACCAGCCGAAAGGCCCUUGGCAGC           */
  const char *fold;     /* e.g. (...((.(((....)))....))...)           */
  double energy_kcal;   /* kT*ln(Kd_uM/1e6), reference scoring.cc:98   */
} orc_motif;

/* Energy (dcal/mol) of `structure` on `seq` under the dangles=2 Turner model. */
int orc_eval_structure(const orc_model *, int n, const char *seq, const char *structure);

/* motif contribution c (dcal/mol) as vrna_sc_add_hi_motif derives it */
int orc_motif_contribution(const orc_model *, const orc_motif *);

/* Zuker MFE + traceback.  cst / motif may be NULL.  structure: n+1 bytes.  Returns 0, or -1 if the
 * constraint is malformed.  *e_dcal == ORC_INF when no structure satisfies the constraint. */
#define ORC_INF 10000000
int orc_mfe(const orc_model *, int n, const char *seq, const char *cst, const orc_motif *motif,
            int32_t *e_dcal, char *structure);

/* McCaskill partition function.  dG_kcal: ensemble free energy in kcal/mol (double; the caller
 * rounds through float when it wants vrna_pf's return type).  bppm (may be NULL): upper triangle,
 * row-major, 0-based, INCLUDING the diagonal: tri(i,j) = i*n - i*(i-1)/2 + (j-i), n(n+1)/2 doubles.
 * dG = +inf when the constrained ensemble is empty. */
int orc_pf(const orc_model *, int n, const char *seq, const char *cst, const orc_motif *motif,
           double *dG_kcal, double *bppm);

/* reference ViennaRnaFold::macrostate_prob (src/scoring.cc:53-71): unconstrained pf, then the same
 * compound with the hard constraint added.  float_dg != 0 rounds both energies through float as
 * vrna_pf's return type does. */
double orc_macrostate_prob(const orc_model *, int n, const char *seq, const char *cst,
                           const orc_motif *motif, int float_dg);

/* recurrence-term counts of SURVEY 8(d): W_in, W_out for an (unconstrained) sequence */
void orc_count_terms(int n, const char *seq, double *w_in, double *w_out);

/* ------------------------------------------------------------------------------------------
 * Design problem / score / Monte Carlo (reference src/scoring.cc:114-158,233-259 and
 * src/sampling.cc:22-107,170-303,305-338).  Mirrors adp_problem of include/addapt_b200.h. */
typedef struct {                 /* mirrors adp_bp_term: a term over RnaFold::base_pair_prob (reference include/scoring.hh:48) */
  int kind;                      /* 0: x = P(i pairs j); 1: x = mean over pos[] of P(position is paired) */
  int cond, favorable;           /* value = log(x) or log(1-x) in the apo (0) / holo (1) unconstrained ensemble */
  double weight;
  int i, j;
  int n_pos;
  const int *pos;
} orc_bp_term;

typedef struct {
  int n;
  const char *init_seq;          /* mixed case; upper = mutable (sampling.cc:170-175) */
  int n_macro;
  const char *const *macro_cst;  /* each n chars of . x ( ) */
  int n_terms;
  const int *term_macro;         /* index into macro_cst */
  const int *term_cond;          /* 0 apo, 1 holo */
  const int *term_favorable;     /* 1: log(p), 0: log(1-p) */
  const double *term_weight;
  orc_motif motif;               /* motif.seq == NULL: no aptamer (holo == apo) */
  int n_ctx;                     /* 0: bare device */
  const char *const *ctx_before;
  const char *const *ctx_after;
  int n_bp_terms;                /* evaluated after the macrostate terms */
  const orc_bp_term *bp_terms;
} orc_problem;

/* score of one upper-case sequence (n chars); terms: (n_terms + n_bp_terms) * max(n_ctx,1) values or NULL */
double orc_score(const orc_model *, const orc_problem *, const char *seq, int float_dg, double *terms);

/* Philox4x32-10, key = (seed_lo, seed_hi), counter = (ctr0, ctr1, ctr2, ctr3) */
void orc_philox4x32(uint32_t key0, uint32_t key1, const uint32_t ctr[4], uint32_t out[4]);

/* One trajectory of the Metropolis maximiser with the engine's RNG convention:
 * thermostat_kind 0 fixed {T}, 1 annealing {T_hi, T_lo, cycle_len}.  Outputs: final/best sequence
 * (n bytes each), scores, best step, outcome counters [REJECT, WORSENED, UNCHANGED, IMPROVED]. */
int orc_mc_run(const orc_model *, const orc_problem *, uint64_t seed, uint64_t traj_id, int n_steps,
               int thermostat_kind, const double *thermostat_args, int float_dg,
               char *final_seq, double *final_score, char *best_seq, double *best_score,
               int32_t *best_step, uint32_t *outcomes, double *score_trace /* n_steps or NULL */);

/* orc_mc_run plus per-step records in the columns of TsvTrajectoryReporter (reference src/sampling.cc:430-489): every
 * interval-th step writes rec_vals[8 + nterms] = {step, current_score, proposed_score, score_diff, temperature,
 * metropolis_criterion, random_threshold, outcome, term values of the last evaluated proposal} and rec_seqs[2n] = current,
 * proposed sequence; proposal columns keep the last scored step's values across ACCEPT_UNCHANGED steps (0 before the first). */
int orc_mc_run_trace(const orc_model *, const orc_problem *, uint64_t seed, uint64_t traj_id, int n_steps,
                     int thermostat_kind, const double *thermostat_args, int float_dg,
                     char *final_seq, double *final_score, char *best_seq, double *best_score,
                     int32_t *best_step, uint32_t *outcomes, int interval, double *rec_vals, char *rec_seqs);

/* ≙ mutate_recursively(device, pos, nucleotide) (reference src/sampling.cc:205-282) on seq (n chars, mixed case, modified in
 * place); returns 0, or -1 mismatched brackets, -2 partner not mutable, -3 contradictory pairing */
int orc_mutate(const orc_problem *, char *seq, int pos, char nucleotide);

#ifdef __cplusplus
}
#endif
#endif
