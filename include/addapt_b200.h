/* addapt_b200.h -- C ABI of the B200 fold/score/Monte-Carlo engine (libaddapt_b200.so).
 *
 * This is the drop-in boundary for kalekundert/addapt's hot path.  In the reference the path sits
 * behind the abstract class RnaFold (include/scoring.hh:42-55) whose only implementation,
 * ViennaRnaFold (src/scoring.cc:17-103), calls ViennaRNA's libRNA on one CPU core; above it sit
 * ScoreFunction::evaluate (src/scoring.cc:114-158) and MonteCarlo::apply (src/sampling.cc:22-107).
 * Every entry point below names the reference interface it replaces.  INTEGRATION.md shows the
 * C++ binding a maintainer of the reference would add (addapt_b200/host/ is that binding).
 *
 * Conventions
 *  - plain pointers and sizes; no C++ / torch types; no exceptions cross the boundary.
 *  - every function returns 0 on success or a negative adp_status; adp_last_error() has the text.
 *  - data pointers may be HOST or DEVICE memory (detected with cudaPointerGetAttributes); host
 *    buffers are staged through the engine's pinned buffers, device buffers are used in place.
 *  - an engine belongs to one GPU and one host thread at a time; calls are synchronous (the
 *    engine's stream is synchronised before returning) unless adp_engine_set_async(e, 1).
 *  - sequences are rows of n bytes, ASCII, either case, T == U, no terminator.
 *  - there is NO CPU fallback: without a CUDA device adp_engine_create fails.
 */
#ifndef ADDAPT_B200_H
#define ADDAPT_B200_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct adp_engine adp_engine;

typedef enum {
  ADP_OK = 0,
  ADP_ERR_ARG = -1,      /* malformed argument (bad constraint string, length mismatch, n too large ...) */
  ADP_ERR_CUDA = -2,     /* CUDA runtime error; text in adp_last_error */
  ADP_ERR_PARAM = -3,    /* energy parameter file missing or malformed */
  ADP_ERR_NOMEM = -4
} adp_status;

#define ADP_INF_DCAL 10000000   /* MFE energy reported when no structure satisfies the constraint */
#define ADP_MAX_LEN 320         /* longest construct (context included) one call accepts */

/* ---- engine ------------------------------------------------------------------------------------
 * Replaces vrna_md_set_default + the per-call table set-up inside vrna_fold_compound
 * (reference src/scoring.cc:80-88): the Turner tables are parsed and scaled ONCE, then live in HBM.
 * par_file: a ViennaRNA "parameter file v2.0"; NULL = $ADDAPT_B200_PAR or the file shipped in params/. */
int adp_engine_create(int device, const char *par_file, adp_engine **out);
void adp_engine_destroy(adp_engine *);
const char *adp_last_error(const adp_engine *);          /* engine may be NULL: last create() error */
int adp_engine_set_stream(adp_engine *, void *cuda_stream); /* cudaStream_t to launch on (default: own stream) */
int adp_engine_set_async(adp_engine *, int async);       /* 1: do not synchronise before returning (device buffers only) */
int adp_engine_sync(adp_engine *);
/* 1: fold everything with the general kernels (any letters, any length <= ADP_MAX_LEN) instead of the fast
 * ACGU-only kernels; results agree to rounding.  For parity tests of the general path. */
int adp_engine_set_general_only(adp_engine *, int on);
double adp_kT_kcal(const adp_engine *);                  /* fc->exp_params->kT/1000, reference scoring.cc:69,93 */
/* Provenance of the energy tables: "file" (a parameter file used as given) or "reconstructed" (the file declares itself a
 * reconstruction, like the one shipped in params/: int11 / int21 / int22 rule-generated -- results then differ from
 * ViennaRNA's; adp_engine_create prints a warning on stderr unless $ADDAPT_B200_QUIET is set).  adp_param_file: its path. */
const char *adp_param_provenance(const adp_engine *);
const char *adp_param_file(const adp_engine *);
int adp_device_count(void);
uint64_t adp_launch_count(const adp_engine *);           /* kernels launched by this engine so far */

/* ---- ligand motif  ≙ Aptamer (reference include/model.hh:113-137) handed to vrna_sc_add_hi_motif
 * (src/scoring.cc:92-100).  The bonus is kT*ln(kd_uM/1e6). */
typedef struct {
  const char *seq;   /* e.g. GAUACCAGCCGAAAGGCCCUUGGCAGC */
  const char *fold;  /* e.g. (...((.(((....)))....))...) */
  double kd_uM;      /* dissociation constant, micromolar */
} adp_motif;

/* Energy (dcal/mol) of a given structure -- ≙ vrna_eval_structure as used inside vrna_sc_add_hi_motif. */
int adp_eval_structure(adp_engine *, const char *seq, const char *structure, int32_t *e_dcal);
/* The soft-constraint contribution (dcal/mol) the motif adds to its closing hairpin. */
int adp_motif_contribution(adp_engine *, const adp_motif *, int32_t *c_dcal);

/* ---- batched folds: B sequences of n nt each --------------------------------------------------
 * cst (may be NULL): ONE dot-bracket hard constraint of n chars shared by the batch, symbols . x ( ),
 * applied like vrna_constraints_add(fc, cst, VRNA_CONSTRAINT_DB_DEFAULT | VRNA_CONSTRAINT_DB_ENFORCE_BP)
 * (reference src/scoring.cc:61-62).  motif (may be NULL): holo condition. */

/* Zuker MFE + traceback  ≙ vrna_mfe (no call site in the reference; BASELINE.json configs[1]).
 * e_dcal: B int32.  structs: B*n chars of . ( ) or NULL. */
int adp_mfe_batch(adp_engine *, int B, int n, const uint8_t *seqs, const char *cst, const adp_motif *motif,
                  int32_t *e_dcal, char *structs);

/* McCaskill partition function  ≙ vrna_pf (reference src/scoring.cc:43,58,65).
 * dG_kcal: B doubles, ensemble free energy (kcal/mol), +inf when the constrained ensemble is empty.
 * bppm: NULL (inside pass only, compute_bpp=0) or B * n(n+1)/2 doubles -- upper triangle incl. diagonal,
 * row-major, 0-based: tri(i,j) = i*n - i*(i-1)/2 + (j-i)  (the matrix base_pair_prob reads, scoring.cc:37-51). */
int adp_pf_batch(adp_engine *, int B, int n, const uint8_t *seqs, const char *cst, const adp_motif *motif,
                 double *dG_kcal, double *bppm);

/* One "fold-eval" per sequence = apo AND holo, each inside + outside (pf + full bppm on the device),
 * returning the two ensemble energies and the pair probabilities of n_query (i,j) pairs (0-based)
 * -- the batched form of constructing ViennaRnaFold(device) + ViennaRnaFold(device, aptamer) and calling
 * base_pair_prob(i,j) on both (reference src/scoring.cc:37-51,145-146; tests/test_scoring.cc:231-252).
 * dG: B*2 doubles (apo, holo).  probs: B*2*n_query doubles (apo block then holo block per sequence).
 * bppm_full: NULL or B*2*n(n+1)/2 doubles. */
int adp_fold_eval_batch(adp_engine *, int B, int n, const uint8_t *seqs, const adp_motif *motif,
                        int n_query, const int32_t *query_i, const int32_t *query_j,
                        double *dG, double *probs, double *bppm_full);

/* ---- design problem  ≙ Device + ScoreFunction (reference include/model.hh:22-111, include/scoring.hh:97-192) */

/* A score term computed from base-pair probabilities, i.e. from what RnaFold::base_pair_prob returns (reference
 * include/scoring.hh:48; the fake backend of tests/test_scoring.cc:12-43 exists to test such terms).  The mounted snapshot ships only
 * MacrostateProbTerm; BASELINE.json's north_star names the two historic kinds, which the engine evaluates on the device from the
 * outside pass of the UNCONSTRAINED ensemble of the term's condition:
 *   kind 0, pair:     x = P(i pairs with j)                                  ("conditional pairing": want it in one condition only)
 *   kind 1, density:  x = mean over the position set of P(position is paired) ("bppm density" of a region)
 *   term value = log(x) if favorable else log(1 - x); the score adds weight * value, summed over contexts like every term.
 * Positions are 0-based in the RAW device (contexts shift them). */
typedef struct {
  int kind;             /* 0 = pair, 1 = density */
  int cond;             /* 0 = APO, 1 = HOLO */
  int favorable;        /* 1 = log(x), 0 = log(1 - x) */
  double weight;
  int i, j;             /* kind 0 */
  int n_pos;            /* kind 1: 1..64 positions (64 in total over all density terms) */
  const int *pos;
} adp_bp_term;

typedef struct {
  int n;                          /* raw device length */
  const char *init_seq;           /* mixed case; upper case = mutable (src/sampling.cc:170-175) */
  int n_macro;
  const char *const *macro_cst;   /* n_macro constraint strings of n chars (Device::macrostate) */
  int n_terms;                    /* MacrostateProbTerm list (src/scoring.cc:204-259) */
  const int *term_macro;          /* index into macro_cst */
  const int *term_cond;           /* 0 = APO, 1 = HOLO (ConditionEnum) */
  const int *term_favorable;      /* 1 = log(p), 0 = log(1-p) (FavorableEnum) */
  const double *term_weight;
  adp_motif motif;                /* motif.seq == NULL: no aptamer */
  int n_ctx;                      /* 0 = bare device; else contexts, score SUMMED over them (scoring.cc:123-132) */
  const char *const *ctx_before;
  const char *const *ctx_after;
  int float_dg;                   /* 1: round each ensemble energy through float like vrna_pf's return type */
  int n_bp_terms;                 /* base-pair-probability terms (0..8; n_terms + n_bp_terms <= 16), evaluated after the macrostate terms */
  const adp_bp_term *bp_terms;
} adp_problem;

/* ≙ ScoreFunction::evaluate for B candidate sequences of the same device (reference src/scoring.cc:114-158).
 * seqs: B*n bytes.  score: B doubles.  terms: NULL or B * max(n_ctx,1) * (n_terms + n_bp_terms) doubles (unweighted;
 * per context the macrostate terms, then the base-pair-probability terms). */
int adp_score_batch(adp_engine *, const adp_problem *, int B, const uint8_t *seqs, double *score, double *terms);

/* ≙ MonteCarlo::apply (reference src/sampling.cc:22-107) for n_traj independent trajectories, with
 * UnbiasedMutationMove (:287-303) and mutate_recursively (:205-282) proposals, Metropolis acceptance
 * (:76-89) and a Fixed (kind 0: {T}), Annealing (kind 1: {T_high, T_low, cycle_len}, :331-338) or AutoScaling
 * (kind 2: {target acceptance rate in (0,1), training period <= 512, initial T}, :371-401) thermostat, all on the device.  Trajectory t draws from Philox4x32-10 keyed by `seed` with counter (step, 0, id_lo, id_hi),
 * id = traj_id0 + t, so results do not depend on how trajectories are sharded over GPUs -- nor on how the engine splits a
 * long run into kernel launches (1,024 steps each; the state is carried in device memory).
 * Outputs (each may be NULL; host or device): final_seq/best_seq n_traj*n bytes (upper case ACGU, fixed
 * positions as given), final_score/best_score n_traj doubles, best_step n_traj int32 (-1 = initial),
 * outcomes n_traj*4 uint32 in the order REJECT, ACCEPT_WORSENED, ACCEPT_UNCHANGED, ACCEPT_IMPROVED. */
int adp_mc_run(adp_engine *, const adp_problem *, int n_traj, uint64_t seed, uint64_t traj_id0, int n_steps,
               int thermostat_kind, const double *thermostat_args,
               uint8_t *final_seq, double *final_score, uint8_t *best_seq, double *best_score,
               int32_t *best_step, uint32_t *outcomes);

/* adp_mc_run plus per-step trajectory records in the columns of the reference's TsvTrajectoryReporter (src/sampling.cc:430-489), so
 * that a GPU run can be written out in the on-disk format the reference's tools read (tools/utils.py:27-51).  Every
 * trace_interval-th step (step % trace_interval == 0) of every trajectory appends one record, cap = ceil(n_steps / trace_interval)
 * records per trajectory:
 *   trace_values: n_traj * cap * adp_mc_trace_width(problem) doubles -- step, current_score, proposed_score, score_diff,
 *                 temperature, metropolis_criterion, random_threshold, outcome (0 REJECT .. 3 ACCEPT_IMPROVED), then the unweighted
 *                 term values [context][term] of the last evaluated proposal (the reference's score_table)
 *   trace_seqs:   n_traj * cap * 2 * n bytes -- current sequence (after the decision), proposed sequence
 * trace_interval = 0: no records (the buffers may be NULL). */
int adp_mc_trace_width(const adp_problem *);     /* 8 + max(n_ctx,1) * (n_terms + n_bp_terms) */
int adp_mc_run_trace(adp_engine *, const adp_problem *, int n_traj, uint64_t seed, uint64_t traj_id0, int n_steps,
                     int thermostat_kind, const double *thermostat_args,
                     uint8_t *final_seq, double *final_score, uint8_t *best_seq, double *best_score,
                     int32_t *best_step, uint32_t *outcomes,
                     int trace_interval, double *trace_values, uint8_t *trace_seqs);

/* ≙ mutate_recursively(device, position, nucleotide) (reference src/sampling.cc:205-282) and can_be_freely_mutated (:177-193) for B
 * independent (sequence, position, nucleotide) triples of one design problem: out_seqs[b] = seqs[b] with `position` set to
 * `nucleotide` and every base-paired partner, in every macrostate, set to the complementary / same nucleotide in turn.  The
 * closures come from the table compiler and are applied ON THE DEVICE by the code the Monte-Carlo kernel uses for its proposals.
 * positions (0-based) and nucleotides ('A','C','G','U') are host arrays; freely (may be NULL) receives can_be_freely_mutated.
 * Fails (ADP_ERR_ARG, text in adp_last_error) where the reference throws: partner not mutable, unbalanced brackets. */
int adp_mutate_batch(adp_engine *, const adp_problem *, int B, const uint8_t *seqs, const int32_t *positions,
                     const uint8_t *nucleotides, uint8_t *out_seqs, int32_t *freely_mutable);

/* Per-rank reduction at the end of a sharded design run (the reference merges the TSV files of its N processes offline,
 * tools/utils.py:27-51; here every rank keeps its k best trajectories and only those records cross NVLink): the k (<= 1024) best of
 * n scores ON THE DEVICE, best first, ties by lower index, NaN last.  scores: n doubles (device memory in a run; host works too).
 * top_index: min(k,n) int32 positions into scores; top_score: the same many doubles or NULL. */
int adp_top_k(adp_engine *, int n, const double *scores, int k, int32_t *top_index, double *top_score);

/* ---- measurement helpers ---------------------------------------------------------------------- */
/* Sustained FP64 FMA rate of this GPU (GFMA/s = 1e9 fused multiply-adds per second), measured with a
 * register-resident DFMA loop; the denominator of the SM-issue roofline in bench.py. */
int adp_measure_dfma_rate(adp_engine *, double *gfma_per_s);

#ifdef __cplusplus
}
#endif
#endif
