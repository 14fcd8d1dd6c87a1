// Score function and Monte-Carlo step on the device, one design (trajectory) per CTA.
//
//   score_construct  ≙ ScoreFunction::evaluate / evaluate_terms + MacrostateProbTerm::evaluate
//                      (reference src/scoring.cc:114-158, 233-259): for every context, the apo and holo
//                      ensembles are filled unconstrained and under each macrostate constraint
//                      (ViennaRnaFold::macrostate_prob, :53-71), P = exp((G_tot - G_cst)/kT), term = log P or
//                      log(1-P), score = sum of weight * term, summed over contexts.
//   mc_trajectory    ≙ MonteCarlo::apply (src/sampling.cc:22-107) with UnbiasedMutationMove (:287-303),
//                      mutate_recursively (:205-282, precompiled on the host into per-position closures),
//                      Fixed / Annealing thermostats (:305-338) and the Metropolis rule (:76-89).
// Random numbers: Philox4x32-10, key = seed, counter = (step, 0, trajectory id lo, hi); word 0 picks the
// position, word 1 the nucleotide, words 2-3 the 53-bit acceptance threshold.
#pragma once
#include "fold_jobs.cuh"

namespace adp {

constexpr int kMaxCtx = 8;
constexpr int kMaxFlank = 64;
constexpr int kMaxTerms = 16;
constexpr int kMaxMacro = 8;
constexpr int kMaxClosure = 16;     // positions changed together by one mutation (a base-pair chain)
constexpr int kMaxTrain = 512;      // longest training window of the auto-scaling thermostat
constexpr int kMaxBp = 8;           // base-pair-probability terms (adp_bp_term)
constexpr int kMaxDensPos = 64;     // positions of all density terms together

struct ScoreProblemDev {
  int32_t n;                         // raw device length
  int32_t n_ctx;                     // >= 1 (a bare device is one context with empty flanks)
  int32_t nb[kMaxCtx], na[kMaxCtx];
  uint8_t before[kMaxCtx][kMaxFlank], after[kMaxCtx][kMaxFlank];
  int32_t n_macro;
  const ConstraintDev* csts;         // [n_ctx][n_macro], already padded with '.' over the flanks
  int32_t n_terms;
  int32_t term_fav[kMaxTerms];
  int32_t term_v0[kMaxTerms];        // variant index of the unconstrained fill of the term's condition
  int32_t term_v1[kMaxTerms];        // variant index of the constrained fill
  double term_w[kMaxTerms];
  int32_t V;
  VariantDev variants[kMaxVariants]; // cst_id = macrostate index or -1
  int32_t float_dg;
  double kT;                         // kcal/mol
  const MotifDev* motif;
  const PfTables* T;
  int32_t n_max;                     // longest contextual length
  // base-pair-probability terms (adp_bp_term), evaluated from the outside pass of the unconstrained fill of their condition
  int32_t n_bp;
  int32_t bp_kind[kMaxBp];           // 0 pair, 1 density
  int32_t bp_v[kMaxBp];              // variant index of that fill (its `outside` flag is set)
  int32_t bp_fav[kMaxBp];
  int32_t bp_at[kMaxBp];             // kind 0: index of the term's query; kind 1: offset of its positions in the density list
  int32_t bp_npos[kMaxBp];           // kind 1: number of positions
  double bp_w[kMaxBp];
  int32_t n_bpq, n_bpd;              // pair queries / density positions in total
  const int32_t* bp_qi;              // [n_ctx][n_bpq] contextual 0-based positions (device memory)
  const int32_t* bp_qj;
  const int32_t* bp_dpos;            // [n_ctx][n_bpd]
  // Holo fills that are derived from their apo twin (fast path): twin[v] = index of the unconstrained / same-macrostate APO
  // variant of a ligand-bound variant without outside pass, else -1; site_of[v] = 1 for an apo variant some holo variant derives from
  int32_t twin[kMaxVariants], site_of[kMaxVariants];
  ADP_HD int n_tot() const { return n_terms + n_bp; }      // term values per context: macrostate terms, then bp terms
};

// Mutation proposals, precompiled on the host from the macrostate brackets.
struct MoveTableDev {
  int32_t n_free;                            // freely mutable positions (sampling.cc:177-193)
  int16_t free_pos[kMaxLen];
  int8_t closure_len[kMaxLen];               // per free position: how many positions change together
  int16_t closure_pos[kMaxLen][kMaxClosure];
  int8_t closure_comp[kMaxLen][kMaxClosure]; // 0: gets the drawn nucleotide, 1: gets its complement
};

ADP_HD void philox4x32_10(uint32_t k0, uint32_t k1, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t out[4]) {
  for (int r = 0; r < 10; r++) {
    const uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1, n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

ADP_HD uint8_t complement_nt(uint8_t c) { return c == 'A' ? 'U' : c == 'U' ? 'A' : c == 'G' ? 'C' : c == 'C' ? 'G' : 0; }

// The closure of table slot `slot` written into `seq`: the drawn nucleotide, or its complement, per position
// (≙ mutate_recursively, reference src/sampling.cc:205-282; the table is compiled on the host, problem_host.h).
ADP_HD bool closure_changes(const MoveTableDev& mv, int slot, uint8_t nt, const uint8_t* cur) {
  const uint8_t ntc = complement_nt(nt);
  bool changed = false;
  for (int q = 0; q < mv.closure_len[slot]; q++)
    if (cur[mv.closure_pos[slot][q]] != (mv.closure_comp[slot][q] ? ntc : nt)) changed = true;
  return changed;
}
ADP_HD void apply_closure(const MoveTableDev& mv, int slot, uint8_t nt, uint8_t* seq) {
  const uint8_t ntc = complement_nt(nt);
  for (int q = 0; q < mv.closure_len[slot]; q++) seq[mv.closure_pos[slot][q]] = mv.closure_comp[slot][q] ? ntc : nt;
}


// products and sums that must not be contracted into an FMA, so the schedule matches a plain C evaluation
ADP_HD double mul_rn(double a, double b) {
#if defined(__CUDA_ARCH__)
  return __dmul_rn(a, b);
#else
  return a * b;
#endif
}
ADP_HD double add_rn(double a, double b) {
#if defined(__CUDA_ARCH__)
  return __dadd_rn(a, b);
#else
  return a + b;
#endif
}

ADP_HD double thermostat_T(int kind, const double* args, int step) {
  if (kind != 1) return args[0];      // fixed; the auto-scaling kind is handled by its caller
  const int cycle = (int)args[2];
  return add_rn(mul_rn((args[1] - args[0]) / cycle, (double)(step % cycle)), args[0]);   // sampling.cc:331-338
}

// Shared-memory bytes a design CTA needs besides the fold working set.
ADP_HD size_t design_extra_bytes(int n_max) { return 112 + (size_t)(n_max + 16) * 4 + (2 * kMaxVariants + kMaxTerms * kMaxCtx + 8) * sizeof(double); }

struct DesignScratch {
  uint8_t* full;     // contextual sequence  [n_max+16]
  uint8_t* cur;      // current raw sequence [n_max+16]
  uint8_t* prop;     // proposed raw sequence
  uint8_t* best;
  double* dg;        // [kMaxVariants]
  double* terms;     // [kMaxCtx*kMaxTerms]
  double* bcast;     // [8] broadcast slots
  double* ratio;     // [kMaxVariants] Z_holo / Z_apo reported by the apo fills (PfJob::site_ratio)
  double* bp_out;    // [kMaxBp + kMaxDensPos] pair-query / paired-density results of the last fill (device memory; NULL without bp terms)
  double* train;     // [kMaxTrain] score differences the auto-scaling thermostat is trained on: the trajectory's slice of the
                     // device-memory window that is carried from launch to launch (NULL for the other thermostats); touched once a step
};

ADP_HD void carve_design(Carver& cv, DesignScratch& d, int n_max) {
  d.full = cv.take<uint8_t>(n_max + 16);
  d.cur = cv.take<uint8_t>(n_max + 16);
  d.prop = cv.take<uint8_t>(n_max + 16);
  d.best = cv.take<uint8_t>(n_max + 16);
  d.dg = cv.take<double>(kMaxVariants);
  d.terms = cv.take<double>(kMaxCtx * kMaxTerms);
  d.bcast = cv.take<double>(8);
  d.ratio = cv.take<double>(kMaxVariants);
  d.bp_out = nullptr;     // set by the caller
  d.train = nullptr;      // set by the caller (not in shared memory: 4 KB that the fold working set needs)
}

// Score of the raw sequence `raw` (n ASCII bytes, readable by every thread).  Collective; every thread gets
// the score.  Unweighted term values are left in d.terms[ctx*n_terms + t].
// kFast: fold with pf_fast.cuh (fold_small / mats sized by fast_small_bytes / fast_matrix_doubles); *declined is
// set (and the score is meaningless) when a sequence holds letters outside ACGU/T.
template <bool kFast, class E>
ADP_HD double score_construct(const E& ex, const ScoreProblemDev& sp, const uint8_t* raw, DesignScratch& d, void* fold_small, double* mats,
                              bool* declined = nullptr, double* dstate = nullptr, double* qm_ext = nullptr) {
  double score = 0.;
  const int nT = sp.n_tot();
  for (int x = 0; x < sp.n_ctx; x++) {
    const int nb = sp.nb[x], na = sp.na[x], nx = nb + sp.n + na;
    for (int p = ex.tid(); p < nx; p += ex.nthreads())
      d.full[p] = p < nb ? sp.before[x][p] : p < nb + sp.n ? raw[p - nb] : sp.after[x][p - nb - sp.n];
    ex.sync();
    // Two rounds over the fills.  Round 0 folds every ensemble that has to be folded; an apo fill that a ligand-bound variant
    // derives from also reports Z_holo / Z_apo (site mode of the outside pass: a fraction of a fill).  Round 1 completes the
    // ligand-bound variants: dG_holo = dG_apo - kT ln(ratio), or a real fill when the motif occurs more than once (ratio = NaN).
    for (int round = 0; round < (kFast ? 2 : 1); round++)
    for (int v = 0; v < sp.V; v++) {
      const VariantDev var = sp.variants[v];
      const int twin = kFast ? sp.twin[v] : -1;
      if (round == 0 && twin >= 0) continue;
      if (round == 1) {
        if (twin < 0) continue;
        const double r = d.ratio[twin];
        if (r == r) {
          if (ex.tid() == 0) d.dg[v] = d.dg[twin] - sp.kT * log(r);
          continue;
        }
      }
      PfJob job;
      job.seq = d.full;
      job.n = nx;
      job.cst = var.cst_id >= 0 ? sp.csts + (size_t)x * sp.n_macro + var.cst_id : nullptr;
      job.motif = var.use_motif ? sp.motif : nullptr;
      job.T = sp.T;
      job.outside = var.outside != 0;
      job.dG = d.dg + v;
      job.bppm = nullptr;
      job.n_query = 0;
      job.qi = nullptr; job.qj = nullptr; job.qprobs = nullptr;
      if (kFast && sp.site_of[v] && !var.outside) { job.site_motif = sp.motif; job.site_ratio = d.ratio + v; }
      if (var.outside) {      // base-pair-probability terms read this fill's pair queries and paired densities
        job.n_query = sp.n_bpq;
        job.qi = sp.bp_qi + (size_t)x * sp.n_bpq;
        job.qj = sp.bp_qj + (size_t)x * sp.n_bpq;
        job.qprobs = d.bp_out;
        job.n_dens = sp.n_bpd;
        job.dens_pos = sp.bp_dpos + (size_t)x * sp.n_bpd;
        job.dens = d.bp_out + kMaxBp;
      }
      if constexpr (kFast) {
        if (!run_pf_job_fast<E, true>(ex, job, fold_small, mats, dstate, qm_ext)) {
          if (declined) *declined = true;
          return 0.;
        }
      } else {
        run_pf_job(ex, job, fold_small, mats);
      }
      ex.sync();
      if (var.outside && ex.tid() == 0) {
        for (int t = 0; t < sp.n_bp; t++) {
          if (sp.bp_v[t] != v) continue;
          double p;
          if (sp.bp_kind[t] == 0) {
            p = d.bp_out[sp.bp_at[t]];
          } else {
            p = 0.;
            for (int q = 0; q < sp.bp_npos[t]; q++) p += d.bp_out[kMaxBp + sp.bp_at[t] + q];
            p /= sp.bp_npos[t];
          }
          if (!sp.bp_fav[t]) p = 1. - p;
          d.terms[x * nT + sp.n_terms + t] = log(p);
        }
      }
      if (var.outside) ex.sync();      // the next fill clears bp_out
    }
    if (ex.tid() == 0) {
      for (int t = 0; t < sp.n_terms; t++) {
        double g0 = d.dg[sp.term_v0[t]], g1 = d.dg[sp.term_v1[t]];
        if (sp.float_dg) { g0 = (double)(float)g0; g1 = (double)(float)g1; }   // vrna_pf returns float
        double p = exp((g0 - g1) / sp.kT);
        if (!sp.term_fav[t]) p = 1. - p;
        d.terms[x * nT + t] = log(p);
      }
    }
    ex.sync();
    for (int t = 0; t < sp.n_terms; t++) score = add_rn(score, mul_rn(sp.term_w[t], d.terms[x * nT + t]));
    for (int t = 0; t < sp.n_bp; t++) score = add_rn(score, mul_rn(sp.bp_w[t], d.terms[x * nT + sp.n_terms + t]));
  }
  return score;
}

// Everything a trajectory carries from one step to the next besides its current / best sequences.  A run may be split
// into several launches (steps [step0, step1) each): the random stream is keyed by the absolute step, so the split does
// not change a single decision.
struct McState {
  double cur_score, best_score;
  double T_auto, last_diff;          // auto-scaling thermostat: current temperature, most recent score difference
  double last_prop, last_crit, last_u;   // trajectory records: proposed score, Metropolis criterion and threshold of the last SCORED step
  double step_T, step_u;                 // temperature and acceptance threshold of the step being scored (parked here across the fold)
  int32_t best_step, ntrain;         // best_step = -1: the start sequence; ntrain: filled part of the training window
  uint32_t outcomes[4];              // REJECT, ACCEPT_WORSENED, ACCEPT_UNCHANGED, ACCEPT_IMPROVED
};

// Optional per-step records of a trajectory, in the columns of the reference's TsvTrajectoryReporter (src/sampling.cc:430-489):
// every `interval`-th step (step % interval == 0) writes, AFTER the accept / reject decision,
//   vals[kTraceScalars + nt]: step, current_score, proposed_score, score_diff, temperature, metropolis_criterion,
//                             random_threshold, outcome, then the nt = n_ctx * n_terms unweighted term values of the LAST evaluated
//                             proposal (the reference's score_table: it still holds a rejected proposal's terms)
//   seqs[2 n]:                current sequence, proposed sequence
// Like in the reference, the proposal columns keep the values of the last scored step across ACCEPT_UNCHANGED steps (which skip
// the evaluation); before the first scored step the reference leaves score_diff / criterion / threshold uninitialised, here they are 0.
constexpr int kTraceScalars = 8;
struct McTrace {
  int interval;            // 0: no records
  int cap;                 // records per trajectory: ceil(n_steps / interval)
  double* vals;            // this trajectory's [cap * (kTraceScalars + nt)]
  uint8_t* seqs;           // this trajectory's [cap * 2 * n]
  double* terms_state;     // this trajectory's [nt]: term values carried from launch to launch
};

// Steps [step0, step1) of one trajectory.  step0 == 0: d.cur holds the start sequence (upper case) and *st is initialised
// here; otherwise d.cur / d.best / d.train / *st hold what the previous launch left.  *st lives in shared memory; it is
// read by every thread at entry and rewritten by thread 0 at exit.  Collective.  kFast requires ACGU/T-only start
// sequence and contexts (checked on the host).
template <bool kFast, class E>
ADP_HD void mc_trajectory(const E& ex, const ScoreProblemDev& sp, const MoveTableDev& mv, uint64_t seed, uint64_t traj_id, int step0, int step1,
                          int thermo_kind, const double* thermo_args, DesignScratch& d, void* fold_small, double* mats, McState* st,
                          double* dstate = nullptr, const McTrace* trace = nullptr, double* qm_ext = nullptr) {
  const int n = sp.n;
  const int nt_terms = sp.n_ctx * sp.n_tot();
  const bool tracing = trace && trace->interval > 0;
  // The trajectory's scalars stay in *st (shared memory), not in registers: they are live across the fold, where every register
  // counts (with them in registers the kernel spilled 18 doubles and a Monte-Carlo step scored 29 % slower than a batch score
  // of the same sequence).  Thread 0 writes them; a barrier at the top of every step publishes the writes.
  if (step0 == 0) {
    // the start sequence is scored by the loop below as "step -1" (one call site of the fold code keeps the kernel small)
    if (ex.tid() == 0) {
      st->cur_score = st->best_score = 0.;
      st->best_step = -1;
      st->outcomes[0] = st->outcomes[1] = st->outcomes[2] = st->outcomes[3] = 0;
      st->ntrain = 0;
      st->T_auto = thermo_kind == 2 ? thermo_args[2] : 0.;
      st->last_diff = st->last_prop = st->last_crit = st->last_u = 0.;
    }
  } else if (tracing) {
    for (int t = ex.tid(); t < nt_terms; t += ex.nthreads()) d.terms[t] = trace->terms_state[t];
  }
  ex.sync();
  // one record, after the decision of `step` (collective; see McTrace)
  auto record = [&](int step, double T, int outcome) {
    if (!tracing || step % trace->interval != 0) return;
    const int r = step / trace->interval;
    if (r >= trace->cap) return;
    double* v = trace->vals + (size_t)r * (kTraceScalars + nt_terms);
    uint8_t* q = trace->seqs + (size_t)r * 2 * n;
    if (ex.tid() == 0) {
      v[0] = step; v[1] = st->cur_score; v[2] = st->last_prop; v[3] = st->last_diff; v[4] = T; v[5] = st->last_crit; v[6] = st->last_u; v[7] = outcome;
    }
    for (int t = ex.tid(); t < nt_terms; t += ex.nthreads()) v[kTraceScalars + t] = d.terms[t];
    for (int p = ex.tid(); p < n; p += ex.nthreads()) { q[p] = d.cur[p]; q[n + p] = d.prop[p]; }
  };
  // auto-scaling thermostat (kind 2, reference src/sampling.cc:371-401): args = {target acceptance rate, training period,
  // initial temperature}.  Every step appends the most recent score difference (0 before the first scored proposal) to the
  // training set; when it holds `period` values, T = max(median / ln(rate), 0) -- "median" being what nth_element leaves at
  // position size/2 -- and the set is cleared.
  const int period = thermo_kind == 2 ? (int)thermo_args[1] : 0;
  for (int step = step0 == 0 ? -1 : step0; step < step1 && (step < 0 || mv.n_free > 0); step++) {
    const bool start = step < 0;          // score the start sequence: proposal = current, always "accepted", not counted
    ex.sync();                            // thread 0's updates of *st in the previous step are visible from here on
    double T = 0., u = 0.;
    if (start) {
      for (int p = ex.tid(); p < n; p += ex.nthreads()) { d.prop[p] = d.cur[p]; d.best[p] = d.cur[p]; }
      ex.sync();
    } else {
      T = thermostat_T(thermo_kind, thermo_args, step);
      if (thermo_kind == 2) {
        const int ntrain = st->ntrain + 1;
        ex.sync();                        // everyone has read the count before thread 0 advances it
        if (ex.tid() == 0) { d.train[ntrain - 1] = st->last_diff; st->ntrain = ntrain; }
        if (ntrain >= period) {
          ex.sync();
          const int mid = ntrain / 2;
          for (int a = ex.tid(); a < ntrain; a += ex.nthreads()) {      // rank counting: exactly one element has rank mid
            const double va = d.train[a];
            int rank = 0;
            for (int b = 0; b < ntrain; b++) {
              const double vb = d.train[b];
              rank += (vb < va) || (vb == va && b < a);
            }
            if (rank == mid) d.bcast[0] = va;
          }
          ex.sync();
          if (ex.tid() == 0) {
            const double t = d.bcast[0] / log(thermo_args[0]);
            st->T_auto = t > 0. ? t : 0.;
            st->ntrain = 0;
          }
        }
        ex.sync();
        T = st->T_auto;
      }
      uint32_t r[4];
      philox4x32_10((uint32_t)seed, (uint32_t)(seed >> 32), (uint32_t)step, 0u, (uint32_t)traj_id, (uint32_t)(traj_id >> 32), r);
      const int slot = (int)(((uint64_t)r[0] * (uint64_t)mv.n_free) >> 32);
      const uint8_t nt = (uint8_t)("ACGU"[r[1] >> 30]);
      u = ((double)(r[2] >> 5) * 67108864. + (double)(r[3] >> 6)) / 9007199254740992.;
      // proposal = current with the closure of the drawn position rewritten
      for (int p = ex.tid(); p < n; p += ex.nthreads()) d.prop[p] = d.cur[p];
      ex.sync();
      const bool changed = closure_changes(mv, slot, nt, d.cur);
      if (ex.tid() == 0) apply_closure(mv, slot, nt, d.prop);
      ex.sync();
      if (!changed) {                                                          // ACCEPT_UNCHANGED: not scored
        if (ex.tid() == 0) st->outcomes[2]++;
        record(step, T, 2);
        continue;
      }
      if (ex.tid() == 0) { st->step_T = T; st->step_u = u; }                   // parked in shared memory across the fold
    }
    const double prop_s = score_construct<kFast>(ex, sp, d.prop, d, fold_small, mats, nullptr, dstate, qm_ext);
    if (start) {
      if (ex.tid() == 0) st->cur_score = st->best_score = st->last_prop = prop_s;
      continue;
    }
    T = st->step_T;
    u = st->step_u;
    const double cur_s = st->cur_score, best_s = st->best_score;
    const double diff = prop_s - cur_s;
    const double crit = exp(diff / T);
    const bool reject = crit < u;                                              // NaN compares false: accept
    ex.sync();                                                                 // everyone has read the scores
    if (ex.tid() == 0) {
      st->last_diff = diff; st->last_prop = prop_s; st->last_crit = crit; st->last_u = u;
      if (reject) st->outcomes[0]++;
      else {
        st->outcomes[diff > 0 ? 3 : 1]++;
        st->cur_score = prop_s;
        if (prop_s > best_s) { st->best_score = prop_s; st->best_step = step; }
      }
    }
    if (reject) { ex.sync(); record(step, T, 0); continue; }
    for (int p = ex.tid(); p < n; p += ex.nthreads()) d.cur[p] = d.prop[p];
    if (prop_s > best_s)
      for (int p = ex.tid(); p < n; p += ex.nthreads()) d.best[p] = d.prop[p];
    ex.sync();
    record(step, T, diff > 0 ? 3 : 1);
  }
  ex.sync();
  if (tracing)
    for (int t = ex.tid(); t < nt_terms; t += ex.nthreads()) trace->terms_state[t] = d.terms[t];
  ex.sync();
}

}  // namespace adp
