// tests/emu/emu.cc -- TEST INFRASTRUCTURE ONLY; never part of libaddapt_b200.so.
//
// Compiles the engine's kernel bodies (addapt_b200/csrc/*_core.cuh, fold_jobs.cuh) for the host with the
// single-thread SerialExec policy, so the recursions, constraint handling and traceback can be checked
// against the oracle on a machine without a GPU (pytest -m "not gpu").  It exercises the same source the
// sm_100a kernels are built from, but not their parallel decomposition -- that is what the -m gpu tests do.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../addapt_b200/csrc/constraints.h"
#include "../../addapt_b200/csrc/design_core.cuh"
#include "../../addapt_b200/csrc/energy_model.h"
#include "../../addapt_b200/csrc/fold_jobs.cuh"
#include "../../addapt_b200/csrc/problem_host.h"

using namespace adp;

static std::string g_err;

extern "C" {

const char* emu_error() { return g_err.c_str(); }

void* emu_model_load(const char* par) {
  try {
    auto* m = new EnergyModel();
    m->load(par);
    return m;
  } catch (std::string& e) { g_err = e; return nullptr; }
}

double emu_kT(void* model) { return ((EnergyModel*)model)->kT_kcal(); }

int emu_eval(void* model, const char* seq, const char* db, int* e) {
  try { *e = ((EnergyModel*)model)->eval_structure(seq, db); return 0; } catch (std::string& er) { g_err = er; return -1; }
}

int emu_motif_contrib(void* model, const char* mseq, const char* mfold, double kd_uM, int* c) {
  auto* M = (EnergyModel*)model;
  try { *c = M->make_motif(mseq, mfold, M->kT_kcal() * std::log(kd_uM / 1e6)).contrib_dcal; return 0; } catch (std::string& er) { g_err = er; return -1; }
}

// `fast` = 1: the fast path (pf_fast.cuh); returns 1 when it declines the sequence (letters outside ACGU/T)
int emu_pf2(void* model, int n, const char* seq, const char* cst, const char* mseq, const char* mfold, double kd_uM,
            int outside, double* dG, double* bppm, int fast, int n_query, const int32_t* qi, const int32_t* qj, double* qprobs) {
  auto* M = (EnergyModel*)model;
  try {
    ConstraintDev cd;
    MotifDev md;
    if (cst) build_constraint(cst, n, &cd);
    if (mseq) md = M->make_motif(mseq, mfold, M->kT_kcal() * std::log(kd_uM / 1e6));
    PfJob job{};
    job.seq = (const uint8_t*)seq; job.n = n; job.cst = cst ? &cd : nullptr; job.motif = mseq ? &md : nullptr;
    job.T = &M->pf(); job.outside = outside != 0; job.dG = dG; job.bppm = bppm;
    job.n_query = outside ? n_query : 0; job.qi = qi; job.qj = qj; job.qprobs = qprobs;
    SerialExec ex;
    if (fast) {
      std::vector<char> small(fast_small_bytes(n) + 64);
      std::vector<double> mats(fast_matrix_doubles(n) + 2);
      std::vector<double> dstate((size_t)2 * fast_R(n) * kFastLevels + 2);
      return run_pf_job_fast(ex, job, small.data(), mats.data(), dstate.data()) ? 0 : 1;
    }
    std::vector<char> small(pf_small_bytes(n) + 64);
    std::vector<double> mats(pf_matrix_doubles(n, job.outside) + 2);
    run_pf_job(ex, job, small.data(), mats.data());
    return 0;
  } catch (std::string& er) { g_err = er; return -1; }
}

int emu_pf(void* model, int n, const char* seq, const char* cst, const char* mseq, const char* mfold, double kd_uM,
           int outside, double* dG, double* bppm) {
  auto* M = (EnergyModel*)model;
  try {
    ConstraintDev cd;
    MotifDev md;
    if (cst) build_constraint(cst, n, &cd);
    if (mseq) md = M->make_motif(mseq, mfold, M->kT_kcal() * std::log(kd_uM / 1e6));
    PfJob job{};
    job.seq = (const uint8_t*)seq; job.n = n; job.cst = cst ? &cd : nullptr; job.motif = mseq ? &md : nullptr;
    job.T = &M->pf(); job.outside = outside != 0; job.dG = dG; job.bppm = bppm;
    std::vector<char> small(pf_small_bytes(n) + 64);
    std::vector<double> mats(pf_matrix_doubles(n, job.outside) + 2);
    SerialExec ex;
    run_pf_job(ex, job, small.data(), mats.data());
    return 0;
  } catch (std::string& er) { g_err = er; return -1; }
}

// fast = 1: mfe_fast.cuh; returns 1 when it declines the sequence
int emu_mfe2(void* model, int n, const char* seq, const char* cst, const char* mseq, const char* mfold, double kd_uM,
             int32_t* e, char* structure, int fast) {
  auto* M = (EnergyModel*)model;
  try {
    ConstraintDev cd;
    MotifDev md;
    if (cst) build_constraint(cst, n, &cd);
    if (mseq) md = M->make_motif(mseq, mfold, M->kT_kcal() * std::log(kd_uM / 1e6));
    MfeJob job{};
    job.seq = (const uint8_t*)seq; job.n = n; job.cst = cst ? &cd : nullptr; job.motif = mseq ? &md : nullptr;
    job.T = &M->mfe(); job.e_dcal = e; job.structure = structure;
    SerialExec ex;
    if (fast) {
      std::vector<char> small(mfe_fast_small_bytes(n) + 64);
      std::vector<int32_t> mats(mfe_fast_matrix_ints(n) + 2);
      std::vector<int32_t> dstate((size_t)2 * fast_R(n) * kFastLevels + 2);
      return run_mfe_job_fast(ex, job, small.data(), mats.data(), dstate.data()) ? 0 : 1;
    }
    std::vector<char> small(mfe_small_bytes(n) + 64);
    std::vector<int32_t> mats(mfe_matrix_ints(n) + 2);
    run_mfe_job(ex, job, small.data(), mats.data());
    return 0;
  } catch (std::string& er) { g_err = er; return -1; }
}

int emu_mfe(void* model, int n, const char* seq, const char* cst, const char* mseq, const char* mfold, double kd_uM,
            int32_t* e, char* structure) {
  auto* M = (EnergyModel*)model;
  try {
    ConstraintDev cd;
    MotifDev md;
    if (cst) build_constraint(cst, n, &cd);
    if (mseq) md = M->make_motif(mseq, mfold, M->kT_kcal() * std::log(kd_uM / 1e6));
    MfeJob job{};
    job.seq = (const uint8_t*)seq; job.n = n; job.cst = cst ? &cd : nullptr; job.motif = mseq ? &md : nullptr;
    job.T = &M->mfe(); job.e_dcal = e; job.structure = structure;
    std::vector<char> small(mfe_small_bytes(n) + 64);
    std::vector<int32_t> mats(mfe_matrix_ints(n) + 2);
    SerialExec ex;
    run_mfe_job(ex, job, small.data(), mats.data());
    return 0;
  } catch (std::string& er) { g_err = er; return -1; }
}

// score of B sequences of one design problem (adp_problem layout of include/addapt_b200.h)
int emu_score(void* model, const adp_problem* pr, int B, const char* seqs, double* score, double* terms, int fast) {
  auto* M = (EnergyModel*)model;
  try {
    CompiledProblem cp;
    compile_problem(*pr, M->kT_kcal(), &cp);
    MotifDev md;
    if (pr->motif.seq) { md = M->make_motif(pr->motif.seq, pr->motif.fold, M->kT_kcal() * std::log(pr->motif.kd_uM / 1e6)); cp.sp.motif = &md; }
    cp.sp.csts = cp.csts.data();
    cp.sp.T = &M->pf();
    const int nm = cp.sp.n_max;
    std::vector<char> small(std::max(pf_small_bytes(nm), fast_small_bytes(nm)) + 64), extra(design_extra_bytes(nm) + 256);
    std::vector<double> mats(std::max(pf_matrix_doubles(nm, cp.needs_outside()), fast_matrix_doubles(nm)) + 2);
    std::vector<double> dstate((size_t)2 * fast_R(nm) * kFastLevels + 2);
    Carver cv(extra.data());
    DesignScratch d;
    carve_design(cv, d, nm);
    std::vector<double> bp_out(kMaxBp + kMaxDensPos);
    d.bp_out = bp_out.data();
    SerialExec ex;
    const int nt = cp.sp.n_ctx * cp.sp.n_tot();
    for (int b = 0; b < B; b++) {
      std::memcpy(d.cur, seqs + (size_t)b * pr->n, pr->n);
      bool declined = false;
      score[b] = fast ? score_construct<true>(ex, cp.sp, d.cur, d, small.data(), mats.data(), &declined, dstate.data())
                      : score_construct<false>(ex, cp.sp, d.cur, d, small.data(), mats.data());
      if (declined) { g_err = "fast path declined the sequence"; return 1; }
      if (terms) for (int t = 0; t < nt; t++) terms[(size_t)b * nt + t] = d.terms[t];
    }
    return 0;
  } catch (std::string& er) { g_err = er; return -1; }
}

int emu_mc_trace(void* model, const adp_problem* pr, uint64_t seed, uint64_t traj_id, int n_steps, int kind, const double* targs,
                 char* final_seq, double* final_score, char* best_seq, double* best_score, int32_t* best_step, uint32_t* outcomes, int fast,
                 int interval, double* trace_vals, uint8_t* trace_seqs);

int emu_mc(void* model, const adp_problem* pr, uint64_t seed, uint64_t traj_id, int n_steps, int kind, const double* targs,
           char* final_seq, double* final_score, char* best_seq, double* best_score, int32_t* best_step, uint32_t* outcomes, int fast) {
  return emu_mc_trace(model, pr, seed, traj_id, n_steps, kind, targs, final_seq, final_score, best_seq, best_score, best_step, outcomes, fast, 0,
                      nullptr, nullptr);
}

// ≙ adp_mutate_batch for one triple: the host-compiled closure applied by the device code path (apply_closure)
int emu_mutate(const adp_problem* pr, const char* seq, int pos, char nt, char* out, int* freely) {
  try {
    std::vector<MoveTableDev> mv(1);
    std::memset(&mv[0], 0, sizeof(MoveTableDev));
    std::vector<int> order, parity;
    compile_closure(*pr, pos, &order, &parity);
    store_closure(&mv[0], 0, pos, order, parity);
    std::memcpy(out, seq, pr->n);
    apply_closure(mv[0], 0, (uint8_t)nt, (uint8_t*)out);
    if (freely) *freely = freely_mutable(*pr, pos) ? 1 : 0;
    return 0;
  } catch (std::string& er) { g_err = er; return -1; }
}

int emu_mc_trace(void* model, const adp_problem* pr, uint64_t seed, uint64_t traj_id, int n_steps, int kind, const double* targs,
                 char* final_seq, double* final_score, char* best_seq, double* best_score, int32_t* best_step, uint32_t* outcomes, int fast,
                 int interval, double* trace_vals, uint8_t* trace_seqs) {
  auto* M = (EnergyModel*)model;
  try {
    CompiledProblem cp;
    compile_problem(*pr, M->kT_kcal(), &cp);
    MotifDev md;
    if (pr->motif.seq) { md = M->make_motif(pr->motif.seq, pr->motif.fold, M->kT_kcal() * std::log(pr->motif.kd_uM / 1e6)); cp.sp.motif = &md; }
    cp.sp.csts = cp.csts.data();
    cp.sp.T = &M->pf();
    std::vector<MoveTableDev> mv(1);
    compile_moves(*pr, &mv[0]);
    const int nm = cp.sp.n_max;
    std::vector<char> small(std::max(pf_small_bytes(nm), fast_small_bytes(nm)) + 64), extra(design_extra_bytes(nm) + 256);
    std::vector<double> mats(std::max(pf_matrix_doubles(nm, cp.needs_outside()), fast_matrix_doubles(nm)) + 2);
    std::vector<double> dstate((size_t)2 * fast_R(nm) * kFastLevels + 2);
    Carver cv(extra.data());
    DesignScratch d;
    carve_design(cv, d, nm);
    std::vector<double> train(kMaxTrain);
    d.train = train.data();
    std::vector<double> bp_out(kMaxBp + kMaxDensPos), terms_state(kMaxCtx * kMaxTerms);
    d.bp_out = bp_out.data();
    McTrace tr;
    tr.interval = interval;
    tr.cap = interval > 0 ? (n_steps + interval - 1) / interval : 0;
    tr.vals = trace_vals;
    tr.seqs = trace_seqs;
    tr.terms_state = terms_state.data();
    SerialExec ex;
    std::memcpy(d.cur, pr->init_seq, pr->n);
    McState res;
    double ta[3] = {targs[0], kind >= 1 ? targs[1] : 0., kind >= 1 ? targs[2] : 1.};
    // two launches' worth of steps through the carried state, like a chunked adp_mc_run
    const int cut = n_steps / 2;
    for (int part = 0; part < 2; part++) {
      const int s0 = part ? cut : 0, s1 = part ? n_steps : cut;
      if (part && cut == 0) break;
      if (fast) mc_trajectory<true>(ex, cp.sp, mv[0], seed, traj_id, s0, part || cut ? s1 : n_steps, kind, ta, d, small.data(), mats.data(), &res, dstate.data(), &tr);
      else mc_trajectory<false>(ex, cp.sp, mv[0], seed, traj_id, s0, part || cut ? s1 : n_steps, kind, ta, d, small.data(), mats.data(), &res, nullptr, &tr);
    }
    std::memcpy(final_seq, d.cur, pr->n);
    std::memcpy(best_seq, d.best, pr->n);
    *final_score = res.cur_score; *best_score = res.best_score; *best_step = res.best_step;
    for (int k = 0; k < 4; k++) outcomes[k] = res.outcomes[k];
    return 0;
  } catch (std::string& er) { g_err = er; return -1; }
}

// Shared-memory footprint of one construct, as the launch planners size it (fold_kernels.cu plan_pf_fast / plan_mfe_fast,
// design_kernels.cu plan_design): kind 0 = fast PF fold, 1 = fast PF fold inside the score / Monte-Carlo kernels, 2 = fast MFE.
// tests/test_emu_parity.py keeps the benchmarked lengths at two (PF) / four (MFE) constructs per SM.
size_t emu_smem_bytes(int kind, int n) {
  if (kind == 2) return ((mfe_fast_small_bytes(n) + 15) & ~(size_t)15) + mfe_fast_matrix_ints(n) * sizeof(int32_t);
  const size_t small = (fast_small_bytes(n) + 15) & ~(size_t)15;
  const size_t mats = fast_matrix_doubles(n) * sizeof(double);
  return kind == 1 ? small + design_extra_bytes(n) + 64 + mats : small + mats;
}

}  // extern "C"
