// tests/emu/emu_simt.cc -- TEST INFRASTRUCTURE ONLY; never part of libaddapt_b200.so.
//
// The kernel bodies run by an emulated CTA with all of its threads (simt.hh) under a seeded schedule.  Every entry point
// restates the job loop of the __global__ wrapper it stands for (fold_kernels.cu, design_kernels.cu): several jobs back to back
// on one CTA with the wrapper's barriers between them, so the hand-over from one job's epilogue to the next job's prologue is
// part of what is checked.  tests/test_simt_emulation.py asks for the same bits from every schedule and for the oracle's numbers.
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
#include "simt.hh"

using namespace adp;

static std::string g_simt_err;
static long g_drop_barrier = -1;      // fault injection for the next runs (simt.hh: Cta::drop_barrier)

namespace {

struct Jobs {
  std::vector<ConstraintDev> cd;
  MotifDev md;
  bool has_motif = false;
};

void prepare(EnergyModel* M, int J, int n, const char* const* csts, const char* mseq, const char* mfold, double kd_uM, Jobs* out) {
  out->cd.resize(J);
  for (int j = 0; j < J; j++)
    if (csts && csts[j]) build_constraint(csts[j], n, &out->cd[j]);
  if (mseq) { out->md = M->make_motif(mseq, mfold, M->kT_kcal() * std::log(kd_uM / 1e6)); out->has_motif = true; }
}

// Shared and scratch memory start out with whatever the previous kernel left there: fill the working buffers with garbage
// (the same for every schedule, so that what a run leaves behind can be compared), not with the zeros of a fresh allocation.
template <class V>
void poison(V& v, uint64_t key) {
  unsigned char* p = reinterpret_cast<unsigned char*>(v.data());
  uint64_t x = key * 0x9E3779B97F4A7C15ull + 0x632BE59BD9B4E019ull;
  for (size_t k = 0; k < v.size() * sizeof(v[0]); k++) {
    x ^= x << 13; x ^= x >> 7; x ^= x << 17;
    p[k] = (unsigned char)(x >> 32);
  }
}

// FNV-1a over a working buffer as the last job left it: every cell of the matrices, not only the outputs, must be schedule-independent
template <class V>
uint64_t fingerprint(const V& v, uint64_t h = 0xcbf29ce484222325ull) {
  const unsigned char* p = reinterpret_cast<const unsigned char*>(v.data());
  for (size_t k = 0; k < v.size() * sizeof(v[0]); k++) { h ^= p[k]; h *= 0x100000001b3ull; }
  return h;
}

void report(const simt::Cta& cta, long* stats) {
  if (!stats) return;
  stats[0] = cta.n_cta_barriers; stats[1] = cta.n_collectives; stats[2] = cta.n_switches;
}

template <int G>
bool run_general_pf(simt::Cta& cta, int threads, int policy, uint64_t seed, const std::vector<PfJob>& jobs, void* small, double* mats) {
  return cta.run(threads, policy, seed, [&](int t) {
    simt::SimtExec<G> ex{&cta, t};
    for (const PfJob& job : jobs) {       // pf_batch_kernel
      run_pf_job(ex, job, small, mats);
      ex.sync();
    }
  });
}

template <int G>
bool run_general_mfe(simt::Cta& cta, int threads, int policy, uint64_t seed, const std::vector<MfeJob>& jobs, void* small, int32_t* mats) {
  return cta.run(threads, policy, seed, [&](int t) {
    simt::SimtExec<G> ex{&cta, t};
    for (const MfeJob& job : jobs) {      // mfe_batch_kernel
      run_mfe_job(ex, job, small, mats);
      ex.sync();
    }
  });
}

}  // namespace

extern "C" {

const char* emu_simt_error() { return g_simt_err.c_str(); }

// every thread of the following runs skips its k-th CTA barrier (-1: none)
void emu_simt_drop_barrier(long k) { g_drop_barrier = k; }

// J partition-function jobs of one length on one CTA.  path: 0 = general (pf_core.cuh, `group` lanes per cell), 1 = fast
// (pf_fast.cuh), 2 = fast with Qm outside the matrix block (the kQmGlobal kernels), 3 = fast with the level sums in memory
// (constructs longer than threads/2).  declined[j] = 1 where the fast path hands the job back.  Returns 0, or -1 with emu_simt_error().
int emu_simt_pf(void* model, int J, int n, const char* seqs, const char* const* csts, const int* use_motif, const int* outside,
                const char* mseq, const char* mfold, double kd_uM, double* dG, double* bppm, int* declined, int path, int threads,
                int group, int policy, uint64_t seed, long* stats, uint64_t* mats_print) {
  auto* M = (EnergyModel*)model;
  try {
    Jobs pj;
    prepare(M, J, n, csts, mseq, mfold, kd_uM, &pj);
    const size_t ntri = (size_t)n * (n + 1) / 2;
    std::vector<PfJob> jobs(J);
    for (int j = 0; j < J; j++) {
      PfJob job{};
      job.seq = (const uint8_t*)seqs + (size_t)j * n; job.n = n;
      job.cst = (csts && csts[j]) ? &pj.cd[j] : nullptr;
      job.motif = (pj.has_motif && use_motif[j]) ? &pj.md : nullptr;
      job.T = &M->pf(); job.outside = outside[j] != 0; job.dG = dG + j;
      job.bppm = (bppm && job.outside) ? bppm + (size_t)j * ntri : nullptr;
      jobs[j] = job;
      declined[j] = 0;
    }
    simt::Cta cta;
    cta.drop_barrier = g_drop_barrier;
    bool ok;
    if (path == 0) {
      std::vector<char> small(pf_small_bytes(n) + 64);
      std::vector<double> mats(pf_matrix_doubles(n, true) + 2);
      poison(small, (uint64_t)0x5eed); poison(mats, (uint64_t)0x5eed + 1);
      ok = group == 4 ? run_general_pf<4>(cta, threads, policy, seed, jobs, small.data(), mats.data())
         : group == 2 ? run_general_pf<2>(cta, threads, policy, seed, jobs, small.data(), mats.data())
                      : run_general_pf<1>(cta, threads, policy, seed, jobs, small.data(), mats.data());
      *mats_print = fingerprint(mats);
    } else {
      if (path != 3 && n > fast_max_len(threads)) { g_simt_err = "construct too long for register-resident level sums at this thread count"; return -1; }
      std::vector<char> small(fast_small_bytes(n) + 64);
      std::vector<double> mats(fast_matrix_doubles(n) + 2), qm(t4_count(n) + 4), dstate((size_t)2 * fast_R(n) * kFastLevels + 2);
      poison(small, (uint64_t)0x5eed); poison(mats, (uint64_t)0x5eed + 1); poison(qm, (uint64_t)0x5eed + 2); poison(dstate, (uint64_t)0x5eed + 3);
      double* qm_ext = path == 2 ? qm.data() : nullptr;
      double* ds = path == 3 ? dstate.data() : nullptr;
      ok = cta.run(threads, policy, seed, [&](int t) {
        simt::SimtExec<1> ex{&cta, t};
        for (int j = 0; j < J; j++) {       // pf_fast_kernel
          const bool done = run_pf_job_fast(ex, jobs[j], small.data(), mats.data(), ds, qm_ext);
          if (!done && t == 0) declined[j] = 1;
          ex.sync();
        }
      });
      *mats_print = fingerprint(qm, fingerprint(mats));
    }
    report(cta, stats);
    if (!ok) { g_simt_err = cta.error; return -1; }
    return 0;
  } catch (std::string& er) { g_simt_err = er; return -1; }
}

// J MFE jobs (energy + structure) on one CTA.  path: 0 = general (mfe_core.cuh), 1 = fast (mfe_fast.cuh), 3 = fast with the
// level minima in memory.
int emu_simt_mfe(void* model, int J, int n, const char* seqs, const char* const* csts, const int* use_motif, const char* mseq,
                 const char* mfold, double kd_uM, int32_t* e, char* structures, int* declined, int path, int threads, int group, int policy,
                 uint64_t seed, long* stats, uint64_t* mats_print) {
  auto* M = (EnergyModel*)model;
  try {
    Jobs pj;
    prepare(M, J, n, csts, mseq, mfold, kd_uM, &pj);
    std::vector<MfeJob> jobs(J);
    for (int j = 0; j < J; j++) {
      MfeJob job{};
      job.seq = (const uint8_t*)seqs + (size_t)j * n; job.n = n;
      job.cst = (csts && csts[j]) ? &pj.cd[j] : nullptr;
      job.motif = (pj.has_motif && use_motif[j]) ? &pj.md : nullptr;
      job.T = &M->mfe(); job.e_dcal = e + j; job.structure = structures ? structures + (size_t)j * n : nullptr;
      jobs[j] = job;
      declined[j] = 0;
    }
    simt::Cta cta;
    cta.drop_barrier = g_drop_barrier;
    bool ok;
    if (path == 0) {
      std::vector<char> small(mfe_small_bytes(n) + 64);
      std::vector<int32_t> mats(mfe_matrix_ints(n) + 2);
      poison(small, (uint64_t)0x5eed); poison(mats, (uint64_t)0x5eed + 1);
      ok = group == 4 ? run_general_mfe<4>(cta, threads, policy, seed, jobs, small.data(), mats.data())
         : group == 2 ? run_general_mfe<2>(cta, threads, policy, seed, jobs, small.data(), mats.data())
                      : run_general_mfe<1>(cta, threads, policy, seed, jobs, small.data(), mats.data());
      *mats_print = fingerprint(mats);
    } else {
      if (path != 3 && n > fast_max_len(threads)) { g_simt_err = "construct too long for register-resident level minima at this thread count"; return -1; }
      std::vector<char> small(mfe_fast_small_bytes(n) + 64);
      std::vector<int32_t> mats(mfe_fast_matrix_ints(n) + 2), dstate((size_t)2 * fast_R(n) * kFastLevels + 2);
      poison(small, (uint64_t)0x5eed); poison(mats, (uint64_t)0x5eed + 1); poison(dstate, (uint64_t)0x5eed + 3);
      int32_t* ds = path == 3 ? dstate.data() : nullptr;
      ok = cta.run(threads, policy, seed, [&](int t) {
        simt::SimtExec<1> ex{&cta, t};
        for (int j = 0; j < J; j++) {       // mfe_fast_kernel
          const bool done = run_mfe_job_fast(ex, jobs[j], small.data(), mats.data(), ds);
          if (!done && t == 0) declined[j] = 1;
          ex.sync();
        }
      });
      *mats_print = fingerprint(mats);
    }
    report(cta, stats);
    if (!ok) { g_simt_err = cta.error; return -1; }
    return 0;
  } catch (std::string& er) { g_simt_err = er; return -1; }
}

// B designs scored back to back by one CTA (score_batch_kernel).  fast = 1: pf_fast.cuh fills, holo derived from the apo twin.
int emu_simt_score(void* model, const adp_problem* pr, int B, const char* seqs, double* score, double* terms, int fast, int qm_global,
                   int threads, int policy, uint64_t seed, long* stats) {
  auto* M = (EnergyModel*)model;
  try {
    CompiledProblem cp;
    compile_problem(*pr, M->kT_kcal(), &cp);
    MotifDev md;
    if (pr->motif.seq) { md = M->make_motif(pr->motif.seq, pr->motif.fold, M->kT_kcal() * std::log(pr->motif.kd_uM / 1e6)); cp.sp.motif = &md; }
    cp.sp.csts = cp.csts.data();
    cp.sp.T = &M->pf();
    const int nm = cp.sp.n_max, n = pr->n;
    if (fast && nm > fast_max_len(threads)) { g_simt_err = "construct too long for this thread count"; return -1; }
    std::vector<char> small(std::max(pf_small_bytes(nm), fast_small_bytes(nm)) + 64), extra(design_extra_bytes(nm) + 256);
    std::vector<double> mats(std::max(pf_matrix_doubles(nm, cp.needs_outside()), fast_matrix_doubles(nm)) + 2), qm(t4_count(nm) + 4);
    poison(small, (uint64_t)0x5eed); poison(mats, (uint64_t)0x5eed + 1); poison(qm, (uint64_t)0x5eed + 2); poison(extra, (uint64_t)0x5eed + 4);
    Carver cv(extra.data());
    DesignScratch d;
    carve_design(cv, d, nm);
    std::vector<double> bp_out(kMaxBp + kMaxDensPos);
    d.bp_out = bp_out.data();
    const int nt_terms = cp.sp.n_ctx * cp.sp.n_tot();
    std::vector<int> declined_any(1, 0);
    simt::Cta cta;
    cta.drop_barrier = g_drop_barrier;
    const bool ok = cta.run(threads, policy, seed, [&](int t) {
      simt::SimtExec<1> ex{&cta, t};
      for (int b = 0; b < B; b++) {
        for (int p = t; p < n; p += threads) d.cur[p] = (uint8_t)seqs[(size_t)b * n + p];
        ex.sync();
        bool declined = false;
        const double sc = fast ? score_construct<true>(ex, cp.sp, d.cur, d, small.data(), mats.data(), &declined, nullptr, qm_global ? qm.data() : nullptr)
                               : score_construct<false>(ex, cp.sp, d.cur, d, small.data(), mats.data());
        if (declined) {
          if (t == 0) declined_any[0] = 1;
        } else {
          if (t == 0) score[b] = sc;
          if (terms)
            for (int k = t; k < nt_terms; k += threads) terms[(size_t)b * nt_terms + k] = d.terms[k];
        }
        ex.sync();
      }
    });
    report(cta, stats);
    if (!ok) { g_simt_err = cta.error; return -1; }
    if (declined_any[0]) { g_simt_err = "fast path declined a sequence"; return 1; }
    return 0;
  } catch (std::string& er) { g_simt_err = er; return -1; }
}

// n_traj Monte-Carlo trajectories back to back on one CTA (mc_kernel), each in two launches' worth of steps through the carried
// state.  Outputs per trajectory.
int emu_simt_mc(void* model, const adp_problem* pr, uint64_t seed, uint64_t traj_id0, int n_traj, int n_steps, int kind, const double* targs,
                char* final_seq, double* final_score, char* best_seq, double* best_score, int32_t* best_step, uint32_t* outcomes, int fast,
                int threads, int policy, uint64_t sched_seed, long* stats) {
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
    const int nm = cp.sp.n_max, n = pr->n;
    if (fast && nm > fast_max_len(threads)) { g_simt_err = "construct too long for this thread count"; return -1; }
    std::vector<char> small(std::max(pf_small_bytes(nm), fast_small_bytes(nm)) + 64), extra(design_extra_bytes(nm) + 256);
    std::vector<double> mats(std::max(pf_matrix_doubles(nm, cp.needs_outside()), fast_matrix_doubles(nm)) + 2);
    poison(small, (uint64_t)0x5eed); poison(mats, (uint64_t)0x5eed + 1); poison(extra, (uint64_t)0x5eed + 4);
    Carver cv(extra.data());
    DesignScratch d;
    carve_design(cv, d, nm);
    std::vector<double> train((size_t)kMaxTrain * n_traj), bp_out(kMaxBp + kMaxDensPos);
    d.bp_out = bp_out.data();
    std::vector<McState> states(n_traj);
    McState st;                 // __shared__ in mc_kernel
    McTrace tr;                 // __shared__ in mc_kernel
    double ta[3] = {targs[0], kind >= 1 ? targs[1] : 0., kind >= 1 ? targs[2] : 1.};
    const int cut = n_steps / 2;
    simt::Cta cta;
    cta.drop_barrier = g_drop_barrier;
    bool ok = true;
    for (int part = 0; part < 2 && ok; part++) {      // two launches
      const int s0 = part ? cut : 0, s1 = part ? n_steps : cut;
      if (s1 <= s0) continue;
      ok = cta.run(threads, policy, sched_seed + part, [&](int t) {
        simt::SimtExec<1> ex{&cta, t};
        for (int k = 0; k < n_traj; k++) {
          if (s0 == 0) {
            for (int p = t; p < n; p += threads) d.cur[p] = (uint8_t)pr->init_seq[p];
          } else {
            for (int p = t; p < n; p += threads) {
              d.cur[p] = (uint8_t)final_seq[(size_t)k * n + p];
              d.best[p] = (uint8_t)best_seq[(size_t)k * n + p];
            }
            if (t == 0) st = states[k];
          }
          d.train = train.data() + (size_t)k * kMaxTrain;
          if (t == 0) { tr.interval = 0; tr.cap = 0; }
          ex.sync();
          if (fast) mc_trajectory<true>(ex, cp.sp, mv[0], seed, traj_id0 + (uint64_t)k, s0, s1, kind, ta, d, small.data(), mats.data(), &st, nullptr, &tr);
          else mc_trajectory<false>(ex, cp.sp, mv[0], seed, traj_id0 + (uint64_t)k, s0, s1, kind, ta, d, small.data(), mats.data(), &st, nullptr, &tr);
          ex.sync();
          for (int p = t; p < n; p += threads) {
            final_seq[(size_t)k * n + p] = (char)d.cur[p];
            best_seq[(size_t)k * n + p] = (char)d.best[p];
          }
          if (t == 0) {
            states[k] = st;
            final_score[k] = st.cur_score; best_score[k] = st.best_score; best_step[k] = st.best_step;
            for (int q = 0; q < 4; q++) outcomes[(size_t)k * 4 + q] = st.outcomes[q];
          }
          ex.sync();
        }
      });
    }
    report(cta, stats);
    if (!ok) { g_simt_err = cta.error; return -1; }
    return 0;
  } catch (std::string& er) { g_simt_err = er; return -1; }
}

// A deliberately broken body: thread t publishes a value and reads its neighbour's WITHOUT a barrier in between.  The emulator
// must give schedule-dependent results for it (tests/test_simt_emulation.py: the checker itself is checked).
int emu_simt_selftest_missing_barrier(int threads, int policy, uint64_t seed, int with_barrier, long* sum) {
  std::vector<long> cell(threads, -1), got(threads, 0);
  simt::Cta cta;
  const bool ok = cta.run(threads, policy, seed, [&](int t) {
    simt::SimtExec<1> ex{&cta, t};
    cell[t] = t;
    if (with_barrier) ex.sync();
    got[t] = cell[(t + 1) % threads];
    ex.sync();
  });
  if (!ok) { g_simt_err = cta.error; return -1; }
  *sum = 0;
  for (int t = 0; t < threads; t++) *sum += got[t];
  return 0;
}

// ... and one whose lanes disagree about a warp collective: half of every warp skips it
int emu_simt_selftest_divergent_collective(int threads, int policy, uint64_t seed) {
  simt::Cta cta;
  const bool ok = cta.run(threads, policy, seed, [&](int t) {
    simt::SimtExec<1> ex{&cta, t};
    double v = t;
    if (ex.lane() < 16) v = ex.warp_sum(v);
    ex.sync();
    (void)v;
  });
  if (!ok) { g_simt_err = cta.error; return -1; }
  return 0;
}

}  // extern "C"
