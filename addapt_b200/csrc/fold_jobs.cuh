// One fold job = one construct handled by one CTA: carve the working set out of a raw buffer, stage the
// context and tables, run the recursions, write the results.  Shared by the CUDA kernels (fold_kernels.cu)
// and by the single-thread emulation the CPU-only tests use (tests/emu).
#pragma once
#include "mfe_core.cuh"
#include "pf_core.cuh"
#include "pf_fast.cuh"
#include "mfe_fast.cuh"

namespace adp {

// ---------------------------------------------------------------------------------------------------
struct PfJob {
  const uint8_t* seq;
  int n;
  const ConstraintDev* cst;     // NULL = unconstrained
  const MotifDev* motif;        // NULL = apo
  const PfTables* T;
  bool outside;                 // also compute base-pair probabilities
  double* dG;                   // -> ensemble free energy, kcal/mol
  double* bppm;                 // NULL or n(n+1)/2 doubles, row-major upper triangle incl. diagonal, 0-based
  int n_query;                  // pair-probability queries (0-based i,j), answered into qprobs
  const int32_t* qi;
  const int32_t* qj;
  double* qprobs;
  int n_dens = 0;               // positions (0-based) whose probability of being paired is wanted, answered into dens
  const int32_t* dens_pos = nullptr;
  double* dens = nullptr;
  // Fast path, apo jobs only (motif == NULL): also report Z_holo / Z_apo for this ligand motif into *site_ratio -- 1 when the
  // motif does not occur, 1 + (boltz - 1) P(plain hairpin at the site) when it occurs once (site mode of fast_outside: only the
  // cells enclosing the site), NaN when it occurs more than once (the caller then folds the holo ensemble itself).
  const MotifDev* site_motif = nullptr;
  double* site_ratio = nullptr;
};

ADP_HD size_t pf_small_bytes(int n) { return sizeof(PfSmall) + 16 + ctx_bytes(n) + (size_t)9 * (n + 4) * sizeof(double); }
ADP_HD size_t pf_matrix_doubles(int n, bool outside) { return (size_t)(outside ? 3 : 2) * ((size_t)n * (n + 1) / 2); }

// `small` : shared memory (pf_small_bytes).  `mats`: shared or global memory (pf_matrix_doubles).
template <class E>
ADP_HD void run_pf_job(const E& ex, const PfJob& job, void* small, double* mats) {
  const int n = job.n;
  const size_t ntri = (size_t)n * (n + 1) / 2;
  Carver cv(small);
  PfSmall* P = cv.take<PfSmall>(1);
  SeqCtx c;
  carve_ctx(cv, c, n);
  PfWork w;
  w.qm1a = cv.take<double>(n + 4);
  w.qm1b = cv.take<double>(n + 4);
  w.q1 = cv.take<double>(n + 4);
  w.qln = cv.take<double>(n + 4);
  w.ext = cv.take<double>(n + 4);
  w.va = cv.take<double>(n + 4);
  w.vb = cv.take<double>(n + 4);
  w.vab = cv.take<double>(n + 4);
  w.qb = mats;
  w.qm = mats + ntri;
  w.pr = job.outside ? mats + 2 * ntri : nullptr;
  w.scale = job.T->scale;
  w.eMLb = job.T->eMLb;
  c.has_cst = job.cst != nullptr;
  c.motif_len = job.motif ? job.motif->len : 0;
  const double mboltz = job.motif ? job.motif->boltz : 1.;
  stage_small(ex, P, job.T);
  load_context(ex, c, job.seq, job.cst, job.motif);   // ends with a barrier

  const double Z = pf_inside(ex, c, *P, job.T, mboltz, w);
  if (ex.tid() == 0)
    *job.dG = Z > 0. ? (-log(Z) - n * log(job.T->pf_scale)) * job.T->kT / 1000. : (double)INFINITY;
  if (!job.outside) return;

  if (Z > 0.) {
    pf_outside(ex, c, *P, job.T, job.motif, mboltz, w);
  } else {
    for (size_t k = ex.tid(); k < ntri; k += ex.nthreads()) w.pr[k] = 0.;
    ex.sync();
  }
  if (job.bppm) {
    // row-major output: row i0 (0-based) starts at i0*n - i0*(i0-1)/2, entry (i0,j0) at + (j0-i0)
    for (int i = 1; i <= n; i++) {
      const size_t row = (size_t)(i - 1) * n - (size_t)(i - 1) * (i - 2) / 2;
      for (int j = i + ex.tid(); j <= n; j += ex.nthreads()) job.bppm[row + (j - i)] = j > i ? w.pr[tri(i, j)] : 0.;
    }
  }
  for (int q = ex.tid(); q < job.n_query; q += ex.nthreads()) {
    int a = job.qi[q], b = job.qj[q];
    if (a > b) { const int t = a; a = b; b = t; }
    job.qprobs[q] = (a != b && a >= 0 && b < n) ? w.pr[tri(a + 1, b + 1)] : 0.;
  }
  for (int q = ex.tid(); q < job.n_dens; q += ex.nthreads()) {
    const int p = job.dens_pos[q] + 1;
    double s = 0.;
    if (p >= 1 && p <= n) {
      for (int k = 1; k < p; k++) s += w.pr[tri(k, p)];
      for (int k = p + 1; k <= n; k++) s += w.pr[tri(p, k)];
    }
    job.dens[q] = s;
  }
}

// ---------------------------------------------------------------------------------------------------
// The fast path (pf_fast.cuh) for the same job.  `small`: fast_small_bytes(n) of shared memory; `mats`:
// fast_matrix_doubles(n) doubles (shared or global); `dstate`: host emulation only.  Returns false (on every
// thread, nothing written) when the sequence holds a letter outside ACGU/T -- the caller then uses run_pf_job.
// `qm_ext` (optional): put Qm there (global scratch, t4_count(n)+2 doubles) instead of behind Qb in `mats`.
// kOneCstPath: fold unconstrained jobs through the constrained instantiation too (valid: without a constraint the run lengths
// span the sequence and every pair is allowed).  The score / Monte-Carlo kernels alternate constrained and unconstrained fills;
// with one instantiation they keep ONE copy of the sweep (~30 KB of SASS) hot instead of two against a 32 KB instruction cache
// (ncu, round 2: stall_no_instruction 3.4 -> see profiles/r2_design_icache.md).
template <class E, bool kOneCstPath = false>
ADP_HD bool run_pf_job_fast(const E& ex, const PfJob& job, void* small, double* mats, double* dstate = nullptr, double* qm_ext = nullptr) {
  const int n = job.n;
#if defined(ADP_PHASE_PROFILE) && defined(__CUDA_ARCH__)
  const long long tp0_ = clock64();
#endif
  OffsetCarver cv(small);
  PfHot* H = cv.take<PfHot>(1);
  SeqCtx c;
  carve_ctx(cv, c, n);
  FastWork w;
  w.ld = fast_ld(n);
  w.R = fast_R(n);
  w.vbase = cv.take<double>((size_t)kFastVectors * w.ld);
  w.cn = cv.take<uint8_t>(w.ld);
  w.cp = cv.take<uint8_t>(w.ld);
  w.cells = cv.take<uint16_t>((size_t)2 * w.ld);
  w.iscell = cv.take<uint8_t>((size_t)2 * w.ld);
  w.counter = cv.take<int>(4);
  w.spx = cv.take<int16_t>((size_t)3 * w.ld);
  w.lvoff = cv.take<int>(64);
  w.qctr = cv.take<int>(4);
  w.qb = mats;
  w.qm = qm_ext ? qm_ext : mats + t4_count(n) + 1;
  w.scaleG = job.T->scale;
  w.dstate = dstate;
  c.has_cst = job.cst != nullptr;
  c.motif_len = job.motif ? job.motif->len : 0;
  const double mboltz = job.motif ? job.motif->boltz : 1.;
  if (ex.tid() == 0) w.counter[3] = 0;
  stage_hot(ex, H, job.T);
  for (int k = ex.tid(); k < w.ld; k += ex.nthreads()) w.vecp(kVEMLb)[k] = job.T->eMLb[k];
  load_context(ex, c, job.seq, job.cst, job.motif ? job.motif : job.site_motif);   // ends with a barrier; marks the motif sites
  if (!fast_codes(ex, c, w)) w.counter[3] = 1;
  fast_special_hairpins(ex, c, job.T, w);
  ex.sync();
  if (w.counter[3]) return false;
#if defined(ADP_PHASE_PROFILE) && defined(__CUDA_ARCH__)
  if (blockIdx.x == 0 && threadIdx.x == 0) printf("prologue %lld cycles\n", clock64() - tp0_);
#endif

  const double Z = (kOneCstPath || c.has_cst) ? fast_inside<true>(ex, c, *H, job.T, mboltz, w) : fast_inside<false>(ex, c, *H, job.T, mboltz, w);
  if (ex.tid() == 0)
    *job.dG = Z > 0. ? (-log(Z) - n * log(job.T->pf_scale)) * job.T->kT / 1000. : (double)INFINITY;
  if (job.site_ratio && job.site_motif && !job.motif) {
    // Z_holo / Z_apo from the apo matrices (see FastOut): where does the motif occur?
    const int mlen = job.site_motif->len;
    if (ex.tid() == 0) { w.counter[0] = 0; w.counter[1] = 0; }      // every thread is past the fill's last use of the counters
    ex.sync();
    for (int p = 1 + ex.tid(); p + mlen - 1 <= n; p += ex.nthreads())
      if (c.motif_at[p]) { ex.atomic_inc(&w.counter[0]); w.counter[1] = p; }   // the position is used for a single site only
    ex.sync();
    if (ex.tid() == 0) *job.site_ratio = w.counter[0] == 0 || !(Z > 0.) ? 1. : (double)NAN;      // no site, or an empty ensemble: nothing to reweight
    if (w.counter[0] == 1 && Z > 0.) {
      const int at = w.counter[1];
      ex.sync();                      // the counters are reused by the outside pass
      FastOut site;
      site.bppm = nullptr;
      site.n_query = 0; site.qi = nullptr; site.qj = nullptr; site.qprobs = nullptr;
      site.n_dens = 0; site.dens_pos = nullptr; site.dens = nullptr;
      site.site_k = at;
      site.site_l = at + mlen - 1;
      site.site_ph = job.site_ratio;  // receives P(plain hairpin at the site) ...
      if (kOneCstPath || c.has_cst) fast_outside<true>(ex, c, *H, job.T, nullptr, 1., w, site);
      else fast_outside<false>(ex, c, *H, job.T, nullptr, 1., w, site);
      if (ex.tid() == 0) *job.site_ratio = 1. + (job.site_motif->boltz - 1.) * *job.site_ratio;      // ... and is turned into the ratio
      ex.sync();
    }
    return true;                      // site mode replaces the outside pass of the job
  }
  if (!job.outside) return true;
  FastOut out;
  out.bppm = job.bppm;
  out.n_query = job.n_query;
  out.qi = job.qi;
  out.qj = job.qj;
  out.qprobs = job.qprobs;
  out.n_dens = job.n_dens;
  out.dens_pos = job.dens_pos;
  out.dens = job.dens;
  if (Z > 0.) {
    if (kOneCstPath || c.has_cst) fast_outside<true>(ex, c, *H, job.T, job.motif, mboltz, w, out);
    else fast_outside<false>(ex, c, *H, job.T, job.motif, mboltz, w, out);
  } else {
    if (job.bppm)
      for (size_t k = ex.tid(); k < (size_t)n * (n + 1) / 2; k += ex.nthreads()) job.bppm[k] = 0.;
    for (int q = ex.tid(); q < job.n_query; q += ex.nthreads()) job.qprobs[q] = 0.;
    for (int q = ex.tid(); q < job.n_dens; q += ex.nthreads()) job.dens[q] = 0.;
    ex.sync();
  }
  return true;
}

// ---------------------------------------------------------------------------------------------------
struct MfeJob {
  const uint8_t* seq;
  int n;
  const ConstraintDev* cst;
  const MotifDev* motif;
  const IntTables* T;
  int32_t* e_dcal;        // -> minimum free energy
  char* structure;        // NULL or n chars
};

ADP_HD size_t mfe_small_bytes(int n) { return 16 + ctx_bytes(n) + (size_t)4 * (n + 4) * sizeof(int32_t) + (size_t)(4 * n + 16) * 3 * sizeof(int16_t) + (n + 16); }
ADP_HD size_t mfe_matrix_ints(int n) { return (size_t)2 * ((size_t)n * (n + 1) / 2); }

template <class E>
ADP_HD void run_mfe_job(const E& ex, const MfeJob& job, void* small, int32_t* mats) {
  const int n = job.n;
  const size_t ntri = (size_t)n * (n + 1) / 2;
  Carver cv(small);
  SeqCtx c;
  carve_ctx(cv, c, n);
  MfeWork w;
  w.fm1a = cv.take<int32_t>(n + 4);
  w.fm1b = cv.take<int32_t>(n + 4);
  w.f5 = cv.take<int32_t>(n + 4);
  w.ext = cv.take<int32_t>(n + 4);
  w.stack = cv.take<int16_t>((size_t)(4 * n + 16) * 3);
  w.out = cv.take<char>(n + 16);
  w.C = mats;
  w.F = mats + ntri;
  c.has_cst = job.cst != nullptr;
  c.motif_len = job.motif ? job.motif->len : 0;
  const int mdcal = job.motif ? job.motif->contrib_dcal : 0;
  load_context(ex, c, job.seq, job.cst, job.motif);
  const int e = mfe_fill(ex, c, *job.T, mdcal, w);
  if (ex.tid() == 0) *job.e_dcal = e;
  if (!job.structure) return;
  if (ex.warp() == 0) mfe_traceback(ex, c, *job.T, job.motif, mdcal, TriView{w.C, w.F}, w);
  ex.sync();
  for (int p = ex.tid(); p < n; p += ex.nthreads()) job.structure[p] = w.out[p];
}

// The fast path (mfe_fast.cuh) for the same job.  `small`: mfe_fast_small_bytes(n); `mats`: mfe_fast_matrix_ints(n) int32
// (shared or global); `dstate`: see run_pf_job_fast.  Returns false (nothing written) for letters outside ACGU/T.
template <class E>
ADP_HD bool run_mfe_job_fast(const E& ex, const MfeJob& job, void* small, int32_t* mats, int32_t* dstate = nullptr) {
  const int n = job.n;
  Carver cv(small);
  MfeHot* H = cv.take<MfeHot>(1);
  SeqCtx c;
  carve_ctx(cv, c, n);
  MfeFastWork w;
  w.ld = fast_ld(n);
  w.R = mfe_fast_R(n);
  w.vbase = cv.take<int32_t>((size_t)kMfeFastVectors * w.ld);
  w.cn = cv.take<uint8_t>(w.ld);
  w.cp = cv.take<uint8_t>(w.ld);
  w.cells = cv.take<uint16_t>(w.ld);
  w.counter = cv.take<int>(4);
  w.hpw = cv.take<int32_t>(w.ld);
  w.spx = cv.take<int16_t>((size_t)3 * w.ld);
  w.lvoff = cv.take<int>(32);
  MfeWork tb;                              // what the shared traceback needs besides the matrices
  tb.stack = cv.take<int16_t>((size_t)(4 * n + 16) * 3);
  tb.out = cv.take<char>(n + 16);
  tb.C = nullptr; tb.F = nullptr; tb.fm1a = nullptr; tb.fm1b = nullptr; tb.ext = nullptr;
  w.C = mats;
  w.F = mats + t4_count(n) + 2;
  w.dstate = dstate;
  c.has_cst = job.cst != nullptr;
  c.motif_len = job.motif ? job.motif->len : 0;
  const int mdcal = job.motif ? job.motif->contrib_dcal : 0;
  if (ex.tid() == 0) w.counter[2] = 0;
  stage_mfe_hot(ex, H, job.T);
  load_context(ex, c, job.seq, job.cst, job.motif);   // ends with a barrier
  {
    // position codes, shared with the PF fast path
    FastWork codes;
    codes.ld = w.ld;
    codes.cn = w.cn;
    codes.cp = w.cp;
    if (!fast_codes(ex, c, codes)) w.counter[2] = 1;
  }
  mfe_special_hairpins(ex, c, job.T, w);
  ex.sync();
  if (w.counter[2]) return false;

  const int e = c.has_cst ? mfe_fast_fill<true>(ex, c, *H, *job.T, mdcal, w) : mfe_fast_fill<false>(ex, c, *H, *job.T, mdcal, w);
  if (ex.tid() == 0) *job.e_dcal = e;
  if (!job.structure) return true;
  tb.f5 = w.vecp(6);
  if (ex.warp() == 0) mfe_traceback(ex, c, *job.T, job.motif, mdcal, FastMfeView{w.C, w.F, H->FI, w.cn, w.cp, n}, tb);
  ex.sync();
  for (int p = ex.tid(); p < n; p += ex.nthreads()) job.structure[p] = tb.out[p];
  return true;
}

}  // namespace adp
