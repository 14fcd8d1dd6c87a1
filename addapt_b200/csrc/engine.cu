// C-ABI layer of libaddapt_b200.so (see include/addapt_b200.h for the contract and the reference
// interfaces each entry point replaces).  Owns the device tables, the stream, staging buffers and the
// launch plans; all arithmetic of the hot path happens in the kernels (fold_kernels.cu, design_kernels.cu).
// There is no CPU implementation behind these entry points: without a CUDA device they fail.
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/addapt_b200.h"
#include "constraints.h"
#include "design_kernels.h"
#include "energy_model.h"
#include "kernels.h"

using namespace adp;

namespace {

std::string g_create_error;

struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
  cudaError_t ensure(size_t bytes) {
    if (bytes <= cap) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
    size_t want = bytes + bytes / 8 + 256;
    cudaError_t e = cudaMalloc(&p, want);
    if (e == cudaSuccess) cap = want;
    return e;
  }
  void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

bool on_device(const void* ptr) {
  if (!ptr) return false;
  cudaPointerAttributes at;
  if (cudaPointerGetAttributes(&at, ptr) != cudaSuccess) { cudaGetLastError(); return false; }
  return at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged;
}

}  // namespace

struct adp_engine {
  int device = 0;
  cudaStream_t own_stream = nullptr, stream = nullptr;
  bool async = false;
  EnergyModel model;
  PfTables* d_pf = nullptr;
  IntTables* d_int = nullptr;
  DeviceInfo info{};
  std::string err;
  uint64_t launches = 0;
  // staging / scratch, grown on demand
  DevBuf b_seqs, b_ctxseqs, b_dg, b_bppm, b_qi, b_qj, b_qprobs, b_csts, b_motif, b_gmats, b_e, b_structs;
  DevBuf b_score, b_terms, b_prob, b_mc[8], b_mcs[8], b_probe, b_redo, b_gmats2, b_bp, b_bpout;
  bool force_general = false;   // tests: route everything through the general kernels
  MotifDev h_motif{};           // last staged motif record (host copy) and its key
  std::string motif_seq, motif_fold;
  double motif_kd = 0.;
  struct PendingCopy { void* dst; const void* src; size_t bytes; };
  std::vector<PendingCopy> pending;
};

namespace {

#define ADP_CUDA(e, call)                                                                         \
  do {                                                                                            \
    cudaError_t ce_ = (call);                                                                     \
    if (ce_ != cudaSuccess) {                                                                     \
      (e)->err = std::string(#call) + ": " + cudaGetErrorString(ce_);                             \
      (e)->pending.clear(); /* queued D2H copies of a failed call must never run in a later one */ \
      return ADP_ERR_CUDA;                                                                        \
    }                                                                                             \
  } while (0)

// device view of an input: the caller's pointer if it already lives on the device, else a staged copy
template <class T>
int stage_in(adp_engine* e, DevBuf& buf, const T* user, size_t count, const T** out) {
  if (on_device(user)) { *out = user; return ADP_OK; }
  ADP_CUDA(e, buf.ensure(count * sizeof(T)));
  ADP_CUDA(e, cudaMemcpyAsync(buf.p, user, count * sizeof(T), cudaMemcpyHostToDevice, e->stream));
  *out = (const T*)buf.p;
  return ADP_OK;
}

// device destination of an output; host destinations get a device buffer and a copy queued for finish()
template <class T>
int stage_out(adp_engine* e, DevBuf& buf, T* user, size_t count, T** out) {
  if (!user) { *out = nullptr; return ADP_OK; }
  if (on_device(user)) { *out = user; return ADP_OK; }
  ADP_CUDA(e, buf.ensure(count * sizeof(T)));
  *out = (T*)buf.p;
  e->pending.push_back({user, buf.p, count * sizeof(T)});
  return ADP_OK;
}

int finish(adp_engine* e) {
  bool had_host = !e->pending.empty();
  for (auto& pc : e->pending) {
    cudaError_t ce = cudaMemcpyAsync(pc.dst, pc.src, pc.bytes, cudaMemcpyDeviceToHost, e->stream);
    if (ce != cudaSuccess) { e->pending.clear(); e->err = std::string("D2H copy: ") + cudaGetErrorString(ce); return ADP_ERR_CUDA; }
  }
  e->pending.clear();
  if (!e->async || had_host) ADP_CUDA(e, cudaStreamSynchronize(e->stream));
  return ADP_OK;
}

int fail_arg(adp_engine* e, const std::string& msg) {
  e->pending.clear();
  e->err = msg;
  return ADP_ERR_ARG;
}

int check_batch(adp_engine* e, int B, int n, const void* seqs) {
  if (!e) return ADP_ERR_ARG;
  e->pending.clear();      // every batched entry point starts here: nothing queued by an earlier (failed) call survives
  if (B < 0 || !seqs) return fail_arg(e, "null or negative batch");
  if (n < 1 || n > ADP_MAX_LEN) return fail_arg(e, "sequence length must be 1.." + std::to_string(ADP_MAX_LEN));
  return ADP_OK;
}

// uploads a motif record (or none) and returns its device pointer.  The record of the last call is cached, so a
// loop that keeps folding with the same aptamer costs neither a table evaluation nor a host synchronisation.
int stage_motif(adp_engine* e, const adp_motif* motif, const MotifDev** d_motif, MotifDev* host_copy = nullptr) {
  *d_motif = nullptr;
  if (!motif || !motif->seq) return ADP_OK;
  if (!motif->fold) return fail_arg(e, "aptamer without a fold");
  const bool cached = e->b_motif.p && e->motif_seq == motif->seq && e->motif_fold == motif->fold && e->motif_kd == motif->kd_uM;
  if (!cached) {
    try {
      e->h_motif = e->model.make_motif(motif->seq, motif->fold, e->model.kT_kcal() * std::log(motif->kd_uM / 1e6));
    } catch (std::string& s) { return fail_arg(e, s); }
    ADP_CUDA(e, e->b_motif.ensure(sizeof(MotifDev)));
    ADP_CUDA(e, cudaStreamSynchronize(e->stream));   // kernels of earlier calls may still read the old record
    ADP_CUDA(e, cudaMemcpyAsync(e->b_motif.p, &e->h_motif, sizeof(MotifDev), cudaMemcpyHostToDevice, e->stream));
    e->motif_seq = motif->seq;
    e->motif_fold = motif->fold;
    e->motif_kd = motif->kd_uM;
  }
  *d_motif = (const MotifDev*)e->b_motif.p;
  if (host_copy) *host_copy = e->h_motif;
  return ADP_OK;
}

int stage_constraints(adp_engine* e, const std::vector<std::string>& csts, int n, const ConstraintDev** d_csts) {
  *d_csts = nullptr;
  if (csts.empty()) return ADP_OK;
  std::vector<ConstraintDev> host(csts.size());
  try {
    for (size_t k = 0; k < csts.size(); k++) build_constraint(csts[k], n, &host[k]);
  } catch (std::string& s) { return fail_arg(e, s); }
  ADP_CUDA(e, e->b_csts.ensure(host.size() * sizeof(ConstraintDev)));
  ADP_CUDA(e, cudaMemcpyAsync(e->b_csts.p, host.data(), host.size() * sizeof(ConstraintDev), cudaMemcpyHostToDevice, e->stream));
  ADP_CUDA(e, cudaStreamSynchronize(e->stream));
  *d_csts = (const ConstraintDev*)e->b_csts.p;
  return ADP_OK;
}

std::string default_par_path() {
  if (const char* env = std::getenv("ADDAPT_B200_PAR")) return env;
#ifdef ADP_DEFAULT_PAR
  return ADP_DEFAULT_PAR;
#else
  return "params/rna_turner2004_reconstructed.par";
#endif
}

}  // namespace

extern "C" {

int adp_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

int adp_engine_create(int device, const char* par_file, adp_engine** out) {
  if (!out) return ADP_ERR_ARG;
  *out = nullptr;
  adp_engine* e = new adp_engine();
  try {
    e->model.load(par_file ? std::string(par_file) : default_par_path());
  } catch (std::string& s) {
    g_create_error = s;
    delete e;
    return ADP_ERR_PARAM;
  }
  if (e->model.provenance() == "reconstructed" && !std::getenv("ADDAPT_B200_QUIET")) {
    static bool warned = false;
    if (!warned) {
      warned = true;
      std::fprintf(stderr,
                   "addapt_b200: WARNING: energy parameters come from a RECONSTRUCTED Turner-2004 file (%s):\n"
                   "addapt_b200: WARNING: int11 / int21 / int22 are rule-generated, so energies, probabilities and scores differ from\n"
                   "addapt_b200: WARNING: ViennaRNA's (rhf(6) apo ensemble: -29.75 vs -29.58 kcal/mol).  Pass a genuine rna_turner2004.par as\n"
                   "addapt_b200: WARNING: par_file or in $ADDAPT_B200_PAR for reference-exact numbers (adp_param_provenance() reports which).\n",
                   e->model.path().c_str());
    }
  }
  auto bail = [&](const char* what, cudaError_t ce) {
    g_create_error = std::string(what) + ": " + cudaGetErrorString(ce) + " (this engine has no CPU fallback)";
    delete e;
    return ADP_ERR_CUDA;
  };
  cudaError_t ce;
  int count = 0;
  if ((ce = cudaGetDeviceCount(&count)) != cudaSuccess || count == 0) {
    cudaGetLastError();
    return bail("no CUDA device", ce == cudaSuccess ? cudaErrorNoDevice : ce);
  }
  if (device < 0 || device >= count) { g_create_error = "device index out of range"; delete e; return ADP_ERR_ARG; }
  e->device = device;
  if ((ce = cudaSetDevice(device)) != cudaSuccess) return bail("cudaSetDevice", ce);
  cudaDeviceProp prop;
  if ((ce = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) return bail("cudaGetDeviceProperties", ce);
  e->info.sm_count = prop.multiProcessorCount;
  e->info.smem_optin = prop.sharedMemPerBlockOptin;
  e->info.smem_per_sm = prop.sharedMemPerMultiprocessor;
  if ((ce = cudaStreamCreateWithFlags(&e->own_stream, cudaStreamNonBlocking)) != cudaSuccess) return bail("cudaStreamCreate", ce);
  e->stream = e->own_stream;
  if ((ce = cudaMalloc(&e->d_pf, sizeof(PfTables))) != cudaSuccess) return bail("cudaMalloc", ce);
  if ((ce = cudaMalloc(&e->d_int, sizeof(IntTables))) != cudaSuccess) return bail("cudaMalloc", ce);
  if ((ce = cudaMemcpy(e->d_pf, &e->model.pf(), sizeof(PfTables), cudaMemcpyHostToDevice)) != cudaSuccess) return bail("cudaMemcpy", ce);
  if ((ce = cudaMemcpy(e->d_int, &e->model.mfe(), sizeof(IntTables), cudaMemcpyHostToDevice)) != cudaSuccess) return bail("cudaMemcpy", ce);
  *out = e;
  return ADP_OK;
}

void adp_engine_destroy(adp_engine* e) {
  if (!e) return;
  cudaSetDevice(e->device);
  if (e->own_stream) { cudaStreamSynchronize(e->own_stream); }
  DevBuf* bufs[] = {&e->b_seqs, &e->b_ctxseqs, &e->b_dg, &e->b_bppm, &e->b_qi, &e->b_qj, &e->b_qprobs, &e->b_csts, &e->b_motif,
                    &e->b_gmats, &e->b_e, &e->b_structs, &e->b_score, &e->b_terms, &e->b_prob, &e->b_probe, &e->b_redo, &e->b_gmats2, &e->b_bp, &e->b_bpout};
  for (DevBuf* b : bufs) b->release();
  for (auto& b : e->b_mc) b.release();
  for (auto& b : e->b_mcs) b.release();
  if (e->d_pf) cudaFree(e->d_pf);
  if (e->d_int) cudaFree(e->d_int);
  if (e->own_stream) cudaStreamDestroy(e->own_stream);
  delete e;
}

const char* adp_last_error(const adp_engine* e) { return e ? e->err.c_str() : g_create_error.c_str(); }

int adp_engine_set_stream(adp_engine* e, void* s) {
  if (!e) return ADP_ERR_ARG;
  e->stream = s ? (cudaStream_t)s : e->own_stream;
  return ADP_OK;
}
int adp_engine_set_async(adp_engine* e, int a) { if (!e) return ADP_ERR_ARG; e->async = a != 0; return ADP_OK; }
int adp_engine_set_general_only(adp_engine* e, int on) { if (!e) return ADP_ERR_ARG; e->force_general = on != 0; return ADP_OK; }
int adp_engine_sync(adp_engine* e) {
  if (!e) return ADP_ERR_ARG;
  ADP_CUDA(e, cudaStreamSynchronize(e->stream));
  return ADP_OK;
}
double adp_kT_kcal(const adp_engine* e) { return e ? e->model.kT_kcal() : 0.; }
const char* adp_param_provenance(const adp_engine* e) { return e ? e->model.provenance().c_str() : ""; }
const char* adp_param_file(const adp_engine* e) { return e ? e->model.path().c_str() : ""; }
uint64_t adp_launch_count(const adp_engine* e) { return e ? e->launches : 0; }

int adp_eval_structure(adp_engine* e, const char* seq, const char* structure, int32_t* e_dcal) {
  if (!e || !seq || !structure || !e_dcal) return ADP_ERR_ARG;
  try { *e_dcal = e->model.eval_structure(seq, structure); } catch (std::string& s) { return fail_arg(e, s); }
  return ADP_OK;
}

int adp_motif_contribution(adp_engine* e, const adp_motif* m, int32_t* c) {
  if (!e || !m || !m->seq || !m->fold || !c) return ADP_ERR_ARG;
  try { *c = e->model.make_motif(m->seq, m->fold, e->model.kT_kcal() * std::log(m->kd_uM / 1e6)).contrib_dcal; }
  catch (std::string& s) { return fail_arg(e, s); }
  return ADP_OK;
}

int adp_mfe_batch(adp_engine* e, int B, int n, const uint8_t* seqs, const char* cst, const adp_motif* motif,
                  int32_t* e_dcal, char* structs) {
  int rc = check_batch(e, B, n, seqs);
  if (rc) return rc;
  if (!e_dcal) return fail_arg(e, "e_dcal is NULL");
  if (B == 0) return ADP_OK;
  ADP_CUDA(e, cudaSetDevice(e->device));
  MfeBatchArgs a{};
  a.B = B; a.n = n; a.T = e->d_int;
  if ((rc = stage_motif(e, motif, &a.motif))) return rc;
  std::vector<std::string> cs;
  if (cst) cs.emplace_back(cst);
  if ((rc = stage_constraints(e, cs, n, &a.cst))) return rc;
  if ((rc = stage_in(e, e->b_seqs, seqs, (size_t)B * n, &a.seqs))) return rc;
  if ((rc = stage_out(e, e->b_e, e_dcal, (size_t)B, &a.e_dcal))) return rc;
  if ((rc = stage_out(e, e->b_structs, structs, (size_t)B * n, &a.structs))) return rc;
  LaunchPlan pfast = plan_mfe_fast(e->info, n, B);
  if (pfast.threads && !e->force_general) {
    // fast kernel first; it flags the sequences it declines (letters outside ACGU/T), which the general kernel then folds
    ADP_CUDA(e, e->b_redo.ensure((size_t)(B + 1) * sizeof(int)));
    ADP_CUDA(e, cudaMemsetAsync(e->b_redo.p, 0, (size_t)(B + 1) * sizeof(int), e->stream));
    a.redo = (int*)e->b_redo.p;
    if (pfast.gmat_stride) {
      ADP_CUDA(e, e->b_gmats.ensure((size_t)pfast.grid * pfast.gmat_stride * sizeof(int32_t)));
      a.gmats = (int32_t*)e->b_gmats.p;
      a.gmat_stride = pfast.gmat_stride;
    }
    ADP_CUDA(e, launch_mfe_fast(a, pfast, &e->model.mfe(), e->stream));
    e->launches++;
    a.only_redo = 1;
  }
  LaunchPlan p = plan_mfe(e->info, n, B);
  if (!p.mats_in_smem) {
    ADP_CUDA(e, e->b_gmats2.ensure((size_t)p.grid * p.gmat_stride * sizeof(int32_t)));
    a.gmats = (int32_t*)e->b_gmats2.p;
    a.gmat_stride = p.gmat_stride;
  }
  ADP_CUDA(e, launch_mfe(a, p, e->stream));
  e->launches++;
  return finish(e);
}

static int run_pf(adp_engine* e, int B, int n, const uint8_t* d_seqs, int V, const VariantDev* variants,
                  const ConstraintDev* d_csts, const MotifDev* d_motif, double* d_dG, double* d_bppm,
                  int n_query, const int32_t* d_qi, const int32_t* d_qj, double* d_qprobs) {
  PfBatchArgs a{};
  a.seqs = d_seqs; a.B = B; a.n = n; a.V = V;
  bool any_outside = false;
  for (int v = 0; v < V; v++) { a.variants[v] = variants[v]; any_outside |= variants[v].outside != 0; }
  a.csts = d_csts; a.motif = d_motif; a.T = e->d_pf;
  a.dG = d_dG; a.bppm = d_bppm; a.n_query = n_query; a.qi = d_qi; a.qj = d_qj; a.qprobs = d_qprobs;
  const int jobs = B * V;
  LaunchPlan pfast = plan_pf_fast(e->info, n, jobs);
  if (pfast.threads && !e->force_general) {
    // fast kernel first; it flags the jobs it declines (letters outside ACGU/T), which the general kernel then folds
    ADP_CUDA(e, e->b_redo.ensure((size_t)(jobs + 1) * sizeof(int)));
    ADP_CUDA(e, cudaMemsetAsync(e->b_redo.p, 0, (size_t)(jobs + 1) * sizeof(int), e->stream));
    a.redo = (int*)e->b_redo.p;
    if (pfast.gmat_stride) {
      ADP_CUDA(e, e->b_gmats.ensure((size_t)pfast.grid * pfast.gmat_stride * sizeof(double)));
      a.gmats = (double*)e->b_gmats.p;
      a.gmat_stride = pfast.gmat_stride;
    }
    ADP_CUDA(e, launch_pf_fast(a, pfast, &e->model.pf(), e->stream));
    e->launches++;
    a.only_redo = 1;
  }
  LaunchPlan p = plan_pf(e->info, n, jobs, any_outside);
  if (!p.mats_in_smem) {
    ADP_CUDA(e, e->b_gmats2.ensure((size_t)p.grid * p.gmat_stride * sizeof(double)));
    a.gmats = (double*)e->b_gmats2.p;
    a.gmat_stride = p.gmat_stride;
  }
  ADP_CUDA(e, launch_pf(a, p, e->stream));
  e->launches++;
  return ADP_OK;
}

int adp_pf_batch(adp_engine* e, int B, int n, const uint8_t* seqs, const char* cst, const adp_motif* motif,
                 double* dG, double* bppm) {
  int rc = check_batch(e, B, n, seqs);
  if (rc) return rc;
  if (!dG) return fail_arg(e, "dG_kcal is NULL");
  if (B == 0) return ADP_OK;
  ADP_CUDA(e, cudaSetDevice(e->device));
  const MotifDev* d_motif;
  const ConstraintDev* d_csts;
  if ((rc = stage_motif(e, motif, &d_motif))) return rc;
  std::vector<std::string> cs;
  if (cst) cs.emplace_back(cst);
  if ((rc = stage_constraints(e, cs, n, &d_csts))) return rc;
  const uint8_t* d_seqs;
  double *d_dG, *d_bppm;
  const size_t ntri = (size_t)n * (n + 1) / 2;
  if ((rc = stage_in(e, e->b_seqs, seqs, (size_t)B * n, &d_seqs))) return rc;
  if ((rc = stage_out(e, e->b_dg, dG, (size_t)B, &d_dG))) return rc;
  if ((rc = stage_out(e, e->b_bppm, bppm, (size_t)B * ntri, &d_bppm))) return rc;
  VariantDev var{cst ? 0 : -1, d_motif ? 1 : 0, bppm ? 1 : 0, 0};
  if ((rc = run_pf(e, B, n, d_seqs, 1, &var, d_csts, d_motif, d_dG, d_bppm, 0, nullptr, nullptr, nullptr))) { e->pending.clear(); return rc; }
  return finish(e);
}

int adp_fold_eval_batch(adp_engine* e, int B, int n, const uint8_t* seqs, const adp_motif* motif,
                        int n_query, const int32_t* query_i, const int32_t* query_j,
                        double* dG, double* probs, double* bppm_full) {
  int rc = check_batch(e, B, n, seqs);
  if (rc) return rc;
  if (!dG) return fail_arg(e, "dG is NULL");
  if (n_query < 0 || (n_query > 0 && (!query_i || !query_j || !probs))) return fail_arg(e, "bad pair-probability query");
  if (B == 0) return ADP_OK;
  ADP_CUDA(e, cudaSetDevice(e->device));
  const MotifDev* d_motif;
  if ((rc = stage_motif(e, motif, &d_motif))) return rc;
  const uint8_t* d_seqs;
  const int32_t *d_qi = nullptr, *d_qj = nullptr;
  double *d_dG, *d_probs, *d_bppm;
  const size_t ntri = (size_t)n * (n + 1) / 2;
  if ((rc = stage_in(e, e->b_seqs, seqs, (size_t)B * n, &d_seqs))) return rc;
  if (n_query) {
    if ((rc = stage_in(e, e->b_qi, query_i, (size_t)n_query, &d_qi))) return rc;
    if ((rc = stage_in(e, e->b_qj, query_j, (size_t)n_query, &d_qj))) return rc;
  }
  if ((rc = stage_out(e, e->b_dg, dG, (size_t)B * 2, &d_dG))) return rc;
  if ((rc = stage_out(e, e->b_qprobs, n_query ? probs : nullptr, (size_t)B * 2 * n_query, &d_probs))) return rc;
  if ((rc = stage_out(e, e->b_bppm, bppm_full, (size_t)B * 2 * ntri, &d_bppm))) return rc;
  VariantDev vars[2] = {{-1, 0, 1, 0}, {-1, d_motif ? 1 : 0, 1, 0}};
  if ((rc = run_pf(e, B, n, d_seqs, 2, vars, nullptr, d_motif, d_dG, d_bppm, n_query, d_qi, d_qj, d_probs))) { e->pending.clear(); return rc; }
  return finish(e);
}

int adp_measure_dfma_rate(adp_engine* e, double* gfma_per_s) {
  if (!e || !gfma_per_s) return ADP_ERR_ARG;
  ADP_CUDA(e, cudaSetDevice(e->device));
  const int threads = 512, grid = e->info.sm_count * 4, iters = 1 << 16;
  ADP_CUDA(e, e->b_probe.ensure((size_t)threads * grid * sizeof(double)));
  cudaEvent_t t0, t1;
  ADP_CUDA(e, cudaEventCreate(&t0));
  ADP_CUDA(e, cudaEventCreate(&t1));
  double best = 0.;
  for (int rep = 0; rep < 4; rep++) {
    ADP_CUDA(e, cudaEventRecord(t0, e->stream));
    ADP_CUDA(e, launch_dfma_probe((double*)e->b_probe.p, iters, grid, threads, e->stream));
    ADP_CUDA(e, cudaEventRecord(t1, e->stream));
    ADP_CUDA(e, cudaEventSynchronize(t1));
    float ms = 0.f;
    ADP_CUDA(e, cudaEventElapsedTime(&ms, t0, t1));
    e->launches++;
    const double fma = (double)threads * grid * (double)iters * 8.;
    const double rate = fma / (ms * 1e-3) / 1e9;
    if (rep > 0 && rate > best) best = rate;
  }
  cudaEventDestroy(t0);
  cudaEventDestroy(t1);
  *gfma_per_s = best;
  return ADP_OK;
}

}  // extern "C"

#include "engine_design.inc"
