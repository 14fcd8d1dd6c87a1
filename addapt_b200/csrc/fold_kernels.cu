// sm_100a kernels for batched folds: one construct per CTA, persistent CTAs striding over the job list.
// The recursions live in pf_core.cuh / mfe_core.cuh; this file binds them to a CTA, shared memory and the
// batch layout.  No tensor cores (the recursions are not dense contractions), no library calls.
#include <cstdlib>

#include <vector>

#include "fold_jobs.cuh"
#include "kernels.h"

namespace adp {


extern __shared__ __align__(16) unsigned char adp_smem[];

template <int G, bool kSmemMats>
__global__ void __launch_bounds__(512) pf_batch_kernel(const PfBatchArgs a) {
  CtaExec<G> ex;
  const int n = a.n;
  const size_t ntri = (size_t)n * (n + 1) / 2;
  const int jobs = a.B * a.V;
  void* small = adp_smem;
  double* mats;
  if (kSmemMats) {
    size_t off = (pf_small_bytes(n) + 15) & ~(size_t)15;
    mats = reinterpret_cast<double*>(adp_smem + off);
  } else {
    mats = a.gmats + (size_t)blockIdx.x * a.gmat_stride;
  }
  if (a.only_redo && a.redo[jobs] == 0) return;
  for (int jid = blockIdx.x; jid < jobs; jid += gridDim.x) {
    if (a.only_redo && !a.redo[jid]) continue;
    const int b = jid / a.V, v = jid - b * a.V;
    const VariantDev var = a.variants[v];
    PfJob job;
    job.seq = a.seqs + (size_t)b * n;
    job.n = n;
    job.cst = var.cst_id >= 0 ? a.csts + var.cst_id : nullptr;
    job.motif = var.use_motif ? a.motif : nullptr;
    job.T = a.T;
    job.outside = var.outside != 0;
    job.dG = a.dG + jid;
    job.bppm = (a.bppm && job.outside) ? a.bppm + (size_t)jid * ntri : nullptr;
    job.n_query = job.outside ? a.n_query : 0;
    job.qi = a.qi;
    job.qj = a.qj;
    job.qprobs = a.qprobs ? a.qprobs + (size_t)jid * a.n_query : nullptr;
    run_pf_job(ex, job, small, mats);
    __syncthreads();   // the next job reuses the working set
  }
}

template <int G, bool kSmemMats>
__global__ void __launch_bounds__(512) mfe_batch_kernel(const MfeBatchArgs a) {
  CtaExec<G> ex;
  const int n = a.n;
  void* small = adp_smem;
  int32_t* mats;
  if (kSmemMats) {
    size_t off = (mfe_small_bytes(n) + 15) & ~(size_t)15;
    mats = reinterpret_cast<int32_t*>(adp_smem + off);
  } else {
    mats = a.gmats + (size_t)blockIdx.x * a.gmat_stride;
  }
  if (a.only_redo && a.redo[a.B] == 0) return;
  for (int b = blockIdx.x; b < a.B; b += gridDim.x) {
    if (a.only_redo && !a.redo[b]) continue;
    MfeJob job;
    job.seq = a.seqs + (size_t)b * n;
    job.n = n;
    job.cst = a.cst;
    job.motif = a.motif;
    job.T = a.T;
    job.e_dcal = a.e_dcal + b;
    job.structure = a.structs ? a.structs + (size_t)b * n : nullptr;
    run_mfe_job(ex, job, small, mats);
    __syncthreads();
  }
}

__device__ __forceinline__ PfJob make_pf_job(const PfBatchArgs& a, int jid) {
  const int n = a.n;
  const size_t ntri = (size_t)n * (n + 1) / 2;
  const int b = jid / a.V, v = jid - b * a.V;
  const VariantDev var = a.variants[v];
  PfJob job;
  job.seq = a.seqs + (size_t)b * n;
  job.n = n;
  job.cst = var.cst_id >= 0 ? a.csts + var.cst_id : nullptr;
  job.motif = var.use_motif ? a.motif : nullptr;
  job.T = a.T;
  job.outside = var.outside != 0;
  job.dG = a.dG + jid;
  job.bppm = (a.bppm && job.outside) ? a.bppm + (size_t)jid * ntri : nullptr;
  job.n_query = job.outside ? a.n_query : 0;
  job.qi = a.qi;
  job.qj = a.qj;
  job.qprobs = a.qprobs ? a.qprobs + (size_t)jid * a.n_query : nullptr;
  return job;
}

// Fast path: one construct per CTA, level sums of the generic interior loops in registers (pf_fast.cuh).
// kThreads = 256 serves n <= 128 with two CTAs per SM at 100 nt; 512 serves longer constructs.
// kQmGlobal: Qb stays in shared memory, Qm moves to the CTA's slab of (L2-resident) global scratch: constructs of 101-128 nt, whose
// two matrices leave room for one CTA per SM only, keep two CTAs per SM that way (plan_pf_fast).
template <int kThreads, bool kSmemMats, bool kQmGlobal>
__global__ void __launch_bounds__(kThreads, kThreads == 256 ? 2 : 1) pf_fast_kernel(const PfBatchArgs a) {
  CtaExec<1> ex;
  const int n = a.n;
  const int jobs = a.B * a.V;
  void* small = adp_smem;
  double* mats;
  double* qm_ext = nullptr;
  size_t gused = 0;                  // doubles of this CTA's global slab already handed out
  double* slab = a.gmats + (size_t)blockIdx.x * a.gmat_stride;
  if (kSmemMats) {
    size_t off = (fast_small_bytes(n) + 15) & ~(size_t)15;
    mats = reinterpret_cast<double*>(adp_smem + off);
    if (kQmGlobal) { qm_ext = slab; gused = t4_count(n) + 2; }
  } else {
    mats = slab;
    gused = fast_matrix_doubles(n);
  }
  // constructs longer than threads/2 have more anti-diagonals than threads: their level sums live in global scratch too
  double* dstate = n > fast_max_len(kThreads) ? slab + gused : nullptr;
  // Large batches are walked design-major: a CTA folds all variants of a sequence (apo, holo) back to back.  Job-major order gave
  // every CTA only apo or only holo jobs (the grid size is even), i.e. a static imbalance between the CTAs, and the second variant
  // now finds the sequence's tables warm; measured +4 % fold-evals/s.  Small batches stay job-major (every job on its own CTA).
  const bool by_design = a.B >= (int)gridDim.x;
  const int units = by_design ? a.B : jobs, per = by_design ? a.V : 1;
  for (int u = blockIdx.x; u < units; u += gridDim.x)
    for (int v = 0; v < per; v++) {
      const int jid = u * per + v;
      const PfJob job = make_pf_job(a, jid);
      const bool done = run_pf_job_fast(ex, job, small, mats, dstate, qm_ext);
      if (!done && threadIdx.x == 0) { a.redo[jid] = 1; atomicAdd(a.redo + jobs, 1); }
      __syncthreads();   // the next job reuses the working set
    }
}

// Fast MFE path (mfe_fast.cuh): int32 matrices are half the size of the PF ones, so up to four constructs share an SM.
template <int kThreads, bool kSmemMats>
__global__ void __launch_bounds__(kThreads, kThreads == 256 ? 4 : 1) mfe_fast_kernel(const MfeBatchArgs a) {
  CtaExec<1> ex;
  const int n = a.n;
  void* small = adp_smem;
  int32_t* mats;
  if (kSmemMats) {
    size_t off = (mfe_fast_small_bytes(n) + 15) & ~(size_t)15;
    mats = reinterpret_cast<int32_t*>(adp_smem + off);
  } else {
    mats = a.gmats + (size_t)blockIdx.x * a.gmat_stride;
  }
  int32_t* dstate = nullptr;
  if (n > fast_max_len(kThreads))
    dstate = a.gmats + (size_t)blockIdx.x * a.gmat_stride + (kSmemMats ? 0 : mfe_fast_matrix_ints(n));
  for (int b = blockIdx.x; b < a.B; b += gridDim.x) {
    MfeJob job;
    job.seq = a.seqs + (size_t)b * n;
    job.n = n;
    job.cst = a.cst;
    job.motif = a.motif;
    job.T = a.T;
    job.e_dcal = a.e_dcal + b;
    job.structure = a.structs ? a.structs + (size_t)b * n : nullptr;
    const bool done = run_mfe_job_fast(ex, job, small, mats, dstate);
    if (!done && threadIdx.x == 0) { a.redo[b] = 1; atomicAdd(a.redo + a.B, 1); }
    __syncthreads();
  }
}

// Register-resident FP64 FMA loop: the measured denominator of the SM-issue roofline.
__global__ void dfma_probe_kernel(double* out, int iters) {
  double a0 = 1.0 + threadIdx.x * 1e-9, a1 = a0 + 1e-9, a2 = a0 + 2e-9, a3 = a0 + 3e-9;
  double a4 = a0 + 4e-9, a5 = a0 + 5e-9, a6 = a0 + 6e-9, a7 = a0 + 7e-9;
  const double m = 0.999999999, c = 1e-12;
  for (int i = 0; i < iters; i++) {
    a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
    a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
  }
  out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
}

// ---------------------------------------------------------------------------------------------------
static void pick_shape(int n, int* threads, int* group) {
  // lanes per cell x cell slots: keep one pass per column where the thread budget allows it
  if (n <= 68) { *threads = 256; *group = 4; }
  else if (n <= 132) { *threads = 256; *group = 2; }
  else if (n <= 260) { *threads = 512; *group = 2; }
  else { *threads = 512; *group = 1; }
}

static LaunchPlan plan_common(const DeviceInfo& d, int n, int jobs, size_t small, size_t mats_bytes) {
  LaunchPlan p{};
  pick_shape(n, &p.threads, &p.group);
  const size_t small_al = (small + 15) & ~(size_t)15;
  const size_t both = small_al + mats_bytes;
  p.mats_in_smem = both <= d.smem_optin;
  p.smem_bytes = p.mats_in_smem ? both : small_al;
  size_t per_sm = d.smem_per_sm / (p.smem_bytes + 1024);   // 1 KB reserved per resident CTA
  size_t by_threads = 2048 / p.threads;
  size_t occ = per_sm < by_threads ? per_sm : by_threads;
  if (occ < 1) occ = 1;
  long grid = (long)d.sm_count * (long)occ;
  if (grid > jobs) grid = jobs;
  if (grid < 1) grid = 1;
  p.grid = (int)grid;
  return p;
}

LaunchPlan plan_pf(const DeviceInfo& d, int n, int jobs, bool any_outside) {
  const size_t nd = pf_matrix_doubles(n, any_outside);
  LaunchPlan p = plan_common(d, n, jobs, pf_small_bytes(n), nd * sizeof(double));
  p.gmat_stride = p.mats_in_smem ? 0 : nd;
  return p;
}

LaunchPlan plan_mfe(const DeviceInfo& d, int n, int jobs) {
  const size_t ni = mfe_matrix_ints(n);
  LaunchPlan p = plan_common(d, n, jobs, mfe_small_bytes(n), ni * sizeof(int32_t));
  p.gmat_stride = p.mats_in_smem ? 0 : ni;
  return p;
}

template <class K>
static cudaError_t set_smem(K kernel, size_t bytes) {
  return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
}

#define ADP_LAUNCH(KERNEL, ARGS)                                                        \
  do {                                                                                  \
    cudaError_t e_ = set_smem(KERNEL, p.smem_bytes);                                    \
    if (e_ != cudaSuccess) return e_;                                                   \
    KERNEL<<<p.grid, p.threads, p.smem_bytes, stream>>>(ARGS);                          \
    return cudaGetLastError();                                                          \
  } while (0)

cudaError_t launch_pf(const PfBatchArgs& a, const LaunchPlan& p, cudaStream_t stream) {
  if (p.mats_in_smem) {
    switch (p.group) {
      case 4: ADP_LAUNCH((pf_batch_kernel<4, true>), a);
      case 2: ADP_LAUNCH((pf_batch_kernel<2, true>), a);
      default: ADP_LAUNCH((pf_batch_kernel<1, true>), a);
    }
  }
  switch (p.group) {
    case 4: ADP_LAUNCH((pf_batch_kernel<4, false>), a);
    case 2: ADP_LAUNCH((pf_batch_kernel<2, false>), a);
    default: ADP_LAUNCH((pf_batch_kernel<1, false>), a);
  }
}

LaunchPlan plan_pf_fast(const DeviceInfo& d, int n, int jobs) {
  LaunchPlan p{};
  // 256 threads keep the level sums of n <= 128 in registers; 512 threads serve n <= 256 the same way and longer constructs
  // (up to kFastMaxN) with the sums in global scratch
  p.threads = n <= fast_max_len(256) ? 256 : n <= kFastMaxN ? 512 : 0;
  if (!p.threads) return p;
  p.group = 1;
  const size_t small_al = (fast_small_bytes(n) + 15) & ~(size_t)15;
  const size_t mats_bytes = fast_matrix_doubles(n) * sizeof(double);
  const size_t qb_bytes = (t4_count(n) + 2) * sizeof(double);
  p.mats_in_smem = small_al + mats_bytes <= d.smem_optin;
  p.smem_bytes = p.mats_in_smem ? small_al + mats_bytes : small_al;
  size_t gdoubles = p.mats_in_smem ? 0 : fast_matrix_doubles(n);
  size_t by_threads = p.threads == 256 ? 2 : 1;            // __launch_bounds__ of pf_fast_kernel
  // Occupancy cliff above ~100 nt: two FP64 triangles + the vectors exceed half an SM's shared memory, i.e. ONE resident CTA (8 warps)
  // per SM.  With Qm in the CTA's slab of L2-resident global scratch (group = 3 selects that kernel variant) the 256-thread kernel
  // keeps two CTAs per SM up to 128 nt.  ADDAPT_B200_QM_GLOBAL=0 switches it off (measurements).
  const char* qmg = std::getenv("ADDAPT_B200_QM_GLOBAL");
  const bool want_qmg = qmg ? std::atoi(qmg) != 0 : true;
  if (want_qmg && p.threads == 256 && p.mats_in_smem && d.smem_per_sm / (p.smem_bytes + 1024) < 2 &&
      d.smem_per_sm / (small_al + qb_bytes + 1024) >= 2) {
    p.group = 3;
    p.smem_bytes = small_al + qb_bytes;
    gdoubles = t4_count(n) + 2;
  }
  p.gmat_stride = gdoubles + (n > fast_max_len(p.threads) ? fast_state_doubles(n) : 0);
  size_t per_sm = d.smem_per_sm / (p.smem_bytes + 1024);   // 1 KB reserved per resident CTA
  size_t occ = per_sm < by_threads ? per_sm : by_threads;
  if (const char* cap = std::getenv("ADDAPT_B200_MAX_CTAS_PER_SM")) {   // experiments only
    const size_t c = (size_t)std::atoi(cap);
    if (c >= 1 && c < occ) occ = c;
  }
  if (occ < 1) occ = 1;
  long grid = (long)d.sm_count * (long)occ;
  if (grid > jobs) grid = jobs;
  if (grid < 1) grid = 1;
  p.grid = (int)grid;
  return p;
}

// code-indexed corner table (pf_fast.cuh, build_corner_table): the table of the current device, built and uploaded on first use
const double* device_corner_table(const PfTables* host_tables) {
  static double* tables[64] = {nullptr};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
  if (!tables[dev]) {
    std::vector<double> h(kCornerEntries);
    build_corner_table(host_tables, h.data());
    double* d = nullptr;
    if (cudaMalloc(&d, kCornerEntries * sizeof(double)) != cudaSuccess) return nullptr;
    if (cudaMemcpy(d, h.data(), kCornerEntries * sizeof(double), cudaMemcpyHostToDevice) != cudaSuccess) return nullptr;
    tables[dev] = d;
  }
  return tables[dev];
}

cudaError_t launch_pf_fast(const PfBatchArgs& a, const LaunchPlan& p, const PfTables* host_tables, cudaStream_t stream) {
  FastConst fc;
  fill_fast_const(&fc, host_tables);
  fc.corner = device_corner_table(host_tables);
  if (!fc.corner) return cudaErrorMemoryAllocation;
  cudaError_t ce = cudaMemcpyToSymbolAsync(c_fast_const, &fc, sizeof fc, 0, cudaMemcpyHostToDevice, stream);
  if (ce != cudaSuccess) return ce;
  if (p.threads == 256) {
    if (p.group == 3) ADP_LAUNCH((pf_fast_kernel<256, true, true>), a);
    if (p.mats_in_smem) ADP_LAUNCH((pf_fast_kernel<256, true, false>), a);
    ADP_LAUNCH((pf_fast_kernel<256, false, false>), a);
  }
  if (p.mats_in_smem) ADP_LAUNCH((pf_fast_kernel<512, true, false>), a);
  ADP_LAUNCH((pf_fast_kernel<512, false, false>), a);
}

LaunchPlan plan_mfe_fast(const DeviceInfo& d, int n, int jobs) {
  LaunchPlan p{};
  p.threads = n <= fast_max_len(256) ? 256 : n <= kFastMaxN ? 512 : 0;
  if (!p.threads) return p;
  p.group = 1;
  const size_t small_al = (mfe_fast_small_bytes(n) + 15) & ~(size_t)15;
  const size_t mats_bytes = mfe_fast_matrix_ints(n) * sizeof(int32_t);
  p.mats_in_smem = small_al + mats_bytes <= d.smem_optin;
  p.smem_bytes = p.mats_in_smem ? small_al + mats_bytes : small_al;
  p.gmat_stride = (p.mats_in_smem ? 0 : mfe_fast_matrix_ints(n)) + (n > fast_max_len(p.threads) ? fast_state_doubles(n) : 0);
  size_t per_sm = d.smem_per_sm / (p.smem_bytes + 1024);
  size_t by_threads = p.threads == 256 ? 4 : 1;            // __launch_bounds__ of mfe_fast_kernel
  size_t occ = per_sm < by_threads ? per_sm : by_threads;
  if (occ < 1) occ = 1;
  long grid = (long)d.sm_count * (long)occ;
  if (grid > jobs) grid = jobs;
  if (grid < 1) grid = 1;
  p.grid = (int)grid;
  return p;
}

static const int32_t* device_corner_table_i(const IntTables* host_tables) {
  static int32_t* tables[64] = {nullptr};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
  if (!tables[dev]) {
    std::vector<int32_t> h(kCornerEntriesI);
    build_corner_table_i(host_tables, h.data());
    int32_t* d = nullptr;
    if (cudaMalloc(&d, kCornerEntriesI * sizeof(int32_t)) != cudaSuccess) return nullptr;
    if (cudaMemcpy(d, h.data(), kCornerEntriesI * sizeof(int32_t), cudaMemcpyHostToDevice) != cudaSuccess) return nullptr;
    tables[dev] = d;
  }
  return tables[dev];
}

cudaError_t launch_mfe_fast(const MfeBatchArgs& a, const LaunchPlan& p, const IntTables* host_tables, cudaStream_t stream) {
  FastConstI fc;
  fill_fast_const_i(&fc, host_tables);
  fc.corner = device_corner_table_i(host_tables);
  if (!fc.corner) return cudaErrorMemoryAllocation;
  cudaError_t ce = cudaMemcpyToSymbolAsync(c_fast_const_i, &fc, sizeof fc, 0, cudaMemcpyHostToDevice, stream);
  if (ce != cudaSuccess) return ce;
  if (p.threads == 256) {
    if (p.mats_in_smem) ADP_LAUNCH((mfe_fast_kernel<256, true>), a);
    ADP_LAUNCH((mfe_fast_kernel<256, false>), a);
  }
  if (p.mats_in_smem) ADP_LAUNCH((mfe_fast_kernel<512, true>), a);
  ADP_LAUNCH((mfe_fast_kernel<512, false>), a);
}

cudaError_t launch_mfe(const MfeBatchArgs& a, const LaunchPlan& p, cudaStream_t stream) {
  if (p.mats_in_smem) {
    switch (p.group) {
      case 4: ADP_LAUNCH((mfe_batch_kernel<4, true>), a);
      case 2: ADP_LAUNCH((mfe_batch_kernel<2, true>), a);
      default: ADP_LAUNCH((mfe_batch_kernel<1, true>), a);
    }
  }
  switch (p.group) {
    case 4: ADP_LAUNCH((mfe_batch_kernel<4, false>), a);
    case 2: ADP_LAUNCH((mfe_batch_kernel<2, false>), a);
    default: ADP_LAUNCH((mfe_batch_kernel<1, false>), a);
  }
}

cudaError_t launch_dfma_probe(double* out, int iters, int grid, int threads, cudaStream_t stream) {
  dfma_probe_kernel<<<grid, threads, 0, stream>>>(out, iters);
  return cudaGetLastError();
}

}  // namespace adp
