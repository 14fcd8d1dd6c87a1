// sm_100a kernels for the design loop: one candidate design (score) or one trajectory (Monte Carlo) per
// CTA.  Sequences, scores, RNG state and acceptance never leave the GPU inside the loop.
#include "design_kernels.h"

namespace adp {

extern __shared__ __align__(16) unsigned char adp_dsmem[];

struct DesignSmem {
  void* fold_small;
  double* mats;
  double* dstate;      // fast folds of constructs longer than threads/2: level sums in global scratch (else NULL)
  DesignScratch d;
};

template <bool kFast, bool kSmemMats>
__device__ __forceinline__ DesignSmem carve_design_smem(int n_max, double* gmats, size_t gmat_stride) {
  DesignSmem s;
  s.fold_small = adp_dsmem;
  size_t off = ((kFast ? fast_small_bytes(n_max) : pf_small_bytes(n_max)) + 15) & ~(size_t)15;
  Carver cv(adp_dsmem + off);
  carve_design(cv, s.d, n_max);
  if (kSmemMats) {
    s.mats = cv.take<double>(0);
  } else {
    s.mats = gmats + (size_t)blockIdx.x * gmat_stride;
  }
  s.dstate = nullptr;
  if (kFast && n_max > fast_max_len((int)blockDim.x))
    s.dstate = gmats + (size_t)blockIdx.x * gmat_stride + (kSmemMats ? 0 : fast_matrix_doubles(n_max));
  return s;
}

// kFast: folds with pf_fast.cuh (G must be 1); designs it declines are flagged in a.redo for the general variant.
template <bool kFast, int G, bool kSmemMats>
__global__ void __launch_bounds__(512) score_batch_kernel(const ScoreBatchArgs a) {
  CtaExec<G> ex;
  DesignSmem s = carve_design_smem<kFast, kSmemMats>(a.sp.n_max, a.gmats, a.gmat_stride);
  const int n = a.sp.n, nt_terms = a.sp.n_ctx * a.sp.n_tot();
  s.d.bp_out = a.bp_out ? a.bp_out + (size_t)blockIdx.x * (kMaxBp + kMaxDensPos) : nullptr;
  if (!kFast && a.only_redo && a.redo[a.B] == 0) return;
  for (int b = blockIdx.x; b < a.B; b += gridDim.x) {
    if (!kFast && a.only_redo && !a.redo[b]) continue;
    for (int p = threadIdx.x; p < n; p += blockDim.x) s.d.cur[p] = a.seqs[(size_t)b * n + p];
    __syncthreads();
    bool declined = false;
    const double sc = score_construct<kFast>(ex, a.sp, s.d.cur, s.d, s.fold_small, s.mats, &declined, s.dstate);
    if (declined) {
      if (threadIdx.x == 0) { a.redo[b] = 1; atomicAdd(a.redo + a.B, 1); }
    } else {
      if (threadIdx.x == 0) a.score[b] = sc;
      if (a.terms)
        for (int t = threadIdx.x; t < nt_terms; t += blockDim.x) a.terms[(size_t)b * nt_terms + t] = s.d.terms[t];
    }
    __syncthreads();
  }
}

template <bool kFast, int G, bool kSmemMats>
__global__ void __launch_bounds__(512) mc_kernel(const McArgs a) {
  CtaExec<G> ex;
  DesignSmem s = carve_design_smem<kFast, kSmemMats>(a.sp.n_max, a.gmats, a.gmat_stride);
  __shared__ McState st;
  const int n = a.sp.n;
  s.d.bp_out = a.bp_out ? a.bp_out + (size_t)blockIdx.x * (kMaxBp + kMaxDensPos) : nullptr;
  for (int t = blockIdx.x; t < a.n_traj; t += gridDim.x) {
    if (a.step0 == 0) {
      for (int p = threadIdx.x; p < n; p += blockDim.x) s.d.cur[p] = a.init_seq[p];
    } else {
      // resume: sequences and scalar state
      for (int p = threadIdx.x; p < n; p += blockDim.x) {
        s.d.cur[p] = a.final_seq[(size_t)t * n + p];
        s.d.best[p] = a.best_seq[(size_t)t * n + p];
      }
      if (threadIdx.x == 0) st = a.state[t];
    }
    s.d.train = a.train_state ? a.train_state + (size_t)t * a.train_stride : nullptr;     // the training window stays in device memory
    McTrace tr;
    tr.interval = a.trace_interval;
    tr.cap = a.trace_cap;
    if (a.trace_interval > 0) {
      const int nt_terms = a.sp.n_ctx * a.sp.n_tot();
      tr.vals = a.trace_vals + (size_t)t * a.trace_cap * (kTraceScalars + nt_terms);
      tr.seqs = a.trace_seqs + (size_t)t * a.trace_cap * 2 * n;
      tr.terms_state = a.trace_terms_state + (size_t)t * nt_terms;
    }
    __syncthreads();
    mc_trajectory<kFast>(ex, a.sp, *a.moves, a.seed, a.traj_id0 + (uint64_t)t, a.step0, a.step1, a.thermo_kind, a.thermo_args, s.d, s.fold_small,
                         s.mats, &st, s.dstate, &tr);
    __syncthreads();
    for (int p = threadIdx.x; p < n; p += blockDim.x) {
      a.final_seq[(size_t)t * n + p] = s.d.cur[p];
      a.best_seq[(size_t)t * n + p] = s.d.best[p];
    }
    if (threadIdx.x == 0) {
      a.state[t] = st;
      if (a.final_score) a.final_score[t] = st.cur_score;
      if (a.best_score) a.best_score[t] = st.best_score;
      if (a.best_step) a.best_step[t] = st.best_step;
      if (a.outcomes)
        for (int k = 0; k < 4; k++) a.outcomes[(size_t)t * 4 + k] = st.outcomes[k];
    }
    __syncthreads();
  }
}

// One (sequence, slot, nucleotide) triple per thread: the move table's closure applied by the code the Monte-Carlo kernel uses.
__global__ void mutate_kernel(const MutateArgs a) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= a.B) return;
  uint8_t* out = a.out + (size_t)b * a.n;
  const uint8_t* in = a.seqs + (size_t)b * a.n;
  for (int p = 0; p < a.n; p++) out[p] = in[p];
  if (a.slots[b] >= 0) apply_closure(*a.moves, a.slots[b], a.nts[b], out);
}

cudaError_t launch_mutate(const MutateArgs& a, cudaStream_t stream) {
  mutate_kernel<<<(a.B + 127) / 128, 128, 0, stream>>>(a);
  return cudaGetLastError();
}

LaunchPlan plan_design(const DeviceInfo& d, int n_max, int jobs, bool fast, bool outside) {
  LaunchPlan p = fast ? plan_pf_fast(d, n_max, jobs) : plan_pf(d, n_max, jobs, outside);
  if (fast && !p.threads) return p;
  // same shape as an inside-only fold, plus the design scratch
  const size_t fold_small = fast ? fast_small_bytes(n_max) : pf_small_bytes(n_max);
  const size_t nmat = fast ? fast_matrix_doubles(n_max) : pf_matrix_doubles(n_max, outside);
  const size_t small = ((fold_small + 15) & ~(size_t)15) + design_extra_bytes(n_max) + 64;
  const size_t mats = nmat * sizeof(double);
  p.mats_in_smem = small + mats <= d.smem_optin;
  p.smem_bytes = p.mats_in_smem ? small + mats : small;
  p.gmat_stride = (p.mats_in_smem ? 0 : nmat) + ((fast && n_max > fast_max_len(p.threads)) ? fast_state_doubles(n_max) : 0);
  size_t per_sm = d.smem_per_sm / (p.smem_bytes + 1024);
  size_t by_threads = 2048 / p.threads;
  if (fast && by_threads > 2) by_threads = 2;               // 128 registers per thread
  size_t occ = per_sm < by_threads ? per_sm : by_threads;
  if (occ < 1) occ = 1;
  long grid = (long)d.sm_count * (long)occ;
  if (grid > jobs) grid = jobs;
  if (grid < 1) grid = 1;
  p.grid = (int)grid;
  return p;
}

#define ADP_DLAUNCH(KERNEL, ARGS)                                                                         \
  do {                                                                                                    \
    cudaError_t e_ = cudaFuncSetAttribute(KERNEL, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem_bytes); \
    if (e_ != cudaSuccess) return e_;                                                                     \
    KERNEL<<<p.grid, p.threads, p.smem_bytes, stream>>>(ARGS);                                            \
    return cudaGetLastError();                                                                            \
  } while (0)

static cudaError_t upload_fast_const(const PfTables* host_tables, cudaStream_t stream) {
  FastConst fc;
  fill_fast_const(&fc, host_tables);
  fc.corner = device_corner_table(host_tables);
  if (!fc.corner) return cudaErrorMemoryAllocation;
  return cudaMemcpyToSymbolAsync(c_fast_const, &fc, sizeof fc, 0, cudaMemcpyHostToDevice, stream);
}

cudaError_t launch_score(const ScoreBatchArgs& a, const LaunchPlan& p, bool fast, const PfTables* host_tables, cudaStream_t stream) {
  if (fast) {
    cudaError_t ce = upload_fast_const(host_tables, stream);
    if (ce != cudaSuccess) return ce;
    if (p.mats_in_smem) ADP_DLAUNCH((score_batch_kernel<true, 1, true>), a);
    ADP_DLAUNCH((score_batch_kernel<true, 1, false>), a);
  }
  if (p.mats_in_smem) {
    switch (p.group) {
      case 4: ADP_DLAUNCH((score_batch_kernel<false, 4, true>), a);
      case 2: ADP_DLAUNCH((score_batch_kernel<false, 2, true>), a);
      default: ADP_DLAUNCH((score_batch_kernel<false, 1, true>), a);
    }
  }
  switch (p.group) {
    case 4: ADP_DLAUNCH((score_batch_kernel<false, 4, false>), a);
    case 2: ADP_DLAUNCH((score_batch_kernel<false, 2, false>), a);
    default: ADP_DLAUNCH((score_batch_kernel<false, 1, false>), a);
  }
}

cudaError_t launch_mc(const McArgs& a, const LaunchPlan& p, bool fast, const PfTables* host_tables, cudaStream_t stream) {
  if (fast) {
    cudaError_t ce = upload_fast_const(host_tables, stream);
    if (ce != cudaSuccess) return ce;
    if (p.mats_in_smem) ADP_DLAUNCH((mc_kernel<true, 1, true>), a);
    ADP_DLAUNCH((mc_kernel<true, 1, false>), a);
  }
  if (p.mats_in_smem) {
    switch (p.group) {
      case 4: ADP_DLAUNCH((mc_kernel<false, 4, true>), a);
      case 2: ADP_DLAUNCH((mc_kernel<false, 2, true>), a);
      default: ADP_DLAUNCH((mc_kernel<false, 1, true>), a);
    }
  }
  switch (p.group) {
    case 4: ADP_DLAUNCH((mc_kernel<false, 4, false>), a);
    case 2: ADP_DLAUNCH((mc_kernel<false, 2, false>), a);
    default: ADP_DLAUNCH((mc_kernel<false, 1, false>), a);
  }
}

}  // namespace adp
