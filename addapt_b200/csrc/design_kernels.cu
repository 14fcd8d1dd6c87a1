// sm_100a kernels for the design loop: one candidate design (score) or one trajectory (Monte Carlo) per
// CTA.  Sequences, scores, RNG state and acceptance never leave the GPU inside the loop.
#include <cstdlib>

#include "design_kernels.h"

namespace adp {

extern __shared__ __align__(16) unsigned char adp_dsmem[];

struct DesignSmem {
  void* fold_small;
  double* mats;
  double* qm_ext;      // fast folds of 101-128 nt: Qm in global scratch so that two CTAs share an SM (else NULL)
  double* dstate;      // fast folds of constructs longer than threads/2: level sums in global scratch (else NULL)
  DesignScratch d;
};

template <bool kFast, bool kSmemMats, bool kQmGlobal = false>
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
  s.qm_ext = kQmGlobal ? gmats + (size_t)blockIdx.x * gmat_stride : nullptr;
  s.dstate = nullptr;
  if (kFast && n_max > fast_max_len((int)blockDim.x))
    s.dstate = gmats + (size_t)blockIdx.x * gmat_stride + (kSmemMats ? (kQmGlobal ? t4_count(n_max) + 2 : 0) : fast_matrix_doubles(n_max));
  return s;
}

// kFast: folds with pf_fast.cuh (G must be 1); designs it declines are flagged in a.redo for the general variant.
template <bool kFast, int G, bool kSmemMats, bool kQmGlobal = false>
__global__ void __launch_bounds__(512) score_batch_kernel(const ScoreBatchArgs a) {
  CtaExec<G> ex;
  DesignSmem s = carve_design_smem<kFast, kSmemMats, kQmGlobal>(a.sp.n_max, a.gmats, a.gmat_stride);
  const int n = a.sp.n, nt_terms = a.sp.n_ctx * a.sp.n_tot();
  s.d.bp_out = a.bp_out ? a.bp_out + (size_t)blockIdx.x * (kMaxBp + kMaxDensPos) : nullptr;
  if (!kFast && a.only_redo && a.redo[a.B] == 0) return;
  for (int b = blockIdx.x; b < a.B; b += gridDim.x) {
    if (!kFast && a.only_redo && !a.redo[b]) continue;
    for (int p = threadIdx.x; p < n; p += blockDim.x) s.d.cur[p] = a.seqs[(size_t)b * n + p];
    __syncthreads();
    bool declined = false;
    const double sc = score_construct<kFast>(ex, a.sp, s.d.cur, s.d, s.fold_small, s.mats, &declined, s.dstate, s.qm_ext);
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

template <bool kFast, int G, bool kSmemMats, bool kQmGlobal = false>
__global__ void __launch_bounds__(512) mc_kernel(const McArgs a) {
  CtaExec<G> ex;
  DesignSmem s = carve_design_smem<kFast, kSmemMats, kQmGlobal>(a.sp.n_max, a.gmats, a.gmat_stride);
  __shared__ McState st;
  __shared__ McTrace tr;
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
    if (threadIdx.x == 0) {            // in shared memory like the rest of the trajectory's state: registers are for the fold
      tr.interval = a.trace_interval;
      tr.cap = a.trace_cap;
      if (a.trace_interval > 0) {
        const int nt_terms = a.sp.n_ctx * a.sp.n_tot();
        tr.vals = a.trace_vals + (size_t)t * a.trace_cap * (kTraceScalars + nt_terms);
        tr.seqs = a.trace_seqs + (size_t)t * a.trace_cap * 2 * n;
        tr.terms_state = a.trace_terms_state + (size_t)t * nt_terms;
      }
    }
    __syncthreads();
    mc_trajectory<kFast>(ex, a.sp, *a.moves, a.seed, a.traj_id0 + (uint64_t)t, a.step0, a.step1, a.thermo_kind, a.thermo_args, s.d, s.fold_small,
                         s.mats, &st, s.dstate, &tr, s.qm_ext);
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

// ---------------------------------------------------------------------------------------------------
// Per-rank reduction of a design run ("each rank reduces to its local top-K records, then one all-gather"):
// the k best of n scores without leaving the device.  One CTA of 1,024 threads:
//   1. radix select, 8 passes of 8 bits over an order-preserving 64-bit key (larger score = larger key, NaN = 0): the key T of
//      the k-th best score and the number of scores strictly better;
//   2. one pass in index order that keeps every score better than T and the first (k - better) scores equal to T
//      (block-wide prefix sums keep the index order: ties go to the lower index, whatever the thread count);
//   3. bitonic sort of the k (key, index) records in shared memory: best first, ties by lower index.
__device__ __forceinline__ unsigned long long topk_key(double x) {
  if (x != x) return 0ull;                                   // NaN ranks below everything, -inf included
  if (x == 0.) x = 0.;                                       // -0.0 and +0.0 are one score
  unsigned long long u = (unsigned long long)__double_as_longlong(x);
  u = (u >> 63) ? ~u : (u | 0x8000000000000000ull);          // ascending in x
  return u == 0ull ? 1ull : u;                               // keep 0 for NaN (only -NaN patterns map to 0 anyway)
}

__global__ void __launch_bounds__(1024) top_k_kernel(const double* __restrict__ scores, int n, int k, int32_t* top_index, double* top_score) {
  __shared__ unsigned int hist[256];
  __shared__ unsigned long long s_prefix;
  __shared__ int s_need, s_base[2], warp_tot[2][32];
  __shared__ unsigned long long rec_key[kTopKMax];
  __shared__ int rec_idx[kTopKMax];
  const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, wid = tid >> 5;
  if (k > n) k = n;
  if (tid == 0) { s_prefix = 0ull; s_need = k; }
  __syncthreads();
  // ---- 1. threshold key
  for (int shift = 56; shift >= 0; shift -= 8) {
    for (int b = tid; b < 256; b += nt) hist[b] = 0u;
    __syncthreads();
    const unsigned long long prefix = s_prefix;
    const unsigned long long mask = shift == 56 ? 0ull : (~0ull << (shift + 8));
    for (int i = tid; i < n; i += nt) {
      const unsigned long long key = topk_key(scores[i]);
      if ((key & mask) == prefix) atomicAdd(&hist[(key >> shift) & 255ull], 1u);
    }
    __syncthreads();
    if (tid == 0) {
      int need = s_need, b = 255;
      for (; b > 0; b--) {                    // from the best bucket down
        if ((int)hist[b] >= need) break;
        need -= (int)hist[b];
      }
      s_need = need;                          // still wanted inside bucket b
      s_prefix = prefix | ((unsigned long long)b << shift);
    }
    __syncthreads();
  }
  const unsigned long long T = s_prefix;
  const int need_eq = s_need;                 // how many scores equal to T are kept (>= 1)
  // ---- 2. collect, in index order
  if (tid == 0) { s_base[0] = 0; s_base[1] = 0; }
  __syncthreads();
  for (int i0 = 0; i0 < n; i0 += nt) {
    const int i = i0 + tid;
    unsigned long long key = 0ull;
    bool gt = false, eq = false;
    if (i < n) { key = topk_key(scores[i]); gt = key > T; eq = key == T; }
    const unsigned bg = __ballot_sync(0xffffffffu, gt), be = __ballot_sync(0xffffffffu, eq);
    if (lane == 0) { warp_tot[0][wid] = __popc(bg); warp_tot[1][wid] = __popc(be); }
    __syncthreads();
    int pre_g = 0, pre_e = 0;
    for (int w = 0; w < wid; w++) { pre_g += warp_tot[0][w]; pre_e += warp_tot[1][w]; }
    const unsigned below = (1u << lane) - 1u;
    const int rank_g = s_base[0] + pre_g + __popc(bg & below);       // among the better scores so far
    const int rank_e = s_base[1] + pre_e + __popc(be & below);       // among the equal scores so far
    if (gt) { rec_key[rank_g] = key; rec_idx[rank_g] = i; }          // fewer than k of them by construction
    if (eq && rank_e < need_eq) { rec_key[k - need_eq + rank_e] = key; rec_idx[k - need_eq + rank_e] = i; }
    __syncthreads();
    if (tid == 0) {
      int tg = 0, te = 0;
      for (int w = 0; w < (nt >> 5); w++) { tg += warp_tot[0][w]; te += warp_tot[1][w]; }
      s_base[0] += tg;
      s_base[1] += te;
    }
    __syncthreads();
  }
  // ---- 3. sort the k records: key descending, index ascending (pad to a power of two with the worst possible record)
  int m = 1;
  while (m < k) m <<= 1;
  for (int t = k + tid; t < m; t += nt) { rec_key[t] = 0ull; rec_idx[t] = 0x7fffffff; }
  __syncthreads();
  for (int size = 2; size <= m; size <<= 1)
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int t = tid; t < m; t += nt) {
        const int p = t ^ stride;
        if (p > t) {
          const bool up = (t & size) == 0;            // this pair sorts "best first"
          const unsigned long long ka = rec_key[t], kb = rec_key[p];
          const int ia = rec_idx[t], ib = rec_idx[p];
          const bool a_first = ka > kb || (ka == kb && ia < ib);
          if (a_first != up) { rec_key[t] = kb; rec_key[p] = ka; rec_idx[t] = ib; rec_idx[p] = ia; }
        }
      }
      __syncthreads();
    }
  for (int t = tid; t < k; t += nt) {
    top_index[t] = rec_idx[t];
    if (top_score) top_score[t] = scores[rec_idx[t]];
  }
}

cudaError_t launch_top_k(const double* scores, int n, int k, int32_t* top_index, double* top_score, cudaStream_t stream) {
  top_k_kernel<<<1, 1024, 0, stream>>>(scores, n, k, top_index, top_score);
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
  size_t gdoubles = p.mats_in_smem ? 0 : nmat;
  p.group = fast ? 1 : p.group;
  // 101-128 nt (e.g. rhf(6) with a 20-nt spacer): Qm in global scratch keeps two constructs per SM (see plan_pf_fast)
  const size_t qb_bytes = (t4_count(n_max) + 2) * sizeof(double);
  const char* qmg = std::getenv("ADDAPT_B200_QM_GLOBAL");
  if (fast && (qmg ? std::atoi(qmg) != 0 : true) && p.threads == 256 && p.mats_in_smem && d.smem_per_sm / (p.smem_bytes + 1024 + 128) < 2 &&
      d.smem_per_sm / (small + qb_bytes + 1024 + 128) >= 2) {
    p.group = 3;
    p.smem_bytes = small + qb_bytes;
    gdoubles = t4_count(n_max) + 2;
  }
  p.gmat_stride = gdoubles + ((fast && n_max > fast_max_len(p.threads)) ? fast_state_doubles(n_max) : 0);
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
    if (p.group == 3) ADP_DLAUNCH((score_batch_kernel<true, 1, true, true>), a);
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
    if (p.group == 3) ADP_DLAUNCH((mc_kernel<true, 1, true, true>), a);
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
