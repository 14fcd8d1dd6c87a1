// Launch interface between the C-ABI layer (engine.cc) and the sm_100a kernels (fold_kernels.cu,
// design_kernels.cu).  Plain structs of device pointers; every launch_* enqueues on `stream` and returns
// the cudaError_t of the launch.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "fold_common.cuh"
#include "tables.h"

namespace adp {

struct PfBatchArgs {
  const uint8_t* seqs;         // B rows of n bytes
  int B, n, V;
  VariantDev variants[kMaxVariants];
  const ConstraintDev* csts;
  const MotifDev* motif;
  const PfTables* T;
  double* dG;                  // [B*V]
  double* bppm;                // NULL or [B*V*n(n+1)/2]
  int n_query;
  const int32_t* qi;
  const int32_t* qj;
  double* qprobs;              // [B*V*n_query]
  double* gmats;               // global scratch for the matrices when they do not fit in shared memory
  size_t gmat_stride;          // doubles per CTA
  int* redo;                   // [B*V + 1] or NULL: jobs the fast kernel declined (non-ACGU letters); last = their count
  int only_redo;               // general kernel: fold only the jobs flagged in `redo`
};

struct MfeBatchArgs {
  const uint8_t* seqs;
  int B, n;
  const ConstraintDev* cst;    // NULL = unconstrained
  const MotifDev* motif;       // NULL = apo
  const IntTables* T;
  int32_t* e_dcal;             // [B]
  char* structs;               // NULL or [B*n]
  int32_t* gmats;
  size_t gmat_stride;          // ints per CTA
  int* redo;                   // [B + 1] or NULL: folds the fast kernel declined (non-ACGU letters); last = their count
  int only_redo;               // general kernel: fold only the flagged sequences
};

struct LaunchPlan {
  int grid, threads, group;
  size_t smem_bytes;
  bool mats_in_smem;
  size_t gmat_stride;          // elements per CTA of global scratch (0 when mats_in_smem)
};

struct DeviceInfo {
  int sm_count;
  size_t smem_optin;           // max dynamic shared memory per CTA
  size_t smem_per_sm;
};

LaunchPlan plan_pf(const DeviceInfo&, int n, int jobs, bool any_outside);
// fast path (pf_fast.cuh); plan.threads == 0 when the length is outside its range
LaunchPlan plan_pf_fast(const DeviceInfo&, int n, int jobs);
LaunchPlan plan_mfe_fast(const DeviceInfo&, int n, int jobs);
cudaError_t launch_mfe_fast(const MfeBatchArgs&, const LaunchPlan&, const IntTables* host_tables, cudaStream_t);
// host_tables: the engine's Boltzmann tables on the host (their size-indexed factors go to the constant bank)
cudaError_t launch_pf_fast(const PfBatchArgs&, const LaunchPlan&, const PfTables* host_tables, cudaStream_t);
LaunchPlan plan_mfe(const DeviceInfo&, int n, int jobs);
cudaError_t launch_pf(const PfBatchArgs&, const LaunchPlan&, cudaStream_t);
cudaError_t launch_mfe(const MfeBatchArgs&, const LaunchPlan&, cudaStream_t);
cudaError_t launch_dfma_probe(double* out, int iters, int grid, int threads, cudaStream_t);
const double* device_corner_table(const PfTables* host_tables);     // see pf_fast.cuh build_corner_table

}  // namespace adp
