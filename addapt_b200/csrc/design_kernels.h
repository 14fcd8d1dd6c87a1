// Launch interface of the design kernels (score function, Monte-Carlo trajectories).
#pragma once
#include "design_core.cuh"
#include "kernels.h"

namespace adp {

struct ScoreBatchArgs {
  ScoreProblemDev sp;
  const uint8_t* seqs;       // B rows of sp.n bytes
  int B;
  double* score;             // [B]
  double* terms;             // NULL or [B * n_ctx * n_terms]
  double* gmats;
  size_t gmat_stride;
  int* redo;                 // [B + 1]: designs the fast kernel declined (letters outside ACGU/T); last = their count
  int only_redo;             // general kernel: score only the flagged designs
  double* bp_out;            // [grid][kMaxBp + kMaxDensPos] scratch of the base-pair-probability terms, or NULL
};

struct McArgs {
  ScoreProblemDev sp;
  const MoveTableDev* moves;
  const uint8_t* init_seq;   // sp.n bytes
  int n_traj;
  uint64_t seed, traj_id0;
  int step0, step1;          // this launch runs steps [step0, step1) of every trajectory (step0 > 0: resumes from `state`)
  int thermo_kind;
  double thermo_args[3];
  McState* state;            // [n_traj] carried between launches (never NULL)
  double* train_state;       // [n_traj * train_stride] training windows of the auto-scaling thermostat, or NULL
  int train_stride;
  uint8_t* final_seq;        // current sequences (never NULL: they are part of the carried state)
  double* final_score;       // the remaining outputs may be NULL
  uint8_t* best_seq;         // best sequences (never NULL)
  double* best_score;
  int32_t* best_step;
  uint32_t* outcomes;
  double* gmats;
  size_t gmat_stride;
  // optional per-step records (McTrace): interval 0 = none; buffers are [n_traj][cap][...] / [n_traj][nt]
  int trace_interval, trace_cap;
  double* trace_vals;
  uint8_t* trace_seqs;
  double* trace_terms_state;
  double* bp_out;            // [grid][kMaxBp + kMaxDensPos] scratch of the base-pair-probability terms, or NULL
};

// ≙ mutate_recursively for B (sequence, table slot, nucleotide) triples, see adp_mutate_batch
struct MutateArgs {
  const MoveTableDev* moves;
  int B, n;
  const uint8_t* seqs;       // B rows of n bytes
  const int32_t* slots;      // [B] slot of the move table (-1: copy the sequence unchanged)
  const uint8_t* nts;        // [B] A, C, G or U
  uint8_t* out;              // B rows of n bytes
};

// fast = fold with pf_fast.cuh; plan.threads == 0 when n_max is outside its range
// outside: the fills also run the outside pass (base-pair-probability terms): the general path then needs a third matrix
LaunchPlan plan_design(const DeviceInfo&, int n_max, int jobs, bool fast, bool outside = false);
// host_tables: the engine's Boltzmann tables on the host (constant-bank factors of the fast kernels)
cudaError_t launch_score(const ScoreBatchArgs&, const LaunchPlan&, bool fast, const PfTables* host_tables, cudaStream_t);
cudaError_t launch_mc(const McArgs&, const LaunchPlan&, bool fast, const PfTables* host_tables, cudaStream_t);
cudaError_t launch_mutate(const MutateArgs&, cudaStream_t);
// the k (<= kTopKMax) best of n scores, best first, ties by lower index, NaN last: one CTA (radix select + bitonic sort)
constexpr int kTopKMax = 1024;
cudaError_t launch_top_k(const double* scores, int n, int k, int32_t* top_index, double* top_score, cudaStream_t);

}  // namespace adp
