// Host facade of addapt_b200 -- the binding to the C ABI (include/addapt_b200.h).  This is the code a maintainer
// of the reference would add to switch its hot path to the GPU engine:
//   B200RnaFold        replaces ViennaRnaFold (reference src/scoring.cc:17-103) behind the RnaFold interface;
//   B200ScoreFunction  evaluates many candidate sequences per call (adp_score_batch);
//   BatchedMonteCarlo  runs n independent MonteCarlo::apply trajectories in one kernel launch (adp_mc_run).
// Errors of the C ABI are rethrown as std::string, the reference's exception convention.
#pragma once
#include <cstdint>

#include "addapt/sampling.hh"
#include "addapt/scoring.hh"

struct adp_engine;

namespace addapt {

// RAII owner of one adp_engine (one GPU).  Not thread safe, like the engine itself.
class B200Engine {
 public:
  explicit B200Engine(int device = 0, string par_file = "");
  ~B200Engine();
  B200Engine(B200Engine const&) = delete;
  B200Engine& operator=(B200Engine const&) = delete;
  adp_engine* handle() const { return my_handle; }
  double kT() const;                      // kcal/mol
  void check(int rc) const;               // throws std::string(adp_last_error) when rc != 0
  static int device_count();

 private:
  adp_engine* my_handle;
};
using B200EnginePtr = std::shared_ptr<B200Engine>;

// process-wide default engine on cuda:0, created on first use; throws std::string without a GPU
B200EnginePtr default_engine();
void default_engine(B200EnginePtr);
FoldFactory b200_fold_factory(B200EnginePtr engine = nullptr);

class B200RnaFold : public RnaFold {
 public:
  // Snapshots the device's contextual sequence like ViennaRnaFold does (reference src/scoring.cc:17-29).
  B200RnaFold(DeviceConstPtr, AptamerConstPtr = nullptr, B200EnginePtr = nullptr);
  double base_pair_prob(int a, int b) const;
  double macrostate_prob(string constraint) const;
  // vrna_pf returns float: probabilities are computed from float-rounded ensemble energies unless disabled
  void float_energies(bool on) { my_float = on; }

 private:
  double ensemble_energy(const char* cst) const;
  B200EnginePtr my_engine;
  AptamerConstPtr my_aptamer;
  string my_seq;
  bool my_float = true;
  mutable vector<double> my_bppm;       // lazily filled by the first base_pair_prob call
};

// ScoreFunction whose terms are MacrostateProbTerm, PairProbTerm or PairedDensityTerm: whole batches of candidate sequences of one device are
// scored by one kernel launch.  evaluate(DeviceConstPtr, table&) keeps working (batch of one).
class B200ScoreFunction : public ScoreFunction {
 public:
  explicit B200ScoreFunction(B200EnginePtr = nullptr);
  using ScoreFunction::evaluate;
  double evaluate(DeviceConstPtr, EvaluatedScoreFunction&) const override;
  // raw sequences (device->raw_len() letters each) -> scores; optionally the unweighted terms, row-major
  // [sequence][context][term]
  vector<double> evaluate_batch(DeviceConstPtr device, vector<string> const& raw_seqs, vector<double>* terms = nullptr) const;
  B200EnginePtr engine() const { return my_engine; }
  void float_energies(bool on) { my_float = on; }
  bool float_energies() const { return my_float; }

 private:
  B200EnginePtr my_engine;
  bool my_float = true;
};

struct BatchedTrajectory {
  string final_seq, best_seq;
  double final_score, best_score;
  int best_step;                          // -1 = the start sequence
  std::map<OutcomeEnum, int> outcome_counters;
};

// n independent Monte-Carlo design trajectories on the GPU: UnbiasedMutationMove proposals, Metropolis acceptance,
// Fixed, Annealing or AutoScaling thermostat, all inside one kernel (reference src/sampling.cc:22-107 per trajectory).
// Trajectory t draws from Philox4x32-10 keyed by (seed, first_id + t), so results do not depend on the sharding.
class BatchedMonteCarlo {
 public:
  BatchedMonteCarlo(std::shared_ptr<B200ScoreFunction> scorefxn, int num_steps, ThermostatPtr thermostat);
  vector<BatchedTrajectory> apply(DeviceConstPtr device, int n_traj, uint64_t seed, uint64_t first_id = 0) const;
  // Also write every trajectory to a file in TsvTrajectoryReporter's layout (reference src/sampling.cc:430-489), one row every
  // `interval` steps, from the records the kernel keeps (adp_mc_run_trace).  `path` names the file of a single trajectory;
  // with several, "%d" in it is replaced by the trajectory id, or the id is inserted before the extension (traj.tsv -> traj.17.tsv).
  void trajectory_files(string path, int interval);
  static string trajectory_path(string path, uint64_t id, int n_traj);

 private:
  std::shared_ptr<B200ScoreFunction> my_scorefxn;
  int my_steps;
  ThermostatPtr my_thermostat;
  string my_tsv_path;
  int my_tsv_interval = 0;
};

}  // namespace addapt
