// Host facade of addapt_b200 -- Monte-Carlo sampling.  Same interface as the reference's include/sampling.hh:
// MonteCarlo (:29-93), MonteCarloStep (:95-104), can_be_mutated / can_be_freely_mutated / mutate_recursively
// (:110-120), Move / UnbiasedMutationMove (:123-143), Thermostat x3 (:146-221), Reporter x2 (:224-259).
// MonteCarlo::apply is the reference's single-trajectory host loop (every score is one GPU call); the throughput
// path is BatchedMonteCarlo in addapt/b200.hh, which runs the same loop for thousands of trajectories in one kernel.
#pragma once
#include <fstream>
#include <random>

#include "addapt/model.hh"
#include "addapt/scoring.hh"

namespace addapt {

class MonteCarlo;
using MonteCarloPtr = std::shared_ptr<MonteCarlo>;
class Move;
using MovePtr = std::shared_ptr<Move>;
using MoveList = std::vector<MovePtr>;
class Thermostat;
using ThermostatPtr = std::shared_ptr<Thermostat>;
class Reporter;
using ReporterPtr = std::shared_ptr<Reporter>;
using ReporterList = std::vector<ReporterPtr>;

enum class OutcomeEnum { REJECT, ACCEPT_WORSENED, ACCEPT_UNCHANGED, ACCEPT_IMPROVED };

struct MonteCarloStep {
  int i, num_steps;
  DevicePtr current_device, proposed_device;
  MovePtr move;
  EvaluatedScoreFunction score_table;
  double current_score, proposed_score, score_diff;
  double temperature, metropolis_criterion, random_threshold;
  OutcomeEnum outcome;
  std::map<OutcomeEnum, int> outcome_counters;
};

class MonteCarlo {
 public:
  MonteCarlo();
  // Metropolis maximisation; returns the LAST accepted device (reference src/sampling.cc:22-107)
  DevicePtr apply(DevicePtr, std::mt19937&) const;
  int num_steps() const;
  void num_steps(int);
  ThermostatPtr thermostat() const;
  void thermostat(ThermostatPtr);
  ScoreFunctionPtr scorefxn() const;
  void scorefxn(ScoreFunctionPtr);
  MoveList moves() const;
  void add_move(MovePtr);
  void operator+=(MovePtr);
  ReporterList reporters() const;
  void add_reporter(ReporterPtr);
  void operator+=(ReporterPtr);

 private:
  int my_steps;
  ThermostatPtr my_thermostat;
  ScoreFunctionPtr my_scorefxn;
  MoveList my_moves;
  ReporterList my_reporters;
};

map<char, char> const COMPLEMENTARY_NUCS = {{'A', 'U'}, {'G', 'C'}, {'C', 'G'}, {'U', 'A'}};

// upper case = mutable (reference src/sampling.cc:170-175)
bool can_be_mutated(DeviceConstPtr, int position);
// mutable and not the 3' end ')' of a pair in any macrostate (reference src/sampling.cc:177-193)
bool can_be_freely_mutated(DeviceConstPtr, int position);
// point mutation plus Watson-Crick complements through every macrostate's bracket partners
// (reference src/sampling.cc:205-282).  Throws std::string for an unbalanced bracket or an immutable partner,
// const char* when two constraints contradict each other.
void mutate_recursively(DevicePtr, int const position, char const mutation);
void mutate_recursively(DevicePtr, int const position, char const mutation, vector<bool>& already_mutated);

class Move {
 public:
  virtual ~Move() {}
  virtual string name() const = 0;
  virtual void apply(DevicePtr, std::mt19937&) const = 0;
};

// uniform freely-mutable position x uniform ACGU, identity included (reference src/sampling.cc:287-303)
class UnbiasedMutationMove : public Move {
 public:
  UnbiasedMutationMove();
  string name() const { return "UnbiasedMutation"; }
  void apply(DevicePtr, std::mt19937&) const;
};

class Thermostat {
 public:
  virtual ~Thermostat() {}
  virtual double adjust(MonteCarloStep const&) = 0;
};

class FixedThermostat : public Thermostat {
 public:
  FixedThermostat(double);
  double adjust(MonteCarloStep const&);
  double temperature() const;
  void temperature(double);

 private:
  double my_temperature;
};

// saw-tooth T(i) = (T_lo - T_hi)/N * (i mod N) + T_hi (reference src/sampling.cc:331-338)
class AnnealingThermostat : public Thermostat {
 public:
  AnnealingThermostat(int cycle_len, double max_temperature, double min_temperature);
  double adjust(MonteCarloStep const&);
  int cycle_len() const;
  void cycle_len(int);
  double max_temperature() const;
  void max_temperature(double);
  double min_temperature() const;
  void min_temperature(double);

 private:
  int my_cycle_len;
  double my_max_temperature, my_min_temperature;
};

// median score difference of a training window / ln(target acceptance rate) (reference src/sampling.cc:371-401)
class AutoScalingThermostat : public Thermostat {
 public:
  AutoScalingThermostat(double target_acceptance_rate = 0.5, unsigned training_period = 100, double initial_temperature = 1.0);
  double adjust(MonteCarloStep const&);
  void initial_temperature(double);
  void target_acceptance_rate(double);
  void training_period(unsigned);
  // read access (not in the reference's header; BatchedMonteCarlo needs it to hand the thermostat to the GPU)
  double temperature() const { return my_temperature; }
  double target_acceptance_rate() const { return my_target_acceptance_rate; }
  unsigned training_period() const { return my_training_period; }

 private:
  double my_temperature, my_target_acceptance_rate;
  unsigned my_training_period;
  vector<double> my_training_set;
};

class Reporter {
 public:
  virtual ~Reporter() {}
  virtual void start(MonteCarloStep const&) {}
  virtual void update(MonteCarloStep const&) {}
  virtual void finish(MonteCarloStep const&) {}
};

class ProgressReporter : public Reporter {
 public:
  void update(MonteCarloStep const&);
};

// Trajectory file in the reference's column layout (reference src/sampling.cc:430-489): a "#\tinitial_seq\t..."
// line, the header, then one row every `interval` steps.
class TsvTrajectoryReporter : public Reporter {
 public:
  TsvTrajectoryReporter(string path, int interval);
  void start(MonteCarloStep const&);
  void update(MonteCarloStep const&);
  void finish(MonteCarloStep const&);

 private:
  string my_path;
  int my_interval;
  std::ofstream my_tsv;
};

}  // namespace addapt

namespace std {
ostream& operator<<(ostream&, const addapt::OutcomeEnum&);
}
