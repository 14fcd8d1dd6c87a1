// Host facade of addapt_b200 -- Monte-Carlo sampling.  Same interface as the reference's include/sampling.hh:
// MonteCarlo (:29-93), MonteCarloStep (:95-104), can_be_mutated / can_be_freely_mutated / mutate_recursively
// (:110-120), Move / UnbiasedMutationMove (:123-143), Thermostat x3 (:146-221), Reporter x2 (:224-259).
// MonteCarlo::apply is the reference's single-trajectory host loop (every score is one GPU call); the throughput
// path is BatchedMonteCarlo in addapt/b200.hh, which runs the same loop for thousands of trajectories in one kernel.
#pragma once
#include "addapt/mc_step.hh"
#include "addapt/moves.hh"
#include "addapt/reporters.hh"
#include "addapt/thermostats.hh"

namespace addapt {

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

}  // namespace addapt
