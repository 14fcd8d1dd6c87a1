// Host facade of addapt_b200 -- the per-step record shared by the Monte-Carlo loop, thermostats and reporters
// (reference include/sampling.hh:20-27,86-104).
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

}  // namespace addapt

namespace std {
ostream& operator<<(ostream&, const addapt::OutcomeEnum&);
}
