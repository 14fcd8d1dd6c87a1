#include <cmath>
#include <functional>

#include "addapt/sampling.hh"

namespace addapt {

MonteCarlo::MonteCarlo()
    : my_steps(0), my_thermostat(make_shared<FixedThermostat>(1)), my_scorefxn(make_shared<ScoreFunction>()) {}

DevicePtr MonteCarlo::apply(DevicePtr device, std::mt19937& rng) const {
  if (my_moves.empty()) return device;

  MonteCarloStep step;
  step.num_steps = my_steps;
  step.i = -1;
  step.current_device = device;
  step.score_diff = step.metropolis_criterion = step.random_threshold = step.temperature = 0;
  step.outcome = OutcomeEnum::ACCEPT_UNCHANGED;

  // The reference binds COPIES of the generator into its two distributions (src/sampling.cc:36-37): acceptance
  // draws and move selection run on private generators that start in the caller's state; only the move itself
  // advances the caller's generator.  Kept as is, so a seed means the same thing here.
  auto uniform01 = std::bind(std::uniform_real_distribution<>(), rng);
  auto pick_move = std::bind(std::uniform_int_distribution<>(0, (int)my_moves.size() - 1), rng);

  step.current_score = my_scorefxn->evaluate(step.current_device, step.score_table);
  step.proposed_score = step.current_score;
  for (OutcomeEnum o : {OutcomeEnum::REJECT, OutcomeEnum::ACCEPT_WORSENED, OutcomeEnum::ACCEPT_UNCHANGED, OutcomeEnum::ACCEPT_IMPROVED})
    step.outcome_counters[o] = 0;
  for (auto const& r : my_reporters) r->start(step);

  for (step.i = 0; step.i < step.num_steps; step.i++) {
    step.temperature = my_thermostat->adjust(step);           // every iteration, scored or not
    step.proposed_device = step.current_device->copy();
    step.move = my_moves[pick_move()];
    step.move->apply(step.proposed_device, rng);

    if (step.current_device->seq() == step.proposed_device->seq()) {
      step.outcome = OutcomeEnum::ACCEPT_UNCHANGED;            // identity proposal: not scored
    } else {
      step.proposed_score = my_scorefxn->evaluate(step.proposed_device, step.score_table);
      step.score_diff = step.proposed_score - step.current_score;
      step.metropolis_criterion = std::exp(step.score_diff / step.temperature);
      step.random_threshold = uniform01();
      if (step.metropolis_criterion < step.random_threshold) {  // NaN (0/0 at T = 0) compares false: accepted
        step.outcome = OutcomeEnum::REJECT;
      } else {
        step.outcome = step.score_diff > 0 ? OutcomeEnum::ACCEPT_IMPROVED : OutcomeEnum::ACCEPT_WORSENED;
        step.current_device = step.proposed_device;
        step.current_score = step.proposed_score;
      }
    }
    step.outcome_counters[step.outcome] += 1;
    for (auto const& r : my_reporters) r->update(step);
  }
  for (auto const& r : my_reporters) r->finish(step);
  return step.current_device;
}

int MonteCarlo::num_steps() const { return my_steps; }
void MonteCarlo::num_steps(int n) { my_steps = n; }
ThermostatPtr MonteCarlo::thermostat() const { return my_thermostat; }
void MonteCarlo::thermostat(ThermostatPtr t) { my_thermostat = t; }
ScoreFunctionPtr MonteCarlo::scorefxn() const { return my_scorefxn; }
void MonteCarlo::scorefxn(ScoreFunctionPtr s) { my_scorefxn = s; }
MoveList MonteCarlo::moves() const { return my_moves; }
void MonteCarlo::add_move(MovePtr m) { my_moves.push_back(m); }
void MonteCarlo::operator+=(MovePtr m) { add_move(m); }
ReporterList MonteCarlo::reporters() const { return my_reporters; }
void MonteCarlo::add_reporter(ReporterPtr r) { my_reporters.push_back(r); }
void MonteCarlo::operator+=(ReporterPtr r) { add_reporter(r); }

}  // namespace addapt


namespace std {
ostream& operator<<(ostream& out, const addapt::OutcomeEnum& o) {
  switch (o) {
    case addapt::OutcomeEnum::REJECT: return out << "REJECT";
    case addapt::OutcomeEnum::ACCEPT_WORSENED: return out << "ACCEPT_WORSENED";
    case addapt::OutcomeEnum::ACCEPT_UNCHANGED: return out << "ACCEPT_UNCHANGED";
    default: return out << "ACCEPT_IMPROVED";
  }
}
}
