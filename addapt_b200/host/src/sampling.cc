#include "addapt/sampling.hh"

#include <unistd.h>

#include <cctype>
#include <cmath>
#include <cstdio>
#include <functional>
#include <iostream>

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

bool can_be_mutated(DeviceConstPtr device, int position) { return std::isupper((unsigned char)device->seq()[position]) != 0; }

bool can_be_freely_mutated(DeviceConstPtr device, int position) {
  if (!can_be_mutated(device, position)) return false;
  for (auto const& ms : device->macrostates())
    if (ms.second[position] == ')') return false;          // mutated together with its 5' partner instead
  return true;
}

void mutate_recursively(DevicePtr device, int const position, char const mutation) {
  vector<bool> done(device->len(), false);
  mutate_recursively(device, position, mutation, done);
}

void mutate_recursively(DevicePtr device, int const position, char const mutation, vector<bool>& already_mutated) {
  device->mutate(position, mutation);
  already_mutated[position] = true;

  for (auto const& ms : device->macrostates()) {
    string const& brackets = ms.second;
    const char here = brackets[position];
    if (here != '(' && here != ')') continue;
    // walk towards the partner, counting nesting depth
    const int dir = here == '(' ? 1 : -1;
    const char same = here, other = here == '(' ? ')' : '(';
    int depth = 1, partner = position;
    while (depth != 0) {
      partner += dir;
      if (partner < 0 || partner >= (int)brackets.size())
        throw msg("mismatched base-pair in '", ms.first, "' macrostate: '", brackets, "'");
      if (brackets[partner] == same) depth++;
      else if (brackets[partner] == other) depth--;
    }
    if (!can_be_mutated(device, partner))
      throw msg("position '", position, "' can be mutated, but it's base-paired to position '", partner, "' which can't be.");
    // Watson-Crick only: GU pairs would bias the nucleotide distribution (reference src/sampling.cc:220-222)
    const char complement = COMPLEMENTARY_NUCS.at(mutation);
    if (!already_mutated[partner]) mutate_recursively(device, partner, complement, already_mutated);
    else if (device->seq()[partner] != complement) throw "no way to satisfy all base pairing constraints.";
  }
}

UnbiasedMutationMove::UnbiasedMutationMove() {}

void UnbiasedMutationMove::apply(DevicePtr device, std::mt19937& rng) const {
  vector<int> free_positions;
  for (int i = 0; i < device->len(); i++)
    if (can_be_freely_mutated(device, i)) free_positions.push_back(i);
  const int where = free_positions[std::uniform_int_distribution<>(0, (int)free_positions.size() - 1)(rng)];
  const char what = "ACGU"[std::uniform_int_distribution<>(0, 3)(rng)];
  mutate_recursively(device, where, what);
}

FixedThermostat::FixedThermostat(double t) : my_temperature(t) {}
double FixedThermostat::adjust(MonteCarloStep const&) { return my_temperature; }
double FixedThermostat::temperature() const { return my_temperature; }
void FixedThermostat::temperature(double t) { my_temperature = t; }

AnnealingThermostat::AnnealingThermostat(int cycle_len, double max_temperature, double min_temperature)
    : my_cycle_len(cycle_len), my_max_temperature(max_temperature), my_min_temperature(min_temperature) {}
double AnnealingThermostat::adjust(MonteCarloStep const& step) {
  return ((my_min_temperature - my_max_temperature) / my_cycle_len) * (step.i % my_cycle_len) + my_max_temperature;
}
int AnnealingThermostat::cycle_len() const { return my_cycle_len; }
void AnnealingThermostat::cycle_len(int v) { my_cycle_len = v; }
double AnnealingThermostat::max_temperature() const { return my_max_temperature; }
void AnnealingThermostat::max_temperature(double v) { my_max_temperature = v; }
double AnnealingThermostat::min_temperature() const { return my_min_temperature; }
void AnnealingThermostat::min_temperature(double v) { my_min_temperature = v; }

AutoScalingThermostat::AutoScalingThermostat(double target_acceptance_rate, unsigned training_period, double initial_temperature)
    : my_temperature(initial_temperature), my_target_acceptance_rate(target_acceptance_rate), my_training_period(training_period) {}

double AutoScalingThermostat::adjust(MonteCarloStep const& step) {
  my_training_set.push_back(step.score_diff);
  if (my_training_set.size() >= my_training_period) {
    const size_t mid = my_training_set.size() / 2;
    std::nth_element(my_training_set.begin(), my_training_set.begin() + mid, my_training_set.end());
    my_temperature = std::max(my_training_set[mid] / std::log(my_target_acceptance_rate), 0.0);
    my_training_set.clear();
  }
  return my_temperature;
}
void AutoScalingThermostat::initial_temperature(double v) { my_temperature = v; }
void AutoScalingThermostat::target_acceptance_rate(double v) { my_target_acceptance_rate = v; }
void AutoScalingThermostat::training_period(unsigned v) { my_training_period = v; }

void ProgressReporter::update(MonteCarloStep const& step) {
  if (!isatty(fileno(stdout))) return;
  std::cout << "\033[2K\r[" << (step.i + 1) << "/" << step.num_steps << "]";
  if (step.i + 1 == step.num_steps) std::cout << std::endl;
  else std::cout << std::flush;
}

TsvTrajectoryReporter::TsvTrajectoryReporter(string path, int interval) : my_path(path), my_interval(interval) {}

void TsvTrajectoryReporter::start(MonteCarloStep const& step) {
  my_tsv.open(my_path);
  if (!my_tsv.is_open()) throw msg("couldn't open '", my_path, "' for writing");
  my_tsv << "#\tinitial_seq\t" << step.current_device->seq() << "\n";
  my_tsv << "step\tnum_steps\tcurrent_score\tproposed_score\t";
  for (auto const& row : step.score_table) my_tsv << "term_weight[" << row.name << "]\tterm_value[" << row.name << "]\t";
  my_tsv << "score_diff\ttemperature\tmetropolis_criterion\trandom_threshold\tmove\toutcome\tcurrent_seq\tproposed_seq\t" << std::endl;
}

void TsvTrajectoryReporter::update(MonteCarloStep const& step) {
  if (step.i % my_interval != 0) return;
  my_tsv << step.i << "\t" << step.num_steps << "\t" << step.current_score << "\t" << step.proposed_score << "\t";
  for (auto const& row : step.score_table) my_tsv << row.weight << "\t" << row.term << "\t";
  my_tsv << step.score_diff << "\t" << step.temperature << "\t" << step.metropolis_criterion << "\t" << step.random_threshold << "\t"
         << step.move->name() << "\t" << step.outcome << "\t" << step.current_device->seq() << "\t" << step.proposed_device->seq() << "\t"
         << std::endl;
}

void TsvTrajectoryReporter::finish(MonteCarloStep const&) { my_tsv.close(); }

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
