#include <algorithm>
#include <cmath>

#include "addapt/thermostats.hh"

namespace addapt {

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

}  // namespace addapt
