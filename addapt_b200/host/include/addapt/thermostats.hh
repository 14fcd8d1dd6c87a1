// Host facade of addapt_b200 -- thermostats (reference include/sampling.hh:146-221, src/sampling.cc:305-411).
#pragma once
#include "addapt/mc_step.hh"

namespace addapt {

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

}  // namespace addapt
