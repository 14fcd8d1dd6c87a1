// Host facade of addapt_b200 -- reporters (reference include/sampling.hh:224-259, src/sampling.cc:414-494).
#pragma once
#include "addapt/mc_step.hh"

namespace addapt {

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
