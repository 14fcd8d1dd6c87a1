#include <unistd.h>

#include <cstdio>
#include <iostream>

#include "addapt/moves.hh"
#include "addapt/reporters.hh"

namespace addapt {

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
