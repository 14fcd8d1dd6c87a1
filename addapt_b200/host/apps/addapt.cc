// addapt -- the design driver of the reference (apps/addapt.cc:18-93) on the B200 engine.
//
//   addapt <config>... [-n <num>] [-T <schedule>] [-r <seed>] [-o <path>] [-i <steps>] [--trajectories <k>] [--device <d>]
//
// Same configuration files (sequence / macrostates / objective / aptamer / contexts / thermostat sections, reference
// src/config.cc:54-175), same options and defaults, same trajectory file (TsvTrajectoryReporter's layout).  What differs is the
// engine: the Monte-Carlo loop -- proposals, the four partition-function fills of every score, the Metropolis rule -- runs inside
// one CUDA kernel (BatchedMonteCarlo -> adp_mc_run_trace), and --trajectories runs k independent trajectories at once (ids 0..k-1
// of the Philox stream keyed by the seed; one file each, traj.tsv -> traj.<id>.tsv), which is what the reference does with k
// processes and k seeds (data/20160705_production_run_favor_wt/production_run.sh:5-8).  The random stream is the engine's
// (Philox4x32-10), not std::mt19937: runs are reproducible per seed, not sequence-identical with the reference binary.
// docopt and yaml-cpp are not in this image: options are parsed here, YAML by addapt/config.hh.
#include <cstring>
#include <iostream>
#include <string>

#include "addapt/b200.hh"
#include "addapt/config.hh"

using namespace addapt;

static char const USAGE[] = R"""(Insert the theophylline aptamer into the first hairpin of the sgRNA and run a
Monte Carlo design simulation to optimize variable regions in the nexus, the
ruler, and the rest of the first hairpin.  The design goal is to only form the
wildtype nexus and hairpin base pairs when theophylline is bound.

Usage:
  addapt <config>... [options]

Options:
  -n <num>, --num-moves <num>                [default: 100]
    The number of moves to attempt in the design simulation.

  -T <schedule>, --temperature <schedule>
    The temperature to use for the Metropolis criterion: a fixed temperature
    (e.g. "5"), a multi-cooled simulated annealing schedule (e.g. "1 to 0 in
    500 steps"), or a schedule that tries to achieve a certain acceptance
    rate (e.g. "auto 50%").  Default: the `thermostat` section, else 1.

  -r <seed>, --random-seed <seed>            [default: 0]
    The seed for the random number generator.

  -o <path>, --output <path>                 [default: traj.tsv]
    The path where the trajectory of the design simulation will be saved.

  -i <steps>, --output-interval <steps>      [default: 1]
    How often a new snapshot in the trajectory should be recorded.

  --trajectories <k>                         [default: 1]
    Run k independent trajectories on the GPU at once (one file each).

  --device <index>                           [default: 0]
    The CUDA device to run on.

  --version
  -h, --help
)""";

int main(int argc, char** argv) {
  try {
    vector<string> config_files;
    string num_moves = "100", schedule, seed = "0", output = "traj.tsv", interval = "1", n_traj = "1", dev = "0";
    bool have_schedule = false;
    for (int k = 1; k < argc; k++) {
      const string a = argv[k];
      auto value = [&](string* dst) {
        if (k + 1 >= argc) throw "option " + a + " needs a value";
        *dst = argv[++k];
      };
      if (a == "-h" || a == "--help") { std::cout << USAGE; return 0; }
      else if (a == "--version") { std::cout << "addapt (B200 engine) 0.2" << std::endl; return 0; }
      else if (a == "-n" || a == "--num-moves") value(&num_moves);
      else if (a == "-T" || a == "--temperature") { value(&schedule); have_schedule = true; }
      else if (a == "-r" || a == "--random-seed") value(&seed);
      else if (a == "-o" || a == "--output") value(&output);
      else if (a == "-i" || a == "--output-interval") value(&interval);
      else if (a == "--trajectories") value(&n_traj);
      else if (a == "--device") value(&dev);
      else if (a.size() > 1 && a[0] == '-') throw "unknown option " + a;
      else config_files.push_back(a);
    }
    if (config_files.empty()) { std::cerr << USAGE; return 1; }

    DevicePtr device = device_from_yaml(config_files);
    auto scorefxn = std::dynamic_pointer_cast<B200ScoreFunction>(scorefxn_from_yaml(config_files));
    if (!scorefxn) throw string("the configured score function cannot run on the GPU engine");
    ThermostatPtr thermostat = have_schedule ? thermostat_from_str(schedule) : thermostat_from_yaml(config_files);
    default_engine(make_shared<B200Engine>(std::stoi(dev)));

    BatchedMonteCarlo sampler(scorefxn, std::stoi(num_moves), thermostat);
    sampler.trajectory_files(output, std::stoi(interval));
    const int k = std::stoi(n_traj);
    vector<BatchedTrajectory> runs = sampler.apply(device, k, (uint64_t)std::stoull(seed), 0);
    for (int t = 0; t < k; t++)
      std::cout << "trajectory " << t << ": best score " << runs[t].best_score << " at step " << runs[t].best_step << "  " << runs[t].best_seq
                << "  -> " << BatchedMonteCarlo::trajectory_path(output, (uint64_t)t, k) << std::endl;
    return 0;
  } catch (string error_message) {
    std::cerr << "Error: " << error_message << std::endl;
    return 1;
  } catch (const char* error_message) {
    std::cerr << "Error: " << error_message << std::endl;
    return 1;
  } catch (std::exception& e) {
    std::cerr << "Error: " << e.what() << std::endl;
    return 1;
  }
}
