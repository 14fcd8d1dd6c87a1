// The reference's tests/test_sampling.cc cases (mutability rules, the mutate_recursively known-answer table, its
// error cases) plus checks of the pieces the reference leaves untested: thermostats, the thermostat / objective
// mini-languages, MonteCarlo::apply with a fake score function, the TSV trajectory layout.
#include <cstdio>
#include <fstream>

#include "addapt/config.hh"
#include "addapt/sampling.hh"
#include "check.hh"

using namespace addapt;

TEST_CASE(sampling_mutable, "mutable / freely mutable positions (reference tests/test_sampling.cc:8-45)") {
  {
    DevicePtr device = make_shared<Device>("UUUuuu");
    device->add_macrostate("a", "(.)(.)");
    const bool mut[] = {true, true, true, false, false, false}, freely[] = {true, true, false, false, false, false};
    for (int k = 0; k < 6; k++) { CHECK(can_be_mutated(device, k) == mut[k]); CHECK(can_be_freely_mutated(device, k) == freely[k]); }
  }
  {
    DevicePtr device = make_shared<Device>("UUUU");
    device->add_macrostate("a", "().)");
    device->add_macrostate("b", "(.))");
    const bool freely[] = {true, false, false, false};
    for (int k = 0; k < 4; k++) { CHECK(can_be_mutated(device, k) == true); CHECK(can_be_freely_mutated(device, k) == freely[k]); }
  }
}

TEST_CASE(sampling_recursive, "mutate_recursively known answers (reference tests/test_sampling.cc:47-99)") {
  struct Row { string sequence; vector<string> macrostates; string mutations; vector<string> expected; };
  const vector<Row> table = {
      {"N", {"."}, "A", {"A"}},
      {"NN", {"()"}, "AG", {"AU", "CG"}},
      {"NNN", {"(.)"}, "AGU", {"ANU", "NGN", "ANU"}},
      {"NN", {"()", "()"}, "AG", {"AU", "CG"}},
      {"NNN", {"().", ".()"}, "AGU", {"AUA", "CGC", "UAU"}},
      {"NNN", {"().", "(.)"}, "AGU", {"AUU", "CGG", "AUU"}},
      {"NNNN", {"()..", "(())"}, "AGUC", {"AUAU", "CGCG", "UAUA", "GCGC"}},
      {"NNNN", {"(.).", "(())"}, "AGUC", {"AAUU", "GGCC", "AAUU", "GGCC"}},
  };
  for (Row const& row : table)
    for (size_t i = 0; i < row.mutations.size(); i++) {
      DevicePtr device = make_shared<Device>(row.sequence);
      for (size_t x = 0; x < row.macrostates.size(); x++) device->add_macrostate(std::to_string(x), row.macrostates[x]);
      mutate_recursively(device, (int)i, row.mutations[i]);
      CHECK(device->seq() == row.expected[i]);
    }
}

TEST_CASE(sampling_recursive_errors, "mutate_recursively error cases (reference tests/test_sampling.cc:101-112)") {
  DevicePtr device = make_shared<Device>("Nn");
  device->add_macrostate("not mutable", "()");
  CHECK_THROWS(mutate_recursively(device, 0, 'G'));
  device = make_shared<Device>("N");
  device->add_macrostate("extra open", "(");
  CHECK_THROWS(mutate_recursively(device, 0, 'G'));
  device = make_shared<Device>("N");
  device->add_macrostate("extra close", ")");
  CHECK_THROWS(mutate_recursively(device, 0, 'G'));
}

TEST_CASE(sampling_thermostats, "thermostats (reference src/sampling.cc:305-401)") {
  MonteCarloStep step;
  step.score_diff = 0;
  FixedThermostat fixed(2.5);
  step.i = 7;
  CHECK(fixed.adjust(step) == 2.5);
  AnnealingThermostat saw(300, 5.0, 0.0);
  step.i = 0;   CHECK(saw.adjust(step) == 5.0);
  step.i = 150; CHECK(approx(saw.adjust(step), 2.5));
  step.i = 299; CHECK(approx(saw.adjust(step), 5.0 - 5.0 * 299 / 300));     // never reaches T_lo
  step.i = 300; CHECK(saw.adjust(step) == 5.0);
  AutoScalingThermostat autoT(0.5, 4, 1.0);
  const double diffs[] = {-1.0, -3.0, -2.0, -4.0};
  double T = 0;
  for (int k = 0; k < 4; k++) { step.score_diff = diffs[k]; T = autoT.adjust(step); if (k < 3) CHECK(T == 1.0); }
  CHECK(approx(T, -2.0 / std::log(0.5)));       // nth_element at n/2 of {-4,-3,-2,-1} is -2
}

TEST_CASE(config_strings, "thermostat_from_str / score_term_from_str (reference src/config.cc:104-175)") {
  auto fixed = std::dynamic_pointer_cast<FixedThermostat>(thermostat_from_str("0.5"));
  CHECK(fixed && fixed->temperature() == 0.5);
  auto saw = std::dynamic_pointer_cast<AnnealingThermostat>(thermostat_from_str("5 to 0 in 300 steps"));
  CHECK(saw && saw->cycle_len() == 300 && saw->max_temperature() == 5 && saw->min_temperature() == 0);
  CHECK(std::dynamic_pointer_cast<AutoScalingThermostat>(thermostat_from_str("auto")) != nullptr);
  CHECK(std::dynamic_pointer_cast<AutoScalingThermostat>(thermostat_from_str("auto 30% 50 2.0")) != nullptr);
  CHECK_THROWS(thermostat_from_str("hot"));
  auto term = score_term_from_str(ConditionEnum::APO, "not active");
  CHECK(term->name() == "apo: not active");
  CHECK(score_term_from_str(ConditionEnum::HOLO, "active")->name() == "holo: active");
  CHECK_THROWS(score_term_from_str(ConditionEnum::APO, "not very active"));
}

namespace {
// score = number of G's; no folding
class CountG : public ScoreFunction {
 public:
  using ScoreFunction::evaluate;
  double evaluate(DeviceConstPtr device, EvaluatedScoreFunction& table) const override {
    table.clear();
    double g = 0;
    for (char ch : device->seq()) g += ch == 'G';
    table.push_back({"count_g", 1.0, g});
    return g;
  }
};
}  // namespace

TEST_CASE(sampling_monte_carlo, "MonteCarlo::apply: greedy at T -> 0, outcome accounting, TSV layout (reference src/sampling.cc:22-107,430-489)") {
  DevicePtr device = make_shared<Device>("AAAAaaaa");
  device->add_macrostate("free", "........");
  MonteCarlo mc;
  mc.num_steps(400);
  mc.scorefxn(make_shared<CountG>());
  mc.thermostat(make_shared<FixedThermostat>(1e-9));
  mc += make_shared<UnbiasedMutationMove>();
  const string path = "/tmp/addapt_b200_host_test.tsv";
  mc += make_shared<TsvTrajectoryReporter>(path, 100);
  std::mt19937 rng(7);
  DevicePtr out = mc.apply(device, rng);
  CHECK(out->seq() == "GGGGaaaa");                 // only upper-case positions move; maximisation
  CHECK(device->seq() == "AAAAaaaa");              // the input device is not modified
  std::ifstream tsv(path);
  string line;
  std::getline(tsv, line);
  CHECK(line == "#\tinitial_seq\tAAAAaaaa");
  std::getline(tsv, line);
  CHECK(line == "step\tnum_steps\tcurrent_score\tproposed_score\tterm_weight[count_g]\tterm_value[count_g]\tscore_diff\ttemperature\t"
                "metropolis_criterion\trandom_threshold\tmove\toutcome\tcurrent_seq\tproposed_seq\t");
  int rows = 0;
  while (std::getline(tsv, line)) rows += !line.empty();
  CHECK(rows == 4);
  std::remove(path.c_str());

  MonteCarlo idle;                                   // no moves: returns the device untouched
  idle.num_steps(10);
  CHECK(idle.apply(device, rng) == device);
}

TEST_CASE(config_pair_objective, "bp_term_from_str: the optional pair_objective mini-language") {
  auto a = std::dynamic_pointer_cast<PairProbTerm>(bp_term_from_str("holo pair 47 79"));
  CHECK(a && a->i() == 47 && a->j() == 79 && a->condition() == ConditionEnum::HOLO && a->favorable() == FavorableEnum::YES && a->weight() == 1.0);
  auto b = std::dynamic_pointer_cast<PairProbTerm>(bp_term_from_str("apo not pair 47 79 x 2.5"));
  CHECK(b && b->favorable() == FavorableEnum::NO && b->condition() == ConditionEnum::APO && b->weight() == 2.5);
  auto c = std::dynamic_pointer_cast<PairedDensityTerm>(bp_term_from_str("apo density 30-33"));
  CHECK(c && c->positions() == vector<int>({30, 31, 32, 33}));
  CHECK_THROWS(bp_term_from_str("holo pairs 1 2"));
  CHECK_THROWS(bp_term_from_str("apo density 5-4"));
}

TEST_CASE(config_yaml, "device / score function / thermostat from YAML files (reference src/config.cc:24-122)") {
  const string a = "/tmp/addapt_b200_cfg_a.yml", b = "/tmp/addapt_b200_cfg_b.yml";
  {
    std::ofstream f(a);
    f << "# design problem\nsequence: acgUUUacgu\nmacrostates:\n  active: \"(........)\"\n  idle: ..........\n"
         "thermostat: 5 to 0 in 300 steps\n";
    std::ofstream g(b);
    g << "objective: {apo: not active, holo: active}\naptamer:\n  sequence: GAUACCAGCCGAAAGGCCCUUGGCAGC\n"
         "  fold: (...((.(((....)))....))...)\n  affinity: 0.32\ncontexts:\n  none: ['', '']\n  rfp: [AACTTT, \"\"]\n";
  }
  DevicePtr device = device_from_yaml({a, b});
  CHECK(device->seq() == "acgUUUacgu");
  CHECK(device->macrostate("active") == "(........)");
  CHECK(device->macrostate("idle") == "..........");
  ScoreFunctionPtr fx = scorefxn_from_yaml({a, b});
  CHECK(fx->terms().size() == 2);
  CHECK(fx->terms()[0]->name() == "apo: not active");
  CHECK(fx->terms()[1]->name() == "holo: active");
  CHECK(fx->aptamer()->affinity() == 0.32);
  CHECK(fx->context("rfp")->before() == "AACTTT");
  CHECK(fx->context("none")->after() == "");
  CHECK(std::dynamic_pointer_cast<AnnealingThermostat>(thermostat_from_yaml({a, b})) != nullptr);
  CHECK(std::dynamic_pointer_cast<FixedThermostat>(thermostat_from_yaml({b})) != nullptr);    // default "1"
  CHECK_THROWS(device_from_yaml({b}));               // no `sequence` section
  CHECK_THROWS(device_from_yaml({a, a}));            // duplicated section
  std::remove(a.c_str());
  std::remove(b.c_str());
}
