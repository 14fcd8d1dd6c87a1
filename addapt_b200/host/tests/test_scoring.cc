// The reference's tests/test_scoring.cc: the CPU-side cases use a fake RnaFold (DummyRnaFold, :12-43); the cases
// that the reference runs against ViennaRNA (:83-259) run here against B200RnaFold on the GPU (--gpu).
#include <cmath>
#include <fstream>
#include <set>

#include "addapt/b200.hh"
#include "addapt/scoring.hh"
#include "check.hh"

using namespace addapt;
using bp = pair<int, int>;

namespace {

class DummyRnaFold : public RnaFold {
 public:
  DummyRnaFold(double p = 0) : my_macrostate_prob(p) {}
  double& operator[](bp key) { return my_probs[{std::min(key.first, key.second), std::max(key.first, key.second)}]; }
  double base_pair_prob(int a, int b) const {
    auto it = my_probs.find({std::min(a, b), std::max(a, b)});
    return it == my_probs.end() ? 0.0 : it->second;
  }
  double macrostate_prob(string) const { return my_macrostate_prob; }

 private:
  map<bp, double> my_probs;
  double my_macrostate_prob;
};

FoldFactory dummy_factory(double p = 0) {
  return [p](DeviceConstPtr, AptamerConstPtr) -> RnaFoldPtr { return make_shared<DummyRnaFold>(p); };
}

AptamerConstPtr theo_aptamer() { return make_shared<Aptamer>("GAUACCAGCCGAAAGGCCCUUGGCAGC", "(...((.(((....)))....))...)", 0.32); }

DevicePtr rhf_6_device() {
  DevicePtr d = make_shared<Device>("guuuuagagcuagaaauagcaaguuaaaauaaggcuaguccCuUUUCGCCgauaccagccgaaaggcccuuggcagcGACggcaccgagucggugcuuuuuu");
  d->add_macrostate("active", "(............................)xx..xxxxx..xxxxxx(...............................).(.............)......");
  return d;
}

}  // namespace

TEST_CASE(scoring_dummy_fold, "DummyRnaFold (reference tests/test_scoring.cc:66-81)") {
  DummyRnaFold fold;
  CHECK(fold.base_pair_prob(1, 2) == 0.0);
  fold[{1, 2}] = 0.75;
  CHECK(fold.base_pair_prob(1, 2) == 0.75);
  CHECK(fold.base_pair_prob(2, 1) == 0.75);
  fold[{2, 1}] = 0.50;
  CHECK(fold.base_pair_prob(1, 2) == 0.50);
}

TEST_CASE(scoring_function, "ScoreFunction sums and weights (reference tests/test_scoring.cc:261-300)") {
  class DummyTerm : public ScoreTerm {
   public:
    DummyTerm(double score, double weight) : ScoreTerm("dummy", weight), my_score(score) {}
    double evaluate(DeviceConstPtr, RnaFold const&, RnaFold const&) const { return my_score; }
   private:
    double my_score;
  };
  DevicePtr device = make_shared<Device>("UUUU");
  {
    ScoreFunction fx;
    fx.fold_factory(dummy_factory());
    CHECK(approx(fx.evaluate(device), 0));
    fx += make_shared<DummyTerm>(10, 1);
    CHECK(approx(fx.evaluate(device), 10));
    fx += make_shared<DummyTerm>(5, 1);
    CHECK(approx(fx.evaluate(device), 15));
  }
  {
    ScoreFunction fx;
    fx.fold_factory(dummy_factory());
    fx += make_shared<DummyTerm>(10, 1);
    fx += make_shared<DummyTerm>(10, 0.5);
    EvaluatedScoreFunction table;
    CHECK(approx(fx.evaluate(device, table), 15));
    CHECK(table.size() == 2 && table[1].name == "dummy" && table[1].weight == 0.5 && table[1].term == 10);
  }
}

TEST_CASE(scoring_contexts, "the score is the SUM over contexts (reference tests/test_scoring.cc:302-330)") {
  class LenTerm : public ScoreTerm {
   public:
    double evaluate(DeviceConstPtr device, RnaFold const&, RnaFold const&) const { return device->len(); }
  };
  ScoreFunction fx;
  fx.fold_factory(dummy_factory());
  fx += make_shared<LenTerm>();
  DevicePtr device = make_shared<Device>("U");
  CHECK(approx(fx.evaluate(device), 1.0));
  fx.add_context("1", make_shared<Context>("a", ""));
  CHECK(approx(fx.evaluate(device), 2.0));
  fx.add_context("2", make_shared<Context>("a", "a"));
  CHECK(approx(fx.evaluate(device), 5.0));
  fx.add_context("3", make_shared<Context>("a", "aa"));
  EvaluatedScoreFunction table;
  CHECK(approx(fx.evaluate(device, table), 9.0));
  CHECK(table.size() == 3 && table[0].name == "1: " && table[2].term == 4);
}

TEST_CASE(scoring_macrostate_term, "MacrostateProbTerm names and values (reference tests/test_scoring.cc:332-384)") {
  DevicePtr device = make_shared<Device>("");
  device->add_macrostate("dummy", "");
  struct Row { string name; ConditionEnum condition; FavorableEnum favorable; double apo, holo, expected; };
  const vector<Row> table = {
      {"apo: not dummy", ConditionEnum::APO, FavorableEnum::NO, 0.2, 0.2, std::log(0.8)},
      {"apo: not dummy", ConditionEnum::APO, FavorableEnum::NO, 0.9, 0.2, std::log(0.1)},
      {"apo: not dummy", ConditionEnum::APO, FavorableEnum::NO, 0.2, 0.9, std::log(0.8)},
      {"apo: not dummy", ConditionEnum::APO, FavorableEnum::NO, 0.9, 0.9, std::log(0.1)},
      {"apo: dummy", ConditionEnum::APO, FavorableEnum::YES, 0.2, 0.2, std::log(0.2)},
      {"apo: dummy", ConditionEnum::APO, FavorableEnum::YES, 0.9, 0.2, std::log(0.9)},
      {"apo: dummy", ConditionEnum::APO, FavorableEnum::YES, 0.2, 0.9, std::log(0.2)},
      {"apo: dummy", ConditionEnum::APO, FavorableEnum::YES, 0.9, 0.9, std::log(0.9)},
      {"holo: not dummy", ConditionEnum::HOLO, FavorableEnum::NO, 0.2, 0.2, std::log(0.8)},
      {"holo: not dummy", ConditionEnum::HOLO, FavorableEnum::NO, 0.9, 0.2, std::log(0.8)},
      {"holo: not dummy", ConditionEnum::HOLO, FavorableEnum::NO, 0.2, 0.9, std::log(0.1)},
      {"holo: not dummy", ConditionEnum::HOLO, FavorableEnum::NO, 0.9, 0.9, std::log(0.1)},
      {"holo: dummy", ConditionEnum::HOLO, FavorableEnum::YES, 0.2, 0.2, std::log(0.2)},
      {"holo: dummy", ConditionEnum::HOLO, FavorableEnum::YES, 0.9, 0.2, std::log(0.2)},
      {"holo: dummy", ConditionEnum::HOLO, FavorableEnum::YES, 0.2, 0.9, std::log(0.9)},
      {"holo: dummy", ConditionEnum::HOLO, FavorableEnum::YES, 0.9, 0.9, std::log(0.9)},
  };
  for (Row const& row : table) {
    DummyRnaFold apo(row.apo), holo(row.holo);
    MacrostateProbTerm term("dummy", row.condition, row.favorable);
    CHECK(term.name() == row.name);
    CHECK(approx(term.evaluate(device, apo, holo), row.expected));
  }
}

// ---- GPU: the cases the reference runs against ViennaRNA -------------------------------------------------------

GPU_TEST_CASE(gpu_hairpin, "B200RnaFold: hairpin without an aptamer (reference tests/test_scoring.cc:83-147)") {
  DevicePtr hairpin = make_shared<Device>("ACGUGAAAACGU");
  const std::set<bp> expected = {{0, 11}, {1, 10}, {2, 9}, {3, 8}};
  map<bp, double> thresholds = {{{0, 11}, 0.70}, {{1, 10}, 0.95}, {{2, 9}, 0.95}, {{3, 8}, 0.95}};
  B200RnaFold fold(hairpin);
  for (int i = 0; i < hairpin->len(); i++)
    for (int j = i; j < hairpin->len(); j++) {
      const double p = fold.base_pair_prob(i, j);
      if (expected.count({i, j})) CHECK(p > thresholds[{i, j}]); else REQUIRE(p < 0.1);
      REQUIRE(p == fold.base_pair_prob(j, i));
    }
  CHECK(fold.base_pair_prob(-1, 0) == fold.base_pair_prob(0, 11));
  CHECK_THROWS(fold.base_pair_prob(0, 12));
  CHECK_THROWS(fold.base_pair_prob(0, -13));
  CHECK(fold.macrostate_prob("...(....)...") > 0.95);
  CHECK(fold.macrostate_prob("..((....))..") > 0.85);
  CHECK(fold.macrostate_prob(".(((....))).") > 0.75);
  CHECK(fold.macrostate_prob("((((....))))") > 0.65);
  CHECK(fold.macrostate_prob("(((......)))") > 0.65);
  CHECK(fold.macrostate_prob("((........))") > 0.65);
  CHECK(fold.macrostate_prob("(..........)") > 0.65);
  CHECK(fold.macrostate_prob("xxxxxxxxxxxx") < 0.05);
  CHECK(fold.macrostate_prob("xxxx........") < 0.05);
  CHECK(fold.macrostate_prob("....xxxx....") > 0.95);
  CHECK(fold.macrostate_prob("........xxxx") < 0.05);
  CHECK_THROWS(fold.macrostate_prob("(((....."));     // wrong length / unbalanced
}

GPU_TEST_CASE(gpu_aptamer, "B200RnaFold: a device that is just the aptamer, apo vs holo (reference tests/test_scoring.cc:149-195)") {
  DevicePtr hairpin = make_shared<Device>("GAUACCAGCCGAAAGGCCCUUGGCAGC");
  const std::set<bp> apo_pairs = {{4, 22}, {5, 21}, {6, 20}, {7, 16}, {8, 15}, {9, 14}};
  const std::set<bp> holo_pairs = {{0, 26}, {4, 22}, {5, 21}, {7, 16}, {8, 15}, {9, 14}};
  B200RnaFold apo(hairpin), holo(hairpin, theo_aptamer());
  for (int i = 0; i < hairpin->len(); i++)
    for (int j = i; j < hairpin->len(); j++) {
      const double pa = apo.base_pair_prob(i, j), ph = holo.base_pair_prob(i, j);
      if (apo_pairs.count({i, j})) CHECK(pa > 0.7); else REQUIRE(pa < 0.3);
      if (holo_pairs.count({i, j})) CHECK(ph > 0.7); else REQUIRE(ph < 0.3);
    }
  CHECK(apo.macrostate_prob("....((((((....)))...)))....") > 0.65);
  CHECK(apo.macrostate_prob("(.........................)") < 0.05);
  CHECK(holo.macrostate_prob("....((((((....)))...)))....") < 0.01);
  CHECK(holo.macrostate_prob("(.........................)") > 0.95);
}

GPU_TEST_CASE(gpu_rhf6, "B200RnaFold: rhf(6) sgRNA (reference tests/test_scoring.cc:197-259)") {
  DevicePtr rhf_6 = rhf_6_device();
  const map<bp, double> constitutive = {{{0, 29}, 0.75}, {{1, 28}, 0.80}, {{2, 27}, 0.85}, {{3, 26}, 0.85}, {{4, 25}, 0.85}, {{5, 24}, 0.85},
                                        {{6, 23}, 0.80}, {{8, 19}, 0.95}, {{9, 18}, 0.95}, {{10, 17}, 0.95}, {{11, 16}, 0.95}, {{81, 95}, 0.90},
                                        {{82, 94}, 0.90}, {{83, 93}, 0.90}, {{84, 92}, 0.90}, {{85, 91}, 0.90}, {{86, 90}, 0.75}};
  const map<bp, double> apo_only = {{{33, 73}, 0.30}, {{34, 72}, 0.30}, {{35, 71}, 0.30}, {{36, 70}, 0.25}, {{37, 69}, 0.10}, {{40, 65}, 0.35},
                                    {{41, 64}, 0.35}, {{42, 63}, 0.35}, {{43, 62}, 0.35}, {{44, 61}, 0.30}, {{45, 60}, 0.25}, {{47, 58}, 0.35},
                                    {{48, 57}, 0.40}};
  const map<bp, pair<double, double>> holo_only = {{{30, 43}, {0.55, 0.25}}, {{31, 42}, {0.75, 0.35}}, {{32, 41}, {0.75, 0.35}},
                                                   {{33, 40}, {0.75, 0.35}}, {{46, 80}, {0.50, 0.05}}, {{47, 79}, {0.60, 0.05}},
                                                   {{49, 77}, {0.85, 0.05}}, {{50, 76}, {0.95, 0.05}}, {{54, 72}, {0.95, 0.20}},
                                                   {{55, 71}, {0.95, 0.20}}, {{57, 66}, {0.95, 0.50}}, {{58, 65}, {0.95, 0.50}},
                                                   {{59, 64}, {0.95, 0.50}}};
  B200RnaFold apo(rhf_6), holo(rhf_6, theo_aptamer());
  for (auto const& kv : constitutive) {
    REQUIRE(apo.base_pair_prob(kv.first.first, kv.first.second) > kv.second);
    REQUIRE(holo.base_pair_prob(kv.first.first, kv.first.second) > kv.second);
  }
  for (auto const& kv : apo_only) {
    REQUIRE(apo.base_pair_prob(kv.first.first, kv.first.second) > kv.second);
    REQUIRE(holo.base_pair_prob(kv.first.first, kv.first.second) < 1e-3);
  }
  for (auto const& kv : holo_only) {
    CHECK(apo.base_pair_prob(kv.first.first, kv.first.second) < kv.second.second);
    CHECK(holo.base_pair_prob(kv.first.first, kv.first.second) > kv.second.first);
  }
  CHECK(apo.macrostate_prob(rhf_6->macrostate("active")) < 7e-5);
  CHECK(holo.macrostate_prob(rhf_6->macrostate("active")) > 4e-3);
}

GPU_TEST_CASE(gpu_score_paths, "ScoreFunction through RnaFold objects == B200ScoreFunction batch == per-term definition") {
  DevicePtr rhf_6 = rhf_6_device();
  auto batched = make_shared<B200ScoreFunction>();
  batched->aptamer(theo_aptamer());
  batched->add_term(make_shared<MacrostateProbTerm>("active", ConditionEnum::APO, FavorableEnum::NO));
  batched->add_term(make_shared<MacrostateProbTerm>("active", ConditionEnum::HOLO, FavorableEnum::YES));
  ScoreFunction generic;                                   // default factory: B200RnaFold per object
  generic.aptamer(theo_aptamer());
  for (auto const& t : batched->terms()) generic.add_term(t);
  EvaluatedScoreFunction ta, tb;
  const double sa = generic.evaluate(rhf_6, ta), sb = batched->evaluate(rhf_6, tb);
  CHECK(approx(sa, sb, 1e-6));
  CHECK(ta.size() == 2 && tb.size() == 2 && ta[0].name == "apo: not active" && tb[1].name == "holo: active");
  CHECK(approx(ta[1].term, tb[1].term, 1e-6));
  CHECK(sb > std::log(4e-3) - 1e-3 && sb < 0);             // log(1 - P_apo) + log(P_holo), thresholds of the rhf(6) test
  // contexts: sum over contexts, rows prefixed
  batched->add_context("none", make_shared<Context>("", ""));
  batched->add_context("rfp", make_shared<Context>("AACTTTCAGTTTAGCGGTCT", ""));
  generic.add_context("none", make_shared<Context>("", ""));
  generic.add_context("rfp", make_shared<Context>("AACTTTCAGTTTAGCGGTCT", ""));
  const double ca = generic.evaluate(rhf_6, ta), cb = batched->evaluate(rhf_6, tb);
  CHECK(approx(ca, cb, 1e-6));
  CHECK(tb.size() == 4 && tb[0].name == "none: apo: not active" && tb[3].name == "rfp: holo: active");
  CHECK(approx(tb[0].term + tb[1].term, sb, 1e-6));
  vector<double> many = batched->evaluate_batch(rhf_6, {rhf_6->raw_seq(), rhf_6->raw_seq()});
  CHECK(many.size() == 2 && many[0] == many[1] && approx(many[0], cb, 1e-12));
}

GPU_TEST_CASE(gpu_batched_mc, "BatchedMonteCarlo: sharding-invariant trajectories, outcome accounting, improvement") {
  DevicePtr rhf_6 = rhf_6_device();
  auto fx = make_shared<B200ScoreFunction>();
  fx->aptamer(theo_aptamer());
  fx->add_term(make_shared<MacrostateProbTerm>("active", ConditionEnum::APO, FavorableEnum::NO));
  fx->add_term(make_shared<MacrostateProbTerm>("active", ConditionEnum::HOLO, FavorableEnum::YES));
  const double start = fx->evaluate(rhf_6);
  BatchedMonteCarlo mc(fx, 60, make_shared<AnnealingThermostat>(300, 5.0, 0.0));
  auto whole = mc.apply(rhf_6, 6, 2016, 0);
  auto tail = mc.apply(rhf_6, 3, 2016, 3);
  CHECK(whole.size() == 6);
  for (int t = 0; t < 3; t++) {
    CHECK(whole[3 + t].final_seq == tail[t].final_seq);
    CHECK(whole[3 + t].best_score == tail[t].best_score);
  }
  for (auto const& tr : whole) {
    int total = 0;
    for (auto const& kv : tr.outcome_counters) total += kv.second;
    CHECK(total == 60);
    CHECK(tr.best_score >= start - 1e-12);
    CHECK(tr.best_score >= tr.final_score - 1e-12);
    CHECK(approx(fx->evaluate_batch(rhf_6, {tr.best_seq})[0], tr.best_score, 1e-6));
    for (size_t p = 0; p < tr.final_seq.size(); p++)
      if (std::islower((unsigned char)rhf_6->raw_seq()[p])) CHECK(rhf_6->raw_seq()[p] == tr.final_seq[p]);     // fixed positions stay as given
  }
  // the auto-scaling thermostat runs on the device too; an unknown thermostat class is refused
  auto scaled = BatchedMonteCarlo(fx, 60, make_shared<AutoScalingThermostat>(0.5, 20, 1.0)).apply(rhf_6, 2, 7);
  CHECK(scaled.size() == 2 && scaled[0].best_score >= start - 1e-12);
  class OddThermostat : public Thermostat { double adjust(MonteCarloStep const&) { return 1; } };
  CHECK_THROWS(BatchedMonteCarlo(fx, 10, make_shared<OddThermostat>()).apply(rhf_6, 1, 0));
}

TEST_CASE(scoring_bp_terms, "PairProbTerm / PairedDensityTerm over a fake base_pair_prob (interface: reference include/scoring.hh:48)") {
  DevicePtr device = make_shared<Device>("ACGUACGU");
  auto apo = make_shared<DummyRnaFold>(), holo = make_shared<DummyRnaFold>();
  (*apo)[{0, 7}] = 0.25;
  (*holo)[{0, 7}] = 0.9;
  (*holo)[{1, 6}] = 0.5;
  (*holo)[{1, 5}] = 0.25;
  PairProbTerm want(0, 7, ConditionEnum::HOLO), unwanted(7, 0, ConditionEnum::APO, FavorableEnum::NO);
  CHECK(want.name() == "holo: pair 0-7" && unwanted.name() == "apo: not pair 7-0");
  CHECK(approx(want.evaluate(device, *apo, *holo), std::log(0.9)));
  CHECK(approx(unwanted.evaluate(device, *apo, *holo), std::log(0.75)));
  PairedDensityTerm dens({0, 1, 2}, ConditionEnum::HOLO);
  CHECK(dens.name() == "holo: paired density 0,1,2");
  CHECK(approx(dens.evaluate(device, *apo, *holo), std::log((0.9 + 0.75 + 0.0) / 3)));
  PairedDensityTerm open_({1}, ConditionEnum::HOLO, FavorableEnum::NO);
  CHECK(approx(open_.evaluate(device, *apo, *holo), std::log(0.25)));
  // a context shifts the raw positions by its 5' flank
  DevicePtr flanked = device->copy();
  flanked->context(make_shared<Context>("GG", "C"));
  auto shifted = make_shared<DummyRnaFold>();
  (*shifted)[{2, 9}] = 0.6;
  CHECK(approx(want.evaluate(flanked, *apo, *shifted), std::log(0.6)));
  CHECK_THROWS(PairProbTerm(0, 8, ConditionEnum::APO).evaluate(device, *apo, *holo));
  CHECK_THROWS(PairedDensityTerm({}, ConditionEnum::APO).evaluate(device, *apo, *holo));
  // through a score function: weights and table rows
  ScoreFunction fx;
  fx.fold_factory([&](DeviceConstPtr, AptamerConstPtr apt) -> RnaFoldPtr { return apt ? holo : apo; });
  fx.aptamer(theo_aptamer());
  auto t1 = make_shared<PairProbTerm>(0, 7, ConditionEnum::HOLO);
  t1->weight(2.0);
  fx.add_term(t1);
  fx.add_term(make_shared<PairProbTerm>(0, 7, ConditionEnum::APO, FavorableEnum::NO));
  EvaluatedScoreFunction table;
  CHECK(approx(fx.evaluate(device, table), 2 * std::log(0.9) + std::log(0.75)));
  CHECK(table.size() == 2 && table[0].name == "holo: pair 0-7" && table[0].weight == 2.0);
}

GPU_TEST_CASE(gpu_bp_terms, "bp terms: batched device path == per-object path over B200RnaFold::base_pair_prob") {
  DevicePtr rhf_6 = rhf_6_device();
  auto batched = make_shared<B200ScoreFunction>();
  batched->float_energies(false);
  batched->aptamer(theo_aptamer());
  batched->add_term(make_shared<MacrostateProbTerm>("active", ConditionEnum::HOLO, FavorableEnum::YES));
  batched->add_term(make_shared<PairProbTerm>(47, 79, ConditionEnum::HOLO));
  batched->add_term(make_shared<PairProbTerm>(47, 79, ConditionEnum::APO, FavorableEnum::NO));
  batched->add_term(make_shared<PairedDensityTerm>(vector<int>{30, 31, 32, 33, 40, 41, 42, 43}, ConditionEnum::APO));
  ScoreFunction generic;
  generic.aptamer(theo_aptamer());
  for (auto const& t : batched->terms()) generic.add_term(t);
  EvaluatedScoreFunction ta, tb;
  const double sa = generic.evaluate(rhf_6, ta), sb = batched->evaluate(rhf_6, tb);
  CHECK(ta.size() == 4 && tb.size() == 4);
  for (size_t k = 1; k < 4 && k < ta.size() && k < tb.size(); k++) {      // row 0 goes through float-rounded energies in the generic path
    CHECK(ta[k].name == tb[k].name);
    CHECK(approx(ta[k].term, tb[k].term, 1e-6));
  }
  CHECK(approx(sa - ta[0].term, sb - tb[0].term, 1e-6));
  CHECK(tb[1].term > std::log(0.5));                       // the holo-only pair of reference tests/test_scoring.cc:219 forms with the ligand
  // with contexts the batched path shifts the positions like the per-object path does
  batched->add_context("rfp", make_shared<Context>("AACTTTCAGTTTAGCGGTCT", ""));
  generic.add_context("rfp", make_shared<Context>("AACTTTCAGTTTAGCGGTCT", ""));
  generic.evaluate(rhf_6, ta);
  batched->evaluate(rhf_6, tb);
  CHECK(tb.size() == 4 && tb[1].name == "rfp: holo: pair 47-79");
  for (size_t k = 1; k < 4 && k < ta.size() && k < tb.size(); k++) CHECK(approx(ta[k].term, tb[k].term, 1e-6));
}

GPU_TEST_CASE(gpu_device_with_own_context, "a device that carries a context is scored with its flanks (reference src/scoring.cc:136-152)") {
  DevicePtr rhf_6 = rhf_6_device();
  auto batched = make_shared<B200ScoreFunction>();
  batched->aptamer(theo_aptamer());
  batched->add_term(make_shared<MacrostateProbTerm>("active", ConditionEnum::APO, FavorableEnum::NO));
  batched->add_term(make_shared<MacrostateProbTerm>("active", ConditionEnum::HOLO, FavorableEnum::YES));
  ScoreFunction generic;
  generic.aptamer(theo_aptamer());
  for (auto const& t : batched->terms()) generic.add_term(t);
  const double bare = batched->evaluate(rhf_6);
  DevicePtr flanked = rhf_6->copy();
  flanked->context(make_shared<Context>("GGGGCCACTAGGGACAGGAT", ""));
  EvaluatedScoreFunction ta, tb;
  const double sa = generic.evaluate(flanked, ta), sb = batched->evaluate(flanked, tb);
  CHECK(approx(sa, sb, 1e-4));
  CHECK(!approx(sb, bare, 1e-6));
  CHECK(tb.size() == 2 && tb[0].name == "apo: not active" && tb[1].name == "holo: active");
}

GPU_TEST_CASE(gpu_batched_tsv, "BatchedMonteCarlo writes TsvTrajectoryReporter's file from the kernel's records (reference src/sampling.cc:430-489)") {
  DevicePtr rhf_6 = rhf_6_device();
  auto fx = make_shared<B200ScoreFunction>();
  fx->aptamer(theo_aptamer());
  fx->add_term(make_shared<MacrostateProbTerm>("active", ConditionEnum::APO, FavorableEnum::NO));
  fx->add_term(make_shared<MacrostateProbTerm>("active", ConditionEnum::HOLO, FavorableEnum::YES));
  BatchedMonteCarlo mc(fx, 25, make_shared<AnnealingThermostat>(10, 5.0, 0.0));
  const string path = "/tmp/addapt_b200_host_test_traj.tsv";
  mc.trajectory_files(path, 2);
  auto runs = mc.apply(rhf_6, 2, 42, 0);
  CHECK(BatchedMonteCarlo::trajectory_path("traj.tsv", 17, 4) == "traj.17.tsv");
  CHECK(BatchedMonteCarlo::trajectory_path("out/t%d.tsv", 3, 4) == "out/t3.tsv");
  CHECK(BatchedMonteCarlo::trajectory_path("traj.tsv", 0, 1) == "traj.tsv");
  for (int t = 0; t < 2; t++) {
    std::ifstream in(BatchedMonteCarlo::trajectory_path(path, t, 2));
    CHECK(in.is_open());
    string line;
    std::getline(in, line);
    CHECK(line == "#\tinitial_seq\t" + rhf_6->seq());
    std::getline(in, line);
    CHECK(line.rfind("step\tnum_steps\tcurrent_score\tproposed_score\tterm_weight[apo: not active]\tterm_value[apo: not active]\t", 0) == 0);
    CHECK(line.size() > 30 && line.substr(line.size() - 25) == "current_seq\tproposed_seq\t");
    int rows = 0;
    string last;
    while (std::getline(in, line)) { rows++; last = line; }
    CHECK(rows == 13);                                     // steps 0, 2, ..., 24
    CHECK(last.rfind("24\t25\t", 0) == 0);
    CHECK(last.find("\tUnbiasedMutation\t") != string::npos);
    // the current sequence of the last record is the final sequence when the record is the last step
    CHECK(last.find("\t" + runs[t].final_seq + "\t") != string::npos);
  }
}
