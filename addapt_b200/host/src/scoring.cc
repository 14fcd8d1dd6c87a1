#include "addapt/scoring.hh"

#include <cmath>
#include <ostream>

#include "addapt/b200.hh"

namespace addapt {

ScoreFunction::ScoreFunction() : my_factory(b200_fold_factory()) {}

double ScoreFunction::evaluate(DeviceConstPtr device) const {
  EvaluatedScoreFunction scratch;
  return evaluate(device, scratch);
}

double ScoreFunction::evaluate(DeviceConstPtr device, EvaluatedScoreFunction& table) const {
  table.clear();
  if (my_contexts.empty()) return evaluate_terms(device, table);
  double total = 0;
  DevicePtr flanked = device->copy();
  for (auto const& named : my_contexts) {        // std::map: name order, like the reference
    flanked->context(named.second);
    total += evaluate_terms(flanked, table, named.first + ": ");
  }
  return total;
}

double ScoreFunction::evaluate_terms(DeviceConstPtr device, EvaluatedScoreFunction& table, string prefix) const {
  double total = 0;
  if (my_terms.empty()) return total;          // nothing to fold for
  if (!my_factory) throw string("score function has terms but no fold engine");
  // one apo and one holo engine per (contextual) device, shared by all terms -- reference src/scoring.cc:145-146
  RnaFoldPtr apo = my_factory(device, nullptr);
  RnaFoldPtr holo = my_factory(device, my_aptamer);
  for (ScoreTermPtr const& term : my_terms) {
    EvaluatedScoreTerm row{prefix + term->name(), term->weight(), term->evaluate(device, *apo, *holo)};
    table.push_back(row);
    total += row.weight * row.term;
  }
  return total;
}

void ScoreFunction::add_term(ScoreTermPtr term) { my_terms.push_back(term); }
void ScoreFunction::operator+=(ScoreTermPtr term) { add_term(term); }
ScoreTermList ScoreFunction::terms() const { return my_terms; }
AptamerConstPtr ScoreFunction::aptamer() const { return my_aptamer; }
void ScoreFunction::aptamer(AptamerConstPtr aptamer) { my_aptamer = aptamer; }
ContextConstPtr ScoreFunction::context(string name) const { return my_contexts.at(name); }
void ScoreFunction::add_context(string name, ContextConstPtr ctx) { my_contexts[name] = ctx; }
map<string, ContextConstPtr> const& ScoreFunction::contexts() const { return my_contexts; }
void ScoreFunction::fold_factory(FoldFactory f) { my_factory = f; }
FoldFactory ScoreFunction::fold_factory() const { return my_factory; }

ScoreTerm::ScoreTerm(string name, double weight) : my_name(name), my_weight(weight) {}
string ScoreTerm::name() const { return my_name; }
void ScoreTerm::name(string name) { my_name = name; }
double ScoreTerm::weight() const { return my_weight; }
void ScoreTerm::weight(double weight) { my_weight = weight; }

MacrostateProbTerm::MacrostateProbTerm(string macrostate, ConditionEnum condition, FavorableEnum favorable)
    : ScoreTerm("active_macrostate"), my_macrostate(macrostate), my_condition(condition), my_favorable(favorable) {
  name(string(condition == ConditionEnum::APO ? "apo" : "holo") + ": " + (favorable == FavorableEnum::NO ? "not " : "") + macrostate);
}

double MacrostateProbTerm::evaluate(DeviceConstPtr device, RnaFold const& apo, RnaFold const& holo) const {
  RnaFold const& engine = my_condition == ConditionEnum::APO ? apo : holo;
  double p = engine.macrostate_prob(device->macrostate(my_macrostate));
  if (my_favorable == FavorableEnum::NO) p = 1 - p;
  return std::log(p);        // may be -inf, exactly like the reference (src/scoring.cc:256-258)
}

namespace {
string cond_prefix(ConditionEnum c, FavorableEnum f) { return string(c == ConditionEnum::APO ? "apo" : "holo") + ": " + (f == FavorableEnum::NO ? "not " : ""); }
int flank5(DeviceConstPtr device) { return device->context() ? (int)device->context()->before().size() : 0; }
}  // namespace

PairProbTerm::PairProbTerm(int i, int j, ConditionEnum condition, FavorableEnum favorable)
    : ScoreTerm("pair"), my_i(i), my_j(j), my_condition(condition), my_favorable(favorable) {
  name(cond_prefix(condition, favorable) + "pair " + std::to_string(i) + "-" + std::to_string(j));
}

double PairProbTerm::evaluate(DeviceConstPtr device, RnaFold const& apo, RnaFold const& holo) const {
  if (my_i < 0 || my_j < 0 || my_i >= device->raw_len() || my_j >= device->raw_len()) throw msg("pair term position out of range");
  RnaFold const& engine = my_condition == ConditionEnum::APO ? apo : holo;
  const int off = flank5(device);
  double p = engine.base_pair_prob(my_i + off, my_j + off);
  if (my_favorable == FavorableEnum::NO) p = 1 - p;
  return std::log(p);
}

PairedDensityTerm::PairedDensityTerm(vector<int> positions, ConditionEnum condition, FavorableEnum favorable)
    : ScoreTerm("density"), my_positions(positions), my_condition(condition), my_favorable(favorable) {
  string list;
  for (size_t k = 0; k < positions.size(); k++) list += (k ? "," : "") + std::to_string(positions[k]);
  name(cond_prefix(condition, favorable) + "paired density " + list);
}

double PairedDensityTerm::evaluate(DeviceConstPtr device, RnaFold const& apo, RnaFold const& holo) const {
  if (my_positions.empty()) throw string("density term without positions");
  RnaFold const& engine = my_condition == ConditionEnum::APO ? apo : holo;
  const int off = flank5(device), n = device->len();
  double sum = 0;
  for (int pos : my_positions) {
    if (pos < 0 || pos >= device->raw_len()) throw msg("density term position out of range");
    for (int k = 0; k < n; k++)
      if (k != pos + off) sum += engine.base_pair_prob(pos + off, k);
  }
  double x = sum / (double)my_positions.size();
  if (my_favorable == FavorableEnum::NO) x = 1 - x;
  return std::log(x);
}

}  // namespace addapt

namespace std {
ostream& operator<<(ostream& out, const addapt::ConditionEnum& c) { return out << (c == addapt::ConditionEnum::APO ? "APO" : "HOLO"); }
ostream& operator<<(ostream& out, const addapt::FavorableEnum& f) { return out << (f == addapt::FavorableEnum::NO ? "NO" : "YES"); }
}
