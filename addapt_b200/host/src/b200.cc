#include "addapt/b200.hh"

#include <cmath>
#include <cstring>

#include "../../../include/addapt_b200.h"

namespace addapt {

namespace {
B200EnginePtr g_default_engine;

string upper(string s) {
  for (char& ch : s) ch = (char)std::toupper((unsigned char)ch);
  return s;
}
}  // namespace

B200Engine::B200Engine(int device, string par_file) : my_handle(nullptr) {
  int rc = adp_engine_create(device, par_file.empty() ? nullptr : par_file.c_str(), &my_handle);
  if (rc != ADP_OK) throw msg("can't create the B200 fold engine: ", adp_last_error(nullptr));
}
B200Engine::~B200Engine() { adp_engine_destroy(my_handle); }
double B200Engine::kT() const { return adp_kT_kcal(my_handle); }
void B200Engine::check(int rc) const {
  if (rc != ADP_OK) throw string(adp_last_error(my_handle));
}
int B200Engine::device_count() { return adp_device_count(); }

B200EnginePtr default_engine() {
  if (!g_default_engine) g_default_engine = make_shared<B200Engine>(0);
  return g_default_engine;
}
void default_engine(B200EnginePtr e) { g_default_engine = e; }

FoldFactory b200_fold_factory(B200EnginePtr engine) {
  return [engine](DeviceConstPtr device, AptamerConstPtr aptamer) -> RnaFoldPtr {
    return make_shared<B200RnaFold>(device, aptamer, engine);
  };
}

B200RnaFold::B200RnaFold(DeviceConstPtr device, AptamerConstPtr aptamer, B200EnginePtr engine)
    : my_engine(engine ? engine : default_engine()), my_aptamer(aptamer), my_seq(upper(device->seq())) {}

double B200RnaFold::ensemble_energy(const char* cst) const {
  adp_motif motif{nullptr, nullptr, 0.};
  string mseq, mfold;
  if (my_aptamer) {
    mseq = my_aptamer->seq();
    mfold = my_aptamer->fold();
    motif = adp_motif{mseq.c_str(), mfold.c_str(), my_aptamer->affinity()};
  }
  double dG = 0;
  my_engine->check(adp_pf_batch(my_engine->handle(), 1, (int)my_seq.size(), (const uint8_t*)my_seq.data(), cst,
                                my_aptamer ? &motif : nullptr, &dG, nullptr));
  return my_float ? (double)(float)dG : dG;
}

double B200RnaFold::macrostate_prob(string constraint) const {
  if (constraint.size() != my_seq.size()) throw msg("constraint length ", constraint.size(), " != sequence length ", my_seq.size());
  if (my_seq.empty()) return 1.0;
  const double g_total = ensemble_energy(nullptr);
  const double g_state = ensemble_energy(constraint.c_str());
  return std::exp((g_total - g_state) / my_engine->kT());      // exp(-inf) = 0 for an empty constrained ensemble
}

double B200RnaFold::base_pair_prob(int a, int b) const {
  auto range = normalize_range(my_seq, a, b, IndexEnum::ITEM);      // throws std::string when out of bounds
  const int n = (int)my_seq.size(), i = range.first, j = range.second;
  if (i == j) return 0.0;
  if (my_bppm.empty()) {
    adp_motif motif{nullptr, nullptr, 0.};
    string mseq, mfold;
    if (my_aptamer) {
      mseq = my_aptamer->seq();
      mfold = my_aptamer->fold();
      motif = adp_motif{mseq.c_str(), mfold.c_str(), my_aptamer->affinity()};
    }
    my_bppm.assign((size_t)n * (n + 1) / 2, 0.0);
    double dG = 0;
    my_engine->check(adp_pf_batch(my_engine->handle(), 1, n, (const uint8_t*)my_seq.data(), nullptr, my_aptamer ? &motif : nullptr,
                                  &dG, my_bppm.data()));
  }
  return my_bppm[(size_t)i * n - (size_t)i * (i - 1) / 2 + (size_t)(j - i)];
}

B200ScoreFunction::B200ScoreFunction(B200EnginePtr engine) : my_engine(engine) {
  B200EnginePtr* slot = &my_engine;
  // engines are created lazily so that building a score function needs no GPU
  fold_factory([slot](DeviceConstPtr device, AptamerConstPtr aptamer) -> RnaFoldPtr {
    if (!*slot) *slot = default_engine();
    return make_shared<B200RnaFold>(device, aptamer, *slot);
  });
}

namespace {

// everything adp_problem points to, kept alive for the duration of a call
struct ProblemPack {
  string init_seq;
  vector<string> macro_names, macro_csts, before, after;
  vector<const char*> macro_ptrs, before_ptrs, after_ptrs;
  vector<int> term_macro, term_cond, term_fav;
  vector<double> term_weight;
  string mseq, mfold;
  adp_problem pr;
};

void pack_problem(B200ScoreFunction const& fx, DeviceConstPtr device, ProblemPack* pk) {
  pk->init_seq = device->raw_seq();
  DevicePtr bare = device->copy();
  bare->remove_context();
  for (auto const& ms : bare->macrostates()) { pk->macro_names.push_back(ms.first); pk->macro_csts.push_back(ms.second); }
  for (auto const& t : fx.terms()) {
    auto term = std::dynamic_pointer_cast<MacrostateProbTerm>(t);
    if (!term) throw string("B200ScoreFunction batches only MacrostateProbTerm terms");
    auto it = std::find(pk->macro_names.begin(), pk->macro_names.end(), term->macrostate());
    if (it == pk->macro_names.end()) throw msg("no macrostate '", term->macrostate(), "'");
    pk->term_macro.push_back((int)(it - pk->macro_names.begin()));
    pk->term_cond.push_back(term->condition() == ConditionEnum::HOLO ? 1 : 0);
    pk->term_fav.push_back(term->favorable() == FavorableEnum::YES ? 1 : 0);
    pk->term_weight.push_back(term->weight());
  }
  for (auto const& named : fx.contexts()) { pk->before.push_back(named.second->before()); pk->after.push_back(named.second->after()); }
  for (auto& s : pk->macro_csts) pk->macro_ptrs.push_back(s.c_str());
  for (auto& s : pk->before) pk->before_ptrs.push_back(s.c_str());
  for (auto& s : pk->after) pk->after_ptrs.push_back(s.c_str());
  std::memset(&pk->pr, 0, sizeof pk->pr);
  pk->pr.n = (int)pk->init_seq.size();
  pk->pr.init_seq = pk->init_seq.c_str();
  pk->pr.n_macro = (int)pk->macro_ptrs.size();
  pk->pr.macro_cst = pk->macro_ptrs.data();
  pk->pr.n_terms = (int)pk->term_macro.size();
  pk->pr.term_macro = pk->term_macro.data();
  pk->pr.term_cond = pk->term_cond.data();
  pk->pr.term_favorable = pk->term_fav.data();
  pk->pr.term_weight = pk->term_weight.data();
  if (fx.aptamer()) {
    pk->mseq = fx.aptamer()->seq();
    pk->mfold = fx.aptamer()->fold();
    pk->pr.motif = adp_motif{pk->mseq.c_str(), pk->mfold.c_str(), fx.aptamer()->affinity()};
  }
  pk->pr.n_ctx = (int)pk->before.size();
  pk->pr.ctx_before = pk->before_ptrs.data();
  pk->pr.ctx_after = pk->after_ptrs.data();
  pk->pr.float_dg = fx.float_energies() ? 1 : 0;
}

}  // namespace

vector<double> B200ScoreFunction::evaluate_batch(DeviceConstPtr device, vector<string> const& raw_seqs, vector<double>* terms) const {
  B200EnginePtr eng = my_engine ? my_engine : default_engine();
  ProblemPack pk;
  pack_problem(*this, device, &pk);
  const size_t n = pk.init_seq.size(), B = raw_seqs.size();
  string flat;
  flat.reserve(n * B);
  for (auto const& s : raw_seqs) {
    if (s.size() != n) throw msg("sequence length ", s.size(), " != device length ", n);
    flat += upper(s);
  }
  vector<double> score(B, 0.0);
  const size_t per = (size_t)std::max(pk.pr.n_ctx, 1) * pk.pr.n_terms;
  if (terms) terms->assign(B * per, 0.0);
  eng->check(adp_score_batch(eng->handle(), &pk.pr, (int)B, (const uint8_t*)flat.data(), score.data(), terms ? terms->data() : nullptr));
  return score;
}

double B200ScoreFunction::evaluate(DeviceConstPtr device, EvaluatedScoreFunction& table) const {
  for (auto const& t : terms())
    if (!std::dynamic_pointer_cast<MacrostateProbTerm>(t)) return ScoreFunction::evaluate(device, table);   // generic terms: per-object path
  table.clear();
  vector<double> values;
  DevicePtr bare = device->copy();
  bare->remove_context();
  const double score = evaluate_batch(bare, {device->raw_seq()}, &values)[0];
  size_t k = 0;
  auto rows = [&](string prefix) {
    for (auto const& t : terms()) table.push_back(EvaluatedScoreTerm{prefix + t->name(), t->weight(), values[k++]});
  };
  if (contexts().empty()) rows("");
  else for (auto const& named : contexts()) rows(named.first + ": ");
  return score;
}

BatchedMonteCarlo::BatchedMonteCarlo(std::shared_ptr<B200ScoreFunction> scorefxn, int num_steps, ThermostatPtr thermostat)
    : my_scorefxn(scorefxn), my_steps(num_steps), my_thermostat(thermostat) {}

vector<BatchedTrajectory> BatchedMonteCarlo::apply(DeviceConstPtr device, int n_traj, uint64_t seed, uint64_t first_id) const {
  B200EnginePtr eng = my_scorefxn->engine() ? my_scorefxn->engine() : default_engine();
  ProblemPack pk;
  pack_problem(*my_scorefxn, device, &pk);
  int kind;
  double args[3] = {0, 0, 1};
  if (auto fixed = std::dynamic_pointer_cast<FixedThermostat>(my_thermostat)) {
    kind = 0;
    args[0] = fixed->temperature();
  } else if (auto anneal = std::dynamic_pointer_cast<AnnealingThermostat>(my_thermostat)) {
    kind = 1;
    args[0] = anneal->max_temperature();
    args[1] = anneal->min_temperature();
    args[2] = anneal->cycle_len();
  } else if (auto scaling = std::dynamic_pointer_cast<AutoScalingThermostat>(my_thermostat)) {
    kind = 2;
    args[0] = scaling->target_acceptance_rate();
    args[1] = scaling->training_period();
    args[2] = scaling->temperature();
  } else {
    throw string("the batched Monte Carlo kernel supports the Fixed, Annealing and AutoScaling thermostats");
  }
  const size_t n = pk.init_seq.size();
  vector<uint8_t> fin(n * n_traj), best(n * n_traj);
  vector<double> fscore(n_traj), bscore(n_traj);
  vector<int32_t> bstep(n_traj);
  vector<uint32_t> counts((size_t)4 * n_traj);
  eng->check(adp_mc_run(eng->handle(), &pk.pr, n_traj, seed, first_id, my_steps, kind, args, fin.data(), fscore.data(), best.data(),
                        bscore.data(), bstep.data(), counts.data()));
  vector<BatchedTrajectory> out(n_traj);
  for (int t = 0; t < n_traj; t++) {
    out[t].final_seq.assign((const char*)fin.data() + n * t, n);
    out[t].best_seq.assign((const char*)best.data() + n * t, n);
    out[t].final_score = fscore[t];
    out[t].best_score = bscore[t];
    out[t].best_step = bstep[t];
    out[t].outcome_counters[OutcomeEnum::REJECT] = (int)counts[4 * t + 0];
    out[t].outcome_counters[OutcomeEnum::ACCEPT_WORSENED] = (int)counts[4 * t + 1];
    out[t].outcome_counters[OutcomeEnum::ACCEPT_UNCHANGED] = (int)counts[4 * t + 2];
    out[t].outcome_counters[OutcomeEnum::ACCEPT_IMPROVED] = (int)counts[4 * t + 3];
  }
  return out;
}

}  // namespace addapt
