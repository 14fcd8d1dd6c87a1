#include "addapt/b200.hh"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <fstream>

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
  vector<adp_bp_term> bp_terms;                 // PairProbTerm / PairedDensityTerm, in the order they were added
  vector<vector<int>> bp_positions;
  vector<string> row_names;                     // score-table rows of one context: macrostate terms, then bp terms (the engine's order)
  vector<double> row_weights;
  string mseq, mfold;
  adp_problem pr;
};

void pack_problem(B200ScoreFunction const& fx, DeviceConstPtr device, ProblemPack* pk) {
  pk->init_seq = device->raw_seq();
  DevicePtr bare = device->copy();
  bare->remove_context();
  for (auto const& ms : bare->macrostates()) { pk->macro_names.push_back(ms.first); pk->macro_csts.push_back(ms.second); }
  vector<string> bp_names;
  vector<double> bp_weights;
  pk->bp_positions.reserve(fx.terms().size());            // pointers into the inner vectors are handed to the C ABI
  for (auto const& t : fx.terms()) {
    if (auto term = std::dynamic_pointer_cast<MacrostateProbTerm>(t)) {
      auto it = std::find(pk->macro_names.begin(), pk->macro_names.end(), term->macrostate());
      if (it == pk->macro_names.end()) throw msg("no macrostate '", term->macrostate(), "'");
      pk->term_macro.push_back((int)(it - pk->macro_names.begin()));
      pk->term_cond.push_back(term->condition() == ConditionEnum::HOLO ? 1 : 0);
      pk->term_fav.push_back(term->favorable() == FavorableEnum::YES ? 1 : 0);
      pk->term_weight.push_back(term->weight());
      pk->row_names.push_back(term->name());
      pk->row_weights.push_back(term->weight());
    } else if (auto pair_term = std::dynamic_pointer_cast<PairProbTerm>(t)) {
      adp_bp_term b{0, pair_term->condition() == ConditionEnum::HOLO ? 1 : 0, pair_term->favorable() == FavorableEnum::YES ? 1 : 0,
                    pair_term->weight(), pair_term->i(), pair_term->j(), 0, nullptr};
      pk->bp_terms.push_back(b);
      bp_names.push_back(pair_term->name());
      bp_weights.push_back(pair_term->weight());
    } else if (auto dens = std::dynamic_pointer_cast<PairedDensityTerm>(t)) {
      pk->bp_positions.push_back(dens->positions());
      adp_bp_term b{1, dens->condition() == ConditionEnum::HOLO ? 1 : 0, dens->favorable() == FavorableEnum::YES ? 1 : 0, dens->weight(), 0, 0,
                    (int)pk->bp_positions.back().size(), pk->bp_positions.back().data()};
      pk->bp_terms.push_back(b);
      bp_names.push_back(dens->name());
      bp_weights.push_back(dens->weight());
    } else {
      throw string("B200ScoreFunction batches MacrostateProbTerm, PairProbTerm and PairedDensityTerm terms");
    }
  }
  pk->row_names.insert(pk->row_names.end(), bp_names.begin(), bp_names.end());
  pk->row_weights.insert(pk->row_weights.end(), bp_weights.begin(), bp_weights.end());
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
  pk->pr.n_bp_terms = (int)pk->bp_terms.size();
  pk->pr.bp_terms = pk->bp_terms.empty() ? nullptr : pk->bp_terms.data();
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
  const size_t per = (size_t)std::max(pk.pr.n_ctx, 1) * (pk.pr.n_terms + pk.pr.n_bp_terms);
  if (terms) terms->assign(B * per, 0.0);
  eng->check(adp_score_batch(eng->handle(), &pk.pr, (int)B, (const uint8_t*)flat.data(), score.data(), terms ? terms->data() : nullptr));
  return score;
}

double B200ScoreFunction::evaluate(DeviceConstPtr device, EvaluatedScoreFunction& table) const {
  for (auto const& t : terms())
    if (!std::dynamic_pointer_cast<MacrostateProbTerm>(t) && !std::dynamic_pointer_cast<PairProbTerm>(t) && !std::dynamic_pointer_cast<PairedDensityTerm>(t))
      return ScoreFunction::evaluate(device, table);   // generic terms: per-object path
  if (contexts().empty() && device->context() &&
      !(device->context()->before().empty() && device->context()->after().empty())) {
    // A device that already carries a context, scored without named contexts: the reference folds device->seq(), flanks
    // included (src/scoring.cc:136-152).  The batched kernel takes contexts from the score function, so hand it this one.
    B200ScoreFunction flanked(my_engine);
    flanked.float_energies(my_float);
    flanked.aptamer(aptamer());
    for (auto const& t : terms()) flanked.add_term(t);
    flanked.add_context("", device->context());
    const double score = flanked.evaluate(device, table);
    for (auto& row : table) row.name = row.name.substr(2);        // drop the ": " prefix of the unnamed context
    return score;
  }
  table.clear();
  vector<double> values;
  DevicePtr bare = device->copy();
  bare->remove_context();
  ProblemPack pk;
  pack_problem(*this, bare, &pk);
  const double score = evaluate_batch(bare, {device->raw_seq()}, &values)[0];
  // the engine reports [context][macrostate terms, bp terms]; the table keeps that order (a score function that mixes the
  // kinds lists its macrostate terms first)
  size_t k = 0;
  auto rows = [&](string prefix) {
    for (size_t r = 0; r < pk.row_names.size(); r++) table.push_back(EvaluatedScoreTerm{prefix + pk.row_names[r], pk.row_weights[r], values[k++]});
  };
  if (contexts().empty()) rows("");
  else for (auto const& named : contexts()) rows(named.first + ": ");
  return score;
}

BatchedMonteCarlo::BatchedMonteCarlo(std::shared_ptr<B200ScoreFunction> scorefxn, int num_steps, ThermostatPtr thermostat)
    : my_scorefxn(scorefxn), my_steps(num_steps), my_thermostat(thermostat) {}

void BatchedMonteCarlo::trajectory_files(string path, int interval) {
  if (interval < 1) throw string("the trajectory interval must be >= 1");
  my_tsv_path = path;
  my_tsv_interval = interval;
}

string BatchedMonteCarlo::trajectory_path(string path, uint64_t id, int n_traj) {
  const size_t pct = path.find("%d");
  if (pct != string::npos) return path.substr(0, pct) + std::to_string(id) + path.substr(pct + 2);
  if (n_traj <= 1) return path;
  const size_t slash = path.find_last_of('/'), dot = path.find_last_of('.');
  if (dot == string::npos || (slash != string::npos && dot < slash)) return path + "." + std::to_string(id);
  return path.substr(0, dot) + "." + std::to_string(id) + path.substr(dot);
}

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
  const bool tsv = my_tsv_interval > 0 && !my_tsv_path.empty() && my_steps > 0;
  const int width = adp_mc_trace_width(&pk.pr), cap = tsv ? (my_steps + my_tsv_interval - 1) / my_tsv_interval : 0;
  vector<double> rec_vals(tsv ? (size_t)n_traj * cap * width : 0);
  vector<uint8_t> rec_seqs(tsv ? (size_t)n_traj * cap * 2 * n : 0);
  eng->check(adp_mc_run_trace(eng->handle(), &pk.pr, n_traj, seed, first_id, my_steps, kind, args, fin.data(), fscore.data(), best.data(),
                              bscore.data(), bstep.data(), counts.data(), tsv ? my_tsv_interval : 0, tsv ? rec_vals.data() : nullptr,
                              tsv ? rec_seqs.data() : nullptr));
  if (tsv) {
    // TsvTrajectoryReporter's layout, column for column (reference src/sampling.cc:436-489); numbers through the same ostream defaults
    static const char* outcome_names[4] = {"REJECT", "ACCEPT_WORSENED", "ACCEPT_UNCHANGED", "ACCEPT_IMPROVED"};
    vector<string> ctx_names;
    for (auto const& named : my_scorefxn->contexts()) ctx_names.push_back(named.first);
    const size_t n_rows = pk.row_names.size(), n_ctx = ctx_names.empty() ? 1 : ctx_names.size();
    for (int t = 0; t < n_traj; t++) {
      const string path = trajectory_path(my_tsv_path, first_id + (uint64_t)t, n_traj);
      std::ofstream out(path);
      if (!out.is_open()) throw msg("couldn't open '", path, "' for writing");
      out << "#\t" << "initial_seq\t" << device->seq() << "\n";
      out << "step\tnum_steps\tcurrent_score\tproposed_score\t";
      for (size_t x = 0; x < n_ctx; x++)
        for (size_t r = 0; r < n_rows; r++) {
          const string nm = (ctx_names.empty() ? string("") : ctx_names[x] + ": ") + pk.row_names[r];
          out << "term_weight[" << nm << "]\tterm_value[" << nm << "]\t";
        }
      out << "score_diff\ttemperature\tmetropolis_criterion\trandom_threshold\tmove\toutcome\tcurrent_seq\tproposed_seq\t" << std::endl;
      for (int r = 0; r < cap; r++) {
        const double* v = rec_vals.data() + ((size_t)t * cap + r) * width;
        const char* q = (const char*)rec_seqs.data() + ((size_t)t * cap + r) * 2 * n;
        out << (int)v[0] << "\t" << my_steps << "\t" << v[1] << "\t" << v[2] << "\t";
        for (size_t x = 0; x < n_ctx; x++)
          for (size_t k = 0; k < n_rows; k++) out << pk.row_weights[k] << "\t" << v[8 + x * n_rows + k] << "\t";
        out << v[3] << "\t" << v[4] << "\t" << v[5] << "\t" << v[6] << "\t" << "UnbiasedMutation" << "\t" << outcome_names[(int)v[7]] << "\t"
            << string(q, n) << "\t" << string(q + n, n) << "\t" << std::endl;
      }
    }
  }
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
