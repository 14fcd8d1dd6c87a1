// Host-side compilation of a design problem (adp_problem of include/addapt_b200.h) into the flat records
// the design kernels read: contexts and their padded constraints, the list of distinct ensemble fills
// ("variants") the score terms need, and the mutation closures.
//   contexts / macrostate padding   ≙ Device::context, Device::macrostate      (reference src/model.cc:43-51)
//   term -> (condition, macrostate) ≙ MacrostateProbTerm                        (reference src/scoring.cc:204-259)
//   free positions                  ≙ can_be_mutated / can_be_freely_mutated    (reference src/sampling.cc:170-193)
//   closures                        ≙ mutate_recursively                        (reference src/sampling.cc:205-282)
#pragma once
#include <cctype>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/addapt_b200.h"
#include "constraints.h"
#include "design_core.cuh"
#include "energy_model.h"

namespace adp {

struct CompiledProblem {
  ScoreProblemDev sp;                    // sp.csts / sp.motif / sp.T are filled in by the caller
  std::vector<ConstraintDev> csts;       // [n_ctx][n_macro]
  bool has_motif = false;
  std::vector<int32_t> bp_qi, bp_qj, bp_dpos;   // [n_ctx][n_bpq] / [n_ctx][n_bpd]: contextual positions of the bp terms (sp.bp_* point at device copies)
  bool needs_outside() const { return sp.n_bp > 0; }
};

// Throws std::string on malformed input.
inline void compile_problem(const adp_problem& pr, double kT_kcal, CompiledProblem* out) {
  ScoreProblemDev& sp = out->sp;
  std::memset(&sp, 0, sizeof sp);
  if (pr.n < 1 || pr.n > ADP_MAX_LEN) throw std::string("device length out of range");
  if (pr.n_macro < 0 || pr.n_macro > kMaxMacro) throw std::string("too many macrostates (max 8)");
  if (pr.n_terms < 0 || pr.n_terms > kMaxTerms) throw std::string("too many score terms (max 16)");
  if (pr.n_bp_terms < 0 || pr.n_bp_terms > kMaxBp || pr.n_terms + pr.n_bp_terms > kMaxTerms)
    throw std::string("too many base-pair-probability terms (max 8, and 16 terms in total)");
  if (pr.n_bp_terms > 0 && !pr.bp_terms) throw std::string("missing base-pair-probability terms");
  if (pr.n_ctx < 0 || pr.n_ctx > kMaxCtx) throw std::string("too many contexts (max 8)");
  sp.n = pr.n;
  sp.n_ctx = pr.n_ctx > 0 ? pr.n_ctx : 1;
  sp.n_macro = pr.n_macro;
  sp.n_terms = pr.n_terms;
  sp.float_dg = pr.float_dg;
  sp.kT = kT_kcal;
  out->has_motif = pr.motif.seq != nullptr;
  sp.n_max = 0;
  for (int x = 0; x < sp.n_ctx; x++) {
    std::string bef = pr.n_ctx > 0 && pr.ctx_before && pr.ctx_before[x] ? pr.ctx_before[x] : "";
    std::string aft = pr.n_ctx > 0 && pr.ctx_after && pr.ctx_after[x] ? pr.ctx_after[x] : "";
    if (bef.size() > (size_t)kMaxFlank || aft.size() > (size_t)kMaxFlank) throw std::string("context flank longer than 64 nt");
    sp.nb[x] = (int)bef.size();
    sp.na[x] = (int)aft.size();
    for (size_t k = 0; k < bef.size(); k++) sp.before[x][k] = (uint8_t)bef[k];
    for (size_t k = 0; k < aft.size(); k++) sp.after[x][k] = (uint8_t)aft[k];
    const int nx = sp.nb[x] + sp.n + sp.na[x];
    if (nx > ADP_MAX_LEN) throw std::string("device + context longer than the engine's maximum length");
    if (nx > sp.n_max) sp.n_max = nx;
  }
  // An enforced pair between two FIXED nucleotides that cannot pair canonically: ViennaRNA keeps such a pair as "pair type 7"
  // (vrna_hc_add_bp resets the type of an enforced non-canonical pair), these kernels would give it weight 0 (P = 0, log P = -inf).
  // That divergence is refused loudly rather than returned silently; mutable pairs are kept complementary by the moves.
  for (int m = 0; m < pr.n_macro; m++) {
    if (!pr.macro_cst || !pr.macro_cst[m] || (int)std::strlen(pr.macro_cst[m]) != pr.n) continue;      // reported below
    std::vector<int> open;
    for (int p = 0; p < pr.n; p++) {
      const char ch = pr.macro_cst[m][p];
      if (ch == '(') open.push_back(p);
      else if (ch == ')' && !open.empty()) {
        const int q = open.back();
        open.pop_back();
        const char a = pr.init_seq[q], b = pr.init_seq[p];
        if (std::islower((unsigned char)a) && std::islower((unsigned char)b)) {
          const int ta = EnergyModel::encode(a), tb = EnergyModel::encode(b);
          if (ta && tb && !EnergyModel::pair_type(ta, tb))
            throw std::string("macrostate ") + std::to_string(m) + " enforces the non-canonical pair " + a + std::to_string(q) + "-" + b + std::to_string(p) +
                " between fixed positions: ViennaRNA folds it as pair type 7, this engine does not support that";
        }
      }
    }
  }
  out->csts.resize((size_t)sp.n_ctx * (sp.n_macro > 0 ? sp.n_macro : 1));
  for (int x = 0; x < sp.n_ctx; x++)
    for (int m = 0; m < sp.n_macro; m++) {
      if (!pr.macro_cst || !pr.macro_cst[m]) throw std::string("missing macrostate constraint");
      std::string c(pr.macro_cst[m]);
      if ((int)c.size() != pr.n) throw std::string("macrostate constraint length differs from the device length");
      std::string padded = std::string(sp.nb[x], '.') + c + std::string(sp.na[x], '.');     // model.cc:43-51
      build_constraint(padded, (int)padded.size(), &out->csts[(size_t)x * sp.n_macro + m]);
    }
  // distinct fills: (condition, none) and (condition, macrostate)
  auto find_or_add = [&](int cond, int macro) {
    const int use_motif = (cond == 1 && out->has_motif) ? 1 : 0;
    for (int v = 0; v < sp.V; v++)
      if (sp.variants[v].cst_id == macro && sp.variants[v].use_motif == use_motif) return v;
    if (sp.V >= kMaxVariants) throw std::string("score function needs more than 8 distinct ensemble fills");
    sp.variants[sp.V] = VariantDev{macro, use_motif, 0, 0};
    return sp.V++;
  };
  for (int t = 0; t < sp.n_terms; t++) {
    const int macro = pr.term_macro[t], cond = pr.term_cond[t];
    if (macro < 0 || macro >= sp.n_macro) throw std::string("score term refers to an unknown macrostate");
    if (cond != 0 && cond != 1) throw std::string("score term condition must be 0 (apo) or 1 (holo)");
    sp.term_fav[t] = pr.term_favorable[t] ? 1 : 0;
    sp.term_w[t] = pr.term_weight[t];
    sp.term_v0[t] = find_or_add(cond, -1);
    sp.term_v1[t] = find_or_add(cond, macro);
  }
  // base-pair-probability terms: pair queries and density positions of the raw device, shifted per context
  sp.n_bp = pr.n_bp_terms;
  std::vector<int> qi, qj, dpos;
  for (int t = 0; t < sp.n_bp; t++) {
    const adp_bp_term& b = pr.bp_terms[t];
    if (b.cond != 0 && b.cond != 1) throw std::string("score term condition must be 0 (apo) or 1 (holo)");
    sp.bp_kind[t] = b.kind;
    sp.bp_fav[t] = b.favorable ? 1 : 0;
    sp.bp_w[t] = b.weight;
    sp.bp_v[t] = find_or_add(b.cond, -1);
    sp.variants[sp.bp_v[t]].outside = 1;
    if (b.kind == 0) {
      if (b.i < 0 || b.i >= pr.n || b.j < 0 || b.j >= pr.n) throw std::string("pair term position out of range");
      sp.bp_at[t] = (int)qi.size();
      sp.bp_npos[t] = 0;
      qi.push_back(b.i);
      qj.push_back(b.j);
    } else if (b.kind == 1) {
      if (b.n_pos < 1 || !b.pos) throw std::string("density term without positions");
      sp.bp_at[t] = (int)dpos.size();
      sp.bp_npos[t] = b.n_pos;
      for (int q = 0; q < b.n_pos; q++) {
        if (b.pos[q] < 0 || b.pos[q] >= pr.n) throw std::string("density term position out of range");
        dpos.push_back(b.pos[q]);
      }
      if ((int)dpos.size() > kMaxDensPos) throw std::string("density terms list more than 64 positions");
    } else {
      throw std::string("base-pair-probability term kind must be 0 (pair) or 1 (density)");
    }
  }
  // ligand-bound fills without outside pass derive from their apo twin (fold_jobs.cuh, PfJob::site_ratio); an apo twin that has an
  // outside pass of its own (bp terms) cannot report the ratio, so its holo variant is folded
  for (int v = 0; v < kMaxVariants; v++) { sp.twin[v] = -1; sp.site_of[v] = 0; }
  const int V0 = sp.V;
  for (int v = 0; v < V0; v++) {
    if (!sp.variants[v].use_motif || sp.variants[v].outside) continue;
    int u = -1;
    for (int k = 0; k < sp.V; k++)
      if (!sp.variants[k].use_motif && sp.variants[k].cst_id == sp.variants[v].cst_id) u = k;
    if (u < 0) {
      if (sp.V >= kMaxVariants) continue;             // no room for the twin: fold the holo ensemble itself
      sp.variants[sp.V] = VariantDev{sp.variants[v].cst_id, 0, 0, 0};
      u = sp.V++;
    }
    if (sp.variants[u].outside) continue;
    sp.twin[v] = u;
    sp.site_of[u] = 1;
  }
  sp.n_bpq = (int)qi.size();
  sp.n_bpd = (int)dpos.size();
  for (int x = 0; x < sp.n_ctx; x++) {
    for (size_t q = 0; q < qi.size(); q++) { out->bp_qi.push_back(qi[q] + sp.nb[x]); out->bp_qj.push_back(qj[q] + sp.nb[x]); }
    for (size_t q = 0; q < dpos.size(); q++) out->bp_dpos.push_back(dpos[q] + sp.nb[x]);
  }
  sp.bp_qi = out->bp_qi.data();      // host pointers (the emulation uses them as they are); the engine patches in device copies
  sp.bp_qj = out->bp_qj.data();
  sp.bp_dpos = out->bp_dpos.data();
}

// Closure of position i under mutate_recursively (reference src/sampling.cc:205-282): every position that has to change with
// it, alternating nucleotide / complement along every bracket partner in every macrostate.  Throws std::string where the
// reference would throw (partner not mutable, unbalanced brackets, contradictory pairing, or -- here -- more than kMaxClosure positions).
inline void compile_closure(const adp_problem& pr, int i, std::vector<int>* order_out, std::vector<int>* parity_out) {
  const int n = pr.n;
  const std::string seq(pr.init_seq, pr.init_seq + n);
  std::vector<std::string> macros;
  for (int m = 0; m < pr.n_macro; m++) macros.emplace_back(pr.macro_cst[m]);
  auto mutable_at = [&](int p) { return std::isupper((unsigned char)seq[p]) != 0; };
  if (i < 0 || i >= n) throw std::string("position out of range");
  if (!mutable_at(i)) throw std::string("position ") + std::to_string(i) + " is not mutable";
  std::vector<int> parity(n, -1), order;
  std::vector<std::pair<int, int>> todo{{i, 0}};
  parity[i] = 0;
  order.push_back(i);
  for (size_t head = 0; head < todo.size(); head++) {
    const int p = todo[head].first, par = todo[head].second;
    for (auto& ms : macros) {
      char open, close;
      int dir;
      if (ms[p] == '(') { open = '('; close = ')'; dir = 1; }
      else if (ms[p] == ')') { open = ')'; close = '('; dir = -1; }
      else continue;
      int level = 1, q = p;
      while (level) {
        q += dir;
        if (q < 0 || q >= n) throw std::string("mismatched base-pair in a macrostate");
        if (ms[q] == open) level++;
        else if (ms[q] == close) level--;
      }
      if (!mutable_at(q)) throw std::string("position ") + std::to_string(i) + " cannot be mutated: its partner " + std::to_string(q) + " is not mutable";
      if (parity[q] < 0) {
        parity[q] = par ^ 1;
        order.push_back(q);
        todo.emplace_back(q, par ^ 1);
      } else if (parity[q] != (par ^ 1)) {
        throw std::string("macrostates impose unsatisfiable base-pairing on position ") + std::to_string(i);
      }
    }
  }
  if ((int)order.size() > kMaxClosure) throw std::string("a mutation would change more than 16 positions");
  *order_out = order;
  parity_out->assign(order.size(), 0);
  for (size_t k = 0; k < order.size(); k++) (*parity_out)[k] = parity[order[k]];
}

inline void store_closure(MoveTableDev* mv, int slot, int i, const std::vector<int>& order, const std::vector<int>& parity) {
  mv->free_pos[slot] = (int16_t)i;
  mv->closure_len[slot] = (int8_t)order.size();
  for (size_t k = 0; k < order.size(); k++) {
    mv->closure_pos[slot][k] = (int16_t)order[k];
    mv->closure_comp[slot][k] = (int8_t)parity[k];
  }
}

// ≙ can_be_freely_mutated (reference src/sampling.cc:177-193): upper case and not the 3' end of a pair in any macrostate
inline bool freely_mutable(const adp_problem& pr, int i) {
  if (i < 0 || i >= pr.n || !std::isupper((unsigned char)pr.init_seq[i])) return false;
  for (int m = 0; m < pr.n_macro; m++)
    if (pr.macro_cst[m][i] == ')') return false;
  return true;
}

// The proposal table of the Monte-Carlo kernel: one slot per freely mutable position (UnbiasedMutationMove, reference
// src/sampling.cc:287-303).  Throws std::string when a drawn mutation could not be applied (the reference would throw at that step).
inline void compile_moves(const adp_problem& pr, MoveTableDev* mv) {
  std::memset(mv, 0, sizeof *mv);
  std::vector<int> order, parity;
  for (int i = 0; i < pr.n; i++) {
    if (!freely_mutable(pr, i)) continue;
    compile_closure(pr, i, &order, &parity);
    store_closure(mv, mv->n_free++, i, order, parity);
  }
}

}  // namespace adp
