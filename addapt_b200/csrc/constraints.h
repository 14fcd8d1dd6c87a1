// Host-side preprocessing of a dot-bracket hard constraint into the per-position arrays the kernels use.
// ≙ vrna_constraints_add(fc, cst, VRNA_CONSTRAINT_DB_DEFAULT | VRNA_CONSTRAINT_DB_ENFORCE_BP)
// (reference src/scoring.cc:61-62); symbols the reference uses: . x ( )  (tests/test_scoring.cc:58,134-145).
#pragma once
#include <string>

#include "fold_common.cuh"

namespace adp {

// Throws std::string on a malformed constraint (wrong length, unknown symbol, unbalanced brackets).
inline void build_constraint(const std::string& cst, int n, ConstraintDev* out) {
  if ((int)cst.size() != n) throw std::string("constraint length ") + std::to_string(cst.size()) + " != sequence length " + std::to_string(n);
  if (n > kMaxLen) throw std::string("construct too long");
  for (int p = 0; p <= n + 1; p++) { out->partner[p] = 0; out->enc[p] = 0; out->flag[p] = 0; }
  int open_pos[kMaxLen + 2], open_id[kMaxLen + 2], depth = 0, next_id = 1;
  for (int p = 1; p <= n; p++) {
    const char ch = cst[p - 1];
    const int parent = depth ? open_id[depth - 1] : 0;
    if (ch == '(') {
      out->enc[p] = (uint16_t)parent;
      out->flag[p] = 2;
      open_pos[depth] = p;
      open_id[depth] = next_id++;
      depth++;
    } else if (ch == ')') {
      if (!depth) throw std::string("unbalanced brackets in constraint");
      depth--;
      const int q = open_pos[depth];
      out->partner[p] = (uint16_t)q;
      out->partner[q] = (uint16_t)p;
      out->flag[p] = 2;
      out->enc[p] = (uint16_t)(depth ? open_id[depth - 1] : 0);
    } else if (ch == 'x') {
      out->flag[p] = 1;
      out->enc[p] = (uint16_t)parent;
    } else if (ch == '.') {
      out->enc[p] = (uint16_t)parent;
    } else {
      throw std::string("unsupported constraint symbol '") + ch + "'";
    }
  }
  if (depth) throw std::string("unbalanced brackets in constraint");
  // run lengths of positions that may stay unpaired ('.' and 'x'): upr counts p, p+1, ..., upl counts p, p-1, ...
  out->upr[n + 1] = out->upr[n + 2] = 0;
  for (int p = n; p >= 1; p--) out->upr[p] = (out->flag[p] & 2) ? 0 : (uint16_t)(out->upr[p + 1] + 1);
  out->upr[0] = out->upl[0] = out->upl[n + 1] = 0;
  for (int p = 1; p <= n; p++) out->upl[p] = (out->flag[p] & 2) ? 0 : (uint16_t)(out->upl[p - 1] + 1);
}

}  // namespace adp
