// Zuker minimum-free-energy fold + traceback for one construct per CTA (integer dcal/mol, bit exact).
//
// The reference never calls vrna_mfe (reference src/scoring.cc includes ViennaRNA/fold.h at :9 but has no
// call site); BASELINE.json configs[1] asks for it, and the oracle (oracle/fold_oracle.c orc_mfe) fixes the
// recurrences and the traceback tie-break order this kernel reproduces exactly:
//   C(i,j)   = min( hairpin(+motif), min_{k,l} interior + C(k,l), min_k F(i+1,k-1)+F(k,j-1) + closing )
//   F(i,j)   = min energy of >= 1 stems in [i,j] inside a multiloop
//   f5(j)    = min( f5(j-1) [j unpaired], min_i f5(i-1) + C(i,j) + ExtStem(i,j) )
// Organisation: same column-parallel scheme as pf_core.cuh, with F built from the rolling column
// Fm1(k,j) (one stem starting at k, unpaired tail) so that a whole column is independent.  Integer min is
// order-independent, so C, F and f5 are bit-identical to the oracle's; the traceback then follows the
// oracle's order of checks, lanes testing candidates in that order and the first hit winning.
#pragma once
#include "fold_common.cuh"

namespace adp {

struct MfeWork {
  int32_t* C;      // [n(n+1)/2]
  int32_t* F;      // [n(n+1)/2]
  int32_t* fm1a;   // [n+2] rolling Fm1 columns
  int32_t* fm1b;
  int32_t* f5;     // [n+2]
  int32_t* ext;    // [n+2] per-row scratch
  int16_t* stack;  // traceback sections: 3 x int16 each, capacity 4n+8
  char* out;       // [n] structure under construction
};

ADP_HD int stem_e_at(const IntTables& P, const SeqCtx& c, int type, int i, int j, bool ext) {
  return stem_e(P, type, i > 1, c.S[i - 1], j < c.n, c.S[j + 1], ext);
}

ADP_HD int hairpin_cell_e(const IntTables& P, const SeqCtx& c, int motif_dcal, int i, int j, int type) {
  const int u = j - i - 1;
  if ((int)c.upr[i + 1] < u) return kInf;
  int e = hairpin_e(P, c.S, i, j, type);
  if (is_motif_site(c, i, j)) e += motif_dcal;
  return e;
}

template <class E>
ADP_HD int mfe_fill(const E& ex, const SeqCtx& c, const IntTables& P, int motif_dcal, MfeWork& w) {
  const int n = c.n, nt = ex.nthreads(), tid = ex.tid();
  constexpr int Gs = E::kGroup;
  const int sub = tid % Gs, ncellslots = nt / Gs;
  const int ntri = n * (n + 1) / 2;
  for (int k = tid; k < ntri; k += nt) { w.C[k] = kInf; w.F[k] = kInf; }
  for (int k = tid; k <= n + 1; k += nt) { w.fm1a[k] = kInf; w.fm1b[k] = kInf; w.ext[k] = kInf; }
  if (tid == 0) w.f5[0] = 0;
  ex.sync();
  int32_t* fm1_prev = w.fm1a;
  int32_t* fm1_cur = w.fm1b;
  const int mlb = P.ML_base;

  for (int j = 1; j <= n; j++) {
    const int ncell = j - kTurn - 1;
    const bool cuj = can_unpair(c, j);
    for (int slot = tid / Gs; slot < ((ncell + ncellslots - 1) / ncellslots) * ncellslots; slot += ncellslots) {
      const int i = slot + 1;
      const bool valid = slot < ncell;
      const int type = valid ? allowed_type(c, i, j) : 0;
      int best = kInf;
      if (type) {
        if (sub == 0) best = hairpin_cell_e(P, c, motif_dcal, i, j, type);
        const int si1 = c.S[i + 1], sj1 = c.S[j - 1];
        int u1max = j - i - 6;
        if (u1max > kMaxLoop) u1max = kMaxLoop;
        if (u1max > (int)c.upr[i + 1]) u1max = c.upr[i + 1];
        const int u2run = c.upl[j - 1];
        for (int u1 = sub; u1 <= u1max; u1 += Gs) {
          const int k = i + 1 + u1;
          int u2max = j - k - 5;
          if (u2max > kMaxLoop - u1) u2max = kMaxLoop - u1;
          if (u2max > u2run) u2max = u2run;
          for (int u2 = 0; u2 <= u2max; u2++) {
            const int l = j - 1 - u2;
            const int v = w.C[tri(k, l)];
            if (v < kInf) {
              const int t2 = pair_type(c.S[k], c.S[l]);
              const int e = v + interior_e(P, u1, u2, type, rtype(t2), si1, sj1, c.S[k - 1], c.S[l + 1]);
              best = e < best ? e : best;
            }
          }
        }
        int dec = kInf;
        for (int k = i + 6 + sub; k <= j - 5; k += Gs) {
          const int a = w.F[tri(i + 1, k - 1)], b = fm1_prev[k];
          if (a < kInf && b < kInf && a + b < dec) dec = a + b;
        }
        if (dec < kInf) {
          const int e = dec + P.ML_closing + stem_e(P, rtype(type), true, sj1, true, si1, false);
          best = e < best ? e : best;
        }
      }
      best = ex.group_min(best);
      if (valid && sub == 0) {
        w.C[tri(i, j)] = best;
        int m1 = (cuj && fm1_prev[i] < kInf) ? fm1_prev[i] + mlb : kInf;
        int e = kInf;
        if (type && best < kInf) {
          const int s = best + stem_e_at(P, c, type, i, j, false);
          m1 = s < m1 ? s : m1;
          if (w.f5[i - 1] < kInf) e = w.f5[i - 1] + best + stem_e_at(P, c, type, i, j, true);
        }
        fm1_cur[i] = m1;
        w.ext[i] = e;
      }
    }
    ex.sync();
    for (int slot = tid / Gs; slot < ((ncell + ncellslots - 1) / ncellslots) * ncellslots; slot += ncellslots) {
      const int i = slot + 1;
      const bool valid = slot < ncell;
      int t = kInf;
      if (valid) {
        const int run = c.upr[i];
        for (int k = i + 1 + sub; k <= j - 4; k += Gs) {
          const int u = k - i;
          int left = w.F[tri(i, k - 1)];
          if (u <= run && u * mlb < left) left = u * mlb;
          const int right = fm1_cur[k];
          if (left < kInf && right < kInf && left + right < t) t = left + right;
        }
      }
      t = ex.group_min(t);
      if (valid && sub == 0) {
        const int own = fm1_cur[i];
        w.F[tri(i, j)] = own < t ? own : t;
      }
    }
    if (ex.warp() == 0) {
      int e = kInf;
      for (int i = 1 + ex.lane(); i <= ncell; i += E::kWarp) e = w.ext[i] < e ? w.ext[i] : e;
      e = ex.warp_min(e);
      if (ex.lane() == 0) {
        const int unp = (cuj && w.f5[j - 1] < kInf) ? w.f5[j - 1] : kInf;
        w.f5[j] = unp < e ? unp : e;
      }
    }
    ex.sync();
    int32_t* tmp = fm1_prev; fm1_prev = fm1_cur; fm1_cur = tmp;
  }
  return w.f5[n];
}

// How the traceback reads the filled matrices: c(i,j) = C(i,j), f(i,j) = F(i,j), kInf where no structure exists.
struct TriView {
  const int32_t* C;
  const int32_t* F;
  ADP_HD int c(int i, int j) const { return C[tri(i, j)]; }
  ADP_HD int f(int i, int j) const { return F[tri(i, j)]; }
};

// Traceback by warp 0 of the CTA (all its lanes must call).  Section kinds: 0 = exterior prefix f5(j),
// 1 = multiloop part F(i,j), 2 = pair (i,j).  `view` abstracts the matrix layout (general kernel: full triangles;
// fast kernel: trimmed, premultiplied triangles); w supplies f5, the section stack and the output string.
template <class E, class View>
ADP_HD void mfe_traceback(const E& ex, const SeqCtx& c, const IntTables& P, const MotifDev* motif, int motif_dcal, const View& view, MfeWork& w) {
  const int n = c.n, lane = ex.lane();
  constexpr int W = E::kWarp;
  for (int p = lane; p < n; p += W) w.out[p] = '.';
  if (w.f5[n] >= kInf) return;
  int sp = 0;
#define ADP_PUSH(a, b, kind)                                                            \
  do {                                                                                  \
    if (lane == 0) { w.stack[3 * sp] = (int16_t)(a); w.stack[3 * sp + 1] = (int16_t)(b); w.stack[3 * sp + 2] = (int16_t)(kind); } \
    sp++;                                                                               \
  } while (0)
#define ADP_WSYNC() ex.warp_sync()
  ADP_PUSH(1, n, 0);
  const int mlb = P.ML_base;
  while (sp > 0) {
    ADP_WSYNC();
    sp--;
    int i = w.stack[3 * sp], j = w.stack[3 * sp + 1];
    const int kind = w.stack[3 * sp + 2];
    ADP_WSYNC();
    if (kind == 0) {
      while (j > 0 && can_unpair(c, j) && w.f5[j - 1] < kInf && w.f5[j] == w.f5[j - 1]) j--;
      if (j == 0) continue;
      // partner k of j: from j-4 DOWN to 1, first match
      int found = -1;
      for (int base = j - kTurn - 1; base >= 1 && found < 0; base -= W) {
        const int k = base - lane;
        bool hit = false;
        if (k >= 1) {
          const int type = allowed_type(c, k, j);
          const int cv = view.c(k, j);
          if (type && cv < kInf && w.f5[k - 1] < kInf) hit = w.f5[j] == w.f5[k - 1] + cv + stem_e_at(P, c, type, k, j, true);
        }
        const int first = ex.warp_first(hit);
        if (first >= 0) found = base - first;
      }
      if (found < 0) continue;   // cannot happen for a consistent fill
      ADP_PUSH(1, found - 1, 0);
      ADP_PUSH(found, j, 2);
    } else if (kind == 2) {
      if (lane == 0) { w.out[i - 1] = '('; w.out[j - 1] = ')'; }
      const int type = allowed_type(c, i, j);
      const int e = view.c(i, j);
      if (e == hairpin_cell_e(P, c, motif_dcal, i, j, type)) {
        if (motif && is_motif_site(c, i, j) && lane == 0)
          for (int q = 0; q < motif->n_pairs; q++) { w.out[i - 1 + motif->pair_i[q]] = '('; w.out[i - 1 + motif->pair_j[q]] = ')'; }
        continue;
      }
      const int si1 = c.S[i + 1], sj1 = c.S[j - 1];
      int u1max = j - i - 6;
      if (u1max > kMaxLoop) u1max = kMaxLoop;
      if (u1max > (int)c.upr[i + 1]) u1max = c.upr[i + 1];
      const int u2run = c.upl[j - 1];
      int fk = -1, fl = -1;
      // interior: k ascending, l descending == u1 ascending, u2 ascending; one u1 per step, lanes over u2
      for (int u1 = 0; u1 <= u1max && fk < 0; u1++) {
        const int k = i + 1 + u1;
        int u2max = j - k - 5;
        if (u2max > kMaxLoop - u1) u2max = kMaxLoop - u1;
        if (u2max > u2run) u2max = u2run;
        for (int base = 0; base <= u2max && fk < 0; base += W) {
          const int u2 = base + lane;
          bool hit = false;
          if (u2 <= u2max) {
            const int l = j - 1 - u2;
            const int v = view.c(k, l);
            if (v < kInf) {
              const int t2 = pair_type(c.S[k], c.S[l]);
              hit = e == v + interior_e(P, u1, u2, type, rtype(t2), si1, sj1, c.S[k - 1], c.S[l + 1]);
            }
          }
          const int first = ex.warp_first(hit);
          if (first >= 0) { fk = k; fl = j - 1 - (base + first); }
        }
      }
      if (fk >= 0) { ADP_PUSH(fk, fl, 2); continue; }
      // multiloop: split k ascending over the oracle's two-sided F
      const int cl = P.ML_closing + stem_e(P, rtype(type), true, sj1, true, si1, false);
      int found = -1;
      for (int base = i + 2; base <= j - 1 && found < 0; base += W) {
        const int k = base + lane;
        bool hit = false;
        if (k <= j - 1) {
          const int a = view.f(i + 1, k - 1), b = view.f(k, j - 1);
          hit = a < kInf && b < kInf && e == a + b + cl;
        }
        const int first = ex.warp_first(hit);
        if (first >= 0) found = base + first;
      }
      if (found < 0) continue;
      ADP_PUSH(i + 1, found - 1, 1);
      ADP_PUSH(found, j - 1, 1);
    } else {
      const int e = view.f(i, j);
      if (j - 1 >= i && can_unpair(c, j) && view.f(i, j - 1) < kInf && e == view.f(i, j - 1) + mlb) { ADP_PUSH(i, j - 1, 1); continue; }
      if (i + 1 <= j && can_unpair(c, i) && view.f(i + 1, j) < kInf && e == view.f(i + 1, j) + mlb) { ADP_PUSH(i + 1, j, 1); continue; }
      const int type = allowed_type(c, i, j);
      if (type && view.c(i, j) < kInf && e == view.c(i, j) + stem_e_at(P, c, type, i, j, false)) { ADP_PUSH(i, j, 2); continue; }
      int found = -1;
      for (int base = i + 1; base <= j && found < 0; base += W) {
        const int k = base + lane;
        bool hit = false;
        if (k <= j) {
          const int a = view.f(i, k - 1), b = view.f(k, j);
          hit = a < kInf && b < kInf && e == a + b;
        }
        const int first = ex.warp_first(hit);
        if (first >= 0) found = base + first;
      }
      if (found < 0) continue;
      ADP_PUSH(i, found - 1, 1);
      ADP_PUSH(found, j, 1);
    }
  }
#undef ADP_PUSH
#undef ADP_WSYNC
}

}  // namespace adp
