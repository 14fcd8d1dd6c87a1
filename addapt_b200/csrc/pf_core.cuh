// McCaskill partition function (inside) and base-pair probabilities (outside) for one construct per CTA.
//
// Computes what the reference obtains from vrna_pf(fc, NULL) (reference src/scoring.cc:43,58,65) under the
// hard constraint of vrna_constraints_add (:61-62) and the ligand-motif soft constraint of
// vrna_sc_add_hi_motif (:94-99).  Organisation (this engine's, not libRNA's):
//   * column-parallel recursion: all cells (i,j) of one column j are independent given the columns before
//     it, so a column is three barrier-separated phases -- Qb(.,j) | Qm1(.,j) + exterior | Qm(.,j);
//   * Qb, Qm (and P for the outside pass) are column-major triangular matrices: a lane group per cell reads
//     consecutive rows of one column (bank-conflict free), and rows across consecutive columns hit the
//     triangular-number stride, which is also conflict free for 8-byte words;
//   * the exterior partition function is kept only as the prefix q1[j] = Q(1,j) and suffix qln[k] = Q(k,n)
//     vectors (O(n^2) work) instead of libRNA's full Q matrix (O(n^3)); same ensemble, fewer terms.
#pragma once
#include "fold_common.cuh"

namespace adp {

struct PfWork {
  double* qb;     // Qb(i,j): i and j pair                              [n(n+1)/2]
  double* qm;     // Qm(i,j): at least one stem inside a multiloop      [n(n+1)/2]
  double* pr;     // outside weight / probability of (i,j)              [n(n+1)/2]  (outside pass only)
  double* qm1a;   // Qm1(.,j-1) and Qm1(.,j): exactly one stem starting at the row, rolling columns   [n+2] each
  double* qm1b;
  double* q1;     // q1[j] = Q(1,j), q1[0] = 1                          [n+2]
  double* qln;    // qln[k] = Q(k,n), qln[n+1] = 1                      [n+2]
  double* ext;    // per-row scratch of the current column              [n+2]
  double* va;     // outside: A(i,l)                                    [n+2]
  double* vb;     // outside: B(i,l)                                    [n+2]
  double* vab;    // outside: A+B                                       [n+2]
  const double* scale;  // pf_scale^-k (global)
  const double* eMLb;   // multiloop unpaired-base weight ^k (global)
};

// exterior / multiloop stem weight of the pair (i,j): neighbours dangle whenever they exist
ADP_HD double stem_at(const PfSmall& P, const SeqCtx& c, int type, int i, int j, bool ext) {
  return stem_x(P, type, i > 1, c.S[i - 1], j < c.n, c.S[j + 1], ext);
}

ADP_HD double hairpin_w(const PfSmall& P, const PfTables* G, const SeqCtx& c, const PfWork& w, double motif_boltz, int i, int j, int type) {
  const int u = j - i - 1;
  if ((int)c.upr[i + 1] < u) return 0.;
  double q = hairpin_x(P, G, c.S, i, j, type) * w.scale[u + 2];
  if (is_motif_site(c, i, j)) q *= motif_boltz;   // the bound motif REPLACES the empty loop
  return q;
}

// Inside pass.  Returns Z = Q(1,n) (scaled) on every thread.
template <class E>
ADP_HD double pf_inside(const E& ex, const SeqCtx& c, const PfSmall& P, const PfTables* G, double motif_boltz, PfWork& w) {
  const int n = c.n, nt = ex.nthreads(), tid = ex.tid();
  constexpr int Gs = E::kGroup;
  const int sub = tid % Gs, ncellslots = nt / Gs;
  const int ntri = n * (n + 1) / 2;
  for (int k = tid; k < ntri; k += nt) { w.qb[k] = 0.; w.qm[k] = 0.; }
  for (int k = tid; k <= n + 1; k += nt) { w.qm1a[k] = 0.; w.qm1b[k] = 0.; w.ext[k] = 0.; }
  if (tid == 0) w.q1[0] = 1.;
  ex.sync();
  double* qm1_prev = w.qm1a;
  double* qm1_cur = w.qm1b;
  const double s1 = P.scale1, eb1 = P.eMLb1;

  for (int j = 1; j <= n; j++) {
    const int ncell = j - kTurn - 1;       // rows 1 .. j-4 can pair with j
    const bool cuj = can_unpair(c, j);
    // ---- phase A: Qb(i,j), then Qm1(i,j) and the exterior term of row i
    for (int slot = tid / Gs; slot < ((ncell + ncellslots - 1) / ncellslots) * ncellslots; slot += ncellslots) {
      const int i = slot + 1;
      const bool valid = slot < ncell;
      const int type = valid ? allowed_type(c, i, j) : 0;
      double acc = 0.;
      if (type) {
        if (sub == 0) acc = hairpin_w(P, G, c, w, motif_boltz, i, j, type);
        const int si1 = c.S[i + 1], sj1 = c.S[j - 1];
        // interior loops closed by (i,j) with inner pair (k,l): u1 = k-i-1, u2 = j-l-1, u1+u2 <= 30
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
            const double v = w.qb[tri(k, l)];
            if (v != 0.) {
              const int t2 = pair_type(c.S[k], c.S[l]);
              acc += v * w.scale[u1 + u2 + 2] * interior_x(P, G, u1, u2, type, rtype(t2), si1, sj1, c.S[k - 1], c.S[l + 1]);
            }
          }
        }
        // multiloop closed by (i,j): Qm(i+1,k-1) * Qm1(k,j-1); non-zero only for i+6 <= k <= j-5
        double ml = 0.;
        for (int k = i + 6 + sub; k <= j - 5; k += Gs) ml += w.qm[tri(i + 1, k - 1)] * qm1_prev[k];
        acc += ml * P.MLclosing * stem_x(P, rtype(type), true, sj1, true, si1, false) * P.scale2;
      }
      acc = ex.group_sum(acc);
      if (valid && sub == 0) {
        w.qb[tri(i, j)] = acc;
        double m1 = cuj ? qm1_prev[i] * eb1 : 0.;
        double e = 0.;
        if (type) {
          m1 += acc * stem_at(P, c, type, i, j, false);
          e = acc * stem_at(P, c, type, i, j, true) * w.q1[i - 1];
        }
        qm1_cur[i] = m1;
        w.ext[i] = e;
      }
    }
    ex.sync();
    // ---- phase C: Qm(i,j) = sum_{k=i..j-4} (U(i,k-1) + Qm(i,k-1)) * Qm1(k,j), U = all-unpaired weight
    for (int slot = tid / Gs; slot < ((ncell + ncellslots - 1) / ncellslots) * ncellslots; slot += ncellslots) {
      const int i = slot + 1;
      const bool valid = slot < ncell;
      double t = 0.;
      if (valid) {
        const int run = c.upr[i];
        for (int k = i + 1 + sub; k <= j - 4; k += Gs) {
          const int u = k - i;
          double wgt = w.qm[tri(i, k - 1)];
          if (u <= run) wgt += w.eMLb[u];
          t += wgt * qm1_cur[k];
        }
      }
      t = ex.group_sum(t);
      if (valid && sub == 0) w.qm[tri(i, j)] = t + qm1_cur[i];
    }
    // exterior prefix: Q(1,j) = [j unpaired] Q(1,j-1)*scale + sum_i Q(1,i-1) * Qb(i,j) * ExtStem(i,j)
    if (ex.warp() == 0) {
      double e = 0.;
      for (int i = 1 + ex.lane(); i <= ncell; i += E::kWarp) e += w.ext[i];
      e = ex.warp_sum(e);
      if (ex.lane() == 0) w.q1[j] = (cuj ? w.q1[j - 1] * s1 : 0.) + e;
    }
    ex.sync();
    double* tmp = qm1_prev; qm1_prev = qm1_cur; qm1_cur = tmp;
  }
  return w.q1[n];
}

// Outside pass: fills w.pr with base-pair probabilities P(i,j) (already multiplied by Qb and corrected for
// the motif's inner pairs).  Needs qb, qm, q1 from pf_inside.
template <class E>
ADP_HD void pf_outside(const E& ex, const SeqCtx& c, const PfSmall& P, const PfTables* G, const MotifDev* motif, double motif_boltz, PfWork& w) {
  const int n = c.n, nt = ex.nthreads(), tid = ex.tid();
  constexpr int Gs = E::kGroup;
  const int sub = tid % Gs, ncellslots = nt / Gs;
  const int ntri = n * (n + 1) / 2;
  const double s1 = P.scale1, eb1 = P.eMLb1;
  for (int k = tid; k < ntri; k += nt) w.pr[k] = 0.;
  for (int k = tid; k <= n + 1; k += nt) { w.va[k] = 0.; w.vb[k] = 0.; w.vab[k] = 0.; }
  // exterior suffix Q(k,n): position k is unpaired, or pairs with some l
  if (ex.warp() == 0) {
    if (ex.lane() == 0) w.qln[n + 1] = 1.;
    for (int k = n; k >= 1; k--) {
      double e = 0.;
      for (int l = k + 4 + ex.lane(); l <= n; l += E::kWarp) {
        const double v = w.qb[tri(k, l)];
        if (v != 0.) e += v * stem_at(P, c, pair_type(c.S[k], c.S[l]), k, l, true) * w.qln[l + 1];
      }
      e = ex.warp_sum(e);
      // all lanes keep the running value in registers would need a broadcast; lane 0 owns the array
      if (ex.lane() == 0) w.qln[k] = (can_unpair(c, k) ? w.qln[k + 1] * s1 : 0.) + e;
      ex.warp_sync();
    }
  }
  ex.sync();
  const double Z = w.q1[n];

  for (int l = n; l >= kTurn + 2; l--) {
    const int ncell = l - kTurn - 1;      // rows k = 1 .. l-4
    if (l < n) {
      // ---- step 1: multiloop carriers of column l, for every possible 5' closing position i <= l-5
      //   A(i,l) = MLclosing * sum_{j>=l+6} P(i,j) * MLstem'(i,j) * Qm(l+1,j-1)      (>= 1 stem right of l)
      //   B(i,l) = c(i,l+1) + [l+1 unpaired] * eMLb * B(i,l+1)                        (nothing right of l)
      const int nrow = l - 5;
      const bool cul1 = can_unpair(c, l + 1);
      for (int slot = tid / Gs; slot < ((nrow + ncellslots - 1) / ncellslots) * ncellslots; slot += ncellslots) {
        const int i = slot + 1;
        const bool valid = slot < nrow;
        double a = 0.;
        if (valid) {
          const int si1 = c.S[i + 1];
          for (int j = l + 6 + sub; j <= n; j += Gs) {
            const double v = w.pr[tri(i, j)];
            if (v != 0.) {
              const int tt = pair_type(c.S[i], c.S[j]);
              a += v * stem_x(P, rtype(tt), true, c.S[j - 1], true, si1, false) * w.qm[tri(l + 1, j - 1)];
            }
          }
        }
        a = ex.group_sum(a) * P.MLclosing;
        if (valid && sub == 0) {
          double c1 = 0.;
          const double v = w.pr[tri(i, l + 1)];
          if (v != 0.) {
            const int tt = pair_type(c.S[i], c.S[l + 1]);
            c1 = v * P.MLclosing * stem_x(P, rtype(tt), true, c.S[l], true, c.S[i + 1], false);
          }
          const double b = (cul1 ? w.vb[i] * eb1 : 0.) + c1;
          w.va[i] = a;
          w.vb[i] = b;
          w.vab[i] = a + b;
        }
      }
      ex.sync();
    }
    // ---- step 2: outside weight of every pair (k,l) of this column
    for (int slot = tid / Gs; slot < ((ncell + ncellslots - 1) / ncellslots) * ncellslots; slot += ncellslots) {
      const int k = slot + 1;
      const bool valid = slot < ncell;
      const int type = valid ? allowed_type(c, k, l) : 0;
      const bool live = type && w.qb[tri(k, l)] != 0.;
      double val = 0.;
      if (live) {
        if (sub == 0) val = w.q1[k - 1] * w.qln[l + 1] / Z * stem_at(P, c, type, k, l, true);
        const int t2r = rtype(type);
        const int sp1 = c.S[k - 1], sq1 = c.S[l + 1];
        // (k,l) enclosed by (i,j) in an interior loop
        int u1max = k - 2;
        if (u1max > kMaxLoop) u1max = kMaxLoop;
        if (u1max > (int)c.upl[k - 1]) u1max = c.upl[k - 1];
        const int u2run = c.upr[l + 1];
        for (int u1 = sub; u1 <= u1max; u1 += Gs) {
          const int i = k - 1 - u1;
          int u2max = n - l - 1;
          if (u2max > kMaxLoop - u1) u2max = kMaxLoop - u1;
          if (u2max > u2run) u2max = u2run;
          for (int u2 = 0; u2 <= u2max; u2++) {
            const int j = l + 1 + u2;
            const double v = w.pr[tri(i, j)];
            if (v != 0.) {
              const int tij = pair_type(c.S[i], c.S[j]);
              val += v * w.scale[u1 + u2 + 2] * interior_x(P, G, u1, u2, tij, t2r, c.S[i + 1], c.S[j - 1], sp1, sq1);
            }
          }
        }
        // (k,l) is a stem of a multiloop closed by some (i,j), i < k, j > l
        if (l < n && k >= 2) {
          double m = 0.;
          for (int i = 1 + sub; i <= k - 1; i += Gs) {
            const int u = k - 1 - i;                     // unpaired bases between i and k
            double t = 0.;
            if (u <= (int)c.upr[i + 1]) t = w.eMLb[u] * w.va[i];
            if (u >= 5) t += w.qm[tri(i + 1, k - 1)] * w.vab[i];
            m += t;
          }
          val += m * stem_x(P, type, true, sp1, true, sq1, false) * P.scale2;
        }
      }
      val = ex.group_sum(val);
      if (live && sub == 0) w.pr[tri(k, l)] = val;
    }
    ex.sync();
  }
  // P(k,l) = outside(k,l) * Qb(k,l); then the motif's inner pairs inherit the probability of the bound hairpin
  if (motif && motif->len > 0) {
    // stash the bound-hairpin probabilities before pr is overwritten (sites are few; reuse w.ext)
    for (int i = 1 + tid; i + motif->len - 1 <= n; i += nt) {
      double pm = 0.;
      if (c.motif_at[i]) {
        const int j = i + motif->len - 1;
        const int type = allowed_type(c, i, j);
        if (type && w.qb[tri(i, j)] != 0.) pm = w.pr[tri(i, j)] * hairpin_w(P, G, c, w, motif_boltz, i, j, type);
      }
      w.ext[i] = pm;
    }
    ex.sync();
  }
  for (int k = tid; k < ntri; k += nt) w.pr[k] *= w.qb[k];
  ex.sync();
  if (motif && motif->len > 0 && tid == 0) {
    for (int i = 1; i + motif->len - 1 <= n; i++) {
      if (!c.motif_at[i] || w.ext[i] == 0.) continue;
      for (int q = 0; q < motif->n_pairs; q++) w.pr[tri(i + motif->pair_i[q], i + motif->pair_j[q])] += w.ext[i];
    }
  }
  ex.sync();
}

}  // namespace adp
