// Fast McCaskill inside / outside pass, one construct per CTA (ACGU-only sequences, n <= fast_max_len(threads)).
// Same ensemble as pf_core.cuh (the general path) and as vrna_pf in the reference (src/scoring.cc:43,58,65);
// what differs is how the work is organised for the SM:
//
//  * Qb and Qm are TRIMMED column-major triangles (cells with j-i < 4 do not exist): 37 KB each at 100 nt.  The
//    outside pass overwrites Qb column by column with the outside weights (a finished column of Qb is only needed
//    through the exterior suffix, which is accumulated on the way), so both passes live in two matrices and two
//    constructs are resident per SM at 100 nt.
//  * Generic interior loops (both sides >= 2 unpaired: 372 of the 496 candidates of a closing pair) are NOT
//    enumerated.  Their weight factorises as  a(i,j) * I[s] * N[|u1-u2|] * b(k,l)  with s = u1+u2, so the
//    per-size sums  D_s(i,j) = sum_{u1+u2=s} N[|u1-u2|] b(k,l)  obey
//        D_s(i,j) = [i+1, j-1 unpaired] D_{s-2}(i+1,j-1) + N[s-4] * (the two candidates with u1 = 2 or u2 = 2)
//    i.e. 27 updates per cell instead of 372 multiply-adds.  (i,j) -> (i+1,j-1) keeps i+j, so a thread that owns
//    the anti-diagonal i+j keeps its 14 level sums IN REGISTERS while the column sweep walks along it: the
//    recursion costs no shared memory and no barrier.  The outside pass runs the mirrored recursion over the
//    enclosing pairs, (k,l) -> (k-1,l+1).
//  * The candidates on the loop's axes (bulges, 1xn loops: 112 per cell) are four 1-D filters; the two that run
//    down a column read a per-column vector premultiplied once, the two that run along a row are evaluated by the
//    cell's lanes.  The 9 corner candidates (stack, bulge-1, 1x1, 1x2, 2x2, 2x3) use the full tables.
//  * All per-pair mismatch factors come from 16x16 tables indexed by two one-byte position codes
//    (nucleotide + 3' neighbour, nucleotide + 5' neighbour), staged once per CTA.
//  * Only cells that can pair are evaluated: each column compacts them into a list, and a power-of-two lane group
//    per cell (sized so the list fills the CTA) splits the cell's terms and reduces with shuffles.
#pragma once
#include "fold_common.cuh"

namespace adp {

constexpr int kFastLevels = 14;       // level sums per thread: s = 4+par, 6+par, ..., 30+par

struct PfHot {                        // shared memory, staged once per CTA
  double FI[256], F1n[256], FM[256];  // mismatch_interior, mismatch_interior_1n, multiloop stem factor (closing orientation)
  double ip[32];                      // interior[s] * scale[s+2]
  double nin[32];                     // ninio factor of asymmetry d
  double tapB[32];                    // bulge[u] * scale[u+2]                      (u >= 2)
  double tap1n[32];                   // interior[u+1] * ninio[u-1] * scale[u+3]    (1 x u loop, u >= 3)
  double scale1, scale2, eMLb1, MLclosing, TermAU, pad;
};

template <class E>
ADP_HD void stage_hot(const E& ex, PfHot* H, const PfTables* T) {
  for (int q = ex.tid(); q < 256; q += ex.nthreads()) {
    const int a = q >> 4, b = q & 15;
    const int x = (a >> 2) + 1, u = (a & 3) + 1, y = (b >> 2) + 1, v = (b & 3) + 1;
    const int t = pair_type(x, y);
    H->FI[q] = t ? T->mmI[t][u][v] : 0.;
    H->F1n[q] = t ? T->mm1nI[t][u][v] : 0.;
    H->FM[q] = t ? T->mmM[rtype(t)][v][u] * (t > 2 ? T->TermAU : 1.) * T->MLintern : 0.;
  }
  for (int q = ex.tid(); q < 32; q += ex.nthreads()) {
    H->ip[q] = q <= 30 ? T->interior[q] * T->scale[q + 2] : 0.;
    H->nin[q] = q <= 30 ? T->ninio[q] : 0.;
    H->tapB[q] = (q >= 2 && q <= 30) ? T->bulge[q] * T->scale[q + 2] : 0.;
    H->tap1n[q] = (q >= 3 && q <= 29) ? T->interior[q + 1] * T->ninio[q - 1] * T->scale[q + 3] : 0.;
  }
  if (ex.tid() == 0) {
    H->scale1 = T->scale[1]; H->scale2 = T->scale[2]; H->eMLb1 = T->eMLb[1];
    H->MLclosing = T->MLclosing; H->TermAU = T->TermAU; H->pad = 0.;
  }
}

constexpr int kFastVectors = 16;

struct FastWork {
  double* qb;      // Qb, then (outside) the outside weights: trimmed triangle
  double* qm;      // Qm: trimmed triangle
  double* vbase;   // kFastVectors vectors of ld doubles each, contiguous; roles are named where they are used
  ADP_HD double* vecp(int v) const { return vbase + (size_t)v * ld; }
  const double* scaleG;        // pf_scale^-k (global)
  uint8_t* cn;     // [ld] code of (S[p], S[p+1])
  uint8_t* cp;     // [ld] code of (S[p], S[p-1])
  uint16_t* cells; // [ld] rows of the current column that can pair
  int* counter;    // [4]
  double* dstate;  // host emulation only: the level sums every "thread" keeps in registers  [2*R*kFastLevels]
  int ld;          // vector length (>= n + 4)
  int R;           // anti-diagonal slots (fast_R); the CTA needs 2*R threads
};

// trimmed column-major triangle: cell (i,j), 1 <= i <= j-4
ADP_HD int t4(int i, int j) { return (((j - 4) * (j - 5)) >> 1) + (i - 1); }
ADP_HD size_t t4_count(int n) { return n >= 5 ? (size_t)(n - 4) * (n - 3) / 2 : 0; }
ADP_HD int fast_ld(int n) { return (n + 6) & ~1; }
ADP_HD int fast_R(int n) { return n < 4 ? 4 : n; }      // anti-diagonal keys i+j <= 2n-4 < 2R, live window < R
ADP_HD int fast_max_len(int threads) { return threads / 2; }

ADP_HD size_t fast_small_bytes(int n) {
  const size_t ld = (size_t)fast_ld(n);
  return sizeof(PfHot) + 64 + ctx_bytes(n) + kFastVectors * ld * sizeof(double) + 2 * ld + 2 * ld + 64 + 16 * 12;
}
ADP_HD size_t fast_matrix_doubles(int n) { return 2 * t4_count(n) + 2; }

// weak pairs (AU, UA, GU, UG) pay the terminal penalty; C*G is the only nucleotide-code product equal to 6
ADP_HD double au_factor(const PfHot& H, int a, int b) { return a * b == 6 ? 1. : H.TermAU; }

// power-of-two lanes per cell such that m cells fill (at most) the CTA
ADP_HD int group_for(int m, int nthreads, int cap) {
  int g = 1;
  while (g < cap && m * (g << 1) <= nthreads) g <<= 1;
  return g;
}

// multiloop / exterior stem factor of the pair (i,j) seen from outside (neighbours i-1, j+1 dangle when they exist)
ADP_HD double fast_stem_ml(const SeqCtx& c, const PfHot& H, const PfTables* G, const FastWork& w, int type, int i, int j) {
  if (i > 1 && j < c.n) return H.FM[(w.cn[j] << 4) | w.cp[i]];
  return stem_x(*G, type, i > 1, c.S[i - 1], j < c.n, c.S[j + 1], false);
}
ADP_HD double fast_stem_ext(const SeqCtx& c, const PfTables* G, int type, int i, int j) {
  return stem_x(*G, type, i > 1, c.S[i - 1], j < c.n, c.S[j + 1], true);
}

ADP_HD double fast_hairpin(const SeqCtx& c, const PfTables* G, const FastWork& w, double motif_boltz, int i, int j, int type) {
  const int u = j - i - 1;
  if ((int)c.upr[i + 1] < u) return 0.;
  double q = hairpin_x(*G, G, c.S, i, j, type) * w.scaleG[u + 2];
  if (is_motif_site(c, i, j)) q *= motif_boltz;   // the bound motif REPLACES the empty loop
  return q;
}

// the 9 interior-loop candidates that go through the full tables: (u1,u2) =
// (0,0)(0,1) (1,0)(1,1)(1,2) (2,1)(2,2)(2,3) (3,2)
ADP_HD void corner_u(int q, int* u1, int* u2) {
  *u1 = q < 2 ? 0 : q < 5 ? 1 : q < 8 ? 2 : 3;
  *u2 = q < 2 ? q : q < 5 ? q - 2 : q < 8 ? q - 4 : 2;
}

template <class E>
ADP_HD bool fast_codes(const E& ex, const SeqCtx& c, FastWork& w) {
  const int n = c.n;
  bool ok = true;
  for (int p = ex.tid(); p < w.ld; p += ex.nthreads()) {
    int a = 0, b = 0;
    if (p >= 1 && p <= n) {
      if (c.S[p] == 0) ok = false;
      const int s = (c.S[p] - 1) & 3;
      if (p < n) a = (s << 2) | ((c.S[p + 1] - 1) & 3);
      if (p > 1) b = (s << 2) | ((c.S[p - 1] - 1) & 3);
    }
    w.cn[p] = (uint8_t)a;
    w.cp[p] = (uint8_t)b;
  }
  return ok;     // per thread; the caller combines
}

// One step of the level recursion for the thread that owns an anti-diagonal.  d[q] holds the sum of level
// s = 4 + par + 2q.  BORDER1 / BORDER2 add the two candidates with u1 = 2 / u2 = 2 to `border`; they differ
// between the inside and the outside pass.  The 14 levels are expanded explicitly, highest first, so that every
// index into d[] is a compile-time constant and the sums stay in registers.
#define ADP_LEVEL_STEP(Q, BORDER1, BORDER2)                                           \
  {                                                                                   \
    const int s = 4 + par + 2 * (Q);                                                  \
    double v = 0.;                                                                    \
    if (s <= smax) {                                                                  \
      double border = 0.;                                                             \
      BORDER1;                                                                        \
      if (s > 4) { BORDER2; }                                                         \
      v = H.nin[s - 4] * border;                                                      \
      if ((Q) > 0 && g) v += d[(Q) > 0 ? (Q) - 1 : 0];                                \
      if (s >= 6) gen += H.ip[s] * v;                                                 \
    }                                                                                 \
    d[Q] = v;                                                                         \
  }
#define ADP_LEVEL_LOOP(B1, B2)                                                                            \
  ADP_LEVEL_STEP(13, B1, B2) ADP_LEVEL_STEP(12, B1, B2) ADP_LEVEL_STEP(11, B1, B2) ADP_LEVEL_STEP(10, B1, B2) \
  ADP_LEVEL_STEP(9, B1, B2) ADP_LEVEL_STEP(8, B1, B2) ADP_LEVEL_STEP(7, B1, B2) ADP_LEVEL_STEP(6, B1, B2)     \
  ADP_LEVEL_STEP(5, B1, B2) ADP_LEVEL_STEP(4, B1, B2) ADP_LEVEL_STEP(3, B1, B2) ADP_LEVEL_STEP(2, B1, B2)     \
  ADP_LEVEL_STEP(1, B1, B2) ADP_LEVEL_STEP(0, B1, B2)
#define ADP_LEVEL_ZERO()                                                                                  \
  d[0] = d[1] = d[2] = d[3] = d[4] = d[5] = d[6] = d[7] = d[8] = d[9] = d[10] = d[11] = d[12] = d[13] = 0.;

// ---------------------------------------------------------------------------------------------------
// inside pass.  Returns Z = Q(1,n) (scaled) on every thread.
// vector roles: 0/1 Qm1 rolling columns, 2 exterior terms, 3 xA (Qb * terminal factor, column j-1),
// 4/5 x1n (Qb * 1xn mismatch, columns j-2 / j-1 by parity), 6/7/8 xI (Qb * generic mismatch, columns j-3..j-1 mod 3),
// 9 q1, 10 eMLb powers, 11/12 generic-loop sums per level parity.
template <class E>
ADP_HD double fast_inside(const E& ex, const SeqCtx& c, const PfHot& H, const PfTables* G, double motif_boltz, FastWork& w) {
  const int n = c.n, nt = ex.nthreads(), tid = ex.tid(), ld = w.ld, R = w.R;
  const int ntri = (int)t4_count(n);
  for (int k = tid; k < ntri; k += nt) { w.qb[k] = 0.; w.qm[k] = 0.; }
  for (int v = 0; v <= 12; v++)
    if (v != 10)
      for (int k = tid; k < ld; k += nt) w.vecp(v)[k] = 0.;
  double* q1 = w.vecp(9);
  const double* eMLbS = w.vecp(10);
  double* ext = w.vecp(2);
  double* xA = w.vecp(3);
  double d[kFastLevels];     // level sums of this thread's anti-diagonal (registers)
  ADP_LEVEL_ZERO()
  if constexpr (E::kSerial)
    for (int k = 0; k < 2 * R * kFastLevels; k++) w.dstate[k] = 0.;
  ex.sync();
  if (tid == 0) q1[0] = 1.;
  double* qm1_prev = w.vecp(0);
  double* qm1_cur = w.vecp(1);
  const double s1 = H.scale1, eb1 = H.eMLb1;

  for (int j = 1; j <= n; j++) {
    const int nr = j - 4;                       // rows 1 .. nr
    const bool cuj = can_unpair(c, j);
    if (nr >= 1) {
      double* x1n_rd = w.vecp(4 + (j & 1));      // column j-2; rewritten with column j at the end of the step
      double* xI_rd = w.vecp(6 + (j % 3));       // column j-3; rewritten with column j at the end of the step
      const int upl_j1 = c.upl[j - 1];
      const bool cuj1 = can_unpair(c, j - 1);
      // ---- phase D: level sums of the generic interior loops, one (anti-diagonal, parity) per thread
      for (int task = tid; task < 2 * R; task += nt) {
        if constexpr (E::kSerial)
          for (int q = 0; q < kFastLevels; q++) d[q] = w.dstate[(size_t)task * kFastLevels + q];
        const int par = task >= R ? 1 : 0;
        int key = task - par * R;
        if (key < j + 1) key += R;
        const int i = key - j;
        if (i >= 1 && i <= nr) {
          int smax = j - i - 6;
          if (smax > kMaxLoop) smax = kMaxLoop;
          double gen = 0.;
          if (smax >= 4) {
            const int upr_i1 = c.upr[i + 1];
            const bool g = upr_i1 >= 1 && cuj1;
            const int cpA = w.cp[i + 3];
            const bool ok1 = upr_i1 >= 2;
            ADP_LEVEL_LOOP(
                if (ok1 && upl_j1 >= s - 2) { const int l = j + 1 - s; border = w.qb[t4(i + 3, l)] * H.FI[(w.cn[l] << 4) | cpA]; },
                if (upr_i1 >= s - 2 && upl_j1 >= 2) border += xI_rd[i + s - 1];)
          } else {
            ADP_LEVEL_ZERO()
          }
          w.vecp(11 + par)[i] = gen;
        }
        if constexpr (E::kSerial)
          for (int q = 0; q < kFastLevels; q++) w.dstate[(size_t)task * kFastLevels + q] = d[q];
      }
      // this column's Qm1 starts from the previous one extended by the unpaired base j
      if (tid == 0) w.counter[0] = 0;
      for (int r = tid; r < nr; r += nt) { ext[r + 1] = 0.; qm1_cur[r + 1] = cuj ? qm1_prev[r + 1] * eb1 : 0.; }
      ex.sync();
      // compaction of the rows that can pair with j
      for (int r = tid; r < nr; r += nt)
        if (allowed_type(c, r + 1, j)) w.cells[ex.atomic_inc(w.counter)] = (uint16_t)(r + 1);
      ex.sync();
      // ---- phase A: Qb(i,j) for the listed cells, a lane group per cell
      const int m = w.counter[0];
      const int Gs = group_for(m > 0 ? m : 1, nt, E::kWarp);
      const int slots = nt / Gs, sub = tid % Gs;
      const double* x1n_prev = x1n_rd;
      for (int slot = tid / Gs; slot < ((m + slots - 1) / slots) * slots; slot += slots) {
        const bool valid = slot < m;
        const int i = valid ? w.cells[slot] : 1;
        double acc = 0.;
        int type = 0;
        if (valid) {
          type = pair_type(c.S[i], c.S[j]);
          const int si1 = c.S[i + 1], sj1 = c.S[j - 1];
          const int codeij = (w.cn[i] << 4) | w.cp[j];
          const int room = j - i - 6;                 // u1 + u2 <= room keeps the inner pair >= 4 apart
          int u1cap = c.upr[i + 1], u2cap = upl_j1;
          if (u1cap > kMaxLoop) u1cap = kMaxLoop;
          if (u2cap > kMaxLoop) u2cap = kMaxLoop;
          double other = 0., accB = 0., acc1n = 0., accML = 0.;
          if (sub == 0) {
            other = fast_hairpin(c, G, w, motif_boltz, i, j, type);
            other += H.FI[codeij] * (w.vecp(11)[i] + w.vecp(12)[i]);
          }
          if (room >= 0) {
            // corner candidates through the full tables
            for (int q = sub; q < 9; q += Gs) {
              int u1, u2;
              corner_u(q, &u1, &u2);
              if (u1 <= u1cap && u2 <= u2cap && u1 + u2 <= room) {
                const int k = i + 1 + u1, l = j - 1 - u2;
                const double v = w.qb[t4(k, l)];
                if (v != 0.)
                  other += v * w.scaleG[u1 + u2 + 2] * interior_x(*G, G, u1, u2, type, rtype(pair_type(c.S[k], c.S[l])), si1, sj1, c.S[k - 1], c.S[l + 1]);
              }
            }
            // bulges with the unpaired run on the 3' side: k = i+1, l = j-1-u2, u2 >= 2
            {
              int hi = room < u2cap ? room : u2cap;
              const int k = i + 1, sk = c.S[k];
              for (int u2 = 2 + sub; u2 <= hi; u2 += Gs) {
                const int l = j - 1 - u2;
                accB += w.qb[t4(k, l)] * au_factor(H, sk, c.S[l]) * H.tapB[u2];
              }
            }
            // 1 x u2 loops: k = i+2, u2 >= 3
            if (u1cap >= 1) {
              int hi = room - 1 < u2cap ? room - 1 : u2cap;
              if (hi > kMaxLoop - 1) hi = kMaxLoop - 1;
              const int k = i + 2, cpk = w.cp[k];
              for (int u2 = 3 + sub; u2 <= hi; u2 += Gs) {
                const int l = j - 1 - u2;
                acc1n += w.qb[t4(k, l)] * H.F1n[(w.cn[l] << 4) | cpk] * H.tap1n[u2];
              }
            }
            // bulges with the run on the 5' side: l = j-1, k = i+1+u1, u1 >= 2 -- premultiplied column vector
            {
              int hi = room < u1cap ? room : u1cap;
              for (int u1 = 2 + sub; u1 <= hi; u1 += Gs) accB += xA[i + 1 + u1] * H.tapB[u1];
            }
            // u1 x 1 loops: l = j-2, u1 >= 3
            if (u2cap >= 1) {
              int hi = room - 1 < u1cap ? room - 1 : u1cap;
              if (hi > kMaxLoop - 1) hi = kMaxLoop - 1;
              for (int u1 = 3 + sub; u1 <= hi; u1 += Gs) acc1n += x1n_prev[i + 1 + u1] * H.tap1n[u1];
            }
          }
          // multiloop closed by (i,j): Qm(i+1,k-1) * Qm1(k,j-1)
          for (int k = i + 6 + sub; k <= j - 5; k += Gs) accML += w.qm[t4(i + 1, k - 1)] * qm1_prev[k];
          acc = other + au_factor(H, c.S[i], c.S[j]) * accB + H.F1n[codeij] * acc1n + accML * (H.MLclosing * H.FM[codeij] * H.scale2);
        }
        acc = ex.group_sum_rt(acc, Gs);
        if (valid && sub == 0) {
          w.qb[t4(i, j)] = acc;
          qm1_cur[i] += acc * fast_stem_ml(c, H, G, w, type, i, j);
          ext[i] = acc * fast_stem_ext(c, G, type, i, j) * q1[i - 1];
        }
      }
      ex.sync();
      // ---- phase C: premultiplied vectors of the new column; Qm(i,j) for every row (rows paired so that every
      //      lane group sums the same number of terms); exterior prefix
      {
        double* x1n_wr = w.vecp(4 + (j & 1));
        double* xI_wr = w.vecp(6 + (j % 3));
        const int cnj = w.cn[j], sj = c.S[j];
        for (int r = tid; r < nr; r += nt) {
          const int i = r + 1;
          const double v = w.qb[t4(i, j)];
          const int code = (cnj << 4) | w.cp[i];
          xA[i] = v * au_factor(H, c.S[i], sj);
          x1n_wr[i] = v * H.F1n[code];
          xI_wr[i] = v * H.FI[code];
        }
        const int npair = (nr + 1) >> 1;
        const int Gc = group_for(npair, nt, E::kWarp);
        const int slotsC = nt / Gc, subC = tid % Gc;
        for (int slot = tid / Gc; slot < ((npair + slotsC - 1) / slotsC) * slotsC; slot += slotsC) {
          const bool valid = slot < npair;
          const int iA = slot + 1, iB = nr - slot;       // iA <= iB; equal for the middle row of an odd count
          double tA = 0., tB = 0.;
          if (valid) {
            const int lenA = nr - iA, lenB = iB > iA ? nr - iB : 0;
            const int runA = c.upr[iA], runB = c.upr[iB];
            for (int t = subC; t < lenA + lenB; t += Gc) {
              const bool inA = t < lenA;
              const int i = inA ? iA : iB;
              const int u = inA ? t + 1 : t - lenA + 1;          // k = i + u
              const int k = i + u;
              double wgt = u >= 5 ? w.qm[t4(i, k - 1)] : 0.;
              if (u <= (inA ? runA : runB)) wgt += eMLbS[u];
              const double term = wgt * qm1_cur[k];
              if (inA) tA += term; else tB += term;
            }
          }
          tA = ex.group_sum_rt(tA, Gc);
          tB = ex.group_sum_rt(tB, Gc);
          if (valid && subC == 0) {
            w.qm[t4(iA, j)] = tA + qm1_cur[iA];
            if (iB > iA) w.qm[t4(iB, j)] = tB + qm1_cur[iB];
          }
        }
      }
    }
    if (ex.warp() == 0) {
      double e = 0.;
      for (int i = 1 + ex.lane(); i <= nr; i += E::kWarp) e += ext[i];
      e = ex.warp_sum(e);
      if (ex.lane() == 0) q1[j] = (cuj ? q1[j - 1] * s1 : 0.) + e;
    }
    ex.sync();
    double* t0 = qm1_prev; qm1_prev = qm1_cur; qm1_cur = t0;
  }
  return q1[n];
}

// ---------------------------------------------------------------------------------------------------
// outside pass.  Overwrites Qb column by column (l = n .. 5) with the outside weights; the probabilities of a
// column go to `bppm` / the pair queries as soon as the column is complete.
// vector roles: 0 A(i,l) (multiloop carrier with >= 1 stem right of l), 1 B(i,l) (nothing right of l),
// 2 P(.,l) of the finished column, 3 xA (Pout * terminal factor, column l+1), 4/5 x1n (columns l+2 / l+1 by parity),
// 6/7/8 xI (columns l+3..l+1 mod 3), 9 q1, 10 eMLb powers, 11/12 generic sums, 13 qln, 14 exterior-suffix
// accumulators, 15 bound-motif probabilities per site.
struct FastOut {
  double* bppm;          // NULL or n(n+1)/2 doubles, row-major upper triangle incl. diagonal, 0-based
  int n_query;
  const int32_t* qi;
  const int32_t* qj;
  double* qprobs;
};

ADP_HD size_t bppm_index(int n, int i, int j) {      // 1-based i <= j
  return (size_t)(i - 1) * n - (size_t)(i - 1) * (i - 2) / 2 + (size_t)(j - i);
}

template <class E>
ADP_HD void fast_outside(const E& ex, const SeqCtx& c, const PfHot& H, const PfTables* G, const MotifDev* motif, double motif_boltz,
                         FastWork& w, const FastOut& out) {
  const int n = c.n, nt = ex.nthreads(), tid = ex.tid(), ld = w.ld, R = w.R;
  const int outvecs[] = {0, 1, 2, 3, 4, 5, 6, 7, 8, 11, 12, 14, 15};
  for (int v : outvecs)
    for (int k = tid; k < ld; k += nt) w.vecp(v)[k] = 0.;
  double* vA = w.vecp(0);
  double* vB = w.vecp(1);
  double* pcol = w.vecp(2);
  double* xA = w.vecp(3);
  const double* q1 = w.vecp(9);
  const double* eMLbS = w.vecp(10);
  double* qln = w.vecp(13);
  double* accq = w.vecp(14);
  double* pm = w.vecp(15);
  if (out.bppm) {
    const size_t ntot = (size_t)n * (n + 1) / 2;
    for (size_t k = tid; k < ntot; k += nt) out.bppm[k] = 0.;
  }
  for (int q = tid; q < out.n_query; q += nt) out.qprobs[q] = 0.;
  double d[kFastLevels];     // level sums of this thread's anti-diagonal (registers)
  ADP_LEVEL_ZERO()
  if constexpr (E::kSerial)
    for (int k = 0; k < 2 * R * kFastLevels; k++) w.dstate[k] = 0.;
  if (tid == 0) { qln[n + 1] = 1.; qln[n + 2] = 0.; }
  ex.sync();
  const double Z = q1[n];
  const double invZ = 1. / Z;
  const double s1 = H.scale1, eb1 = H.eMLb1;
  const bool has_motif = motif && motif->len > 0;

  for (int l = n; l >= kTurn + 2; l--) {
    const int nr = l - 4;                       // rows k = 1 .. nr
    double* x1n_rd = w.vecp(4 + (l & 1));        // column l+2
    double* xI_rd = w.vecp(6 + (l % 3));         // column l+3
    const int upr_l1 = c.upr[l + 1];            // upr[n+1] = 0
    // ---- step 1: everything that only needs the finished columns > l
    // (a) exterior suffix Q(l,n): position l unpaired, or paired with some l' >= l+4 (already accumulated)
    if (tid == 0) qln[l] = (can_unpair(c, l) ? qln[l + 1] * s1 : 0.) + accq[l];
    // (b) multiloop carriers for every possible 5' closing position i <= l-5
    if (l < n) {
      const int nrow = l - 5;
      const bool cul1 = can_unpair(c, l + 1);
      const int Ga = group_for(nrow > 0 ? nrow : 1, nt, E::kWarp);
      const int slotsA = nt / Ga, subA = tid % Ga;
      const int cpl1 = w.cp[l + 1];
      for (int slot = tid / Ga; slot < ((nrow + slotsA - 1) / slotsA) * slotsA; slot += slotsA) {
        const int i = slot + 1;
        const bool valid = slot < nrow;
        double a = 0.;
        if (valid) {
          const int cni = w.cn[i] << 4;
          for (int j = l + 6 + subA; j <= n; j += Ga)
            a += w.qb[t4(i, j)] * H.FM[cni | w.cp[j]] * w.qm[t4(l + 1, j - 1)];
        }
        a = ex.group_sum_rt(a, Ga) * H.MLclosing;
        if (valid && subA == 0) {
          const double c1 = w.qb[t4(i, l + 1)] * H.MLclosing * H.FM[(w.cn[i] << 4) | cpl1];
          vB[i] = (cul1 ? vB[i] * eb1 : 0.) + c1;
          vA[i] = a;
        }
      }
    }
    // (c) level sums of the generic interior loops that enclose (k,l)
    for (int task = tid; task < 2 * R; task += nt) {
      if constexpr (E::kSerial)
        for (int q = 0; q < kFastLevels; q++) d[q] = w.dstate[(size_t)task * kFastLevels + q];
      const int par = task >= R ? 1 : 0;
      int key = task - par * R;
      if (key < l + 1) key += R;
      const int k = key - l;
      if (k >= 1 && k <= nr) {
        int smax = (k - 2) + (n - l - 1);          // u1 <= k-2, u2 <= n-l-1
        if (smax > kMaxLoop) smax = kMaxLoop;
        double gen = 0.;
        if (k >= 4 && l + 3 <= n && smax >= 4) {
          const int upl_k1 = c.upl[k - 1];
          const bool g = upl_k1 >= 1 && upr_l1 >= 1;
          const int cnA = w.cn[k - 3] << 4;
          const bool ok1 = upl_k1 >= 2;
          ADP_LEVEL_LOOP(
              if (ok1 && upr_l1 >= s - 2 && l + s - 1 <= n) { const int j = l + s - 1; border = w.qb[t4(k - 3, j)] * H.FI[cnA | w.cp[j]]; },
              if (upl_k1 >= s - 2 && upr_l1 >= 2 && k + 1 - s >= 1) border += xI_rd[k + 1 - s];)
        } else {
          ADP_LEVEL_ZERO()
        }
        w.vecp(11 + par)[k] = gen;
      }
      if constexpr (E::kSerial)
        for (int q = 0; q < kFastLevels; q++) w.dstate[(size_t)task * kFastLevels + q] = d[q];
    }
    // (d) the cells of this column that carry weight
    if (tid == 0) w.counter[0] = 0;
    for (int r = tid; r < nr + 4; r += nt) pcol[r + 1] = 0.;
    ex.sync();
    for (int r = tid; r < nr; r += nt)
      if (w.qb[t4(r + 1, l)] != 0.) w.cells[ex.atomic_inc(w.counter)] = (uint16_t)(r + 1);
    ex.sync();
    // ---- step 2: outside weight of every listed pair (k,l)
    {
      const int m = w.counter[0];
      const int Gs = group_for(m > 0 ? m : 1, nt, E::kWarp);
      const int slots = nt / Gs, sub = tid % Gs;
      const int sq1 = c.S[l + 1];
      const int room2 = n - l - 1;                // u2 <= room2 keeps j <= n
      int u2cap = upr_l1;
      if (u2cap > kMaxLoop) u2cap = kMaxLoop;
      if (u2cap > room2) u2cap = room2;
      for (int slot = tid / Gs; slot < ((m + slots - 1) / slots) * slots; slot += slots) {
        const bool valid = slot < m;
        const int k = valid ? w.cells[slot] : 1;
        double val = 0.;
        int type = 0;
        if (valid) {
          type = pair_type(c.S[k], c.S[l]);
          const int sp1 = c.S[k - 1];
          const int codekl = (w.cn[l] << 4) | w.cp[k];          // inner orientation
          int u1cap = c.upl[k - 1];
          if (u1cap > kMaxLoop) u1cap = kMaxLoop;
          if (u1cap > k - 2) u1cap = k - 2;                     // i = k-1-u1 >= 1
          double other = 0., accB = 0., acc1n = 0., accML = 0.;
          if (sub == 0) {
            other = q1[k - 1] * qln[l + 1] * invZ * fast_stem_ext(c, G, type, k, l);
            other += H.FI[codekl] * (w.vecp(11)[k] + w.vecp(12)[k]);
          }
          if (k >= 2 && l < n) {
            const int t2r = rtype(type);
            for (int q = sub; q < 9; q += Gs) {
              int u1, u2;
              corner_u(q, &u1, &u2);
              if (u1 <= u1cap && u2 <= u2cap && u1 + u2 <= kMaxLoop) {
                const int i = k - 1 - u1, j = l + 1 + u2;
                const double v = w.qb[t4(i, j)];
                if (v != 0.)
                  other += v * w.scaleG[u1 + u2 + 2] * interior_x(*G, G, u1, u2, pair_type(c.S[i], c.S[j]), t2r, c.S[i + 1], c.S[j - 1], sp1, sq1);
              }
            }
            // enclosing bulges with the run on the 3' side: i = k-1, j = l+1+u2, u2 >= 2
            {
              const int i = k - 1, si = c.S[i];
              for (int u2 = 2 + sub; u2 <= u2cap; u2 += Gs) {
                const int j = l + 1 + u2;
                accB += w.qb[t4(i, j)] * au_factor(H, si, c.S[j]) * H.tapB[u2];
              }
            }
            // enclosing 1 x u2 loops: i = k-2, u2 >= 3
            if (u1cap >= 1) {
              const int hi = u2cap < kMaxLoop - 1 ? u2cap : kMaxLoop - 1;
              const int i = k - 2, cni = w.cn[i] << 4;
              for (int u2 = 3 + sub; u2 <= hi; u2 += Gs) {
                const int j = l + 1 + u2;
                acc1n += w.qb[t4(i, j)] * H.F1n[cni | w.cp[j]] * H.tap1n[u2];
              }
            }
            // run on the 5' side: j = l+1, i = k-1-u1, u1 >= 2 -- premultiplied column vector
            for (int u1 = 2 + sub; u1 <= u1cap; u1 += Gs) accB += xA[k - 1 - u1] * H.tapB[u1];
            // u1 x 1 loops: j = l+2
            if (u2cap >= 1) {
              const int hi = u1cap < kMaxLoop - 1 ? u1cap : kMaxLoop - 1;
              for (int u1 = 3 + sub; u1 <= hi; u1 += Gs) acc1n += x1n_rd[k - 1 - u1] * H.tap1n[u1];
            }
            // (k,l) is a stem of a multiloop closed by (i,j), i < k, j > l
            for (int i = 1 + sub; i <= k - 1; i += Gs) {
              const int u = k - 1 - i;                         // unpaired bases between i and k
              double t = 0.;
              if (u <= (int)c.upr[i + 1]) t = eMLbS[u] * vA[i];
              if (u >= 5) t += w.qm[t4(i + 1, k - 1)] * (vA[i] + vB[i]);
              accML += t;
            }
          }
          val = other + au_factor(H, c.S[k], c.S[l]) * accB + H.F1n[codekl] * acc1n;
          if (k >= 2 && l < n) val += accML * (H.FM[codekl] * H.scale2);
        }
        val = ex.group_sum_rt(val, Gs);
        if (valid && sub == 0) {
          const double qbv = w.qb[t4(k, l)];
          pcol[k] = val * qbv;
          accq[k] += qbv * fast_stem_ext(c, G, type, k, l) * qln[l + 1];
          if (has_motif && is_motif_site(c, k, l)) pm[k] = val * fast_hairpin(c, G, w, motif_boltz, k, l, type);
          w.qb[t4(k, l)] = val;
        }
      }
    }
    ex.sync();
    // ---- step 3: premultiplied vectors of the finished column; results
    {
      double* x1n_wr = w.vecp(4 + (l & 1));
      double* xI_wr = w.vecp(6 + (l % 3));
      const int cpl = w.cp[l], sl = c.S[l];
      for (int r = tid; r < nr + 3; r += nt) {
        const int k = r + 1;
        double v = 0.;
        int code = 0;
        if (k <= nr) { v = w.qb[t4(k, l)]; code = (w.cn[k] << 4) | cpl; }     // outer orientation
        xA[k] = v * au_factor(H, c.S[k], sl);
        x1n_wr[k] = v * H.F1n[code];
        xI_wr[k] = v * H.FI[code];
      }
      if (out.bppm)
        for (int r = tid; r < nr; r += nt)
          if (pcol[r + 1] != 0.) out.bppm[bppm_index(n, r + 1, l)] = pcol[r + 1];
      for (int q = tid; q < out.n_query; q += nt) {
        int a = out.qi[q], b = out.qj[q];
        if (a > b) { const int t = a; a = b; b = t; }
        if (b + 1 == l && a >= 0 && a + 1 <= nr) out.qprobs[q] = pcol[a + 1];
      }
    }
    ex.sync();
  }
  // the motif's inner pairs inherit the probability of the bound hairpin
  if (has_motif && tid == 0) {
    for (int i = 1; i + motif->len - 1 <= n; i++) {
      if (!c.motif_at[i] || pm[i] == 0.) continue;
      for (int q = 0; q < motif->n_pairs; q++) {
        const int a = i + motif->pair_i[q], b = i + motif->pair_j[q];
        if (out.bppm) out.bppm[bppm_index(n, a, b)] += pm[i];
        for (int z = 0; z < out.n_query; z++) {
          int qa = out.qi[z], qb_ = out.qj[z];
          if (qa > qb_) { const int t = qa; qa = qb_; qb_ = t; }
          if (qa + 1 == a && qb_ + 1 == b) out.qprobs[z] += pm[i];
        }
      }
    }
  }
  ex.sync();
}

}  // namespace adp
