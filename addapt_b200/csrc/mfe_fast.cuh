// Fast Zuker MFE fill, one construct per CTA: the inside pass of pf_fast.cuh transcribed to the (min, +) semiring over
// int32 dcal/mol.  Same organisation -- trimmed column-major triangles C and F (19 KB each at 100 nt), C stored with the
// generic interior mismatch of the pair already added, the generic interior loops through the register-resident level
// recursion  D_s(i,j) = min([i+1, j-1 unpaired] D_{s-2}(i+1,j-1), N[s-4] + min(two border candidates)), the loop's axes as
// four 1-D filters, compacted cell lists, two barriers per column -- and the same results as mfe_core.cuh, because integer
// min / + are exact whatever the order.  "No structure" is kInf; sums that contain it are clamped back to kInf when a cell
// is stored, so at most three kInf-sized terms are ever added (no overflow) and every stored value is either a real energy
// or exactly kInf.  The traceback is mfe_core.cuh's, reading the matrices through FastMfeView.
//
// Replaces vrna_mfe (no call site in the reference; BASELINE.json configs[1]); the oracle (oracle/fold_oracle.c orc_mfe)
// fixes recurrences and tie-break order.
#pragma once
#include "mfe_core.cuh"
#include "pf_fast.cuh"

namespace adp {

constexpr int kInfHalf = kInf / 2;    // anything above is "no structure"

struct MfeHot {                       // shared memory, staged once per job
  int32_t FI[256], F1n[256], FM[256]; // mismatch_interior, mismatch_interior_1n, multiloop stem energy (closing orientation)
  int32_t RI[256], RB[256], R1[256];  // -FI, terminal-AU - FI, F1n - FI: what readers add to a stored C + FI
  int32_t tapB[32];                   // bulge[u]                      (u >= 2)
  int32_t tap1n[32];                  // interior[u+1] + ninio[u-1]    (1 x u loop, u >= 3)
  int32_t ML_base, ML_closing, TermAU, inf;
};

struct FastConstI {
  int32_t ip[32];                     // interior[s]
  int32_t nin[32];                    // ninio(d)
};

ADP_HD void fill_fast_const_i(FastConstI* fc, const IntTables* T) {
  for (int q = 0; q < 32; q++) {
    fc->ip[q] = q <= 30 ? T->interior[q] : kInf;
    fc->nin[q] = q <= 30 ? T->ninio[q] : kInf;
  }
}

#if defined(__CUDACC__)
static __constant__ FastConstI c_fast_const_i;
#endif
static FastConstI h_fast_const_i;
#if defined(__CUDA_ARCH__)
#define ADP_FCI c_fast_const_i
#else
#define ADP_FCI h_fast_const_i
#endif

template <class E>
ADP_HD void stage_mfe_hot(const E& ex, MfeHot* H, const IntTables* T) {
  for (int q = ex.tid(); q < 256; q += ex.nthreads()) {
    const int a = q >> 4, b = q & 15;
    const int x = (a >> 2) + 1, u = (a & 3) + 1, y = (b >> 2) + 1, v = (b & 3) + 1;
    const int t = pair_type(x, y);
    const int fi = t ? T->mmI[t][u][v] : 0;
    H->FI[q] = t ? fi : kInf;
    H->F1n[q] = t ? T->mm1nI[t][u][v] : kInf;
    H->FM[q] = t ? T->mmM[rtype(t)][v][u] + (t > 2 ? T->TerminalAU : 0) + T->ML_intern : kInf;
    H->RI[q] = t ? -fi : kInf;
    H->RB[q] = t ? (t > 2 ? T->TerminalAU : 0) - fi : kInf;
    H->R1[q] = t ? T->mm1nI[t][u][v] - fi : kInf;
  }
  for (int q = ex.tid(); q < 32; q += ex.nthreads()) {
    H->tapB[q] = (q >= 2 && q <= 30) ? T->bulge[q] : kInf;
    H->tap1n[q] = (q >= 3 && q <= 29) ? T->interior[q + 1] + T->ninio[q - 1] : kInf;
  }
  if (ex.tid() == 0) { H->ML_base = T->ML_base; H->ML_closing = T->ML_closing; H->TermAU = T->TerminalAU; H->inf = kInf; }
#if !defined(__CUDA_ARCH__)
  fill_fast_const_i(&h_fast_const_i, T);
#endif
}

constexpr int kMfeFastVectors = 9;    // int32 vectors of ld entries: fm1 x2, ext, xA, x1n x2, f5, gen x2

struct MfeFastWork {
  int32_t* C;      // C(i,j) + FI(i,j) (inner orientation; plain at the sequence ends), kInf = cannot pair: trimmed triangle
  int32_t* F;      // F(i,j): >= 1 stem in [i,j] inside a multiloop: trimmed triangle
  int32_t* vbase;
  ADP_HD int32_t* vecp(int v) const { return vbase + (size_t)v * ld; }
  uint8_t* cn;
  uint8_t* cp;
  uint16_t* cells;
  int* counter;
  int32_t* dstate; // see FastWork::dstate
  int ld, R;
};

ADP_HD size_t mfe_fast_small_bytes(int n) {
  const size_t ld = (size_t)fast_ld(n);
  return sizeof(MfeHot) + 64 + ctx_bytes(n) + kMfeFastVectors * ld * sizeof(int32_t) + 2 * ld + 2 * ld + 64 + 16 * 12
         + (size_t)(4 * n + 16) * 3 * sizeof(int16_t) + (n + 16) + 64;      // + traceback stack and output string
}
ADP_HD size_t mfe_fast_matrix_ints(int n) { return 2 * t4_count(n) + 4; }

ADP_HD int imin(int a, int b) { return a < b ? a : b; }
ADP_HD int au_e(const MfeHot& H, int a, int b) { return a * b == 6 ? 0 : H.TermAU; }

ADP_HD int mfe_stem_ml(const SeqCtx& c, const MfeHot& H, const IntTables& P, const MfeFastWork& w, int type, int i, int j) {
  if (i > 1 && j < c.n) return H.FM[(w.cn[j] << 4) | w.cp[i]];
  return stem_e(P, type, i > 1, c.S[i - 1], j < c.n, c.S[j + 1], false);
}

// matrix view for the traceback: plain C(i,j), F(i,j), kInf where the trimmed triangles have no cell
struct FastMfeView {
  const int32_t* C;
  const int32_t* F;
  const int32_t* FI;
  const uint8_t* cn;
  const uint8_t* cp;
  int n;
  ADP_HD int c(int i, int j) const {
    if (j - i < 4) return kInf;
    const int v = C[t4(i, j)];
    if (v >= kInfHalf) return kInf;
    return (i > 1 && j < n) ? v - FI[(cn[j] << 4) | cp[i]] : v;
  }
  ADP_HD int f(int i, int j) const { return j - i < 4 ? kInf : F[t4(i, j)]; }
};

// One level of the (min,+) recursion; see ADP_LEVEL_STEP in pf_fast.cuh.  `zp` points at an always-kInf word.
#define ADP_MLEVEL_STEP(PAR, Q, P1, C1, P2, C2)                                        \
  if (4 + (PAR) + 2 * (Q) <= kMaxLoop) {                                              \
    constexpr int S = 4 + (PAR) + 2 * (Q);                                            \
    const bool on = S <= smax;                                                        \
    const int32_t* p1 = (on && (C1)) ? (P1) : zp;                                     \
    const int32_t* p2 = (S > 4 && on && (C2)) ? (P2) : zp;                            \
    int v = imin(ADP_FCI.nin[S - 4 > 0 ? S - 4 : 0] + imin(*p1, *p2), kInf);          \
    if ((Q) > 0) v = imin(v, (on && g) ? d[(Q) > 0 ? (Q) - 1 : 0] : kInf);            \
    if (S >= 6) gen = imin(gen, ADP_FCI.ip[S] + v);                                   \
    d[Q] = v;                                                                         \
  }
#define ADP_MLEVEL_PASS(PAR, B1, C1, B2, C2)                                                        \
  ADP_MLEVEL_STEP(PAR, 13, B1, C1, B2, C2) ADP_MLEVEL_STEP(PAR, 12, B1, C1, B2, C2)                \
  ADP_MLEVEL_STEP(PAR, 11, B1, C1, B2, C2) ADP_MLEVEL_STEP(PAR, 10, B1, C1, B2, C2)                \
  ADP_MLEVEL_STEP(PAR, 9, B1, C1, B2, C2) ADP_MLEVEL_STEP(PAR, 8, B1, C1, B2, C2)                  \
  ADP_MLEVEL_STEP(PAR, 7, B1, C1, B2, C2) ADP_MLEVEL_STEP(PAR, 6, B1, C1, B2, C2)                  \
  ADP_MLEVEL_STEP(PAR, 5, B1, C1, B2, C2) ADP_MLEVEL_STEP(PAR, 4, B1, C1, B2, C2)                  \
  ADP_MLEVEL_STEP(PAR, 3, B1, C1, B2, C2) ADP_MLEVEL_STEP(PAR, 2, B1, C1, B2, C2)                  \
  ADP_MLEVEL_STEP(PAR, 1, B1, C1, B2, C2) ADP_MLEVEL_STEP(PAR, 0, B1, C1, B2, C2)
#define ADP_MLEVEL_INF() \
  d[0] = d[1] = d[2] = d[3] = d[4] = d[5] = d[6] = d[7] = d[8] = d[9] = d[10] = d[11] = d[12] = d[13] = kInf;

// this lane's share of  min_{k=i+1}^{nr} ( min(F(i,k-1), (k-i)*ML_base [k-i <= run]) + Fm1(k,j) )
ADP_HD int f_row_terms(const int32_t* F, const int32_t* fm1, int mlb, int i, int nr, int run, int sub, int G) {
  int best = kInf;
  {
    int umax = nr - i;
    if (umax > run) umax = run;
    for (int u = 1 + sub; u <= umax; u += G) best = imin(best, u * mlb + fm1[i + u]);
  }
  {
    int k = i + 5 + sub;
    int off = t4(i, k - 1);
    int step = G * (k - 5) + ((G * (G - 1)) >> 1);
    const int step2 = G * G;
    for (; k <= nr; k += G) {
      best = imin(best, F[off] + fm1[k]);
      off += step;
      step += step2;
    }
  }
  return best;
}

// vector roles: 0/1 Fm1 columns by parity, 2 exterior terms of the listed cells, 3 xA (C + terminal-AU, column j-1),
// 4/5 x1n (C + 1xn mismatch, by column parity), 6 f5, 7/8 generic-loop minima per level parity.
template <bool kCst, class E>
ADP_HD int mfe_fast_fill(const E& ex, const SeqCtx& c, const MfeHot& H, const IntTables& P, int motif_dcal, MfeFastWork& w) {
  const int n = c.n, nt = ex.nthreads(), tid = ex.tid(), ld = w.ld, R = w.R;
  const int nt_lg = ilog2_floor((unsigned)nt), warp_lg = ilog2_floor((unsigned)E::kWarp);
  const bool mem_state = E::kSerial || w.dstate != nullptr;
  const int half = mem_state ? R : (nt >> 1);
  const int ntri = (int)t4_count(n);
  int32_t* const C = w.C;
  int32_t* const F = w.F;
  const uint8_t* const cn = w.cn;
  const uint8_t* const cp = w.cp;
  for (int k = tid; k < ntri; k += nt) { C[k] = kInf; F[k] = kInf; }
  for (int k = tid; k < kMfeFastVectors * ld; k += nt) w.vbase[k] = kInf;
  int32_t* const ext = w.vecp(2);
  int32_t* const xA = w.vecp(3);
  int32_t* const f5 = w.vecp(6);
  int32_t* const gen0 = w.vecp(7);
  int32_t* const gen1 = w.vecp(8);
  int d[kFastLevels];
  const int32_t* const zp = &H.inf;
  ADP_MLEVEL_INF()
  if (mem_state)
    for (int k = tid; k < 2 * R * kFastLevels; k += nt) w.dstate[k] = kInf;
  if (tid == 0) { w.counter[0] = 0; w.counter[1] = 0; }
  ex.sync();
  const int mlb = H.ML_base;
  if (tid == 0) {
    f5[0] = 0;
    for (int j = 1; j <= 4 && j <= n; j++) f5[j] = ((!kCst || can_unpair(c, j)) && f5[j - 1] < kInfHalf) ? f5[j - 1] : kInf;
  }

  auto prepare_column = [&](int jn) {
    const int nr = jn - 4;
    const bool cuj = !kCst || can_unpair(c, jn);
    const bool cuj1 = !kCst || can_unpair(c, jn - 1);
    const int upl_j1 = kCst ? (int)c.upl[jn - 1] : 1000;
    const int32_t* col3 = C + (((jn - 7) * (jn - 8)) >> 1) - 2;   // C+FI(k, jn-3) = col3[k + 1]
    const int32_t* fm1_prev = w.vecp((jn - 1) & 1);
    int32_t* fm1_cur = w.vecp(jn & 1);
    int* counter = w.counter + (jn & 1);
    for (int task = tid; task < 2 * half; task += nt) {
      const int par = task >= half ? 1 : 0;
      int key = task - par * half;
      if (key >= R) continue;
      if (mem_state)
        for (int q = 0; q < kFastLevels; q++) d[q] = w.dstate[(size_t)task * kFastLevels + q];
      if (key < jn + 1) key += R;
      const int i = key - jn;
      if (i >= 1 && i <= nr) {
        int smax = jn - i - 6;
        if (smax > kMaxLoop) smax = kMaxLoop;
        int gen = kInf;
        if (smax >= 4) {
          const int upr_i1 = kCst ? (int)c.upr[i + 1] : 1000;
          const bool g = upr_i1 >= 1 && cuj1;
          const bool ok1 = upr_i1 >= 2, ok2 = upl_j1 >= 2;
          const int32_t* xi = col3 + i;
          const int32_t* qrow = C + i + 2;
#define ADP_P1 (qrow + (((jn - 3 - S) * (jn - 4 - S)) >> 1))
#define ADP_C1 (!kCst || (ok1 && upl_j1 >= S - 2))
#define ADP_P2 (xi + S)
#define ADP_C2 (!kCst || (ok2 && upr_i1 >= S - 2))
          if (par) { ADP_MLEVEL_PASS(1, ADP_P1, ADP_C1, ADP_P2, ADP_C2) } else { ADP_MLEVEL_PASS(0, ADP_P1, ADP_C1, ADP_P2, ADP_C2) }
#undef ADP_P1
#undef ADP_C1
#undef ADP_P2
#undef ADP_C2
        } else {
          ADP_MLEVEL_INF()
        }
        (par ? gen1 : gen0)[i] = gen;
      }
      if (mem_state)
        for (int q = 0; q < kFastLevels; q++) w.dstate[(size_t)task * kFastLevels + q] = d[q];
    }
    for (int r = tid; r < nr; r += nt) {
      const int prev = fm1_prev[r + 1];
      fm1_cur[r + 1] = (cuj && prev < kInfHalf) ? prev + mlb : kInf;
      if (kCst ? allowed_type(c, r + 1, jn) != 0 : pair_type(c.S[r + 1], c.S[jn]) != 0) w.cells[ex.atomic_inc(counter)] = (uint16_t)(r + 1);
    }
  };

  ex.sync();
  for (int j = 4; j <= n; j++) {
    const int nr = j - 4;
    const bool cuj = !kCst || can_unpair(c, j);
    const int32_t* fm1_prev = w.vecp((j - 1) & 1);
    int32_t* fm1_cur = w.vecp(j & 1);
    const int m = w.counter[j & 1];
    // ---- interval 1: C(i,j) for the listed cells
    {
      if (tid == 0) w.counter[(j + 1) & 1] = 0;
      const int lg = group_lg(m, nt_lg, warp_lg);
      const int Gs = 1 << lg, slots = nt >> lg, sub = tid & (Gs - 1);
      const int32_t* x1n_rd = w.vecp(4 + (j & 1));     // column j-2
      const int upl_j1 = kCst ? (int)c.upl[j - 1] : 1000;
      const int sj = c.S[j], sj1 = c.S[j - 1], cpj = cp[j];
      for (int slot0 = 0; slot0 < m; slot0 += slots) {
        const int slot = slot0 + (tid >> lg);
        const bool valid = slot < m;
        const int i = valid ? w.cells[slot] : 1;
        int best = kInf;
        int type = 0;
        if (valid) {
          const int si = c.S[i], si1 = c.S[i + 1];
          type = pair_type(si, sj);
          const int codeij = (cn[i] << 4) | cpj;
          const int room = j - i - 6;
          int u1cap = kCst ? (int)c.upr[i + 1] : kMaxLoop, u2cap = upl_j1;
          if (u1cap > kMaxLoop) u1cap = kMaxLoop;
          if (u2cap > kMaxLoop) u2cap = kMaxLoop;
          int other = kInf, accB = kInf, acc1n = kInf, accML = kInf;
          if (sub == 0) {
            other = hairpin_cell_e(P, c, motif_dcal, i, j, type);
            other = imin(other, H.FI[codeij] + imin(gen0[i], gen1[i]));
          }
          if (room >= 0) {
            for (int q = sub; q < 9; q += Gs) {
              int u1, u2;
              corner_u(q, &u1, &u2);
              if (u1 <= u1cap && u2 <= u2cap && u1 + u2 <= room) {
                const int k = i + 1 + u1, l = j - 1 - u2;
                const int v = C[t4(k, l)];
                if (v < kInfHalf)
                  other = imin(other, v + H.RI[(cn[l] << 4) | cp[k]] +
                                          interior_e(P, u1, u2, type, rtype(pair_type(c.S[k], c.S[l])), si1, sj1, c.S[k - 1], c.S[l + 1]));
              }
            }
            {
              const int hi = room < u2cap ? room : u2cap;
              const int cpk = cp[i + 1];
              for (int u2 = 2 + sub; u2 <= hi; u2 += Gs) {
                const int l = j - 1 - u2;
                accB = imin(accB, C[t4(i + 1, l)] + H.RB[(cn[l] << 4) | cpk] + H.tapB[u2]);
              }
            }
            if (u1cap >= 1) {
              int hi = room - 1 < u2cap ? room - 1 : u2cap;
              if (hi > kMaxLoop - 1) hi = kMaxLoop - 1;
              const int cpk = cp[i + 2];
              for (int u2 = 3 + sub; u2 <= hi; u2 += Gs) {
                const int l = j - 1 - u2;
                acc1n = imin(acc1n, C[t4(i + 2, l)] + H.R1[(cn[l] << 4) | cpk] + H.tap1n[u2]);
              }
            }
            {
              const int hi = room < u1cap ? room : u1cap;
              const int32_t* xa = xA + i + 1;
              for (int u1 = 2 + sub; u1 <= hi; u1 += Gs) accB = imin(accB, xa[u1] + H.tapB[u1]);
            }
            if (u2cap >= 1) {
              int hi = room - 1 < u1cap ? room - 1 : u1cap;
              if (hi > kMaxLoop - 1) hi = kMaxLoop - 1;
              const int32_t* xn = x1n_rd + i + 1;
              for (int u1 = 3 + sub; u1 <= hi; u1 += Gs) acc1n = imin(acc1n, xn[u1] + H.tap1n[u1]);
            }
          }
          {
            int k = i + 6 + sub;
            int off = t4(i + 1, k - 1);
            int step = Gs * (k - 5) + ((Gs * (Gs - 1)) >> 1);
            const int step2 = Gs * Gs;
            for (; k <= j - 5; k += Gs) {
              accML = imin(accML, F[off] + fm1_prev[k]);
              off += step;
              step += step2;
            }
          }
          best = imin(imin(other, au_e(H, si, sj) + accB), imin(H.F1n[codeij] + acc1n, accML + H.ML_closing + H.FM[codeij]));
        }
        best = ex.group_min_rt(best, Gs);
        if (valid && sub == 0) {
          if (best >= kInfHalf) best = kInf;
          C[t4(i, j)] = (best < kInf && i > 1 && j < n) ? best + H.FI[(cn[j] << 4) | cp[i]] : best;     // stored premultiplied
          int e = kInf;
          if (best < kInf) {
            fm1_cur[i] = imin(fm1_cur[i], best + mfe_stem_ml(c, H, P, w, type, i, j));
            if (f5[i - 1] < kInfHalf) e = f5[i - 1] + best + stem_e_at(P, c, type, i, j, true);
          }
          ext[slot] = e;
        }
      }
    }
    ex.sync();
    // ---- interval 2: shifted copies of the new column, F(i,j) for every row, exterior prefix, preparation of column j+1
    {
      int32_t* x1n_wr = w.vecp(4 + (j & 1));
      const int cnj = cn[j] << 4;
      for (int r = tid; r < nr; r += nt) {
        const int i = r + 1;
        const int v = C[t4(i, j)];
        const int code = cnj | cp[i];
        xA[i] = v < kInfHalf ? v + H.RB[code] : kInf;
        x1n_wr[i] = v < kInfHalf ? v + H.R1[code] : kInf;
      }
      const int npair = (nr + 1) >> 1;
      const int lg = group_lg(npair, nt_lg, warp_lg);
      const int Gc = 1 << lg, slotsC = nt >> lg, subC = tid & (Gc - 1);
      for (int slot0 = 0; slot0 < npair; slot0 += slotsC) {
        const int slot = slot0 + (tid >> lg);
        const bool valid = slot < npair;
        const int iA = slot + 1, iB = nr - slot;
        int tA = kInf, tB = kInf;
        if (valid) {
          tA = f_row_terms(F, fm1_cur, mlb, iA, nr, kCst ? (int)c.upr[iA] : 1000, subC, Gc);
          if (iB > iA) tB = f_row_terms(F, fm1_cur, mlb, iB, nr, kCst ? (int)c.upr[iB] : 1000, subC, Gc);
        }
        tA = ex.group_min_rt(tA, Gc);
        tB = ex.group_min_rt(tB, Gc);
        if (valid && subC == 0) {
          int fa = imin(tA, fm1_cur[iA]);
          F[t4(iA, j)] = fa >= kInfHalf ? kInf : fa;
          if (iB > iA) {
            int fb = imin(tB, fm1_cur[iB]);
            F[t4(iB, j)] = fb >= kInfHalf ? kInf : fb;
          }
        }
      }
      if (ex.warp() == 0) {
        int e = kInf;
        for (int s = ex.lane(); s < m; s += E::kWarp) e = imin(e, ext[s]);
        e = ex.warp_min(e);
        if (ex.lane() == 0) {
          const int unp = (cuj && f5[j - 1] < kInfHalf) ? f5[j - 1] : kInf;
          const int v = imin(unp, e);
          f5[j] = v >= kInfHalf ? kInf : v;
        }
      }
      if (j < n) prepare_column(j + 1);
    }
    ex.sync();
  }
  return f5[n];
}

}  // namespace adp
