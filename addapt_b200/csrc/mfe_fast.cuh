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
  // what hairpins, stems at the sequence ends, exterior stems and the corner candidates read (see PfHot): only int11 / int21 /
  // int22 and the special-hairpin energies stay in global memory
  int32_t stack[64];                  // [type][type2]
  int32_t mm23[96], mmH[96], mmExt[96];       // compacted mismatch tables, see mm_index (pf_fast.cuh)
  int32_t dangle5[40], dangle3[40];
  int32_t cfe[16];                    // per corner candidate q: bulge[1] | interior[5] + ninio[1] | 0
  int32_t zero, ML_intern, pad0, pad1;
};

struct FastConstI {
  int32_t ip[32];                     // [parity][q]: interior[s] of level s = 4 + parity + 2q (kInf beyond kMaxLoop)
  int32_t nin[32];                    // [parity][q]: ninio(s - 4)
  const int32_t* corner;              // code-indexed corner table (build_corner_table_i): [q-1][closing code][inner code] energies
};

constexpr size_t kCornerEntriesI = (size_t)9 << 16;      // all nine candidates (the MFE loop keeps the stack in the loop)
inline void build_corner_table_i(const IntTables* T, int32_t* out) {
  for (int q = 0; q < 9; q++) {
    const int u1 = (int)((0x322211100ull >> (4 * q)) & 15), u2 = (int)((0x232121010ull >> (4 * q)) & 15);
    for (int co = 0; co < 256; co++) {
      const int a = co >> 4, b = co & 15;
      const int si = (a >> 2) + 1, si1 = (a & 3) + 1, sj = (b >> 2) + 1, sj1 = (b & 3) + 1;
      const int type = pair_type(si, sj);
      for (int ci = 0; ci < 256; ci++) {
        const int a2 = ci >> 4, b2 = ci & 15;      // cn[l], cp[k]
        const int sl = (a2 >> 2) + 1, sl1 = (a2 & 3) + 1, sk = (b2 >> 2) + 1, sk1 = (b2 & 3) + 1;
        const int t2 = pair_type(sk, sl);
        out[((size_t)q << 16) | ((size_t)co << 8) | ci] = (type && t2) ? interior_e(*T, u1, u2, type, rtype(t2), si1, sj1, sk1, sl1) : 10000000;
      }
    }
  }
}

ADP_HD void fill_fast_const_i(FastConstI* fc, const IntTables* T) {
  for (int t = 0; t < 32; t++) {
    const int s = 4 + (t >> 4) + 2 * (t & 15);
    fc->ip[t] = s <= kMaxLoop ? T->interior[s] : kInf;
    fc->nin[t] = s <= kMaxLoop ? T->ninio[s - 4] : kInf;
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
  for (int q = ex.tid(); q < 96; q += ex.nthreads()) {
    const int t = q / 16 + 1, a = ((q >> 2) & 3) + 1, b = (q & 3) + 1;
    H->mm23[q] = T->mm23I[t][a][b];
    H->mmH[q] = T->mmH[t][a][b];
    H->mmExt[q] = T->mmExt[t][a][b];
    if (q < 64) H->stack[q] = (&T->stack[0][0])[q];
    if (q < 40) { H->dangle5[q] = (&T->dangle5[0][0])[q]; H->dangle3[q] = (&T->dangle3[0][0])[q]; }
    if (q < 16) H->cfe[q] = (q == 1 || q == 2) ? T->bulge[1] : (q == 7 || q == 8) ? T->interior[5] + T->ninio[1] : 0;
  }
  if (ex.tid() == 0) {
    H->ML_base = T->ML_base; H->ML_closing = T->ML_closing; H->TermAU = T->TerminalAU; H->inf = kInf;
    H->zero = 0; H->ML_intern = T->ML_intern; H->pad0 = 0; H->pad1 = 0;
  }
#if !defined(__CUDA_ARCH__)
  fill_fast_const_i(&h_fast_const_i, T);
  {
    static const IntTables* built_for = nullptr;
    static int32_t* table = nullptr;
    if (built_for != T) {
      if (!table) table = new int32_t[kCornerEntriesI];
      build_corner_table_i(T, table);
      built_for = T;
    }
    h_fast_const_i.corner = table;
  }
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
  int32_t* hpw;    // [ld]: hairpin[u] by loop size
  int16_t* spx;    // [3][ld]: index of the special hairpin (tri / tetra / hexa loop) closed at i, or -1
  int* lvoff;      // [32]: byte offsets of the row candidates of the level recursion, [parity][level]
  int32_t* dstate; // see FastWork::dstate
  int ld, R;
};

ADP_HD size_t mfe_fast_small_bytes(int n) {
  const size_t ld = (size_t)fast_ld(n);
  return sizeof(MfeHot) + 64 + ctx_bytes(n) + kMfeFastVectors * ld * sizeof(int32_t) + 2 * ld + 2 * ld + 64 + 16 * 12
         + (size_t)(4 * n + 16) * 3 * sizeof(int16_t) + (n + 16) + 64       // + traceback stack and output string
         + ld * sizeof(int32_t) + 3 * ld * sizeof(int16_t) + 32 * sizeof(int) + 48;
}
ADP_HD size_t mfe_fast_matrix_ints(int n) { return 2 * t4_count(n) + 4; }

ADP_HD int imin(int a, int b) { return a < b ? a : b; }
ADP_HD int au_e(const MfeHot& H, int a, int b) { return a * b == 6 ? 0 : H.TermAU; }

// stems as stem_e (fold_common.cuh), from the staged tables
ADP_HD int mfe_stem_ml(const SeqCtx& c, const MfeHot& H, const MfeFastWork& w, int type, int i, int j) {
  if (i > 1 && j < c.n) return H.FM[(w.cn[j] << 4) | w.cp[i]];
  int e = 0;
  if (i > 1) e = H.dangle5[type * 5 + c.S[i - 1]];
  else if (j < c.n) e = H.dangle3[type * 5 + c.S[j + 1]];
  if (type > 2) e += H.TermAU;
  return e + H.ML_intern;
}
ADP_HD int mfe_stem_ext(const SeqCtx& c, const MfeHot& H, int type, int i, int j) {
  int e = 0;
  if (i > 1 && j < c.n) e = H.mmExt[mm_index(type, c.S[i - 1], c.S[j + 1])];
  else if (i > 1) e = H.dangle5[type * 5 + c.S[i - 1]];
  else if (j < c.n) e = H.dangle3[type * 5 + c.S[j + 1]];
  if (type > 2) e += H.TermAU;
  return e;
}

// special hairpins of the job and the hairpin energies by size, found / copied once per job
template <class E>
ADP_HD void mfe_special_hairpins(const E& ex, const SeqCtx& c, const IntTables* P, MfeFastWork& w) {
  const int n = c.n, ld = w.ld;
  for (int t = ex.tid(); t < 3 * ld; t += ex.nthreads()) {
    const int cls = t / ld, i = t - cls * ld;
    const int u = cls == 0 ? 3 : cls == 1 ? 4 : 6;
    int idx = -1;
    uint32_t code;
    if (i >= 1 && i + u + 1 <= n && pack_window(c.S, i, u + 2, &code)) {
      const uint32_t* codes = u == 4 ? P->tetra_code : u == 6 ? P->hex_code : P->tri_code;
      const int cnt = u == 4 ? P->n_tetra : u == 6 ? P->n_hex : P->n_tri;
      for (int k = 0; k < cnt; k++)
        if (codes[k] == code) { idx = k; break; }
    }
    w.spx[t] = (int16_t)idx;
  }
  for (int u = ex.tid(); u < ld; u += ex.nthreads()) w.hpw[u] = P->hairpin[u];
}

// hairpin_cell_e (mfe_core.cuh) from the staged tables
ADP_HD int mfe_fast_hairpin(const SeqCtx& c, const MfeHot& H, const IntTables& P, const MfeFastWork& w, int motif_dcal, int i, int j, int type) {
  const int u = j - i - 1;
  if ((int)c.upr[i + 1] < u) return kInf;
  int sx = -1;
  if (u == 3) sx = w.spx[i];
  else if (u == 4) sx = w.spx[w.ld + i];
  else if (u == 6) sx = w.spx[2 * w.ld + i];
  int e;
  if (sx >= 0) e = (u == 4 ? P.tetra_e : u == 6 ? P.hex_e : P.tri_e)[sx];
  else if (u == 3) e = w.hpw[3] + (type > 2 ? H.TermAU : 0);
  else e = w.hpw[u] + H.mmH[mm_index(type, c.S[i + 1], c.S[j - 1])];
  if (is_motif_site(c, i, j)) e += motif_dcal;
  return e;
}

// Corner candidate q (see corner_entry in pf_fast.cuh): the energy is  *add + (*returned address); same cases as interior_e.
ADP_HD const int32_t* mfe_corner_entry(const MfeHot& H, const IntTables& P, int q, int type, int type2, int si1, int sj1, int sp1, int sq1, int* add) {
  int a = H.cfe[q];
  const int32_t* p;
  if (q <= 2) p = &H.stack[type * 8 + type2];
  else if (q == 3) p = &P.int11[type][type2][si1][sj1];
  else if (q == 4) p = &P.int21[type][type2][si1][sq1][sj1];
  else if (q == 5) p = &P.int21[type2][type][sq1][si1][sp1];
  else if (q == 6) p = &P.int22[type][type2][si1][sp1][sq1][sj1];
  else {
    a += H.mm23[mm_index(type, si1, sj1)] + H.mm23[mm_index(type2, sq1, sp1)];
    p = &H.zero;
  }
  *add = a;
  return p;
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

// One level of the (min,+) recursion; see ADP_LEVEL_STEP in pf_fast.cuh (run-time level parity, one code path).  `zp` points at
// an always-kInf word.
#define ADP_MLEVEL_STEP(Q, ADDR1, C1, ADDR2, C2)                                       \
  {                                                                                   \
    const int32_t* p1 = (4 + 2 * (Q) <= sm1 && C1(Q)) ? ADDR1(Q) : zp;                \
    const int32_t* p2 = (4 + 2 * (Q) <= sm2 && ((Q) > 0 || par != 0) && C2(Q)) ? ADDR2(Q) : zp; \
    int v = imin(fcn[Q] + imin(*p1, *p2), kInf);                                      \
    if ((Q) > 0) v = imin(v, (!kCst || (4 + 2 * (Q) <= sm && g)) ? d[(Q) > 0 ? (Q) - 1 : 0] : kInf); \
    if ((Q) > 0) gen = imin(gen, fci[Q] + v);                                         \
    d[Q] = v;                                                                         \
  }
#define ADP_MLEVEL_PASS(A1, C1, A2, C2)                                                \
  ADP_MLEVEL_STEP(13, A1, C1, A2, C2) ADP_MLEVEL_STEP(12, A1, C1, A2, C2) ADP_MLEVEL_STEP(11, A1, C1, A2, C2) \
  ADP_MLEVEL_STEP(10, A1, C1, A2, C2) ADP_MLEVEL_STEP(9, A1, C1, A2, C2) ADP_MLEVEL_STEP(8, A1, C1, A2, C2)  \
  ADP_MLEVEL_STEP(7, A1, C1, A2, C2) ADP_MLEVEL_STEP(6, A1, C1, A2, C2) ADP_MLEVEL_STEP(5, A1, C1, A2, C2)  \
  ADP_MLEVEL_STEP(4, A1, C1, A2, C2) ADP_MLEVEL_STEP(3, A1, C1, A2, C2) ADP_MLEVEL_STEP(2, A1, C1, A2, C2)  \
  ADP_MLEVEL_STEP(1, A1, C1, A2, C2) ADP_MLEVEL_STEP(0, A1, C1, A2, C2)
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
  for (int t = tid; t < 32; t += nt) w.lvoff[t] = 0;
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
    const int kbase = ((jn + 1) / R) * R;
    for (int task = tid; task < 2 * half; task += nt) {
      const int par = task >= half ? 1 : 0;
      const int slot = task - par * half;
      if (slot >= R) continue;
      if (mem_state)
        for (int q = 0; q < kFastLevels; q++) d[q] = w.dstate[(size_t)task * kFastLevels + q];
      const int i = fast_slot_key(slot, kbase, jn, R) - jn;
      if (i >= 1 && i <= nr) {
        int sm = jn - i - 6;
        if (sm > kMaxLoop) sm = kMaxLoop;
        sm -= par;                                     // level Q is on iff 4 + 2Q <= sm
        int gen = kInf;
        if (sm + par >= 4) {
          const int upr_i1 = kCst ? (int)c.upr[i + 1] : 1000;
          const bool g = upr_i1 >= 1 && cuj1;
          const bool ok1 = upr_i1 >= 2, ok2 = upl_j1 >= 2;
          const int sm1 = sm, sm2 = sm;
          const int lim1 = upl_j1 - par - 2, lim2 = upr_i1 - par - 2;
          const int32_t* xi = col3 + i + par + 4;      // C+FI(i+S-1, jn-3) = xi[2Q]
          const char* qrow = reinterpret_cast<const char*>(C + i + 2);    // C+FI(i+3, l), l = jn+1-S: offset from the table
          const int32_t* fcn = ADP_FCI.nin + (par << 4);
          const int32_t* fci = ADP_FCI.ip + (par << 4);
          ADP_LEVEL_OFFSETS(w.lvoff + (par << 4))
#define ADP_A1(Q) reinterpret_cast<const int32_t*>(qrow + lo[Q])
#define ADP_C1(Q) (!kCst || (ok1 && 2 * (Q) <= lim1))
#define ADP_A2(Q) (xi + 2 * (Q))
#define ADP_C2(Q) (!kCst || (ok2 && 2 * (Q) <= lim2))
          ADP_MLEVEL_PASS(ADP_A1, ADP_C1, ADP_A2, ADP_C2)
#undef ADP_A1
#undef ADP_C1
#undef ADP_A2
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
      ADP_NOUNROLL
      for (int t = tid; t < 32; t += nt) {
        // byte offsets of the row candidates of the level recursion of column j+1: level S reads C+FI(i+3, j+2-S)
        const int S = 4 + (t >> 4) + 2 * (t & 15);
        const int l = j + 2 - S;
        w.lvoff[t] = (((l - 4) * (l - 5)) >> 1) * (int)sizeof(int32_t);
      }
      const int lg = group_lg(m, nt_lg, warp_lg);
      const int Gs = 1 << lg, slots = nt >> lg, sub = tid & (Gs - 1);
      const int32_t* x1n_rd = w.vecp(4 + (j & 1));     // column j-2
      const int upl_j1 = kCst ? (int)c.upl[j - 1] : 1000;
      const int sj = c.S[j], cpj = cp[j];
      for (int slot0 = 0; slot0 < m; slot0 += slots) {
        const int slot = slot0 + (tid >> lg);
        const bool valid = slot < m;
        const int i = valid ? w.cells[slot] : 1;
        int best = kInf;
        int type = 0;
        if (valid) {
          const int si = c.S[i];
          type = pair_type(si, sj);
          const int codeij = (cn[i] << 4) | cpj;
          const int room = j - i - 6;
          int u1cap = kCst ? (int)c.upr[i + 1] : kMaxLoop, u2cap = upl_j1;
          if (u1cap > kMaxLoop) u1cap = kMaxLoop;
          if (u2cap > kMaxLoop) u2cap = kMaxLoop;
          int other = kInf, accB = kInf, acc1n = kInf, accML = kInf;
          int cadd = kInf, ctab = 0;
          if (sub == 0) {
            other = mfe_fast_hairpin(c, H, P, w, motif_dcal, i, j, type);
            other = imin(other, H.FI[codeij] + imin(gen0[i], gen1[i]));
          }
          if (room >= 0) {
            // corner candidates, software-pipelined and rolled like in pf_fast.cuh: the table load of a round is consumed in the next
            ADP_NOUNROLL
            for (int q = sub; q < 9; q += Gs) {
              int u1, u2;
              corner_u(q, &u1, &u2);
              int a2 = kInf;
              const int32_t* p2 = &H.zero;
              if (u1 <= u1cap && u2 <= u2cap && u1 + u2 <= room) {
                const int k = i + 1 + u1, l = j - 1 - u2;
                const int v = C[t4(k, l)];
                if (v < kInfHalf) {
                  const int cin = (cn[l] << 4) | cp[k];
                  p2 = ADP_FCI.corner + ((size_t)q << 16) + ((size_t)codeij << 8) + cin;
                  a2 = v + H.RI[cin];
                }
              }
              const int t2 = *p2;
              other = imin(other, cadd + ctab);
              cadd = a2;
              ctab = t2;
            }
            // the four 1-D filters along the loop's axes fused into one loop (see pf_fast.cuh)
            {
              const int hiB3 = room < u2cap ? room : u2cap;
              const int hiB5 = room < u1cap ? room : u1cap;
              int hi13 = room - 1 < u2cap ? room - 1 : u2cap, hi31 = room - 1 < u1cap ? room - 1 : u1cap;
              if (hi13 > kMaxLoop - 1) hi13 = kMaxLoop - 1;
              if (hi31 > kMaxLoop - 1) hi31 = kMaxLoop - 1;
              if (u1cap < 1) hi13 = -1;
              if (u2cap < 1) hi31 = -1;
              const int hmax = hiB3 > hiB5 ? hiB3 : hiB5;
              const int cpk1 = cp[i + 1], cpk2 = cp[i + 2];
              const int32_t* xa = xA + i + 1;
              const int32_t* xn = x1n_rd + i + 1;
              for (int u = 2 + sub; u <= hmax; u += Gs) {
                const int l = j - 1 - u;
                const int cl = cn[l] << 4;
                const int32_t* pb = C + t4(i + 1, l);
                const int b3 = pb[0] + H.RB[cl | cpk1], n3 = pb[1] + H.R1[cl | cpk2];
                const int b5 = xa[u], n5 = xn[u];
                accB = imin(accB, H.tapB[u] + imin(u <= hiB3 ? b3 : kInf, u <= hiB5 ? b5 : kInf));
                acc1n = imin(acc1n, H.tap1n[u] + imin(u <= hi13 ? n3 : kInf, u <= hi31 ? n5 : kInf));
              }
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
          other = imin(other, cadd + ctab);
          best = imin(imin(other, au_e(H, si, sj) + accB), imin(H.F1n[codeij] + acc1n, accML + H.ML_closing + H.FM[codeij]));
        }
        best = ex.group_min_rt(best, Gs);
        if (valid && sub == 0) {
          if (best >= kInfHalf) best = kInf;
          C[t4(i, j)] = (best < kInf && i > 1 && j < n) ? best + H.FI[(cn[j] << 4) | cp[i]] : best;     // stored premultiplied
          int e = kInf;
          if (best < kInf) {
            fm1_cur[i] = imin(fm1_cur[i], best + mfe_stem_ml(c, H, w, type, i, j));
            if (f5[i - 1] < kInfHalf) e = f5[i - 1] + best + mfe_stem_ext(c, H, type, i, j);
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
