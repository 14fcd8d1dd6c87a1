// Fast McCaskill inside / outside pass, one construct per CTA (ACGU/T-only sequences, n <= kFastMaxN).
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
//    enclosing pairs, (k,l) -> (k-1,l+1).  Level parity is a run-time, warp-uniform value: both parities share ONE code
//    path (code size, see below), the size-dependent factors come from [parity][level] tables in the constant bank and the
//    row offsets of a column from a 32-entry table in shared memory.
//    Levels that are off for a thread (no room, or closed by the hard constraint) read an always-zero word instead, so
//    the recursion is branch-free and every lane of a warp executes the same instructions.
//  * Qb (and, in the outside pass, the outside weight) is stored PREMULTIPLIED by the generic mismatch factor of its
//    pair, so the two border candidates of every level are single loads; the other readers multiply by a ratio from a
//    16x16 table (1/FI, terminal-AU/FI, F1n/FI, FM/FI).
//  * The candidates on the loop's axes (bulges, 1xn loops: 112 per cell) are four 1-D filters; the two that run
//    down a column read a per-column vector premultiplied once, the two that run along a row read the matrix; all four
//    are evaluated in ONE loop per lane so that their load chains overlap.  The 9 corner candidates (stack, bulge-1,
//    1x1, 1x2, 2x2, 2x3) use the full tables: the stack as straight-line code, the other eight in a rolled loop whose
//    table load is issued one round ahead of its use.
//  * All per-pair factors come from 16x16 tables indexed by two one-byte position codes
//    (nucleotide + 3' neighbour, nucleotide + 5' neighbour), staged once per job -- together with every other small
//    table a cell reads (stack, hairpin / exterior / 2x3 mismatches, dangles, hairpin weights by size, the special
//    hairpin closed at each position): only int11 / int21 / int22 are read from global memory inside the sweep.
//  * Only cells that can pair are evaluated: each column compacts them into a list, and a power-of-two lane group
//    per cell (sized so the list fills the CTA) splits the cell's terms and reduces with shuffles.
//  * Two barriers per column: the sparse per-cell phase alone, then everything that is independent of it (the level
//    recursion / carriers / compaction / per-row hairpin and stem factors of the NEXT column, and Qm of the PREVIOUS
//    column from a queue of warp-sized chunks that evens out the warps) in one interval.
//  * The code is kept small on purpose: the SM's instruction caches hold 6 KB (L0) and 32 KB (L1.5), the two resident
//    CTAs are usually in different passes, and the kernel's speed follows its SASS size.
#pragma once
#include "fold_common.cuh"

#if defined(__CUDA_ARCH__)
#define ADP_UNROLL _Pragma("unroll")
#define ADP_NOUNROLL _Pragma("unroll 1")
#else
#define ADP_UNROLL
#define ADP_NOUNROLL
#endif

namespace adp {

// Phase timing for tuning (tools/variant_build.sh prof -DADP_PHASE_PROFILE, then tools/quick_time.py --phase): every warp of CTA 0 accumulates clock64()
// deltas per phase of the column loop and prints them for the first jobs.  Off in the product build.
#if defined(ADP_PHASE_PROFILE) && defined(__CUDACC__)
static __device__ int g_prof_reports;
#endif
#if defined(ADP_PHASE_PROFILE) && defined(__CUDA_ARCH__)
#define ADP_PROF_DECL long long prof_[14] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0}; long long tprev_ = clock64();
#define ADP_TICK(s_) { const long long t_ = clock64(); prof_[s_] += t_ - tprev_; tprev_ = t_; }
#define ADP_PROF_REPORT(tag_)                                                                                                   \
  if (blockIdx.x == 0 && (threadIdx.x & 31) == 0 && atomicAdd(&g_prof_reports, 1) < 48)                                          \
    printf(tag_ " warp %d: cells %lld drain1 %lld wait1 %lld vec %lld levels %lld rows %lld drain2 %lld wait2 %lld | head %lld corners %lld axis %lld ml %lld reduce %lld store %lld\n", (int)(threadIdx.x >> 5), \
           prof_[0], prof_[1], prof_[2], prof_[3], prof_[4], prof_[5], prof_[6], prof_[7], prof_[8], prof_[9], prof_[10], prof_[11], prof_[12], prof_[13]);
#else
#define ADP_PROF_DECL
#define ADP_TICK(s_)
#define ADP_PROF_REPORT(tag_)
#endif

constexpr int kFastLevels = 14;       // level sums per thread: s = 4+par, 6+par, ..., 30+par
constexpr int kFastMaxN = 320;        // longest construct the fast path takes (ADP_MAX_LEN)

struct PfHot {                        // shared memory, staged once per CTA
  double FI[256];                     // mismatch_interior (closing orientation); mismatch_interior_1n = R1 * FI; multiloop stem factor = RM * FI
  // The matrices hold Qb (inside) / the outside weight PREMULTIPLIED by FI of the pair (inner orientation inside, outer
  // orientation outside), so that the level recursion reads its candidates with one load.  Ratios for the other readers:
  double RI[256], RB[256], R1[256], RM[256];   // 1/FI, terminal-AU factor / FI, F1n / FI, FM / FI  (0 for non-pairs)
  double tapB[32];                    // bulge[u] * scale[u+2]                      (u >= 2)
  double tap1n[32];                   // interior[u+1] * ninio[u-1] * scale[u+3]    (1 x u loop, u >= 3)
  double scale1, scale2, eMLb1, MLclosing, TermAU, pad;
  // what hairpins, stems at the sequence ends, exterior stems and the corner candidates read: staged so that no L2-latency
  // load sits on the path of a cell (only int11 / int21 / int22 and the special-hairpin weights stay in global memory)
  double stack[64];                   // [type][type2]
  double mm23[96], mmH[96], mmExt[96];        // [type 1..6][5' side neighbour 1..4][3' side neighbour 1..4], see mm_index
  double dangle5[40], dangle3[40];
  double cfac[16];                    // per corner candidate q: scale[u1+u2+2] * (bulge[1] | interior[5] * ninio[1] | 1)
  double one, MLintern;
  uint8_t tyr[256];                   // by position codes (nucleotide + 3' neighbour, nucleotide + 5' neighbour): pair type of the
                                      // two NEIGHBOURS, 3' one first -- the reversed type of the stacked inner pair (inside), the type
                                      // of the stacked enclosing pair (outside)
};

// size-indexed factors of the generic interior loops; read with compile-time indices
struct FastConst {
  double ip[32];                      // [parity][q]: interior[s] * scale[s+2] of level s = 4 + parity + 2q (0 beyond kMaxLoop)
  double nin[32];                     // [parity][q]: ninio factor of asymmetry s - 4
  const double* corner;               // see build_corner_table
};

// Code-indexed corner table (round 2: +11 % fold-evals/s, +10 % MFE folds/s over per-candidate decoding, profiles/r2_experiments.md).
// The weight of corner candidate q = 1..8 depends on eight nucleotides only, and the cell already holds them as two bytes: the closing pair's code (nucleotide + 3' neighbour of i, nucleotide +
// 5' neighbour of j) and the inner pair's code (the same of l and k), the byte it looks RI up with.  One table of
// 8 x 256 x 256 doubles (4 MB, L2-resident), indexed [q-1][closing code][inner code] and holding
// scale[u1+u2+2] * interior_x(...), replaces the per-candidate type / neighbour decoding and the six-way case split of
// corner_entry: every lane of a group runs the same short path.
constexpr size_t kCornerEntries = (size_t)8 << 16;
inline void build_corner_table(const PfTables* T, double* out) {
  for (int q = 1; q < 9; q++) {
    const int u1 = (int)((0x322211100ull >> (4 * q)) & 15), u2 = (int)((0x232121010ull >> (4 * q)) & 15);
    for (int co = 0; co < 256; co++) {
      const int a = co >> 4, b = co & 15;
      const int si = (a >> 2) + 1, si1 = (a & 3) + 1, sj = (b >> 2) + 1, sj1 = (b & 3) + 1;
      const int type = pair_type(si, sj);
      for (int ci = 0; ci < 256; ci++) {
        const int a2 = ci >> 4, b2 = ci & 15;      // cn[l], cp[k]
        const int sl = (a2 >> 2) + 1, sl1 = (a2 & 3) + 1, sk = (b2 >> 2) + 1, sk1 = (b2 & 3) + 1;
        const int t2 = pair_type(sk, sl);
        out[((size_t)(q - 1) << 16) | ((size_t)co << 8) | ci] =
            (type && t2) ? T->scale[u1 + u2 + 2] * interior_x(*T, T, u1, u2, type, rtype(t2), si1, sj1, sk1, sl1) : 0.;
      }
    }
  }
}

ADP_HD void fill_fast_const(FastConst* fc, const PfTables* T) {
  for (int t = 0; t < 32; t++) {
    const int s = 4 + (t >> 4) + 2 * (t & 15);
    fc->ip[t] = s <= kMaxLoop ? T->interior[s] * T->scale[s + 2] : 0.;
    fc->nin[t] = s <= kMaxLoop ? T->ninio[s - 4] : 0.;
  }
}

#if defined(__CUDACC__)
static __constant__ FastConst c_fast_const;       // one copy per translation unit; uploaded before every launch
#endif
static FastConst h_fast_const;                    // host emulation
#if defined(__CUDA_ARCH__)
#define ADP_FC c_fast_const
#else
#define ADP_FC h_fast_const
#endif

// the 9 interior-loop candidates that go through the full tables: (u1,u2) =
// (0,0)(0,1) (1,0)(1,1)(1,2) (2,1)(2,2)(2,3) (3,2), packed 4 bits each
ADP_HD void corner_u(int q, int* u1, int* u2) {
  *u1 = (int)((0x322211100ull >> (4 * q)) & 15);
  *u2 = (int)((0x232121010ull >> (4 * q)) & 15);
}

// index of (pair type 1..6, nucleotide 1..4, nucleotide 1..4) in the compacted mismatch tables of PfHot
ADP_HD int mm_index(int type, int a, int b) { return type * 16 + a * 4 + b - 21; }

template <class E>
ADP_HD void stage_hot(const E& ex, PfHot* H, const PfTables* T) {
  for (int q = ex.tid(); q < 256; q += ex.nthreads()) {
    const int a = q >> 4, b = q & 15;
    const int x = (a >> 2) + 1, u = (a & 3) + 1, y = (b >> 2) + 1, v = (b & 3) + 1;
    const int t = pair_type(x, y);
    H->FI[q] = t ? T->mmI[t][u][v] : 0.;
    const double fm = t ? T->mmM[rtype(t)][v][u] * (t > 2 ? T->TermAU : 1.) * T->MLintern : 0.;   // multiloop stem factor
    const double inv = t ? 1. / T->mmI[t][u][v] : 0.;
    H->RI[q] = inv;
    H->RB[q] = t ? (t > 2 ? T->TermAU : 1.) * inv : 0.;
    H->R1[q] = t ? T->mm1nI[t][u][v] * inv : 0.;
    H->RM[q] = t ? fm * inv : 0.;
    H->tyr[q] = (uint8_t)pair_type(v, u);
  }
  for (int q = ex.tid(); q < 32; q += ex.nthreads()) {
    H->tapB[q] = (q >= 2 && q <= 30) ? T->bulge[q] * T->scale[q + 2] : 0.;
    H->tap1n[q] = (q >= 3 && q <= 29) ? T->interior[q + 1] * T->ninio[q - 1] * T->scale[q + 3] : 0.;
  }
  for (int q = ex.tid(); q < 200; q += ex.nthreads()) {
    if (q < 96) {
      const int t = q / 16 + 1, a = ((q >> 2) & 3) + 1, b = (q & 3) + 1;
      H->mm23[q] = T->mm23I[t][a][b];
      H->mmH[q] = T->mmH[t][a][b];
      H->mmExt[q] = T->mmExt[t][a][b];
    }
    if (q < 64) H->stack[q] = (&T->stack[0][0])[q];
    if (q < 40) { H->dangle5[q] = (&T->dangle5[0][0])[q]; H->dangle3[q] = (&T->dangle3[0][0])[q]; }
    if (q < 16) {
      int u1 = 0, u2 = 0;
      if (q < 9) corner_u(q, &u1, &u2);
      double f = q < 9 ? T->scale[u1 + u2 + 2] : 0.;
      if (q == 1 || q == 2) f *= T->bulge[1];
      if (q == 7 || q == 8) f *= T->interior[5] * T->ninio[1];
      H->cfac[q] = f;
    }
  }
  if (ex.tid() == 0) {
    H->scale1 = T->scale[1]; H->scale2 = T->scale[2]; H->eMLb1 = T->eMLb[1];
    H->MLclosing = T->MLclosing; H->TermAU = T->TermAU; H->pad = 0.;
    H->one = 1.; H->MLintern = T->MLintern;
  }
#if !defined(__CUDA_ARCH__)
  fill_fast_const(&h_fast_const, T);
  {
    static const PfTables* built_for = nullptr;
    static double* table = nullptr;
    if (built_for != T) {
      if (!table) table = new double[kCornerEntries];
      build_corner_table(T, table);
      built_for = T;
    }
    h_fast_const.corner = table;
  }
#endif
}

// The unpaired-prefix half of Qm(i,j), U(i) = sum_u eMLb^u Qm1(i+u,j), obeys U(i) = [i may stay unpaired] eMLb
// (Qm1(i+1,j) + U(i+1)): one warp evaluates it for the whole column as a scan (two vectors, by column parity), and the
// dot products of the chunk queue keep only the other half of their terms.
constexpr int kFastVectors = 20;      // see FastVec (inside pass) / FastVecOut (outside pass)

struct FastWork {
  double* qb;      // Qb, then (outside) the outside weights: trimmed triangle
  double* qm;      // Qm: trimmed triangle
  double* vbase;   // kFastVectors vectors of ld doubles each, contiguous; roles are named where they are used
  ADP_HD double* vecp(int v) const { return vbase + (size_t)v * ld; }
  const double* scaleG;        // pf_scale^-k (global)
  uint8_t* cn;     // [ld] code of (S[p], S[p+1])
  uint8_t* cp;     // [ld] code of (S[p], S[p-1])
  uint16_t* cells; // [2][ld] rows of a column that can pair, by column parity
  uint8_t* iscell; // [2][ld] whether row i of a column is in its list, by column parity
  int* counter;    // [4]: 0..2 list lengths by column mod 3, 3 "declined" flag
  int16_t* spx;    // [3][ld]: index of the special hairpin (tri / tetra / hexa loop) closed at i, or -1
  int* lvoff;      // [2][32]: byte offsets of the row candidates of the level recursion, [column parity][level parity][level]
  int* qctr;       // [3]: chunk counters of the deferred work queue, by column mod 3
  double* dstate;  // NULL: every thread keeps the level sums of its anti-diagonal in registers (needs 2*R <= threads);
                   // else [2*R*kFastLevels] doubles of memory holding them (host emulation; constructs longer than threads/2)
  int ld;          // vector length (>= n + 4)
  int R;           // anti-diagonal slots (fast_R); the CTA needs 2*R threads
};

// trimmed column-major triangle: cell (i,j), 1 <= i <= j-4
ADP_HD int t4(int i, int j) { return (((j - 4) * (j - 5)) >> 1) + (i - 1); }
ADP_HD size_t t4_count(int n) { return n >= 5 ? (size_t)(n - 4) * (n - 3) / 2 : 0; }
ADP_HD int fast_ld(int n) { return (n + 6) & ~1; }
// Anti-diagonal slots: at column j the live anti-diagonals are the keys i+j of the rows i = 1 .. j-4, a window of at most n-4
// consecutive keys, so any R >= n-4 slots (key mod R) never collide.  R = n-4 packs the level tasks of a 100-nt construct into
// exactly three warps per parity; measured on one box (round 2): inside fills +2.7 %, MFE folds +3.3 %, but fold-evals -0.3 %,
// scores -3 % and Monte-Carlo steps -5.5 % -- so the PF kernels keep R = n and the MFE kernel (mfe_fast_R) takes n-4.
ADP_HD int fast_R(int n) { return n < 4 ? 4 : n; }
ADP_HD int mfe_fast_R(int n) { return n < 8 ? 4 : n - 4; }
// the live key of `slot` at column j: the only key in [j+1, j+R] congruent to the slot; kbase = ((j+1)/R)*R
ADP_HD int fast_slot_key(int slot, int kbase, int j, int R) { const int key = kbase + slot; return key < j + 1 ? key + R : key; }
ADP_HD int fast_max_len(int threads) { return threads / 2; }      // longest construct with register-resident level sums
ADP_HD size_t fast_state_doubles(int n) { return (size_t)2 * fast_R(n) * kFastLevels; }

ADP_HD size_t fast_small_bytes(int n) {
  const size_t ld = (size_t)fast_ld(n);
  return sizeof(PfHot) + 64 + ctx_bytes(n) + kFastVectors * ld * sizeof(double) + 2 * ld /* cn, cp */ + 2 * 2 * ld /* cells */ + 2 * ld /* iscell */ + 64 +
         16 * 12 + (4 + 64 + 4) * sizeof(int) + 32 + 3 * ld * sizeof(int16_t) /* spx */ + 32;
}
ADP_HD size_t fast_matrix_doubles(int n) { return 2 * t4_count(n) + 2; }

// weak pairs (AU, UA, GU, UG) pay the terminal penalty; C*G is the only nucleotide-code product equal to 6
ADP_HD double au_factor(const PfHot& H, int a, int b) { return a * b == 6 ? 1. : H.TermAU; }

ADP_HD int ilog2_floor(unsigned v) {
#if defined(__CUDA_ARCH__)
  return 31 - __clz((int)v);
#else
  return 31 - __builtin_clz(v);
#endif
}

// log2 of the lanes per cell such that m cells fill (at most) a CTA of 2^nt_lg threads; at most a warp per cell
ADP_HD int group_lg(int m, int nt_lg, int cap_lg) {
  int lg = nt_lg;
  if (m > 1) lg -= ilog2_floor((unsigned)(m - 1)) + 1;      // ceil(log2 m)
  if (lg > cap_lg) lg = cap_lg;
  return lg < 0 ? 0 : lg;
}

// multiloop / exterior stem factor of the pair (i,j) seen from outside (neighbours i-1, j+1 dangle when they exist); same
// products in the same order as stem_x (fold_common.cuh)
ADP_HD double fast_stem_ml(const SeqCtx& c, const PfHot& H, const FastWork& w, int type, int i, int j) {
  if (i > 1 && j < c.n) { const int q = (w.cn[j] << 4) | w.cp[i]; return H.RM[q] * H.FI[q]; }
  double e = 1.;
  if (i > 1) e = H.dangle5[type * 5 + c.S[i - 1]];
  else if (j < c.n) e = H.dangle3[type * 5 + c.S[j + 1]];
  if (type > 2) e *= H.TermAU;
  return e * H.MLintern;
}
ADP_HD double fast_stem_ext(const SeqCtx& c, const PfHot& H, int type, int i, int j) {
  double e = 1.;
  if (i > 1 && j < c.n) e = H.mmExt[mm_index(type, c.S[i - 1], c.S[j + 1])];
  else if (i > 1) e = H.dangle5[type * 5 + c.S[i - 1]];
  else if (j < c.n) e = H.dangle3[type * 5 + c.S[j + 1]];
  if (type > 2) e *= H.TermAU;
  return e;
}

// special hairpins of the job: index into the tri / tetra / hexa loop tables per 5' end, found once per job
template <class E>
ADP_HD void fast_special_hairpins(const E& ex, const SeqCtx& c, const PfTables* G, FastWork& w) {
  const int n = c.n, ld = w.ld;
  for (int t = ex.tid(); t < 3 * ld; t += ex.nthreads()) {
    const int cls = t / ld, i = t - cls * ld;
    const int u = cls == 0 ? 3 : cls == 1 ? 4 : 6;
    int idx = -1;
    uint32_t code;
    if (i >= 1 && i + u + 1 <= n && pack_window(c.S, i, u + 2, &code)) {
      const uint32_t* codes = u == 4 ? G->tetra_code : u == 6 ? G->hex_code : G->tri_code;
      const int cnt = u == 4 ? G->n_tetra : u == 6 ? G->n_hex : G->n_tri;
      for (int k = 0; k < cnt; k++)
        if (codes[k] == code) { idx = k; break; }
    }
    w.spx[t] = (int16_t)idx;
  }
}

template <bool kCst>
ADP_HD double fast_hairpin(const SeqCtx& c, const PfHot& H, const PfTables* G, const FastWork& w, double motif_boltz, int i, int j, int type) {
  const int u = j - i - 1;
  if (kCst && (int)c.upr[i + 1] < u) return 0.;
  int sx = -1;
  if (u == 3) sx = w.spx[i];
  else if (u == 4) sx = w.spx[w.ld + i];
  else if (u == 6) sx = w.spx[2 * w.ld + i];
  double q;
  if (sx >= 0) q = (u == 4 ? G->tetra_x : u == 6 ? G->hex_x : G->tri_x)[sx] * w.scaleG[u + 2];
  else if (u == 3) q = type > 2 ? (G->hairpin[3] * w.scaleG[5]) * H.TermAU : G->hairpin[3] * w.scaleG[5];
  else q = (G->hairpin[u] * w.scaleG[u + 2]) * H.mmH[mm_index(type, c.S[i + 1], c.S[j - 1])];      // hairpin weight by loop size: two global loads, off the critical path
  if (is_motif_site(c, i, j)) q *= motif_boltz;   // the bound motif REPLACES the empty loop
  return q;
}

// Corner candidate q of the closing pair (type; neighbours si1, sj1 inside) with the inner pair (type2 already reversed;
// neighbours sp1 = S[k-1], sq1 = S[l+1]): the weight is  *fac * (*returned address)  -- same cases as interior_x
// (fold_common.cuh), split so that the one load that may go to global memory (int11 / int21 / int22) can be issued early.
ADP_HD const double* corner_entry(const PfHot& H, const PfTables* G, int q, int type, int type2, int si1, int sj1, int sp1, int sq1, double* fac) {
  double f = H.cfac[q];
  const double* p;
  if (q <= 2) p = &H.stack[type * 8 + type2];
  else if (q == 3) p = &G->int11[type][type2][si1][sj1];
  else if (q == 4) p = &G->int21[type][type2][si1][sq1][sj1];
  else if (q == 5) p = &G->int21[type2][type][sq1][si1][sp1];
  else if (q == 6) p = &G->int22[type][type2][si1][sp1][sq1][sj1];
  else {
    f *= H.mm23[mm_index(type, si1, sj1)] * H.mm23[mm_index(type2, sq1, sp1)];
    p = &H.one;
  }
  *fac = f;
  return p;
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

struct alignas(16) Int4 { int x, y, z, w; };

// One level of the recursion for the thread that owns an anti-diagonal.  d[Q] holds the sum of level S = 4 + par + 2Q; the level
// parity `par` is a run-time (warp-uniform) value, so both parities share one code path: everything that depends on it comes
// from tables indexed [par][Q] (size factors in the constant bank, row offsets in shared memory) or is folded into the limits
// (sm1 / sm2 = largest 4 + 2Q that is on for the row / column candidate).  ADDR1 / ADDR2 are the ADDRESSES of the two candidates
// with u1 = 2 / u2 = 2; a candidate that is off reads the always-zero word `zp` instead, so the step is branch-free: every lane of
// the warp runs the same instructions whatever its row.  Q is a compile-time constant, so d[] stays in registers.
// Hard constraints cost nothing per level: the gate of a candidate ("the S - 2 bases on the other side may stay unpaired") is
// monotone in the level, so the caller folds it into sm1 / sm2, and the gate of the carried sum ("i+1 and j-1 may stay unpaired")
// does not depend on the level, so the caller clears d[] instead.  The carried sum itself needs no limit: a level that is off
// here was off one step earlier on the anti-diagonal too, i.e. it carries zero.
#define ADP_LEVEL_STEP(Q, ADDR1, ADDR2)                                                \
  {                                                                                   \
    const double* p1 = (4 + 2 * (Q) <= sm1) ? ADDR1(Q) : zp;                          \
    const double* p2 = (4 + 2 * (Q) <= sm2 && ((Q) > 0 || par != 0)) ? ADDR2(Q) : zp; \
    double v = fcn[Q] * (*p1 + *p2);                                                  \
    if ((Q) > 0) v += d[(Q) > 0 ? (Q) - 1 : 0];                                       \
    if ((Q) > 0) gen += fci[Q] * v;                                                   \
    d[Q] = v;                                                                         \
  }
#define ADP_LEVEL_PASS(A1, A2)                                                         \
  ADP_LEVEL_STEP(13, A1, A2) ADP_LEVEL_STEP(12, A1, A2) ADP_LEVEL_STEP(11, A1, A2) ADP_LEVEL_STEP(10, A1, A2) ADP_LEVEL_STEP(9, A1, A2)   \
  ADP_LEVEL_STEP(8, A1, A2) ADP_LEVEL_STEP(7, A1, A2) ADP_LEVEL_STEP(6, A1, A2) ADP_LEVEL_STEP(5, A1, A2) ADP_LEVEL_STEP(4, A1, A2)     \
  ADP_LEVEL_STEP(3, A1, A2) ADP_LEVEL_STEP(2, A1, A2) ADP_LEVEL_STEP(1, A1, A2) ADP_LEVEL_STEP(0, A1, A2)
// the row offsets of one parity: 16 ints, four vector loads
#define ADP_LEVEL_OFFSETS(LVO)                                                                            \
  int lo[16];                                                                                             \
  {                                                                                                       \
    const Int4* lv4 = reinterpret_cast<const Int4*>(LVO);                                                 \
    const Int4 o0 = lv4[0], o1 = lv4[1], o2 = lv4[2], o3 = lv4[3];                                        \
    lo[0] = o0.x; lo[1] = o0.y; lo[2] = o0.z; lo[3] = o0.w; lo[4] = o1.x; lo[5] = o1.y; lo[6] = o1.z; lo[7] = o1.w;       \
    lo[8] = o2.x; lo[9] = o2.y; lo[10] = o2.z; lo[11] = o2.w; lo[12] = o3.x; lo[13] = o3.y; lo[14] = o3.z; lo[15] = o3.w; \
  }
#define ADP_LEVEL_ZERO()                                                                                  \
  d[0] = d[1] = d[2] = d[3] = d[4] = d[5] = d[6] = d[7] = d[8] = d[9] = d[10] = d[11] = d[12] = d[13] = 0.;

// Partial sum (this lane's share) of  sum_{k=i+1}^{nr} (Qm(i,k-1) [k-i >= 5] + eMLb[k-i] [k-i <= run]) * Qm1(k,j)  -- the terms
// that extend row i of Qm to the new column.  Lane `sub` of a group of G takes every G-th term.  Two plain loops
// (unpaired prefix; earlier multiloop part walked along row i with a second-order offset increment).
ADP_HD double qm_row_terms(const double* qm, const double* qm1, const double* eMLbS, int i, int nr, int run, int sub, int G) {
  double s0 = 0., s1 = 0.;
  {
    int umax = nr - i;
    if (umax > run) umax = run;
    const double* pe = eMLbS + 1 + sub;
    const double* pq = qm1 + i + 1 + sub;
    int u = 1 + sub;
    for (; u + G <= umax; u += 2 * G) {
      s0 += pe[0] * pq[0];
      s1 += pe[G] * pq[G];
      pe += 2 * G;
      pq += 2 * G;
    }
    if (u <= umax) s0 += pe[0] * pq[0];
  }
  {
    int k = i + 5 + sub;
    int off = t4(i, k - 1);
    int step = G * (k - 5) + ((G * (G - 1)) >> 1);      // column k-1 -> k-1+G
    const int step2 = G * G;
    const double* pq = qm1 + k;
    for (; k <= nr; k += G) {
      s1 += qm[off] * pq[0];
      off += step;
      step += step2;
      pq += G;
    }
  }
  return s0 + s1;
}

// ---------------------------------------------------------------------------------------------------
// Deferred work queue.  Work that is not on the column-to-column critical path (Qm of the previous column in the inside pass,
// the multiloop carrier A of the next column in the outside pass) is cut into warp-sized chunks; a warp that has nothing else
// to do in either interval of a column takes the next chunk.  One counter per column parity, reset a column ahead.
template <class E>
ADP_HD int grab_chunk(const E& ex, int* ctr) {
  int c = 0;
  if (ex.lane() == 0) c = ex.atomic_inc(ctr);
  return ex.warp_bcast(c, 0);
}


// ---------------------------------------------------------------------------------------------------
// inside pass.  Returns Z = Q(1,n) (scaled) on every thread.
//
// ONE barrier per column (round 2; round 1 had two).  Everything a thread does between barrier j-1 and barrier j ("phase j")
// reads only what was complete at barrier j-1, so the pieces of a phase may run in any order and on any warp:
//   cells(j)        Qb(i,j) of the listed cells, a lane group per cell; the group's first lane also writes the cell's share of
//                   Qm1(.,j) and of the premultiplied column vectors xA / x1n of column j
//   levels(j+1)     level sums of the generic loops of the NEXT column (reads columns <= j-2); thread-bound (registers)
//   list(j+1)       the cells of the next column and their hairpin weights
//   rest(j)         rows of column j that cannot pair: unpaired extension of Qm1, zeros in xA / x1n
//   q1(j-1)         exterior prefix from the finished column of Qb (one warp)
//   scan(j-1)       unpaired-prefix half of Qm(., j-1) from the finished Qm1 column (one warp)
//   drain(j-2)      the other half of Qm(., j-2), from a queue of warp-sized chunks: whoever is idle takes the next one
// Cells are assigned from thread 0 upwards; the list / rest rows, the two single-warp jobs and (being last in program order)
// the chunk queue fall to the warps that hold no cells, so the phase ends when the busiest warp does.  Buffers that are
// written for the next column while the current one is read are double-buffered by column parity, those with a third
// live column (Qm1, x1n, the list and queue counters) rotate mod 3.
enum FastVec : int {
  kVQ1 = 0, kVEMLb = 1,                     // shared with the outside pass: exterior prefix q1, eMLb powers
  kVQm1 = 2,                                // +3: Qm1 columns by column mod 3
  kVXA = 5,                                 // +2: Qb * terminal factor (bulge filter), by column parity
  kVX1n = 7,                                // +3: Qb * 1xn mismatch ratio, by column mod 3
  kVHp = 10,                                // +2: hairpin weight of the listed cells, by column parity
  kVGen0 = 12, kVGen1 = 14,                 // +2 each: generic-loop sums per level parity, by column parity
  kVU = 16                                  // +2: unpaired-prefix half of Qm by column parity
};

template <bool kCst, class E>
ADP_HD double fast_inside(const E& ex, const SeqCtx& c, const PfHot& H, const PfTables* G, double motif_boltz, FastWork& w) {
  const int n = c.n, nt = ex.nthreads(), tid = ex.tid(), ld = w.ld, R = w.R;
  const int nt_lg = ilog2_floor((unsigned)nt), warp_lg = ilog2_floor((unsigned)E::kWarp);
  const int nwarp = nt >> warp_lg;
  const bool mem_state = E::kSerial || w.dstate != nullptr;   // level sums in memory: host emulation, or more anti-diagonals than threads
  const int half = mem_state ? R : (nt >> 1);        // anti-diagonal slots per level parity (>= R)
  const int ntri = (int)t4_count(n);
  double* const qb = w.qb;
  double* const qm = w.qm;
  const uint8_t* const cn = w.cn;
  const uint8_t* const cp = w.cp;
  for (int k = tid; k < ntri; k += nt) { qb[k] = 0.; qm[k] = 0.; }
  for (int k = tid; k < kFastVectors * ld; k += nt)
    if (k < kVEMLb * ld || k >= (kVEMLb + 1) * ld) w.vbase[k] = 0.;
  double* const q1 = w.vecp(kVQ1);
  const double* const eMLbS = w.vecp(kVEMLb);
  double d[kFastLevels];     // level sums of this thread's anti-diagonal (registers)
  const double* const zp = &H.pad;   // an always-zero word for levels that are off
  ADP_LEVEL_ZERO()
  ADP_PROF_DECL
  if (mem_state)
    for (int k = tid; k < 2 * R * kFastLevels; k += nt) w.dstate[k] = 0.;
  if (tid == 0)
    for (int k = 0; k < 3; k++) { w.counter[k] = 0; w.qctr[k] = 0; }
  for (int t = tid; t < 64; t += nt) w.lvoff[t] = 0;
  ex.sync();
  const double s1 = H.scale1, eb1 = H.eMLb1;
  if (tid == 0) {
    // columns without cells: only the exterior prefix moves
    q1[0] = 1.;
    for (int j = 1; j <= 4 && j <= n; j++) q1[j] = (!kCst || can_unpair(c, j)) ? q1[j - 1] * s1 : 0.;
  }

  // level sums of the generic interior loops of column jn (reads Qb columns <= jn-3)
  auto levels = [&](int jn) {
    const int nr = jn - 4;
    const bool cuj1 = !kCst || can_unpair(c, jn - 1);
    const int upl_j1 = kCst ? (int)c.upl[jn - 1] : 1000;
    const double* col3 = qb + (((jn - 7) * (jn - 8)) >> 1) - 2;   // Qb*FI(k, jn-3) = col3[k + 1]
    double* const gen0 = w.vecp(kVGen0 + (jn & 1));
    double* const gen1 = w.vecp(kVGen1 + (jn & 1));
    const int* const lvo = w.lvoff + ((jn & 1) << 5);
    const int kbase = ((jn + 1) / R) * R;
    for (int task = tid; task < 2 * half; task += nt) {
      const int par = task >= half ? 1 : 0;          // warp-uniform: the two parities are separate code paths
      const int slot = task - par * half;
      if (slot >= R) continue;
      if (mem_state)
        for (int q = 0; q < kFastLevels; q++) d[q] = w.dstate[(size_t)task * kFastLevels + q];
      const int i = fast_slot_key(slot, kbase, jn, R) - jn;
      if (i >= 1 && i <= nr) {
        int sm = jn - i - 6;
        if (sm > kMaxLoop) sm = kMaxLoop;
        sm -= par;                                     // level Q is on iff 4 + 2Q <= sm
        double gen = 0.;
        if (sm + par >= 4) {
          int sm1 = sm, sm2 = sm;
          if (kCst) {
            const int upr_i1 = (int)c.upr[i + 1];
            // constraint gates: two unpaired bases on the candidate's own side, S - 2 on the other side (monotone in the level)
            const int lim1 = 4 + upl_j1 - par - 2, lim2 = 4 + upr_i1 - par - 2;
            sm1 = upr_i1 >= 2 ? (sm < lim1 ? sm : lim1) : -1;
            sm2 = upl_j1 >= 2 ? (sm < lim2 ? sm : lim2) : -1;
            if (!(upr_i1 >= 1 && cuj1)) { ADP_LEVEL_ZERO() }      // i+1 or j-1 may not stay unpaired: nothing is carried
          }
          const double* xi = col3 + i + par + 4;       // Qb*FI(i+S-1, jn-3) = xi[2Q]
          const char* qrow = reinterpret_cast<const char*>(qb + i + 2);   // Qb*FI(i+3, l) = qrow[(l-4)(l-5)/2], l = jn+1-S: offset from the table
          const double* fcn = ADP_FC.nin + (par << 4);
          const double* fci = ADP_FC.ip + (par << 4);
          ADP_LEVEL_OFFSETS(lvo + (par << 4))
#define ADP_A1(Q) reinterpret_cast<const double*>(qrow + lo[Q])
#define ADP_A2(Q) (xi + 2 * (Q))
          ADP_LEVEL_PASS(ADP_A1, ADP_A2)
#undef ADP_A1
#undef ADP_A2
        } else {
          ADP_LEVEL_ZERO()
        }
        (par ? gen1 : gen0)[i] = gen;
      }
      if (mem_state)
        for (int q = 0; q < kFastLevels; q++) w.dstate[(size_t)task * kFastLevels + q] = d[q];
    }
  };

  // the cells of column jn and their hairpin weights, a thread per row, rows handed out from the last thread downwards (the
  // first threads hold the cells of the current column)
  auto list_rows = [&](int jn) {
    const int nr = jn - 4;
    double* const hp = w.vecp(kVHp + (jn & 1));
    uint16_t* const cells = w.cells + (jn & 1) * ld;
    uint8_t* const iscell = w.iscell + (jn & 1) * ld;
    int* const counter = w.counter + (jn % 3);
    ADP_NOUNROLL
    for (int r = nt - 1 - tid; r < nr; r += nt) {
      const int i = r + 1;
      const int type = kCst ? allowed_type(c, i, jn) : pair_type(c.S[i], c.S[jn]);
      iscell[i] = type != 0;
      if (type != 0) {
        cells[ex.atomic_inc(counter)] = (uint16_t)i;
        hp[i] = fast_hairpin<kCst>(c, H, G, w, motif_boltz, i, jn, type);
      }
    }
  };

  // rows of column j that hold no cell
  auto rest_rows = [&](int j) {
    const int nr = j - 4;
    const bool cuj = !kCst || can_unpair(c, j);
    const double* const qm1_prev = w.vecp(kVQm1 + (j - 1) % 3);
    double* const qm1_cur = w.vecp(kVQm1 + j % 3);
    double* const xA = w.vecp(kVXA + (j & 1));
    double* const x1n = w.vecp(kVX1n + j % 3);
    const uint8_t* const iscell = w.iscell + (j & 1) * ld;
    ADP_NOUNROLL
    for (int r = nt - 1 - tid; r < nr; r += nt) {
      const int i = r + 1;
      if (iscell[i]) continue;
      qm1_cur[i] = cuj ? qm1_prev[i] * eb1 : 0.;
      xA[i] = 0.;
      x1n[i] = 0.;
    }
  };

  // exterior prefix of the finished column jq: q1[jq] = [jq unpaired] q1[jq-1] scale + sum_i q1[i-1] Qb(i,jq) stem(i,jq), by one
  // warp straight from the column of Qb (summed in row order: the result does not depend on the list order)
  auto q1_task = [&](int jq) {
    const double* const col = qb + t4(1, jq) - 1;        // Qb (premultiplied) (i, jq) = col[i]
    const int cnq = cn[jq] << 4, sq = c.S[jq];
    double e = 0.;
    ADP_NOUNROLL
    for (int i = 1 + ex.lane(); i <= jq - 4; i += E::kWarp) {
      const double v = col[i];
      if (v != 0.) {
        const double plain = (i > 1 && jq < n) ? v * H.RI[cnq | cp[i]] : v;      // back to the plain Qb(i,jq)
        e += plain * fast_stem_ext(c, H, pair_type(c.S[i], sq), i, jq) * q1[i - 1];
      }
    }
    e = ex.warp_sum(e);
    if (ex.lane() == 0) q1[jq] = ((!kCst || can_unpair(c, jq)) ? q1[jq - 1] * s1 : 0.) + e;
  };

  // U(., jq) for the finished Qm1 column jq, by one warp: every lane folds a block of rows into an affine map, a suffix scan
  // over the lanes composes the maps, a second pass over the block writes the values
  auto scan_unpaired = [&](int jq) {
    const int nrq = jq - 4;
    if (nrq < 1) return;
    double* Uv = w.vecp(kVU + (jq & 1));
    const double* x = w.vecp(kVQm1 + jq % 3);
    const int lane = ex.lane();
    const int B = (nrq + E::kWarp - 1) / E::kWarp;
    const int lo = lane * B + 1;
    int hi = lo + B - 1;
    if (hi > nrq) hi = nrq;
    double A = 1., Cc = 0.;                       // U(lo) = A * U(hi+1) + Cc
    ADP_NOUNROLL
    for (int i = hi; i >= lo; i--) {
      const double m = (i < nrq && (!kCst || can_unpair(c, i))) ? eb1 : 0.;      // U(nrq) = 0: no row below it
      Cc = m * (Cc + x[i + 1]);
      A = m * A;
    }
    ADP_NOUNROLL
    for (int o = 1; o < E::kWarp; o <<= 1) {      // higher lanes hold the higher rows: suffix scan
      const double A2 = ex.shfl_down_d(A, o), C2 = ex.shfl_down_d(Cc, o);
      if (lane + o < E::kWarp) { Cc = A * C2 + Cc; A = A * A2; }
    }
    double un = ex.shfl_down_d(Cc, 1);            // U at the first row of the next lane's block
    if (lane + 1 >= E::kWarp) un = 0.;
    ADP_NOUNROLL
    for (int i = hi; i >= lo; i--) {
      const double m = (i < nrq && (!kCst || can_unpair(c, i))) ? eb1 : 0.;
      un = m * (un + x[i + 1]);
      Uv[i] = un;
    }
  };

  // Qm(., jq) from the finished Qm1 column jq and its scan: row pairs (rows paired so that every pair sums the same number of
  // terms), a few lanes per pair, a warp-sized chunk of pairs at a time from the queue.  Qm(., jq) is first read by column jq + 5.
  auto drain_qm = [&](int jq) {
    const int nrq = jq - 4;
    if (nrq < 1) return;
    const int npair = (nrq + 1) >> 1;
    const int lgc = E::kSerial ? 0 : (nrq >= 24 ? 2 : 3);
    const int Gc = 1 << lgc, per = E::kWarp >> lgc, subC = ex.lane() & (Gc - 1);
    const int nchunk = (npair + per - 1) / per;
    const double* qm1q = w.vecp(kVQm1 + jq % 3);
    const double* Uq = w.vecp(kVU + (jq & 1));
    int* ctr = w.qctr + (jq % 3);
    for (;;) {
      const int ch = grab_chunk(ex, ctr);
      if (ch >= nchunk) break;
      const int slot = ch * per + (ex.lane() >> lgc);
      const bool valid = slot < npair;
      const int iA = slot + 1, iB = nrq - slot;       // iA <= iB; equal for the middle row of an odd count
      double tA = 0., tB = 0.;
      if (valid) {      // run = 0: the unpaired-prefix terms come from the scan
        tA = qm_row_terms(qm, qm1q, eMLbS, iA, nrq, 0, subC, Gc);
        if (iB > iA) tB = qm_row_terms(qm, qm1q, eMLbS, iB, nrq, 0, subC, Gc);
      }
      ex.group_sum2_rt(tA, tB, Gc);
      if (valid && subC == 0) {
        qm[t4(iA, jq)] = tA + qm1q[iA] + Uq[iA];
        if (iB > iA) qm[t4(iB, jq)] = tB + qm1q[iB] + Uq[iB];
      }
    }
  };

  ex.sync();
  // j = 4 is a warm-up turn without cells: it only prepares column 5; j = n+1 is a run-out turn without cells: it finishes the
  // exterior prefix of column n and Qm(., n-1), which the outside pass reads (one call site of every piece keeps the code small)
  for (int j = 4; j <= n + 1; j++) {
    const bool cuj = !kCst || can_unpair(c, j);
    const double* qm1_prev = w.vecp(kVQm1 + (j - 1) % 3);
    double* qm1_cur = w.vecp(kVQm1 + j % 3);
    const int m = w.counter[j % 3];
    if (tid == 0) { w.counter[(j + 2) % 3] = 0; w.qctr[(j + 2) % 3] = 0; }      // the slots of phase j+2 (nobody reads them in this phase)
    ADP_NOUNROLL
    for (int t = tid; t < 32; t += nt) {
      // byte offsets of the row candidates of the level recursion of column j+2 (evaluated in phase j+1): level S reads Qb*FI(i+3, j+3-S)
      const int S = 4 + (t >> 4) + 2 * (t & 15);
      const int l = j + 3 - S;
      w.lvoff[((j & 1) << 5) + t] = (((l - 4) * (l - 5)) >> 1) * (int)sizeof(double);
    }
    // level sums first: every warp runs the same code at the same time (instruction-cache locality), then the warps split
    // into those that hold cells and those that do the rest of the phase
    if (j < n) levels(j + 1);
    // ---- cells(j): Qb(i,j) for the listed cells, a lane group per cell
    if (m > 0) {
      const double* const hp = w.vecp(kVHp + (j & 1));
      const double* const gen0 = w.vecp(kVGen0 + (j & 1));
      const double* const gen1 = w.vecp(kVGen1 + (j & 1));
      const double* const xA_rd = w.vecp(kVXA + ((j - 1) & 1));      // column j-1
      const double* const x1n_rd = w.vecp(kVX1n + (j - 2) % 3);      // column j-2
      double* const xA_wr = w.vecp(kVXA + (j & 1));
      double* const x1n_wr = w.vecp(kVX1n + j % 3);
      const uint16_t* const cells = w.cells + (j & 1) * ld;
      const int lg = group_lg(m, nt_lg, warp_lg);
      const int Gs = 1 << lg, slots = nt >> lg, sub = tid & (Gs - 1);
      const int upl_j1 = kCst ? (int)c.upl[j - 1] : 1000;
      const int sj = c.S[j], cpj = cp[j], cnj = cn[j] << 4;
      for (int slot0 = 0; slot0 < m; slot0 += slots) {
        const int slot = slot0 + (tid >> lg);
        const bool valid = slot < m;
        const int i = valid ? cells[slot] : 1;
        double acc = 0.;
        if (valid) {
          const int si = c.S[i];
          const int type = pair_type(si, sj);
          const int codeij = (cn[i] << 4) | cpj;
          const int room = j - i - 6;                 // u1 + u2 <= room keeps the inner pair >= 4 apart
          int u1cap = kCst ? (int)c.upr[i + 1] : kMaxLoop, u2cap = upl_j1;
          if (u1cap > kMaxLoop) u1cap = kMaxLoop;
          if (u2cap > kMaxLoop) u2cap = kMaxLoop;
          double other = 0., accB = 0., acc1n = 0., accML = 0.;
          double cval = 0., ctab = 0.;
          if (sub == 0) other = hp[i] + H.FI[codeij] * (gen0[i] + gen1[i]);
          ADP_TICK(8)
          if (room >= 0) {
            // corner candidates through the code-indexed table, one per lane and round.  Software-pipelined and kept rolled (code
            // size): the table load of a round is issued in that round and consumed in the next, the last one after the filters below.
            if (sub == Gs - 1) {
              // the stack (u1 = u2 = 0) as straight-line code on the group's last lane: both types come from the cell's own codes
              const double v = qb[t4(i + 1, j - 1)] * H.RI[(cn[j - 1] << 4) | cp[i + 1]];
              other += v * H.cfac[0] * H.stack[type * 8 + H.tyr[codeij]];
            }
            const double* const ctq = ADP_FC.corner + ((size_t)codeij << 8) - ((size_t)1 << 16);
            ADP_NOUNROLL
            for (int q = 1 + sub; q < 9; q += Gs) {
              int u1, u2;
              corner_u(q, &u1, &u2);
              double v2 = 0.;
              const double* p2 = zp;
              if (u1 <= u1cap && u2 <= u2cap && u1 + u2 <= room) {
                const int k = i + 1 + u1, l = j - 1 - u2;
                const int cin = (cn[l] << 4) | cp[k];
                const double v = qb[t4(k, l)] * H.RI[cin];
                if (v != 0.) {
                  p2 = ctq + ((size_t)q << 16) + cin;
                  v2 = v;
                }
              }
              const double t2 = *p2;
              other += cval * ctab;
              cval = v2;
              ctab = t2;
            }
            ADP_TICK(9)
            // The four 1-D filters along the loop's axes, fused into one loop so that their load chains overlap:
            //   bulges with the unpaired run on the 3' side (k = i+1, l = j-1-u, u >= 2) and on the 5' side (l = j-1, k = i+1+u:
            //   premultiplied column vector xA);  1 x u loops (k = i+2, l = j-1-u, u >= 3) and u x 1 loops (l = j-2, k = i+1+u:
            //   column vector x1n).  tap1n[2] = 0.  A term beyond its own limit is selected away after the (always valid) loads.
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
              const double* xa = xA_rd + i + 1;
              const double* xn = x1n_rd + i + 1;
              for (int u = 2 + sub; u <= hmax; u += Gs) {
                const int l = j - 1 - u;                       // >= i + 5: (i+1,l) is a cell; (i+2,l) is one unless u = room
                const int cl = cn[l] << 4;
                const double* pb = qb + t4(i + 1, l);
                const double b3 = pb[0] * H.RB[cl | cpk1], n3 = pb[1] * H.R1[cl | cpk2];
                const double b5 = xa[u], n5 = xn[u];
                accB += H.tapB[u] * ((u <= hiB3 ? b3 : 0.) + (u <= hiB5 ? b5 : 0.));
                acc1n += H.tap1n[u] * ((u <= hi13 ? n3 : 0.) + (u <= hi31 ? n5 : 0.));
              }
            }
          }
          ADP_TICK(10)
          // multiloop closed by (i,j): Qm(i+1,k-1) * Qm1(k,j-1), k = i+6 .. j-5
          {
            int k = i + 6 + sub;
            int off = t4(i + 1, k - 1);
            int step = Gs * (k - 5) + ((Gs * (Gs - 1)) >> 1);     // column k-1 -> k-1+Gs
            const int step2 = Gs * Gs;
            const double* pq = qm1_prev + k;
            for (; k <= j - 5; k += Gs) {
              accML += qm[off] * pq[0];
              off += step;
              step += step2;
              pq += Gs;
            }
          }
          // ---- cell total
          other += cval * ctab;
          acc = other + au_factor(H, si, sj) * accB + H.R1[codeij] * H.FI[codeij] * acc1n + accML * (H.MLclosing * (H.RM[codeij] * H.FI[codeij]) * H.scale2);
        }
        ADP_TICK(11)
        acc = ex.group_sum_rt(acc, Gs);
        ADP_TICK(12)
        if (valid && sub == 0) {
          const int code = cnj | cp[i];
          const double pre = (i > 1 && j < n) ? acc * H.FI[code] : acc;     // stored premultiplied; rows / columns at the sequence
          qb[t4(i, j)] = pre;                                               // ends are never read back by the loop terms
          xA_wr[i] = pre * H.RB[code];
          x1n_wr[i] = pre * H.R1[code];
          qm1_cur[i] = (cuj ? qm1_prev[i] * eb1 : 0.) + acc * fast_stem_ml(c, H, w, pair_type(c.S[i], sj), i, j);
        }
        ADP_TICK(13)
      }
    }
    ADP_TICK(0)
    // ---- everything else of the phase
    ADP_TICK(4)
    if (j < n) list_rows(j + 1);
    if (j >= 5 && j <= n) rest_rows(j);
    ADP_TICK(5)
    if (j >= 6 && ex.warp() == nwarp - 1) q1_task(j - 1);
    if (j >= 6 && j <= n && ex.warp() == (nwarp > 1 ? nwarp - 2 : 0)) scan_unpaired(j - 1);
    ADP_TICK(3)
    if (j >= 7) drain_qm(j - 2);
    ADP_TICK(6)
    ex.sync();
    ADP_TICK(7)
  }
  ADP_PROF_REPORT("inside ")
  return q1[n];
}

// ---------------------------------------------------------------------------------------------------
// outside pass.  Overwrites Qb column by column (l = n .. 5) with the outside weights; the probabilities of a
// column go to `bppm` / the pair queries one phase after the column is complete.
//
// ONE barrier per column, like the inside pass.  Phase l (between barrier l+1 and barrier l) holds
//   cells(l)        outside weight of every listed pair (k,l), a lane group per cell; the group's first lane also writes the
//                   pair probability, the exterior-suffix share, the cell's entries of the premultiplied column vectors xA / x1n
//                   and of the multiloop carrier B of column l-1
//   rest(l)         rows of column l without a cell: zeros in P / xA / x1n, decay of the carrier B
//   levels(l-1)     level sums of the generic loops enclosing the cells of the NEXT column (reads columns >= l+2); thread-bound
//   list(l-1)       the cells of the next column (pairs with a non-zero inside weight)
//   qln(l-1)        exterior suffix (one thread)
//   results(l+1)    probabilities of the finished column to bppm / the queries
//   carriers(l-1)   multiloop carrier A of the next column (reads columns >= l+5), from a queue of warp-sized chunks
// Double-buffered by column parity: A, B, P, xA, the generic sums, the lists; mod 3: x1n and the counters.
enum FastVecOut : int {
  kOA = 2,                                  // +2: A(i,l): multiloop carrier with >= 1 stem right of l, by column parity
  kOB = 4,                                  // +2: B(i,l): nothing right of l, by column parity
  kOP = 6,                                  // +2: P(.,l), by column parity
  kOXA = 8,                                 // +2: Pout * terminal factor, by column parity
  kOX1n = 10,                               // +3: Pout * 1xn mismatch ratio, by column mod 3
  kOGen0 = 13, kOGen1 = 15,                 // +2 each: generic-loop sums per level parity, by column parity
  kOQln = 17,                               // exterior suffix Q(l,n)
  kOAcc = 18,                               // exterior-suffix accumulators
  kOPm = 19                                 // bound-motif probabilities per site
};

struct FastOut {
  double* bppm;          // NULL or n(n+1)/2 doubles, row-major upper triangle incl. diagonal, 0-based
  int n_query;
  const int32_t* qi;
  const int32_t* qj;
  double* qprobs;
  int n_dens;            // positions (0-based) whose probability of being paired is wanted
  const int32_t* dens_pos;
  double* dens;
  // SITE MODE (site_k > 0): only the outside weights of the cells that enclose the pair (site_k, site_l) are evaluated -- rows
  // k <= site_k, columns l >= site_l, which depend on nothing outside that corner -- and *site_ph receives the probability that
  // (site_k, site_l) closes a plain hairpin loop.  With it the ligand-bound ensemble of a construct with ONE motif site follows
  // from the apo fill:  Z_holo = Z_apo (1 + (boltz - 1) P_h), because binding only reweights the structures that hold that hairpin.
  int site_k = 0, site_l = 0;
  double* site_ph = nullptr;
};

ADP_HD size_t bppm_index(int n, int i, int j) {      // 1-based i <= j
  return (size_t)(i - 1) * n - (size_t)(i - 1) * (i - 2) / 2 + (size_t)(j - i);
}

template <bool kCst, class E>
ADP_HD void fast_outside(const E& ex, const SeqCtx& c, const PfHot& H, const PfTables* G, const MotifDev* motif, double motif_boltz,
                         FastWork& w, const FastOut& out) {
  const int n = c.n, nt = ex.nthreads(), tid = ex.tid(), ld = w.ld, R = w.R;
  const int nt_lg = ilog2_floor((unsigned)nt), warp_lg = ilog2_floor((unsigned)E::kWarp);
  const bool mem_state = E::kSerial || w.dstate != nullptr;   // level sums in memory: host emulation, or more anti-diagonals than threads
  const int half = mem_state ? R : (nt >> 1);        // anti-diagonal slots per level parity (>= R)
  double* const qb = w.qb;
  const double* const qm = w.qm;
  const uint8_t* const cn = w.cn;
  const uint8_t* const cp = w.cp;
  for (int k = tid; k < kFastVectors * ld; k += nt)
    if (k >= 2 * ld) w.vbase[k] = 0.;      // everything but q1 and the eMLb powers
  const double* const q1 = w.vecp(kVQ1);
  const double* const eMLbS = w.vecp(kVEMLb);
  double* const qln = w.vecp(kOQln);
  double* const accq = w.vecp(kOAcc);
  double* const pm = w.vecp(kOPm);
  if (out.bppm) {
    const size_t ntot = (size_t)n * (n + 1) / 2;
    for (size_t k = tid; k < ntot; k += nt) out.bppm[k] = 0.;
  }
  for (int q = tid; q < out.n_query; q += nt) out.qprobs[q] = 0.;
  for (int q = tid; q < out.n_dens; q += nt) out.dens[q] = 0.;
  double d[kFastLevels];
  const double* const zp = &H.pad;
  ADP_LEVEL_ZERO()
  ADP_PROF_DECL
  if (mem_state)
    for (int k = tid; k < 2 * R * kFastLevels; k += nt) w.dstate[k] = 0.;
  if (tid == 0)
    for (int k = 0; k < 3; k++) { w.counter[k] = 0; w.qctr[k] = 0; }
  for (int t = tid; t < 64; t += nt) w.lvoff[t] = 0;
  ex.sync();
  if (tid == 0) { qln[n + 1] = 1.; qln[n + 2] = 0.; }
  const double Z = q1[n];
  const double invZ = 1. / Z;
  const double s1 = H.scale1, eb1 = H.eMLb1;
  const bool has_motif = motif && motif->len > 0;
  const bool site_mode = out.site_k > 0;
  const int row_max = site_mode ? out.site_k : n;          // rows / columns that are evaluated at all
  const int l_min = site_mode ? out.site_l : kTurn + 1;
  if (site_mode) {
    // the exterior suffix Q(l,n) right of the site, which the full pass accumulates from rows that are skipped here: position l
    // unpaired, or paired with some l' (one warp, a row of Qb per step)
    ex.sync();
    if (ex.warp() == 0) {
      for (int l = n; l > out.site_l; l--) {
        double e = 0.;
        const int cpl = cp[l], sl = c.S[l];
        ADP_NOUNROLL
        for (int j = l + kTurn + 1 + ex.lane(); j <= n; j += E::kWarp) {
          const double v = qb[t4(l, j)];
          if (v != 0.) {
            const double plain = (l > 1 && j < n) ? v * H.RI[(cn[j] << 4) | cpl] : v;
            e += plain * fast_stem_ext(c, H, pair_type(sl, c.S[j]), l, j) * qln[j + 1];
          }
        }
        e = ex.warp_sum(e);
        if (ex.lane() == 0) qln[l] = ((!kCst || can_unpair(c, l)) ? qln[l + 1] * s1 : 0.) + e;
        ex.warp_sync();
      }
    }
    if (tid == 0) *out.site_ph = 0.;
    ex.sync();
  }

  // multiloop carrier A(., ln) for every possible 5' closing position i <= ln-5: rows from a queue, a warp-sized chunk at a time
  auto carriers = [&](int ln) {
    const int nrow = ln - 5 < row_max - 1 ? ln - 5 : row_max - 1;      // rows a listed cell can read: i < k <= row_max
    if (ln >= n || nrow < 1) return;
    double* const vA = w.vecp(kOA + (ln & 1));
    const int lg = E::kSerial ? 0 : group_lg(nrow, nt_lg, warp_lg);      // rows fill the CTA once (1, 2, 4 and 8 lanes per row measured: all slower)
    const int Ga = 1 << lg, per = E::kWarp >> lg, subA = ex.lane() & (Ga - 1);
    const int gstep = (Ga * (Ga - 1)) >> 1;
    const int nchunk = (nrow + per - 1) / per;
    int* ctr = w.qctr + (ln % 3);
    for (;;) {
      const int ch = grab_chunk(ex, ctr);
      if (ch >= nchunk) break;
      const int i = ch * per + (ex.lane() >> lg) + 1;
      const bool valid = i <= nrow;
      double a = 0.;
      if (valid) {
        const int cni = cn[i] << 4;
        int j = ln + 6 + subA;
        int offp = t4(i, j), offm = t4(ln + 1, j - 1);
        int stepp = Ga * (j - 4) + gstep;          // column j -> j+Ga (and j-1 -> j-1+Ga is Ga less)
        const int step2 = Ga * Ga;
        const uint8_t* pc = cp + j;
        for (; j <= n; j += Ga) {
          a += qb[offp] * H.RM[cni | pc[0]] * qm[offm];
          offp += stepp;
          offm += stepp - Ga;
          stepp += step2;
          pc += Ga;
        }
      }
      a = ex.group_sum_rt(a, Ga) * H.MLclosing;
      if (valid && subA == 0) vA[i] = a;
    }
  };

  // level sums of the generic interior loops that enclose (k,ln) (reads columns >= ln+3)
  auto levels = [&](int ln) {
    const int nr = ln - 4;
    const int upr_l1 = kCst ? (int)c.upr[ln + 1] : n - ln;           // unpaired run right of ln (0 at the 3' end)
    const double* col3 = qb + (((ln - 1) * (ln - 2)) >> 1);           // Pout*FI(i, ln+3) = col3[i - 1]
    double* const gen0 = w.vecp(kOGen0 + (ln & 1));
    double* const gen1 = w.vecp(kOGen1 + (ln & 1));
    const int* const lvo = w.lvoff + ((ln & 1) << 5);
    const int kbase = ((ln + 1) / R) * R;
    for (int task = tid; task < 2 * half; task += nt) {
      const int par = task >= half ? 1 : 0;
      const int slot = task - par * half;
      if (slot >= R) continue;
      if (mem_state)
        for (int q = 0; q < kFastLevels; q++) d[q] = w.dstate[(size_t)task * kFastLevels + q];
      const int k = fast_slot_key(slot, kbase, ln, R) - ln;
      if (k >= 1 && k <= nr && k <= row_max) {
        int sm = (k - 2) + (n - ln - 1);            // u1 <= k-2, u2 <= n-ln-1
        if (sm > kMaxLoop) sm = kMaxLoop;
        sm -= par;                                  // level Q is on iff 4 + 2Q <= sm
        double gen = 0.;
        if (k >= 4 && ln + 3 <= n && sm + par >= 4) {
          const int room1 = n + 1 - ln - par;        // row candidate needs j = ln+S-1 <= n
          const int room2 = k - par;                 // column candidate needs i = k+1-S >= 1
          int sm1 = sm < room1 ? sm : room1, sm2 = sm < room2 ? sm : room2;
          if (kCst) {
            const int upl_k1 = (int)c.upl[k - 1];
            // constraint gates: two unpaired bases on the candidate's own side, S - 2 on the other side (monotone in the level)
            const int lim1 = 4 + upr_l1 - par - 2, lim2 = 4 + upl_k1 - par - 2;
            sm1 = upl_k1 >= 2 ? (sm1 < lim1 ? sm1 : lim1) : -1;
            sm2 = upr_l1 >= 2 ? (sm2 < lim2 ? sm2 : lim2) : -1;
            if (!(upl_k1 >= 1 && upr_l1 >= 1)) { ADP_LEVEL_ZERO() }      // k-1 or ln+1 may not stay unpaired: nothing is carried
          }
          const char* qrow = reinterpret_cast<const char*>(qb + k - 4);   // Pout*FI(k-3, j) = qrow[(j-4)(j-5)/2], j = ln+S-1: offset from the table
          const double* xi = col3 + k - par - 4;     // Pout*FI(k+1-S, ln+3) = xi[-2Q]
          const double* fcn = ADP_FC.nin + (par << 4);
          const double* fci = ADP_FC.ip + (par << 4);
          ADP_LEVEL_OFFSETS(lvo + (par << 4))
#define ADP_A1(Q) reinterpret_cast<const double*>(qrow + lo[Q])
#define ADP_A2(Q) (xi - 2 * (Q))
          ADP_LEVEL_PASS(ADP_A1, ADP_A2)
#undef ADP_A1
#undef ADP_A2
        } else {
          ADP_LEVEL_ZERO()
        }
        (par ? gen1 : gen0)[k] = gen;
      }
      if (mem_state)
        for (int q = 0; q < kFastLevels; q++) w.dstate[(size_t)task * kFastLevels + q] = d[q];
    }
  };

  // the cells of column ln that carry weight, rows handed out from the last thread downwards
  auto list_rows = [&](int ln) {
    const int nr = ln - 4;
    uint16_t* const cells = w.cells + (ln & 1) * ld;
    uint8_t* const iscell = w.iscell + (ln & 1) * ld;
    int* const counter = w.counter + (ln % 3);
    const double* const col = qb + t4(1, ln) - 1;
    ADP_NOUNROLL
    for (int r = nt - 1 - tid; r < nr; r += nt) {
      const int k = r + 1;
      const bool on = k <= row_max && col[k] != 0.;
      iscell[k] = on;
      if (on) cells[ex.atomic_inc(counter)] = (uint16_t)k;
    }
  };

  // rows of column l without a cell (and the three rows behind the column's end, which the filters of later columns may load)
  auto rest_rows = [&](int l) {
    const int nr = l - 4;
    const bool cul = !kCst || can_unpair(c, l);
    double* const pcol = w.vecp(kOP + (l & 1));
    double* const xA = w.vecp(kOXA + (l & 1));
    double* const x1n = w.vecp(kOX1n + l % 3);
    const double* const vB_old = w.vecp(kOB + (l & 1));
    double* const vB_new = w.vecp(kOB + ((l - 1) & 1));
    const uint8_t* const iscell = w.iscell + (l & 1) * ld;
    ADP_NOUNROLL
    for (int r = nt - 1 - tid; r < nr + 3; r += nt) {
      const int k = r + 1;
      if (k <= nr && iscell[k]) continue;
      xA[k] = 0.;
      x1n[k] = 0.;
      if (k <= nr) {
        pcol[k] = 0.;
        if (l < n + 1 && k <= l - 6) vB_new[k] = cul ? vB_old[k] * eb1 : 0.;
      }
    }
  };

  // probabilities of the finished column lq: to the full matrix and to the pair queries
  auto results = [&](int lq) {
    const int nr = lq - 4;
    const double* const pcol = w.vecp(kOP + (lq & 1));
    if (out.bppm)
      ADP_NOUNROLL
      for (int r = nt - 1 - tid; r < nr; r += nt)
        if (pcol[r + 1] != 0.) out.bppm[bppm_index(n, r + 1, lq)] = pcol[r + 1];
    for (int q = nt - 1 - tid; q < out.n_query; q += nt) {
      int a = out.qi[q], b = out.qj[q];
      if (a > b) { const int t = a; a = b; b = t; }
      if (b + 1 == lq && a >= 0 && a + 1 <= nr) out.qprobs[q] = pcol[a + 1];
    }
    // paired probability of the listed positions: position lq collects its whole column, a row position its entry of the
    // column (one thread per listed position, columns in descending order: a fixed summation order)
    for (int q = nt - 1 - tid; q < out.n_dens; q += nt) {
      const int p = out.dens_pos[q] + 1;
      double s = 0.;
      if (p == lq) {
        for (int k = 1; k <= nr; k++) s += pcol[k];
      } else if (p >= 1 && p <= nr) {
        s = pcol[p];
      }
      if (s != 0.) out.dens[q] += s;
    }
  };

  ex.sync();
  ADP_TICK(1)
  // l = n+1 is a warm-up turn without cells: it only prepares column n; l = 4 is a run-out turn: it writes the results of
  // column 5 (one call site of every piece keeps the code small)
  for (int l = n + 1; l >= l_min; l--) {
    const int m = l <= n ? w.counter[l % 3] : 0;
    if (tid == 0) { w.counter[(l + 1) % 3] = 0; w.qctr[(l + 1) % 3] = 0; }      // the slots of phase l-2 (nobody reads them in this phase)
    ADP_NOUNROLL
    for (int t = tid; t < 32; t += nt) {
      // byte offsets of the row candidates of the level recursion of column l-2 (evaluated in phase l-1): level S reads Pout*FI(k-3, l+S-3)
      const int S = 4 + (t >> 4) + 2 * (t & 15);
      const int j = l + S - 3;
      w.lvoff[((l & 1) << 5) + t] = (((j - 4) * (j - 5)) >> 1) * (int)sizeof(double);
    }
    const bool more = l > kTurn + 2 && l > l_min;      // another column follows
    if (more) levels(l - 1);      // first, by every warp at the same time (see the inside pass)
    // ---- cells(l): outside weight of every listed pair (k,l)
    if (m > 0) {
      const double* const gen0 = w.vecp(kOGen0 + (l & 1));
      const double* const gen1 = w.vecp(kOGen1 + (l & 1));
      const double* const vA = w.vecp(kOA + (l & 1));
      const double* const vB = w.vecp(kOB + (l & 1));
      double* const vB_new = w.vecp(kOB + ((l - 1) & 1));
      double* const pcol = w.vecp(kOP + (l & 1));
      const double* const xA_rd = w.vecp(kOXA + ((l + 1) & 1));      // column l+1
      const double* const x1n_rd = w.vecp(kOX1n + (l + 2) % 3);      // column l+2
      double* const xA_wr = w.vecp(kOXA + (l & 1));
      double* const x1n_wr = w.vecp(kOX1n + l % 3);
      const uint16_t* const cells = w.cells + (l & 1) * ld;
      const int lg = group_lg(m, nt_lg, warp_lg);
      const int Gs = 1 << lg, slots = nt >> lg, sub = tid & (Gs - 1);
      const int upr_l1 = kCst ? (int)c.upr[l + 1] : n - l;
      const bool cul = !kCst || can_unpair(c, l);
      const int sl = c.S[l], cnl = cn[l] << 4, cpl = cp[l];
      int u2cap = upr_l1;                         // also keeps j = l+1+u2 <= n when unconstrained
      if (u2cap > kMaxLoop) u2cap = kMaxLoop;
      if (u2cap > n - l - 1) u2cap = n - l - 1;
      const double extl = qln[l + 1] * invZ;
      for (int slot0 = 0; slot0 < m; slot0 += slots) {
        const int slot = slot0 + (tid >> lg);
        const bool valid = slot < m;
        const int k = valid ? cells[slot] : 1;
        double val = 0., sx = 0.;      // sx: exterior stem factor of the cell (first lane of the group)
        if (valid) {
          const int sk = c.S[k];
          const int type = pair_type(sk, sl);
          const int codekl = cnl | cp[k];          // inner orientation
          int u1cap = kCst ? (int)c.upl[k - 1] : kMaxLoop;
          if (u1cap > kMaxLoop) u1cap = kMaxLoop;
          if (u1cap > k - 2) u1cap = k - 2;                     // i = k-1-u1 >= 1
          double other = 0., accB = 0., acc1n = 0., accML = 0.;
          double cval = 0., ctab = 0.;
          if (sub == 0) {
            sx = fast_stem_ext(c, H, type, k, l);
            other = q1[k - 1] * extl * sx + H.FI[codekl] * (gen0[k] + gen1[k]);
          }
          ADP_TICK(8)
          if (k >= 2 && l < n) {
            const int t2r = rtype(type);
            if (sub == Gs - 1) {
              // the enclosing stack (i,j) = (k-1,l+1) as straight-line code on the group's last lane
              const double v = qb[t4(k - 1, l + 1)] * H.RI[(cn[k - 1] << 4) | cp[l + 1]];
              other += v * H.cfac[0] * H.stack[H.tyr[codekl] * 8 + t2r];
            }
            ADP_NOUNROLL
            for (int q = 1 + sub; q < 9; q += Gs) {       // software-pipelined like the inside pass
              int u1, u2;
              corner_u(q, &u1, &u2);
              double v2 = 0.;
              const double* p2 = zp;
              if (u1 <= u1cap && u2 <= u2cap) {
                const int i = k - 1 - u1, j = l + 1 + u2;
                const int cout = (cn[i] << 4) | cp[j];
                const double v = qb[t4(i, j)] * H.RI[cout];
                if (v != 0.) {
                  p2 = ADP_FC.corner + ((size_t)(q - 1) << 16) + ((size_t)cout << 8) + codekl;
                  v2 = v;
                }
              }
              const double t2 = *p2;
              other += cval * ctab;
              cval = v2;
              ctab = t2;
            }
            ADP_TICK(9)
            // the four 1-D filters fused into one loop like in the inside pass: enclosing bulges (i = k-1, j = l+1+u / j = l+1,
            // i = k-1-u: column vector xA), enclosing 1 x u loops (i = k-2, j = l+1+u, u >= 3) and u x 1 loops (j = l+2,
            // i = k-1-u: column vector x1n).
            {
              int hi13 = u2cap < kMaxLoop - 1 ? u2cap : kMaxLoop - 1, hi31 = u1cap < kMaxLoop - 1 ? u1cap : kMaxLoop - 1;
              if (u1cap < 1) hi13 = -1;
              if (u2cap < 1) hi31 = -1;
              const int hmax = u2cap > u1cap ? u2cap : u1cap;
              const int cni1 = cn[k - 1] << 4, cni2 = k >= 3 ? cn[k - 2] << 4 : 0;
              const double* xa = xA_rd + k - 1;
              const double* xn = x1n_rd + k - 1;
              for (int u = 2 + sub; u <= hmax; u += Gs) {
                const bool in3 = u <= u2cap;                   // j = l+1+u <= n
                const int j = l + 1 + (in3 ? u : 0);
                const int cj = cp[j];
                const double* pb = in3 ? qb + t4(k - 1, j) : zp;       // (k-1,j) is a cell; (k-2,j) is the one before it when k >= 3
                const double b3 = pb[0] * H.RB[cni1 | cj];
                const double n3 = u <= hi13 ? pb[-1] * H.R1[cni2 | cj] : 0.;
                const int ub = u <= u1cap ? u : 0;             // row k-1-u >= 1
                const double b5 = xa[-ub], n5 = xn[-ub];
                accB += H.tapB[u] * ((in3 ? b3 : 0.) + (u <= u1cap ? b5 : 0.));
                acc1n += H.tap1n[u] * (n3 + (u <= hi31 ? n5 : 0.));
              }
            }
            ADP_TICK(10)
            // (k,l) is a stem of a multiloop closed by (i,j), i < k, j > l:  u = k-1-i unpaired bases left of k
            {
              const double* qcol = qm + t4(1, k - 1) - 1;       // Qm(i+1,k-1) = qcol[i+1], valid for i+1 <= k-5
              for (int i = 1 + sub; i <= k - 1; i += Gs) {
                const int u = k - 1 - i;
                double t = 0.;
                if (!kCst || u <= (int)c.upr[i + 1]) t = eMLbS[u] * vA[i];
                if (u >= 5) t += qcol[i + 1] * (vA[i] + vB[i]);
                accML += t;
              }
            }
          }
          // ---- cell total
          other += cval * ctab;
          val = other + au_factor(H, sk, sl) * accB + H.R1[codekl] * H.FI[codekl] * acc1n;
          if (k >= 2 && l < n) val += accML * (H.RM[codekl] * H.FI[codekl] * H.scale2);
        }
        ADP_TICK(11)
        val = ex.group_sum_rt(val, Gs);
        ADP_TICK(12)
        if (valid && sub == 0) {
          const int idx = t4(k, l);
          const int type = pair_type(c.S[k], sl);
          const double qbv = (k > 1 && l < n) ? qb[idx] * H.RI[cnl | cp[k]] : qb[idx];     // back to the plain Qb(k,l)
          pcol[k] = val * qbv;
          accq[k] += qbv * sx * qln[l + 1];
          if (has_motif && is_motif_site(c, k, l)) pm[k] = val * fast_hairpin<kCst>(c, H, G, w, motif_boltz, k, l, type);
          if (site_mode && k == out.site_k && l == out.site_l) *out.site_ph = val * fast_hairpin<kCst>(c, H, G, w, 1., k, l, type);
          const int code = (cn[k] << 4) | cpl;             // outer orientation
          const double pre = val * H.FI[code];             // outside weight, premultiplied
          qb[idx] = pre;
          xA_wr[k] = pre * H.RB[code];
          x1n_wr[k] = pre * H.R1[code];
          if (k <= l - 6) vB_new[k] = (cul ? vB[k] * eb1 : 0.) + pre * H.MLclosing * H.RM[code];      // carrier B of column l-1
        }
        ADP_TICK(13)
      }
    }
    ADP_TICK(0)
    // ---- everything else of the phase
    ADP_TICK(4)
    if (more) list_rows(l - 1);
    if (l <= n && l > kTurn + 1) rest_rows(l);
    if (l <= n - 1 && !site_mode) results(l + 1);
    // exterior suffix Q(l-1,n): position l-1 unpaired, or paired with some l' >= l+3 (already accumulated)
    if (more && !site_mode && tid == nt - 1) qln[l - 1] = ((!kCst || can_unpair(c, l - 1)) ? qln[l] * s1 : 0.) + accq[l - 1];
    ADP_TICK(5)
    if (more) carriers(l - 1);
    ADP_TICK(6)
    ex.sync();
    ADP_TICK(7)
  }

  // the motif's inner pairs inherit the probability of the bound hairpin
  ex.sync();
  if (has_motif && !site_mode && tid == 0) {
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
        for (int z = 0; z < out.n_dens; z++)
          if (out.dens_pos[z] + 1 == a || out.dens_pos[z] + 1 == b) out.dens[z] += pm[i];
      }
    }
  }
  ex.sync();
  ADP_PROF_REPORT("outside")
}

}  // namespace adp
