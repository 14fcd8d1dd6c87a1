// Shared pieces of the fold kernels: the execution policy (real CTA vs. single-thread host emulation used
// only by tests/emu), the per-construct context (sequence, hard constraints, motif sites) and the loop
// energy / Boltzmann-weight functions of the Turner-2004 dangles=2 model.
//
// What these functions compute is what libRNA computes inside vrna_pf / vrna_mfe for the reference
// (reference src/scoring.cc:43,58,65); how it is organised (column-parallel recursion, column-major
// triangular matrices in shared memory, lane groups per cell) is specific to this engine.
#pragma once
#include <math.h>
#include <stdint.h>

#include "tables.h"

#if defined(__CUDACC__)
#define ADP_HD __host__ __device__ __forceinline__
#else
#define ADP_HD inline
#endif

namespace adp {

// ---------------------------------------------------------------------------------------------------
// Execution policies.  Kernel bodies are written once against this interface.
//   tid / nthreads : position in the CTA            sync()        : CTA barrier
//   kGroup         : lanes cooperating on one cell  group_sum/min : reduction over those lanes
//   lane / warp    : position in the warp           warp_sum      : reduction over the warp
#if defined(__CUDACC__)
template <int G>
struct CtaExec {
  static constexpr int kGroup = G;
  static constexpr int kWarp = 32;
  static constexpr bool kSerial = false;
  __device__ __forceinline__ int tid() const { return threadIdx.x; }
  __device__ __forceinline__ int nthreads() const { return blockDim.x; }
  __device__ __forceinline__ int lane() const { return threadIdx.x & 31; }
  __device__ __forceinline__ int warp() const { return threadIdx.x >> 5; }
  __device__ __forceinline__ void sync() const { __syncthreads(); }
  __device__ __forceinline__ double group_sum(double v) const {
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
  }
  __device__ __forceinline__ int group_min(int v) const {
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) v = min(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
  }
  __device__ __forceinline__ double warp_sum(double v) const {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
  }
  __device__ __forceinline__ int warp_min(int v) const {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = min(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
  }
  // index of the lowest lane whose flag is set, or -1
  __device__ __forceinline__ int warp_first(bool flag) const {
    unsigned m = __ballot_sync(0xffffffffu, flag);
    return m ? __ffs(m) - 1 : -1;
  }
  __device__ __forceinline__ int warp_bcast(int v, int src) const { return __shfl_sync(0xffffffffu, v, src); }
  // reduction over an aligned group of `g` lanes (g = power of two <= 32, uniform over the warp)
  __device__ __forceinline__ double group_sum_rt(double v, int g) const {
    // g is uniform over the warp: five straight-line butterfly steps, each skipped as a whole when the group is smaller
    if (g > 16) v += __shfl_xor_sync(0xffffffffu, v, 16);
    if (g > 8) v += __shfl_xor_sync(0xffffffffu, v, 8);
    if (g > 4) v += __shfl_xor_sync(0xffffffffu, v, 4);
    if (g > 2) v += __shfl_xor_sync(0xffffffffu, v, 2);
    if (g > 1) v += __shfl_xor_sync(0xffffffffu, v, 1);
    return v;
  }
  __device__ __forceinline__ int group_min_rt(int v, int g) const {
    if (g > 16) v = min(v, __shfl_xor_sync(0xffffffffu, v, 16));
    if (g > 8) v = min(v, __shfl_xor_sync(0xffffffffu, v, 8));
    if (g > 4) v = min(v, __shfl_xor_sync(0xffffffffu, v, 4));
    if (g > 2) v = min(v, __shfl_xor_sync(0xffffffffu, v, 2));
    if (g > 1) v = min(v, __shfl_xor_sync(0xffffffffu, v, 1));
    return v;
  }
  // two independent sums reduced together (the shuffles of one hide the latency of the other)
  __device__ __forceinline__ void group_sum2_rt(double& a, double& b, int g) const {
    if (g > 16) { a += __shfl_xor_sync(0xffffffffu, a, 16); b += __shfl_xor_sync(0xffffffffu, b, 16); }
    if (g > 8) { a += __shfl_xor_sync(0xffffffffu, a, 8); b += __shfl_xor_sync(0xffffffffu, b, 8); }
    if (g > 4) { a += __shfl_xor_sync(0xffffffffu, a, 4); b += __shfl_xor_sync(0xffffffffu, b, 4); }
    if (g > 2) { a += __shfl_xor_sync(0xffffffffu, a, 2); b += __shfl_xor_sync(0xffffffffu, b, 2); }
    if (g > 1) { a += __shfl_xor_sync(0xffffffffu, a, 1); b += __shfl_xor_sync(0xffffffffu, b, 1); }
  }
  __device__ __forceinline__ int atomic_inc(int* ctr) const { return atomicAdd(ctr, 1); }
  __device__ __forceinline__ double shfl_down_d(double v, int o) const { return __shfl_down_sync(0xffffffffu, v, o); }
  __device__ __forceinline__ void warp_sync() const { __syncwarp(); }
};
#endif

// One virtual thread: every strided loop covers its whole range, barriers and reductions are no-ops.
// Used by tests/emu to check the recursions on a machine without a GPU; never part of the product library.
struct SerialExec {
  static constexpr int kGroup = 1;
  static constexpr int kWarp = 1;
  static constexpr bool kSerial = true;
  int tid() const { return 0; }
  int nthreads() const { return 1; }
  int lane() const { return 0; }
  int warp() const { return 0; }
  void sync() const {}
  double group_sum(double v) const { return v; }
  int group_min(int v) const { return v; }
  double warp_sum(double v) const { return v; }
  int warp_min(int v) const { return v; }
  int warp_first(bool flag) const { return flag ? 0 : -1; }
  int warp_bcast(int v, int) const { return v; }
  double group_sum_rt(double v, int) const { return v; }
  void group_sum2_rt(double&, double&, int) const {}
  int group_min_rt(int v, int) const { return v; }
  int atomic_inc(int* ctr) const { return (*ctr)++; }
  double shfl_down_d(double v, int) const { return v; }
  void warp_sync() const {}
};

// ---------------------------------------------------------------------------------------------------
// Hard constraint, preprocessed on the host from a dot-bracket string
// (≙ vrna_constraints_add(..., DB_DEFAULT | DB_ENFORCE_BP), reference src/scoring.cc:61-62):
//   'x'      position may not pair                     -> flag bit 0
//   '(' ')'  the pair MUST form: both ends lose the right to stay unpaired (flag bit 1), may pair only with
//            each other (partner), and no pair may cross it.  `enc` is the id of the innermost enforced pair
//            strictly enclosing a position (0 = none); two free positions may pair iff their enc ids agree.
struct ConstraintDev {
  uint16_t partner[kMaxLen + 2];
  uint16_t enc[kMaxLen + 2];
  uint8_t flag[kMaxLen + 2];
  uint16_t upr[kMaxLen + 3];   // run lengths of positions that may stay unpaired (see SeqCtx::upr / upl); a function of the
  uint16_t upl[kMaxLen + 2];   // constraint alone, so build_constraint computes them once instead of every fill
};

constexpr int kMaxVariants = 8;

// One variant = one fill of the same sequence under a different condition.
struct VariantDev {
  int32_t cst_id;      // index into csts, -1 = unconstrained
  int32_t use_motif;   // 1 = holo (ligand motif soft constraint active)
  int32_t outside;     // 1 = also base-pair probabilities
  int32_t pad;
};

// Per-construct working set (pointers into shared memory on the device).  Indices are 1-based like the
// recursions; S[0] = S[n+1] = 0.
struct SeqCtx {
  int n;
  bool has_cst;
  uint8_t* S;          // nucleotide codes
  uint8_t* flag;       // constraint flags (see ConstraintDev)
  uint16_t* partner;
  uint16_t* enc;
  uint16_t* upr;       // upr[p]: number of consecutive positions p, p+1, ... that may stay unpaired; upr[n+1] = 0
  uint16_t* upl;       // upl[p]: number of consecutive positions p, p-1, ... that may stay unpaired; upl[0] = 0
  uint8_t* motif_at;   // 1 where an occurrence of the ligand motif starts
  int motif_len;
};

ADP_HD int encode_nt(uint8_t c) {
  switch (c) {
    case 'A': case 'a': return 1;
    case 'C': case 'c': return 2;
    case 'G': case 'g': return 3;
    case 'U': case 'u': case 'T': case 't': return 4;
    default: return 0;
  }
}

// pair type of nucleotide codes (a,b): CG1 GC2 GU3 UG4 AU5 UA6, 0 = no pair.  3-bit LUT over (a-1)*4+(b-1).
ADP_HD int pair_type(int a, int b) {
  const unsigned long long lut = (5ull << 9) | (1ull << 18) | (2ull << 27) | (3ull << 33) | (6ull << 36) | (4ull << 42);
  if (a == 0 || b == 0) return 0;
  return (int)((lut >> (3 * (((a - 1) << 2) | (b - 1)))) & 7ull);
}
ADP_HD int rtype(int t) { return (t == 0 || t == 7) ? t : (((t - 1) ^ 1) + 1); }

// pair type of (i,j), i < j, or 0 if the pair is too short, non-canonical or forbidden by the hard constraint
ADP_HD int allowed_type(const SeqCtx& c, int i, int j) {
  if (j - i <= kTurn) return 0;
  int t = pair_type(c.S[i], c.S[j]);
  if (t == 0 || !c.has_cst) return t;
  if ((c.flag[i] | c.flag[j]) & 1) return 0;
  int pi = c.partner[i], pj = c.partner[j];
  if (pi | pj) return pi == j ? t : 0;
  return c.enc[i] == c.enc[j] ? t : 0;
}

ADP_HD bool can_unpair(const SeqCtx& c, int p) { return !c.has_cst || !(c.flag[p] & 2); }

// Fills the context from the ASCII sequence row and the optional constraint / motif.  Collective: every
// thread of the CTA calls it; ends with a barrier.
template <class E>
ADP_HD void load_context(const E& ex, SeqCtx& c, const uint8_t* seq, const ConstraintDev* cst, const MotifDev* motif) {
  const int n = c.n;
  for (int p = ex.tid(); p <= n + 2; p += ex.nthreads()) {
    const bool in = p >= 1 && p <= n;
    if (p <= n + 1) {
      c.S[p] = in ? (uint8_t)encode_nt(seq[p - 1]) : 0;
      if (cst) {
        c.flag[p] = in ? cst->flag[p] : 0;
        c.partner[p] = in ? cst->partner[p] : 0;
        c.enc[p] = in ? cst->enc[p] : 0;
      } else {
        c.flag[p] = 0; c.partner[p] = 0; c.enc[p] = 0;
      }
      c.motif_at[p] = 0;
      c.upl[p] = in ? (cst ? cst->upl[p] : (uint16_t)p) : 0;
    }
    c.upr[p] = in ? (cst ? cst->upr[p] : (uint16_t)(n - p + 1)) : 0;     // run lengths: see ConstraintDev
  }
  ex.sync();
  if (motif && motif->len > 0) {
    const int m = motif->len;
    for (int p = 1 + ex.tid(); p + m - 1 <= n; p += ex.nthreads()) {
      bool hit = true;
      for (int q = 0; q < m; q++)
        if (c.S[p + q] != motif->code[q]) { hit = false; break; }
      c.motif_at[p] = hit ? 1 : 0;
    }
  }
  ex.sync();
}

// ---------------------------------------------------------------------------------------------------
// carving a raw (shared-memory) buffer into the per-construct arrays
struct Carver {
  char* p;
  ADP_HD explicit Carver(void* base) : p((char*)base) {}
  // 16-byte aligned array of `count` T.  Only pointer arithmetic on `p` (no integer -> pointer cast), so that the
  // compiler keeps the address space of the buffer (shared memory loads instead of generic ones).
  template <class T>
  ADP_HD T* take(size_t count) {
    p += (16 - ((size_t)p & 15)) & 15;
    T* r = reinterpret_cast<T*>(p);
    p += count * sizeof(T);
    return r;
  }
};

// Same layout, but every array is the (aligned) base plus an offset that is tracked as an integer: with many arrays live
// (the fast PF path) the compiler then rebuilds addresses from one base register instead of carrying a pointer per array
// (measured: PF fast +10 %; the MFE kernel, with fewer arrays, is faster with Carver).
struct OffsetCarver {
  char* p;
  size_t off;
  ADP_HD explicit OffsetCarver(void* base) : p((char*)base), off((16 - ((size_t)base & 15)) & 15) {}
  template <class T>
  ADP_HD T* take(size_t count) {
    off = (off + 15) & ~(size_t)15;
    T* r = reinterpret_cast<T*>(p + off);
    off += count * sizeof(T);
    return r;
  }
};

// bytes of the per-construct context arrays
ADP_HD size_t ctx_bytes(int n) {
  return 16 * 8 + (size_t)(n + 4) * (1 + 1 + 1) + (size_t)(n + 4) * 2 * 4;
}

template <class CV>
ADP_HD void carve_ctx(CV& cv, SeqCtx& c, int n) {
  c.n = n;
  c.S = cv.template take<uint8_t>(n + 4);
  c.flag = cv.template take<uint8_t>(n + 4);
  c.motif_at = cv.template take<uint8_t>(n + 4);
  c.partner = cv.template take<uint16_t>(n + 4);
  c.enc = cv.template take<uint16_t>(n + 4);
  c.upr = cv.template take<uint16_t>(n + 4);
  c.upl = cv.template take<uint16_t>(n + 4);
}

ADP_HD bool is_motif_site(const SeqCtx& c, int i, int j) { return c.motif_len > 0 && (j - i + 1) == c.motif_len && c.motif_at[i]; }

// column-major triangular index of cell (i,j), 1 <= i <= j: columns are contiguous in i
ADP_HD int tri(int i, int j) { return ((j * (j - 1)) >> 1) + (i - 1); }

// 2-bit packed code of S[i..i+len-1]; returns false when a non-ACGU letter is inside
ADP_HD bool pack_window(const uint8_t* S, int i, int len, uint32_t* code) {
  uint32_t v = 0;
  for (int p = 0; p < len; p++) {
    int s = S[i + p];
    if (!s) return false;
    v |= (uint32_t)(s - 1) << (2 * p);
  }
  *code = v;
  return true;
}

// ---------------------------------------------------------------------------------------------------
// Integer loop energies (dcal/mol) -- MFE

ADP_HD int hairpin_e(const IntTables& P, const uint8_t* S, int i, int j, int type) {
  const int u = j - i - 1;
  const int e = P.hairpin[u];
  if (u == 3 || u == 4 || u == 6) {
    uint32_t code;
    if (pack_window(S, i, u + 2, &code)) {
      const uint32_t* codes = u == 4 ? P.tetra_code : u == 6 ? P.hex_code : P.tri_code;
      const int32_t* es = u == 4 ? P.tetra_e : u == 6 ? P.hex_e : P.tri_e;
      const int cnt = u == 4 ? P.n_tetra : u == 6 ? P.n_hex : P.n_tri;
      for (int k = 0; k < cnt; k++)
        if (codes[k] == code) return es[k];
    }
  }
  if (u == 3) return e + (type > 2 ? P.TerminalAU : 0);
  return e + P.mmH[type][S[i + 1]][S[j - 1]];
}

ADP_HD int interior_e(const IntTables& P, int u1, int u2, int type, int type2, int si1, int sj1, int sp1, int sq1) {
  const int big = u1 > u2 ? u1 : u2, small = u1 > u2 ? u2 : u1;
  if (big == 0) return P.stack[type][type2];
  if (small == 0) {
    const int e = P.bulge[big];
    if (big == 1) return e + P.stack[type][type2];
    return e + (type > 2 ? P.TerminalAU : 0) + (type2 > 2 ? P.TerminalAU : 0);
  }
  if (small == 1) {
    if (big == 1) return P.int11[type][type2][si1][sj1];
    if (big == 2) return u1 == 1 ? P.int21[type][type2][si1][sq1][sj1] : P.int21[type2][type][sq1][si1][sp1];
    return P.interior[big + 1] + P.ninio[big - 1] + P.mm1nI[type][si1][sj1] + P.mm1nI[type2][sq1][sp1];
  }
  if (small == 2 && big == 2) return P.int22[type][type2][si1][sp1][sq1][sj1];
  if (small == 2 && big == 3) return P.interior[5] + P.ninio[1] + P.mm23I[type][si1][sj1] + P.mm23I[type2][sq1][sp1];
  return P.interior[u1 + u2] + P.ninio[big - small] + P.mmI[type][si1][sj1] + P.mmI[type2][sq1][sp1];
}

// stem (i,j) of pair type `type` inside a multiloop (ext = false) or the exterior loop (ext = true);
// have5 / have3 say whether the 5' / 3' neighbour exists (dangles = 2: both dangle whenever present)
ADP_HD int stem_e(const IntTables& P, int type, bool have5, int s5, bool have3, int s3, bool ext) {
  int e = 0;
  if (have5 && have3) e = ext ? P.mmExt[type][s5][s3] : P.mmM[type][s5][s3];
  else if (have5) e = P.dangle5[type][s5];
  else if (have3) e = P.dangle3[type][s3];
  if (type > 2) e += P.TerminalAU;
  if (!ext) e += P.ML_intern;
  return e;
}

// ---------------------------------------------------------------------------------------------------
// Boltzmann weights -- partition function.  `P` may be the full table set in global memory or a staged
// copy; int11 / int21 / int22 are always read from `G` (global, L2-resident).

// The subset of PfTables small enough to stage in shared memory once per CTA.
struct PfSmall {
  double stack[8][8];
  double hairpin[31], bulge[31], interior[31], ninio[31];
  double mmH[8][5][5], mmI[8][5][5], mm1nI[8][5][5], mm23I[8][5][5], mmM[8][5][5], mmExt[8][5][5];
  double dangle5[8][5], dangle3[8][5];
  double MLclosing, MLintern, TermAU, scale1, scale2, eMLb1;
};

template <class E>
ADP_HD void stage_small(const E& ex, PfSmall* dst, const PfTables* T) {
  // the first members of PfTables are laid out differently (hairpin[kMaxLen+1]); copy field by field
  double* d = &dst->stack[0][0];
  const double* s = &T->stack[0][0];
  for (int k = ex.tid(); k < 64; k += ex.nthreads()) d[k] = s[k];
  for (int k = ex.tid(); k < 31; k += ex.nthreads()) {
    dst->hairpin[k] = T->hairpin[k]; dst->bulge[k] = T->bulge[k]; dst->interior[k] = T->interior[k]; dst->ninio[k] = T->ninio[k];
  }
  for (int k = ex.tid(); k < 200; k += ex.nthreads()) {
    (&dst->mmH[0][0][0])[k] = (&T->mmH[0][0][0])[k];
    (&dst->mmI[0][0][0])[k] = (&T->mmI[0][0][0])[k];
    (&dst->mm1nI[0][0][0])[k] = (&T->mm1nI[0][0][0])[k];
    (&dst->mm23I[0][0][0])[k] = (&T->mm23I[0][0][0])[k];
    (&dst->mmM[0][0][0])[k] = (&T->mmM[0][0][0])[k];
    (&dst->mmExt[0][0][0])[k] = (&T->mmExt[0][0][0])[k];
  }
  for (int k = ex.tid(); k < 40; k += ex.nthreads()) {
    (&dst->dangle5[0][0])[k] = (&T->dangle5[0][0])[k];
    (&dst->dangle3[0][0])[k] = (&T->dangle3[0][0])[k];
  }
  if (ex.tid() == 0) {
    dst->MLclosing = T->MLclosing; dst->MLintern = T->MLintern; dst->TermAU = T->TermAU;
    dst->scale1 = T->scale[1]; dst->scale2 = T->scale[2]; dst->eMLb1 = T->eMLb[1];
  }
}

// unscaled hairpin weight (the caller multiplies by scale[u+2] and by the motif factor)
template <class Tab>
ADP_HD double hairpin_x(const Tab& P, const PfTables* G, const uint8_t* S, int i, int j, int type) {
  const int u = j - i - 1;
  const double q = u <= 30 ? P.hairpin[u] : G->hairpin[u];
  if (u == 3 || u == 4 || u == 6) {
    uint32_t code;
    if (pack_window(S, i, u + 2, &code)) {
      const uint32_t* codes = u == 4 ? G->tetra_code : u == 6 ? G->hex_code : G->tri_code;
      const double* xs = u == 4 ? G->tetra_x : u == 6 ? G->hex_x : G->tri_x;
      const int cnt = u == 4 ? G->n_tetra : u == 6 ? G->n_hex : G->n_tri;
      for (int k = 0; k < cnt; k++)
        if (codes[k] == code) return xs[k];
    }
  }
  if (u == 3) return type > 2 ? q * P.TermAU : q;
  return q * P.mmH[type][S[i + 1]][S[j - 1]];
}

template <class Tab>
ADP_HD double interior_x(const Tab& P, const PfTables* G, int u1, int u2, int type, int type2, int si1, int sj1, int sp1, int sq1) {
  const int big = u1 > u2 ? u1 : u2, small = u1 > u2 ? u2 : u1;
  if (big == 0) return P.stack[type][type2];
  if (small == 0) {
    double z = P.bulge[big];
    if (big == 1) return z * P.stack[type][type2];
    if (type > 2) z *= P.TermAU;
    if (type2 > 2) z *= P.TermAU;
    return z;
  }
  if (small == 1) {
    if (big == 1) return G->int11[type][type2][si1][sj1];
    if (big == 2) return u1 == 1 ? G->int21[type][type2][si1][sq1][sj1] : G->int21[type2][type][sq1][si1][sp1];
    return P.interior[big + 1] * P.mm1nI[type][si1][sj1] * P.mm1nI[type2][sq1][sp1] * P.ninio[big - 1];
  }
  if (small == 2 && big == 2) return G->int22[type][type2][si1][sp1][sq1][sj1];
  if (small == 2 && big == 3) return P.interior[5] * P.mm23I[type][si1][sj1] * P.mm23I[type2][sq1][sp1] * P.ninio[1];
  return P.interior[u1 + u2] * P.mmI[type][si1][sj1] * P.mmI[type2][sq1][sp1] * P.ninio[big - small];
}

template <class Tab>
ADP_HD double stem_x(const Tab& P, int type, bool have5, int s5, bool have3, int s3, bool ext) {
  double e = 1.;
  if (have5 && have3) e = ext ? P.mmExt[type][s5][s3] : P.mmM[type][s5][s3];
  else if (have5) e = P.dangle5[type][s5];
  else if (have3) e = P.dangle3[type][s3];
  if (type > 2) e *= P.TermAU;
  if (!ext) e *= P.MLintern;
  return e;
}

}  // namespace adp
