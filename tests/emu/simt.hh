// tests/emu/simt.hh -- TEST INFRASTRUCTURE ONLY; never part of libaddapt_b200.so.
//
// A CTA in host emulation WITH its threads: every CUDA thread of the CTA is a fiber (ucontext) that runs the kernel body
// through the same execution-policy interface as CtaExec (fold_common.cuh), and a seeded scheduler decides which fiber runs
// next.  A fiber runs until it reaches a collective:
//   * sync()              the CTA barrier: nobody passes before every thread has arrived;
//   * a warp collective   (the shuffles / ballots behind warp_sum, group_sum_rt, warp_bcast, ...): every lane of the warp
//                         deposits its operand, nobody reads before all 32 have arrived, and the butterfly is evaluated in the
//                         order of the __shfl_xor_sync sequence of CtaExec, so the sums round like the GPU's.
// Between two collectives the threads of a CTA may run in ANY order on the GPU; here the order is what the policy picks
// (random thread, ascending, descending, random warp).  A kernel body that is correct under the CUDA model gives the same bits
// for every order; one with a missing barrier (a value read in the phase that writes it), with a collective that not every
// lane reaches, or with a barrier after some threads have left does not, and is reported.  This is the race / divergence check
// that compute-sanitizer would do on a GPU (the tool is closed on the pool these kernels were developed on), for exactly the
// source the sm_100a kernels are compiled from.  What it cannot see: hardware memory-ordering effects inside a phase and anything
// in the __global__ wrappers themselves (those are restated in emu.cc next to each entry point).
#pragma once
#include <ucontext.h>

#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <random>
#include <string>
#include <vector>

namespace simt {

struct Cta {
  enum State { kRun = 0, kAtCtaBarrier = 1, kAtWarpBarrier = 2, kDone = 3 };
  enum Policy { kRandomThread = 0, kAscending = 1, kDescending = 2, kRandomWarp = 3 };
  static constexpr size_t kStackBytes = 512 * 1024;

  int nt = 0, policy = 0;
  std::mt19937_64 rng;
  std::vector<ucontext_t> ctx;
  ucontext_t main_ctx;
  std::vector<char*> stacks;
  std::vector<int> state, where, runnable, warp_wait, warp_live;
  std::vector<uint8_t> parity;        // per thread: which of the two exchange buffers its next collective uses
  std::vector<uint64_t> slot_a, slot_b;   // [warp][buffer][lane]
  std::vector<uint32_t> slot_tag;     // opcode and group size of the collective a lane is in (all lanes must agree)
  int cta_wait = 0, live = 0, sticky_warp = -1;
  long n_cta_barriers = 0, n_collectives = 0, n_switches = 0;
  // Fault injection (tests only): every thread returns from its k-th sync() call at once, as if that barrier were not in the
  // source.  Shows that the schedule comparison does notice a missing barrier in the real kernel bodies.
  long drop_barrier = -1;
  std::vector<long> sync_calls;
  std::string error;
  std::function<void(int)> body;

  ~Cta() { for (char* s : stacks) std::free(s); }

  static void trampoline(unsigned lo, unsigned hi, int tid) {
    Cta* c = reinterpret_cast<Cta*>(((uintptr_t)hi << 32) | (uintptr_t)lo);
    c->body(tid);
    c->finish(tid);
  }

  // Runs body(tid) for tid = 0 .. threads-1 to completion.  false + `error` on a deadlock / divergence.
  bool run(int threads, int policy_, uint64_t seed, std::function<void(int)> body_) {
    nt = threads; policy = policy_; rng.seed(seed); body = std::move(body_);
    if (nt <= 0 || nt % 32) { error = "thread count must be a positive multiple of 32"; return false; }
    const int nw = nt / 32;
    for (char* s : stacks) std::free(s);
    ctx.resize(nt); stacks.assign(nt, nullptr); state.assign(nt, kRun); where.assign(nt, -1); runnable.clear();
    warp_wait.assign(nw, 0); warp_live.assign(nw, 32); parity.assign(nt, 0);
    slot_a.assign((size_t)nw * 64, 0); slot_b.assign((size_t)nw * 64, 0); slot_tag.assign((size_t)nw * 64, 0);
    cta_wait = 0; live = nt; sticky_warp = -1; error.clear();
    sync_calls.assign(nt, 0);
    for (int t = 0; t < nt; t++) {
      stacks[t] = (char*)std::malloc(kStackBytes);
      getcontext(&ctx[t]);
      ctx[t].uc_stack.ss_sp = stacks[t];
      ctx[t].uc_stack.ss_size = kStackBytes;
      ctx[t].uc_link = &main_ctx;
      const uintptr_t self = reinterpret_cast<uintptr_t>(this);
      makecontext(&ctx[t], (void (*)())trampoline, 3, (unsigned)(self & 0xffffffffu), (unsigned)(self >> 32), t);
      make_runnable(t);
    }
    while (live > 0 && error.empty()) {
      if (runnable.empty()) { error = describe_deadlock(); break; }
      const int t = pick();
      n_switches++;
      swapcontext(&main_ctx, &ctx[t]);
    }
    return error.empty();
  }

  // ---- called from inside a fiber -------------------------------------------------------------------------------------
  void cta_barrier(int t) {
    if (sync_calls[t]++ == drop_barrier) return;
    if (live < nt) fail("a thread reached the CTA barrier after another thread had returned");
    block(t, kAtCtaBarrier);
    if (++cta_wait == live) {
      cta_wait = 0;
      n_cta_barriers++;
      for (int u = 0; u < nt; u++)
        if (state[u] == kAtCtaBarrier) make_runnable(u);
    }
    yield(t);
  }

  void warp_barrier(int t) {
    const int w = t >> 5;
    if (warp_live[w] < 32) fail("a lane reached a warp collective after another lane of its warp had returned");
    block(t, kAtWarpBarrier);
    if (++warp_wait[w] == warp_live[w]) {
      warp_wait[w] = 0;
      n_collectives++;
      for (int u = w * 32; u < w * 32 + 32; u++)
        if (state[u] == kAtWarpBarrier) make_runnable(u);
    }
    yield(t);
  }

  // One warp collective: every lane deposits (a, b) and a tag; after all 32 have arrived `va` / `vb` point to the 32 operands.
  // Two buffers alternate, so a lane that runs ahead into its next collective cannot overwrite operands still being read.
  void exchange(int t, uint64_t a, uint64_t b, uint32_t tag, const uint64_t** va, const uint64_t** vb) {
    const int w = t >> 5, l = t & 31;
    const size_t base = ((size_t)w * 2 + parity[t]) * 32;
    slot_a[base + l] = a; slot_b[base + l] = b; slot_tag[base + l] = tag;
    parity[t] ^= 1;
    warp_barrier(t);
    for (int k = 0; k < 32; k++)
      if (slot_tag[base + k] != tag) { fail("the lanes of a warp are in different collectives (or pass different group sizes)"); break; }
    *va = &slot_a[base]; *vb = &slot_b[base];
  }

 private:
  void make_runnable(int t) { state[t] = kRun; where[t] = (int)runnable.size(); runnable.push_back(t); }
  void block(int t, int st) {
    const int k = where[t], last = runnable.back();
    runnable[k] = last; where[last] = k; runnable.pop_back(); where[t] = -1;
    state[t] = st;
  }
  void yield(int t) { swapcontext(&ctx[t], &main_ctx); }
  void finish(int t) {
    block(t, kDone);
    live--; warp_live[t >> 5]--;
    if (cta_wait > 0) fail("a thread returned while others wait at the CTA barrier");
    if (warp_wait[t >> 5] > 0) fail("a lane returned while others of its warp wait in a collective");
  }
  void fail(const char* what) { if (error.empty()) error = what; }
  int pick() {
    const int m = (int)runnable.size();
    if (policy == kRandomThread) return runnable[rng() % m];
    if (policy == kAscending || policy == kDescending) {
      int best = runnable[0];
      for (int t : runnable) if (policy == kAscending ? t < best : t > best) best = t;
      return best;
    }
    // random warp: stay with one warp (lowest lane first) until none of its lanes can run, then draw another
    int best = -1;
    for (int t : runnable) if ((t >> 5) == sticky_warp && (best < 0 || t < best)) best = t;
    if (best >= 0) return best;
    sticky_warp = runnable[rng() % m] >> 5;
    for (int t : runnable) if ((t >> 5) == sticky_warp && (best < 0 || t < best)) best = t;
    return best;
  }
  std::string describe_deadlock() const {
    int at_cta = 0, at_warp = 0;
    for (int t = 0; t < nt; t++) { at_cta += state[t] == kAtCtaBarrier; at_warp += state[t] == kAtWarpBarrier; }
    return "deadlock: " + std::to_string(at_cta) + " threads at the CTA barrier, " + std::to_string(at_warp) +
           " lanes in warp collectives that the rest of their warp never reaches, " + std::to_string(nt - live) + " returned";
  }
};

// The execution policy (same interface as CtaExec<G> / SerialExec in fold_common.cuh).
template <int G>
struct SimtExec {
  static constexpr int kGroup = G;
  static constexpr int kWarp = 32;
  static constexpr bool kSerial = false;
  Cta* cta;
  int t;
  enum Op : uint32_t { kSumD = 1, kMinI, kBallot, kBcast, kSum2, kShflDown, kSync };

  int tid() const { return t; }
  int nthreads() const { return cta->nt; }
  int lane() const { return t & 31; }
  int warp() const { return t >> 5; }
  void sync() const { cta->cta_barrier(t); }

  static uint64_t bits(double v) { uint64_t u; std::memcpy(&u, &v, 8); return u; }
  static double dbl(uint64_t u) { double v; std::memcpy(&v, &u, 8); return v; }

  // v += __shfl_xor_sync(v, o) for o = g/2 .. 1, evaluated for the whole warp like the hardware does
  double butterfly_sum(double v, int g, uint32_t op, const uint64_t* other = nullptr, double* second = nullptr) const {
    const uint64_t *va, *vb;
    cta->exchange(t, bits(v), other ? *other : 0, (op << 8) | (uint32_t)g, &va, &vb);
    double x[32], y[32], nx[32], ny[32];
    for (int k = 0; k < 32; k++) { x[k] = dbl(va[k]); y[k] = dbl(vb[k]); }
    for (int o = g / 2; o > 0; o >>= 1) {
      for (int k = 0; k < 32; k++) { nx[k] = x[k] + x[k ^ o]; ny[k] = y[k] + y[k ^ o]; }
      std::memcpy(x, nx, sizeof x); std::memcpy(y, ny, sizeof y);
    }
    if (second) *second = y[lane()];
    return x[lane()];
  }
  int butterfly_min(int v, int g) const {
    const uint64_t *va, *vb;
    cta->exchange(t, (uint64_t)(int64_t)v, 0, ((uint32_t)kMinI << 8) | (uint32_t)g, &va, &vb);
    int x[32], nx[32];
    for (int k = 0; k < 32; k++) x[k] = (int)(int64_t)va[k];
    for (int o = g / 2; o > 0; o >>= 1) {
      for (int k = 0; k < 32; k++) nx[k] = x[k] < x[k ^ o] ? x[k] : x[k ^ o];
      std::memcpy(x, nx, sizeof x);
    }
    return x[lane()];
  }

  double group_sum(double v) const { return G > 1 ? butterfly_sum(v, G, kSumD) : v; }
  int group_min(int v) const { return G > 1 ? butterfly_min(v, G) : v; }
  double warp_sum(double v) const { return butterfly_sum(v, 32, kSumD); }
  int warp_min(int v) const { return butterfly_min(v, 32); }
  // the run-time group size is uniform over the warp, and every lane executes the (possibly empty) shuffle sequence
  double group_sum_rt(double v, int g) const { return butterfly_sum(v, g, kSumD); }
  int group_min_rt(int v, int g) const { return butterfly_min(v, g); }
  void group_sum2_rt(double& a, double& b, int g) const {
    const uint64_t bb = bits(b);
    double second;
    a = butterfly_sum(a, g, kSum2, &bb, &second);
    b = second;
  }
  int warp_first(bool flag) const {
    const uint64_t *va, *vb;
    cta->exchange(t, flag ? 1 : 0, 0, (uint32_t)kBallot << 8, &va, &vb);
    for (int k = 0; k < 32; k++) if (va[k]) return k;
    return -1;
  }
  int warp_bcast(int v, int src) const {
    const uint64_t *va, *vb;
    cta->exchange(t, (uint64_t)(int64_t)v, 0, (uint32_t)kBcast << 8, &va, &vb);
    return (int)(int64_t)va[src & 31];
  }
  double shfl_down_d(double v, int o) const {
    const uint64_t *va, *vb;
    cta->exchange(t, bits(v), 0, ((uint32_t)kShflDown << 8) | (uint32_t)o, &va, &vb);
    return lane() + o < 32 ? dbl(va[lane() + o]) : v;
  }
  void warp_sync() const {
    const uint64_t *va, *vb;
    cta->exchange(t, 0, 0, (uint32_t)kSync << 8, &va, &vb);
  }
  int atomic_inc(int* ctr) const { return (*ctr)++; }      // fibers are not preempted: a plain increment is atomic
};

}  // namespace simt
