#!/usr/bin/env python3
"""bench.py -- BASELINE.json's metric: sgRNA fold-evals/sec (apo+holo pf+bppm, 100 nt).

One "step" = one pass of the hot path over one batch of 65,536 synthetic sgRNA-templated 100-nt designs
(BASELINE.json configs[2]); one fold-eval = apo AND holo, each McCaskill inside + outside (full bppm on the
device).  Under torchrun every rank folds its own 65,536 designs (weak scaling, no data-path collective).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload fold_eval|mfe|pf|score|mc]

`value`  : device-resident inputs/outputs, CUDA-event time of the K steps, max over ranks.
`e2e`    : the same metric through the C-ABI call with HOST (pinned) buffers: H2D of the sequences and D2H of
           the two ensemble energies + the queried pair probabilities inside the timed region, every step.
`roofline`: this path is bound by SM FP64 issue (+ shared memory), not HBM or tensor cores (BASELINE.json
           north_star), so `achieved` = algorithmic recurrence terms (1 term = 1 FMA) / kernel time
           and `peak` = the FP64 FMA rate measured on this GPU by a register-resident DFMA loop.
`cpu_baseline` / `--impl reference`: the reference's CPU path for this metric.  Its arithmetic lives in
           ViennaRNA libRNA, which is not in this image, so the in-repo CPU restatement (oracle/, "port") is timed
           on all host cores on a bounded sample of the same workload.
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "sgRNA fold-evals/sec (apo+holo pf+bppm, 100 nt)"
N_NT = 100
BATCH = 65536
RHF6 = "guuuuagagcuagaaauagcaaguuaaaauaaggcuaguccCuUUUCGCCgauaccagccgaaaggcccuuggcagcGACggcaccgagucggugcuuuuuu"
RHF6_ACTIVE = "(............................)xx..xxxxx..xxxxxx(...............................).(.............)......"
THEO = ("GAUACCAGCCGAAAGGCCCUUGGCAGC", "(...((.(((....)))....))...)", 0.32)
# pairs the reference's own rhf(6) test queries through base_pair_prob (tests/test_scoring.cc:201-222), those < 100
QUERY_PAIRS = [(0, 29), (1, 28), (2, 27), (3, 26), (4, 25), (5, 24), (6, 23), (8, 19), (9, 18), (10, 17), (11, 16), (81, 95), (82, 94),
               (83, 93), (84, 92), (85, 91), (86, 90), (33, 73), (34, 72), (35, 71), (36, 70), (37, 69), (40, 65), (41, 64), (42, 63),
               (43, 62), (44, 61), (45, 60), (47, 58), (48, 57), (30, 43), (31, 42), (32, 41), (33, 40), (46, 80), (47, 79), (49, 77),
               (50, 76), (54, 72), (55, 71), (57, 66), (58, 65), (59, 64)]


def templated_designs(B, seed, n=N_NT):
    """rhf(6) trimmed to n nt; the 10 freely mutable positions resampled, the slaved partner complemented."""
    rng = np.random.Generator(np.random.PCG64(seed))
    base = np.frombuffer(RHF6[:n].upper().encode(), dtype=np.uint8)
    a = np.tile(base, (B, 1))
    nts = np.frombuffer(b"ACGU", dtype=np.uint8)
    free = [41, 43, 44, 45, 46, 47, 48, 49, 77, 78]
    draw = rng.integers(0, 4, (B, len(free)))
    for c, p in enumerate(free):
        a[:, p] = nts[draw[:, c]]
    a[:, 79] = np.frombuffer(b"UGCA", dtype=np.uint8)[draw[:, free.index(47)]]
    return np.ascontiguousarray(a)


def random_seqs(B, seed, n=N_NT):
    rng = np.random.Generator(np.random.PCG64(seed))
    return np.ascontiguousarray(np.frombuffer(b"ACGU", dtype=np.uint8)[rng.integers(0, 4, (B, n))])


def count_terms(seq_row):
    """Algorithmic recurrence terms of SURVEY.md 8(d) for one sequence: (W_in, W_out)."""
    code = {65: 1, 67: 2, 71: 3, 85: 4}
    S = np.array([code.get(int(c), 0) for c in seq_row])
    n = len(S)
    pair = np.zeros((5, 5), dtype=bool)
    for a, b in ((1, 4), (4, 1), (2, 3), (3, 2), (3, 4), (4, 3)):
        pair[a, b] = True
    I, J = np.meshgrid(np.arange(n), np.arange(n), indexing="ij")
    far = (J - I) >= 4
    P = pair[S[:, None], S[None, :]] & far
    d = (J - I)[far]
    w_in = 3.0 * d.sum() + (1 + (J - I - 2))[P].sum()
    n_il = 0
    for u1 in range(31):
        for u2 in range(31 - u1):
            a, b = 1 + u1, 1 + u2                       # k = i + a, l = j - b
            if a + b + 4 > n:
                continue
            n_il += np.count_nonzero(P[:n - a, b:] & P[a:, :n - b])
    w_in += n_il
    w_out = float(n_il)
    for l in range(5, n):                                # 1-based l = 5 .. n-1
        w_out += max(0, l - 4 - 1) * (n - l - 1)
    w_out += (I[P] + 1 - 2).clip(min=0).sum()
    return float(w_in), float(w_out)


class ClockSampler(threading.Thread):
    """samples SM clock and throttle reasons of one GPU while the timed region runs"""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.samples, self.reasons, self.max_mhz = index, False, [], set(), None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if not self.nv:
            return
        nv = self.nv
        names = {getattr(nv, k): k for k in dir(nv) if k.startswith("nvmlClocksThrottleReason") and isinstance(getattr(nv, k), int)}
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if bit and (r & bit) and name not in ("nvmlClocksThrottleReasonNone", "nvmlClocksThrottleReasonAll"):
                        self.reasons.add(name.replace("nvmlClocksThrottleReason", ""))
            except Exception:
                pass
            time.sleep(0.1)

    def summary(self):
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(x for x in self.reasons if x not in ("GpuIdle", "ApplicationsClocksSetting"))}


def cpu_port_rate(sample_seqs, threads=0, want_seconds=12.0):
    """the reference's CPU path for this metric: the oracle port, one sequence per thread on all host cores"""
    from oracle.oracle import Oracle
    orc = Oracle()
    cores = orc.num_threads() if threads <= 0 else threads
    rows = [bytes(r).decode() for r in sample_seqs]
    probe = rows[:4 * cores]
    t0 = time.perf_counter()
    orc.pf_batch(probe, bppm=True, threads=cores)
    orc.pf_batch(probe, motif=THEO, bppm=True, threads=cores)
    rate = len(probe) / (time.perf_counter() - t0)
    m = int(max(len(probe), min(len(rows), rate * want_seconds)))
    m -= m % cores
    t0 = time.perf_counter()
    orc.pf_batch(rows[:m], bppm=True, threads=cores)
    orc.pf_batch(rows[:m], motif=THEO, bppm=True, threads=cores)
    dt = time.perf_counter() - t0
    return m / dt, cores, m, dt


def run_reference(args, rank):
    if rank != 0:
        return
    seqs = templated_designs(4096, 20260706)
    from oracle.oracle import Oracle
    orc = Oracle()
    cores = orc.num_threads()
    rows = [bytes(r).decode() for r in seqs]
    t0 = time.perf_counter()
    orc.pf_batch(rows[:2 * cores], bppm=True, threads=cores)
    orc.pf_batch(rows[:2 * cores], motif=THEO, bppm=True, threads=cores)
    rate = 2 * cores / (time.perf_counter() - t0)
    budget = 120.0 / max(1, args.steps + args.warmup)            # whole run <= ~2 minutes
    m = int(max(cores, min(len(rows), rate * min(budget, float(os.environ.get("ADDAPT_BENCH_REF_SECONDS", "15"))))))
    m -= m % cores
    times = []
    for s in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        orc.pf_batch(rows[:m], bppm=True, threads=cores)
        orc.pf_batch(rows[:m], motif=THEO, bppm=True, threads=cores)
        if s >= args.warmup:
            times.append(time.perf_counter() - t0)
    total = sum(times)
    value = m * args.steps / total
    sample = "%d of the 65,536 templated 100-nt designs per step, apo+holo pf+bppm, oracle port (libRNA absent), %d threads" % (m, cores)
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": "fold-evals/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "configs[2]: apo+holo McCaskill pf+bppm of sgRNA-templated 100-nt designs (bounded sample)", "n_nt": N_NT,
                   "designs_per_step": m},
        "cpu_baseline": {"value": value, "unit": "fold-evals/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "fold-evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="fold_eval", choices=["fold_eval", "mfe", "pf", "score", "mc", "sweep"])
    ap.add_argument("--batch", type=int, default=BATCH)
    ap.add_argument("--n", type=int, default=N_NT)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--mc-steps", type=int, default=100)
    ap.add_argument("--no-mc-leg", action="store_true", help="skip the second timed leg (Monte-Carlo design run + top-K gather)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    if args.impl == "reference":
        run_reference(args, rank)
        return

    # stdout carries ONE JSON line, so NCCL's log goes to stderr -- it is NOT switched off: the communicator lines
    # ("... rank R nranks N ...") are the record that N ranks really formed one communicator.  A caller's NCCL_DEBUG /
    # NCCL_DEBUG_FILE are respected; the default is INFO restricted to the INIT subsystem.  Set before torch is imported.
    from addapt_b200.sharding import nccl_log_begin, nccl_log_end
    nccl_log = nccl_log_begin(world)
    import torch
    import torch.distributed as dist
    import addapt_b200
    from addapt_b200.capi import Problem

    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    eng = addapt_b200.Engine(local)          # raises without the CUDA library / device: no fallback
    # one explicit stream for torch (events, L2 flush) AND the engine's kernels: torch.cuda.Event only sees the stream it is
    # recorded on, and the legacy default stream (handle 0) would make the engine fall back to its own stream
    stream = torch.cuda.Stream(device=local)
    torch.cuda.set_stream(stream)
    assert stream.cuda_stream != 0
    eng.set_stream(stream.cuda_stream)
    dev = torch.device("cuda", local)
    B, n = args.batch, args.n
    wl = args.workload

    # ---- synthetic inputs (rank-specific seed: every rank folds its own designs) -------------------------
    if wl in ("mfe", "pf"):
        host_seqs = random_seqs(B, 20260705 + rank, n)
    else:
        host_seqs = templated_designs(B, 20260706 + rank, n) if n <= 102 else random_seqs(B, 20260707 + rank, n)
    pairs = [(i, j) for i, j in QUERY_PAIRS if j < n]
    nq = len(pairs)
    d_seqs = torch.from_numpy(host_seqs).to(dev)
    d_qi = torch.tensor([p[0] for p in pairs], dtype=torch.int32, device=dev)
    d_qj = torch.tensor([p[1] for p in pairs], dtype=torch.int32, device=dev)
    d_dg = torch.zeros(B, 2, dtype=torch.float64, device=dev)
    d_pr = torch.zeros(B, 2, max(nq, 1), dtype=torch.float64, device=dev)
    d_e = torch.zeros(B, dtype=torch.int32, device=dev)
    d_st = torch.zeros(B, n, dtype=torch.uint8, device=dev)
    h_seqs = torch.from_numpy(host_seqs).pin_memory()
    h_qi, h_qj = d_qi.cpu().pin_memory(), d_qj.cpu().pin_memory()
    h_dg = torch.zeros(B, 2, dtype=torch.float64).pin_memory()
    h_pr = torch.zeros(B, 2, max(nq, 1), dtype=torch.float64).pin_memory()
    h_e = torch.zeros(B, dtype=torch.int32).pin_memory()
    h_st = torch.zeros(B, n, dtype=torch.uint8).pin_memory()
    problem = Problem(RHF6[:min(n, 102)], {"active": RHF6_ACTIVE[:min(n, 102)]},
                      [("active", "apo", False, 1.0), ("active", "holo", True, 1.0)], motif=THEO) if wl in ("score", "mc") else None
    mc_traj = min(B, 8192)
    d_best = torch.zeros(mc_traj, dtype=torch.float64, device=dev)
    d_bseq = torch.zeros(mc_traj, min(n, 102), dtype=torch.uint8, device=dev)
    h_best = torch.zeros(mc_traj, dtype=torch.float64).pin_memory()
    h_bseq = torch.zeros(mc_traj, min(n, 102), dtype=torch.uint8).pin_memory()

    def step_device():
        if wl == "fold_eval" or wl == "sweep":
            eng.fold_eval_raw(B, n, d_seqs, THEO, nq, d_qi, d_qj, d_dg, d_pr)
        elif wl == "mfe":
            eng.mfe_raw(B, n, d_seqs, d_e, d_st)
        elif wl == "pf":
            eng.pf_raw(B, n, d_seqs, d_dg)
        elif wl == "score":
            eng.score_raw(problem, B, d_seqs, d_dg)
        elif wl == "mc":
            eng.mc_run_raw(problem, mc_traj, 2016, rank * mc_traj, args.mc_steps, ("annealing", 5.0, 0.0, 300), best_seq=d_bseq, best_score=d_best)

    def step_host():
        if wl == "fold_eval" or wl == "sweep":
            eng.fold_eval_raw(B, n, h_seqs, THEO, nq, h_qi, h_qj, h_dg, h_pr)
            return B * n + 2 * nq * 4, B * 2 * 8 * (1 + nq)
        if wl == "mfe":
            eng.mfe_raw(B, n, h_seqs, h_e, h_st)
            return B * n, B * 4 + B * n
        if wl == "pf":
            eng.pf_raw(B, n, h_seqs, h_dg)
            return B * n, B * 8
        if wl == "score":
            eng.score_raw(problem, B, h_seqs, h_dg)
            return B * n, B * 8
        eng.mc_run_raw(problem, mc_traj, 2016, rank * mc_traj, args.mc_steps, ("annealing", 5.0, 0.0, 300), best_seq=h_bseq, best_score=h_best)
        return n, mc_traj * (8 + n)

    units_per_step = {"fold_eval": B, "sweep": B, "mfe": B, "pf": B, "score": B, "mc": mc_traj * args.mc_steps}[wl]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)      # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(step_fn, K, is_async):
        eng.set_async(1 if is_async else 0)
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
        io = (0, 0)
        barrier()
        for a, b in ev:
            flush.fill_(1)                                            # evict L2 between timed iterations
            a.record(stream)
            r = step_fn()
            io = r if r else io
            b.record(stream)
        barrier()
        ms = sum(a.elapsed_time(b) for a, b in ev)
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        eng.set_async(0)
        return ms, io

    for _ in range(args.warmup):
        step_device()
    torch.cuda.synchronize()
    for _ in range(1):
        step_host()
    sampler = ClockSampler(local)
    sampler.start()
    launches0 = eng.launch_count()
    ms_dev, _ = timed(step_device, args.steps, True)
    launches = eng.launch_count() - launches0
    ms_e2e, io = timed(step_host, args.steps, False)
    sampler.stop_flag = True
    sampler.join(timeout=2)

    total_units = units_per_step * args.steps * world
    value = total_units / (ms_dev * 1e-3)
    e2e_value = total_units / (ms_e2e * 1e-3)
    unit = {"fold_eval": "fold-evals/s", "sweep": "fold-evals/s", "mfe": "folds/s", "pf": "fills/s", "score": "scores/s", "mc": "MC steps/s"}[wl]

    # ---- roofline of the dominant kernel (pf_fast_kernel: one launch per step) --------------------------
    roofline = cpu = None
    if rank == 0:
        sample = host_seqs[:: max(1, B // 48)][:48]
        w = np.array([count_terms(r) for r in sample])
        w_in, w_out = float(w[:, 0].mean()), float(w[:, 1].mean())
        terms_per_unit = {"fold_eval": 2 * (w_in + w_out), "sweep": 2 * (w_in + w_out), "mfe": w_in, "pf": w_in, "score": 4 * w_in, "mc": 0.75 * 4 * w_in}[wl]
        peak = eng.dfma_rate()                                        # GFMA/s measured on this GPU, this run
        kernel_ms = ms_dev / args.steps                               # one kernel launch per step; events bracket only it
        achieved = terms_per_unit * units_per_step / (kernel_ms * 1e-3) / 1e9
        roofline = {"bound": "sm_fp64_issue", "achieved": achieved, "peak": peak, "unit": "GFMA/s", "frac": achieved / peak,
                    "traffic": None, "traffic_note": "dram bytes of the kernel are ~1e-6 of HBM peak (profiles/: 0.7-0.8 MB read, 0 written per 4,096-design launch); the matrices never leave shared memory",
                    "terms_per_unit": terms_per_unit, "w_in": w_in, "w_out": w_out,
                    "peak_source": "adp_measure_dfma_rate: register-resident DFMA loop on this GPU (MEASURED_PEAKS.json holds only HBM / bf16 peaks, which do not bound this path)",
                    "hbm_frac_of_measured": None}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
            algo_bytes = {"fold_eval": n + 16 * (1 + nq)}.get(wl, n + 8) * units_per_step
            roofline["hbm_frac_of_measured"] = algo_bytes / (kernel_ms * 1e-3) / 1e9 / peaks["hbm_gbs"]
        except Exception:
            pass
        cpu = None
        if world == 1 and not args.no_cpu_baseline and wl == "fold_eval":
            v, cores, m, dt = cpu_port_rate(host_seqs[:4096])
            cpu = {"value": v, "unit": "fold-evals/s", "cores": cores, "kind": "port",
                   "sample": "first %d of this step's designs, apo+holo pf+bppm, %.1f s of CPU work on %d threads; oracle port (ViennaRNA libRNA is not in the image)" % (m, dt, cores)}
    # ---- second timed leg (BASELINE.json configs[3], bounded): annealing Monte-Carlo trajectories of the rhf(6) design on every rank
    #      (ids sharded by rank, no traffic inside the loop), then the one exchange step: per-rank top-K on the device (adp_top_k) and
    #      an all_gather of the K records over NCCL
    mc_leg = None
    if wl == "fold_eval" and not args.no_mc_leg:
        try:
            from addapt_b200.sharding import gather_top_designs_device
            leg_traj, leg_steps, leg_k = 4096, 60, 1024
            prob = Problem(RHF6, {"active": RHF6_ACTIVE}, [("active", "apo", False, 1.0), ("active", "holo", True, 1.0)], motif=THEO)
            lb = torch.zeros(leg_traj, dtype=torch.float64, device=dev)
            lq = torch.zeros(leg_traj, len(RHF6), dtype=torch.uint8, device=dev)
            run = lambda: eng.mc_run_raw(prob, leg_traj, 2016, rank * leg_traj, leg_steps, ("annealing", 5.0, 0.0, 300), best_seq=lq, best_score=lb)
            run()                                                         # warm-up of the kernel ...
            gather_top_designs_device(eng, lb, lq, rank * leg_traj, leg_k, dist if world > 1 else None)     # ... and of the exchange (lazy module loads, NCCL channels)
            l0 = eng.launch_count()
            ea, eb, ec = (torch.cuda.Event(enable_timing=True) for _ in range(3))
            barrier()
            ea.record(stream)
            run()
            eb.record(stream)
            top_s, top_q, top_i = gather_top_designs_device(eng, lb, lq, rank * leg_traj, leg_k, dist if world > 1 else None)
            ec.record(stream)
            barrier()
            t = torch.tensor([ea.elapsed_time(eb), eb.elapsed_time(ec)], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            mc_ms, gather_ms = float(t[0].item()), float(t[1].item())
            mc_leg = {"mc_steps_per_s": world * leg_traj * leg_steps / (mc_ms * 1e-3), "gather_ms": gather_ms, "top_k": int(top_s.shape[0]),
                      "mc_config": "configs[3] bounded: %d annealing trajectories x %d steps per GPU of the 102-nt rhf(6) design, thermostat '5 to 0 in 300 steps'; "
                                   "per-rank top-%d on the device (adp_top_k), %s" % (leg_traj, leg_steps, leg_k, "all_gather of the K records over NCCL" if world > 1 else "single rank: no collective"),
                      "mc_best_score": float(top_s[0]), "mc_gpu_launches": eng.launch_count() - l0}
        except Exception as exc:                                   # the headline line must not depend on the second leg
            if world > 1:
                raise
            mc_leg = {"mc_leg_error": "%s: %s" % (type(exc).__name__, exc)}
    if rank == 0:
        line = {
            "metric": METRIC if wl == "fold_eval" and n == N_NT else "%s (%s, %d nt)" % (unit, wl, n),
            "value": value, "unit": unit, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64" if wl != "mfe" else "int32",
            "data": "synthetic",
            "config": {"workload": "configs[2]: batched apo+holo McCaskill pf+bppm of 65,536 sgRNA-templated 100-nt designs per GPU" if wl == "fold_eval" and n == N_NT and B == BATCH else "%s B=%d n=%d" % (wl, B, n),
                       "designs_per_gpu_per_step": B, "n_nt": n, "l2": "256 MiB flush between timed iterations", "pair_queries": nq,
                       "sharding": "independent designs per rank, no data-path collective",
                       "energy_params": "%s (%s)" % (eng.param_provenance, os.path.basename(eng.param_file))},
            "e2e": {"value": e2e_value, "unit": unit, "h2d_bytes_per_step": io[0], "d2h_bytes_per_step": io[1]},
            "gpu_launches": launches,
            "clocks": sampler.summary(),
            "roofline": roofline,
        }
        if cpu:
            line["cpu_baseline"] = cpu
        if mc_leg:
            line.update(mc_leg)
        print(json.dumps(line))
    if wl == "mc":
        # the one exchange step of the design run: every rank's best designs, merged to a global top-K over NCCL
        from addapt_b200.sharding import gather_top_designs_device
        t0 = time.perf_counter()
        top_s, top_q, top_i = gather_top_designs_device(eng, d_best, d_bseq, rank * mc_traj, min(1024, mc_traj), dist if world > 1 else None)
        if rank == 0:
            print(json.dumps({"top_designs_gather": {"k": int(top_s.shape[0]), "ranks": world, "best_score": float(top_s[0]),
                                                     "best_seq": bytes(top_q[0]).decode(), "seconds": time.perf_counter() - t0}}), file=sys.stderr)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    nccl_log_end(nccl_log)
    eng.close()


if __name__ == "__main__":
    main()
