#!/usr/bin/env python3
"""One-off large randomized run of the EMULATED CTA (tests/emu/simt.hh: the kernel source executed by all threads of a CTA as fibers,
seeded thread schedules) beyond the pytest cases: random lengths, alphabets, hard constraints, motif placements, job lists, code
paths (fast / Qm outside the matrix block / level sums in memory / general with 1, 2, 4 lanes per cell; 128-512 threads) and three
random schedules per case.  Every case must give the same bits (outputs and matrix fingerprint) under its schedules and the oracle's
numbers.  A CPU-only race / divergence hunt; prints a summary line and exits non-zero on the first mismatch.
   python tools/stress_simt.py [seconds] [seed] [--design]
--design: the score and Monte-Carlo kernel bodies instead (random designs of the rhf(6) problem with random flanking contexts, and a
44-nt problem with base-pair-probability terms; random thermostats, trajectory ids, step counts)."""
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
subprocess.check_call(["make", "-C", os.path.join(ROOT, "tests", "emu"), "-s"])
from oracle.oracle import Oracle  # noqa: E402
from tests import cases  # noqa: E402
from tests.emu_binding import Emu  # noqa: E402

args = [a for a in sys.argv[1:] if not a.startswith("--")]
budget = float(args[0]) if len(args) > 0 else 120.0
emu, orc = Emu(), Oracle()
rng = np.random.Generator(np.random.PCG64(int(args[1]) if len(args) > 1 else 31337))
POLICIES = ["random", "random", "random_warp", "ascending", "descending"]
t0 = time.time()


def design_stress():
    from addapt_b200.capi import Problem
    n_score = n_mc = n_sched = 0
    small_seq = "gggaaaCCUGAUACCAGCCGAAAGGCCCUUGGCAGCAGGuuucc"
    small_act = ("." * 6 + "(" + "." * 31 + ")" + "." * 5)[:len(small_seq)]
    while time.time() - t0 < budget:
        if rng.random() < 0.5:
            ctx = [cases.SPACERS[k] for k in sorted(rng.choice(5, int(rng.integers(1, 3)), replace=False))] if rng.random() < 0.5 else None
            pr = cases.rhf6_problem(contexts=ctx)
            designs = cases.templated_designs(int(rng.integers(1, 3)), int(rng.integers(0, 1 << 30)), n=102)
            qm_global = bool(ctx) and rng.random() < 0.5
        else:
            bps = [("pair", "holo", True, 0.5, 6, 38), ("density", "apo", False, 1.5, [0, 1, 2, 40, 41]), ("pair", "apo", False, 1.0, 7, 37)]
            pr = Problem(small_seq, {"act": small_act}, [("act", "apo", False, 1.0), ("act", "holo", True, 2.0)], motif=cases.THEO,
                         bp_terms=bps if rng.random() < 0.7 else None)
            base = list(small_seq.upper())
            designs = []
            for _ in range(int(rng.integers(1, 4))):
                d = list(base)
                for p in rng.choice(np.arange(6, 39), 3, replace=False):
                    d[p] = "ACGU"[int(rng.integers(0, 4))]
                designs.append("".join(d))
            qm_global = False
        fast = rng.random() < 0.8
        if rng.random() < 0.5:
            s0, t0_ = orc.score_batch(pr.as_oracle_dict(), designs)
            first = None
            for _ in range(3):
                policy, sseed = POLICIES[int(rng.integers(0, len(POLICIES)))], int(rng.integers(0, 1 << 30))
                s1, t1, _st = emu.simt_score(pr, designs, fast=fast, qm_global=qm_global and fast, threads=256, policy=policy, seed=sseed)
                fin = np.isfinite(s0)
                assert np.array_equal(np.isfinite(s1), fin) and np.allclose(s1[fin], s0[fin], rtol=1e-6), ("score", designs, policy, sseed)
                key = s1.tobytes() + t1.tobytes()
                assert first is None or key == first, ("schedule-dependent score", designs, fast, policy, sseed)
                first = key
                n_sched += 1
            n_score += 1
        else:
            thermo = [("annealing", 5.0, 0.0, 300), ("fixed", float(rng.uniform(0.2, 3.0))), ("auto", 0.5, int(rng.integers(1, 5)), 1.0)][int(rng.integers(0, 3))]
            seed, tid, n_traj, steps = int(rng.integers(0, 1 << 30)), int(rng.integers(0, 1 << 40)), int(rng.integers(1, 3)), int(rng.integers(2, 8))
            want = [orc.mc_run(pr.as_oracle_dict(), seed, tid + k, steps, thermostat=thermo) for k in range(n_traj)]
            first = None
            for _ in range(3):
                policy, sseed = POLICIES[int(rng.integers(0, len(POLICIES)))], int(rng.integers(0, 1 << 30))
                got = emu.simt_mc(pr, seed, tid, n_traj, steps, thermostat=thermo, fast=fast, threads=256, policy=policy, sched_seed=sseed)
                for a, b in zip(want, got):
                    assert b["final_seq"] == a["final_seq"] and b["outcomes"] == a["outcomes"] and b["best_step"] == a["best_step"], ("mc", thermo, seed, tid, policy, sseed)
                    assert abs(b["final_score"] - a["final_score"]) <= 1e-6 * abs(a["final_score"]), ("mc score", thermo, seed, tid)
                assert first is None or got == first, ("schedule-dependent trajectory", thermo, seed, tid, policy, sseed)
                first = got
                n_sched += 1
            n_mc += 1
    print("emulated-CTA design stress ok: %d score cases, %d Monte-Carlo cases, %d schedules; no schedule-dependent bit, no mismatch against the oracle"
          % (n_score, n_mc, n_sched))


if "--design" in sys.argv:
    design_stress()
    sys.exit(0)
n_case = n_job = n_sched = 0
by_path = {}
while time.time() - t0 < budget:
    kind = "pf" if rng.random() < 0.6 else "mfe"
    r = rng.random()
    if r < 0.45:
        path, threads, group, n = "fast", 256, 1, int(rng.integers(5, 129))
    elif r < 0.55 and kind == "pf":
        threads = 256 if rng.random() < 0.7 else 512
        path, group, n = "fast_qm_ext", 1, int(rng.integers(90, 129 if threads == 256 else 180))
    elif r < 0.65:
        threads = int(rng.choice([64, 128, 256]))
        path, group, n = "fast_mem_levels", 1, int(rng.integers(threads // 2 + 1, min(threads + 40, 200)))
    elif r < 0.75:
        path, threads, group, n = "fast", 512, 1, int(rng.integers(100, 200))
    else:
        path, threads, group, n = "general", int(rng.choice([128, 256])), int(rng.choice([1, 2, 4])), int(rng.integers(5, 90))
    J = int(rng.integers(1, 5))
    alphabet = ["ACGU", "GC", "GU", "ACGU", "AU", "ACGU"][int(rng.integers(0, 6))]
    seqs = ["".join(alphabet[k] for k in rng.integers(0, len(alphabet), n)) for _ in range(J)]
    use_motif = [int(rng.random() < 0.5) for _ in range(J)]
    m = cases.THEO[0]
    for j in range(J):
        if use_motif[j] and n >= len(m) + 4 and rng.random() < 0.7:
            at = int(rng.integers(0, n - len(m) + 1))
            seqs[j] = seqs[j][:at] + m + seqs[j][at + len(m):]
    csts = [cases.random_constraint(orc, s, rng) if rng.random() < 0.4 else None for s in seqs]
    outside = [int(rng.random() < 0.8) for _ in range(J)]
    first = None
    for _ in range(3):
        policy, sseed = POLICIES[int(rng.integers(0, len(POLICIES)))], int(rng.integers(0, 1 << 30))
        what = (kind, path, threads, group, n, seqs, csts, use_motif, outside, policy, sseed)
        if kind == "pf":
            got = emu.simt_pf(seqs, csts=csts, motif=cases.THEO, use_motif=use_motif, outside=outside, path=path, threads=threads, group=group,
                              policy=policy, seed=sseed)
            key = (got["dG"].tobytes(), got["bppm"].tobytes(), got["matrices"])
        else:
            got = emu.simt_mfe(seqs, csts=csts, motif=cases.THEO, use_motif=use_motif, path=path if path != "fast_qm_ext" else "fast",
                               threads=threads, group=group, policy=policy, seed=sseed)
            key = (got["e"].tobytes(), tuple(got["structures"]), got["matrices"])
        assert not got["declined"].any(), what
        n_sched += 1
        if first is None:
            first = key
            for j, s in enumerate(seqs):
                mo = cases.THEO if use_motif[j] else None
                if kind == "pf":
                    d0, b0 = orc.pf(s, cst=csts[j], motif=mo, bppm=True)
                    if np.isinf(d0):
                        assert got["dG"][j] == d0, ("finiteness", what)
                        continue
                    assert abs(got["dG"][j] - d0) <= 1e-9 * max(abs(d0), 1e-3), ("dG", what)
                    if outside[j]:
                        assert float((np.abs(got["bppm"][j] - b0) - 1e-6 * np.abs(b0)).max()) <= 1e-9, ("bppm", what)
                else:
                    assert (int(got["e"][j]), got["structures"][j]) == orc.mfe(s, cst=csts[j], motif=mo), ("mfe", what)
        else:
            assert key == first, ("schedule-dependent result", what)
    n_case += 1
    n_job += J
    by_path[(kind, path)] = by_path.get((kind, path), 0) + 1
print("emulated-CTA stress ok: %d cases (%d jobs, %d schedules), lengths 5-199, no schedule-dependent bit, no mismatch against the oracle"
      % (n_case, n_job, n_sched))
print("cases by path:", ", ".join("%s %s %d" % (k[0], k[1], v) for k, v in sorted(by_path.items())))
