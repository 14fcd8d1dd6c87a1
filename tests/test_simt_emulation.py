"""Race / divergence check of the kernel source without a GPU: the kernel bodies run by an emulated CTA with ALL of its threads
(tests/emu/simt.hh: one fiber per CUDA thread, real CTA barriers, warp collectives with the hardware's butterfly order) under
different thread schedules.  CUDA leaves the order of the threads between two barriers open, so a correct body must give

  * the same bits for every schedule (random thread, ascending, descending, one random warp at a time), and
  * the oracle's numbers (MFE energy + structure exactly, ensemble dG 1e-9 relative, bppm 1e-9 abs + 1e-6 rel, scores 1e-6),

and every lane of a warp must reach every warp collective, every thread every barrier.  A value read in the phase that writes it,
a buffer reused one phase too early, a counter reset while a slow thread still reads it, or a collective inside a divergent
branch shows up as differing bits, a wrong number or a reported deadlock.  compute-sanitizer, which would check this on the GPU,
is closed on the pool these kernels were developed on; the -m gpu suite adds bitwise reproducibility on the hardware itself.
Job lists mirror the __global__ wrappers: several jobs back to back on one CTA (apo, holo, constrained, inside-only), so the
hand-over between jobs is covered too."""
import os

import numpy as np
import pytest

from addapt_b200.capi import Problem
from tests import cases

SCHEDULES = [("random", 1), ("random", 2), ("random", 3), ("ascending", 0), ("descending", 0), ("random_warp", 5), ("random_warp", 6)]
DG_RTOL, BP_ATOL, BP_RTOL, SC_RTOL = 1e-9, 1e-9, 1e-6, 1e-6
# The default run (part of `pytest -m "not gpu"`) keeps every code path but only three schedules for the expensive cases and
# leaves the longest constructs out; ADP_SIMT_FULL=1 runs the whole matrix (about 8 minutes; profiles/r2_simt_emulation_full.txt).
FULL = bool(os.environ.get("ADP_SIMT_FULL"))
full_only = pytest.mark.skipif(not FULL, reason="ADP_SIMT_FULL=1 runs the whole schedule / length matrix")


def schedules(heavy):
    return SCHEDULES if FULL or not heavy else [SCHEDULES[0], SCHEDULES[4], SCHEDULES[5]]


# ---- the checker itself ----------------------------------------------------------------------------------------------------
def test_emulator_sees_a_missing_barrier(emu):
    n = 64
    want = n * (n - 1) // 2
    for policy, seed in SCHEDULES:
        assert emu.simt_selftest_missing_barrier(n, policy, seed, with_barrier=True) == want
    got = {emu.simt_selftest_missing_barrier(n, policy, seed, with_barrier=False) for policy, seed in SCHEDULES}
    assert len(got) > 1      # the unsynchronised read depends on the schedule


def test_emulator_sees_a_divergent_collective(emu):
    with pytest.raises(RuntimeError, match="deadlock"):
        emu.simt_selftest_divergent_collective(64, "random", 0)


@pytest.mark.parametrize("kind", ["pf", "mfe"])
def test_a_barrier_dropped_from_the_kernel_body_is_noticed(emu, kind):
    """Fault injection: the same fold with ONE of its CTA barriers skipped by every thread.  Almost every barrier of the column
    loops is load-bearing, so the schedule comparison must notice: other bits in the outputs or in the matrices the job leaves
    behind, or a deadlock.  The barriers of the LAST job are dropped (the matrix fingerprint is taken after it); the few that go
    unnoticed sit at the job boundary, where the wrapper's barrier and the body's own follow each other, and in the first
    columns, which hold no cells."""
    seqs = cases.random_seqs(2, 40, 5)
    run = (lambda **kw: emu.simt_pf(seqs, path="fast", **kw)) if kind == "pf" else (lambda **kw: emu.simt_mfe(seqs, path="fast", **kw))
    keys = ("dG", "bppm", "matrices") if kind == "pf" else ("e", "structures", "matrices")
    base = run(policy="random", seed=1)
    n_barriers = base["stats"][0]
    noticed = total = 0
    try:
        for k in range(n_barriers // 2 + 8, n_barriers - 2, 2 if FULL else 5):
            emu.simt_drop_barrier(k)
            found = False
            for policy, sseed in (("random", 1), ("descending", 0), ("random_warp", 5), ("ascending", 0)):
                try:
                    r = run(policy=policy, seed=sseed)
                except RuntimeError:
                    found = True
                    break
                if any(not np.array_equal(r[key], base[key]) for key in keys):
                    found = True
                    break
            total += 1
            noticed += found
    finally:
        emu.simt_drop_barrier(-1)
    assert total >= 10 and noticed >= 0.8 * total, (noticed, total)


# ---- folds -----------------------------------------------------------------------------------------------------------------
def fold_jobs(oracle, n, seed):
    """(seqs, csts, use_motif, outside): apo, holo, constrained apo, constrained holo, inside-only -- back to back on one CTA"""
    rng = np.random.Generator(np.random.PCG64(seed))
    seqs = ["".join("ACGU"[k] for k in rng.integers(0, 4, n)) for _ in range(5)]
    m = cases.THEO[0]
    if n >= len(m) + 8:
        at = int(rng.integers(2, n - len(m) - 2))
        seqs[1] = seqs[1][:at] + m + seqs[1][at + len(m):]
        seqs[3] = seqs[1]
    csts = [None, None, cases.random_constraint(oracle, seqs[2], rng), cases.random_constraint(oracle, seqs[3], rng), None]
    return seqs, csts, [0, 1, 0, 1, 0], [1, 1, 1, 1, 0]


def check_pf(oracle, emu, n, seed, **how):
    seqs, csts, use_motif, outside = fold_jobs(oracle, n, seed)
    first = None
    for policy, sseed in schedules(heavy=n >= 100):
        r = emu.simt_pf(seqs, csts=csts, motif=cases.THEO, use_motif=use_motif, outside=outside, policy=policy, seed=sseed, **how)
        assert not r["declined"].any()
        if first is None:
            first = r
            for j, s in enumerate(seqs):
                d0, b0 = oracle.pf(s, cst=csts[j], motif=cases.THEO if use_motif[j] else None, bppm=True)
                if np.isinf(d0):
                    assert r["dG"][j] == d0
                    continue
                assert r["dG"][j] == pytest.approx(d0, rel=DG_RTOL, abs=1e-12)
                if outside[j]:
                    np.testing.assert_allclose(r["bppm"][j], b0, rtol=BP_RTOL, atol=BP_ATOL)
            assert first["stats"][0] > n and first["stats"][1] > 0          # it did go through barriers and warp collectives
        else:
            assert np.array_equal(r["dG"], first["dG"]), (n, policy, sseed)
            assert np.array_equal(r["bppm"], first["bppm"]), (n, policy, sseed)
            assert r["matrices"] == first["matrices"], (n, policy, sseed)        # every cell of the matrix block, not only the outputs


@pytest.mark.parametrize("n", [5, 9, 17, 33, 48, 64, 100, pytest.param(102, marks=full_only), 128])
def test_fast_partition_function_is_schedule_independent(oracle, emu, n):
    check_pf(oracle, emu, n, 100 + n, path="fast", threads=256)


@pytest.mark.parametrize("n,threads", [(101, 256), pytest.param(128, 256, marks=full_only), pytest.param(140, 512, marks=full_only), pytest.param(200, 512, marks=full_only), (136, 512)])
def test_fast_partition_function_with_qm_outside_the_matrix_block(oracle, emu, n, threads):
    check_pf(oracle, emu, n, 300 + n, path="fast_qm_ext", threads=threads)      # the kQmGlobal kernels (101-128 nt) and beyond


@pytest.mark.parametrize("n,threads", [(40, 64), (150, 256), pytest.param(270, 512, marks=full_only)])
def test_fast_partition_function_with_level_sums_in_memory(oracle, emu, n, threads):
    check_pf(oracle, emu, n, 500 + n, path="fast_mem_levels", threads=threads)   # constructs longer than threads / 2


@pytest.mark.parametrize("n,threads,group", [(30, 128, 1), (47, 128, 2), (61, 256, 4), (100, 512, 4)])
def test_general_partition_function_is_schedule_independent(oracle, emu, n, threads, group):
    check_pf(oracle, emu, n, 700 + n, path="general", threads=threads, group=group)


def check_mfe(oracle, emu, n, seed, **how):
    seqs, csts, use_motif, _ = fold_jobs(oracle, n, seed)
    want = [oracle.mfe(s, cst=csts[j], motif=cases.THEO if use_motif[j] else None) for j, s in enumerate(seqs)]
    prints = set()
    for policy, sseed in schedules(heavy=n >= 100):
        r = emu.simt_mfe(seqs, csts=csts, motif=cases.THEO, use_motif=use_motif, policy=policy, seed=sseed, **how)
        assert not r["declined"].any()
        assert [(int(e), st) for e, st in zip(r["e"], r["structures"])] == want, (n, policy, sseed)
        prints.add(r["matrices"])
    assert len(prints) == 1, (n, "the matrices depend on the schedule")


@pytest.mark.parametrize("n", [5, 12, 31, 64, 100, 128])
def test_fast_mfe_is_schedule_independent(oracle, emu, n):
    check_mfe(oracle, emu, n, 900 + n, path="fast", threads=256)


@pytest.mark.parametrize("n,threads", [(150, 256), pytest.param(300, 512, marks=full_only)])
def test_fast_mfe_with_level_minima_in_memory(oracle, emu, n, threads):
    check_mfe(oracle, emu, n, 1100 + n, path="fast_mem_levels", threads=threads)


@pytest.mark.parametrize("n,threads,group", [(29, 128, 1), (52, 128, 2), (100, 256, 4)])
def test_general_mfe_is_schedule_independent(oracle, emu, n, threads, group):
    check_mfe(oracle, emu, n, 1300 + n, path="general", threads=threads, group=group)


def test_fast_path_hands_back_foreign_letters(emu):
    r = emu.simt_pf(["ACGUNACGUACGUACGU", "ACGUAACGUACGUACGU"], path="fast", policy="random", seed=4)
    assert list(r["declined"]) == [1, 0]          # the declined job leaves the CTA in a state the next job can start from
    r = emu.simt_mfe(["ACGUNACGUACGUACGU", "ACGUAACGUACGUACGU"], path="fast", policy="descending")
    assert list(r["declined"]) == [1, 0]


# ---- design kernels --------------------------------------------------------------------------------------------------------
def small_problem(bp):
    seq = "gggaaaCCUGAUACCAGCCGAAAGGCCCUUGGCAGCAGGuuucc"
    act = ("." * 6 + "(" + "." * 31 + ")" + "." * 5)[:len(seq)]
    terms = [("act", "apo", False, 1.0), ("act", "holo", True, 2.0)]
    bps = [("pair", "holo", True, 0.5, 6, 38), ("density", "apo", False, 1.5, [0, 1, 2, 40, 41]), ("pair", "apo", False, 1.0, 7, 37)] if bp else None
    return Problem(seq, {"act": act}, terms, motif=cases.THEO, bp_terms=bps)


@pytest.mark.parametrize("which", ["rhf6", "rhf6_spacers", "small_bp"])
def test_score_kernel_body_is_schedule_independent(oracle, emu, which):
    if which == "small_bp":
        pr = small_problem(bp=True)
        base = "GGGAAACCUGAUACCAGCCGAAAGGCCCUUGGCAGCAGGUUUCC"
        designs = [base, base.replace("CCUGAUACC", "CGUGAUACG", 1)]
    else:
        pr = cases.rhf6_problem(contexts=cases.SPACERS[:3] if which == "rhf6_spacers" else None)
        designs = cases.templated_designs(3 if FULL else 2, 21, n=102)
    s0, t0 = oracle.score_batch(pr.as_oracle_dict(), designs)
    first = None
    for fast in (True, False):
        for policy, sseed in schedules(heavy=True)[:3 if fast else 2]  if not FULL else (SCHEDULES if fast else SCHEDULES[:4]):
            s1, t1, stats = emu.simt_score(pr, designs, fast=fast, threads=256, policy=policy, seed=sseed)
            np.testing.assert_allclose(s1, s0, rtol=SC_RTOL)
            np.testing.assert_allclose(t1, t0, rtol=SC_RTOL, atol=1e-12)
            if first is None or first[0] != fast:
                first = (fast, s1, t1)
            else:
                assert np.array_equal(s1, first[1]) and np.array_equal(t1, first[2]), (which, fast, policy, sseed)


def test_score_kernel_body_with_qm_outside_the_matrix_block(oracle, emu):
    pr = cases.rhf6_problem(contexts=cases.SPACERS[1:2])          # 122-nt construct: the kQmGlobal design kernels
    designs = cases.templated_designs(2, 22, n=102)
    s0, t0 = oracle.score_batch(pr.as_oracle_dict(), designs)
    seen = set()
    for policy, sseed in schedules(heavy=True):
        s1, t1, _ = emu.simt_score(pr, designs, fast=True, qm_global=True, threads=256, policy=policy, seed=sseed)
        np.testing.assert_allclose(s1, s0, rtol=SC_RTOL)
        seen.add(s1.tobytes() + t1.tobytes())
    assert len(seen) == 1


@pytest.mark.parametrize("which,thermo", [("rhf6", ("annealing", 5.0, 0.0, 300)), pytest.param("rhf6", ("fixed", 1.0), marks=full_only),
                                          pytest.param("rhf6", ("auto", 0.5, 6, 1.0), marks=full_only),
                                          ("small", ("fixed", 1.0)), ("small", ("auto", 0.5, 6, 1.0))])
def test_monte_carlo_kernel_body_is_schedule_independent(oracle, emu, which, thermo):
    """trajectories back to back on one CTA, two launches each (carried state), against the oracle step for step"""
    pr = cases.rhf6_problem() if which == "rhf6" else small_problem(bp=False)
    n_traj, steps = (3, 12) if FULL else ((2, 8) if which == "small" else (2, 4))
    want = [oracle.mc_run(pr.as_oracle_dict(), 2016, 40 + k, steps, thermostat=thermo) for k in range(n_traj)]
    first = None
    for policy, sseed in schedules(heavy=True)[:5]:
        got = emu.simt_mc(pr, 2016, 40, n_traj, steps, thermostat=thermo, fast=True, threads=256, policy=policy, sched_seed=sseed)
        for a, b in zip(want, got):
            assert b["final_seq"] == a["final_seq"] and b["best_seq"] == a["best_seq"]
            assert b["outcomes"] == a["outcomes"] and b["best_step"] == a["best_step"] and sum(b["outcomes"]) == steps
            assert b["final_score"] == pytest.approx(a["final_score"], rel=SC_RTOL)
            assert b["best_score"] == pytest.approx(a["best_score"], rel=SC_RTOL)
        if first is None:
            first = got
        else:
            assert got == first, (thermo, policy, sseed)


def test_monte_carlo_general_path_and_bp_terms(oracle, emu):
    pr = small_problem(bp=True)
    want = oracle.mc_run(pr.as_oracle_dict(), 7, 3, 10)
    seen = []
    for fast in (True, False):
        for policy, sseed in SCHEDULES[:3] if FULL or fast else SCHEDULES[:2]:
            got = emu.simt_mc(pr, 7, 3, 1, 10, fast=fast, threads=128 if not fast else 256, policy=policy, sched_seed=sseed)[0]
            assert got["final_seq"] == want["final_seq"] and got["outcomes"] == want["outcomes"]
            assert got["final_score"] == pytest.approx(want["final_score"], rel=SC_RTOL)
            seen.append((fast, got["final_score"], got["best_score"]))
    assert len({s for s in seen if s[0]}) == 1 and len({s for s in seen if not s[0]}) == 1
