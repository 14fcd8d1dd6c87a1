"""Kernel bodies (addapt_b200/csrc/*_core.cuh) compiled for the host with one virtual thread, against the
oracle.  Runs without a GPU; checks the recursions, constraint handling, motif, traceback, score and
Monte-Carlo logic of the very source the sm_100a kernels are built from (the parallel decomposition itself is
covered by the -m gpu tests).  Tolerances: MFE energy + structure bit exact; ensemble energy 1e-9 relative;
bppm 1e-9 absolute + 1e-6 relative; scores 1e-6 relative."""
import math

import numpy as np
import pytest

from tests import cases

DG_RTOL, BP_ATOL, BP_RTOL, SC_RTOL = 1e-9, 1e-9, 1e-6, 1e-6


def check_fold(oracle, emu, s, cst=None, motif=None):
    d0, b0 = oracle.pf(s, cst=cst, motif=motif, bppm=True)
    d1, b1 = emu.pf(s, cst=cst, motif=motif, bppm=True)
    if math.isinf(d0):
        assert d1 == d0
    else:
        assert d1 == pytest.approx(d0, rel=DG_RTOL, abs=1e-12)
        np.testing.assert_allclose(b1, b0, rtol=BP_RTOL, atol=BP_ATOL)
    assert emu.mfe(s, cst=cst, motif=motif) == oracle.mfe(s, cst=cst, motif=motif)
    if len(s) <= 320 and set(s.upper()) <= set("ACGUT"):
        # the fast path (pf_fast.cuh: register-resident generic-loop levels, trimmed matrices, Qb overwritten by the
        # outside pass) must describe the same ensemble
        d2, b2 = emu.pf(s, cst=cst, motif=motif, bppm=True, fast=True)
        if math.isinf(d0):
            assert d2 == d0
        else:
            assert d2 == pytest.approx(d0, rel=DG_RTOL, abs=1e-12)
            np.testing.assert_allclose(b2, b0, rtol=BP_RTOL, atol=BP_ATOL)
            assert emu.pf(s, cst=cst, motif=motif, fast=True) == d2          # inside-only run
        # ... and the (min,+) transcription of the same scheme the same optimum and the same traceback
        assert emu.mfe(s, cst=cst, motif=motif, fast=True) == oracle.mfe(s, cst=cst, motif=motif)


def test_reference_sequences(oracle, emu):
    for s in (cases.HAIRPIN, cases.THEO[0], cases.RHF6):
        check_fold(oracle, emu, s)
        check_fold(oracle, emu, s, motif=cases.THEO)
    check_fold(oracle, emu, cases.RHF6, cst=cases.RHF6_ACTIVE)
    check_fold(oracle, emu, cases.RHF6, cst=cases.RHF6_ACTIVE, motif=cases.THEO)
    for c in ("...(....)...", "xxxxxxxxxxxx", "....xxxx....", "((((....))))", "(..........)", "xxxx........"):
        check_fold(oracle, emu, cases.HAIRPIN, cst=c)


def test_random_lengths(oracle, emu):
    rng = np.random.Generator(np.random.PCG64(1))
    for _ in range(120):
        n = int(rng.integers(1, 150))
        check_fold(oracle, emu, "".join("ACGU"[k] for k in rng.integers(0, 4, n)))


def test_long_hairpin_extrapolation(oracle, emu):
    # hairpin loops longer than 30 nt use the logarithmic extrapolation
    s = "GGGGG" + "A" * 60 + "CCCCC"
    check_fold(oracle, emu, s)
    check_fold(oracle, emu, s, cst="(((((" + "x" * 60 + ")))))")


def test_random_constraints_and_motif(oracle, emu):
    rng = np.random.Generator(np.random.PCG64(2))
    for t in range(80):
        n = int(rng.integers(20, 110))
        s = "".join("ACGU"[k] for k in rng.integers(0, 4, n))
        check_fold(oracle, emu, s, cst=cases.random_constraint(oracle, s, rng), motif=cases.THEO if t % 2 else None)
    for t in range(40):
        a, b = int(rng.integers(5, 40)), int(rng.integers(5, 40))
        s = "".join("ACGU"[k] for k in rng.integers(0, 4, a)) + cases.THEO[0] + "".join("ACGU"[k] for k in rng.integers(0, 4, b))
        check_fold(oracle, emu, s, motif=cases.THEO)
        check_fold(oracle, emu, s, cst=cases.random_constraint(oracle, s, rng), motif=cases.THEO)


def test_fast_path_queries_and_decline(oracle, emu):
    s = cases.RHF6[:100].upper()
    qs = [(0, 29), (47, 79), (79, 47), (50, 76), (54, 72), (5, 5), (3, 200), (10, 12)]
    d, b, q = emu.pf(s, motif=cases.THEO, bppm=True, fast=True, queries=qs)
    d0, b0 = oracle.pf(s, motif=cases.THEO, bppm=True)
    from addapt_b200.capi import tri_index
    want = [b0[tri_index(100, i, j)] if i != j and max(i, j) < 100 else 0.0 for i, j in qs]
    np.testing.assert_allclose(q, want, rtol=BP_RTOL, atol=BP_ATOL)
    with pytest.raises(NotImplementedError):
        emu.pf("ACGNACGUACGU", fast=True)
    rng = np.random.Generator(np.random.PCG64(5))
    for alphabet in ("GC", "GU", "AU"):                   # densest pairing: every level of the loop recursion is busy
        for n in (31, 64, 100, 128):
            check_fold(oracle, emu, "".join(alphabet[k] for k in rng.integers(0, 2, n)))


def test_lowercase_and_T(oracle, emu):
    assert emu.pf("acgtgaaaacgt") == pytest.approx(oracle.pf(cases.HAIRPIN), rel=1e-12)
    assert emu.mfe("acgtgaaaacgt") == oracle.mfe(cases.HAIRPIN)


def test_empty_ensemble_and_errors(oracle, emu):
    assert emu.pf("AAAAAAAAAA", cst="(........)") == math.inf
    assert emu.mfe("AAAAAAAAAA", cst="(........)")[0] == 10000000
    for bad in ("((..", "..))", "ab..", "..."):
        with pytest.raises(ValueError):
            emu.pf("ACGU", cst=bad)


def test_eval_and_motif_contribution(oracle, emu):
    assert emu.kT == oracle.kT
    assert emu.motif_contribution(cases.THEO) == oracle.motif_contribution(cases.THEO)
    rng = np.random.Generator(np.random.PCG64(3))
    for _ in range(50):
        n = int(rng.integers(10, 120))
        s = "".join("ACGU"[k] for k in rng.integers(0, 4, n))
        e, st = oracle.mfe(s)
        assert emu.eval_structure(s, st) == oracle.eval_structure(s, st) == e


def test_score(oracle, emu):
    designs = cases.templated_designs(12, 4, n=102)
    for float_dg in (False, True):
        pr = cases.rhf6_problem(float_dg=float_dg)
        s0, t0 = oracle.score_batch(pr.as_oracle_dict(), designs, float_dg=float_dg)
        for fast in (False, True):
            s1, t1 = emu.score(pr, designs, want_terms=True, fast=fast)
            np.testing.assert_allclose(s1, s0, rtol=1e-4 if float_dg else SC_RTOL)
            np.testing.assert_allclose(t1, t0, rtol=1e-4 if float_dg else SC_RTOL)


def test_score_with_contexts(oracle, emu):
    designs = cases.templated_designs(2, 5, n=102)
    pr = cases.rhf6_problem(contexts=cases.SPACERS)
    s0, t0 = oracle.score_batch(pr.as_oracle_dict(), designs)
    s1, t1 = emu.score(pr, designs, want_terms=True)
    np.testing.assert_allclose(s1, s0, rtol=SC_RTOL)
    np.testing.assert_allclose(t1, t0, rtol=SC_RTOL)
    s2, t2 = emu.score(pr, designs, want_terms=True, fast=True)
    np.testing.assert_allclose(s2, s0, rtol=SC_RTOL)
    np.testing.assert_allclose(t2, t0, rtol=SC_RTOL)
    # the score is the SUM over contexts (reference src/scoring.cc:123-132, tests/test_scoring.cc:302-330)
    single = sum(emu.score(cases.rhf6_problem(contexts=[c]), designs) for c in cases.SPACERS)
    np.testing.assert_allclose(s1, single, rtol=1e-12)


def test_monte_carlo_trajectory(oracle, emu):
    pr = cases.rhf6_problem()
    for seed, traj, thermo in ((2016, 7, ("annealing", 5.0, 0.0, 300)), (1, 123456789012, ("fixed", 1.0)), (3, 0, ("annealing", 2.0, 0.0, 7)),
                               (4, 5, ("auto", 0.5, 6, 1.0)), (9, 11, ("auto", 0.3, 1, 0.25))):
        a = oracle.mc_run(pr.as_oracle_dict(), seed, traj, 30, thermostat=thermo)
        for fast in (False, True):
            b = emu.mc_run(pr, seed, traj, 30, thermostat=thermo, fast=fast)
            assert b["final_seq"] == a["final_seq"] and b["best_seq"] == a["best_seq"]
            assert b["outcomes"] == a["outcomes"] and b["best_step"] == a["best_step"]
            assert b["final_score"] == pytest.approx(a["final_score"], rel=SC_RTOL)
            assert b["best_score"] == pytest.approx(a["best_score"], rel=SC_RTOL)
            assert sum(b["outcomes"]) == 30


def test_move_table_errors(emu):
    from addapt_b200.capi import Problem
    bad = Problem("Nn", {"m": "()"}, [("m", "apo", True, 1.0)])       # partner not mutable: tests/test_sampling.cc:101-103
    with pytest.raises(ValueError):
        emu.mc_run(bad, 0, 0, 1)


def test_shared_memory_budget(emu):
    """Occupancy guard: a B200 SM has 228 KB of shared memory, 1 KB of it reserved per resident CTA.  The benchmarked constructs
    must keep two (PF: 100-nt folds, the 102-nt rhf(6) design problem inside the score / Monte-Carlo kernels) and four (MFE, 100 nt)
    constructs per SM -- losing one halves the occupancy (measured: -22 % on the design run)."""
    per_sm = 228 * 1024
    assert emu.smem_bytes(0, 100) <= per_sm // 2 - 1024
    assert emu.smem_bytes(1, 102) <= per_sm // 2 - 1024
    assert emu.smem_bytes(2, 100) <= per_sm // 4 - 1024
