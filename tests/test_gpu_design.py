"""-m gpu twins of tests/test_moves_and_trace.py: the same three scope rows through the C ABI on the GPU."""
import json
import os

import numpy as np
import pytest

from addapt_b200 import trajectory
from tests import cases
from tests.test_moves_and_trace import REF, SC_RTOL, _check_trace, small_problem, table_problem

pytestmark = pytest.mark.gpu


def test_mutate_recursively_table_on_the_device(engine, oracle):
    """reference tests/test_sampling.cc:47-112 through adp_mutate_batch: closures compiled by the table compiler, applied by
    mutate_kernel with the code mc_kernel uses for its proposals"""
    import addapt_b200
    for seq, macros, muts, expected in REF["mutate_recursively"]["table"]:
        pr = table_problem(seq, macros)
        got, _ = engine.mutate(pr, [seq] * len(muts), list(range(len(muts))), list(muts))
        assert got == expected, (seq, macros)
    for seq, macros, i, nt in REF["mutate_recursively"]["throws"]:
        with pytest.raises(addapt_b200.EngineError):
            engine.mutate(table_problem(seq, macros), [seq], [i], [nt])
    for case in REF["mutable_positions"]["cases"]:      # tests/test_sampling.cc:8-45
        pr = table_problem(case["seq"], case["macrostates"])
        for i, (mut, free) in enumerate(zip(case["can_be_mutated"], case["can_be_freely_mutated"])):
            if not mut:
                with pytest.raises(addapt_b200.EngineError):
                    engine.mutate(pr, [case["seq"]], [i], ["A"])
                continue
            try:
                _, f = engine.mutate(pr, [case["seq"]], [i], ["A"])
            except addapt_b200.EngineError:
                assert not free
                continue
            assert f == [bool(free)]
    # the rhf(6) design problem: every freely mutable position x every nucleotide against the oracle's recursive restatement
    pr = cases.rhf6_problem()
    free = [41, 43, 44, 45, 46, 47, 48, 49, 77, 78]
    triples = [(p, nt) for p in free for nt in "ACGU"]
    got, flags = engine.mutate(pr, [cases.RHF6] * len(triples), [p for p, _ in triples], [nt for _, nt in triples])
    assert all(flags)
    for (p, nt), g in zip(triples, got):
        assert g == oracle.mutate(pr.as_oracle_dict(), cases.RHF6, p, nt)
    assert engine.mutate(pr, [cases.RHF6], [79], ["A"])[1] == [False]          # the 3' partner is mutable, not freely mutable


def test_monte_carlo_draws_the_table_moves(engine, oracle):
    """one step per trajectory on the reference's chained cases: the proposal is always accepted (the score is -inf on both
    sides: NaN criterion), so the final sequence is the closure of the Philox-drawn (position, nucleotide)"""
    from addapt_b200.capi import Problem
    for seq, macros in (("NNNN", ["()..", "(())"]), ("NNNN", ["(.).", "(())"]), ("NNN", ["().", ".()"])):
        start = seq.replace("N", "A")
        pr = Problem(start, {str(k): m for k, m in enumerate(macros)}, [(str(0), "apo", True, 1.0)])
        free = [i for i in range(len(seq)) if all(m[i] != ")" for m in macros)]
        r = engine.mc_run(pr, 32, 99, 1, thermostat=("fixed", 1.0))
        for t in range(32):
            w = oracle.philox(99, 0, [0, 0, t, 0])
            pos, nt = free[(w[0] * len(free)) >> 32], "ACGU"[w[1] >> 30]
            assert r["final_seq"][t] == oracle.mutate(pr.as_oracle_dict(), start, pos, nt), (seq, macros, t)


def test_trajectory_records_on_the_device(engine, oracle, tmp_path):
    pr = small_problem()
    for thermo in (("annealing", 5.0, 0.0, 7), ("auto", 0.5, 4, 1.0)):
        r = engine.mc_run(pr, 3, 11, 30, thermostat=thermo, traj_id0=2, trace_interval=1)
        for t in range(3):
            o = oracle.mc_run(pr.as_oracle_dict(), 11, 2 + t, 30, thermostat=thermo, trace_interval=1)
            _check_trace((r["trace_values"][t], r["trace_seqs"][t]), (o["trace_values"], o["trace_seqs"]), pr.n_term_values)
    # a strided record, and a run split into launches of 7 steps: the records must not depend on the split
    whole = engine.mc_run(pr, 4, 5, 45, trace_interval=4)
    os.environ["ADDAPT_B200_MC_CHUNK"] = "7"
    try:
        parts = engine.mc_run(pr, 4, 5, 45, trace_interval=4)
    finally:
        del os.environ["ADDAPT_B200_MC_CHUNK"]
    assert np.array_equal(whole["trace_values"], parts["trace_values"], equal_nan=True) and np.array_equal(whole["trace_seqs"], parts["trace_seqs"])
    o = oracle.mc_run(pr.as_oracle_dict(), 5, 1, 45, trace_interval=4)
    _check_trace((whole["trace_values"][1], whole["trace_seqs"][1]), (o["trace_values"], o["trace_seqs"]), pr.n_term_values)
    # the GPU trajectory in the reference's on-disk format, row for row against the oracle's
    a, b = str(tmp_path / "gpu.tsv"), str(tmp_path / "oracle.tsv")
    r = engine.mc_run(pr, 1, 5, 45, traj_id0=1, trace_interval=1)
    o = oracle.mc_run(pr.as_oracle_dict(), 5, 1, 45, trace_interval=1)
    trajectory.write_tsv(a, pr, r["trace_values"][0], r["trace_seqs"][0], 45)
    trajectory.write_tsv(b, pr, o["trace_values"], o["trace_seqs"], 45)
    la, lb = open(a).read().split("\n"), open(b).read().split("\n")
    assert la[:2] == lb[:2] and len(la) == len(lb) == 48
    for x, y in zip(la[2:-1], lb[2:-1]):
        cx, cy = x.split("\t"), y.split("\t")
        assert len(cx) == len(cy)
        for u, v in zip(cx, cy):
            if u != v:                                    # numbers printed with 6 digits may differ in the last one
                assert float(u) == pytest.approx(float(v), rel=2e-5, abs=1e-9), (x, y)


def test_bp_terms_on_the_device(engine, oracle):
    pr = small_problem(bp=True)
    seqs = [pr.seq.upper(), pr.seq.upper().replace("CCUG", "GCUG", 1), pr.seq.upper().replace("AGG", "UGG", 1)] * 3
    for general in (False, True):
        engine.set_general_only(general)
        try:
            s1, t1 = engine.score(pr, seqs, want_terms=True)
        finally:
            engine.set_general_only(False)
        for b, s in enumerate(seqs[:3]):
            s0, t0 = oracle.score(pr.as_oracle_dict(), s, want_terms=True)
            assert s1[b] == pytest.approx(s0, rel=SC_RTOL)
            np.testing.assert_allclose(t1[b], t0, rtol=SC_RTOL)
        assert np.array_equal(s1[:3], s1[3:6]) and np.array_equal(t1[:3], t1[6:9])
    # the rhf(6) design with the reference's pair sets as terms: the holo-only pairs wanted in holo, unwanted in apo
    g = REF["rhf6"]
    holo_pairs = [(i, j) for i, j, *_ in g["holo_pairs"]][:3]
    bps = [("pair", "holo", True, 1.0, i, j) for i, j in holo_pairs] + [("pair", "apo", False, 1.0, i, j) for i, j in holo_pairs[:2]]
    bps.append(("density", "apo", True, 0.5, list(range(30, 46))))
    base = cases.rhf6_problem(contexts=cases.SPACERS[:2])
    from addapt_b200.capi import Problem
    pr6 = Problem(base.seq, base.macrostates, base.terms, motif=cases.THEO, contexts=base.contexts, bp_terms=bps)
    designs = cases.templated_designs(6, 4, n=102)
    s1, t1 = engine.score(pr6, designs, want_terms=True)
    s0, t0 = oracle.score_batch(pr6.as_oracle_dict(), designs)
    np.testing.assert_allclose(s1, s0, rtol=SC_RTOL)
    np.testing.assert_allclose(t1, t0, rtol=SC_RTOL)
    r = engine.mc_run(pr6, 2, 3, 10, trace_interval=1)
    for t in range(2):
        o = oracle.mc_run(pr6.as_oracle_dict(), 3, t, 10, trace_interval=1)
        assert r["final_seq"][t] == o["final_seq"] and list(r["outcomes"][t]) == o["outcomes"]
        _check_trace((r["trace_values"][t], r["trace_seqs"][t]), (o["trace_values"], o["trace_seqs"]), pr6.n_term_values)


def test_top_k_on_the_device(engine):
    """adp_top_k against a host sort: best first, ties by lower index, -inf and NaN at the end, k > n, k = 1"""
    from addapt_b200.sharding import local_top_k
    rng = np.random.Generator(np.random.PCG64(8))
    for n, k in ((65536, 1024), (5000, 1024), (1000, 16), (7, 1024), (1, 1), (4096, 1), (70000, 1000)):
        s = np.round(rng.normal(size=n), 1)                     # rounded: many ties, also across the threshold
        s[rng.integers(0, n, max(1, n // 50))] = -np.inf
        idx, top = engine.top_k(s, k)
        want = local_top_k(s, k)
        assert np.array_equal(idx, want), (n, k)
        assert np.array_equal(top, s[want])
    s = rng.normal(size=1100)
    s[::7] = np.nan
    s[5] = s[6] = np.nanmax(s)
    idx, top = engine.top_k(s, 1024)
    fin = ~np.isnan(s)
    want = [i for i in np.lexsort((np.arange(1100), -np.where(fin, s, -np.inf))) if fin[i]]
    assert len(want) < 1024 and list(idx[:len(want)]) == want and np.isnan(top[len(want):]).all()
    assert np.all(np.diff(idx[len(want):]) > 0)                 # the NaN tail in index order
    assert list(engine.top_k(np.zeros(2000), 5)[0]) == [0, 1, 2, 3, 4]
    assert list(engine.top_k(np.array([0.0, -0.0, 0.0]), 3)[0]) == [0, 1, 2]


def test_gather_top_designs_on_the_device(engine):
    """the design run's exchange step with the per-rank reduction on the device equals the host path"""
    import torch
    from addapt_b200.sharding import gather_top_designs, gather_top_designs_device
    rng = np.random.Generator(np.random.PCG64(9))
    m, n, k = 8192, 102, 1024
    scores = np.round(rng.normal(size=m), 2)
    seqs = rng.integers(65, 70, (m, n)).astype(np.uint8)
    a = gather_top_designs(scores, seqs, np.arange(500, 500 + m), k)
    b = gather_top_designs_device(engine, torch.from_numpy(scores).cuda(), torch.from_numpy(seqs).cuda(), 500, k)
    for x, y in zip(a, b):
        assert np.array_equal(x, y)


def test_golden_trajectory_file_on_the_device(engine, tmp_path):
    """the frozen oracle trajectory (tests/golden/oracle_trajectory.tsv) reproduced by the Monte-Carlo kernel"""
    from tests.test_moves_and_trace import GOLDEN_TSV, _same_tsv
    pr = small_problem()
    r = engine.mc_run(pr, 1, 2016, 30, traj_id0=7, trace_interval=1)
    path = str(tmp_path / "gpu.tsv")
    trajectory.write_tsv(path, pr, r["trace_values"][0], r["trace_seqs"][0], 30)
    _same_tsv(path, GOLDEN_TSV)
