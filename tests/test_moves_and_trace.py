"""CPU side of three rows of the hot-path scope, through the host emulation of the kernel source
against the oracle and the reference's own known-answer tables:
  * mutate_recursively / can_be_freely_mutated (reference tests/test_sampling.cc:8-112) through the table compiler and the DEVICE
    closure code (apply_closure) -- the -m gpu twin is in tests/test_gpu_design.py;
  * trajectory records in TsvTrajectoryReporter's layout (reference src/sampling.cc:430-489), readable by the reference's loader
    (tools/utils.py:6-51);
  * score terms over base-pair probabilities (RnaFold::base_pair_prob, reference include/scoring.hh:48)."""
import json
import os

import numpy as np
import pytest

from addapt_b200 import trajectory
from addapt_b200.capi import Problem
from addapt_b200.sharding import pick_best_scores
from tests import cases

GOLD = os.path.join(os.path.dirname(__file__), "golden")
REF = json.load(open(os.path.join(GOLD, "reference_thresholds.json")))
SC_RTOL = 1e-6


def table_problem(seq, macros):
    return Problem(seq, {str(k): m for k, m in enumerate(macros)}, [])


def test_mutate_recursively_table_through_the_device_closure_code(emu, oracle):
    """reference tests/test_sampling.cc:47-99: every row, through compile_closure + apply_closure (what mc_kernel runs)"""
    for seq, macros, muts, expected in REF["mutate_recursively"]["table"]:
        pr = table_problem(seq, macros)
        for i, (nt, want) in enumerate(zip(muts, expected)):
            got, _ = emu.mutate(pr, seq, i, nt)
            assert got == want, (seq, macros, i, nt)
            assert oracle.mutate(pr.as_oracle_dict(), seq, i, nt) == want
    for seq, macros, i, nt in REF["mutate_recursively"]["throws"]:
        with pytest.raises(ValueError):
            emu.mutate(table_problem(seq, macros), seq, i, nt)
        with pytest.raises(ValueError):
            oracle.mutate(table_problem(seq, macros).as_oracle_dict(), seq, i, nt)


def test_can_be_freely_mutated_table(emu):
    """reference tests/test_sampling.cc:8-45"""
    for case in REF["mutable_positions"]["cases"]:
        pr = table_problem(case["seq"], case["macrostates"])
        for i, (mut, free) in enumerate(zip(case["can_be_mutated"], case["can_be_freely_mutated"])):
            if not mut:
                with pytest.raises(ValueError):          # lower case: mutate_recursively refuses (the reference's Device::mutate throws)
                    emu.mutate(pr, case["seq"], i, "A")
                continue
            try:
                _, f = emu.mutate(pr, case["seq"], i, "A")
            except ValueError:
                assert not free                           # a partner is lower case: the reference would throw at that step
                continue
            assert f == bool(free), (case["seq"], i)


def _check_trace(a, b, n_terms):
    """a, b: (values [cap, 8+nt], seqs [cap, 2, n]) of two implementations"""
    va, qa = a
    vb, qb = b
    assert va.shape == vb.shape and np.array_equal(qa, qb)
    assert np.array_equal(va[:, 0], vb[:, 0]) and np.array_equal(va[:, 7], vb[:, 7])          # step, outcome
    assert np.array_equal(va[:, 6], vb[:, 6])                                                 # random threshold: bitwise (integer arithmetic)
    np.testing.assert_allclose(va[:, 4], vb[:, 4], rtol=1e-5)          # temperature: the annealing schedule is bit-equal (checked in
                                                                       # test_gpu_parity), the auto-scaling one is a median of score differences
    for col in (1, 2, 3, 5) + tuple(range(8, 8 + n_terms)):
        fin = np.isfinite(vb[:, col])
        assert np.array_equal(np.isfinite(va[:, col]), fin)
        np.testing.assert_allclose(va[fin, col], vb[fin, col], rtol=1e-5, atol=1e-9)


def small_problem(bp=False):
    seq = "gggaaaCCUGAUACCAGCCGAAAGGCCCUUGGCAGCAGGuuucc"
    act = "....." + "." * 0 + "(" + "." * 31 + ")" + "......"
    act = ("." * 6 + "(" + "." * 31 + ")" + "." * 5)[:len(seq)]
    terms = [("act", "apo", False, 1.0), ("act", "holo", True, 2.0)]
    bps = [("pair", "holo", True, 0.5, 6, 38), ("density", "apo", False, 1.5, [0, 1, 2, 40, 41]), ("pair", "apo", False, 1.0, 7, 37)] if bp else None
    return Problem(seq, {"act": act}, terms, motif=cases.THEO, bp_terms=bps)


@pytest.mark.parametrize("fast", [False, True])
def test_trajectory_records_match_the_oracle(emu, oracle, fast):
    pr = small_problem()
    for interval, steps in ((1, 24), (5, 23)):
        e = emu.mc_run(pr, 11, 3, steps, fast=fast, trace_interval=interval)
        o = oracle.mc_run(pr.as_oracle_dict(), 11, 3, steps, trace_interval=interval)
        assert e["final_seq"] == o["final_seq"] and e["outcomes"] == o["outcomes"]
        _check_trace((e["trace_values"], e["trace_seqs"]), (o["trace_values"], o["trace_seqs"]), pr.n_term_values)
        assert e["trace_values"].shape[0] == (steps + interval - 1) // interval
        assert list(e["trace_values"][:, 0]) == list(range(0, steps, interval))


def test_tsv_is_readable_by_the_reference_loader(emu, tmp_path):
    """the file has the reference's layout: tools/utils.py:6-31 reads it with pandas.read_table(comment='#') and takes the
    initial sequence from the first line; tools/utils.py:36-51 picks peaks over the current_score column"""
    pd = pytest.importorskip("pandas")
    pr = small_problem()
    steps = 40
    r = emu.mc_run(pr, 5, 0, steps, trace_interval=1)
    path = str(tmp_path / "traj.tsv")
    trajectory.write_tsv(path, pr, r["trace_values"], r["trace_seqs"], steps)
    lines = open(path).read().split("\n")
    assert lines[0] == "#\tinitial_seq\t" + pr.seq
    assert lines[1].split("\t")[:4] == ["step", "num_steps", "current_score", "proposed_score"]
    assert lines[1].split("\t")[4:8] == ["term_weight[apo: not act]", "term_value[apo: not act]", "term_weight[holo: act]", "term_value[holo: act]"]
    assert lines[1].endswith("current_seq\tproposed_seq\t") and lines[2].endswith("\t")
    with open(path) as f:                                     # load_initial_seq
        _, label, value = f.readline().split("\t")
    assert label == "initial_seq" and value.strip() == pr.seq
    traj = pd.read_table(path, comment="#")                   # load_trajectory
    assert len(traj) == steps and list(traj["step"]) == list(range(steps)) and set(traj["num_steps"]) == {steps}
    assert set(traj["outcome"]) <= set(trajectory.OUTCOMES) and set(traj["move"]) == {"UnbiasedMutation"}
    assert traj["current_seq"].iloc[-1] == r["final_seq"]
    np.testing.assert_allclose(traj["current_score"].to_numpy(), r["trace_values"][:, 1], rtol=1e-5)
    counts = [int((traj["outcome"] == name).sum()) for name in trajectory.OUTCOMES]
    assert counts == r["outcomes"]
    # peak picking over the column, the reference's way (tools/utils.py:36-51) and the repo's restatement agree
    scores = traj["current_score"].to_numpy()
    if scores.min() < np.percentile(scores, 75):
        ref_scores, peaks = scores.copy(), []
        blocked, thr = ref_scores.min(), np.percentile(ref_scores, 75)
        while ref_scores.max() >= thr:
            i = int(ref_scores.argmax())
            peaks.append(i)
            ref_scores[max(i - 3, 0):min(i + 3, len(ref_scores))] = blocked
        assert pick_best_scores(scores, 3) == sorted(peaks)


@pytest.mark.parametrize("fast", [False, True])
def test_bp_terms_match_the_oracle(emu, oracle, fast):
    pr = small_problem(bp=True)
    seqs = [pr.seq.upper(), pr.seq.upper().replace("CCUG", "GCUG", 1), pr.seq.upper().replace("AGG", "UGG", 1)]
    s1, t1 = emu.score(pr, seqs, want_terms=True, fast=fast)
    for b, s in enumerate(seqs):
        s0, t0 = oracle.score(pr.as_oracle_dict(), s, want_terms=True)
        assert s1[b] == pytest.approx(s0, rel=SC_RTOL)
        np.testing.assert_allclose(t1[b], t0, rtol=SC_RTOL)
    assert t1.shape[1] == 5
    # with contexts the positions shift with the 5' flank
    prc = Problem(pr.seq, pr.macrostates, pr.terms, motif=cases.THEO, contexts=[("", ""), ("ACGUAC", "GG")], bp_terms=pr.bp_terms)
    s1, t1 = emu.score(prc, seqs[:1], want_terms=True, fast=fast)
    s0, t0 = oracle.score(prc.as_oracle_dict(), seqs[0], want_terms=True)
    assert s1[0] == pytest.approx(s0, rel=SC_RTOL)
    np.testing.assert_allclose(t1[0], t0, rtol=SC_RTOL)
    # and a Monte-Carlo run over the extended score function, step for step
    e = emu.mc_run(pr, 3, 1, 12, fast=fast, trace_interval=1)
    o = oracle.mc_run(pr.as_oracle_dict(), 3, 1, 12, trace_interval=1)
    assert e["final_seq"] == o["final_seq"] and e["outcomes"] == o["outcomes"]
    _check_trace((e["trace_values"], e["trace_seqs"]), (o["trace_values"], o["trace_seqs"]), pr.n_term_values)


def test_holo_fills_derived_from_the_apo_twin(emu, oracle):
    """The fast path derives a ligand-bound ensemble from its apo fill, Z_holo = Z_apo (1 + (boltz - 1) P(plain hairpin at the
    motif site)) with the site probability from an outside pass over the cells enclosing the site (fold_jobs.cuh, PfJob::site_ratio);
    the oracle folds the holo ensemble directly.  One site (the shortcut), no site (ratio 1), two sites (falls back to a real fill),
    contexts, and the float-rounded reference mode."""
    pr = cases.rhf6_problem()
    designs = cases.templated_designs(4, 4, n=102)
    s1, t1 = emu.score(pr, designs, want_terms=True, fast=True)
    s0, t0 = oracle.score_batch(pr.as_oracle_dict(), designs)
    np.testing.assert_allclose(s1, s0, rtol=1e-12)
    np.testing.assert_allclose(t1, t0, rtol=1e-12)
    m = cases.THEO[0]
    twice = "GGGAAA" + m + "AAAA" + m + "UUUCCC"
    cst = "(" + "." * (len(twice) - 2) + ")"
    pr2 = Problem(twice, {"m": cst}, [("m", "apo", False, 1.0), ("m", "holo", True, 1.0)], motif=cases.THEO)
    assert emu.score(pr2, [twice], fast=True)[0] == pytest.approx(oracle.score(pr2.as_oracle_dict(), twice), rel=1e-12)
    none = "GGGAAACCCUUUGGGAAACCCAAAGGGUUUCCC"
    pr3 = Problem(none, {"m": "(" + "." * (len(none) - 2) + ")"}, [("m", "apo", False, 1.0), ("m", "holo", True, 1.0)], motif=cases.THEO)
    assert emu.score(pr3, [none], fast=True)[0] == pytest.approx(oracle.score(pr3.as_oracle_dict(), none), rel=1e-12)
    prc = cases.rhf6_problem(contexts=cases.SPACERS[:3])
    np.testing.assert_allclose(emu.score(prc, designs[:2], fast=True), oracle.score_batch(prc.as_oracle_dict(), designs[:2])[0], rtol=1e-12)
    prf = cases.rhf6_problem(float_dg=True)
    np.testing.assert_allclose(emu.score(prf, designs[:2], fast=True), oracle.score_batch(prf.as_oracle_dict(), designs[:2], float_dg=True)[0], rtol=1e-4)
    # only holo terms: the apo twin is added behind the scenes
    prh = Problem(pr.seq, pr.macrostates, [("active", "holo", True, 1.0)], motif=cases.THEO)
    np.testing.assert_allclose(emu.score(prh, designs[:2], fast=True), oracle.score_batch(prh.as_oracle_dict(), designs[:2])[0], rtol=1e-12)


def test_enforced_non_canonical_fixed_pair_is_refused(emu):
    """ViennaRNA keeps an enforced non-canonical pair as pair type 7; the kernels here would silently give it weight 0.  A design
    problem whose macrostate enforces such a pair between two fixed positions is refused with an error instead."""
    seq = "gggaaaCCCaaaUUUaga"            # g0 .. a17 enforced: G-A
    cst = "(" + "." * (len(seq) - 2) + ")"
    with pytest.raises(ValueError, match="non-canonical"):
        emu.score(Problem(seq, {"m": cst}, [("m", "apo", True, 1.0)]), [seq.upper()])
    ok = "gggaaaCCCaaaUUUagc"             # G-C: fine
    emu.score(Problem(ok, {"m": cst}, [("m", "apo", True, 1.0)]), [ok.upper()])


def test_picks_from_trajectory_files(emu, tmp_path):
    """the pick / merge half of SURVEY.md 8(f1): trajectory files of several runs -> the designs the reference's tools would
    select (tools/utils.py:27-51), read back without pandas"""
    pr = small_problem()
    paths = []
    for t in range(3):
        r = emu.mc_run(pr, 5, t, 40, trace_interval=1)
        p = str(tmp_path / ("traj.%d.tsv" % t))
        trajectory.write_tsv(p, pr, r["trace_values"], r["trace_seqs"], 40)
        paths.append(p)
        traj = trajectory.load_trajectory(p)
        assert trajectory.load_initial_seq(p) == pr.seq
        assert traj["step"] == [float(s) for s in range(40)] and traj["current_seq"][-1] == r["final_seq"]
        np.testing.assert_allclose(traj["current_score"], r["trace_values"][:, 1], rtol=1e-5)
        assert set(traj["outcome"]) <= set(trajectory.OUTCOMES)
    picks = trajectory.pick_best_seqs(paths, 3)
    assert picks and {p[0] for p in picks} <= set(paths)
    for path, i, step, score, seq in picks:
        traj = trajectory.load_trajectory(path)
        assert step == i and seq == traj["current_seq"][i] and score == traj["current_score"][i]
        assert score >= np.percentile(traj["current_score"], 75) - 1e-12       # a peak is at least at the 75th percentile
        assert len(seq) == len(pr.seq)


def _same_tsv(path_a, path_b):
    la, lb = open(path_a).read().split("\n"), open(path_b).read().split("\n")
    assert la[:2] == lb[:2] and len(la) == len(lb)
    for x, y in zip(la[2:], lb[2:]):
        cx, cy = x.split("\t"), y.split("\t")
        assert len(cx) == len(cy)
        for u, v in zip(cx, cy):
            if u != v:                                    # numbers printed with 6 digits may differ in the last one
                assert float(u) == pytest.approx(float(v), rel=2e-5, abs=1e-9), (x, y)


GOLDEN_TSV = os.path.join(GOLD, "oracle_trajectory.tsv")


def test_golden_trajectory_file(emu, oracle, tmp_path):
    """tests/golden/oracle_trajectory.tsv (tests/golden/make_trajectory_golden.py): the reference's on-disk format, the
    random-stream convention and the scores, frozen; a fresh oracle run and the emulation of the kernel source reproduce it"""
    pr = small_problem()
    head = open(GOLDEN_TSV).read().split("\n")[:2]
    assert head[0] == "#\tinitial_seq\t" + pr.seq
    assert head[1] == ("step\tnum_steps\tcurrent_score\tproposed_score\tterm_weight[apo: not act]\tterm_value[apo: not act]\tterm_weight[holo: act]\t"
                       "term_value[holo: act]\tscore_diff\ttemperature\tmetropolis_criterion\trandom_threshold\tmove\toutcome\tcurrent_seq\tproposed_seq\t")
    o = oracle.mc_run(pr.as_oracle_dict(), 2016, 7, 30, trace_interval=1)
    a = str(tmp_path / "oracle.tsv")
    trajectory.write_tsv(a, pr, o["trace_values"], o["trace_seqs"], 30)
    _same_tsv(a, GOLDEN_TSV)
    for fast in (False, True):
        e = emu.mc_run(pr, 2016, 7, 30, fast=fast, trace_interval=1)
        b = str(tmp_path / ("emu%d.tsv" % fast))
        trajectory.write_tsv(b, pr, e["trace_values"], e["trace_seqs"], 30)
        _same_tsv(b, GOLDEN_TSV)


def test_nccl_log_helpers(tmp_path, monkeypatch, capsys):
    from addapt_b200 import sharding
    for var in ("NCCL_DEBUG", "NCCL_DEBUG_SUBSYS", "NCCL_DEBUG_FILE"):
        monkeypatch.delenv(var, raising=False)
    assert sharding.nccl_log_begin(1) is None and "NCCL_DEBUG" not in os.environ            # one rank: nothing to log
    monkeypatch.setenv("NCCL_DEBUG", "VERSION")                                              # quieter than INFO: raised
    path = sharding.nccl_log_begin(2)
    assert os.environ["NCCL_DEBUG"] == "INFO" and os.environ["NCCL_DEBUG_SUBSYS"] == "INIT" and os.environ["NCCL_DEBUG_FILE"] == path
    open(path, "w").write("host:1:1 [0] NCCL INFO comm 0x1 rank 0 nranks 2\n")
    sharding.nccl_log_end(path)
    assert "nranks 2" in capsys.readouterr().err and not os.path.exists(path)
    monkeypatch.setenv("NCCL_DEBUG_FILE", "/somewhere/else")                                 # the caller's own file is respected
    assert sharding.nccl_log_begin(2) is None and os.environ["NCCL_DEBUG_FILE"] == "/somewhere/else"
