"""-m gpu: the sm_100a kernels through the C ABI (libaddapt_b200.so) against the oracle on the same seeded inputs,
against the committed golden vectors, and -- at BASELINE.json's full batch sizes -- through size-independent
properties.  Tolerances: MFE energy and structure bit exact; ensemble energy 1e-9 relative; bppm
1e-9 absolute + 1e-6 relative; probabilities / scores 1e-6 relative (1e-4 in the float-rounded reference mode)."""
import json
import math
import os

import numpy as np
import pytest

from tests import cases

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
REF = json.load(open(os.path.join(GOLD, "reference_thresholds.json")))
VEC = json.load(open(os.path.join(GOLD, "oracle_vectors.json")))
DG_RTOL, BP_ATOL, BP_RTOL, SC_RTOL = 1e-9, 1e-9, 1e-6, 1e-6


def check_batch(oracle, engine, seqs, cst=None, motif=None):
    d0, b0 = oracle.pf_batch(seqs, cst=cst, motif=motif, bppm=True)
    d1, b1 = engine.pf(seqs, cst=cst, motif=motif, bppm=True)
    fin = np.isfinite(d0)
    assert np.array_equal(np.isfinite(d1), fin)
    np.testing.assert_allclose(d1[fin], d0[fin], rtol=DG_RTOL, atol=1e-12)
    np.testing.assert_allclose(b1[fin], b0[fin], rtol=BP_RTOL, atol=BP_ATOL)
    d2 = engine.pf(seqs, cst=cst, motif=motif, bppm=False)          # inside-only launch
    np.testing.assert_allclose(d2[fin], d0[fin], rtol=DG_RTOL, atol=1e-12)
    e0, s0 = oracle.mfe_batch(seqs, cst=cst, motif=motif)
    e1, s1 = engine.mfe(seqs, cst=cst, motif=motif)
    assert np.array_equal(e1, e0)
    assert s1 == s0


def test_native_library_is_loaded(engine):
    maps = open("/proc/self/maps").read()
    assert "libaddapt_b200.so" in maps
    assert engine.dfma_rate() > 1000.0        # GFMA/s: a real GPU ran the probe kernel
    assert engine.launch_count() >= 4


def test_reference_sequences(oracle, engine):
    for s in (cases.HAIRPIN, cases.THEO[0], cases.RHF6):
        check_batch(oracle, engine, [s])
        check_batch(oracle, engine, [s], motif=cases.THEO)
    check_batch(oracle, engine, [cases.RHF6], cst=cases.RHF6_ACTIVE)
    check_batch(oracle, engine, [cases.RHF6], cst=cases.RHF6_ACTIVE, motif=cases.THEO)
    for c, _, _ in REF["hairpin"]["macrostates"]:
        check_batch(oracle, engine, [cases.HAIRPIN], cst=c)


def test_reference_thresholds_through_the_engine(engine):
    """reference tests/test_scoring.cc:197-259 run against the GPU path itself"""
    g = REF["rhf6"]
    n = len(g["seq"])
    pairs = [(i, j) for i, j, *_ in g["constitutive"] + g["apo_pairs"] + g["holo_pairs"]]
    dg, pr, _ = engine.fold_eval([g["seq"]], cases.THEO, queries=pairs)
    P = {q: (pr[0, 0, k], pr[0, 1, k]) for k, q in enumerate(pairs)}
    for i, j, t in g["constitutive"]:
        assert P[(i, j)][0] > t and P[(i, j)][1] > t
    for i, j, t in g["apo_pairs"]:
        assert P[(i, j)][0] > t and P[(i, j)][1] < g["apo_pairs_holo_lt"]
    for i, j, th, ta in g["holo_pairs"]:
        assert P[(i, j)][0] < ta and P[(i, j)][1] > th
    pr_ = cases.rhf6_problem()
    _, terms = engine.score(pr_, [g["seq"]], want_terms=True)
    p_apo, p_holo = 1 - math.exp(terms[0, 0]), math.exp(terms[0, 1])
    assert p_apo < g["p_apo_active_lt"] and p_holo > g["p_holo_active_gt"]


def test_golden_vectors(engine):
    for f in VEC["folds"]:
        s = f["seq"]
        e, st = engine.mfe([s])
        assert [int(e[0]), st[0]] == f["mfe"]
        e, st = engine.mfe([s], motif=cases.THEO)
        assert [int(e[0]), st[0]] == f["mfe_holo"]
        e, st = engine.mfe([s], cst=f["cst"])
        assert [int(e[0]), st[0]] == f["mfe_cst"]
        dg, bp = engine.pf([s], bppm=True)
        assert dg[0] == pytest.approx(f["dG"], rel=DG_RTOL)
        assert engine.pf([s], motif=cases.THEO)[0] == pytest.approx(f["dG_holo"], rel=DG_RTOL)
        assert engine.pf([s], cst=f["cst"])[0] == pytest.approx(f["dG_cst"], rel=DG_RTOL)
        assert bp[0].sum() == pytest.approx(f["bppm_sum"], rel=1e-8)
        for k, v in f["bppm_top"]:
            assert bp[0][k] == pytest.approx(v, rel=BP_RTOL)
    g = VEC["score"]
    sc, terms = engine.score(cases.rhf6_problem(), g["designs"], want_terms=True)
    np.testing.assert_allclose(sc, g["score"], rtol=SC_RTOL)
    np.testing.assert_allclose(terms, g["terms"], rtol=SC_RTOL)
    g = VEC["score_5ctx"]
    np.testing.assert_allclose(engine.score(cases.rhf6_problem(contexts=cases.SPACERS), g["designs"]), g["score"], rtol=SC_RTOL)
    g = VEC["mc"]
    r = engine.mc_run(cases.rhf6_problem(), 1, g["seed"], g["n_steps"], thermostat=tuple(g["thermostat"]), traj_id0=g["traj_id"])
    assert r["final_seq"][0] == g["final_seq"] and r["best_seq"][0] == g["best_seq"]
    assert list(r["outcomes"][0]) == g["outcomes"] and r["best_step"][0] == g["best_step"]
    assert r["best_score"][0] == pytest.approx(g["best_score"], rel=SC_RTOL)


@pytest.mark.parametrize("n", [1, 4, 5, 17, 64, 69, 100, 133, 150, 200, 261, 300])
def test_random_batches_by_length(oracle, engine, n):
    """every launch shape (lane-group size, shared vs global matrices) incl. the ragged smallest lengths"""
    check_batch(oracle, engine, cases.random_seqs(6 if n > 150 else 24, n, 100 + n))


def test_general_kernels_and_mixed_alphabet(oracle, engine):
    """the fast pf kernels decline sequences with letters outside ACGU/T; the general kernel folds those (and, with
    set_general_only, everything), and both agree with the oracle"""
    seqs = cases.random_seqs(12, 90, 77)
    mixed = list(seqs)
    mixed[3] = mixed[3][:40] + "N" + mixed[3][41:]
    mixed[8] = "N" * 5 + mixed[8][5:]
    check_batch(oracle, engine, mixed)
    check_batch(oracle, engine, mixed, motif=cases.THEO)
    engine.set_general_only(1)
    try:
        check_batch(oracle, engine, seqs)
        check_batch(oracle, engine, cases.templated_designs(8, 5, n=100), cst=cases.RHF6_ACTIVE[:100], motif=cases.THEO)
        check_batch(oracle, engine, cases.random_seqs(4, 140, 78))
    finally:
        engine.set_general_only(0)


def test_long_constructs_with_constraints_and_motif(oracle, engine):
    """the 512-thread fast kernels beyond 128 nt (level sums in registers up to 256 nt, in global scratch above; matrices in
    shared memory up to ~150 nt, in L2 above), with a hard constraint and the ligand motif"""
    rng = np.random.Generator(np.random.PCG64(17))
    for n in (140, 190, 270, 320):
        a, b = n // 3, n - n // 3 - len(cases.THEO[0])
        s = cases.random_seqs(1, a, n)[0] + cases.THEO[0] + cases.random_seqs(1, b, n + 1)[0]
        cst = cases.random_constraint(oracle, s, rng)
        check_batch(oracle, engine, [s, s], cst=cst, motif=cases.THEO)
        check_batch(oracle, engine, [s], motif=cases.THEO)


def test_constraints_and_motif(oracle, engine):
    rng = np.random.Generator(np.random.PCG64(7))
    seqs = cases.templated_designs(16, 8, n=100)
    active = cases.RHF6_ACTIVE[:100]
    check_batch(oracle, engine, seqs, cst=active)
    check_batch(oracle, engine, seqs, cst=active, motif=cases.THEO)
    for _ in range(12):
        n = int(rng.integers(30, 120))
        s = cases.random_seqs(1, n, int(rng.integers(1 << 30)))[0]
        check_batch(oracle, engine, [s] * 3, cst=cases.random_constraint(oracle, s, rng), motif=cases.THEO)
    s = "GGGGG" + "A" * 60 + "CCCCC"                      # > 30-nt hairpin: extrapolated loop energy
    check_batch(oracle, engine, [s])


def test_empty_ensemble_and_errors(engine):
    import addapt_b200
    assert engine.pf(["AAAAAAAAAA"], cst="(........)")[0] == math.inf
    assert engine.mfe(["AAAAAAAAAA"], cst="(........)")[0][0] == 10000000
    for bad in ("((..", "ab..", "..."):
        with pytest.raises(addapt_b200.EngineError):
            engine.pf(["ACGU"], cst=bad)
    with pytest.raises(addapt_b200.EngineError):
        engine.pf(["A" * 400])
    assert engine.pf(["acgtgaaaacgt"])[0] == pytest.approx(engine.pf([cases.HAIRPIN])[0], rel=1e-14)


def test_fold_eval_matches_oracle(oracle, engine):
    seqs = cases.templated_designs(32, 9, n=100)
    queries = [(0, 29), (47, 79), (50, 76), (33, 73), (76, 50), (5, 5)]
    dg, pr, full = engine.fold_eval(seqs, cases.THEO, queries=queries, full=True)
    n = 100
    from addapt_b200.capi import tri_index
    for b, s in enumerate(seqs):
        for c, motif in enumerate((None, cases.THEO)):
            d0, b0 = oracle.pf(s, motif=motif, bppm=True)
            assert dg[b, c] == pytest.approx(d0, rel=DG_RTOL)
            np.testing.assert_allclose(full[b, c], b0, rtol=BP_RTOL, atol=BP_ATOL)
            for k, (i, j) in enumerate(queries):
                want = b0[tri_index(n, i, j)] if i != j else 0.0        # order-free, a == b -> 0 (scoring.cc:46-50)
                assert pr[b, c, k] == pytest.approx(want, rel=BP_RTOL, abs=BP_ATOL)


def test_results_are_bitwise_reproducible(engine):
    """compute-sanitizer is not available on this pool, so data races are hunted the other way round: every result must be
    bit-identical from run to run, whatever the batch composition (which CTA folds which construct, in which order the
    cell lists are compacted) and whichever co-resident CTA shares the SM"""
    rng = np.random.Generator(np.random.PCG64(11))
    seqs = cases.templated_designs(600, 21, n=100)
    q = [(0, 29), (47, 79), (50, 76), (33, 73)]
    dg0, pr0, _ = engine.fold_eval(seqs, cases.THEO, queries=q)
    dg1, pr1, _ = engine.fold_eval(seqs, cases.THEO, queries=q)
    assert np.array_equal(dg0, dg1) and np.array_equal(pr0, pr1)
    perm = rng.permutation(len(seqs))
    dg2, pr2, _ = engine.fold_eval([seqs[k] for k in perm], cases.THEO, queries=q)
    assert np.array_equal(dg2, dg0[perm]) and np.array_equal(pr2, pr0[perm])
    sub = perm[:37]
    dg3, _, full3 = engine.fold_eval([seqs[k] for k in sub], cases.THEO, queries=q, full=True)
    assert np.array_equal(dg3, dg0[sub])
    _, _, full4 = engine.fold_eval([seqs[k] for k in sub], cases.THEO, queries=q, full=True)
    assert np.array_equal(full3, full4)
    cst = cases.RHF6_ACTIVE[:100]
    a = engine.pf(seqs, cst=cst, motif=cases.THEO)
    b = engine.pf([seqs[k] for k in perm], cst=cst, motif=cases.THEO)
    assert np.array_equal(b, a[perm])
    long_seqs = cases.random_seqs(40, 300, 5)              # level sums in global scratch, matrices in L2
    assert np.array_equal(engine.pf(long_seqs), engine.pf(long_seqs))
    pr = cases.rhf6_problem()
    designs = cases.templated_designs(300, 22, n=102)
    assert np.array_equal(engine.score(pr, designs), engine.score(pr, designs))
    e0, s0 = engine.mfe(seqs)
    e1, s1 = engine.mfe([seqs[k] for k in perm])
    assert np.array_equal(e1, e0[perm]) and s1 == [s0[k] for k in perm]
    r0 = engine.mc_run(pr, 48, 7, 25)
    r1 = engine.mc_run(pr, 48, 7, 25)
    assert r0["final_seq"] == r1["final_seq"] and np.array_equal(r0["best_score"], r1["best_score"])


def test_score_and_contexts(oracle, engine):
    designs = cases.templated_designs(24, 4, n=102)
    for float_dg in (False, True):
        pr = cases.rhf6_problem(float_dg=float_dg)
        s0, t0 = oracle.score_batch(pr.as_oracle_dict(), designs, float_dg=float_dg)
        s1, t1 = engine.score(pr, designs, want_terms=True)
        np.testing.assert_allclose(s1, s0, rtol=1e-4 if float_dg else SC_RTOL)
        np.testing.assert_allclose(t1, t0, rtol=1e-4 if float_dg else SC_RTOL)
    pr = cases.rhf6_problem(contexts=cases.SPACERS)
    s0, _ = oracle.score_batch(pr.as_oracle_dict(), designs[:4])
    np.testing.assert_allclose(engine.score(pr, designs[:4]), s0, rtol=SC_RTOL)


def test_monte_carlo_matches_oracle_step_for_step(oracle, engine):
    pr = cases.rhf6_problem()
    n_traj, steps = 6, 40
    for thermo in (("annealing", 5.0, 0.0, 300), ("fixed", 1.0), ("auto", 0.5, 8, 1.0)):
        r = engine.mc_run(pr, n_traj, 2016, steps, thermostat=thermo, traj_id0=1000)
        for t in range(n_traj):
            a = oracle.mc_run(pr.as_oracle_dict(), 2016, 1000 + t, steps, thermostat=thermo)
            assert r["final_seq"][t] == a["final_seq"] and r["best_seq"][t] == a["best_seq"]
            assert list(r["outcomes"][t]) == a["outcomes"] and r["best_step"][t] == a["best_step"]
            assert r["final_score"][t] == pytest.approx(a["final_score"], rel=SC_RTOL)
            assert r["best_score"][t] == pytest.approx(a["best_score"], rel=SC_RTOL)
    # long runs are split into launches that carry the state in device memory: the split must not change anything
    import os
    for thermo in (("annealing", 5.0, 0.0, 300), ("auto", 0.5, 8, 1.0)):
        whole = engine.mc_run(pr, 5, 77, 45, thermostat=thermo)
        os.environ["ADDAPT_B200_MC_CHUNK"] = "7"
        try:
            parts = engine.mc_run(pr, 5, 77, 45, thermostat=thermo)
        finally:
            del os.environ["ADDAPT_B200_MC_CHUNK"]
        assert whole["final_seq"] == parts["final_seq"] and whole["best_seq"] == parts["best_seq"]
        assert np.array_equal(whole["outcomes"], parts["outcomes"]) and np.array_equal(whole["best_step"], parts["best_step"])
        assert np.array_equal(whole["final_score"], parts["final_score"]) and np.array_equal(whole["best_score"], parts["best_score"])
    assert engine.mc_run(pr, 2, 3, 0)["best_step"].tolist() == [-1, -1]          # zero steps: the start sequence, scored
    # sharding invariance: trajectory ids, not launch geometry, define the streams
    whole = engine.mc_run(pr, 4, 5, 25, traj_id0=0)
    parts = [engine.mc_run(pr, 2, 5, 25, traj_id0=0), engine.mc_run(pr, 2, 5, 25, traj_id0=2)]
    assert whole["final_seq"] == parts[0]["final_seq"] + parts[1]["final_seq"]


def test_full_size_mfe_properties(oracle, engine):
    """BASELINE configs[1]: 65,536 random 100-nt MFE folds.  Bit-exact vs the oracle on a 1,024 sample; for every
    sequence: the structure is balanced, pairs are canonical, and its evaluated energy equals the reported MFE."""
    B, n = 65536, 100
    rng = np.random.Generator(np.random.PCG64(20260705))
    a = np.frombuffer(b"ACGU", dtype=np.uint8)[rng.integers(0, 4, (B, n))]
    e = np.zeros(B, dtype=np.int32)
    st = np.zeros((B, n), dtype=np.uint8)
    engine.mfe_raw(B, n, a, e, st)
    idx = rng.choice(B, 1024, replace=False)
    seqs = [bytes(a[i]).decode() for i in idx]
    e0, s0 = oracle.mfe_batch(seqs)
    assert np.array_equal(e[idx], e0)
    assert [bytes(st[i]).decode() for i in idx] == s0
    opens, closes = (st == ord("(")).sum(axis=1), (st == ord(")")).sum(axis=1)
    assert np.array_equal(opens, closes)
    assert (e <= 0).all()
    for i in rng.choice(B, 4096, replace=False):
        assert engine.eval_structure(bytes(a[i]).decode(), bytes(st[i]).decode()) == e[i]


def test_full_size_fold_eval_properties(oracle, engine):
    """BASELINE configs[2]: 65,536 sgRNA-templated 100-nt designs, apo+holo pf+bppm.  Oracle parity on a sample;
    for all: ensemble energy <= MFE energy, holo <= apo (the ligand only adds weight), probabilities in [0,1]."""
    B, n = 65536, 100
    seqs = cases.templated_designs(B, 20260706, n=n)
    pairs = [(i, j) for i, j, *_ in REF["rhf6"]["constitutive"] + REF["rhf6"]["holo_pairs"] if j < n]
    dg, pr, _ = engine.fold_eval(seqs, cases.THEO, queries=pairs)
    assert np.isfinite(dg).all() and (dg[:, 1] <= dg[:, 0] + 1e-9).all()
    assert (pr >= 0).all() and (pr <= 1 + 1e-9).all()
    e = engine.mfe(seqs, structures=False)
    assert (dg[:, 0] <= e / 100.0 + 1e-9).all()
    rng = np.random.Generator(np.random.PCG64(3))
    idx = rng.choice(B, 1024, replace=False)
    sub = [seqs[i] for i in idx]
    d0, b0 = oracle.pf_batch(sub, bppm=True)
    d1, b1 = oracle.pf_batch(sub, motif=cases.THEO, bppm=True)
    np.testing.assert_allclose(dg[idx, 0], d0, rtol=DG_RTOL)
    np.testing.assert_allclose(dg[idx, 1], d1, rtol=DG_RTOL)
    # full bppm of the 1,024 sampled designs (2 x 5,050 probabilities each) against the oracle, and the queried probabilities of
    # the full-size run against the same designs folded in a different batch (bitwise: they must not depend on the batch)
    dg_s, pr_s, full_s = engine.fold_eval(sub, cases.THEO, queries=pairs, full=True)
    assert np.array_equal(dg_s, dg[idx]) and np.array_equal(pr_s, pr[idx])
    np.testing.assert_allclose(full_s[:, 0], b0, rtol=BP_RTOL, atol=BP_ATOL)
    np.testing.assert_allclose(full_s[:, 1], b1, rtol=BP_RTOL, atol=BP_ATOL)
    from addapt_b200.capi import tri_index
    for k, (i, j) in enumerate(pairs):
        np.testing.assert_allclose(pr_s[:, 0, k], b0[:, tri_index(n, i, j)], rtol=BP_RTOL, atol=BP_ATOL)
        np.testing.assert_allclose(pr_s[:, 1, k], b1[:, tri_index(n, i, j)], rtol=BP_RTOL, atol=BP_ATOL)


def test_reference_pairing_strings_through_the_engine(oracle, engine):
    """reference tests/test_scoring.cc:52-55: the RNAfold -p strings of rhf(6), derived from the ENGINE's bppm.  Holo must equal
    the reference's string; apo equals it only with a genuine parameter file (tests/test_oracle.py states the residual of the
    reconstructed one) -- with either file the engine's string must equal the oracle's."""
    g = json.load(open(os.path.join(GOLD, "reference_pairing_strings.json")))
    s, n = g["seq"], len(g["seq"])
    assert engine.param_provenance == oracle.param_provenance
    dg, _, full = engine.fold_eval([s], cases.THEO, full=True)
    sym = [cases.rnafold_symbols(full[0, c], n) for c in (0, 1)]
    _, b_apo = oracle.pf(s, bppm=True)
    _, b_holo = oracle.pf(s, motif=cases.THEO, bppm=True)
    assert sym[0] == cases.rnafold_symbols(b_apo, n) and sym[1] == cases.rnafold_symbols(b_holo, n)
    assert sym[1] == g["holo"]["pairing"]
    if engine.param_provenance == "reconstructed":
        r = g["reconstructed_par_residual"]
        assert [k for k in range(n) if sym[0][k] != g["apo"]["pairing"][k]] == r["apo_mismatch_positions"]
        assert dg[0, 0] == pytest.approx(r["apo_dG"], abs=5e-4) and dg[0, 1] == pytest.approx(r["holo_dG"], abs=5e-4)
    else:
        assert sym[0] == g["apo"]["pairing"]
        assert abs(dg[0, 0] - g["apo"]["dG_ensemble_kcal"]) <= 0.005 and abs(dg[0, 1] - g["holo"]["dG_ensemble_kcal"]) <= 0.005
