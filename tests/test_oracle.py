"""The CPU oracle against everything the reference's own tests pin for this path (tests/golden/
reference_thresholds.json, transcribed from reference tests/test_scoring.cc:83-259) and against its own
committed vectors.  Exact parity with ViennaRNA is UNPINNED (no libRNA in this image, no exact golden numbers in
the reference); what is pinned here are the reference's thresholds and comment-level energies."""
import json
import math
import os

import numpy as np
import pytest

from oracle.oracle import tri
from tests import cases

GOLD = os.path.join(os.path.dirname(__file__), "golden")
REF = json.load(open(os.path.join(GOLD, "reference_thresholds.json")))
VEC = json.load(open(os.path.join(GOLD, "oracle_vectors.json")))


def _cmp(p, op, t):
    return p > t if op == ">" else p < t


def test_kT_and_bonus(oracle):
    assert oracle.kT == pytest.approx(0.61632, abs=1e-5)
    assert oracle.kT * math.log(0.32 / 1e6) == pytest.approx(-9.22, abs=5e-3)  # tests/test_scoring.cc:154


def test_hairpin_thresholds(oracle):
    g = REF["hairpin"]
    s, n = g["seq"], len(g["seq"])
    assert oracle.mfe(s) == (round(g["comment_energy_kcal"] * 100), g["comment_structure"])
    assert oracle.eval_structure(s, g["comment_structure"]) == round(g["comment_energy_kcal"] * 100)
    _, bp = oracle.pf(s, bppm=True)
    expected = {(i, j): t for i, j, t in g["pairs_gt"]}
    for i in range(n):
        for j in range(i, n):
            p = bp[tri(n, i, j)]
            if (i, j) in expected:
                assert p > expected[(i, j)]
            else:
                assert p < g["other_pairs_lt"]
    for c, op, t in g["macrostates"]:
        assert _cmp(oracle.macrostate_prob(s, c), op, t), c
        assert _cmp(oracle.macrostate_prob(s, c, float_dg=True), op, t), c


def test_aptamer_thresholds(oracle):
    g = REF["aptamer"]
    s, n, motif = g["seq"], len(g["seq"]), tuple(g["motif"])
    assert oracle.eval_structure(s, g["comment_apo"][0]) == round(g["comment_apo"][1] * 100)
    assert oracle.mfe(s) == (round(g["comment_apo"][1] * 100), g["comment_apo"][0])
    assert oracle.mfe(s, motif=motif)[1] == g["comment_holo"][0]
    _, apo = oracle.pf(s, bppm=True)
    _, holo = oracle.pf(s, motif=motif, bppm=True)
    A, H = {tuple(p) for p in g["apo_pairs"]}, {tuple(p) for p in g["holo_pairs"]}
    for i in range(n):
        for j in range(i, n):
            pa, ph = apo[tri(n, i, j)], holo[tri(n, i, j)]
            assert (pa > g["present_gt"]) if (i, j) in A else (pa < g["absent_lt"])
            assert (ph > g["present_gt"]) if (i, j) in H else (ph < g["absent_lt"])
    for cond, c, op, t in g["macrostates"]:
        assert _cmp(oracle.macrostate_prob(s, c, motif=motif if cond == "holo" else None), op, t)


def test_rhf6_thresholds(oracle):
    g = REF["rhf6"]
    s, n = g["seq"].upper(), len(g["seq"])
    dga, apo = oracle.pf(s, bppm=True)
    dgh, holo = oracle.pf(s, motif=cases.THEO, bppm=True)
    for i, j, t in g["constitutive"]:
        assert apo[tri(n, i, j)] > t and holo[tri(n, i, j)] > t
    for i, j, t in g["apo_pairs"]:
        assert apo[tri(n, i, j)] > t and holo[tri(n, i, j)] < g["apo_pairs_holo_lt"]
    for i, j, th, ta in g["holo_pairs"]:
        assert apo[tri(n, i, j)] < ta and holo[tri(n, i, j)] > th
    assert oracle.macrostate_prob(s, g["active"]) < g["p_apo_active_lt"]
    assert oracle.macrostate_prob(s, g["active"], motif=cases.THEO) > g["p_holo_active_gt"]


# ---- the reference's per-position data for rhf(6) (reference tests/test_scoring.cc:52-55: two RNAfold -p strings and two ensemble
# energies).  With the RECONSTRUCTED parameter file shipped in params/ the oracle does NOT reproduce them exactly; the tests
# below state the true residual instead of hiding it behind a wide tolerance: they are strict xfails that a genuine
# rna_turner2004.par (ADDAPT_B200_PAR=/path/to/rna_turner2004.par) turns into passes, and the documented residual itself is pinned.
PAIRING = json.load(open(os.path.join(GOLD, "reference_pairing_strings.json")))


def _reconstructed(oracle):
    return oracle.param_provenance == "reconstructed"


def _mismatches(a, b):
    return [k for k in range(len(a)) if a[k] != b[k]]


def test_rhf6_holo_pairing_string_matches_the_reference(oracle):
    s = PAIRING["seq"]
    _, bp = oracle.pf(s, motif=cases.THEO, bppm=True)
    assert cases.rnafold_symbols(bp, len(s)) == PAIRING["holo"]["pairing"]


def test_rhf6_apo_pairing_string_matches_the_reference(oracle, request):
    if _reconstructed(oracle):
        request.applymarker(pytest.mark.xfail(strict=True, reason="reconstructed int11/int21/int22: 7 of 102 apo symbols differ"))
    s = PAIRING["seq"]
    _, bp = oracle.pf(s, bppm=True)
    got = cases.rnafold_symbols(bp, len(s))
    bad = _mismatches(got, PAIRING["apo"]["pairing"])
    print("apo pairing string: %d of %d symbols differ from the reference, at %s" % (len(bad), len(s), bad))
    assert got == PAIRING["apo"]["pairing"]


@pytest.mark.parametrize("cond", ["apo", "holo"])
def test_rhf6_ensemble_energy_at_comment_precision(oracle, request, cond):
    if _reconstructed(oracle):
        request.applymarker(pytest.mark.xfail(strict=True, reason="reconstructed tables: apo -29.752 vs -29.58, holo -33.827 vs -33.82"))
    dg = oracle.pf(PAIRING["seq"], motif=cases.THEO if cond == "holo" else None)
    print("%s ensemble dG %.4f, reference comment %.2f, residual %+.4f kcal/mol" % (cond, dg, PAIRING[cond]["dG_ensemble_kcal"], dg - PAIRING[cond]["dG_ensemble_kcal"]))
    assert abs(dg - PAIRING[cond]["dG_ensemble_kcal"]) <= PAIRING["comment_precision_kcal"]


def test_rhf6_residual_of_the_reconstructed_tables_is_the_documented_one(oracle):
    """pins the size of the gap (profiles/r2_param_sensitivity.txt attribute it to named table entries)"""
    if not _reconstructed(oracle):
        pytest.skip("a genuine parameter file is loaded")
    r, s = PAIRING["reconstructed_par_residual"], PAIRING["seq"]
    dga, apo = oracle.pf(s, bppm=True)
    dgh, holo = oracle.pf(s, motif=cases.THEO, bppm=True)
    assert dga == pytest.approx(r["apo_dG"], abs=5e-4) and dgh == pytest.approx(r["holo_dG"], abs=5e-4)
    assert _mismatches(cases.rnafold_symbols(apo, len(s)), PAIRING["apo"]["pairing"]) == r["apo_mismatch_positions"]
    assert _mismatches(cases.rnafold_symbols(holo, len(s)), PAIRING["holo"]["pairing"]) == r["holo_mismatch_positions"]


def test_term_counts_match_survey(oracle):
    w_in, w_out = oracle.count_terms(cases.RHF6.upper())       # SURVEY.md 8(d) table, row rhf(6)
    assert w_in == 795843
    assert w_out == pytest.approx(404528, rel=0.01)   # the survey's script and the oracle differ by 0.6 % in the outside count


def test_oracle_matches_committed_vectors(oracle):
    for f in VEC["folds"]:
        s = f["seq"]
        assert list(oracle.mfe(s)) == f["mfe"]
        assert list(oracle.mfe(s, motif=cases.THEO)) == f["mfe_holo"]
        assert list(oracle.mfe(s, cst=f["cst"])) == f["mfe_cst"]
        dg, bp = oracle.pf(s, bppm=True)
        assert dg == pytest.approx(f["dG"], rel=1e-12)
        assert oracle.pf(s, motif=cases.THEO) == pytest.approx(f["dG_holo"], rel=1e-12)
        assert oracle.pf(s, cst=f["cst"]) == pytest.approx(f["dG_cst"], rel=1e-12)
        assert bp.sum() == pytest.approx(f["bppm_sum"], rel=1e-10)
        for k, v in f["bppm_top"]:
            assert bp[k] == pytest.approx(v, rel=1e-10)
    assert oracle.philox(*VEC["philox"]["key"], VEC["philox"]["ctr"]) == VEC["philox"]["out"]


def test_philox_known_answer(oracle):
    # Random123 known-answer vectors for Philox4x32-10
    assert oracle.philox(0, 0, [0, 0, 0, 0]) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert oracle.philox(0xffffffff, 0xffffffff, [0xffffffff] * 4) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert oracle.philox(0xa4093822, 0x299f31d0, [0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344]) == [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_bppm_is_a_probability(oracle):
    for s in cases.random_seqs(8, 80, 5):
        n = len(s)
        dg, bp = oracle.pf(s, bppm=True)
        e, _ = oracle.mfe(s)
        assert dg <= e / 100 + 1e-9                       # ensemble free energy <= MFE
        P = np.zeros((n, n))
        for i in range(n):
            for j in range(i + 1, n):
                P[i, j] = P[j, i] = bp[tri(n, i, j)]
        assert P.min() >= 0 and P.sum(axis=1).max() <= 1 + 1e-9


def test_empty_constrained_ensemble(oracle):
    # an enforced pair between two bases that cannot pair: no structure satisfies it
    assert oracle.pf("AAAAAAAAAA", cst="(........)") == math.inf
    assert oracle.mfe("AAAAAAAAAA", cst="(........)")[0] == 10000000
    with pytest.raises(ValueError):
        oracle.pf("ACGU", cst="((..")
