"""The oracle's recursions against first principles: for short sequences EVERY secondary structure is listed, its energy taken
from the oracle's loop-by-loop evaluator (orc_eval_structure, the restatement of vrna_eval_structure that
vrna_sc_add_hi_motif itself calls, reference src/scoring.cc:94-99), and the ensemble summed directly:

    Z = sum_s exp(-E(s)/kT),   P(i,j) = sum_{s contains (i,j)} exp(-E(s)/kT) / Z,   MFE = min_s E(s).

With dangles=2 the energy of a structure is a sum over its loops, so McCaskill's inside/outside recursions (orc_pf) and Zuker's
recursion (orc_mfe) must reproduce these sums exactly (up to rounding) -- for the plain ensemble, under a hard constraint
(DB_DEFAULT | ENFORCE_BP: 'x' unpaired, matching brackets paired), and with a ligand hairpin motif, whose bound state
replaces the empty hairpin at the occurrence and hands its probability to the motif's inner pairs.
One published rule separates the two energy models: libRNA's partition-function parameters pass the dangle energies through
SMOOTH() (params.c), so a stem at a sequence end, whose single dangling neighbour is worth 0 in the table, weighs
exp(+1.698 dcal/kT) in the ensemble but 0 in vrna_eval_structure / vrna_mfe.  `pf_energy` applies that rule, restated here from
the parameter file, to the exterior stems that touch an end; everywhere else (both neighbours present: mismatch tables, all
<= -30 dcal/mol, which SMOOTH leaves alone) the two models coincide.
The recursions and the evaluator share the energy tables but no code path, so this pins the dynamic programming itself;
the GPU kernels are then pinned to the oracle by the -m gpu parity tests."""
import math

import numpy as np
import pytest

from oracle.oracle import tri

PAIRS = {("A", "U"), ("U", "A"), ("G", "C"), ("C", "G"), ("G", "U"), ("U", "G")}
TURN = 3


def all_structures(seq):
    """every non-crossing set of canonical pairs with hairpin loops of at least 3 bases, as sorted tuples of 0-based (i, j)"""
    n = len(seq)
    memo = {}

    def rec(i, j):           # structures on the closed interval [i, j]
        if j - i < TURN + 1:
            return [()]
        key = (i, j)
        if key in memo:
            return memo[key]
        out = list(rec(i + 1, j))                      # i unpaired
        for k in range(i + TURN + 1, j + 1):           # i pairs with k
            if (seq[i], seq[k]) in PAIRS:
                inner = rec(i + 1, k - 1)
                rest = rec(k + 1, j)
                out += [((i, k),) + a + b for a in inner for b in rest]
        memo[key] = out
        return out

    return rec(0, n - 1)


def smooth(x):
    """SMOOTH() of libRNA's params.c (SCALE = 10), x in dcal/mol"""
    if x / 10. < -1.2283697:
        return 0.
    if x / 10. > 0.8660254:
        return x
    return 10. * 0.38490018 * (math.sin(x / 10. - 0.34242663) + 1.) ** 2


_DANGLES = {}


def dangle_tables(par_path):
    """{'dangle5': rows CG GC GU UG AU UA NS x columns N A C G U, 'dangle3': ...} read from the parameter file"""
    if par_path not in _DANGLES:
        tabs, cur = {}, None
        for line in open(par_path):
            t = line.strip()
            if t.startswith("#"):
                cur = t[1:].strip() if t[1:].strip() in ("dangle5", "dangle3") else None
                if cur:
                    tabs[cur] = []
            elif cur and t and not t.startswith("/*"):
                tabs[cur].append([int(v) for v in t.split()])
        _DANGLES[par_path] = tabs
    return _DANGLES[par_path]


PAIR_TYPE = {"CG": 0, "GC": 1, "GU": 2, "UG": 3, "AU": 4, "UA": 5}
BASE_COL = {"A": 1, "C": 2, "G": 3, "U": 4}


def pf_energy(oracle, seq, pairs):
    """energy of the structure in the partition function's model (dcal/mol, float): vrna_eval_structure + the smoothing of the
    single dangle of every exterior stem that touches exactly one sequence end"""
    n = len(seq)
    S = seq.upper().replace("T", "U")
    e = float(oracle.eval_structure(seq, dot_bracket(n, pairs)))
    tabs = dangle_tables(oracle.par_path)
    for i, j in pairs:
        if any(a < i and j < b for a, b in pairs):
            continue                      # enclosed: not an exterior stem
        row = PAIR_TYPE[S[i] + S[j]]
        d = None
        if i > 0 and j == n - 1:
            d = tabs["dangle5"][row][BASE_COL[S[i - 1]]]
        elif i == 0 and j < n - 1:
            d = tabs["dangle3"][row][BASE_COL[S[j + 1]]]
        if d is not None:
            e += -smooth(-d) - d
    return e


def dot_bracket(n, pairs):
    s = ["."] * n
    for i, j in pairs:
        s[i], s[j] = "(", ")"
    return "".join(s)


def satisfies(pairs, cst):
    if cst is None:
        return True
    paired = {}
    for i, j in pairs:
        paired[i] = j
        paired[j] = i
    stack = []
    for p, ch in enumerate(cst):
        if ch == "x" and p in paired:
            return False
        if ch == "(":
            stack.append(p)
        elif ch == ")":
            q = stack.pop()
            if paired.get(q) != p:
                return False
    return True


def motif_pairs(fold):
    stack, out = [], []
    for p, ch in enumerate(fold):
        if ch == "(":
            stack.append(p)
        elif ch == ")":
            out.append((stack.pop(), p))
    return out


def enumerate_ensemble(oracle, seq, cst=None, motif=None):
    """-> (dG kcal/mol, P as dict (i,j)->prob 0-based, minimum energy dcal/mol) by explicit summation"""
    n = len(seq)
    kT_dcal = oracle.kT * 100.
    occ, m_c, inner = [], 0, []
    if motif is not None:
        mseq, mfold, _ = motif
        m = len(mseq)
        occ = [p for p in range(n - m + 1) if seq[p:p + m].upper().replace("T", "U") == mseq]
        m_c = oracle.motif_contribution(motif)
        inner = [pq for pq in motif_pairs(mfold) if pq != (0, m - 1)]
    states = []
    for pairs in all_structures(seq.upper().replace("T", "U")):
        if not satisfies(pairs, cst):
            continue
        e_int = oracle.eval_structure(seq, dot_bracket(n, pairs))
        e = pf_energy(oracle, seq, pairs)
        shown = set(pairs)
        for p in occ:
            q = p + len(motif[0]) - 1
            if (p, q) in shown and not any(p < i < q for i, _ in pairs):      # the occurrence closes an EMPTY hairpin: ligand bound
                e += m_c
                e_int += m_c
                shown |= {(p + a, p + b) for a, b in inner}
        states.append((e, shown, e_int))
    if not states:
        return math.inf, {}, None
    e_min = min(e for e, _, _ in states)
    wts = np.array([math.exp(-(e - e_min) / kT_dcal) for e, _, _ in states])
    Z = wts.sum()
    P = {}
    for (e, shown, _), wt in zip(states, wts):
        for ij in shown:
            P[ij] = P.get(ij, 0.) + wt / Z
    dG = (e_min - kT_dcal * math.log(Z)) / 100.
    return dG, P, min(e_int for _, _, e_int in states)


def check(oracle, seq, cst=None, motif=None, fold=None, tol=1e-10):
    """`fold`: what is checked against the explicit sum -- an object with the oracle's pf / mfe signature (default: the oracle;
    the host emulation of the kernel source and, on the GPU, the engine go through the same check)"""
    fold = fold or oracle
    n = len(seq)
    dG, P, e_min = enumerate_ensemble(oracle, seq, cst, motif)
    g, bp = fold.pf(seq, cst=cst, motif=motif, bppm=True)
    if e_min is None:
        assert math.isinf(g) and g > 0
        return 0
    assert g == pytest.approx(dG, rel=tol, abs=tol), (seq, cst)
    for i in range(n):
        for j in range(i + 1, n):
            assert bp[tri(n, i, j)] == pytest.approx(P.get((i, j), 0.), abs=tol), (seq, cst, i, j)
    e, st = fold.mfe(seq, cst=cst, motif=motif)
    assert e == e_min, (seq, cst, st)
    return len(P)


class _Fast:
    """the fast-path kernel bodies (pf_fast.cuh / mfe_fast.cuh) in host emulation"""

    def __init__(self, emu):
        self.emu = emu

    def pf(self, seq, **kw):
        return self.emu.pf(seq, fast=True, **kw)

    def mfe(self, seq, **kw):
        return self.emu.mfe(seq, fast=True, **kw)


class _Device:
    """the engine on the GPU through the C ABI, one sequence per call"""

    def __init__(self, engine):
        self.engine = engine

    def pf(self, seq, cst=None, motif=None, bppm=True):
        dg, bp = self.engine.pf([seq], cst=cst, motif=motif, bppm=True)
        return float(dg[0]), bp[0]

    def mfe(self, seq, cst=None, motif=None):
        e, st = self.engine.mfe([seq], cst=cst, motif=motif)
        return int(e[0]), st[0]


SEQS = [
    "GGGAAAUCCC", "GCGCAAAAGCGC", "GGGGAAAACCCC", "GUGUGAAAUGCGC", "ACGUGAAAACGU", "GGCGAAAGCCUUUGG",
    "GCUUCGGCAAGCUUCGGC", "GGGAAACCCAGGGAAACCC", "GGACGUUUUCGUCAAGACCUUUUGGUCC"[:22], "UUUUUUUUUUUU", "CCCCCCCCCC",
]


@pytest.mark.parametrize("seq", SEQS)
def test_plain_ensemble_equals_the_explicit_sum(oracle, seq):
    check(oracle, seq)


def test_random_sequences_equal_the_explicit_sum(oracle):
    rng = np.random.Generator(np.random.PCG64(20161019))
    n_pairs = 0
    for alphabet, lengths in (("ACGU", (9, 14, 18, 21)), ("GC", (10, 13, 16)), ("GU", (11, 15)), ("AU", (12, 17)), ("ACGUT", (16,))):
        for n in lengths:
            for _ in range(3):
                seq = "".join(alphabet[k] for k in rng.integers(0, len(alphabet), n))
                n_pairs += check(oracle, seq)
    assert n_pairs > 200          # the sample did exercise pairs


def test_lower_case_and_multiloop(oracle):
    # three hairpins inside one closing pair: the smallest multiloop; also lower case input (reference src/scoring.cc:24-28)
    seq = "GGGAAACCGGGAAACCCGGAAACCC"[:24]
    check(oracle, "gcGGAAACCgGGAAACCcgc"[:20])
    dG, P, e_min = enumerate_ensemble(oracle, seq)
    g = oracle.pf(seq)
    assert g == pytest.approx(dG, rel=1e-10)


@pytest.mark.parametrize("seq,cst", [
    ("GGGAAAUCCC", "(........)"),
    ("GGGAAAUCCC", "..x......."),
    ("GGGAAAUCCC", "xxxxxxxxxx"),
    ("GCGCAAAAGCGCAAGC", ".(........)....."),
    ("GCGCAAAAGCGCAAGC", "x..........xx..."),
    ("GGGAAACCCAGGGAAACCC", "(.......).(.......)"),
    ("GGGAAACCCAGGGAAACCC", "(.................)"),
    ("GGGAAACCCAGGGAAACCC", ".x......x.........."),
    ("GGCGAAAGCCUUUGG", "((.......)....)"),
    ("GGCGAAAGCCUUUGG", "....xxx........"),
])
def test_constrained_ensemble_equals_the_explicit_sum(oracle, seq, cst):
    check(oracle, seq, cst=cst)


def test_constraint_without_any_structure_gives_an_empty_ensemble(oracle):
    # the enforced pair would close a hairpin of two bases: nothing satisfies it
    check(oracle, "GGGAAAUCCC", cst="...(..)...")


# a short hairpin ligand motif in the reference's (sequence, fold, Kd in uM) form: outer pair + one inner pair
MOTIF = ("GCGAAAAGC", "((.....))", 5.0)


@pytest.mark.parametrize("seq,cst", [
    ("GCGAAAAGC", None),
    ("AAGCGAAAAGCUU", None),
    ("GGGCGAAAAGCCC", None),
    ("GGGCGAAAAGCCCAAGCGAAAAGC", None),          # two occurrences
    ("GGGCGAAAAGCCC", "(...........)"),
    ("GGGCGAAAAGCCC", "....xx......."),
    ("gggcgaaaagccc", None),
    ("GGGCGAAAUGCCC", None),                     # no occurrence: the motif changes nothing
])
def test_ligand_motif_ensemble_equals_the_explicit_sum(oracle, seq, cst):
    check(oracle, seq, cst=cst, motif=MOTIF)
    if "GCGAAAAGC" not in seq.upper():
        assert oracle.pf(seq, cst=cst, motif=MOTIF) == oracle.pf(seq, cst=cst)


def test_holo_follows_from_apo_by_the_site_identity(oracle):
    """Z_holo = Z_apo (1 + (beta - 1) P_h) for ONE occurrence, P_h = probability of the empty hairpin at the site in the apo
    ensemble: the identity the score / Monte-Carlo kernels use to derive the holo ensemble from the apo fill."""
    seq = "GGGCGAAAAGCCCAUGGC"
    n = len(seq)
    kT_dcal = oracle.kT * 100.
    at = seq.index(MOTIF[0])
    q = at + len(MOTIF[0]) - 1
    states = [(pf_energy(oracle, seq, s), s) for s in all_structures(seq)]
    Z = sum(math.exp(-e / kT_dcal) for e, _ in states)
    Zh = sum(math.exp(-e / kT_dcal) for e, s in states if (at, q) in s and not any(at < i < q for i, _ in s))
    beta = math.exp(-oracle.motif_contribution(MOTIF) / kT_dcal)
    g_apo, g_holo = oracle.pf(seq), oracle.pf(seq, motif=MOTIF)
    assert g_holo == pytest.approx(g_apo - oracle.kT * math.log(1. + (beta - 1.) * Zh / Z), rel=1e-10)


# ---- the kernel source itself against the explicit sum: host emulation here, the sm_100a build on the GPU ----------------------
KERNEL_CASES = [
    ("GGGAAAUCCC", None, None),
    ("GCGCAAAAGCGC", None, None),
    ("GGGAAACCCAGGGAAACCC", None, None),
    ("gcGGAAACCgGGAAACCcgc", None, None),
    ("GUGUGAAAUGCGCUU", None, None),
    ("GGGAAACCCAGGGAAACCC", "(.................)", None),
    ("GGCGAAAGCCUUUGG", "....xxx........", None),
    ("GGGAAAUCCC", "...(..)...", None),
    ("GGGCGAAAAGCCC", None, MOTIF),
    ("GGGCGAAAAGCCCAAGCGAAAAGC", None, MOTIF),
    ("GGGCGAAAAGCCC", "(...........)", MOTIF),
]


@pytest.mark.parametrize("seq,cst,motif", KERNEL_CASES)
def test_kernel_source_in_emulation_equals_the_explicit_sum(oracle, emu, seq, cst, motif):
    check(oracle, seq, cst=cst, motif=motif, fold=emu, tol=1e-9)              # general path (pf_core.cuh / mfe_core.cuh)
    check(oracle, seq, cst=cst, motif=motif, fold=_Fast(emu), tol=1e-9)       # fast path (pf_fast.cuh / mfe_fast.cuh)


@pytest.mark.gpu
@pytest.mark.parametrize("seq,cst,motif", KERNEL_CASES)
def test_engine_equals_the_explicit_sum(oracle, engine, seq, cst, motif):
    check(oracle, seq, cst=cst, motif=motif, fold=_Device(engine), tol=1e-9)


@pytest.mark.parametrize("seq,cst,motif", [
    ("GGGAAACCCAGGGAAACCC", "(.................)", None),
    ("GGGAAACCCAGGGAAACCC", ".x......x..........", None),
    ("GGGCGAAAAGCCCAUGGC", "..(.......).......", MOTIF),       # the ligand-binding hairpin itself
    ("GGGCGAAAAGCCCAUGGC", "xx................", MOTIF),
])
def test_macrostate_probability_is_the_share_of_the_satisfying_structures(oracle, seq, cst, motif):
    """RnaFold::macrostate_prob (reference src/scoring.cc:53-71): exp((G_all - G_constrained)/kT) = weight of the structures that
    satisfy the constraint / weight of all structures"""
    g_all, _, _ = enumerate_ensemble(oracle, seq, None, motif)
    g_cst, _, _ = enumerate_ensemble(oracle, seq, cst, motif)
    want = math.exp((g_all - g_cst) / oracle.kT)
    assert 0. < want < 1.
    assert oracle.macrostate_prob(seq, cst, motif=motif) == pytest.approx(want, rel=1e-9)
