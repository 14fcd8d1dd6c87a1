"""Shared test inputs: the sequences, constraints and the design problem of the reference's own tests
(reference tests/test_scoring.cc:45-63,83-259) plus seeded synthetic batches."""
import numpy as np

THEO = ("GAUACCAGCCGAAAGGCCCUUGGCAGC", "(...((.(((....)))....))...)", 0.32)
HAIRPIN = "ACGUGAAAACGU"
RHF6 = "guuuuagagcuagaaauagcaaguuaaaauaaggcuaguccCuUUUCGCCgauaccagccgaaaggcccuuggcagcGACggcaccgagucggugcuuuuuu"
RHF6_ACTIVE = "(............................)xx..xxxxx..xxxxxx(...............................).(.............)......"
# the five spacer contexts of the reference's tools/test_seq:10-16
SPACERS = [("", ""), ("AACTTTCAGTTTAGCGGTCT", ""), ("CATCTAATTCAACAAGAATT", ""), ("GGGGCCACTAGGGACAGGAT", ""), ("GGGTGGGGGGAGTTTGCTCC", "")]


def rhf6_problem(contexts=None, float_dg=False, n=102):
    """default objective of the reference: log(1 - P_apo(active)) + log(P_holo(active)) (src/config.cc:77-81)"""
    from addapt_b200.capi import Problem
    return Problem(RHF6[:n], {"active": RHF6_ACTIVE[:n]}, [("active", "apo", False, 1.0), ("active", "holo", True, 1.0)],
                   motif=THEO, contexts=contexts, float_dg=float_dg)


def random_seqs(B, n, seed):
    rng = np.random.Generator(np.random.PCG64(seed))
    a = rng.integers(0, 4, (B, n))
    return ["".join("ACGU"[k] for k in row) for row in a]


def templated_designs(B, seed, n=100):
    """BASELINE config 3: rhf(6) trimmed to n nt; the 10 free positions resampled, the slaved partner complemented."""
    rng = np.random.Generator(np.random.PCG64(seed))
    base = list(RHF6[:n].upper())
    free = [41, 43, 44, 45, 46, 47, 48, 49, 77, 78]
    comp = {"A": "U", "U": "A", "G": "C", "C": "G"}
    out = []
    for _ in range(B):
        s = list(base)
        for p in free:
            s[p] = "ACGU"[int(rng.integers(0, 4))]
        s[79] = comp[s[47]]
        out.append("".join(s))
    return out


def random_constraint(oracle, seq, rng, p_keep=0.3, p_x=0.1):
    """a satisfiable-looking constraint: some pairs of the MFE structure enforced, some free bases forbidden to pair"""
    _, st = oracle.mfe(seq)
    c = ["."] * len(seq)
    stack = []
    for i, ch in enumerate(st):
        if ch == "(":
            stack.append(i)
        elif ch == ")":
            k = stack.pop()
            if rng.random() < p_keep:
                c[k], c[i] = "(", ")"
    for i in range(len(seq)):
        if c[i] == "." and rng.random() < p_x:
            c[i] = "x"
    return "".join(c)


def rnafold_symbols(bppm_tri, n):
    """The per-position pairing string RNAfold -p prints (ViennaRNA bppm_symbol), derived from an upper-triangle bppm in the
    layout base_pair_prob reads (reference src/scoring.cc:49-50; tri(i,j) = i*n - i*(i-1)/2 + (j-i), 0-based)."""
    P = np.zeros((n, n))
    iu = np.triu_indices(n)
    P[iu] = np.asarray(bppm_tri)
    P[np.diag_indices(n)] = 0.0
    out = []
    for k in range(n):
        x1, x2 = P[k, k + 1:].sum(), P[:k, k].sum()
        x0 = 1.0 - x1 - x2
        if x0 > 0.667:
            c = "."
        elif x1 > 0.667:
            c = "("
        elif x2 > 0.667:
            c = ")"
        elif x1 + x2 > x0:
            c = "{" if x1 / (x1 + x2) > 0.667 else "}" if x2 / (x1 + x2) > 0.667 else "|"
        else:
            c = "," if x0 > x1 + x2 else ":"
        out.append(c)
    return "".join(out)
