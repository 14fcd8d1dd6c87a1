#!/usr/bin/env python3
"""Write params/rna_turner2004_reconstructed.par (ViennaRNA "parameter file v2.0" layout).

PROVENANCE -- read this before trusting a number
------------------------------------------------
The reference (kalekundert/addapt, src/scoring.cc:81-88) gets its energy model from
ViennaRNA's built-in Turner-2004 set (`vrna_md_set_default` + `vrna_fold_compound`).
libRNA and its `rna_turner2004.par` are NOT available in this image (no network), so
this script RECONSTRUCTS that parameter set:

* scalars, loop-length arrays, stacking, dangles, the six mismatch families, special
  hairpins: written from the published Turner-2004 / ViennaRNA 2.x values as recalled;
  they reproduce the energies the reference's own tests quote in comments
  (tests/test_scoring.cc:87  ACGUGAAAACGU ((((....)))) = -2.20 kcal/mol;
   tests/test_scoring.cc:153 aptamer ....((((((....)))...))).... = -6.20 kcal/mol).
* int11 / int21 / int22 (1x1, 1x2, 2x2 interior loops; 900+4500+9216 numbers in the
  real file): generated from the nearest-neighbour RULES below, not from the measured
  tables.  These are the entries most likely to differ from the genuine file.

The engine and the oracle both load parameters from a `.par` file at run time, so a
genuine `rna_turner2004.par` can be dropped in (adp_engine_create(par_file=...)) with
no code change.  Every results file repeats: exact parity with libRNA is UNPINNED.
"""
import os

INF = 1000000
PAIRS = ["CG", "GC", "GU", "UG", "AU", "UA"]   # ViennaRNA pair types 1..6
NUC = "NACGU"                                  # 0..4

stack = [
    [-240, -330, -210, -140, -210, -210],
    [-330, -340, -250, -150, -220, -240],
    [-210, -250,  130,  -50, -140, -130],
    [-140, -150,  -50,   30,  -60, -100],
    [-210, -220, -140,  -60, -110,  -90],
    [-210, -240, -130, -100,  -90, -130],
]

hairpin = [INF, INF, INF, 540, 560, 570, 540, 600, 550, 640, 650, 660, 670, 680, 690, 690,
           700, 710, 710, 720, 720, 730, 730, 740, 740, 750, 750, 750, 760, 760, 770]
bulge = [INF, 380, 280, 320, 360, 400, 440, 459, 470, 480, 490, 500, 510, 520, 530, 540,
         540, 550, 550, 560, 570, 570, 580, 580, 580, 590, 590, 600, 600, 600, 610]
interior = [INF, INF, 100, 100, 110, 200, 200, 210, 230, 240, 250, 260, 270, 280, 290, 290,
            300, 310, 310, 320, 330, 330, 340, 340, 350, 350, 350, 360, 360, 370, 370]

mismatch_hairpin = {
    "CG": [[-80, -100, -110, -100, -80], [-140, -150, -150, -140, -150], [-80, -100, -110, -100, -80],
           [-150, -230, -150, -240, -150], [-100, -100, -140, -100, -210]],
    "GC": [[-50, -110, -70, -110, -50], [-110, -110, -150, -130, -150], [-50, -110, -70, -110, -50],
           [-150, -250, -150, -220, -150], [-100, -110, -100, -110, -160]],
    "GU": [[20, 20, -20, -10, -20], [20, 20, -50, -30, -50], [-10, -10, -20, -10, -20],
           [-50, -100, -50, -110, -50], [-10, -10, -30, -10, -100]],
    "UG": [[0, -20, -10, -20, 0], [-30, -50, -30, -60, -30], [0, -20, -10, -20, 0],
           [-30, -90, -30, -110, -30], [-10, -20, -10, -20, -90]],
    "AU": [[-10, -10, -20, -10, -20], [-30, -30, -50, -30, -50], [-10, -10, -20, -10, -20],
           [-30, -120, -30, -110, -30], [-10, -10, -30, -10, -100]],
    "UA": [[0, -20, -10, -20, 0], [-30, -50, -30, -50, -30], [0, -20, -10, -20, 0],
           [-30, -150, -30, -150, -30], [-10, -20, -10, -20, -80]],
}


def mm_bonus_table(strong, ag, ga, gg, uu, weak_pen=70):
    """5x5 block: rows = 5' mismatch base (N,A,C,G,U), cols = 3' mismatch base."""
    base = 0 if strong else weak_pen
    t = [[base] * 5 for _ in range(5)]
    t[1][3] = base + ag
    t[3][1] = base + ga
    t[3][3] = base + gg
    t[4][4] = base + uu
    return t


mismatch_interior = {p: mm_bonus_table(p in ("CG", "GC"), -80, -100, -100, -60) for p in PAIRS}
mismatch_interior_1n = {p: mm_bonus_table(p in ("CG", "GC"), 0, 0, 0, 0) for p in PAIRS}
mismatch_interior_23 = {
    "CG": mm_bonus_table(True, -50, -110, -70, -30),
    "GC": mm_bonus_table(True, 0, -120, -70, -30),
}
for p in ("GU", "UG", "AU", "UA"):
    mismatch_interior_23[p] = mm_bonus_table(False, 0, -110, -70, -30)

mismatch_multi = {
    "CG": [[-50, -110, -50, -140, -70], [-110, -110, -110, -160, -110], [-70, -150, -70, -150, -100],
           [-110, -130, -110, -140, -110], [-50, -150, -50, -150, -70]],
    "GC": [[-80, -140, -80, -140, -100], [-100, -150, -100, -140, -100], [-110, -150, -110, -150, -140],
           [-100, -140, -100, -160, -100], [-80, -150, -80, -150, -120]],
    "GU": [[-50, -80, -50, -50, -50], [-50, -100, -70, -50, -70], [-60, -80, -60, -80, -60],
           [-70, -110, -70, -80, -70], [-50, -80, -50, -80, -50]],
    "UG": [[-30, -30, -60, -60, -60], [-30, -30, -60, -60, -60], [-70, -100, -70, -100, -80],
           [-60, -80, -60, -80, -60], [-60, -80, -60, -80, -60]],
    "AU": [[-50, -80, -50, -80, -50], [-70, -100, -70, -110, -70], [-60, -80, -60, -80, -60],
           [-70, -110, -70, -120, -70], [-50, -80, -50, -80, -50]],
    "UA": [[-60, -80, -60, -80, -60], [-60, -80, -60, -80, -60], [-70, -100, -70, -100, -80],
           [-60, -80, -60, -80, -60], [-70, -100, -70, -100, -80]],
}
mismatch_exterior = mismatch_multi  # same numbers in the Turner-2004 set as recalled

dangle5 = {  # cols N,A,C,G,U
    "CG": [-10, -50, -30, -20, -10], "GC": [0, -20, -30, 0, 0], "GU": [-20, -30, -30, -40, -20],
    "UG": [-10, -30, -10, -20, -20], "AU": [-20, -30, -30, -40, -20], "UA": [-10, -30, -10, -20, -20],
}
dangle3 = {
    "CG": [-40, -110, -40, -130, -60], "GC": [-80, -170, -80, -170, -120], "GU": [-10, -70, -10, -70, -10],
    "UG": [-50, -80, -50, -80, -60], "AU": [-10, -70, -10, -70, -10], "UA": [-50, -80, -50, -80, -60],
}

tetraloops = [("CAACGG", 550), ("CCAAGG", 330), ("CCACGG", 370), ("CCCAGG", 340), ("CCGAGG", 350),
              ("CCGCGG", 360), ("CCUAGG", 370), ("CCUCGG", 250), ("CUAAGG", 360), ("CUACGG", 280),
              ("CUCAGG", 370), ("CUCCGG", 270), ("CUGCGG", 280), ("CUUAGG", 350), ("CUUCGG", 370),
              ("CUUUGG", 370)]
triloops = [("CAACG", 680), ("GUUAC", 690)]
hexaloops = [("ACAGUACU", 280), ("ACAGUGAU", 360), ("ACAGUGCU", 290), ("ACAGUGUU", 180)]

ML_BASE, ML_CLOSING, ML_INTERN = 0, 930, -90
NINIO, MAX_NINIO = 60, 300
DUPLEX_INIT, TERMINAL_AU, LXC = 410, 50, 107.856


def weak(p):
    return p not in ("CG", "GC")


def gen_int11():
    """RULE: 50 + 70 per AU/GU closure; G.G mismatch -260 (CG..CG measured exception -140);
    U.U mismatch -70 (strong/strong: 40); A.A in strong/strong: 90."""
    out = {}
    for p1 in PAIRS:
        for p2 in PAIRS:
            base = 50 + 70 * weak(p1) + 70 * weak(p2)
            t = [[base] * 5 for _ in range(5)]
            nweak = weak(p1) + weak(p2)
            t[3][3] = base - 260
            t[4][4] = base - 70
            if nweak == 0:
                t[1][1] = 90
                t[0][0] = t[0][1] = t[1][0] = 90
                t[4][4] = 40
                if p1 == "CG" and p2 == "CG":
                    t[3][3] = -140
                elif p1 == "GC" and p2 == "GC":
                    t[3][3] = -230
                else:
                    t[3][3] = -220
            out[(p1, p2)] = t
    return out


def gen_int21():
    """RULE: 230 + 70 per AU/GU closure; G.G across the loop (5' single base vs 3'-most base) -120."""
    out = {}
    for p1 in PAIRS:
        for p2 in PAIRS:
            base = 230 + 70 * weak(p1) + 70 * weak(p2)
            blocks = []
            for a in range(5):          # S[i+1]
                t = [[base] * 5 for _ in range(5)]   # [b = S[l+1]][c = S[j-1]]
                if a == 3:
                    for b in range(5):
                        t[b][3] = base - 120
                blocks.append(t)
            out[(p1, p2)] = blocks
    return out


def gen_int22():
    """RULE: generic interior-loop formula applied to 2x2: interior[4] + mismatch_interior of both
    closing pairs (the genuine file holds measured values here)."""
    out = {}
    for p1 in PAIRS:
        for p2 in PAIRS:
            vals = []
            for a in range(1, 5):            # S[i+1]
                for b in range(1, 5):        # S[k-1]
                    for c in range(1, 5):    # S[l+1]
                        for d in range(1, 5):  # S[j-1]
                            e = interior[4] + mismatch_interior[p1][a][d] + mismatch_interior[p2][c][b]
                            vals.append(e)
            out[(p1, p2)] = vals
    return out


def fmt(v):
    return "   INF" if v >= INF else "%6d" % v


def ns_block(blocks):
    """'@' (non-standard pair) block: element-wise max over the six real pair types."""
    return [[max(b[r][c] for b in blocks) for c in range(5)] for r in range(5)]


def write(path):
    L = []
    A = L.append
    A("## RNAfold parameter file v2.0")
    A("")
    A("/* RECONSTRUCTION of the Turner-2004 set (see params/gen_turner2004_par.py).   */")
    A("/* int11/int21/int22 are RULE-GENERATED, not the measured tables.              */")
    A("/* Only the 37 C free energies are given; the engine folds at 37 C only.       */")
    A("")
    A("# stack")
    A("/*  CG     GC     GU     UG     AU     UA     @  */")
    rows = [r + [max(r)] for r in stack]
    rows.append([max(rows[i][j] for i in range(6)) for j in range(7)])
    for r in rows:
        A(" ".join(fmt(v) for v in r))
    A("")

    def mm_section(name, tab):
        A("# " + name)
        blocks = [tab[p] for p in PAIRS]
        for p, b in zip(PAIRS + ["@"], blocks + [ns_block(blocks)]):
            A("/* %s */" % p)
            for r in b:
                A(" ".join(fmt(v) for v in r))
        A("")

    mm_section("mismatch_hairpin", mismatch_hairpin)
    mm_section("mismatch_interior", mismatch_interior)
    mm_section("mismatch_interior_1n", mismatch_interior_1n)
    mm_section("mismatch_interior_23", mismatch_interior_23)
    mm_section("mismatch_multi", mismatch_multi)
    mm_section("mismatch_exterior", mismatch_exterior)

    for name, tab in (("dangle5", dangle5), ("dangle3", dangle3)):
        A("# " + name)
        A("/*  N      A      C      G      U  */")
        rows = [tab[p] for p in PAIRS]
        rows.append([max(r[c] for r in rows) for c in range(5)])
        for r in rows:
            A(" ".join(fmt(v) for v in r))
        A("")

    i11 = gen_int11()
    A("# int11")
    allp = PAIRS + ["@"]
    for p1 in allp:
        for p2 in allp:
            A("/* %s..%s */" % (p1, p2))
            if p1 == "@" or p2 == "@":
                q1 = "AU" if p1 == "@" else p1
                q2 = "AU" if p2 == "@" else p2
                t = i11[(q1, q2)]
            else:
                t = i11[(p1, p2)]
            for r in t:
                A(" ".join(fmt(v) for v in r))
    A("")

    i21 = gen_int21()
    A("# int21")
    for p1 in allp:
        for p2 in allp:
            q1 = "AU" if p1 == "@" else p1
            q2 = "AU" if p2 == "@" else p2
            for a in range(5):
                A("/* %s.%s..%s */" % (p1, NUC[a], p2))
                for r in i21[(q1, q2)][a]:
                    A(" ".join(fmt(v) for v in r))
    A("")

    i22 = gen_int22()
    A("# int22")
    for p1 in PAIRS:
        for p2 in PAIRS:
            vals = i22[(p1, p2)]
            idx = 0
            for a in range(1, 5):
                for b in range(1, 5):
                    A("/* %s.%s%s..%s */" % (p1, NUC[a], NUC[b], p2))
                    for c in range(4):
                        A(" ".join(fmt(v) for v in vals[idx:idx + 4]))
                        idx += 4
    A("")

    for name, arr in (("hairpin", hairpin), ("bulge", bulge), ("interior", interior)):
        A("# " + name)
        for i in range(0, 31, 10):
            A(" ".join(fmt(v) for v in arr[i:i + 10]))
        A("")

    A("# ML_params")
    A("/* F = cu*n_unpaired + cc + ci*loop_degree (+TermAU) */")
    A("/*      cu      cu_dH        cc      cc_dH        ci      ci_dH  */")
    A("%6d %6d %6d %6d %6d %6d" % (ML_BASE, 0, ML_CLOSING, 3000, ML_INTERN, -220))
    A("")
    A("# NINIO")
    A("/* Ninio = MIN(max, m*|n1-n2| */")
    A("/*       m    m_dH     max  */")
    A("%6d %6d %6d" % (NINIO, 320, MAX_NINIO))
    A("")
    A("# Misc")
    A("/* all parameters are pairs of 'energy enthalpy' */")
    A("/*    DuplexInit     TerminalAU   LXC  */")
    A("%6d %6d %6d %6d %10.6f %6d" % (DUPLEX_INIT, 360, TERMINAL_AU, 370, LXC, 0))
    A("")
    A("# Hexaloops")
    for s, e in hexaloops:
        A("%s %6d %6d" % (s, e, 0))
    A("")
    A("# Tetraloops")
    for s, e in tetraloops:
        A("%s %6d %6d" % (s, e, 0))
    A("")
    A("# Triloops")
    for s, e in triloops:
        A("%s %6d %6d" % (s, e, 0))
    A("")
    A("#END")
    with open(path, "w") as f:
        f.write("\n".join(L) + "\n")


if __name__ == "__main__":
    here = os.path.dirname(os.path.abspath(__file__))
    write(os.path.join(here, "rna_turner2004_reconstructed.par"))
    print("wrote", os.path.join(here, "rna_turner2004_reconstructed.par"))
