#!/usr/bin/env python3
"""Attribute the oracle-vs-reference residual on rhf(6) to named Turner-table entries (CPU only; oracle = test infrastructure).

The reference holds two ensemble free energies and two RNAfold -p strings for rhf(6) (reference tests/test_scoring.cc:54-55).
d(ensemble dG)/d(table entry) = the expected number of loops of the ensemble that use the entry, so a finite difference per
entry of the parameter file lists exactly the entries the two numbers depend on.  Entries are named by section, the comment
line of their block (the closing pairs) and row/column inside the block.

   python tools/param_sensitivity.py [par_file] [--min 0.02] [--sections int11,int21,int22,...]
"""
import argparse
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.oracle import Oracle, DEFAULT_PAR  # noqa: E402
from tests import cases  # noqa: E402

DELTA = 10  # dcal/mol = 0.1 kcal/mol


def entries(lines):
    """yield (line index, token index, section, block comment, row in block) for every finite numeric entry"""
    section, block, row = None, "", 0
    for li, line in enumerate(lines):
        s = line.strip()
        if s.startswith("# "):
            section, block, row = s[2:].strip(), "", 0
            continue
        if s.startswith("/*"):
            txt = s[2:-2].strip() if s.endswith("*/") else ""
            if txt and " " not in txt and len(txt) <= 14:      # a block label such as CG..AU, CG.A..GC, CG.AC..UA
                block, row = txt, 0
            continue
        if not s or section is None or s.startswith("#"):
            continue
        toks = line.split()
        numeric = []
        for ti, t in enumerate(toks):
            try:
                int(t)
                numeric.append(ti)
            except ValueError:
                pass
        if not numeric:
            continue
        for ti in numeric:
            yield li, ti, section, block, row
        row += 1


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("par", nargs="?", default=os.environ.get("ADDAPT_B200_PAR", DEFAULT_PAR))
    ap.add_argument("--min", type=float, default=0.02, help="list entries with |d dG / d entry| at least this (expected uses)")
    ap.add_argument("--sections", default="stack,mismatch_hairpin,mismatch_interior,mismatch_interior_1n,mismatch_interior_23,"
                    "mismatch_multi,mismatch_exterior,dangle5,dangle3,int11,int21,int22,hairpin,bulge,interior,ML_params,NINIO,Misc,"
                    "Tetraloops,Triloops,Hexaloops")
    a = ap.parse_args()
    want = set(a.sections.split(","))
    lines = open(a.par).read().split("\n")
    seq = cases.RHF6.upper()
    base = Oracle(a.par)
    g0a, g0h = base.pf(seq), base.pf(seq, motif=cases.THEO)
    print("# par file: %s" % a.par)
    print("# rhf(6) ensemble dG: apo %.4f (reference comment -29.58), holo %.4f (reference comment -33.82)" % (g0a, g0h))
    out = []
    tmp = tempfile.NamedTemporaryFile("w", suffix=".par", delete=False)
    tmp.close()
    seen = 0
    for li, ti, section, block, row in entries(lines):
        if section not in want:
            continue
        toks = lines[li].split()
        v = int(toks[ti])
        if abs(v) >= 100000:           # INF
            continue
        if section in ("Tetraloops", "Triloops", "Hexaloops") and ti != 1:
            continue
        toks2 = list(toks)
        toks2[ti] = str(v + DELTA)
        mod = list(lines)
        mod[li] = " ".join(toks2)
        with open(tmp.name, "w") as f:
            f.write("\n".join(mod))
        o = Oracle(tmp.name)
        da = (o.pf(seq) - g0a) / (DELTA / 100.0)
        dh = (o.pf(seq, motif=cases.THEO) - g0h) / (DELTA / 100.0)
        seen += 1
        if max(abs(da), abs(dh)) >= a.min:
            name = toks[0] if section in ("Tetraloops", "Triloops", "Hexaloops") else "%s row %d col %d" % (block, row, ti)
            out.append((max(abs(da), abs(dh)), section, name, v, da, dh))
    os.unlink(tmp.name)
    out.sort(reverse=True)
    print("# %d entries perturbed by +%.2f kcal/mol; listed: expected uses >= %.3f in either ensemble" % (seen, DELTA / 100.0, a.min))
    print("%-22s %-34s %7s %9s %9s" % ("section", "entry", "value", "d_apo", "d_holo"))
    for _, section, name, v, da, dh in out:
        print("%-22s %-34s %7d %9.4f %9.4f" % (section, name, v, da, dh))


if __name__ == "__main__":
    main()
