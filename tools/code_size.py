#!/usr/bin/env python3
"""Static SASS size of pf_fast_kernel<256,true,false> (or, with --kernel=mfe, mfe_fast_kernel<256,true>) by source region of pf_fast.cuh / mfe_fast.cuh (the instruction caches are small --
L0 ~6 KB, L1.5 32 KB per SM -- so the hot loop bodies are kept compact).   python tools/code_size.py [lib.so]"""
import collections, os, re, subprocess, sys, tempfile
args = [a for a in sys.argv[1:] if not a.startswith("--")]
KERNEL = "mfe" if "--kernel=mfe" in sys.argv else "mc" if "--kernel=mc" in sys.argv else "pf"      # mc: the PF sweep inside mc_kernel<true,1,true,false>
lib = args[0] if args else "addapt_b200/libaddapt_b200.so"
src_path = "addapt_b200/csrc/%s_fast.cuh" % ("mfe" if KERNEL == "mfe" else "pf")
src_name = os.path.basename(src_path)
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "design_kernels" if KERNEL == "mc" else "fold_kernels", os.path.abspath(lib)], cwd=tmp, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
sass = subprocess.run(["nvdisasm", "-gi", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.split("\n")
sec = [i for i, l in enumerate(sass) if re.match(r"\s*\.section\s*\.text\.", l)]
a = [i for i in sec if {"pf": "pf_fast_kernelILi256ELb1ELb0", "mfe": "mfe_fast_kernelILi256ELb1", "mc": "mc_kernelILb1ELi1ELb1ELb0"}[KERNEL] in sass[i]][0]
b = min(i for i in sec + [len(sass)] if i > a)
instrs, chain, pend = [], [], []
for ln in sass[a:b]:
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', ln)
    if m:
        pend.append((m.group(1).split("/")[-1], int(m.group(2))))
        continue
    if re.match(r"\s*/\*[0-9a-f]+\*/\s+.*?;", ln):
        if pend:
            chain, pend = pend, []
        instrs.append(chain)
src = open(src_path).read().split("\n")
def find(sub, start=0):
    for i in range(start, len(src)):
        if sub in src[i]:
            return i + 1
    return None
if KERNEL != "mfe":
    ins_a, out_a = find("ADP_HD double fast_inside("), find("ADP_HD void fast_outside(")
else:
    ins_a, out_a = find("ADP_HD int mfe_fast_fill("), len(src) + 1
MI = [("ADP_HD double fast_inside(", "in.init"), ("auto levels = [&](int jn)", "in.levels"), ("auto list_rows = [&](int jn)", "in.list"), ("auto rest_rows = [&](int j)", "in.rest"),
      ("auto q1_task", "in.q1"), ("auto scan_unpaired", "in.scan"), ("auto drain_qm", "in.qm"),
      ("// j = 4 is a warm-up", "in.colhead"), ("// ---- cells(j): Qb(i,j)", "in.cells.head"), ("// corner candidates through", "in.cells.corners"),
      ("// The four 1-D filters along", "in.cells.axis"), ("// multiloop closed by (i,j)", "in.cells.ml"), ("// ---- cell total", "in.cells.fin"),
      ("// ---- everything else of the phase", "in.tail"), ("// epilogue: the exterior prefix", "in.epilogue")]
MO = [("ADP_HD void fast_outside(", "out.init"), ("auto carriers = [&](int ln)", "out.carriers"), ("auto levels = [&](int ln)", "out.levels"), ("auto list_rows = [&](int ln)", "out.list"),
      ("auto rest_rows = [&](int l)", "out.rest"), ("auto results = [&](int lq)", "out.results"),
      ("// l = n+1 is a warm-up", "out.colhead"), ("// ---- cells(l): outside weight", "out.cells.head"), ("const int t2r = rtype(type);", "out.cells.corners"),
      ("// the four 1-D filters fused into one loop like", "out.cells.axis"), ("// (k,l) is a stem of a multiloop", "out.cells.ml"), ("// ---- cell total", "out.cells.fin"),
      ("// ---- everything else of the phase", "out.tail"), ("// the motif's inner pairs", "out.motif")]
MM = [("ADP_HD int mfe_fast_fill(", "fill.init"), ("auto prepare_column", "fill.levels"), ("      const int prev = fm1_prev[r + 1];", "fill.rows"),
      ("// j = 4 is a warm-up", "fill.colhead"), ("// ---- interval 1: C(i,j)", "fill.cells.head"), ("// corner candidates, software-pipelined", "fill.cells.corners"),
      ("// the four 1-D filters along", "fill.cells.axis"), ("int k = i + 6 + sub;", "fill.cells.ml"), ("other = imin(other, cadd + ctab);", "fill.cells.fin"),
      ("// ---- interval 2: shifted copies", "fill.int2")]
if KERNEL != "mfe":
    marks = sorted([(find(s, ins_a), n) for s, n in MI if find(s, ins_a)] + [(find(s, out_a), n) for s, n in MO if find(s, out_a)])
else:
    marks = sorted([(find(s, ins_a), n) for s, n in MM if find(s, ins_a)])
def rname(l):
    nm = "other"
    for a_, n in marks:
        if a_ <= l:
            nm = n
    return nm
agg = collections.Counter()
for c in instrs:
    best = next((l for f, l in c if f == src_name and l >= ins_a), None)
    agg[rname(best) if best else "other"] += 1
for k, v in agg.most_common():
    print("%-20s %5d instr %6.1f KB" % (k, v, v * 16 / 1024))
print("total %d instr, %.0f KB (both constraint variants)" % (sum(agg.values()), sum(agg.values()) * 16 / 1024))
