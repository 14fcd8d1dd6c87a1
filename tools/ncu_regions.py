#!/usr/bin/env python3
"""Executed instructions and stall samples of pf_fast_kernel<256,true,false> (--kernel=mfe: mfe_fast_kernel<256,true>) by source region, from one
`ncu --set full --import-source on` capture of the CURRENT build (the SASS of the report is matched instruction by instruction with
nvdisasm -gi of addapt_b200/libaddapt_b200.so, which carries the inlining chains).   python tools/ncu_regions.py report.ncu-rep"""
import collections, csv, io, os, re, subprocess, sys, tempfile
_pos = [a for a in sys.argv[1:] if not a.startswith("--")]
rep = _pos[0]
lib = _pos[1] if len(_pos) > 1 else "addapt_b200/libaddapt_b200.so"
exec(open(os.path.join(os.path.dirname(__file__), "code_size.py")).read().split("agg = collections.Counter()")[0].replace('lib = args[0] if args else "addapt_b200/libaddapt_b200.so"', "pass"))
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[2:] if len(r) > 10]
assert len(data) == len(instrs), (len(data), len(instrs))
agg = collections.defaultdict(lambda: [0, 0, 0, collections.Counter()])
STALLS = ["stall_barrier", "stall_wait", "stall_short_sb", "stall_long_sb", "stall_no_inst", "stall_branch_resolving", "stall_math", "stall_not_selected", "stall_selected", "stall_dispatch", "stall_lg", "stall_mio"]
for c, r in zip(instrs, data):
    best = next((l for f, l in c if f == src_name and l >= ins_a), None)
    a = agg[rname(best) if best else "other"]
    a[0] += int(r[ix["Instructions Executed"]] or 0)
    a[1] += int(r[ix["Thread Instructions Executed"]] or 0)
    a[2] += int(r[ix["# Samples"]] or 0)
    for s in STALLS:
        if s in ix:
            a[3][s] += int(r[ix[s]] or 0)
ti = sum(v[0] for v in agg.values()); ts = sum(v[2] for v in agg.values())
print("%-20s %7s %7s %6s  top stall reasons (%% of the region's samples)" % ("region", "inst%", "samp%", "lanes"))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][2]):
    top = ", ".join("%s %.0f" % (s.replace("stall_", ""), 100 * n / max(v[2], 1)) for s, n in v[3].most_common(4))
    print("%-20s %7.2f %7.2f %6.1f  %s" % (k, 100 * v[0] / ti, 100 * v[2] / ts, v[1] / max(v[0], 1), top))
if len(_pos) > 2:
    pat = _pos[2]
    print("\ntop instructions of region", pat)
    out = []
    for (c, r), t in zip(zip(instrs, data), [r[1] for r in data]):
        best = next((l for f, l in c if f == src_name and l >= ins_a), None)
        if (rname(best) if best else "other") == pat:
            out.append((int(r[ix["# Samples"]] or 0), t.strip()[:50], " < ".join("%s:%d" % (f.split(".")[0], l) for f, l in c)[:110], r[ix["stall_barrier"]]))
    for s_, t, cs, b in sorted(out, key=lambda x: -x[0])[:14]:
        print("%6.2f%% %-50s bar=%s | %s" % (100 * s_ / ts, t, b, cs))
