#!/usr/bin/env python3
"""Short wall-clock timings of every entry point on one GPU, for tuning runs (ADDAPT_B200_LIB selects a variant build made with
tools/variant_build.sh; tools/ab.sh runs this for several builds).  Numbers for the record come from bench.py, not from here.

   python tools/quick_time.py [label] [--mc] [--phase]

fold_eval / inside / mfe: 8,192 (16,384 for MFE) templated 100-nt designs.  --mc: is a Monte-Carlo step as fast as a batch score of the
same sequences?  (With the trajectory's scalars in registers it was 29 % slower: they were spilled inside the fold.)  --phase: one
296-construct batch for a -DADP_PHASE_PROFILE build, whose kernel prints clock64() cycles per phase of the column loop."""
import sys
import time

sys.path.insert(0, '.')
import torch  # noqa: E402
import addapt_b200  # noqa: E402
from tests import cases  # noqa: E402

label = next((a for a in sys.argv[1:] if not a.startswith("--")), "")
eng = addapt_b200.Engine(0)


def timed(fn, reps=3):
    for _ in range(reps):
        torch.cuda.synchronize()
        t0 = time.time()
        fn()
        torch.cuda.synchronize()
        dt = time.time() - t0
    return dt


if "--phase" in sys.argv:
    B = 296
    a = torch.from_numpy(addapt_b200.capi.seqs_to_array(cases.templated_designs(B, 1, n=100))).cuda()
    dg = torch.zeros(B, 2, dtype=torch.float64, device='cuda')
    eng.fold_eval_raw(B, 100, a, cases.THEO, 0, None, None, dg, None)
    torch.cuda.synchronize()
    sys.exit(0)

B = 8192
a = torch.from_numpy(addapt_b200.capi.seqs_to_array(cases.templated_designs(B, 1, n=100))).cuda()
dg = torch.zeros(B, 2, dtype=torch.float64, device='cuda')
d1 = torch.zeros(B, dtype=torch.float64, device='cuda')
t_fe = timed(lambda: eng.fold_eval_raw(B, 100, a, cases.THEO, 0, None, None, dg, None))
t_in = timed(lambda: eng.pf_raw(B, 100, a, d1))
print("%s fold_eval %.0f /s   inside %.0f fills/s   checksum %.9f" % (label, B / t_fe, B / t_in, float(dg.sum())))
Bm = 16384
am = torch.from_numpy(addapt_b200.capi.seqs_to_array(cases.templated_designs(Bm, 1, n=100))).cuda()
e = torch.zeros(Bm, dtype=torch.int32, device='cuda')
st = torch.zeros(Bm, 100, dtype=torch.uint8, device='cuda')
t_mfe = timed(lambda: eng.mfe_raw(Bm, 100, am, e, st), reps=4)
print("%s mfe %.0f folds/s  checksum %d" % (label, Bm / t_mfe, int(e.sum())))

if "--mc" in sys.argv:
    pr = cases.rhf6_problem(n=100)
    T, S = 8192, 100
    fs = torch.zeros(T, 100, dtype=torch.uint8, device='cuda')
    bs = torch.zeros_like(fs)
    oc = torch.zeros(T, 4, dtype=torch.int32, device='cuda')
    t_mc = timed(lambda: eng.mc_run_raw(pr, T, 2016, 0, S, ("annealing", 5.0, 0.0, 300), final_seq=fs, best_seq=bs, outcomes=oc), reps=2)
    scored = int(oc[:, [0, 1, 3]].sum()) + T
    print("%s mc: %.0f steps/s, %.0f scored proposals/s (%d scored of %d)" % (label, T * S / t_mc, scored / t_mc, scored, T * S))
    sc = torch.zeros(T, dtype=torch.float64, device='cuda')
    for name, seqs in (("final sequences of the run", fs), ("templated designs", a)):
        print("%s score, %s: %.0f scores/s" % (label, name, T / timed(lambda: eng.score_raw(pr, T, seqs, sc), reps=2)))
