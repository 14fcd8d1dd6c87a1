"""Tuning aid: is a Monte-Carlo step slower than a batch score of the same sequences?  (python tools/mc_vs_score.py)"""
import sys, time
sys.path.insert(0, '.')
import numpy as np, torch
import addapt_b200
from tests import cases
eng = addapt_b200.Engine(0)
pr = cases.rhf6_problem(n=100)
T, S = 8192, 100
fs = torch.zeros(T, 100, dtype=torch.uint8, device='cuda'); bs = torch.zeros_like(fs)
oc = torch.zeros(T, 4, dtype=torch.int32, device='cuda')
for rep in range(2):
    torch.cuda.synchronize(); t0 = time.time()
    eng.mc_run_raw(pr, T, 2016, 0, S, ("annealing", 5.0, 0.0, 300), final_seq=fs, best_seq=bs, outcomes=oc)
    torch.cuda.synchronize(); t1 = time.time()
scored = int(oc[:, [0, 1, 3]].sum()) + T
print("mc: %.0f steps/s, %.0f scored proposals/s (%d scored of %d)" % (T * S / (t1 - t0), scored / (t1 - t0), scored, T * S))
sc = torch.zeros(T, dtype=torch.float64, device='cuda')
for name, seqs in (("final sequences of the run", fs), ("templated designs", torch.from_numpy(addapt_b200.capi.seqs_to_array(cases.templated_designs(T, 1, n=100))).cuda()),
                   ("8192 copies of the start sequence", torch.from_numpy(addapt_b200.capi.seqs_to_array([cases.RHF6[:100].upper()] * T)).cuda())):
    for rep in range(2):
        torch.cuda.synchronize(); t0 = time.time()
        eng.score_raw(pr, T, seqs, sc)
        torch.cuda.synchronize(); t1 = time.time()
    print("score, %s: %.0f scores/s" % (name, T / (t1 - t0)))
for steps in (1, 10):
    torch.cuda.synchronize(); t0 = time.time()
    eng.mc_run_raw(pr, T, 2016, 0, steps, ("annealing", 5.0, 0.0, 300), final_seq=fs, best_seq=bs, outcomes=oc)
    torch.cuda.synchronize(); t1 = time.time()
    print("mc %d steps: %.4f s" % (steps, t1 - t0))
