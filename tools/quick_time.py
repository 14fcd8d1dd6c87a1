import time, sys, numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, '..')
import addapt_b200
from tests import cases
eng = addapt_b200.Engine(0)
print("dfma GFMA/s", eng.dfma_rate())
import torch
for B in (4096, 16384):
    seqs = cases.templated_designs(B, 1, n=100)
    a = torch.from_numpy(addapt_b200.capi.seqs_to_array(seqs)).cuda()
    dg = torch.zeros(B, 2, dtype=torch.float64, device='cuda')
    for rep in range(3):
        torch.cuda.synchronize(); t0 = time.time()
        eng.fold_eval_raw(B, 100, a, cases.THEO, 0, None, None, dg, None)
        torch.cuda.synchronize(); t1 = time.time()
        print("fold_eval B=%d: %.3f s -> %.0f fold-evals/s" % (B, t1 - t0, B / (t1 - t0)))
    e = torch.zeros(B, dtype=torch.int32, device='cuda')
    st = torch.zeros(B, 100, dtype=torch.uint8, device='cuda')
    for rep in range(2):
        torch.cuda.synchronize(); t0 = time.time()
        eng.mfe_raw(B, 100, a, e, st)
        torch.cuda.synchronize(); t1 = time.time()
        print("mfe B=%d: %.3f s -> %.0f folds/s" % (B, t1 - t0, B / (t1 - t0)))
    d1 = torch.zeros(B, dtype=torch.float64, device='cuda')
    for rep in range(2):
        torch.cuda.synchronize(); t0 = time.time()
        eng.pf_raw(B, 100, a, d1)
        torch.cuda.synchronize(); t1 = time.time()
        print("pf inside B=%d: %.3f s -> %.0f fills/s" % (B, t1 - t0, B / (t1 - t0)))
pr = cases.rhf6_problem(n=100)
sc = torch.zeros(4096, dtype=torch.float64, device='cuda')
a = torch.from_numpy(addapt_b200.capi.seqs_to_array(cases.templated_designs(4096, 1, n=100))).cuda()
torch.cuda.synchronize(); t0 = time.time(); eng.score_raw(pr, 4096, a, sc); torch.cuda.synchronize(); print("score 4096: %.3f s" % (time.time() - t0))
t0 = time.time(); r = eng.mc_run(pr, 1024, 1, 20); print("mc 1024x20: %.3f s" % (time.time() - t0), r["outcomes"].sum(axis=0))
