"""Short timing of the fold-eval and inside-only entry points (tuning runs; ADDAPT_B200_LIB selects a variant build)."""
import sys, time
sys.path.insert(0, '.')
import torch
import addapt_b200
from tests import cases
eng = addapt_b200.Engine(0)
B = 8192
seqs = cases.templated_designs(B, 1, n=100)
a = torch.from_numpy(addapt_b200.capi.seqs_to_array(seqs)).cuda()
dg = torch.zeros(B, 2, dtype=torch.float64, device='cuda')
d1 = torch.zeros(B, dtype=torch.float64, device='cuda')
for rep in range(3):
    torch.cuda.synchronize(); t0 = time.time()
    eng.fold_eval_raw(B, 100, a, cases.THEO, 0, None, None, dg, None)
    torch.cuda.synchronize(); t1 = time.time()
    eng.pf_raw(B, 100, a, d1)
    torch.cuda.synchronize(); t2 = time.time()
print("%s fold_eval %.0f /s   inside %.0f fills/s   checksum %.9f" % (sys.argv[1] if len(sys.argv) > 1 else "", B / (t1 - t0), B / (t2 - t1), float(dg.sum())))
