"""Short timing of the MFE entry point (tuning runs; ADDAPT_B200_LIB selects a variant build)."""
import sys, time
sys.path.insert(0, '.')
import torch
import addapt_b200
from tests import cases
eng = addapt_b200.Engine(0)
B = 16384
a = torch.from_numpy(addapt_b200.capi.seqs_to_array(cases.templated_designs(B, 1, n=100))).cuda()
e = torch.zeros(B, dtype=torch.int32, device='cuda')
st = torch.zeros(B, 100, dtype=torch.uint8, device='cuda')
for rep in range(4):
    torch.cuda.synchronize(); t0 = time.time()
    eng.mfe_raw(B, 100, a, e, st)
    torch.cuda.synchronize(); t1 = time.time()
print("%s mfe %.0f folds/s  checksum %d" % (sys.argv[1], B / (t1 - t0), int(e.sum())))
