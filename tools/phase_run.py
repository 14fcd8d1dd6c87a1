"""Per-phase clock64() timing of the fast PF kernel (tools/variant_build.sh prof -DADP_PHASE_PROFILE; ADDAPT_B200_LIB=build/prof/libaddapt_b200_prof.so): runs one small fold-eval batch on the instrumented
build; the kernel prints, for every warp of CTA 0, the cycles spent in each phase of the column loop."""
import sys
sys.path.insert(0, '.')
import torch
import addapt_b200
from tests import cases
eng = addapt_b200.Engine(0)
B = 296
seqs = cases.templated_designs(B, 1, n=100)
a = torch.from_numpy(addapt_b200.capi.seqs_to_array(seqs)).cuda()
dg = torch.zeros(B, 2, dtype=torch.float64, device='cuda')
eng.fold_eval_raw(B, 100, a, cases.THEO, 0, None, None, dg, None)
torch.cuda.synchronize()
