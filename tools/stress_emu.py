#!/usr/bin/env python3
"""One-off large randomized parity run of the HOST EMULATION of the fast kernels (tests/emu: the very source of pf_fast.cuh /
mfe_fast.cuh with one virtual thread) against the oracle: random lengths, alphabets, hard constraints and motif placements.
Covers the recursions and the constraint / motif handling, not the parallel decomposition (that is tools/stress_parity.py on
the GPU).  Prints a summary line and exits non-zero on the first mismatch.   python tools/stress_emu.py [seconds]"""
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
subprocess.check_call(["make", "-C", os.path.join(ROOT, "tests", "emu"), "-s"])
from oracle.oracle import Oracle  # noqa: E402
from tests import cases  # noqa: E402
from tests.emu_binding import Emu  # noqa: E402

budget = float(sys.argv[1]) if len(sys.argv) > 1 else 120.0
emu, orc = Emu(), Oracle()
rng = np.random.Generator(np.random.PCG64(777))
t0 = time.time()
n_seq = 0
worst_dg = worst_bp = 0.0
while time.time() - t0 < budget:
    n = int(rng.integers(5, 200))
    alphabet = ["ACGU", "GC", "GU", "ACGU", "AU", "ACGU"][int(rng.integers(0, 6))]
    s = "".join(alphabet[k] for k in rng.integers(0, len(alphabet), n))
    motif = cases.THEO if rng.random() < 0.5 else None
    if motif and n > 40 and rng.random() < 0.7:         # plant the aptamer
        m = cases.THEO[0]
        at = int(rng.integers(0, n - len(m) + 1))
        s = s[:at] + m + s[at + len(m):]
    cst = cases.random_constraint(orc, s, rng) if rng.random() < 0.5 else None
    d0, b0 = orc.pf(s, cst=cst, motif=motif, bppm=True)
    d1, b1 = emu.pf(s, cst=cst, motif=motif, bppm=True, fast=True)
    if np.isinf(d0):
        assert d1 == d0, ("finiteness", s, cst)
    else:
        rel = abs(d1 - d0) / max(abs(d0), 1e-3)
        err = float((np.abs(b1 - b0) - 1e-6 * np.abs(b0)).max())
        worst_dg, worst_bp = max(worst_dg, rel), max(worst_bp, err)
        assert rel <= 1e-9 and err <= 1e-9, ("pf mismatch", s, cst, motif is not None, rel, err)
    assert emu.mfe(s, cst=cst, motif=motif, fast=True) == orc.mfe(s, cst=cst, motif=motif), ("mfe mismatch", s, cst, motif is not None)
    n_seq += 1
print("emulation stress parity ok: %d sequences, lengths 5-199; worst dG rel err %.2e, worst bppm excess err %.2e" % (n_seq, worst_dg, worst_bp))
