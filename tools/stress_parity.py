#!/usr/bin/env python3
"""One-off large randomized parity run (beyond the pytest cases): GPU engine vs the oracle on many random lengths, sequences,
constraints and motif placements.  Prints a summary line and exits non-zero on the first mismatch.
   python tools/stress_parity.py [seconds]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import addapt_b200  # noqa: E402
from oracle.oracle import Oracle  # noqa: E402
from tests import cases  # noqa: E402

budget = float(sys.argv[1]) if len(sys.argv) > 1 else 120.0
eng, orc = addapt_b200.Engine(0), Oracle()
rng = np.random.Generator(np.random.PCG64(424242))
t0 = time.time()
n_fold = n_seq = 0
worst_dg = worst_bp = 0.0
while time.time() - t0 < budget:
    n = int(rng.integers(5, 200)) if rng.random() < 0.97 else int(rng.integers(200, 321))      # a few long constructs (up to ADP_MAX_LEN)
    B = int(rng.integers(8, 64)) if n < 200 else int(rng.integers(4, 17))
    alphabet = ["ACGU", "GC", "GU", "ACGU", "AU", "ACGU"][int(rng.integers(0, 6))]
    seqs = ["".join(alphabet[k] for k in rng.integers(0, len(alphabet), n)) for _ in range(B)]
    motif = cases.THEO if rng.random() < 0.5 else None
    if motif and n > 40 and rng.random() < 0.7:         # plant the aptamer
        m = cases.THEO[0]
        at = int(rng.integers(0, n - len(m) + 1)) if n >= len(m) else 0
        seqs = [s[:at] + m + s[at + len(m):] if n >= len(m) else s for s in seqs]
    cst = cases.random_constraint(orc, seqs[0], rng) if rng.random() < 0.5 else None
    d0, b0 = orc.pf_batch(seqs, cst=cst, motif=motif, bppm=True)
    d1, b1 = eng.pf(seqs, cst=cst, motif=motif, bppm=True)
    fin = np.isfinite(d0)
    assert np.array_equal(np.isfinite(d1), fin), ("finiteness", n, cst)
    if fin.any():
        rel = np.abs(d1[fin] - d0[fin]) / np.maximum(np.abs(d0[fin]), 1e-3)
        worst_dg = max(worst_dg, float(rel.max()))
        err = np.abs(b1[fin] - b0[fin]) - 1e-6 * np.abs(b0[fin])
        worst_bp = max(worst_bp, float(err.max()))
        assert rel.max() <= 1e-9 and err.max() <= 1e-9, ("pf mismatch", n, cst, motif is not None, float(rel.max()), float(err.max()))
    e0, s0 = orc.mfe_batch(seqs, cst=cst, motif=motif)
    e1, s1 = eng.mfe(seqs, cst=cst, motif=motif)
    assert np.array_equal(e1, e0) and s1 == s0, ("mfe mismatch", n, cst, motif is not None)
    n_fold += 1
    n_seq += B
print("stress parity ok: %d batches, %d sequences, lengths 5-320; worst dG rel err %.2e, worst bppm excess err %.2e" % (n_fold, n_seq, worst_dg, worst_bp))
eng.close()
