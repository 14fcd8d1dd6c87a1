#!/usr/bin/env python3
"""Tiny invocations of every kernel for compute-sanitizer (memcheck / racecheck / initcheck):
   compute-sanitizer --tool racecheck python tools/sanitize_small.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import addapt_b200  # noqa: E402
from tests import cases  # noqa: E402

eng = addapt_b200.Engine(0)
seqs = cases.templated_designs(3, 1, n=100)
dg, pr, full = eng.fold_eval(seqs, cases.THEO, queries=[(0, 29), (47, 79)], full=True)
print("fold_eval", dg[0])
print("pf cst", eng.pf(seqs, cst=cases.RHF6_ACTIVE[:100], motif=cases.THEO)[0])
e, st = eng.mfe(seqs)
print("mfe", e[0], st[0])
print("mfe N", eng.mfe([seqs[0][:50] + "N" + seqs[0][51:]])[0])
print("pf 150", eng.pf(cases.random_seqs(2, 150, 3), bppm=True)[0][0])
print("pf 300", eng.pf(cases.random_seqs(1, 300, 4))[0])
problem = cases.rhf6_problem()
print("score", eng.score(problem, [cases.RHF6, cases.RHF6]))
r = eng.mc_run(problem, 2, 1, 4, thermostat=("auto", 0.5, 2, 1.0))
print("mc", r["outcomes"])
eng.set_general_only(1)
print("general pf", eng.pf(seqs[:1], bppm=True)[0])
print("general mfe", eng.mfe(seqs[:1])[0])
eng.close()
print("done")
