#!/usr/bin/env python3
"""Writes tests/golden/oracle_vectors.json: outputs of the CPU oracle (oracle/) on fixed seeded inputs, so that
the -m gpu tests can also be checked against committed numbers (the oracle itself is re-run live as well).
libRNA is not available (see DESIGN.md "Oracle"), so these are the ORACLE's values, not ViennaRNA's.
Run:  python tests/golden/make_oracle_golden.py
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.oracle import Oracle, tri  # noqa: E402
from tests import cases  # noqa: E402

O = Oracle()
G = {"par_file": "params/rna_turner2004_reconstructed.par", "kT_kcal": O.kT}
seqs = [cases.HAIRPIN, cases.THEO[0], cases.RHF6.upper()] + cases.random_seqs(6, 60, 11) + cases.random_seqs(6, 100, 12) + cases.random_seqs(2, 150, 13)
rng = np.random.Generator(np.random.PCG64(99))
folds = []
for s in seqs:
    n = len(s)
    e, st = O.mfe(s)
    eh, sth = O.mfe(s, motif=cases.THEO)
    dg, bp = O.pf(s, bppm=True)
    dgh, bph = O.pf(s, motif=cases.THEO, bppm=True)
    cst = cases.random_constraint(O, s, rng)
    ec, stc = O.mfe(s, cst=cst)
    dgc = O.pf(s, cst=cst)
    top = np.argsort(bp)[::-1][:8]
    folds.append(dict(seq=s, mfe=[e, st], mfe_holo=[eh, sth], dG=dg, dG_holo=dgh, cst=cst, mfe_cst=[ec, stc], dG_cst=dgc,
                      bppm_sum=float(bp.sum()), bppm_holo_sum=float(bph.sum()),
                      bppm_top=[[int(k), float(bp[k])] for k in top], w_terms=list(O.count_terms(s))))
G["folds"] = folds
prob = dict(seq=cases.RHF6, macrostates={"active": cases.RHF6_ACTIVE},
            terms=[("active", "apo", False, 1.0), ("active", "holo", True, 1.0)], motif=cases.THEO)
designs = cases.templated_designs(8, 21, n=102)
sc, terms = O.score_batch(prob, designs)
G["score"] = dict(designs=designs, score=[float(x) for x in sc], terms=[[float(y) for y in t] for t in terms])
prob5 = dict(prob, contexts=cases.SPACERS)
sc5, terms5 = O.score_batch(prob5, designs[:2])
G["score_5ctx"] = dict(designs=designs[:2], score=[float(x) for x in sc5])
mc = O.mc_run(prob, seed=2016, traj_id=7, n_steps=40)
G["mc"] = dict(seed=2016, traj_id=7, n_steps=40, thermostat=["annealing", 5.0, 0.0, 300], final_seq=mc["final_seq"],
               final_score=mc["final_score"], best_seq=mc["best_seq"], best_score=mc["best_score"], best_step=mc["best_step"],
               outcomes=mc["outcomes"])
G["philox"] = dict(key=[2016, 0], ctr=[5, 0, 7, 0], out=O.philox(2016, 0, [5, 0, 7, 0]))
out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "oracle_vectors.json")
with open(out, "w") as f:
    json.dump(G, f, indent=1)
print("wrote", out, os.path.getsize(out), "bytes")
