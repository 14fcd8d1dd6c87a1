#!/usr/bin/env python3
"""BASELINE.json configs[4]: construct-length sweep 100-300 nt.  For every length: apo+holo pf+bppm fold-evals/s and MFE folds/s of
8,192 constructs on one GPU (device-resident buffers, CUDA events), the algorithmic term count and the fraction of
the measured FP64-FMA roofline.  Constructs: the rhf(6) template with its lower-case ruler region padded with random nucleotides
(so every construct still contains the aptamer), seed 20260707.   python tools/sweep_lengths.py > profiles/r1_sweep_lengths.jsonl"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import addapt_b200  # noqa: E402
from bench import RHF6, THEO, count_terms  # noqa: E402


def padded_constructs(B, n, seed):
    """rhf(6) (102 nt) with n-102 random nucleotides inserted into the ruler (between nexus and hairpin, after position 80)"""
    rng = np.random.Generator(np.random.PCG64(seed))
    base = np.frombuffer(RHF6.upper().encode(), dtype=np.uint8)
    nts = np.frombuffer(b"ACGU", dtype=np.uint8)
    if n <= 102:
        a = np.tile(base[:n], (B, 1))
    else:
        ins = nts[rng.integers(0, 4, (B, n - 102))]
        a = np.concatenate([np.tile(base[:80], (B, 1)), ins, np.tile(base[80:], (B, 1))], axis=1)
    free = [41, 43, 44, 45, 46, 47, 48, 49]
    draw = rng.integers(0, 4, (B, len(free)))
    for c, p in enumerate(free):
        a[:, p] = nts[draw[:, c]]
    return np.ascontiguousarray(a)


def main():
    eng = addapt_b200.Engine(0)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    eng.set_stream(stream.cuda_stream)
    peak = eng.dfma_rate()
    B = 8192
    lengths = (100, 110, 122, 125, 128, 140, 150, 200, 250, 300)
    for n in lengths:
        host = padded_constructs(B, n, 20260707 + n)
        d = torch.from_numpy(host).cuda()
        dg = torch.zeros(B, 2, dtype=torch.float64, device="cuda")
        e = torch.zeros(B, dtype=torch.int32, device="cuda")
        st = torch.zeros(B, n, dtype=torch.uint8, device="cuda")
        eng.set_async(1)
        res = {}
        for name, fn in (("fold_eval", lambda: eng.fold_eval_raw(B, n, d, THEO, 0, None, None, dg, None)), ("mfe", lambda: eng.mfe_raw(B, n, d, e, st))):
            fn()
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            fn()
            fn()
            b.record(stream)
            torch.cuda.synchronize()
            res[name] = 2 * B / (a.elapsed_time(b) * 1e-3)
        eng.set_async(0)
        w = np.array([count_terms(r) for r in host[:: B // 8][:8]])
        w_in, w_out = float(w[:, 0].mean()), float(w[:, 1].mean())
        terms = 2 * (w_in + w_out)
        print(json.dumps({"n_nt": n, "constructs": B, "fold_evals_per_s": res["fold_eval"], "mfe_folds_per_s": res["mfe"],
                          "terms_per_fold_eval": terms, "achieved_gfma_s": terms * res["fold_eval"] / 1e9, "peak_gfma_s": peak,
                          "roofline_frac": terms * res["fold_eval"] / 1e9 / peak,
                          "kernel": ("pf_fast<256>, 2 CTAs/SM" if n <= 100 else "pf_fast<256>, Qm in L2 scratch, 2 CTAs/SM" if n <= 128 else
                                     "pf_fast<512>, 1 CTA/SM" if n <= 256 else "pf_fast<512>, level sums in global scratch")}), flush=True)
    # the occupancy cliff above 100 nt, with and without Qm in L2 scratch (plan_pf_fast), and the reference's 5-spacer production score
    # function (tools/test_seq:10-16: rhf(6) + 20-nt spacers = 122 nt) against the bare device, per partition-function fill
    from addapt_b200.capi import Problem
    from tests import cases
    for n in (110, 122, 125, 128):
        host = padded_constructs(B, n, 20260707 + n)
        d = torch.from_numpy(host).cuda()
        dg = torch.zeros(B, 2, dtype=torch.float64, device="cuda")
        rates = {}
        for flag in ("1", "0"):
            os.environ["ADDAPT_B200_QM_GLOBAL"] = flag
            eng.fold_eval_raw(B, n, d, THEO, 0, None, None, dg, None)
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            eng.fold_eval_raw(B, n, d, THEO, 0, None, None, dg, None)
            b.record(stream)
            torch.cuda.synchronize()
            rates[flag] = B / (a.elapsed_time(b) * 1e-3)
        del os.environ["ADDAPT_B200_QM_GLOBAL"]
        print(json.dumps({"n_nt": n, "fold_evals_per_s_qm_in_l2_2ctas": rates["1"], "fold_evals_per_s_all_smem_1cta": rates["0"]}), flush=True)
    designs = torch.from_numpy(addapt_b200.capi.seqs_to_array(cases.templated_designs(4096, 5, n=102))).cuda()
    sc = torch.zeros(4096, dtype=torch.float64, device="cuda")
    out = {}
    for name, pr in (("bare", cases.rhf6_problem()), ("5_spacers", cases.rhf6_problem(contexts=cases.SPACERS))):
        eng.score_raw(pr, 4096, designs, sc)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        eng.score_raw(pr, 4096, designs, sc)
        b.record(stream)
        torch.cuda.synchronize()
        fills = 4 * max(len(pr.contexts), 1)
        out[name] = {"scores_per_s": 4096 / (a.elapsed_time(b) * 1e-3), "fills_per_s": 4096 * fills / (a.elapsed_time(b) * 1e-3), "fills_per_score": fills}
    out["fill_rate_ratio_bare_over_spacers"] = out["bare"]["fills_per_s"] / out["5_spacers"]["fills_per_s"]
    print(json.dumps({"rhf6_score": out}), flush=True)
    eng.close()


if __name__ == "__main__":
    main()
