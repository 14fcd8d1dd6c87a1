#!/usr/bin/env python3
"""BASELINE.json configs[3]: independent annealing Monte-Carlo design trajectories of the rhf(6) theophylline sgRNA, sharded over the
GPUs of one box (one process per GPU under torchrun), no traffic inside the loop, the top designs gathered over NCCL at the end.

  torchrun --nproc-per-node 8 tools/run_design.py --traj 65536 --steps 1000 --topk 1024 --out gpurun_out/design_run.json

Trajectory t draws from Philox keyed by (seed, t): the result does not depend on the number of GPUs."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--traj", type=int, default=65536)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--topk", type=int, default=1024)
    ap.add_argument("--seed", type=int, default=2016)
    ap.add_argument("--thermostat", default="5 to 0 in 300 steps")
    ap.add_argument("--distinct", action="store_true", help="rank distinct designs instead of trajectories")
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    from addapt_b200.sharding import nccl_log_begin, nccl_log_end      # NCCL's communicator lines are evidence: kept, on stderr
    nccl_log = nccl_log_begin(world)
    import torch
    import torch.distributed as dist
    import addapt_b200
    from addapt_b200.sharding import gather_top_designs, gather_top_designs_device, shard_range
    from tests import cases

    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    hi, lo, cyc = args.thermostat.split(" to ")[0], args.thermostat.split(" to ")[1].split(" in ")[0], args.thermostat.split(" in ")[1].split()[0]
    thermo = ("annealing", float(hi), float(lo), int(cyc))
    eng = addapt_b200.Engine(local)
    problem = cases.rhf6_problem()
    a, b = shard_range(args.traj, rank, world)
    start_score = float(eng.score(problem, [cases.RHF6])[0])
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    # sequences, scores and outcome counters stay on the device: the run's only exchange is the top-K gather below
    m, n = b - a, problem.n
    dev = torch.device("cuda", local)
    d_best_seq = torch.zeros(m, n, dtype=torch.uint8, device=dev)
    d_best = torch.zeros(m, dtype=torch.float64, device=dev)
    d_oc = torch.zeros(m, 4, dtype=torch.int32, device=dev)
    eng.mc_run_raw(problem, m, args.seed, a, args.steps, thermo, best_seq=d_best_seq, best_score=d_best, outcomes=d_oc)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    oc = d_oc.to(torch.int64).sum(dim=0)
    tmax = torch.tensor([dt], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(oc)
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    tg = time.perf_counter()
    if args.distinct:      # ranking distinct designs needs every trajectory's sequence: host path
        top_s, top_q, top_i = gather_top_designs(d_best.cpu().numpy(), d_best_seq.cpu().numpy(), np.arange(a, b), args.topk, dist if world > 1 else None, dev, distinct=True)
    else:                  # per-rank top-K on the device (adp_top_k), K records over NCCL
        top_s, top_q, top_i = gather_top_designs_device(eng, d_best, d_best_seq, a, args.topk, dist if world > 1 else None)
    tg = time.perf_counter() - tg
    if rank == 0:
        top_seqs = [bytes(q).decode() for q in top_q]
        recheck = eng.score(problem, top_seqs[:64])                     # the gathered designs, scored again from scratch
        assert np.allclose(recheck, top_s[:64], rtol=1e-6), "gathered scores do not match a fresh evaluation"
        oc = oc.cpu().numpy()
        total = int(oc.sum())
        out = {"config": "configs[3]: %d trajectories x %d steps, thermostat '%s', seed %d" % (args.traj, args.steps, args.thermostat, args.seed),
               "n_gpus": world, "seconds": float(tmax.item()), "energy_params": eng.param_provenance,
               "exchange": "per-rank top-K on the device (adp_top_k) + all_gather of K records over NCCL" if not args.distinct else "host merge of distinct designs", "mc_steps_per_s": args.traj * args.steps / float(tmax.item()),
               "outcomes": dict(zip(["REJECT", "ACCEPT_WORSENED", "ACCEPT_UNCHANGED", "ACCEPT_IMPROVED"], [int(x) for x in oc])),
               "outcome_fractions": [round(float(x) / max(total, 1), 4) for x in oc],
               "start_score": start_score, "top_k": int(top_s.shape[0]), "gather_seconds": tg,
               "best_score": float(top_s[0]), "score_of_rank_k": float(top_s[-1]),
               "distinct_sequences_in_top_k": len(set(top_seqs)),
               "top5": [{"score": float(s), "trajectory": int(i), "seq": q} for s, i, q in zip(top_s[:5], top_i[:5], top_seqs[:5])]}
        line = json.dumps(out)
        print(line)
        if args.out:
            open(args.out, "w").write(line + "\n")
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    nccl_log_end(nccl_log)
    eng.close()


if __name__ == "__main__":
    main()
