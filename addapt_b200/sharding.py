"""Multi-GPU plumbing of the design path: one process per GPU, trajectories (or sequences) sharded by id with NO traffic
inside the loop; the only exchange is the final gather of each rank's top designs.

In the reference, parallelism is N independent OS processes with different seeds whose TSV files are merged offline
(reference data/20160705_production_run_favor_wt/production_run.sh:5-8, tools/utils.py:27-51).  Here trajectory t draws from
Philox keyed by (seed, t), so any partition of the ids over ranks gives the same trajectories, and the merge is one
all_gather of K fixed-size records per rank (NCCL over NVLink on the GPU box, gloo in the CPU tests)."""
import os
import sys

import numpy as np


def nccl_log_begin(world):
    """Keep NCCL's communicator log as evidence without polluting stdout (which carries result lines): NCCL_DEBUG=INFO / INIT
    (raised from VERSION / WARN if that is what the environment says) into a per-process file that nccl_log_end() copies to stderr.  (NCCL_DEBUG_FILE=/dev/stderr re-opens the descriptor and
    truncates a redirected stderr file.)  A caller's own NCCL_DEBUG / NCCL_DEBUG_FILE are respected.  Call before importing torch."""
    if world <= 1:
        return None
    if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION", "WARN"):      # quieter than INFO (the pool's boxes export VERSION)
        os.environ["NCCL_DEBUG"] = "INFO"
        os.environ.setdefault("NCCL_DEBUG_SUBSYS", "INIT")
    if "NCCL_DEBUG_FILE" in os.environ:
        return None
    path = "/tmp/addapt_b200_nccl_%d.log" % os.getpid()
    os.environ["NCCL_DEBUG_FILE"] = path
    return path


def nccl_log_end(path):
    if not path or not os.path.exists(path):
        return
    with open(path) as f:
        for line in f:
            sys.stderr.write(line)
    sys.stderr.flush()
    os.unlink(path)


def shard_range(n_units, rank, world):
    """contiguous id range [lo, hi) of `rank`: sizes differ by at most one, union is [0, n_units)"""
    base, extra = divmod(n_units, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def local_top_k(scores, k):
    """indices of the k best (largest) scores, best first; ties broken by lower index so the result is deterministic"""
    scores = np.asarray(scores, dtype=np.float64)
    k = min(k, scores.shape[0])
    order = np.lexsort((np.arange(scores.shape[0]), -scores))
    return order[:k]


def _first_of_each_sequence(order, seqs, k):
    """walk `order` (best first) and keep the first occurrence of every distinct sequence, up to k"""
    seen, keep = set(), []
    for idx in order:
        key = seqs[idx].tobytes()
        if key not in seen:
            seen.add(key)
            keep.append(idx)
            if len(keep) == k:
                break
    return np.asarray(keep, dtype=np.int64)


def gather_top_designs(scores, seqs, ids, k, dist=None, device=None, distinct=False):
    """Every rank passes its local trajectories' best scores [m], best sequences [m, n] (uint8) and global ids [m];
    every rank gets the global top-k (scores [<=k], seqs [<=k, n], ids [<=k]), best first, identical on all ranks.
    distinct=True ranks DESIGNS rather than trajectories: of the trajectories that ended on the same sequence only the one
    with the lowest id is kept (many trajectories converge to the same optimum).
    `dist` = torch.distributed (initialised) or None for a single process."""
    scores = np.asarray(scores, dtype=np.float64)
    seqs = np.asarray(seqs, dtype=np.uint8)
    ids = np.asarray(ids, dtype=np.int64)
    if distinct:
        pick = _first_of_each_sequence(np.lexsort((ids, -scores)), seqs, k)
    else:
        pick = local_top_k(scores, k)
    loc_s, loc_q, loc_i = scores[pick], seqs[pick], ids[pick]
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return loc_s, loc_q, loc_i
    import torch
    world = dist.get_world_size()
    n = seqs.shape[1]
    kk = int(k)
    # fixed-size record per rank: k scores (f64), k ids (i64 as f64 bit pattern is avoided: two tensors), k*n bytes; pad short ranks
    pad = kk - loc_s.shape[0]
    s_t = torch.full((kk,), -np.inf, dtype=torch.float64)
    i_t = torch.full((kk,), -1, dtype=torch.int64)
    q_t = torch.zeros((kk, n), dtype=torch.uint8)
    s_t[:kk - pad] = torch.from_numpy(loc_s)
    i_t[:kk - pad] = torch.from_numpy(loc_i)
    q_t[:kk - pad] = torch.from_numpy(np.ascontiguousarray(loc_q))
    if device is not None:
        s_t, i_t, q_t = s_t.to(device), i_t.to(device), q_t.to(device)
    S = [torch.empty_like(s_t) for _ in range(world)]
    I = [torch.empty_like(i_t) for _ in range(world)]
    Q = [torch.empty_like(q_t) for _ in range(world)]
    dist.all_gather(S, s_t)
    dist.all_gather(I, i_t)
    dist.all_gather(Q, q_t)
    all_s = torch.cat(S).cpu().numpy()
    all_i = torch.cat(I).cpu().numpy()
    all_q = torch.cat(Q).cpu().numpy()
    valid = all_i >= 0
    all_s, all_i, all_q = all_s[valid], all_i[valid], all_q[valid]
    order = np.lexsort((all_i, -all_s))               # best score first, ties by global id: independent of the sharding
    order = _first_of_each_sequence(order, all_q, kk) if distinct else order[:kk]
    return all_s[order], all_q[order], all_i[order]


def gather_top_designs_device(engine, d_scores, d_seqs, id0, k, dist=None):
    """The same exchange with the per-rank reduction ON THE DEVICE: `d_scores` [m] float64 and `d_seqs` [m, n] uint8 are torch
    CUDA tensors holding this rank's trajectories (global ids id0 .. id0+m-1); the engine's top-k kernel (adp_top_k) picks the k
    best, only those K records are all-gathered over NCCL, and the G*K gathered records are merged on the host.
    Returns (scores [<=k], seqs [<=k, n], ids [<=k]) as numpy arrays, best first, identical on every rank."""
    import torch
    m, n = d_seqs.shape
    kk = min(int(k), m)
    idx = torch.empty(kk, dtype=torch.int32, device=d_scores.device)
    top = torch.empty(kk, dtype=torch.float64, device=d_scores.device)
    engine.top_k_raw(m, d_scores, kk, idx, top)                      # one kernel launch on the engine's stream, synchronised on return
    s_t = torch.full((int(k),), float("-inf"), dtype=torch.float64, device=d_scores.device)
    i_t = torch.full((int(k),), -1, dtype=torch.int64, device=d_scores.device)
    q_t = torch.zeros((int(k), n), dtype=torch.uint8, device=d_scores.device)
    s_t[:kk], i_t[:kk], q_t[:kk] = top, idx.long() + int(id0), d_seqs.index_select(0, idx.long())
    if dist is not None and dist.is_initialized() and dist.get_world_size() > 1:
        world = dist.get_world_size()
        S = [torch.empty_like(s_t) for _ in range(world)]
        I = [torch.empty_like(i_t) for _ in range(world)]
        Q = [torch.empty_like(q_t) for _ in range(world)]
        dist.all_gather(S, s_t)
        dist.all_gather(I, i_t)
        dist.all_gather(Q, q_t)
        s_t, i_t, q_t = torch.cat(S), torch.cat(I), torch.cat(Q)
    all_s, all_i, all_q = s_t.cpu().numpy(), i_t.cpu().numpy(), q_t.cpu().numpy()
    valid = all_i >= 0
    all_s, all_i, all_q = all_s[valid], all_i[valid], all_q[valid]
    key = np.where(np.isnan(all_s), -np.inf, all_s)
    order = np.lexsort((all_i, np.isnan(all_s), -key))[:int(k)]      # best score first, NaN last, ties by global id
    return all_s[order], all_q[order], all_i[order]


def pick_best_scores(scores, window_size):
    """Peak picking over one trajectory's scores, as the reference's offline tools select designs from a TSV trajectory
    (reference tools/utils.py:36-51): repeatedly take the arg-max, record it, and blank +-window_size steps around it with the
    minimum score, while the maximum is still >= the 75th percentile of the ORIGINAL scores.  Returns the sorted step indices.
    (The reference loops forever when every score is equal -- blanking cannot push the maximum below the threshold; here a
    flat trajectory returns its first step, and a zero-width window still blanks the peak itself.)"""
    s = np.array(scores, dtype=np.float64)
    if s.size == 0:
        return []
    blocked, threshold = s.min(), np.percentile(s, 75)
    if blocked >= threshold:
        return [0]
    peaks = []
    while s.max() >= threshold:
        i = int(s.argmax())
        peaks.append(i)
        s[max(i - window_size, 0):max(min(i + window_size, s.size), i + 1)] = blocked      # [i-w, i+w), never empty
    return sorted(peaks)
