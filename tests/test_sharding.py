"""N > 1 path on CPU: two gloo ranks shard the unit ids, work independently, and merge their top designs with one
all_gather; the merged result must equal the single-process answer whatever the partition."""
import os
import socket
import subprocess
import sys

import numpy as np

from addapt_b200.sharding import gather_top_designs, local_top_k, shard_range

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_ranges_partition_the_ids():
    for n in (0, 1, 7, 64, 65536, 65537):
        for world in (1, 2, 3, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def test_local_top_k_is_deterministic():
    s = np.array([1.0, 3.0, 3.0, -np.inf, 2.0])
    assert list(local_top_k(s, 3)) == [1, 2, 4]
    assert list(local_top_k(s, 10)) == [1, 2, 4, 0, 3]


WORKER = r"""
import os, sys, numpy as np, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
from addapt_b200.sharding import gather_top_designs, shard_range
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
N, n, K = 1000, 12, 16
rng = np.random.Generator(np.random.PCG64(5))
scores = np.round(rng.normal(size=N), 1)              # rounded: plenty of ties across ranks
seqs = rng.integers(65, 70, (N, n)).astype(np.uint8)
lo, hi = shard_range(N, rank, world)
s, q, i = gather_top_designs(scores[lo:hi], seqs[lo:hi], np.arange(lo, hi), K, dist)
order = np.lexsort((np.arange(N), -scores))[:K]
assert np.array_equal(i, order), (rank, i, order)
assert np.array_equal(s, scores[order]) and np.array_equal(q, seqs[order])
# distinct designs: many trajectories end on the same sequence; the merged list holds each sequence once (lowest id wins)
dup = seqs[rng.integers(0, 40, N)]
s2, q2, i2 = gather_top_designs(scores[lo:hi], dup[lo:hi], np.arange(lo, hi), K, dist, distinct=True)
want, seen = [], set()
for idx in np.lexsort((np.arange(N), -scores)):
    if dup[idx].tobytes() not in seen:
        seen.add(dup[idx].tobytes()); want.append(idx)
    if len(want) == K: break
assert np.array_equal(i2, np.array(want)), (rank, i2, want)
assert len({q.tobytes() for q in q2}) == len(q2)
dist.barrier()
dist.destroy_process_group()
print("rank", rank, "ok")
"""


def test_two_gloo_ranks_merge_to_the_single_process_answer(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                        "--master-port", str(port), str(script), ROOT], capture_output=True, text=True, timeout=300, env=env)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-3000:]
    assert p.stdout.count("ok") == 2


def test_single_process_gather_is_local_top_k():
    rng = np.random.Generator(np.random.PCG64(1))
    scores = rng.normal(size=50)
    seqs = rng.integers(65, 70, (50, 8)).astype(np.uint8)
    s, q, i = gather_top_designs(scores, seqs, np.arange(100, 150), 5)
    assert list(i) == [100 + k for k in local_top_k(scores, 5)]
    assert np.array_equal(q, seqs[local_top_k(scores, 5)])


def test_pick_best_scores_matches_the_reference_rule():
    """reference tools/utils.py:36-51 -- peaks >= 75th percentile with an exclusion window"""
    from addapt_b200.sharding import pick_best_scores
    scores = [0, 1, 5, 1, 0, 0, 4, 0, 0, 9, 8, 0]
    # 75th percentile = 4.25; the window is [i-w, i+w): the 8 right after the 9 survives w = 1 but not w = 2
    assert pick_best_scores(scores, 1) == [2, 9, 10]
    assert pick_best_scores(scores, 2) == [2, 9]
    assert pick_best_scores(scores, 0) == [2, 9, 10]         # zero-width window: only the peak itself is blanked
    assert pick_best_scores([3, 3, 3], 2) == [0]             # flat trajectory (the reference would not terminate)
    assert pick_best_scores([], 5) == []
    rng = np.random.Generator(np.random.PCG64(8))
    s = rng.normal(size=400).cumsum()
    picks = pick_best_scores(s, 25)
    thr = np.percentile(s, 75)
    assert all(s[i] >= thr for i in picks)
    assert all(b - a >= 25 for a, b in zip(picks, picks[1:])) and len(picks) >= 1
