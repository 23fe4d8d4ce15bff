"""CPU, world_size 2 over gloo: the multi-GPU host logic (contiguous sharding + the single aggregate exchange)."""
import os
import socket

import numpy as np
import torch
import torch.multiprocessing as mp

from pyhybridcontrol_b200 import distributed as D


def test_shard_range_partitions():
    for B, W in ((100, 1), (100, 8), (1001, 4), (3, 8)):
        spans = [D.shard_range(B, r, W) for r in range(W)]
        assert spans[0][0] == 0 and spans[-1][1] == B
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        sizes = [hi - lo for lo, hi in spans]
        assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.distributed.init_process_group("gloo", rank=rank, world_size=world)
    B, Nt = 11, 7
    rng = np.random.default_rng(0)
    u = (rng.random((B, Nt)) > 0.5).astype(float)
    P = rng.uniform(2700, 3300, B)
    lo, hi = D.shard_range(B, rank, world)
    local = torch.as_tensor((P[lo:hi, None] * u[lo:hi]).sum(axis=0))
    total = D.allreduce_aggregate(local.clone())
    traj = D.allgather_trajectories(torch.as_tensor(P[lo:hi, None] * u[lo:hi]),
                                    counts=[D.shard_range(B, r, world)[1] - D.shard_range(B, r, world)[0] for r in range(world)])
    ref = (P[:, None] * u).sum(axis=0)
    ok = np.allclose(total.numpy(), ref, rtol=1e-13) and np.array_equal(traj.numpy(), P[:, None] * u)
    g = D.grid_evaluate(total, torch.full((Nt,), -9000.0, dtype=torch.float64), torch.full((Nt,), 4000.0, dtype=torch.float64))
    ok = ok and bool(torch.all(g["p_imp"] >= 0)) and bool(torch.allclose(g["p_imp"] + g["p_exp"], g["y"]))
    out[rank] = ok
    torch.distributed.destroy_process_group()


def test_aggregate_exchange_world2():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    assert out[0] and out[1]


def _worker_blocks(rank, world, port, out):
    """the loop skeleton of DewhFleet._best_response on uneven shards: every rank must issue the same number of
    all-reduces per pass, whatever its own shard size"""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.distributed.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = D.shard_range(5, rank, world)          # 3 agents on rank 0, 2 on rank 1
    B = hi - lo
    B_max = D.allreduce_max_int(B)
    calls, covered, G = 0, set(), max(1, min(2, B_max))
    sums = []
    for _ in range(3):                              # the block count doubles up to the LARGEST shard
        for blo, bhi in D.response_blocks(B, G):
            covered.update(range(blo, bhi))
            t = torch.tensor([float(bhi - blo)], dtype=torch.float64)
            D.allreduce_aggregate(t)                # (would hang or mispair if the ranks disagreed on the count)
            sums.append(float(t))
            calls += 1
        G = min(B_max, 2 * G)
    out[rank] = (calls, B_max, sorted(covered) == list(range(B)), sums)
    torch.distributed.destroy_process_group()


def test_best_response_blocks_agree_on_uneven_shards():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker_blocks, args=(2, port, out), nprocs=2, join=True)
    assert out[0][0] == out[1][0] == 2 + 3 + 3 and out[0][1] == out[1][1] == 3
    assert out[0][2] and out[1][2]
    assert out[0][3] == out[1][3]                   # every all-reduce paired the same block on both ranks
