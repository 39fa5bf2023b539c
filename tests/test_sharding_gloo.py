"""world_size-2 run of the multi-GPU host logic on the CPU with the gloo backend: each rank takes its block of the
pair list, "processes" it, and the only collective is the (sum of units, max of time) reduction used by bench.py."""
import os
import socket
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_pairs, q):
    sys.path.insert(0, os.path.join(ROOT, "piv_liteflownet-pytorch_b200"))
    from pivlfn import shard
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard.pair_range(n_pairs, rank, world)
    f_lo, f_hi = shard.frame_range(n_pairs + 1, rank, world)
    done = list(range(lo, hi))
    total, t = shard.gather_counts(len(done), 10.0 * (rank + 1))
    gathered = [None] * world
    dist.all_gather_object(gathered, (done, f_lo, f_hi))
    if rank == 0:
        q.put((total, t, gathered))
    dist.barrier()
    dist.destroy_process_group()


def test_pair_sharding_two_ranks():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    n_pairs = 37
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_pairs, q)) for r in range(2)]
    for p in procs:
        p.start()
    total, t, gathered = q.get(timeout=120)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert total == n_pairs and t == 20.0                    # units summed, time = max over ranks
    all_pairs = sorted(sum((g[0] for g in gathered), []))
    assert all_pairs == list(range(n_pairs))                  # every pair exactly once
    assert gathered[0][2] == gathered[1][1]                   # shared boundary frame
