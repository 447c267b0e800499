"""CPU, world_size 2 over gloo: the N>1 plumbing of bench.py (one independent world per rank, MAX of the per-step
time, SUM of the counts) without a GPU."""
import os
import socket
import sys

import numpy as np
import pytest
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


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import bench
    import orc
    import scenes
    # each rank owns one independent world (seed depends on the rank, like bench.run_ours)
    sc = bench.make_scene(1200, seed=100 + rank)
    ow = scenes.build_oracle(sc, orc.DBVT)
    pairs = ow.step(sc.transforms(0))
    ms = 10.0 + rank  # stand-in for the per-rank device time
    vals = torch.tensor([ms], dtype=torch.float64)
    sums = torch.tensor([float(len(pairs))], dtype=torch.float64)
    dist.all_reduce(vals, op=dist.ReduceOp.MAX)
    dist.all_reduce(sums, op=dist.ReduceOp.SUM)
    dist.barrier()
    q.put((rank, float(vals[0]), float(sums[0]), len(pairs), float(sc.base[7, 9])))
    dist.destroy_process_group()


def test_two_rank_world_sharding_and_reduction():
    world = 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in ps:
        p.start()
    res = sorted(q.get(timeout=240) for _ in range(world))
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, mx0, sum0, n0, x0), (r1, mx1, sum1, n1, x1) = res
    assert mx0 == mx1 == 11.0                      # MAX over ranks
    assert sum0 == sum1 == float(n0 + n1)          # SUM over ranks
    assert x0 != x1                                # different worlds per rank
    assert n0 > 1000 and n1 > 1000
