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


class _FakeWorld:
    """Stands in for GpuCollisionWorld in the host-logic test: records the calls PartitionedStepper makes."""

    def __init__(self, rank, p2p_works):
        self.rank, self.p2p_works, self.calls, self.num_bodies = rank, p2p_works, [], 1000

    def stream(self):
        return 0

    def set_partition(self, rank, nranks):
        self.calls.append(("set_partition", rank, nranks))

    def mgpu_slot_bytes(self, cap):
        return 16 + cap * 424

    def mgpu_halo_slot_bytes(self, cap):
        return 16 + cap * 80

    def mgpu_p2p_init(self, cap, mcap):
        if not self.p2p_works:
            raise RuntimeError("no peer access")
        return bytes([self.rank + 1] * 64), 0x1000 * (self.rank + 1)

    def mgpu_p2p_connect(self, ipc_handles=None, inbox_ptrs=None):
        self.calls.append(("connect", bytes(ipc_handles)))

    def __getattr__(self, name):          # every per-step entry point: just note it
        if name.startswith("mgpu_"):
            return lambda *a: self.calls.append((name,))
        raise AttributeError(name)


def _stepper_worker(rank, world, port, q, broken_rank):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import importlib.util
    spec = importlib.util.spec_from_file_location("partitioned", os.path.join(ROOT, "libgdx-jbullet_b200", "partitioned.py"))
    part = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(part)
    gw = _FakeWorld(rank, p2p_works=(rank != broken_rank))
    st = part.PartitionedStepper(gw, rank, world, dist, torch, "cpu", migrate_cap=64, halo_cap=128)
    st.step()
    q.put((rank, st.halo_mode, [c[0] for c in gw.calls], [c[1] for c in gw.calls if c[0] == "connect"], st.describe()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("broken_rank", [-1, 1])
def test_stepper_exchange_mode_is_agreed_by_all_ranks(broken_rank):
    """PartitionedStepper over gloo with a stand-in world: the IPC handles are all-gathered in rank order and every rank
    runs the collective-free sequence; if ONE rank cannot map its peers, ALL ranks fall back to the all-gather sequence."""
    world = 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_stepper_worker, args=(r, world, port, q, broken_rank)) for r in range(world)]
    for p in ps:
        p.start()
    res = sorted(q.get(timeout=240) for _ in range(world))
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, mode, calls, handles, text in res:
        if broken_rank < 0:
            assert mode == "p2p"
            assert handles == [bytes([1] * 64) + bytes([2] * 64)]              # rank order
            assert calls[-6:] == ["mgpu_p2p_export_halo", "mgpu_p2p_import_halo", "mgpu_broadphase", "mgpu_p2p_export_departed",
                                  "mgpu_p2p_import_arrivals", "mgpu_narrowphase"]
            assert "NO collective" in text
        else:
            assert mode == "nccl"
            assert calls[-6:] == ["mgpu_update_export_halo", "mgpu_import_halo", "mgpu_broadphase", "mgpu_export_departed_slot",
                                  "mgpu_import_arrival_slots", "mgpu_narrowphase"]
            assert "ncclAllGather" in text
