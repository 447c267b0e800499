"""-m gpu: one world partitioned over R ranks must reproduce the single-GPU result exactly.

Run on ONE GPU: R contexts stand in for R ranks and the all-gather is a concatenation of device buffers, so the
partition logic, the manifold migration and the bit-exact union are covered without a multi-GPU box (bench.py
--workload c5 does the same exchange with NCCL)."""
import numpy as np
import pytest

import scenes

pytestmark = pytest.mark.gpu


def run_partitioned(pkg, sc, R, steps, mode=1, slots=False):
    import torch
    single = scenes.build_gpu(pkg, sc, mode=mode)
    ranks = [scenes.build_gpu(pkg, sc, mode=mode) for _ in range(R)]
    for r, w in enumerate(ranks):
        w.set_partition(r, R)
    cap = 1 << 16
    bufs = [(torch.zeros(cap, dtype=torch.int64, device="cuda"), torch.zeros(cap * 8, dtype=torch.int32, device="cuda"),
             torch.zeros(cap * 4 * 24, dtype=torch.int32, device="cuda")) for _ in range(R)]
    scap = 1 << 14
    sbytes = single.mgpu_slot_bytes(scap)
    allslots = torch.zeros(sbytes * R, dtype=torch.uint8, device="cuda")
    migrated_total = 0
    for step in range(steps):
        xf = sc.transforms(step)
        single.setWorldTransforms(xf)
        single.performDiscreteCollisionDetection()
        for w in ranks:
            w.setWorldTransforms(xf)
            w.mgpu_broadphase()
        if slots:
            # sync-free variant: every rank packs its slot straight into the "gathered" buffer, then all ranks scan it
            for r, w in enumerate(ranks):
                w.mgpu_export_departed_slot(allslots.data_ptr() + r * sbytes, scap)
                w.sync_counts()
            tot = int(sum(int(allslots[r * sbytes:r * sbytes + 4].view(torch.int32)[0]) for r in range(R)))
            for w in ranks:
                w.mgpu_import_arrival_slots(allslots.data_ptr(), R, scap)
        else:
            counts = []
            for r, w in enumerate(ranks):
                k, h, p = bufs[r]
                counts.append(w.mgpu_export_departed(k.data_ptr(), h.data_ptr(), p.data_ptr(), cap))
            # "all-gather": concatenate every rank's departed manifolds
            keys = torch.cat([bufs[r][0][:counts[r]] for r in range(R)])
            hdrs = torch.cat([bufs[r][1][:counts[r] * 8] for r in range(R)])
            pts = torch.cat([bufs[r][2][:counts[r] * 96] for r in range(R)])
            torch.cuda.synchronize()
            tot = int(sum(counts))
            for w in ranks:
                w.mgpu_import_arrivals(keys.data_ptr(), hdrs.data_ptr(), pts.data_ptr(), tot)
        for w in ranks:
            w.mgpu_narrowphase()
            w.sync_counts()
        # union of the ranks == the single world
        pr = np.concatenate([w.pairs() for w in ranks])
        order = np.lexsort((pr[:, 1], pr[:, 0]))
        assert np.array_equal(pr[order], single.pairs()), f"pair union differs at step {step}"
        mr = np.concatenate([w.manifolds() for w in ranks])
        mo = np.lexsort((mr["pair_uid1"], mr["pair_uid0"]))
        ms = single.manifolds()
        assert len(mr) == len(ms), f"manifold count differs at step {step}: {len(mr)} vs {len(ms)}"
        assert mr[mo].tobytes() == ms.tobytes(), f"manifolds differ at step {step}"
        # the migrated manifolds of this step that found a new owner
        migrated_total += tot
        sizes = [len(w.pairs()) for w in ranks]
        assert min(sizes) > 0
    return migrated_total


def test_partitioned_spheres_world_matches_single(gpu_pkg):
    sc = scenes.spheres_scene(n=20000, seed=6)
    sc.vel *= 6.0  # enough motion that the sorted order (and with it pair ownership) changes every step
    migrated = run_partitioned(gpu_pkg, sc, R=2, steps=6)
    assert migrated > 0, "no manifold ever departed: the migration path was not exercised"


def test_partitioned_spheres_world_slot_exchange(gpu_pkg):
    sc = scenes.spheres_scene(n=20000, seed=8)
    sc.vel *= 6.0
    migrated = run_partitioned(gpu_pkg, sc, R=4, steps=5, slots=True)
    assert migrated > 0


def test_partitioned_bin_world_with_large_statics(gpu_pkg):
    sc = scenes.bin_scene(n=4000, seed=13)
    sc.vel *= 3.0
    run_partitioned(gpu_pkg, sc, R=3, steps=5)
