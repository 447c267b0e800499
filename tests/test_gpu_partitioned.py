"""-m gpu: one world partitioned over R ranks by slabs with a halo must reproduce the single-GPU result
exactly.

Run on ONE GPU: R contexts stand in for R ranks and each all-gather is "every rank writes its slot straight into the
gathered buffer", so slab ownership, the halo exchange of boundary AABBs, the manifold migration and the bit-exact union
are covered without a multi-GPU box (bench.py's c5 object does the same exchanges with NCCL and checks the same union
inside the real multi-rank run, libgdx-jbullet_b200/partitioned.py:partition_check)."""
import numpy as np
import pytest

import scenes

pytestmark = pytest.mark.gpu


def run_partitioned(pkg, sc, R, steps, mode=1, slots=False, axis=None, planes=None, halo_cap=None, stats=None, p2p=False):
    import torch
    single = scenes.build_gpu(pkg, sc, mode=mode)
    ranks = [scenes.build_gpu(pkg, sc, mode=mode) for _ in range(R)]
    for r, w in enumerate(ranks):
        if planes is None:
            w.set_partition(r, R)
        else:
            w.set_partition_slabs(r, R, axis, planes)
    hcap = halo_cap or max(4096, sc.n)
    if p2p:
        # the halo exchange as peer-to-peer stores: every "rank" maps the others' inboxes by plain device pointer
        inboxes = [w.mgpu_p2p_init(hcap, 1 << 12)[1] for w in ranks]
        for w in ranks:
            w.mgpu_p2p_connect(inbox_ptrs=inboxes)
    hbytes = single.mgpu_halo_slot_bytes(hcap)
    allhalo = torch.zeros(hbytes * R, dtype=torch.uint8, device="cuda")
    halo_total = 0
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
        # every owner updates its proxies and publishes the boundary ones; "all-gather" = all slots in one buffer
        for r, w in enumerate(ranks):
            w.setWorldTransforms(xf)
            if p2p:
                w.mgpu_p2p_export_halo()          # all exports are enqueued before any import starts to wait
            else:
                w.mgpu_update_export_halo(allhalo.data_ptr() + r * hbytes, hcap)
        torch.cuda.synchronize()
        if not p2p:
            halo_total += int(sum(int(allhalo[r * hbytes:r * hbytes + 4].view(torch.int32)[0]) for r in range(R)))
        for w in ranks:
            if p2p:
                w.mgpu_p2p_import_halo()
            else:
                w.mgpu_import_halo(allhalo.data_ptr(), R, hcap)
            w.mgpu_broadphase()
        if p2p:
            for w in ranks:
                w.mgpu_p2p_export_departed()      # every rank pushes its departed manifolds into all the others' inboxes
            torch.cuda.synchronize()
            tot = 1                               # counted on the device only; the bit-exact union below is the check
            for w in ranks:
                w.mgpu_p2p_import_arrivals()
        elif slots:
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
    if stats is not None:
        stats["halo_records"] = halo_total
        stats["pairs_per_rank"] = sizes
        stats["owner"] = ranks[0].partition()[2]
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


def test_partitioned_world_with_peer_to_peer_halo(gpu_pkg):
    """The halo exchange fused into the export kernel (stores into the peers' inboxes, epoch flags) gives the same union as
    the all-gather path: spheres world over 4 ranks with migration, and the bin with its large statics over 3."""
    sc = scenes.spheres_scene(n=20000, seed=23)
    sc.vel *= 6.0
    moved = run_partitioned(gpu_pkg, sc, R=4, steps=5, slots=True, p2p=True)
    assert moved > 0
    sc = scenes.bin_scene(n=4000, seed=29)
    sc.vel *= 3.0
    run_partitioned(gpu_pkg, sc, R=3, steps=4, slots=True, p2p=True)


def test_peer_to_peer_halo_needs_a_connection(gpu_pkg):
    sc = scenes.spheres_scene(n=500, seed=3)
    w = scenes.build_gpu(gpu_pkg, sc, mode=1)
    w.set_partition(0, 2)
    with pytest.raises(Exception, match="not connected"):
        w.mgpu_p2p_export_halo()


def test_a_peer_that_never_publishes_is_reported(gpu_pkg):
    """Failure detection of the peer-to-peer exchange: rank 0 waits for rank 1's epoch flag, rank 1 never exports; after ~10 s
    the wait gives up and the step's counters come back as an error instead of a hang."""
    sc = scenes.spheres_scene(n=2000, seed=5)
    ranks = [scenes.build_gpu(gpu_pkg, sc, mode=1) for _ in range(2)]
    for r, w in enumerate(ranks):
        w.set_partition(r, 2)
    inboxes = [w.mgpu_p2p_init(4096, 256)[1] for w in ranks]
    for w in ranks:
        w.mgpu_p2p_connect(inbox_ptrs=inboxes)
    w0 = ranks[0]
    w0.setWorldTransforms(sc.transforms(0))
    w0.mgpu_p2p_export_halo()
    w0.mgpu_p2p_import_halo()          # rank 1 stays silent
    w0.mgpu_broadphase()
    with pytest.raises(gpu_pkg.B2CError, match="timed out"):
        w0.sync_counts()


def test_slab_ownership_and_halo_size(gpu_pkg):
    """Explicit planes: proxies are owned by the slab their origin lies in, only boundary proxies travel, and every rank
    ends up with a share of the pairs."""
    sc = scenes.spheres_scene(n=27000, seed=9)   # 30^3 lattice
    lo, hi = float(sc.base[:, 11].min()), float(sc.base[:, 11].max())
    planes = [lo + (hi - lo) * k / 3.0 for k in (1, 2)]
    st = {}
    run_partitioned(gpu_pkg, sc, R=3, steps=4, axis=2, planes=planes, stats=st)
    owner = st["owner"]
    z = sc.base[:, 11]
    expect = (z >= np.float32(planes[0])).astype(np.uint8) + (z >= np.float32(planes[1])).astype(np.uint8)
    assert np.array_equal(owner, expect)
    # 2 interior faces, ~900 proxies per lattice layer on either side of a face, 4 steps: far fewer than all proxies
    assert 4 * 900 < st["halo_records"] < 4 * 8 * 900
    assert min(st["pairs_per_rank"]) > 0.2 * max(st["pairs_per_rank"])


def test_halo_slot_overflow_is_reported(gpu_pkg):
    import torch
    sc = scenes.spheres_scene(n=8000, seed=10)
    w = scenes.build_gpu(gpu_pkg, sc, mode=1)
    w.set_partition(1, 2)   # the plane is the median origin: the lattice layer just above it reaches across
    hcap = 8
    hb = w.mgpu_halo_slot_bytes(hcap)
    buf = torch.zeros(2 * hb, dtype=torch.uint8, device="cuda")
    w.setWorldTransforms(sc.transforms(0))
    w.mgpu_update_export_halo(buf.data_ptr(), hcap)
    w.mgpu_import_halo(buf.data_ptr(), 2, hcap)
    w.mgpu_broadphase()
    with pytest.raises(gpu_pkg.B2CError) as e:
        w.sync_counts()
    assert e.value.code == -3 and "halo" in str(e.value)


def test_pair_calculation_needs_the_halo_first(gpu_pkg):
    sc = scenes.spheres_scene(n=2000, seed=11)
    w = scenes.build_gpu(gpu_pkg, sc, mode=1)
    w.set_partition(1, 2)
    with pytest.raises(gpu_pkg.B2CError) as e:
        w.step_device()
    assert e.value.code == -5
    w.set_partition(0, 1)      # a partition of one is the plain world again
    w.step_device()
    assert w.sync_counts()[0] > 0
