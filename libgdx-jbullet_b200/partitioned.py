"""One collision world partitioned over several GPUs (BASELINE config C5) — host side.

One process per GPU (torch.distributed, NCCL over NVLink).  Every call below is enqueue-only on the world's CUDA stream;
the NCCL collectives are enqueued on the same stream, so a step never synchronises with the host.  The device work is
libb2c.so's `b2c_mgpu_*` entry points (include/b2c.h, last section); this module only sequences them and owns the
exchange buffers.  `partition_check` proves, inside a real multi-rank run, that the union over the ranks equals what one
GPU computes for the same world (pairs exactly, manifolds bit for bit).
"""
import numpy as np

MASK64 = (1 << 64) - 1


def default_halo_cap(n_bodies):
    """Records per halo slot: the proxies within one box of a slab face — a few N^(2/3) for a compact world."""
    return max(4096, int(3.5 * float(n_bodies) ** (2.0 / 3.0)))


class PartitionedStepper:
    """Sequences one partitioned step (include/b2c.h, last section):
        updateAabbs of the owned proxies + boundary proxies into this rank's halo slot
        -> the boundary records reach the ranks whose slab they touch: by peer-to-peer stores fused into the export kernel
           (halo="p2p", default when every rank can map every inbox) or by ONE all-gather of fixed-size halo slots
        -> adoption of the records that touch this rank's slab, pair calculation over the local list
        -> manifolds of pairs that changed owner reach their new owner the same way (peer-to-peer push to every other
           rank, or a small migration slot + ONE all-gather) -> adoption
        -> narrowphase on the pairs this rank owns.
    planes = None lets the library cut the world at equal-count quantiles along its longest axis."""

    def __init__(self, gw, rank, nranks, dist, torch, dev, migrate_cap=2048, halo_cap=None, axis=None, planes=None, halo="p2p"):
        self.gw, self.rank, self.nranks, self.dist, self.torch = gw, rank, nranks, dist, torch
        # dev = "cpu": host-logic tests (gloo, a stand-in world object); everything else is a CUDA device index
        self.on_gpu = dev != "cpu"
        self.stream = torch.cuda.ExternalStream(gw.stream(), device=torch.device("cuda", dev)) if self.on_gpu else None
        self.mcap = int(migrate_cap)
        self.hcap = int(halo_cap) if halo_cap else default_halo_cap(gw.num_bodies)
        if planes is None:
            gw.set_partition(rank, nranks)
        else:
            gw.set_partition_slabs(rank, nranks, axis, planes)
        self.slot_bytes = gw.mgpu_slot_bytes(self.mcap)
        self.halo_bytes = gw.mgpu_halo_slot_bytes(self.hcap)
        d = f"cuda:{dev}" if self.on_gpu else "cpu"
        self.my_slot = torch.zeros(self.slot_bytes, dtype=torch.uint8, device=d)
        self.all_slots = torch.zeros(self.slot_bytes * nranks, dtype=torch.uint8, device=d)
        self.my_halo = torch.zeros(self.halo_bytes, dtype=torch.uint8, device=d)
        self.all_halo = torch.zeros(self.halo_bytes * nranks, dtype=torch.uint8, device=d)
        self.extra_launches = 0   # every kernel of a partitioned step is counted by the library (b2c_stats.kernel_launches)
        # halo exchange: peer-to-peer stores fused into the export kernel when every rank can map every inbox (CUDA IPC,
        # one node), else the all-gather.  B2C_HALO=nccl forces the collective (A/B measurements).
        import os
        self.halo_mode = "nccl"
        if halo == "p2p" and os.environ.get("B2C_HALO", "p2p") != "nccl":
            self.halo_mode = self._connect_p2p(d)

    def _connect_p2p(self, d):
        """Every rank publishes the IPC handle of its inbox (one small all-gather at set-up), maps the others', and all ranks
        agree (all-reduce of a success flag) on whether the peer-to-peer path is usable."""
        torch, gw = self.torch, self.gw
        ok, handle, ptr = 1, bytes(64), 0
        try:
            handle, ptr = gw.mgpu_p2p_init(self.hcap, self.mcap)
        except Exception as e:  # noqa: BLE001 — any failure means "use the collective", decided by all ranks together
            self.p2p_error, ok = repr(e), 0
        multi = self.dist is not None and self.nranks > 1
        # every rank takes part in both collectives below whatever happened to it, so that they always match up
        if multi:
            mine = torch.tensor(list(handle), dtype=torch.uint8, device=d)
            allh = torch.zeros(64 * self.nranks, dtype=torch.uint8, device=d)
            self.dist.all_gather_into_tensor(allh, mine)
        if ok:
            try:
                if multi:
                    gw.mgpu_p2p_connect(ipc_handles=bytes(allh.cpu().numpy().tobytes()))
                else:
                    gw.mgpu_p2p_connect(inbox_ptrs=[ptr] * self.nranks)
            except Exception as e:  # noqa: BLE001
                self.p2p_error, ok = repr(e), 0
        if multi:
            flag = torch.tensor([ok], dtype=torch.int32, device=d)
            self.dist.all_reduce(flag, op=self.dist.ReduceOp.MIN)
            ok = int(flag.item())
        return "p2p" if ok else "nccl"

    def all_gather(self, out, inp):
        """One NCCL all-gather on the world's stream; with a single rank it is a copy."""
        torch = self.torch
        import contextlib
        with (torch.cuda.stream(self.stream) if self.on_gpu else contextlib.nullcontext()):
            if self.dist is not None and self.nranks > 1:
                self.dist.all_gather_into_tensor(out, inp)
            else:
                out.copy_(inp)

    def step(self):
        gw = self.gw
        if self.halo_mode == "p2p":
            gw.mgpu_p2p_export_halo()      # k_aabb + boundary records stored straight into the neighbours' inboxes
            gw.mgpu_p2p_import_halo()      # wait for every source's epoch, adopt
        else:
            gw.mgpu_update_export_halo(self.my_halo.data_ptr(), self.hcap)
            self.all_gather(self.all_halo, self.my_halo)
            gw.mgpu_import_halo(self.all_halo.data_ptr(), self.nranks, self.hcap)
        gw.mgpu_broadphase()
        if self.halo_mode == "p2p":
            gw.mgpu_p2p_export_departed()  # manifolds of pairs that changed owner -> every other rank's migration inbox
            gw.mgpu_p2p_import_arrivals()
        else:
            gw.mgpu_export_departed_slot(self.my_slot.data_ptr(), self.mcap)
            self.all_gather(self.all_slots, self.my_slot)
            gw.mgpu_import_arrival_slots(self.all_slots.data_ptr(), self.nranks, self.mcap)
        gw.mgpu_narrowphase()

    def tune_caps(self, headroom=1.5, min_halo=1024, min_migrate=256):
        """Shrink the exchange slots to what the world actually sends: an all-gather moves whole fixed-size slots, so a slot
        sized for the worst case costs its full size in NVLink time every step.  Reads the record counts every rank published
        in the LAST step (first word of each gathered slot — the same values on every rank, so all ranks choose the same new
        capacities without another collective) and re-allocates the buffers with `headroom`.  A later overflow is still
        reported by b2c_sync_counts (B2C_ERR_CAPACITY); call this again, or raise the caps, when the world changes."""
        torch = self.torch
        if self.halo_mode == "p2p":
            return self.hcap, self.mcap   # records travel, not slots: the inbox capacities cost nothing per step
        torch.cuda.synchronize()
        halo = self.all_halo.view(torch.int32)[:: self.halo_bytes // 4][: self.nranks]
        mig = self.all_slots.view(torch.int32)[:: self.slot_bytes // 4][: self.nranks]
        hmax, mmax = int(halo.max().item()), int(mig.max().item())
        hcap = max(min_halo, int(hmax * headroom) + 64)
        mcap = max(min_migrate, int(mmax * headroom) + 64)
        if hcap >= self.hcap and mcap >= self.mcap:
            return self.hcap, self.mcap
        self.hcap, self.mcap = min(hcap, self.hcap), min(mcap, self.mcap)
        self.slot_bytes = self.gw.mgpu_slot_bytes(self.mcap)
        self.halo_bytes = self.gw.mgpu_halo_slot_bytes(self.hcap)
        d = self.my_slot.device
        self.my_slot = torch.zeros(self.slot_bytes, dtype=torch.uint8, device=d)
        self.all_slots = torch.zeros(self.slot_bytes * self.nranks, dtype=torch.uint8, device=d)
        self.my_halo = torch.zeros(self.halo_bytes, dtype=torch.uint8, device=d)
        self.all_halo = torch.zeros(self.halo_bytes * self.nranks, dtype=torch.uint8, device=d)
        return self.hcap, self.mcap

    def describe(self):
        if self.halo_mode == "p2p":
            return (f"slab partition; NO collective in the step: boundary proxies (80-byte records, inbox capacity {self.hcap} per "
                    f"source) and migrating manifolds (capacity {self.mcap} per source) are stored by this library's kernels "
                    f"straight into the inboxes of the ranks that need them (peer memory over NVLink, CUDA IPC), coalesced, with "
                    f"release/acquire epoch flags at system scope")
        return (f"slab partition, 2 ncclAllGather per step: halo slots ({self.hcap} x 80-byte boundary-proxy records = "
                f"{self.halo_bytes} B per rank, {self.halo_bytes * self.nranks} B gathered) and manifold-migration slots "
                f"({self.mcap} manifolds = {self.slot_bytes} B per rank, {self.slot_bytes * self.nranks} B gathered)")

    def close(self):
        self.my_slot = self.all_slots = self.my_halo = self.all_halo = None


def _mix(h):
    h = (h ^ (h >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    h = (h ^ (h >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return h ^ (h >> np.uint64(31))


def multiset_hash(rows_u64):
    """Order-independent 64-bit digest of the rows of a (n, w) uint64 array: sum of a mixed per-row hash (mod 2^64)."""
    a = np.ascontiguousarray(rows_u64, dtype=np.uint64)
    if a.size == 0:
        return 0
    with np.errstate(over="ignore"):
        w = (np.arange(a.shape[1], dtype=np.uint64) * np.uint64(2) + np.uint64(0x9E3779B97F4A7C15))
        h = _mix((a * w[None, :]).sum(axis=1, dtype=np.uint64) + np.uint64(a.shape[1]))
        return int(h.sum(dtype=np.uint64)) & MASK64


def world_digest(gw):
    """(pair count, pair hash, touching manifolds, contact points, manifold hash) of the world's last step.  The manifold
    hash covers the header (pair uids, body order, contact count, algorithm) and all 22 payload words of every live point
    (local / world positions, normal, distance, friction, restitution, lifetime, warm-start slot, triangle ids) from the
    compact contact stream, i.e. equality means bit-identical manifolds."""
    p = gw.pairs().astype(np.uint64)
    ph = multiset_hash(p)
    hdr, pts = gw.contacts()
    rows = np.zeros((len(hdr), 6 + 4 * 11), dtype=np.uint64)
    if len(hdr):
        for k, f in enumerate(("pair_uid0", "pair_uid1", "body0", "body1", "num_contacts", "algorithm")):
            rows[:, k] = hdr[f].astype(np.int64).astype(np.uint64)
        words = np.ascontiguousarray(pts).view(np.uint64).reshape(len(pts), 12)[:, :11]   # 96-byte records, last 8 bytes padding
        idx = hdr["first_point"].astype(np.int64)[:, None] + np.arange(4)[None, :]
        live = np.arange(4)[None, :] < hdr["num_contacts"][:, None]
        w = words[np.clip(idx, 0, max(len(pts) - 1, 0))]
        w[~live] = 0
        rows[:, 6:] = w.reshape(len(hdr), 44)
    return [int(len(p)), ph, int(len(hdr)), int(len(pts)), multiset_hash(rows)]


def partition_check(pkg, make_stepper, build_world, frames, rank, nranks, dist, steps=4):
    """Run `steps` steps of the same world twice inside this multi-rank job — partitioned over all ranks, and whole on rank
    0's GPU — and compare after every step: the ranks' pair lists must be disjoint and their union the single-GPU list, the
    union of their touching manifolds bit-identical to the single-GPU manifolds (order-independent 64-bit digests, summed
    over the ranks).  Returns a summary dict on rank 0 (raises AssertionError on any mismatch), None elsewhere.

    build_world() -> a fresh GpuCollisionWorld of the scene; make_stepper(gw) -> its PartitionedStepper;
    frames[k] = (12, n) float32 transform planes of step k."""
    gw = build_world()
    mg = make_stepper(gw)
    single = build_world() if rank == 0 else None
    report = {"steps": steps, "pairs": [], "touching_manifolds": [], "contact_points": [], "migrated_manifolds": "exercised by the moving trace"}
    for k in range(steps):
        planes = frames[k % len(frames)]
        gw.setWorldTransformPlanes(planes)
        mg.step()
        gw.sync_counts()
        d = world_digest(gw)
        if dist is not None and nranks > 1:
            allv = [None] * nranks
            dist.all_gather_object(allv, d)
        else:
            allv = [d]
        if rank == 0:
            tot = [sum(v[0] for v in allv), sum(v[1] for v in allv) & MASK64, sum(v[2] for v in allv), sum(v[3] for v in allv),
                   sum(v[4] for v in allv) & MASK64]
            single.setWorldTransformPlanes(planes)
            single.step_device()
            single.sync_counts()
            ref = world_digest(single)
            assert tot[0] == ref[0], f"step {k}: union of the ranks has {tot[0]} pairs, one GPU {ref[0]}"
            assert tot[1] == ref[1], f"step {k}: pair sets differ (digest)"
            assert tot[2] == ref[2] and tot[3] == ref[3], f"step {k}: touching manifolds / points {tot[2]}/{tot[3]} vs {ref[2]}/{ref[3]}"
            assert tot[4] == ref[4], f"step {k}: manifold contents differ (digest)"
            report["pairs"].append(ref[0]); report["touching_manifolds"].append(ref[2]); report["contact_points"].append(ref[3])
        if dist is not None and nranks > 1:
            dist.barrier()   # the other ranks wait for rank 0's single-GPU step here, not inside the next step's flag wait
    mg.close()
    gw.close()
    if single is not None:
        single.close()
    if rank != 0:
        return None
    report["result"] = "union over ranks == single GPU (pairs exact, manifolds bit-identical) at every step"
    report["per_rank_pairs_last_step"] = [v[0] for v in allv]
    return report
