#!/usr/bin/env python
"""bench.py — collision-phase benchmark (BASELINE.json: "collision-phase ms/step at 100k bodies; overlap
pairs/s and contacts/s per GPU").

  python bench.py --gpus N --steps K --warmup W          # our CUDA path
  python bench.py --impl reference --steps K --warmup W   # the reference algorithm on host cores (CPU oracle port)

A "step" is one performDiscreteCollisionDetection (AABB update + broadphase pairs + narrowphase + manifolds).

Headline (`value`, `ms_per_step`, `e2e`, `roofline`): the C2 workload — 100 000 mixed boxes / spheres / 16-point hulls in
a closed bin of 5 static boxes (tests/scenes.py:bin_scene, seeded, settled snapshot; synthetic transform trace because the
solver/integrator are not on the path).  N>1: one independent replica of that world per GPU (worlds never interact -> no
data-path collective, weak scaling); `value` is whole-job world-steps per second, `ms_per_step` the per-step device time
(max over ranks).

The same line also carries the two SHARDED configs of BASELINE.json (strong scaling, total work fixed as N grows):
  "c4": 4096 independent 64-body worlds split by world over the N GPUs (no collective),
  "c5": ONE world of 1 M spheres partitioned by slabs of space with a halo; boundary AABBs + transforms and migrating
        manifolds are stored by the library's own kernels straight into the peers' inboxes over NVLink (CUDA IPC; no
        collective in the step — B2C_HALO=nccl selects the two all-gathers instead); with `check` = the union of the
        ranks' pair lists / manifolds compared against a single-GPU run of the same step inside this very run.
One JSON line on stdout from rank 0.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

FRAMES = 8  # distinct transform frames of the trace, traversed back and forth
C2_SEED = 100


def log(*a):
    print(*a, file=sys.stderr, flush=True)


_RESULT_FD = None


def claim_stdout():
    """The contract is ONE JSON line on stdout.  Libraries print there too (NCCL's version banner, for one), so the real
    stdout is set aside for the result line and fd 1 is pointed at stderr for everything else."""
    global _RESULT_FD
    if _RESULT_FD is None:
        sys.stdout.flush()
        _RESULT_FD = os.dup(1)
        os.dup2(2, 1)


def emit(obj):
    line = (json.dumps(obj) + "\n").encode()
    if _RESULT_FD is None:
        sys.stdout.write(line.decode())
        sys.stdout.flush()
    else:
        os.write(_RESULT_FD, line)


def make_scene(n_bodies, seed, workload="c2", worlds=4096):
    import scenes
    if workload == "c4":   # independent 64-body dice worlds
        return scenes.worlds_scene(num_worlds=worlds, seed=seed)
    if workload == "c3":   # 708 x 708 cells = 1 002 528 triangles, 50 % hull-16 / 50 % spheres r=0.4
        cells = max(8, int(round(708 * (n_bodies / 10000.0) ** 0.5)))
        return scenes.terrain_scene(cells=cells, n=n_bodies, seed=seed, mix=(0.5, 0.5, 0.0))
    if workload == "c5":   # spheres r=0.5 at 40 % packing, one world
        return scenes.spheres_scene(n=n_bodies, seed=seed)
    # 49 x 49 footprint at 0.82 spacing = the 40 x 40 bin of SURVEY §8d C2
    return scenes.bin_scene(n=n_bodies, seed=seed, footprint=max(4, int(round((n_bodies / 100000.0) ** 0.5 * 49))))


def workload_name(wl, n_bodies, worlds):
    """The `config.workload` string; both arms print exactly this for the same workload."""
    if wl == "c2":
        return (f"C2: {n_bodies} mixed boxes/spheres/16-pt hulls in a closed bin of 5 static boxes, one world per GPU, "
                "DbvtBroadphase pair semantics, seeded transform trace")
    if wl == "c3":
        tris = 2 * max(8, int(round(708 * (n_bodies / 10000.0) ** 0.5))) ** 2
        return (f"C3: {n_bodies} convex hulls (16 pts) and spheres on a {tris}-triangle BvhTriangleMeshShape heightfield "
                "(quantized BVH), one world per GPU")
    if wl == "c4":
        return f"C4: {worlds} independent 64-body dice worlds, split by world over the GPUs"
    return f"C5: {n_bodies} spheres r=0.5 at 40% packing, ONE world partitioned by sorted-AABB slabs with a halo over the GPUs"


def settle_scene(pkg, sc, dev, max_pairs, iters, log_fn=None):
    """Turn the jittered lattice into a SETTLED snapshot (SURVEY §8d C2 asks for one): a Jacobi position
    relaxation that pushes every pair apart along its contact normal by a fraction of its penetration, run
    with the CUDA path itself as the contact generator.  Only sc.base changes; the measurement then runs on a
    fresh world, and the CPU baseline gets the same relaxed transforms."""
    import scenes
    gw = scenes.build_gpu(pkg, sc, mode=pkg.DBVT, max_pairs=max_pairs, device=dev)
    stat = np.asarray(sc.static, dtype=bool)
    pos = sc.base[:, 9:].astype(np.float64)
    xf = sc.base.copy()
    for it in range(iters):
        xf[:, 9:] = pos.astype(np.float32)
        gw.setWorldTransforms(xf)
        gw.step()
        hdr, pts = gw.contacts()
        if len(hdr) == 0:
            break
        order = np.argsort(hdr["first_point"], kind="stable")
        owner = order[np.repeat(np.arange(len(hdr)), hdr["num_contacts"][order])]
        d = pts["distance"].astype(np.float64)
        pen = np.minimum(d + 0.005, 0.0)            # leave a small resting penetration
        n = pts["normal_on_b"].astype(np.float64)
        b0 = hdr["body0"][owner] - 1
        b1 = hdr["body1"][owner] - 1
        w0 = np.where(stat[b0], 0.0, 1.0)
        w1 = np.where(stat[b1], 0.0, 1.0)
        wsum = np.maximum(w0 + w1, 1.0)
        push = (-pen * 0.6)[:, None] * n
        disp = np.zeros_like(pos)
        cnt = np.zeros(len(pos))
        np.add.at(disp, b0, push * (w0 / wsum)[:, None])
        np.add.at(disp, b1, -push * (w1 / wsum)[:, None])
        np.add.at(cnt, b0, (pen < 0) * 1.0)
        np.add.at(cnt, b1, (pen < 0) * 1.0)
        pos += disp / np.maximum(cnt, 1.0)[:, None] * np.minimum(cnt, 2.0)[:, None]
        if log_fn and (it % 10 == 0 or it == iters - 1):
            st = gw.stats()
            log_fn(f"  settle {it}: pairs {st['num_pairs']} contacts {len(pts)} deep {st['deep_penetration_checks']} "
                   f"max pen {-d.min():.3f}")
    sc.base[:, 9:] = pos.astype(np.float32)
    sc.base[stat] = xf[stat]
    gw.close()
    return sc


def settled_path(n_bodies, seed, iters):
    return os.path.join(ROOT, "tests", "golden", f"c2_settled_n{n_bodies}_seed{seed}_it{iters}.npz")


def load_settled(sc, n_bodies, seed, iters):
    """Apply the committed settled snapshot (origins only; made by `bench.py --save-settled` with settle_scene) so
    that both arms measure the same scene and the reference arm needs no GPU.  Returns False when there is none."""
    p = settled_path(n_bodies, seed, iters)
    if not os.path.exists(p):
        return False
    z = np.load(p)
    if z["pos"].shape != sc.base[:, 9:].shape:
        return False
    sc.base[:, 9:] = z["pos"]
    return True


def c2_scene(args, pkg=None, dev=0, rank=0):
    """The headline scene: EVERY rank (and the reference arm) steps the same seeded, settled 100k-body world."""
    sc = make_scene(args.bodies, seed=C2_SEED, workload="c2")
    settled = False
    if args.settle > 0:
        settled = (not args.save_settled) and load_settled(sc, args.bodies, C2_SEED, args.settle)
        if not settled and pkg is not None:
            sc = settle_scene(pkg, sc, dev, args.max_pairs, args.settle, log if rank == 0 else None)
            settled = True
            if args.save_settled and rank == 0:
                os.makedirs(os.path.dirname(args.save_settled) or ".", exist_ok=True)
                np.savez_compressed(args.save_settled, pos=sc.base[:, 9:])
        if settled:
            sc.vel *= 0.25  # a settled pile creeps; it does not drift
    return sc, settled


def frame_index(step):
    period = 2 * (FRAMES - 1)
    k = step % period
    return k if k < FRAMES else period - k


class ClockSampler:
    """nvidia-smi clock / throttle-reason sampling during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, gpu_index):
        self.rows = []
        self.proc = None
        self.gpu = gpu_index

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); smax.append(float(f[1]))
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": float(max(smax)) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def algorithmic_bytes(stage, N, P, st, key_passes_body, contacts):
    """Algorithmic bytes one launch group moves.  N proxies, P pairs."""
    G = st["gjk_checks"]
    if stage == "aabb":
        return N * (48 + 4 + 1 + 32)
    if stage == "bounds_keys":
        return N * (32 + 1 + 4 + 8)
    if stage == "sort_proxies":
        return N * 4 + key_passes_body * 2 * 8 * N      # 32-bit key (row | qx) + 32-bit payload
    if stage == "gather":
        return N * (8 + 32 + 4 + 32 + 4)
    if stage == "sweep":
        return 9 * N * 4 + N * 32 + 8 * P
    if stage == "large":
        return 0   # a handful of proxies against the sorted AABBs the sweep has just read (L2); hidden beside k_sweep
    if stage == "sort_pairs":
        # pair_rows.cuh: scan of N row counters (r 4N, w 4N), scatter (r 8P, w 4P), row sort (r 4P, w 8P pairs + 8P keys)
        return 8 * N + P * (8 + 4 + 4 + 16)
    if stage == "unpack_carry":
        # k_carry: keys + search in last step's keys, header copy r/w, live point copy r/w
        return P * (8 + 8) + 2 * (32 * P + 96 * contacts)
    if stage == "classify_bin":
        return P * (8 + 2 * (1 + 4 + 4) + 1 + 4) + P * (1 + 4)
    if stage == "closed_form":
        return 0
    if stage == "gjk_mesh":
        # pair 8 B + 2 transforms 96 B + 2 shape records 128 B + manifold header r/w 64 B + raw record 56 B
        return G * (8 + 96 + 128 + 64 + 56) + 96 * contacts * 2
    if stage == "epa_fold_count":
        return st["deep_penetration_checks"] * (120 + 96 + 128 + 56) + P * 8
    return 0


def max_pairs_for(args, wl):
    """Pair capacity per workload: the first C5 step (creation AABBs -> every proxy fattened) has ~9 pairs per sphere."""
    if wl == "c5":
        return max(args.max_pairs, min((1 << 24) - 1, 10 * args.c5_bodies if args.workload == "c2" else 10 * args.bodies))
    return args.max_pairs


BROADPHASE_STAGES = ("aabb", "bounds_keys", "sort_proxies", "gather", "sweep", "large", "sort_pairs")


class Rig:
    """Per-process CUDA plumbing shared by the workloads of one bench run."""

    def __init__(self, args):
        import torch
        import __graft_entry__ as ge
        self.torch = torch
        self.args = args
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.dist = None
        if self.world > 1:
            import torch.distributed as dist
            torch.cuda.set_device(self.local_rank)
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local_rank))
            self.dist = dist
        self.dev = self.local_rank if self.world > 1 else 0
        torch.cuda.set_device(self.dev)
        self.bind_cpus()
        self.pkg = ge.load_package()
        self.L = self.pkg._lib.load()
        if self.L.b2c_device_count() < 1:
            raise SystemExit("no sm_100 device: the CUDA path cannot run and there is no CPU fallback")
        self.flush = torch.empty(192 * 1024 * 1024 // 4, dtype=torch.float32, device=f"cuda:{self.dev}")  # 192 MiB > 126 MB L2

    def bind_cpus(self):
        """Give every rank its own slice of the host cores its GPU is attached to (the GPU's NUMA-local CPU set from NVML when
        it is available, else all cores): the pinned staging buffers are then allocated and touched on that node and the
        ranks do not share cores."""
        self.cpu_binding = None
        try:
            allowed = sorted(os.sched_getaffinity(0))
            local = allowed
            try:
                import pynvml
                pynvml.nvmlInit()
                h = pynvml.nvmlDeviceGetHandleByIndex(self.dev)
                words = pynvml.nvmlDeviceGetCpuAffinity(h, (max(allowed) + 64) // 64)
                near = [c for c in allowed if (words[c // 64] >> (c % 64)) & 1]
                if near:
                    local = near
            except Exception:
                pass
            if self.world > 1:
                per = max(1, len(local) // self.world)
                mine = local[(self.local_rank * per) % len(local):][:per] or local
            else:
                mine = local
            os.sched_setaffinity(0, mine)
            self.cpu_binding = f"{len(mine)} cores [{mine[0]}..{mine[-1]}]"
        except Exception as e:  # not fatal: run unbound
            self.cpu_binding = f"unbound ({type(e).__name__})"

    def barrier(self):
        if self.dist is not None:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def reduce(self, maxes, sums):
        torch = self.torch
        v = torch.tensor(maxes, dtype=torch.float64, device=f"cuda:{self.dev}")
        s = torch.tensor(sums, dtype=torch.float64, device=f"cuda:{self.dev}")
        if self.dist is not None:
            self.dist.all_reduce(v, op=self.dist.ReduceOp.MAX)
            self.dist.all_reduce(s, op=self.dist.ReduceOp.SUM)
        return [float(x) for x in v.tolist()], [float(x) for x in s.tolist()]


def measure(rig, wl, sc, steps, warmup, e2e_steps, want_stages, partitioned=False, sampler=None):
    """Time `steps` collision steps of scene `sc` on this rank's GPU: device-resident arm, optional per-stage pass, and the
    end-to-end arm through the C ABI with pinned host buffers.  Returns this rank's numbers (reduction is the caller's)."""
    torch, pkg, L, args, dev = rig.torch, rig.pkg, rig.L, rig.args, rig.dev
    import scenes
    t0 = time.time()
    P_cap = max_pairs_for(args, wl)
    gw = scenes.build_gpu(pkg, sc, mode=pkg.DBVT, max_pairs=P_cap, device=dev, raw_records=False)
    nb = sc.n
    stream = torch.cuda.ExternalStream(gw.stream(), device=torch.device("cuda", dev))
    mg = None
    if partitioned:
        mg = pkg.PartitionedStepper(gw, rig.rank, rig.world, rig.dist, torch, dev)

    def one_step():
        if mg is not None:
            mg.step()
        else:
            gw.step_device()

    frames = [np.ascontiguousarray(pkg.transforms_to_planes(sc.transforms(k))) for k in range(FRAMES)]
    dframes = [torch.from_numpy(f).to(f"cuda:{dev}") for f in frames]
    hframes = [torch.from_numpy(f).pin_memory() for f in frames]
    log(f"[rank {rig.rank}] {wl}: {nb} proxies built in {time.time() - t0:.1f}s")
    flush = rig.flush
    torch.cuda.synchronize()

    # ---------------- device-resident arm ----------------
    # the timed steps run WITHOUT per-stage events (so the library may replay the step as one CUDA graph); the per-stage
    # times the roofline is computed from come from a short profiled pass afterwards
    step_no = 0
    for _ in range(warmup):
        gw.setWorldTransformsDevice(nb, dframes[frame_index(step_no)].data_ptr())
        one_step()
        gw.sync_counts()
        step_no += 1
    if mg is not None and mg.halo_mode != "p2p":
        # (all-gather path only: the peer-to-peer path sends records, not fixed-size slots)
        # one full pass over the trace with the worst-case slots, then size them to what this world sends (+50 %)
        seen_h = seen_m = 0
        for _ in range(2 * (FRAMES - 1)):
            gw.setWorldTransformsDevice(nb, dframes[frame_index(step_no)].data_ptr())
            one_step()
            gw.sync_counts()
            torch.cuda.synchronize()
            seen_h = max(seen_h, int(mg.all_halo.view(torch.int32)[:: mg.halo_bytes // 4][: rig.world].max().item()))
            seen_m = max(seen_m, int(mg.all_slots.view(torch.int32)[:: mg.slot_bytes // 4][: rig.world].max().item()))
            step_no += 1
        mg.all_halo.view(torch.int32)[0] = seen_h      # tune_caps reads the maxima from the gathered slots
        mg.all_slots.view(torch.int32)[0] = seen_m
        mg.tune_caps()
        for _ in range(3):
            gw.setWorldTransformsDevice(nb, dframes[frame_index(step_no)].data_ptr())
            one_step()
            gw.sync_counts()
            step_no += 1
    rig.barrier()
    if sampler is not None:
        sampler.start()
    ev0 = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
    ev1 = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
    pairs_tot = contacts_tot = manif_tot = launches = 0
    st = None
    rig.barrier()
    for k in range(steps):
        with torch.cuda.stream(stream):
            flush.fill_(float(k))  # L2 flush between timed iterations (not timed)
            ev0[k].record(stream)  # inputs already resident in HBM: the frame is a device tensor
        if args.profile_step and k == 0 and wl == args.profile_workload:
            torch.cuda.synchronize()
            torch.cuda.profiler.start()   # ncu --profile-from-start off: capture exactly one timed step
        gw.setWorldTransformsDevice(nb, dframes[frame_index(step_no)].data_ptr())
        one_step()
        with torch.cuda.stream(stream):
            ev1[k].record(stream)
        p, m, c = gw.sync_counts()
        if args.profile_step and k == 0 and wl == args.profile_workload:
            torch.cuda.profiler.stop()
        pairs_tot += p; manif_tot += m; contacts_tot += c
        st = gw.stats()
        launches += st["kernel_launches"] + (mg.extra_launches if mg is not None else 0)
        step_no += 1
    rig.barrier()
    ms_per_step = float(np.mean([ev0[k].elapsed_time(ev1[k]) for k in range(steps)]))

    stage_ms = {}
    gjk_ms = 0.0
    if want_stages:
        gw.set_profiling(True)
        prof_steps = max(3, min(steps, 10))
        for k in range(prof_steps):
            with torch.cuda.stream(stream):
                flush.fill_(float(k))
            gw.setWorldTransformsDevice(nb, dframes[frame_index(step_no)].data_ptr())
            one_step()
            gw.sync_counts()
            for nm, ms in gw.stage_times().items():
                stage_ms[nm] = stage_ms.get(nm, 0.0) + ms / prof_steps
            gjk_ms += gw.gjk_kernel_ms() / prof_steps
            step_no += 1
        gw.set_profiling(False)
        rig.barrier()

    # ---------------- end-to-end arm through the C ABI with HOST buffers ----------------
    # per step: H2D of the 12 transform planes from pinned memory, the step, D2H of what a host consumes every step —
    # the pair-cache DELTAS (added / removed pairs: what HashedOverlappingPairCache reports to its callbacks; the host
    # mirror applies them) and the packed contact stream (uid-keyed 16-B headers + 48-B solver points)
    e2e_ms, d2h = [], 0
    if e2e_steps > 0:
        D_cap = P_cap // 2
        hdr_host = torch.empty((P_cap, 4), dtype=torch.int32).pin_memory()
        pts_host = torch.empty((2 * P_cap, 12), dtype=torch.int32).pin_memory()
        add_host = torch.empty((D_cap, 2), dtype=torch.int32).pin_memory()
        rem_host = torch.empty((D_cap, 2), dtype=torch.int32).pin_memory()
        gw.set_contact_prefetch(3)   # the uid-keyed packed contact stream is compacted inside the step's graph
        gw.set_pair_delta_prefetch(True)   # and so are the pair-cache add / remove events
        nA, nR, nH, nPt = ctypes.c_int32(), ctypes.c_int32(), ctypes.c_int32(), ctypes.c_int32()
        vp = ctypes.c_void_p
        rig.barrier()
        for k in range(e2e_steps + 2):
            with torch.cuda.stream(stream):
                flush.fill_(float(k))
            torch.cuda.synchronize()
            t_start = time.perf_counter()
            gw.setWorldTransformsHostPtr(nb, hframes[frame_index(step_no)].data_ptr())       # H2D of this step's inputs
            one_step()                                                                        # enqueue only, no host sync
            gw._ck(L.b2c_get_pair_deltas(gw.h, vp(add_host.data_ptr()), D_cap, vp(rem_host.data_ptr()), D_cap,
                                         ctypes.byref(nA), ctypes.byref(nR)))                 # D2H pair-cache events (while the narrowphase runs)
            gw._ck(L.b2c_begin_contact_download(gw.h, vp(hdr_host.data_ptr()), P_cap, vp(pts_host.data_ptr()), 2 * P_cap))
            gw.sync_counts()
            gw._ck(L.b2c_get_packed_contacts_uid(gw.h, vp(hdr_host.data_ptr()), P_cap, vp(pts_host.data_ptr()), 2 * P_cap,
                                                 ctypes.byref(nH), ctypes.byref(nPt)))        # D2H contact stream
            t_end = time.perf_counter()
            step_no += 1
            if k >= 2:
                e2e_ms.append((t_end - t_start) * 1e3)
                d2h = (nA.value + nR.value) * 8 + nH.value * 16 + nPt.value * 48 + 128
        rig.barrier()
    out = dict(ms=ms_per_step, e2e_ms=float(np.mean(e2e_ms)) if e2e_ms else 0.0, pairs=pairs_tot, contacts=contacts_tot,
               manifolds=manif_tot, launches=launches, stats=st, stage_ms=stage_ms, gjk_ms=gjk_ms, d2h=int(d2h), nb=nb, steps=steps,
               mg=(mg.describe() if mg is not None else None))
    if mg is not None:
        mg.close()
    gw.close()
    del dframes, hframes
    torch.cuda.empty_cache()
    if partitioned and args.check:
        # parity inside the REAL multi-rank run: union over the ranks == one GPU, step by step (fresh worlds, NCCL exchange)
        chk = pkg.partition_check(pkg, lambda g: pkg.PartitionedStepper(g, rig.rank, rig.world, rig.dist, torch, dev),
                                  lambda: scenes.build_gpu(pkg, sc, mode=pkg.DBVT, max_pairs=P_cap, device=dev, raw_records=False),
                                  frames, rig.rank, rig.world, rig.dist, steps=4)
        if chk is not None:
            out["check"] = chk
        torch.cuda.empty_cache()
    return out


def run_ours(args):
    rig = Rig(args)
    rank, world, ngpu = rig.rank, rig.world, args.gpus
    wl = args.workload
    sampler = ClockSampler(rig.dev) if rank == 0 else None

    # ---------------- headline workload ----------------
    settled = False
    if wl == "c2":
        sc, settled = c2_scene(args, rig.pkg, rig.dev, rank)
    elif wl == "c4":
        sc = make_scene(0, seed=100 + rank, workload="c4", worlds=max(1, args.worlds // world))
    elif wl == "c5":
        sc = make_scene(args.bodies, seed=100, workload="c5")
    else:
        sc = make_scene(args.bodies, seed=100 + rank, workload=wl)
    strong = wl in ("c4", "c5")
    r = measure(rig, wl, sc, args.steps, args.warmup, max(3, min(args.steps, 20)), True,
                partitioned=(wl == "c5" and world > 1), sampler=sampler)
    (ms_max, e2e_max), (pairs_all, contacts_all, manif_all, launches_all) = rig.reduce(
        [r["ms"], r["e2e_ms"]], [r["pairs"], r["contacts"], r["manifolds"], r["launches"]])
    clocks = sampler.stop() if sampler is not None else None

    # ---------------- the sharded configs ride in the same line (default run only) ----------------
    extra = {}
    if wl == "c2" and not args.no_sharded:
        for w2 in ("c4", "c5"):
            if w2 == "c4":
                sc2 = make_scene(0, seed=100 + rank, workload="c4", worlds=max(1, args.worlds // world))
            else:
                sc2 = make_scene(args.c5_bodies, seed=100, workload="c5")
            r2 = measure(rig, w2, sc2, args.sharded_steps, 3, 5, True, partitioned=(w2 == "c5" and world > 1))
            (m2, e2), (p2, c2, mf2, l2) = rig.reduce([r2["ms"], r2["e2e_ms"]], [r2["pairs"], r2["contacts"], r2["manifolds"], r2["launches"]])
            div = 1 if w2 == "c5" and world > 1 else 1   # c5 ranks own disjoint pair sets: the sums are whole-world totals
            extra[w2] = {
                "workload": workload_name(w2, args.c5_bodies, args.worlds), "scaling": "strong", "n_gpus": ngpu,
                "ms_per_step": m2, "steps_per_s": 1000.0 / m2, "e2e_ms_per_step": e2, "steps": args.sharded_steps,
                "proxies_per_rank": r2["nb"], "pairs_per_step_total": p2 / args.sharded_steps / div,
                "contacts_added_per_step_total": c2 / args.sharded_steps / div, "gpu_launches": int(l2),
                "stage_ms_rank0": {k: round(v, 5) for k, v in r2["stage_ms"].items()},
                "broadphase_ms_rank0": round(sum(r2["stage_ms"].get(s, 0.0) for s in BROADPHASE_STAGES), 5),
                "collective": (r2["mg"] if r2["mg"] else "none (worlds / the single rank share nothing)"),
            }
            if "check" in r2:
                extra[w2]["check"] = r2["check"]
            if w2 == "c5":
                # broadphase moves 180 B/proxy + 8 B/pair
                N5, P5 = args.c5_bodies, p2 / args.sharded_steps
                bp_ms = sum(r2["stage_ms"].get(s, 0.0) for s in BROADPHASE_STAGES)
                extra[w2]["broadphase_hbm"] = {"algorithmic_bytes": 180 * N5 + 8 * P5, "ms_rank0": bp_ms,
                                               "note": "rank 0's share of the slab-partitioned broadphase" if world > 1 else "whole world"}

    if rank != 0:
        if rig.dist is not None:
            rig.dist.destroy_process_group()
        return
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak_gbs = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    st, stage_ms, nb = r["stats"], r["stage_ms"], r["nb"]
    P_avg = r["pairs"] / args.steps
    contacts_live = r["contacts"] / args.steps
    body_bits = 12 + min(20, int(np.ceil(np.log2(2 * nb + 66))))
    abytes = {s: algorithmic_bytes(s, nb, P_avg, st, (body_bits + 7) // 8, contacts_live) for s in stage_ms}
    frac = {s: (abytes[s] / (stage_ms[s] * 1e-3) / 1e9) / peak_gbs if stage_ms[s] > 0 else 0.0 for s in stage_ms}
    # Two rooflines.  The DOMINANT kernel of the step is k_gjk (the GJK iterations of every convex-convex pair that
    # survives the prefilter): FP32-issue / divergence bound, DRAM nearly idle — so its roofline is instruction issue: thread
    # instructions per launch (a property of the workload, counted once by ncu on this same seeded snapshot:
    # profiles/kernel_metrics.json) / the launch time measured LIVE here with its own CUDA-event pair, against
    # 148 SMs x 4 schedulers x 32 lanes x the SM clock sampled during the run.  The HBM-shaped part of the step is the
    # broadphase (sort / sweep / pair ordering): bytes / time against the measured copy bandwidth, reported beside it.
    bp_ms = sum(stage_ms.get(s, 0.0) for s in BROADPHASE_STAGES)
    bp_bytes = 180 * nb + 8 * P_avg            # 180 B/proxy + 8 B/pair
    dom = max(stage_ms, key=stage_ms.get)
    km = {}
    try:
        km = json.load(open(os.path.join(ROOT, "profiles", "kernel_metrics.json")))
    except Exception:
        pass
    traffic = {}
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
    except Exception:
        pass
    # the counted instructions belong to the default snapshot only (100 000 bodies, settled, seed 100)
    gk = km.get("kernels", {}).get("k_gjk") if (wl == "c2" and args.bodies == 100000 and settled) else None
    sm_mhz = (clocks or {}).get("sm_mhz") or (clocks or {}).get("sm_max_mhz") or 1965.0
    issue_peak = 148 * 4 * 32 * sm_mhz * 1e6 / 1e12          # T thread-instructions / s
    gjk_ms = r.get("gjk_ms", 0.0)
    hbm_obj = {"bound": "hbm", "kernel": "broadphase (k_aabb .. pair rows: sort / sweep / pair ordering)",
               "achieved": bp_bytes / (bp_ms * 1e-3) / 1e9 if bp_ms > 0 else 0.0, "peak": peak_gbs, "unit": "GB/s",
               "frac": (bp_bytes / (bp_ms * 1e-3) / 1e9) / peak_gbs if bp_ms > 0 else 0.0,
               "traffic": traffic.get("broadphase"), "peak_source": peak_src,
               "algorithmic_bytes_per_launch": bp_bytes, "ms": bp_ms,
               "formula": "SURVEY 8(d): 180 B/proxy + 8 B/pair over the summed CUDA-event time of the broadphase stages",
               "all_stages_frac": {s_: round(frac[s_], 5) for s_ in stage_ms},
               "note": "latency / issue bound at this size: the stages' own DRAM traffic (ncu, `traffic`) would take ~7 us at the HBM peak"}
    if gk and gjk_ms > 0:
        achieved = gk["thread_inst_per_step"] / (gjk_ms * 1e-3) / 1e12
        roofline = {"bound": "fp32_issue", "kernel": "k_gjk", "achieved": achieved, "peak": issue_peak, "unit": "T thread-inst/s",
                    "frac": achieved / issue_peak, "ms": gjk_ms, "share_of_step": gjk_ms / ms_max if ms_max > 0 else None,
                    "thread_inst_per_launch": gk["thread_inst_per_step"], "warp_inst_per_launch": gk["warp_inst_per_step"],
                    "lane_efficiency": round(gk["threads_per_inst"] / 32.0, 4), "threads_per_inst": gk["threads_per_inst"],
                    "issue_active_pct_ncu": gk["issue_active_pct"], "warps_active_pct_ncu": gk["warps_active_pct"],
                    "regs": gk["regs"], "traffic": gk["dram_bytes_per_step"],
                    "peak_source": f"148 SMs x 4 schedulers x 32 lanes x {sm_mhz:.0f} MHz (SM clock sampled under load)",
                    "counts_source": km.get("source"),
                    "formula": "thread instructions executed per k_gjk launch (ncu, workload property) / live CUDA-event time of the "
                               "launch; frac = issue-slot utilisation x lane efficiency",
                    "hbm": hbm_obj}
    else:
        roofline = dict(hbm_obj)
        roofline["dominant_stage"] = {"stage": dom, "ms": stage_ms[dom],
                                      "bound": "fp32_issue" if dom in ("gjk_mesh", "epa_fold_count") else "hbm",
                                      "hbm_frac": round(frac[dom], 5)}
    if wl == "c2":
        pl = "1 GPU" if ngpu == 1 else "1 replica of the world per GPU, no collective"
    elif wl == "c4":
        pl = f"{args.worlds // world} worlds per GPU, no collective"
    elif wl == "c5":
        pl = "1 GPU" if world == 1 else f"slab partition over {world} GPUs: {r['mg']}"
    else:
        pl = "1 world per GPU, no collective"
    out = {
        "metric": "collision_phase_world_steps_per_s_100k_bodies" if wl == "c2" else f"collision_phase_steps_per_s_{wl}",
        "value": (1000.0 / ms_max) if strong else ngpu * 1000.0 / ms_max,
        "unit": "steps/s (one step = full collision phase of a 100k-body world)" if wl == "c2" else "steps/s (one step = full collision phase of the whole workload)",
        "n_gpus": ngpu, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_max,
        "higher_is_better": True,
        "scaling": "strong" if strong else "weak",
        "vs_baseline": None,
        "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": workload_name(wl, args.bodies, args.worlds),
                   "snapshot": f"settled ({args.settle} relaxation iterations)" if settled else "raw jittered lattice (deep overlaps)",
                   "deep_penetration_checks_per_step": st["deep_penetration_checks"],
                   "proxies": nb, "l2": "flushed between timed iterations (192 MiB write)",
                   "raw_detector_records": "off = the library default (b2c_set_raw_records: an inspection channel of this library, not an output of the reference; the tests switch it on)",
                   "parallelism": pl, "cpu_binding_rank0": rig.cpu_binding},
        "pairs_per_s": pairs_all / args.steps / (ms_max * 1e-3),
        "contacts_per_s": contacts_all / args.steps / (ms_max * 1e-3),
        "pairs_per_step": pairs_all / args.steps / (1 if wl == "c5" else ngpu),
        "contacts_added_per_step": contacts_all / args.steps / (1 if wl == "c5" else ngpu),
        "manifolds_per_step": manif_all / args.steps / (1 if wl == "c5" else ngpu),
        "stage_ms": {k: round(v, 5) for k, v in stage_ms.items()},
        "gpu_launches": int(launches_all),
        "e2e": {"value": (1000.0 / e2e_max) if strong else ngpu * 1000.0 / e2e_max, "unit": "steps/s", "ms_per_step": e2e_max,
                "h2d_bytes_per_step": nb * 48, "d2h_bytes_per_step": int(r["d2h"]),
                "what": "b2c_set_transforms(pinned host planes) + b2c_step_device + b2c_get_pair_deltas (pairs added to / removed "
                        "from the pair cache, overlaps the narrowphase) + b2c_begin_contact_download (overlaps the penetration bin) "
                        "+ b2c_sync_counts + b2c_get_packed_contacts_uid (16-B uid-keyed manifold headers + 48-B solver points: world "
                        "points on A and B, normal, distance, lifetime, warm-start slot, triangle index), all into pinned host buffers"},
        "roofline": roofline,
        "clocks": clocks,
    }
    out.update(extra)
    if ngpu == 1 and not args.no_cpu and wl == "c2":
        out["cpu_baseline"] = cpu_baseline(sc, args)
    emit(out)
    if rig.dist is not None:
        rig.dist.destroy_process_group()


def probe_jvm():
    """the reference is Java — say at run time whether THIS box could run it."""
    import shutil
    exe = shutil.which("java")
    if not exe:
        return "no `java` on PATH of this box"
    try:
        v = subprocess.run([exe, "-version"], capture_output=True, text=True, timeout=20)
        return "java present: " + (v.stderr or v.stdout).strip().splitlines()[0]
    except Exception as e:
        return f"java present but not runnable ({type(e).__name__})"


def _cpu_world_worker(q, scene_args, n_steps, warmup, idx):
    """One reference world on one host core (child process)."""
    try:
        try:
            cores = sorted(os.sched_getaffinity(0))
            os.sched_setaffinity(0, [cores[idx % len(cores)]])
        except Exception:
            pass
        sc, _ = c2_scene(scene_args)
        times, pairs = cpu_steps(sc, n_steps, warmup)
        q.put((idx, times, pairs, sc.n))
    except Exception as e:  # surface the failure to the parent instead of hanging it
        q.put((idx, repr(e), 0, 0))


def cpu_steps(sc, n_steps, warmup=1):
    """Time the CPU oracle (a port of the reference's algorithm; single thread like the reference's world)."""
    import scenes
    ow = scenes.build_oracle(sc, 2)  # 2 = literal restatement of the reference's Dbvt tree broadphase (oracle/dbvt_literal.h)
    times, pairs = [], 0
    for k in range(warmup + n_steps):
        xf = sc.transforms(frame_index(k))
        t, p, m = ow.timed_step(xf)
        if k >= warmup:
            times.append(t)
            pairs = p
    return times, pairs


def cpu_worlds_parallel(args, n_worlds, n_steps, warmup):
    """N independent reference worlds, one process (= one core) each, stepped concurrently: the whole-job CPU rate that
    stands beside N GPUs stepping one world each.  Returns (ms per step as the MAX over the worlds' means, pairs, proxies)."""
    import multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_cpu_world_worker, args=(q, args, n_steps, warmup, i)) for i in range(n_worlds)]
    for p in procs:
        p.start()
    res = [q.get() for _ in procs]
    for p in procs:
        p.join()
    for r in res:
        if isinstance(r[1], str):
            raise RuntimeError("reference worker failed: " + r[1])
    ms = max(float(np.mean(r[1])) * 1e3 for r in res)
    return ms, res[0][2], res[0][3]


def cpu_baseline(sc, args):
    times, pairs = cpu_steps(sc, args.cpu_steps)
    ms = float(np.mean(times)) * 1e3
    return {"value": 1000.0 / ms, "unit": "steps/s", "ms_per_step": ms, "cores": 1, "kind": "port",
            "sample": f"{args.cpu_steps} full steps of the same {sc.n}-proxy C2 world after 1 warm-up (oracle/, g++ -O2, literal Dbvt "
                      "tree broadphase + GJK/EPA narrowphase, single thread: the reference steps one world on one thread)",
            "pairs": int(pairs), "host_cpus": os.cpu_count(), "jvm": probe_jvm()}


def run_reference(args):
    """The reference's own CPU implementation of the path on this box's host cores.  The reference is Java; where no JVM
    can run it (probed here, at run time) the arm times the C++ restatement in oracle/ (`kind: "port"`).  One world per
    requested GPU, one host core per world (the reference steps a world on one thread), all worlds concurrently; --steps
    and --warmup are honoured."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n_worlds = max(1, args.gpus)
    host = os.cpu_count() or 1
    conc = min(n_worlds, host)
    ms, pairs, nprox = cpu_worlds_parallel(args, conc, args.steps, args.warmup)
    # fewer cores than worlds: the remaining worlds would run in further rounds of the same duration
    rounds = (n_worlds + conc - 1) // conc
    v = n_worlds * 1000.0 / (ms * rounds)
    jvm = probe_jvm()
    out = {
        "impl": "reference",
        "metric": "collision_phase_world_steps_per_s_100k_bodies", "value": v,
        "unit": "steps/s (one step = full collision phase of a 100k-body world)",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms * rounds, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name("c2", args.bodies, args.worlds),
                   "snapshot": f"settled ({args.settle} relaxation iterations)" if args.settle > 0 and os.path.exists(
                       settled_path(args.bodies, C2_SEED, args.settle)) else "raw jittered lattice (deep overlaps)",
                   "proxies": nprox, "parallelism": f"{n_worlds} world(s), one host core each, {conc} concurrently"},
        "pairs_per_step": pairs,
        "cpu_baseline": {"value": v, "unit": "steps/s", "cores": conc, "kind": "port", "jvm": jvm,
                         "sample": f"{args.steps} full steps after {args.warmup} warm-up of {n_worlds} copy(ies) of the {nprox}-proxy C2 world, "
                                   f"one per host core; the Java reference itself: {jvm}; timed instead: the C++ restatement in oracle/ "
                                   "(literal Dbvt tree broadphase + GJK/EPA narrowphase + persistent manifolds)"},
        "e2e": {"value": v, "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(out)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--bodies", type=int, default=None)
    ap.add_argument("--max-pairs", type=int, default=3 << 20)
    ap.add_argument("--cpu-steps", type=int, default=4)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--workload", default="c2", choices=["c2", "c3", "c4", "c5"],
                    help="c2 = BASELINE headline (default; the line also carries c4 and c5 objects); c3 = convex bodies on a "
                         "1M-triangle BVH mesh; c4 = batched worlds split by world; c5 = one partitioned world")
    ap.add_argument("--worlds", type=int, default=4096)
    ap.add_argument("--c5-bodies", type=int, default=1000000)
    ap.add_argument("--sharded-steps", type=int, default=10, help="timed steps of the c4 / c5 objects in the default line")
    ap.add_argument("--no-sharded", action="store_true", help="skip the c4 / c5 objects")
    ap.add_argument("--check", action="store_true", default=True,
                    help="c5 on N>1 GPUs: compare the union over the ranks with a single-GPU step of the same world")
    ap.add_argument("--no-check", dest="check", action="store_false")
    ap.add_argument("--profile-step", action="store_true",
                    help="bracket the first timed step with cudaProfilerStart/Stop (for ncu --profile-from-start off)")
    ap.add_argument("--profile-workload", default=None, help="which workload --profile-step brackets (default: --workload)")
    ap.add_argument("--save-settled", default="", help="write the settled origins (npz) here; commit it under tests/golden/")
    ap.add_argument("--settle", type=int, default=60, help="relaxation iterations for the settled snapshot (0 = raw lattice)")
    args = ap.parse_args()
    if args.bodies is None:
        args.bodies = {"c2": 100000, "c3": 10000, "c4": 0, "c5": 1000000}[args.workload]
    if args.profile_workload is None:
        args.profile_workload = args.workload
    claim_stdout()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
