#!/usr/bin/env python
"""bench.py — collision-phase benchmark (BASELINE.json: "collision-phase ms/step at 100k bodies; overlap
pairs/s and contacts/s per GPU").

  python bench.py --gpus N --steps K --warmup W          # our CUDA path
  python bench.py --impl reference --steps K --warmup W   # the reference algorithm on host cores (CPU oracle port)

A "step" is one performDiscreteCollisionDetection (AABB update + broadphase pairs + narrowphase + manifolds)
over the C2 workload: 100 000 mixed boxes / spheres / 16-point hulls in a closed bin of 5 static boxes
(tests/scenes.py:bin_scene, seeded; synthetic transform trace because the solver/integrator are not on the
path).  N>1: one independent 100k-body world per GPU (worlds never interact -> no data-path collective,
weak scaling); `value` is whole-job world-steps per second, `ms_per_step` the per-step device time (max over
ranks).  One JSON line on stdout from rank 0.
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
        owner = np.repeat(np.arange(len(hdr)), hdr["num_contacts"])
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


def algorithmic_bytes(stage, N, P, st, key_passes_body, key_passes_pair, contacts):
    """Algorithmic bytes one launch group moves.  N proxies, P pairs."""
    G = st["gjk_checks"]
    if stage == "aabb":
        return N * (48 + 4 + 1 + 32)
    if stage == "bounds_keys":
        return N * (32 + 1) + N * (32 + 1 + 4 + 8)
    if stage == "sort_proxies":
        return N * 4 + key_passes_body * 2 * 8 * N      # 32-bit key (row | qx) + 32-bit payload
    if stage == "gather":
        return N * (8 + 32 + 4 + 32 + 4)
    if stage == "sweep":
        return 9 * N * 4 + N * 32 + 8 * P
    if stage == "large":
        return st["large_proxies"] * N * 32
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


def run_ours(args):
    import torch
    import __graft_entry__ as ge
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    ngpu = args.gpus
    dist = None
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = local_rank if world > 1 else 0
    torch.cuda.set_device(dev)
    pkg = ge.load_package()
    L = pkg._lib.load()
    if L.b2c_device_count() < 1:
        raise SystemExit("no sm_100 device: the CUDA path cannot run and there is no CPU fallback")

    N = args.bodies
    t0 = time.time()
    wl = args.workload
    if wl == "c4":
        # strong scaling: the 4096-world batch is split by world over the ranks (no cross-world pairs, no collective)
        sc = make_scene(N, seed=100 + rank, workload="c4", worlds=max(1, args.worlds // world))
    elif wl == "c5":
        # strong scaling: ONE world, every rank holds all proxies and owns a slice of the sorted-AABB list
        sc = make_scene(N, seed=100, workload="c5")
    else:
        sc = make_scene(N, seed=100 + rank, workload=wl)
    import scenes
    if args.settle > 0 and wl == "c2":
        if args.save_settled or not load_settled(sc, N, 100 + rank, args.settle):
            sc = settle_scene(pkg, sc, dev, args.max_pairs, args.settle, log if rank == 0 else None)
            if args.save_settled and rank == 0:
                os.makedirs(os.path.dirname(args.save_settled) or ".", exist_ok=True)
                np.savez_compressed(args.save_settled, pos=sc.base[:, 9:])
        sc.vel *= 0.25  # a settled pile creeps; it does not drift
    gw = scenes.build_gpu(pkg, sc, mode=pkg.DBVT, max_pairs=args.max_pairs, device=dev)
    nb = sc.n
    partitioned = wl == "c5" and world > 1
    stream = torch.cuda.ExternalStream(gw.stream(), device=torch.device("cuda", dev))
    if partitioned:
        gw.set_partition(rank, world)
        mcap = args.migrate_cap
        slot_bytes = gw.mgpu_slot_bytes(mcap)
        my_slot = torch.zeros(slot_bytes, dtype=torch.uint8, device=f"cuda:{dev}")
        all_slots = torch.zeros(slot_bytes * world, dtype=torch.uint8, device=f"cuda:{dev}")

    def one_step():
        """One collision step on the resident transforms.  The partitioned world adds the manifold migration: every rank
        packs the manifolds of the pairs it stopped owning into a fixed-size slot, ONE NCCL all-gather (NVLink) enqueued
        on the ctx stream right behind the export, and every rank adopts what it owns now.  No host synchronisation."""
        if not partitioned:
            gw.step_device()
            return
        gw.mgpu_broadphase()
        gw.mgpu_export_departed_slot(my_slot.data_ptr(), mcap)
        with torch.cuda.stream(stream):
            dist.all_gather_into_tensor(all_slots, my_slot)
        gw.mgpu_import_arrival_slots(all_slots.data_ptr(), world, mcap)
        gw.mgpu_narrowphase()

    frames = [np.ascontiguousarray(pkg.transforms_to_planes(sc.transforms(k))) for k in range(FRAMES)]
    log(f"[rank {rank}] scene built: {nb} proxies in {time.time() - t0:.1f}s")

    dframes = []
    for f in frames:
        t = torch.from_numpy(f).to(f"cuda:{dev}")
        dframes.append(t)
    hframes = [torch.from_numpy(f).pin_memory() for f in frames]
    flush = torch.empty(192 * 1024 * 1024 // 4, dtype=torch.float32, device=f"cuda:{dev}")  # 192 MiB > 126 MB L2
    torch.cuda.synchronize()

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- device-resident arm: `value` ----------------
    # the timed steps run WITHOUT per-stage events (so the library may replay the step as one CUDA graph); the per-stage
    # times the roofline is computed from come from a short profiled pass afterwards
    step_no = 0
    for _ in range(args.warmup):
        gw.setWorldTransformsDevice(nb, dframes[frame_index(step_no)].data_ptr())
        one_step()
        gw.sync_counts()
        step_no += 1
    barrier()
    sampler = ClockSampler(dev)
    if rank == 0:
        sampler.start()
    ev0 = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    ev1 = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    stage_sum = {}
    pairs_tot = contacts_tot = manif_tot = 0
    launches = 0
    st = None
    barrier()
    for k in range(args.steps):
        with torch.cuda.stream(stream):
            flush.fill_(float(k))  # L2 flush between timed iterations (not timed)
            # inputs already resident in HBM: the frame is a device tensor
            ev0[k].record(stream)
        if args.profile_step and k == 0:
            torch.cuda.synchronize()
            torch.cuda.profiler.start()   # ncu --profile-from-start off: capture exactly one timed step
        gw.setWorldTransformsDevice(nb, dframes[frame_index(step_no)].data_ptr())
        one_step()
        with torch.cuda.stream(stream):
            ev1[k].record(stream)
        p, m, c = gw.sync_counts()
        if args.profile_step and k == 0:
            torch.cuda.profiler.stop()
        pairs_tot += p; manif_tot += m; contacts_tot += c
        st = gw.stats()
        launches += st["kernel_launches"]
        step_no += 1
    barrier()
    ms_steps = [ev0[k].elapsed_time(ev1[k]) for k in range(args.steps)]
    ms_per_step = float(np.mean(ms_steps))
    # per-stage pass (CUDA events around each kernel group; not part of `value`)
    gw.set_profiling(True)
    prof_steps = max(3, min(args.steps, 10))
    for k in range(prof_steps):
        with torch.cuda.stream(stream):
            flush.fill_(float(k))
        gw.setWorldTransformsDevice(nb, dframes[frame_index(step_no)].data_ptr())
        one_step()
        gw.sync_counts()
        for nm, ms in gw.stage_times().items():
            stage_sum[nm] = stage_sum.get(nm, 0.0) + ms
        step_no += 1
    gw.set_profiling(False)
    barrier()

    # ---------------- end-to-end arm through the C ABI with HOST buffers: `e2e` ----------------
    P_cap = args.max_pairs
    pairs_host = torch.empty((P_cap, 2), dtype=torch.int32).pin_memory()
    hdr_host = torch.empty((P_cap, 4), dtype=torch.int32).pin_memory()
    pts_host = torch.empty((2 * P_cap, 12), dtype=torch.int32).pin_memory()
    gw.set_contact_prefetch(2)   # the packed contact stream is compacted at the end of the step's graph
    e2e_steps = max(3, min(args.steps, 20))
    nP, nH, nPt = ctypes.c_int32(), ctypes.c_int32(), ctypes.c_int32()
    e2e_ms = []
    d2h = 0
    barrier()
    for k in range(e2e_steps + 2):
        with torch.cuda.stream(stream):
            flush.fill_(float(k))
        torch.cuda.synchronize()
        t_start = time.perf_counter()
        gw.setWorldTransformsHostPtr(nb, hframes[frame_index(step_no)].data_ptr())       # H2D of this step's inputs
        one_step()                                                                        # enqueue only, no host sync
        gw._ck(L.b2c_get_pairs(gw.h, ctypes.c_void_p(pairs_host.data_ptr()), P_cap, ctypes.byref(nP)))   # D2H pair list (while the narrowphase runs)
        gw._ck(L.b2c_begin_contact_download(gw.h, ctypes.c_void_p(hdr_host.data_ptr()), P_cap, ctypes.c_void_p(pts_host.data_ptr()),
                                            2 * P_cap))                                   # D2H of the manifolds that are final before the EPA tail
        gw.sync_counts()
        gw._ck(L.b2c_get_packed_contacts(gw.h, ctypes.c_void_p(hdr_host.data_ptr()), P_cap, ctypes.c_void_p(pts_host.data_ptr()),
                                         2 * P_cap, ctypes.byref(nH), ctypes.byref(nPt)))  # D2H contact stream (16-B headers, 48-B points)
        t_end = time.perf_counter()
        step_no += 1
        if k >= 2:
            e2e_ms.append((t_end - t_start) * 1e3)
            d2h = nP.value * 8 + nH.value * 16 + nPt.value * 48 + 128
    barrier()
    e2e_ms_per_step = float(np.mean(e2e_ms))
    clocks = sampler.stop() if rank == 0 else None

    # ---------------- reduce over ranks ----------------
    vals = torch.tensor([ms_per_step, e2e_ms_per_step], dtype=torch.float64, device=f"cuda:{dev}")
    sums = torch.tensor([pairs_tot, contacts_tot, manif_tot, launches], dtype=torch.float64, device=f"cuda:{dev}")
    if dist is not None:
        dist.all_reduce(vals, op=dist.ReduceOp.MAX)
        dist.all_reduce(sums, op=dist.ReduceOp.SUM)
    ms_max, e2e_max = [float(x) for x in vals.tolist()]
    pairs_all, contacts_all, manif_all, launches_all = [float(x) for x in sums.tolist()]

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak_gbs = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    stage_ms = {k: v / prof_steps for k, v in stage_sum.items()}
    dom = max(stage_ms, key=stage_ms.get)
    P_avg = pairs_tot / args.steps
    contacts_live = st["num_manifolds"] and contacts_tot / args.steps
    body_bits = 12 + min(20, int(np.ceil(np.log2(2 * nb + 66))))
    pair_bits = 2 * int(np.ceil(np.log2(nb + 2)))
    abytes = {s: algorithmic_bytes(s, nb, P_avg, st, (body_bits + 7) // 8, (pair_bits + 7) // 8, contacts_live) for s in stage_ms}
    achieved = abytes[dom] / (stage_ms[dom] * 1e-3) / 1e9 if stage_ms[dom] > 0 else 0.0
    traffic = None
    try:
        tr = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        traffic = tr.get(dom)
    except Exception:
        pass
    strong = wl in ("c4", "c5")
    wl_name = {"c2": f"C2: {N} mixed boxes/spheres/16-pt hulls in a closed bin of 5 static boxes, single world per GPU, "
                     "DbvtBroadphase pair semantics, seeded transform trace",
               "c3": f"C3: {N} convex hulls (16 pts) and spheres on a {2 * max(8, int(round(708 * (N / 10000.0) ** 0.5))) ** 2}-triangle "
                     "BvhTriangleMeshShape heightfield (quantized BVH), one world per GPU",
               "c4": f"C4: {args.worlds} independent 64-body dice worlds, split by world over the GPUs",
               "c5": f"C5: {N} spheres r=0.5 at 40% packing, ONE world partitioned by sorted-AABB slices over the GPUs, "
                     "departed manifolds all-gathered with NCCL"}[wl]
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
        "config": {"workload": wl_name,
                   "snapshot": f"settled ({args.settle} relaxation iterations)" if args.settle > 0 else "raw jittered lattice (deep overlaps)",
                   "deep_penetration_checks_per_step": st["deep_penetration_checks"],
                   "proxies": nb, "l2": "flushed between timed iterations (192 MiB write)",
                   "parallelism": ("1 GPU" if ngpu == 1 else
                                   {"c2": "1 world per GPU, no collective", "c3": "1 world per GPU, no collective", "c4": f"{args.worlds // world} worlds per GPU, no collective",
                                    "c5": f"every GPU sorts all proxies, sweeps and dispatches 1/{world} of the sorted list; one "
                                          f"all-gather of {args.migrate_cap}-manifold migration slots per step"}[wl])},
        "pairs_per_s": pairs_all / args.steps / (ms_max * 1e-3),
        "contacts_per_s": contacts_all / args.steps / (ms_max * 1e-3),
        "pairs_per_step": pairs_all / args.steps / ngpu,
        "contacts_added_per_step": contacts_all / args.steps / ngpu,
        "manifolds_per_step": manif_all / args.steps / ngpu,
        "stage_ms": {k: round(v, 5) for k, v in stage_ms.items()},
        "gpu_launches": int(launches_all),
        "e2e": {"value": (1000.0 / e2e_max) if strong else ngpu * 1000.0 / e2e_max, "unit": "steps/s", "ms_per_step": e2e_max,
                "h2d_bytes_per_step": nb * 48, "d2h_bytes_per_step": int(d2h),
                "what": "b2c_set_transforms(pinned host planes) + b2c_step_device + b2c_get_pairs (overlaps the narrowphase) + b2c_begin_contact_download (overlaps the penetration bin) + b2c_sync_counts + b2c_get_packed_contacts (16-B manifold headers + 48-B solver points: world points on A and B, normal, distance, lifetime, warm-start slot, triangle index), all into pinned host buffers"},
        "roofline": {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak_gbs, "unit": "GB/s",
                     "frac": achieved / peak_gbs, "traffic": traffic, "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": abytes[dom],
                     "all_stages_frac": {s: round((abytes[s] / (stage_ms[s] * 1e-3) / 1e9) / peak_gbs, 5) if stage_ms[s] > 0 else 0.0
                                         for s in stage_ms}},
        "clocks": clocks,
    }
    if ngpu == 1 and not args.no_cpu and wl == "c2":
        out["cpu_baseline"] = cpu_baseline(sc, args)
    emit(out)
    if dist is not None:
        dist.destroy_process_group()


def cpu_steps(sc, n_steps, warmup=1):
    """Time the CPU oracle (a port of the reference's algorithm; single thread like the reference's world)."""
    import scenes
    ow = scenes.build_oracle(sc, 2)  # 2 = literal restatement of the reference's Dbvt tree broadphase (oracle/dbvt_literal.h)
    times, pairs, contacts = [], 0, 0
    for k in range(warmup + n_steps):
        xf = sc.transforms(frame_index(k))
        t, p, m = ow.timed_step(xf)
        if k >= warmup:
            times.append(t)
            pairs = p
    return times, pairs


def cpu_baseline(sc, args):
    times, pairs = cpu_steps(sc, args.cpu_steps)
    ms = float(np.mean(times)) * 1e3
    return {"value": 1000.0 / ms, "unit": "steps/s", "ms_per_step": ms, "cores": 1, "kind": "port",
            "sample": f"{args.cpu_steps} full steps of the same {sc.n}-proxy C2 world after 1 warm-up (oracle/, g++ -O2, literal Dbvt "
                      "tree broadphase + GJK/EPA narrowphase, single thread: the reference steps one world on one thread)",
            "pairs": int(pairs), "host_cpus": os.cpu_count()}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sc = make_scene(args.bodies, seed=100)
    settled = args.settle > 0 and load_settled(sc, args.bodies, 100, args.settle)
    if settled:
        sc.vel *= 0.25
    steps = max(1, min(args.steps, args.cpu_steps if args.cpu_steps > 0 else 4))
    times, pairs = cpu_steps(sc, steps, warmup=min(args.warmup, 1))
    ms = float(np.mean(times)) * 1e3
    v = 1000.0 / ms
    out = {
        "impl": "reference",
        "metric": "collision_phase_world_steps_per_s_100k_bodies", "value": v,
        "unit": "steps/s (one step = full collision phase of a 100k-body world)",
        "n_gpus": args.gpus, "steps": steps, "warmup": min(args.warmup, 1), "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"C2: {args.bodies} mixed boxes/spheres/16-pt hulls in a closed bin of 5 static boxes, single world, "
                               "DbvtBroadphase pair semantics, seeded transform trace",
                   "snapshot": f"settled ({args.settle} relaxation iterations)" if settled else "raw jittered lattice (deep overlaps)",
                   "proxies": sc.n},
        "cpu_baseline": {"value": v, "unit": "steps/s", "cores": 1, "kind": "port",
                         "sample": f"{steps} full steps of the {sc.n}-proxy C2 world; Java reference not runnable on this box (no JVM): "
                                   "CPU baseline is the C++ restatement in oracle/"},
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
    ap.add_argument("--bodies", type=int, default=100000)
    ap.add_argument("--max-pairs", type=int, default=3 << 20)
    ap.add_argument("--cpu-steps", type=int, default=4)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--workload", default="c2", choices=["c2", "c3", "c4", "c5"],
                    help="c2 = BASELINE headline (default); c3 = convex bodies on a 1M-triangle BVH mesh (use --bodies 10000); "
                         "c4 = batched worlds split by world; c5 = one partitioned world")
    ap.add_argument("--worlds", type=int, default=4096)
    ap.add_argument("--migrate-cap", type=int, default=8192, help="c5, N>1: manifolds per migration slot")
    ap.add_argument("--profile-step", action="store_true",
                    help="bracket the first timed step with cudaProfilerStart/Stop (for ncu --profile-from-start off)")
    ap.add_argument("--save-settled", default="", help="write the settled origins (npz) here; commit it under tests/golden/")
    ap.add_argument("--settle", type=int, default=60, help="relaxation iterations for the settled snapshot (0 = raw lattice)")
    args = ap.parse_args()
    claim_stdout()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
