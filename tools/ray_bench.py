"""Side measurement: b2c_ray_test_closest on the C2 world (100 k bodies).  Usage: python tools/ray_bench.py [rays]"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as ge  # noqa: E402
import bench  # noqa: E402
import scenes  # noqa: E402


def main():
    nrays = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
    pkg = ge.load_package()
    sc = bench.make_scene(100000, seed=100)
    bench.load_settled(sc, 100000, 100, 60)
    gw = scenes.build_gpu(pkg, sc, mode=pkg.DBVT, max_pairs=3 << 20)
    rng = np.random.default_rng(1)
    ext = float(sc.extent)
    f = rng.uniform(-0.1 * ext, 1.1 * ext, size=(nrays, 3)).astype(np.float32)
    t = rng.uniform(-0.1 * ext, 1.1 * ext, size=(nrays, 3)).astype(np.float32)
    f[:, 1] = ext * 1.2
    out = {"workload": f"{nrays} rays against the settled C2 world ({sc.n} bodies)"}
    for label in ("before_first_broadphase", "after_broadphase"):
        if label == "after_broadphase":
            gw.setWorldTransforms(sc.transforms(0))
            gw.step()
        gw.rayTestClosest(f[:64], t[:64])
        ms = []
        for _ in range(3):
            t0 = time.perf_counter()
            uid, frac, nrm, pt = gw.rayTestClosest(f, t)
            ms.append((time.perf_counter() - t0) * 1e3)
        out[label] = {"ms": round(min(ms), 3), "rays_per_s": round(nrays / (min(ms) * 1e-3)), "hits": int((uid > 0).sum())}
    # CCD-like convex sweeps: every sweep moves a small sphere / box 1.5 units (a fast body's step) from a random point of the pile
    nsw = nrays
    cast = [gw.SphereShape(0.25), gw.BoxShape((0.3, 0.2, 0.25))]
    ids = np.asarray([cast[k % 2] for k in range(nsw)], np.int32)
    sf = rng.uniform(0.0, ext, size=(nsw, 3)).astype(np.float32)
    d = rng.normal(size=(nsw, 3)); d /= np.linalg.norm(d, axis=1, keepdims=True)
    st = (sf + 1.5 * d).astype(np.float32)
    basis = scenes.random_rotations(rng, nsw).astype(np.float32)
    gw.convexSweepTestClosest(ids[:64], basis[:64], sf[:64], st[:64], 1, 1)
    ms = []
    for _ in range(3):
        t0 = time.perf_counter()
        uid, frac, nrm, pt = gw.convexSweepTestClosest(ids, basis, sf, st, 1, 1)       # dynamic bodies only
        ms.append((time.perf_counter() - t0) * 1e3)
    out["convex_sweeps"] = {"what": f"{nsw} translational sweeps of 1.5 units (sphere r=0.25 / box 0.3x0.2x0.25, random bases) inside the pile, callback mask 1",
                            "ms": round(min(ms), 3), "sweeps_per_s": round(nsw / (min(ms) * 1e-3)), "hits": int((uid > 0).sum())}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
