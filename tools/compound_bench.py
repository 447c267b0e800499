"""Side measurement (not the headline bench): one collision step of a compound-heavy scene on the CUDA path.
Usage: python tools/compound_bench.py [n_bodies] [steps]   -> one JSON line."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as ge  # noqa: E402
import scenes  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 30
    pkg = ge.load_package()
    sc = scenes.compound_scene(n=n, seed=8)
    gw = scenes.build_gpu(pkg, sc, mode=pkg.DBVT, max_pairs=max(64 * n, 1 << 16), max_compound_items=max(256 * n, 1 << 16))
    planes = [np.ascontiguousarray(sc.transforms(s).T) for s in range(4)]
    for s in range(6):   # warm-up: both ping-pong parities captured as graphs
        counts = gw.step(planes[s % 4])
    ms = []
    for s in range(steps):
        gw.setWorldTransformPlanes(planes[s % 4])
        t0 = time.perf_counter()
        gw.step_device()
        counts = gw.sync_counts()
        ms.append((time.perf_counter() - t0) * 1e3)
    st = gw.stats()
    m = gw.manifolds()
    kid = m["child0"] >= 0
    print(json.dumps({"workload": f"compound_scene n={n} (half of the bodies are 1-3-child compounds)", "bodies": sc.n,
                      "pairs": counts[0], "manifolds": counts[1], "child_manifolds": int(kid.sum()),
                      "touching_child_manifolds": int((m["num_contacts"][kid] > 0).sum()),
                      "device_ms_per_step": round(float(st["ms_total"]), 4), "host_ms_per_step_median": round(float(np.median(ms)), 4),
                      "kernel_launches": st["kernel_launches"], "deep_penetration_checks": st["deep_penetration_checks"],
                      "epa_failed": st["epa_failed"]}))


if __name__ == "__main__":
    main()
