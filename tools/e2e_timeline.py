"""Host-side timeline of the end-to-end step (what bench.py's `e2e` times): ms spent inside each C-ABI call."""
import ctypes
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as ge  # noqa: E402
import bench  # noqa: E402
import scenes  # noqa: E402


def main():
    pkg = ge.load_package()
    sc = bench.make_scene(100000, seed=100)
    bench.load_settled(sc, 100000, 100, 60)
    sc.vel *= 0.25
    P_cap = 3 << 20
    gw = scenes.build_gpu(pkg, sc, mode=pkg.DBVT, max_pairs=P_cap, raw_records=False)
    L = gw.L
    nb = sc.n
    frames = [torch.from_numpy(np.ascontiguousarray(pkg.transforms_to_planes(sc.transforms(k)))).pin_memory() for k in range(4)]
    D_cap = P_cap // 2
    hdr_host = torch.empty((P_cap, 4), dtype=torch.int32).pin_memory()
    pts_host = torch.empty((2 * P_cap, 12), dtype=torch.int32).pin_memory()
    add_host = torch.empty((D_cap, 2), dtype=torch.int32).pin_memory()
    rem_host = torch.empty((D_cap, 2), dtype=torch.int32).pin_memory()
    nA, nR, nH, nPt = ctypes.c_int32(), ctypes.c_int32(), ctypes.c_int32(), ctypes.c_int32()
    vp = ctypes.c_void_p
    gw.set_contact_prefetch(3)
    gw.set_pair_delta_prefetch(True)
    # the same call sequence bench.py's e2e arm times
    names = ["set_transforms (H2D 4.8 MB)", "step_device (enqueue)", "get_pair_deltas", "begin_contact_download", "sync_counts",
             "get_packed_contacts_uid"]
    acc = np.zeros(len(names))
    steps = 0
    for k in range(24):
        torch.cuda.synchronize()
        t = [time.perf_counter()]
        gw.setWorldTransformsHostPtr(nb, frames[k % 4].data_ptr()); t.append(time.perf_counter())
        gw.step_device(); t.append(time.perf_counter())
        gw._ck(L.b2c_get_pair_deltas(gw.h, vp(add_host.data_ptr()), D_cap, vp(rem_host.data_ptr()), D_cap, ctypes.byref(nA), ctypes.byref(nR))); t.append(time.perf_counter())
        gw._ck(L.b2c_begin_contact_download(gw.h, vp(hdr_host.data_ptr()), P_cap, vp(pts_host.data_ptr()), 2 * P_cap)); t.append(time.perf_counter())
        gw.sync_counts(); t.append(time.perf_counter())
        gw._ck(L.b2c_get_packed_contacts_uid(gw.h, vp(hdr_host.data_ptr()), P_cap, vp(pts_host.data_ptr()), 2 * P_cap, ctypes.byref(nH), ctypes.byref(nPt))); t.append(time.perf_counter())
        if k >= 4:
            acc += np.diff(t) * 1e3
            steps += 1
    acc /= steps
    print({n: round(float(v), 4) for n, v in zip(names, acc)}, "total", round(float(acc.sum()), 4), "deltas", nA.value, nR.value,
          "headers", nH.value, "points", nPt.value, "d2h MB", round(((nA.value + nR.value) * 8 + nH.value * 16 + nPt.value * 48) / 1e6, 2))


if __name__ == "__main__":
    main()
