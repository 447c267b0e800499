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
    gw = scenes.build_gpu(pkg, sc, mode=pkg.DBVT, max_pairs=P_cap)
    L = gw.L
    nb = sc.n
    frames = [torch.from_numpy(np.ascontiguousarray(sc.transforms(k).T)).pin_memory() for k in range(4)]
    pairs_host = torch.empty((P_cap, 2), dtype=torch.int32).pin_memory()
    hdr_host = torch.empty((P_cap, 4), dtype=torch.int32).pin_memory()
    pts_host = torch.empty((2 * P_cap, 12), dtype=torch.int32).pin_memory()
    nP, nH, nPt = ctypes.c_int32(), ctypes.c_int32(), ctypes.c_int32()
    gw.set_contact_prefetch(2)
    names = ["set_transforms", "step_device", "get_pairs", "begin_contact_download", "sync_counts", "get_packed_contacts"]
    acc = np.zeros(len(names))
    steps = 0
    for k in range(14):
        torch.cuda.synchronize()
        t = [time.perf_counter()]
        gw.setWorldTransformsHostPtr(nb, frames[k % 4].data_ptr()); t.append(time.perf_counter())
        gw.step_device(); t.append(time.perf_counter())
        gw._ck(L.b2c_get_pairs(gw.h, ctypes.c_void_p(pairs_host.data_ptr()), P_cap, ctypes.byref(nP))); t.append(time.perf_counter())
        gw._ck(L.b2c_begin_contact_download(gw.h, ctypes.c_void_p(hdr_host.data_ptr()), P_cap, ctypes.c_void_p(pts_host.data_ptr()), 2 * P_cap)); t.append(time.perf_counter())
        gw.sync_counts(); t.append(time.perf_counter())
        gw._ck(L.b2c_get_packed_contacts(gw.h, ctypes.c_void_p(hdr_host.data_ptr()), P_cap, ctypes.c_void_p(pts_host.data_ptr()), 2 * P_cap,
                                         ctypes.byref(nH), ctypes.byref(nPt))); t.append(time.perf_counter())
        if k >= 4:
            acc += np.diff(t) * 1e3
            steps += 1
    acc /= steps
    print({n: round(float(v), 4) for n, v in zip(names, acc)}, "total", round(float(acc.sum()), 4), "pairs", nP.value, "hdr", nH.value, "pts", nPt.value)


if __name__ == "__main__":
    main()
