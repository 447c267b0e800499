"""Scratch script for GPU debugging sessions (not a test)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import scenes
import __graft_entry__ as ge
pkg = ge.load_package()
sc = scenes.spheres_scene(n=20000, seed=8)
sc.vel *= 6.0
R = 4
ranks = [scenes.build_gpu(pkg, sc, mode=1) for _ in range(R)]
for r, w in enumerate(ranks):
    w.set_partition(r, R)
scap = 1 << 14
sbytes = ranks[0].mgpu_slot_bytes(scap)
print("slot bytes", sbytes, 16 + 424 * scap)
allslots = torch.zeros(sbytes * R, dtype=torch.uint8, device="cuda")
cap = 1 << 16
bufs = [(torch.zeros(cap, dtype=torch.int64, device="cuda"), torch.zeros(cap * 8, dtype=torch.int32, device="cuda"),
         torch.zeros(cap * 4 * 24, dtype=torch.int32, device="cuda")) for _ in range(R)]
for step in range(3):
    xf = sc.transforms(step)
    for w in ranks:
        w.setWorldTransforms(xf)
        w.mgpu_broadphase()
    for r, w in enumerate(ranks):
        k, h, p = bufs[r]
        c = w.mgpu_export_departed(k.data_ptr(), h.data_ptr(), p.data_ptr(), cap)
        w.mgpu_export_departed_slot(allslots.data_ptr() + r * sbytes, scap)
        torch.cuda.synchronize()
        print(step, r, "old-api count", c, "slot count", allslots[r * sbytes:r * sbytes + 16].view(torch.int32).tolist())
    for w in ranks:
        w.mgpu_import_arrival_slots(allslots.data_ptr(), R, scap)
        w.mgpu_narrowphase()
        try:
            print(w.sync_counts())
        except Exception as e:
            print("ERR", e)
