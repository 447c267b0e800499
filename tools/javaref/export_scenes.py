"""Export the golden cases of tests/golden/make_golden.py as neutral scene files the Java harness reads
(tools/javaref/DumpGolden.java): tests/golden/scenes/<case>.npz.

    python tools/javaref/export_scenes.py

Arrays (all int32 / float32, C order): mode[1], steps[1], num_worlds[1];
  shape_kind[S] (0 box, 1 sphere, 2 hull, 4 plane, 5 mesh, 6 compound), shape_params[S,4] (box half extents | sphere radius |
  plane normal + constant), hull_off[S+1] + hull_pts[H,3], mesh_off[S+1] (vertices) + mesh_verts[V,3], mesh_toff[S+1]
  (triangles) + mesh_tris[T,3], comp_off[S+1] + comp_child[C] + comp_xf[C,12];
  body_shape[N], body_static[N], body_group[N], body_mask[N], body_world[N]; xf<step>[N,12] (row-major basis + origin).
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import make_golden  # noqa: E402

KIND = {"box": 0, "sphere": 1, "hull": 2, "plane": 4, "mesh": 5, "compound": 6}


def export(name, make, mode, steps, outdir):
    sc = make()
    S = len(sc.shapes)
    kind = np.zeros(S, np.int32)
    params = np.zeros((S, 4), np.float32)
    hull_off = np.zeros(S + 1, np.int32); hull = []
    mesh_off = np.zeros(S + 1, np.int32); mesh_toff = np.zeros(S + 1, np.int32); mv = []; mt = []
    comp_off = np.zeros(S + 1, np.int32); cc = []; cx = []
    for i, s in enumerate(sc.shapes):
        kind[i] = KIND[s[0]]
        if s[0] == "box":
            params[i, :3] = np.asarray(s[1], np.float32)
        elif s[0] == "sphere":
            params[i, 0] = np.float32(s[1])
        elif s[0] == "hull":
            hull.append(np.asarray(s[1], np.float32).reshape(-1, 3))
        elif s[0] == "plane":
            params[i, :3] = np.asarray(s[1], np.float32); params[i, 3] = np.float32(s[2])
        elif s[0] == "mesh":
            mv.append(np.asarray(s[1], np.float32).reshape(-1, 3)); mt.append(np.asarray(s[2], np.int32).reshape(-1, 3))
        elif s[0] == "compound":
            cc.append(np.asarray(s[1], np.int32)); cx.append(np.asarray(s[2], np.float32).reshape(-1, 12))
        hull_off[i + 1] = sum(len(h) for h in hull)
        mesh_off[i + 1] = sum(len(v) for v in mv); mesh_toff[i + 1] = sum(len(t) for t in mt)
        comp_off[i + 1] = sum(len(c) for c in cc)
    cat = lambda xs, shape, dt: np.concatenate(xs).astype(dt) if xs else np.zeros(shape, dt)
    out = dict(mode=np.asarray([mode], np.int32), steps=np.asarray([steps], np.int32), num_worlds=np.asarray([sc.num_worlds], np.int32),
               shape_kind=kind, shape_params=params, hull_off=hull_off, hull_pts=cat(hull, (0, 3), np.float32),
               mesh_off=mesh_off, mesh_verts=cat(mv, (0, 3), np.float32), mesh_toff=mesh_toff, mesh_tris=cat(mt, (0, 3), np.int32),
               comp_off=comp_off, comp_child=cat(cc, (0,), np.int32), comp_xf=cat(cx, (0, 12), np.float32),
               body_shape=np.asarray(sc.body_shape, np.int32), body_static=np.asarray(sc.static, np.int32),
               body_group=np.asarray(sc.group, np.int32), body_mask=np.asarray(sc.mask, np.int32), body_world=np.asarray(sc.world, np.int32))
    for k in range(steps):
        out[f"xf{k}"] = np.ascontiguousarray(sc.transforms(k), np.float32)
    np.savez(os.path.join(outdir, name + ".npz"), **out)   # uncompressed: the Java reader handles STORED entries only


if __name__ == "__main__":
    outdir = os.path.join(ROOT, "tests", "golden", "scenes")
    os.makedirs(outdir, exist_ok=True)
    for name, (make, mode, steps) in make_golden.CASES.items():
        export(name, make, mode, steps, outdir)
        print("wrote", name)
