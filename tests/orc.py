"""ctypes binding of the CPU oracle (oracle/liboracle.so).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_ORACLE_DIR = os.path.join(_ROOT, "oracle")
_LIB = None

f32p = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
u8p = np.ctypeslib.ndpointer(dtype=np.uint8, flags="C_CONTIGUOUS")


def build():
    subprocess.check_call(["make", "-s", "-C", _ORACLE_DIR])


def lib():
    global _LIB
    if _LIB is not None:
        return _LIB
    path = os.path.join(_ORACLE_DIR, "liboracle.so")
    srcs = [os.path.join(_ORACLE_DIR, f) for f in os.listdir(_ORACLE_DIR) if f.endswith((".h", ".cpp"))]
    if not os.path.exists(path) or any(os.path.getmtime(s) > os.path.getmtime(path) for s in srcs):
        build()
    L = C.CDLL(path)
    L.orc_create.restype = C.c_void_p
    L.orc_create.argtypes = [C.c_int]
    L.orc_destroy.argtypes = [C.c_void_p]
    L.orc_set_brute_force.argtypes = [C.c_void_p, C.c_int]
    L.orc_set_no_collide_pairs.argtypes = [C.c_void_p, C.c_int, i32p]
    L.orc_set_world_aabb.argtypes = [C.c_void_p, f32p, f32p]
    L.orc_sap_quantize.argtypes = [C.c_void_p, f32p, C.c_int, np.ctypeslib.ndpointer(dtype=np.uint32, flags="C_CONTIGUOUS")]
    L.orc_set_params.argtypes = [C.c_void_p, C.c_float, C.c_float, C.c_float]
    L.orc_shape_box.argtypes = [C.c_void_p, C.c_float, C.c_float, C.c_float]
    L.orc_shape_sphere.argtypes = [C.c_void_p, C.c_float]
    L.orc_shape_hull.argtypes = [C.c_void_p, f32p, C.c_int]
    L.orc_shape_plane.argtypes = [C.c_void_p, C.c_float, C.c_float, C.c_float, C.c_float]
    L.orc_shape_mesh.argtypes = [C.c_void_p, f32p, C.c_int, i32p, C.c_int]
    L.orc_shape_compound.argtypes = [C.c_void_p, C.c_int, i32p, f32p]
    L.orc_shape_mesh_parts.argtypes = [C.c_void_p, C.c_int, f32p, i32p, i32p, i32p]
    L.orc_mesh_num_nodes.argtypes = [C.c_void_p, C.c_int]
    L.orc_mesh_get_nodes.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
    L.orc_mesh_get_quant.argtypes = [C.c_void_p, C.c_int, f32p]
    L.orc_body_create.argtypes = [C.c_void_p, C.c_int, f32p, C.c_int, C.c_int, C.c_int, C.c_int]
    L.orc_body_destroy.argtypes = [C.c_void_p, C.c_int]
    L.orc_num_bodies.argtypes = [C.c_void_p]
    L.orc_set_transforms.argtypes = [C.c_void_p, C.c_int, C.c_void_p, f32p]
    L.orc_set_active.argtypes = [C.c_void_p, C.c_int, C.c_void_p, u8p]
    L.orc_set_material.argtypes = [C.c_void_p, C.c_int, C.c_float, C.c_float]
    L.orc_update_aabbs.argtypes = [C.c_void_p]
    L.orc_set_aabb.argtypes = [C.c_void_p, C.c_int, f32p, f32p]
    L.orc_get_aabbs.argtypes = [C.c_void_p, f32p]
    L.orc_calculate_overlapping_pairs.argtypes = [C.c_void_p]
    L.orc_get_pairs.argtypes = [C.c_void_p, i32p]
    L.orc_dispatch_all_pairs.argtypes = [C.c_void_p]
    L.orc_ray_test_closest.argtypes = [C.c_void_p, C.c_int, f32p, f32p, C.c_int, C.c_int, i32p, f32p]
    L.orc_convex_sweep_closest.argtypes = [C.c_void_p, C.c_int, i32p, f32p, f32p, f32p, C.c_int, C.c_int, C.c_float, i32p, f32p]
    L.orc_ccd_sweep_not_me.argtypes = [C.c_void_p, C.c_int, i32p, f32p, f32p, C.c_float, i32p, f32p]
    L.orc_pair_deltas.argtypes = [C.c_void_p, i32p, C.c_int, i32p, C.c_int, i32p]
    L.orc_islands.argtypes = [C.c_void_p, i32p]
    L.orc_num_raw.argtypes = [C.c_void_p]
    L.orc_get_raw.argtypes = [C.c_void_p, i32p, f32p]
    L.orc_get_manifolds.argtypes = [C.c_void_p, C.c_int, i32p, f32p, i32p]
    L.orc_get_counters.argtypes = [C.c_void_p, C.c_void_p]
    L.orc_shape_aabb.argtypes = [C.c_void_p, C.c_int, f32p, f32p]
    L.orc_gjk_pair.argtypes = [C.c_void_p, C.c_int, f32p, C.c_int, f32p, i32p, f32p]
    L.orc_support.argtypes = [C.c_void_p, C.c_int, f32p, C.c_int, f32p]
    L.orc_bvh_query.argtypes = [C.c_void_p, C.c_int, f32p, f32p, i32p, C.c_int]
    L.orc_epa_constants.argtypes = [f32p]
    L.orc_timed_step.restype = C.c_double
    L.orc_timed_step.argtypes = [C.c_void_p, C.c_int, f32p, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    _LIB = L
    return L


def xf12(basis=None, origin=(0, 0, 0)):
    """9 row-major basis floats + 3 origin floats."""
    b = np.eye(3, dtype=np.float32) if basis is None else np.asarray(basis, dtype=np.float32).reshape(3, 3)
    return np.concatenate([b.reshape(9), np.asarray(origin, dtype=np.float32)]).astype(np.float32)


TIGHT, DBVT, DBVT_LITERAL = 0, 1, 2
SAP16, SAP32, SAP16_LITERAL, SAP32_LITERAL = 3, 4, 5, 6


class OracleWorld:
    def __init__(self, mode=TIGHT, brute_force=False, world_aabb=None):
        self.L = lib()
        self.h = self.L.orc_create(mode)
        if brute_force:
            self.L.orc_set_brute_force(self.h, 1)
        if world_aabb is not None:
            self.L.orc_set_world_aabb(self.h, np.asarray(world_aabb[0], np.float32), np.asarray(world_aabb[1], np.float32))

    def set_no_collide_pairs(self, pairs):
        p = np.ascontiguousarray(pairs, dtype=np.int32).reshape(-1, 2)
        self.L.orc_set_no_collide_pairs(self.h, len(p), p if len(p) else np.zeros((1, 2), np.int32))

    def sap_quantize(self, p, is_max):
        out = np.zeros(3, dtype=np.uint32)
        self.L.orc_sap_quantize(self.h, np.asarray(p, np.float32), int(is_max), out)
        return out

    def __del__(self):
        try:
            self.L.orc_destroy(self.h)
        except Exception:
            pass

    # shapes
    def box(self, hx, hy, hz):
        return self.L.orc_shape_box(self.h, hx, hy, hz)

    def sphere(self, r):
        return self.L.orc_shape_sphere(self.h, r)

    def hull(self, pts):
        pts = np.ascontiguousarray(pts, dtype=np.float32).reshape(-1, 3)
        return self.L.orc_shape_hull(self.h, pts, len(pts))

    def plane(self, n, c):
        return self.L.orc_shape_plane(self.h, n[0], n[1], n[2], c)

    def compound(self, child_shapes, child_xf):
        """CompoundShape with addChildShape(child_xf[i], child_shapes[i]) in order (sh/CompoundShape.java:50-82)."""
        cs = np.ascontiguousarray(child_shapes, dtype=np.int32)
        xf = np.ascontiguousarray(child_xf, dtype=np.float32).reshape(-1, 12)
        assert len(cs) == len(xf)
        sid = self.L.orc_shape_compound(self.h, len(cs), cs, xf)
        assert sid >= 0, "compound children must be convex shapes registered before"
        return sid

    def mesh(self, verts, idx):
        verts = np.ascontiguousarray(verts, dtype=np.float32).reshape(-1, 3)
        idx = np.ascontiguousarray(idx, dtype=np.int32).reshape(-1, 3)
        return self.L.orc_shape_mesh(self.h, verts, len(verts), idx, len(idx))

    def mesh_parts(self, parts):
        """BvhTriangleMeshShape over a TriangleIndexVertexArray with several IndexedMesh parts: [(verts, idx), ...]
        (indices local to their part, any integer dtype)."""
        v = np.ascontiguousarray(np.concatenate([np.asarray(p[0], np.float32).reshape(-1, 3) for p in parts]))
        i = np.ascontiguousarray(np.concatenate([np.asarray(p[1]).astype(np.int32).reshape(-1, 3) for p in parts]))
        nv = np.asarray([len(np.asarray(p[0]).reshape(-1, 3)) for p in parts], dtype=np.int32)
        nt = np.asarray([len(np.asarray(p[1]).reshape(-1, 3)) for p in parts], dtype=np.int32)
        return self.L.orc_shape_mesh_parts(self.h, len(parts), v, nv, i, nt)

    def mesh_nodes(self, shape):
        n = self.L.orc_mesh_num_nodes(self.h, shape)
        out = np.zeros((n, 4), dtype=np.int32)
        self.L.orc_mesh_get_nodes(self.h, shape, out.ctypes.data)
        q = np.zeros(9, dtype=np.float32)
        self.L.orc_mesh_get_quant(self.h, shape, q)
        return out, q

    # bodies
    def body(self, shape, xf, group=1, mask=-1, static=False, world=0):
        return self.L.orc_body_create(self.h, shape, np.ascontiguousarray(xf, dtype=np.float32), int(group), int(mask),
                                      int(static), int(world))

    def destroy_body(self, uid):
        self.L.orc_body_destroy(self.h, uid)

    @property
    def num_bodies(self):
        return self.L.orc_num_bodies(self.h)

    def set_transforms(self, xf, uids=None):
        xf = np.ascontiguousarray(xf, dtype=np.float32).reshape(-1, 12)
        if uids is None:
            self.L.orc_set_transforms(self.h, len(xf), None, xf)
        else:
            u = np.ascontiguousarray(uids, dtype=np.int32)
            self.L.orc_set_transforms(self.h, len(xf), u.ctypes.data, xf)

    def set_active(self, active, uids=None):
        a = np.ascontiguousarray(active, dtype=np.uint8)
        if uids is None:
            self.L.orc_set_active(self.h, len(a), None, a)
        else:
            u = np.ascontiguousarray(uids, dtype=np.int32)
            self.L.orc_set_active(self.h, len(a), u.ctypes.data, a)

    def update_aabbs(self):
        self.L.orc_update_aabbs(self.h)

    def set_aabb(self, uid, mn, mx):
        self.L.orc_set_aabb(self.h, uid, np.asarray(mn, dtype=np.float32), np.asarray(mx, dtype=np.float32))

    def aabbs(self):
        out = np.zeros((self.num_bodies, 6), dtype=np.float32)
        self.L.orc_get_aabbs(self.h, out)
        return out

    def calculate_overlapping_pairs(self):
        n = self.L.orc_calculate_overlapping_pairs(self.h)
        out = np.zeros((n, 2), dtype=np.int32)
        if n:
            self.L.orc_get_pairs(self.h, out)
        return out

    def pair_deltas(self, cap=1 << 22):
        """(added, removed) pair lists of the last calculate_overlapping_pairs, each sorted (n,2)."""
        a = np.zeros((cap, 2), dtype=np.int32)
        r = np.zeros((cap, 2), dtype=np.int32)
        n2 = np.zeros(2, dtype=np.int32)
        self.L.orc_pair_deltas(self.h, a, cap, r, cap, n2)
        return a[: n2[0]].copy(), r[: n2[1]].copy()

    def islands(self):
        """Island tag per object (index = uid-1), -1 for statics; and the number of islands."""
        t = np.zeros(max(self.num_bodies, 1), dtype=np.int32)
        n = self.L.orc_islands(self.h, t)
        return t[: self.num_bodies], n

    def ray_test_closest(self, ray_from, ray_to, group=1, mask=-1):
        """CollisionWorld.rayTest with a ClosestRayResultCallback per ray: (uid (0 = miss), fraction, normal, point)."""
        f = np.ascontiguousarray(ray_from, dtype=np.float32).reshape(-1, 3)
        t = np.ascontiguousarray(ray_to, dtype=np.float32).reshape(-1, 3)
        uid = np.zeros(len(f), dtype=np.int32)
        out = np.zeros((len(f), 7), dtype=np.float32)
        self.L.orc_ray_test_closest(self.h, len(f), f, t, int(group), int(mask), uid, out)
        return uid, out[:, 0].copy(), out[:, 1:4].copy(), out[:, 4:7].copy()

    def convex_sweep_closest(self, cast_shapes, basis, sweep_from, sweep_to, group=1, mask=-1, allowed_ccd_penetration=0.04):
        """CollisionWorld.convexSweepTest with a ClosestConvexResultCallback per translational sweep: (uid (0 = miss, -1 = the
        static-plane branch the reference throws in), fraction, normal, point)."""
        f = np.ascontiguousarray(sweep_from, dtype=np.float32).reshape(-1, 3)
        t = np.ascontiguousarray(sweep_to, dtype=np.float32).reshape(-1, 3)
        n = len(f)
        ids = np.full(n, int(cast_shapes), np.int32) if np.isscalar(cast_shapes) else np.ascontiguousarray(cast_shapes, dtype=np.int32)
        b = np.ascontiguousarray(basis, dtype=np.float32).reshape(-1, 9)
        if len(b) == 1 and n != 1:
            b = np.ascontiguousarray(np.repeat(b, n, axis=0))
        uid = np.zeros(n, dtype=np.int32)
        out = np.zeros((n, 7), dtype=np.float32)
        self.L.orc_convex_sweep_closest(self.h, n, ids, b, f, t, int(group), int(mask), float(allowed_ccd_penetration), uid, out)
        return uid, out[:, 0].copy(), out[:, 1:4].copy(), out[:, 4:7].copy()

    def ccd_sweep_not_me(self, body_uids, ccd_radius, predicted_origins, allowed_ccd_penetration=0.04):
        """DiscreteDynamicsWorld.integrateTransforms' CCD motion-clamping sweep per listed body (ClosestNotMeConvexResultCallback)."""
        u = np.ascontiguousarray(body_uids, dtype=np.int32).reshape(-1)
        n = len(u)
        r = np.full(n, float(ccd_radius), np.float32) if np.isscalar(ccd_radius) else np.ascontiguousarray(ccd_radius, dtype=np.float32)
        t = np.ascontiguousarray(predicted_origins, dtype=np.float32).reshape(-1, 3)
        uid = np.zeros(n, dtype=np.int32)
        out = np.zeros((n, 7), dtype=np.float32)
        self.L.orc_ccd_sweep_not_me(self.h, n, u, r, t, float(allowed_ccd_penetration), uid, out)
        return uid, out[:, 0].copy(), out[:, 1:4].copy(), out[:, 4:7].copy()

    def dispatch_all_pairs(self):
        return self.L.orc_dispatch_all_pairs(self.h)

    def raw(self):
        n = self.L.orc_num_raw(self.h)
        ints = np.zeros((n, 6), dtype=np.int32)
        fl = np.zeros((n, 7), dtype=np.float32)
        if n:
            self.L.orc_get_raw(self.h, ints, fl)
        return ints, fl

    def manifolds(self):
        hdr0 = np.zeros((1, 7), dtype=np.int32)
        n = self.L.orc_get_manifolds(self.h, 0, hdr0, np.zeros(1, np.float32), np.zeros(1, np.int32))
        hdr = np.zeros((max(n, 1), 7), dtype=np.int32)
        pts = np.zeros((max(n, 1), 4, 18), dtype=np.float32)
        pint = np.zeros((max(n, 1), 4, 6), dtype=np.int32)
        self.L.orc_get_manifolds(self.h, n, hdr, pts, pint)
        return hdr[:n], pts[:n], pint[:n]

    def counters(self):
        out = (C.c_long * 5)()
        self.L.orc_get_counters(self.h, out)
        return dict(zip(["gjk_checks", "deep_penetration_checks", "added_contacts", "bvh_nodes", "triangles"], list(out)))

    def step(self, xf=None):
        if xf is not None:
            self.set_transforms(xf)
        self.update_aabbs()
        pairs = self.calculate_overlapping_pairs()
        self.dispatch_all_pairs()
        return pairs

    def timed_step(self, xf):
        xf = np.ascontiguousarray(xf, dtype=np.float32).reshape(-1, 12)
        a, b = C.c_int(), C.c_int()
        t = self.L.orc_timed_step(self.h, len(xf), xf, C.byref(a), C.byref(b))
        return t, a.value, b.value

    # stand-alone
    def shape_aabb(self, shape, xf):
        out = np.zeros(6, dtype=np.float32)
        self.L.orc_shape_aabb(self.h, shape, np.ascontiguousarray(xf, dtype=np.float32), out)
        return out

    def gjk_pair(self, sa, xa, sb, xb):
        oi = np.zeros(4, dtype=np.int32)
        of = np.zeros(7, dtype=np.float32)
        self.L.orc_gjk_pair(self.h, sa, np.ascontiguousarray(xa, np.float32), sb, np.ascontiguousarray(xb, np.float32), oi, of)
        return dict(has=int(oi[0]), method=int(oi[1]), iters=int(oi[2]), degenerate=int(oi[3]), normal=of[0:3].copy(),
                    point=of[3:6].copy(), depth=float(of[6]))

    def support(self, shape, d, with_margin=False):
        out = np.zeros(3, dtype=np.float32)
        self.L.orc_support(self.h, shape, np.asarray(d, dtype=np.float32), int(with_margin), out)
        return out

    def bvh_query(self, shape, mn, mx, cap=1 << 16):
        out = np.zeros(cap, dtype=np.int32)
        n = self.L.orc_bvh_query(self.h, shape, np.asarray(mn, np.float32), np.asarray(mx, np.float32), out, cap)
        return out[: min(n, cap)]
