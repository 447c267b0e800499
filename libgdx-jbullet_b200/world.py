"""Host-side mirror of the reference's collision interfaces over the C ABI (include/b2c.h).

The reference is Java and no JVM exists in this image, so this module is the test/bench host: it keeps
the reference's names, argument meaning and call order for the path
``CollisionWorld.performDiscreteCollisionDetection`` (disp/CollisionWorld.java:123-151):

* ``GpuCollisionWorld``  ~ disp/CollisionWorld.java (addCollisionObject, updateAabbs,
  performDiscreteCollisionDetection, getBroadphase, getDispatcher)
* ``GpuBroadphase``      ~ bp/BroadphaseInterface.java:33-52 (createProxy, destroyProxy, setAabb,
  calculateOverlappingPairs, getOverlappingPairCache)
* ``GpuPairCache``       ~ bp/OverlappingPairCache.java:34-54 (getOverlappingPairArray, getNumOverlappingPairs)
* ``GpuDispatcher``      ~ bp/Dispatcher.java:38-68 (dispatchAllCollisionPairs, getNumManifolds,
  getManifoldByIndexInternal)

All compute happens in libb2c.so; nothing here touches the oracle and nothing falls back to the CPU.
The Java shim a maintainer would add (Panama FFM) is in java/ and INTEGRATION.md.
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import B2CError, Config, MANIFOLD_DTYPE, RAW_DTYPE, Stats

TIGHT, DBVT, SAP16, SAP32 = 0, 1, 2, 3
DEFAULT_FILTER, STATIC_FILTER, ALL_FILTER = 1, 2, -1  # bp/CollisionFilterGroups.java:33-39


def _vp(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def transforms_to_planes(xf):
    """(n,12) row-major basis + origin  ->  12 SoA planes of n floats (the ABI's layout)."""
    xf = np.ascontiguousarray(xf, dtype=np.float32).reshape(-1, 12)
    return np.ascontiguousarray(xf.T)


class GpuCollisionWorld:
    def __init__(self, mode=DBVT, max_bodies=131072, max_pairs=2 << 20, num_worlds=1, device=0, max_mesh_items=1 << 20,
                 max_hull_points=1 << 20, max_shapes=4096, contact_breaking_threshold=0.02, world_aabb=None,
                 max_compound_items=0, raw_records=False):
        self.L = _lib.load()
        cfg = Config()
        self.L.b2c_default_config(C.byref(cfg))
        cfg.device = device
        cfg.broadphase_mode = mode
        cfg.max_bodies = max_bodies
        cfg.max_pairs = max_pairs
        cfg.num_worlds = num_worlds
        cfg.max_mesh_items = max_mesh_items
        cfg.max_hull_points = max_hull_points
        cfg.max_shapes = max_shapes
        cfg.contact_breaking_threshold = contact_breaking_threshold
        cfg.max_compound_items = max_compound_items
        self.cfg = cfg
        h = C.c_void_p()
        rc = self.L.b2c_create(C.byref(cfg), C.byref(h))
        if rc != 0:
            raise B2CError(rc, "b2c_create failed (no sm_100 device? there is no CPU fallback)")
        self.h = h
        if world_aabb is not None:   # AxisSweep3(worldAabbMin, worldAabbMax) for the SAP modes
            mn = np.ascontiguousarray(world_aabb[0], dtype=np.float32)
            mx = np.ascontiguousarray(world_aabb[1], dtype=np.float32)
            self._ck(self.L.b2c_set_world_aabb(self.h, _vp(mn), _vp(mx)))
        if raw_records:   # inspection channel: the raw detector record of EVERY dispatched pair (tests, debugging)
            self._ck(self.L.b2c_set_raw_records(self.h, 1))
        self.num_bodies = 0
        self._broadphase = GpuBroadphase(self)
        self._dispatcher = GpuDispatcher(self)

    def close(self):
        if getattr(self, "h", None):
            self.L.b2c_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc != 0:
            raise B2CError(rc, self.L.b2c_last_error_string(self.h).decode())

    # ---- shapes (constructors of sh/*Shape.java) ----
    def BoxShape(self, half_extents, margin=-1.0):
        he = np.asarray(half_extents, dtype=np.float32)
        out = C.c_int32()
        self._ck(self.L.b2c_shape_register_box(self.h, _vp(he), margin, C.byref(out)))
        return out.value

    def SphereShape(self, radius):
        out = C.c_int32()
        self._ck(self.L.b2c_shape_register_sphere(self.h, radius, C.byref(out)))
        return out.value

    def ConvexHullShape(self, points, margin=-1.0):
        p = np.ascontiguousarray(points, dtype=np.float32).reshape(-1, 3)
        out = C.c_int32()
        self._ck(self.L.b2c_shape_register_hull(self.h, _vp(p), len(p), margin, C.byref(out)))
        return out.value

    def StaticPlaneShape(self, normal, constant):
        n = np.asarray(normal, dtype=np.float32)
        out = C.c_int32()
        self._ck(self.L.b2c_shape_register_plane(self.h, _vp(n), constant, C.byref(out)))
        return out.value

    def BvhTriangleMeshShape(self, vertices, indices, scaling=(1.0, 1.0, 1.0)):
        v = np.ascontiguousarray(vertices, dtype=np.float32).reshape(-1, 3)
        i = np.ascontiguousarray(indices, dtype=np.int32).reshape(-1, 3)
        s = np.asarray(scaling, dtype=np.float32)
        out = C.c_int32()
        self._ck(self.L.b2c_shape_register_mesh(self.h, _vp(v), len(v), 12, _vp(i), len(i), 12, _vp(s), C.byref(out)))
        return out.value

    def BvhTriangleMeshShapeParts(self, parts, scaling=(1.0, 1.0, 1.0)):
        """BvhTriangleMeshShape over a TriangleIndexVertexArray of several IndexedMesh parts (sh/TriangleIndexVertexArray.java:
        72-100): parts = [(vertices (n,3) float32, indices (m,3) uint16 -> ScalarType.SHORT | int32 -> INTEGER), ...]."""
        keep = []
        arr = (_lib.IndexedMesh * len(parts))()
        for k, (v, i) in enumerate(parts):
            v = np.ascontiguousarray(v, dtype=np.float32).reshape(-1, 3)
            i = np.ascontiguousarray(i)
            if i.dtype not in (np.uint16, np.int32):
                i = np.ascontiguousarray(i, dtype=np.int32)
            i = i.reshape(-1, 3)
            keep += [v, i]
            arr[k] = _lib.IndexedMesh(v.ctypes.data, len(v), 12, i.ctypes.data, len(i), 3 * i.itemsize, i.itemsize)
        s = np.asarray(scaling, dtype=np.float32)
        out = C.c_int32()
        self._ck(self.L.b2c_shape_register_mesh_parts(self.h, arr, len(parts), _vp(s), C.byref(out)))
        return out.value

    def CompoundShape(self, child_shapes, child_transforms):
        """new CompoundShape() + addChildShape(child_transforms[i], child_shapes[i]) in order (sh/CompoundShape.java:50-82)."""
        cs = np.ascontiguousarray(child_shapes, dtype=np.int32)
        xf = np.ascontiguousarray(child_transforms, dtype=np.float32).reshape(-1, 12)
        if len(cs) != len(xf):
            raise ValueError("one transform per child")
        out = C.c_int32()
        self._ck(self.L.b2c_shape_register_compound(self.h, len(cs), _vp(cs), _vp(xf), C.byref(out)))
        return out.value

    def mesh_bvh(self, shape):
        n = C.c_int32()
        q = np.zeros(9, dtype=np.float32)
        self._ck(self.L.b2c_mesh_get_bvh(self.h, shape, None, 0, C.byref(n), _vp(q)))
        nodes = np.zeros((n.value, 4), dtype=np.int32)
        self._ck(self.L.b2c_mesh_get_bvh(self.h, shape, _vp(nodes), n.value, C.byref(n), _vp(q)))
        return nodes, q

    # ---- disp/CollisionWorld.java:98-121 ----
    def addCollisionObject(self, shape, transform12, group=DEFAULT_FILTER, mask=ALL_FILTER, static=False, world=0):
        t = np.ascontiguousarray(transform12, dtype=np.float32)
        out = C.c_int32()
        self._ck(self.L.b2c_proxy_create(self.h, shape, _vp(t), group, mask, 1 if static else 0, world, C.byref(out)))
        self.num_bodies = max(self.num_bodies, out.value)  # a recycled slot returns an old uid
        return out.value

    def addCollisionObjects(self, shapes, transforms, groups=None, masks=None, static=None, worlds=None):
        """Batched addCollisionObject: transforms (n,12)."""
        shapes = np.ascontiguousarray(shapes, dtype=np.int32)
        n = len(shapes)
        planes = transforms_to_planes(transforms)
        groups = np.full(n, DEFAULT_FILTER, np.int16) if groups is None else np.ascontiguousarray(groups, dtype=np.int16)
        masks = np.full(n, ALL_FILTER, np.int16) if masks is None else np.ascontiguousarray(masks, dtype=np.int16)
        flags = np.zeros(n, np.int32) if static is None else np.ascontiguousarray(static, dtype=np.int32)
        wl = None if worlds is None else np.ascontiguousarray(worlds, dtype=np.int32)
        out = C.c_int32()
        self._ck(self.L.b2c_proxy_create_batch(self.h, n, _vp(shapes), _vp(planes), _vp(groups), _vp(masks), _vp(flags), _vp(wl),
                                               C.byref(out)))
        self.num_bodies += n
        return out.value

    def removeCollisionObject(self, uid):
        self._ck(self.L.b2c_proxy_destroy(self.h, uid))

    def setMaterial(self, uid, friction, restitution):
        self._ck(self.L.b2c_proxy_set_material(self.h, uid, friction, restitution))

    # ---- per-step inputs ----
    def setWorldTransforms(self, transforms, uids=None):
        planes = transforms_to_planes(transforms)
        n = planes.shape[1]
        u = None if uids is None else np.ascontiguousarray(uids, dtype=np.int32)
        self._ck(self.L.b2c_set_transforms(self.h, n, _vp(u), _vp(planes)))

    def setWorldTransformPlanes(self, planes):
        """planes: (12,n) float32 C-contiguous (already in the ABI's SoA layout; no host repack)."""
        assert planes.dtype == np.float32 and planes.flags.c_contiguous and planes.shape[0] == 12
        self._ck(self.L.b2c_set_transforms(self.h, planes.shape[1], None, _vp(planes)))

    def setWorldTransformsDevice(self, n, device_ptr):
        """12 planes of n floats already on the device (D2D on the ctx stream)."""
        self._ck(self.L.b2c_set_transforms_device(self.h, n, C.c_void_p(device_ptr)))

    def setWorldTransformsHostPtr(self, n, host_ptr):
        """12 planes of n floats at a raw (pinned) host address."""
        self._ck(self.L.b2c_set_transforms(self.h, n, None, C.c_void_p(host_ptr)))

    def setActivation(self, active, uids=None):
        a = np.ascontiguousarray(active, dtype=np.uint8)
        u = None if uids is None else np.ascontiguousarray(uids, dtype=np.int32)
        self._ck(self.L.b2c_set_activation(self.h, len(a), _vp(u), _vp(a)))

    # ---- the path ----
    def updateAabbs(self):  # disp/CollisionWorld.java:231-245
        self._ck(self.L.b2c_update_aabbs(self.h))

    def getBroadphase(self):
        return self._broadphase

    def getPairCache(self):
        return self._broadphase.getOverlappingPairCache()

    def getDispatcher(self):
        return self._dispatcher

    def performDiscreteCollisionDetection(self):  # disp/CollisionWorld.java:123-151
        self.updateAabbs()
        self._broadphase.calculateOverlappingPairs(self._dispatcher)
        self._dispatcher.dispatchAllCollisionPairs(self._broadphase.getOverlappingPairCache(), None, self._dispatcher)

    def step(self, planes=None):
        """One fused b2c_step: (pairs, manifolds, contacts added)."""
        a, b, c = C.c_int32(), C.c_int32(), C.c_int32()
        n = 0 if planes is None else planes.shape[1]
        self._ck(self.L.b2c_step(self.h, n, _vp(planes), C.byref(a), C.byref(b), C.byref(c)))
        return a.value, b.value, c.value

    def step_device(self):
        self._ck(self.L.b2c_step_device(self.h))

    def rayTestClosest(self, ray_from, ray_to, group=1, mask=-1):
        """CollisionWorld.rayTest + ClosestRayResultCallback for a batch of rays: (uid (0 = miss), fraction, normal, point)."""
        f = np.ascontiguousarray(ray_from, dtype=np.float32).reshape(-1, 3)
        t = np.ascontiguousarray(ray_to, dtype=np.float32).reshape(-1, 3)
        n = len(f)
        uid = np.zeros(n, dtype=np.int32)
        frac = np.zeros(n, dtype=np.float32)
        nrm = np.zeros((n, 3), dtype=np.float32)
        pt = np.zeros((n, 3), dtype=np.float32)
        if n:
            self._ck(self.L.b2c_ray_test_closest(self.h, n, _vp(f), _vp(t), int(group), int(mask), _vp(uid), _vp(frac), _vp(nrm), _vp(pt)))
        return uid, frac, nrm, pt

    def convexSweepTestClosest(self, cast_shapes, basis, sweep_from, sweep_to, group=1, mask=-1, allowed_ccd_penetration=0.04):
        """CollisionWorld.convexSweepTest + ClosestConvexResultCallback (disp/CollisionWorld.java:596-651, 765-800) for a batch
        of translational sweeps: cast_shapes = registered convex shape ids (one, or one per sweep), basis = one 3x3 (shared) or n
        of them, sweep_from / sweep_to = n x 3 origins.  Returns (uid (0 = miss, -1 = the sweep reached the reference's
        broken static-plane branch), fraction, normal, point)."""
        f = np.ascontiguousarray(sweep_from, dtype=np.float32).reshape(-1, 3)
        t = np.ascontiguousarray(sweep_to, dtype=np.float32).reshape(-1, 3)
        n = len(f)
        if len(t) != n:
            raise ValueError("sweep_from and sweep_to must hold the same number of origins")
        ids = np.full(n, int(cast_shapes), np.int32) if np.isscalar(cast_shapes) else np.ascontiguousarray(cast_shapes, dtype=np.int32)
        b = np.ascontiguousarray(basis, dtype=np.float32).reshape(-1, 9)
        if len(b) == 1 and n != 1:
            b = np.ascontiguousarray(np.repeat(b, n, axis=0))
        if len(ids) != n or len(b) != n:
            raise ValueError("one cast shape and one basis per sweep")
        uid = np.zeros(n, dtype=np.int32)
        frac = np.zeros(n, dtype=np.float32)
        nrm = np.zeros((n, 3), dtype=np.float32)
        pt = np.zeros((n, 3), dtype=np.float32)
        if n:
            self._ck(self.L.b2c_convex_sweep_closest(self.h, n, _vp(ids), _vp(b), _vp(f), _vp(t), int(group), int(mask),
                                                     float(allowed_ccd_penetration), _vp(uid), _vp(frac), _vp(nrm), _vp(pt)))
        return uid, frac, nrm, pt

    def ccdSweepNotMe(self, body_uids, ccd_radius, predicted_origins, allowed_ccd_penetration=0.04):
        """The CCD motion-clamping sweeps of DiscreteDynamicsWorld.integrateTransforms (dyn/DiscreteDynamicsWorld.java:700-729)
        for the listed bodies: a sphere of ccd_radius swept from each body's resident transform to its predicted origin with a
        ClosestNotMeConvexResultCallback.  Returns (hit uid (0 = none), closestHitFraction, normal, point)."""
        u = np.ascontiguousarray(body_uids, dtype=np.int32).reshape(-1)
        n = len(u)
        r = np.full(n, float(ccd_radius), np.float32) if np.isscalar(ccd_radius) else np.ascontiguousarray(ccd_radius, dtype=np.float32)
        t = np.ascontiguousarray(predicted_origins, dtype=np.float32).reshape(-1, 3)
        if len(r) != n or len(t) != n:
            raise ValueError("one radius and one predicted origin per body")
        uid = np.zeros(n, dtype=np.int32)
        frac = np.zeros(n, dtype=np.float32)
        nrm = np.zeros((n, 3), dtype=np.float32)
        pt = np.zeros((n, 3), dtype=np.float32)
        if n:
            self._ck(self.L.b2c_ccd_sweep_not_me(self.h, n, _vp(u), _vp(r), _vp(t), float(allowed_ccd_penetration), _vp(uid), _vp(frac),
                                                 _vp(nrm), _vp(pt)))
        return uid, frac, nrm, pt

    def setNoCollidePairs(self, pairs):
        """Body pairs linked by a collision-disabling constraint (dynamics/RigidBody.java:624-639): kept in the pair cache,
        never dispatched."""
        p = np.ascontiguousarray(pairs, dtype=np.int32).reshape(-1, 2)
        self._ck(self.L.b2c_set_no_collide_pairs(self.h, len(p), _vp(p) if len(p) else None))

    def sync_counts(self):
        a, b, c = C.c_int32(), C.c_int32(), C.c_int32()
        self._ck(self.L.b2c_sync_counts(self.h, C.byref(a), C.byref(b), C.byref(c)))
        return a.value, b.value, c.value

    def transforms_written(self, n):
        self._ck(self.L.b2c_transforms_written(self.h, n))

    # ---- one world partitioned over several GPUs (include/b2c.h, last section) ----
    def set_partition(self, rank, nranks):
        self._ck(self.L.b2c_set_partition(self.h, rank, nranks))

    def set_partition_slabs(self, rank, nranks, axis, planes):
        pl = np.ascontiguousarray(planes, dtype=np.float32)
        assert len(pl) == nranks - 1
        self._ck(self.L.b2c_set_partition_slabs(self.h, rank, nranks, axis, _vp(pl) if len(pl) else None))

    def partition(self):
        """(axis, planes, owner rank per proxy) of the slab partition in force."""
        axis = C.c_int32()
        planes = np.zeros(15, dtype=np.float32)
        owner = np.zeros(max(self.num_bodies, 1), dtype=np.uint8)
        self._ck(self.L.b2c_get_partition(self.h, C.byref(axis), _vp(planes), _vp(owner), self.num_bodies))
        return axis.value, planes, owner[: self.num_bodies]

    def mgpu_halo_slot_bytes(self, cap):
        return int(self.L.b2c_mgpu_halo_slot_bytes(cap))

    def mgpu_update_export_halo(self, slot_ptr, cap):
        self._ck(self.L.b2c_mgpu_update_export_halo(self.h, C.c_void_p(slot_ptr), cap))

    def mgpu_import_halo(self, slots_ptr, nslots, cap):
        self._ck(self.L.b2c_mgpu_import_halo(self.h, C.c_void_p(slots_ptr), nslots, cap))

    def mgpu_p2p_init(self, cap, migrate_cap=2048):
        """Allocate this rank's inboxes (halo records, migrating manifolds) for the peer-to-peer exchange: (64-byte CUDA IPC
        handle, device pointer)."""
        handle = (C.c_ubyte * 64)()
        ptr = C.c_void_p()
        self._ck(self.L.b2c_mgpu_p2p_init(self.h, int(cap), int(migrate_cap), handle, C.byref(ptr)))
        return bytes(handle), int(ptr.value)

    def mgpu_p2p_connect(self, ipc_handles=None, inbox_ptrs=None):
        """Map the peers' inboxes: ipc_handles = nranks x 64 bytes in rank order (other processes), or inbox_ptrs = nranks
        device pointers (ranks in this process)."""
        if inbox_ptrs is not None:
            arr = (C.c_void_p * len(inbox_ptrs))(*[C.c_void_p(int(p)) for p in inbox_ptrs])
            self._ck(self.L.b2c_mgpu_p2p_connect(self.h, None, arr))
        else:
            buf = (C.c_ubyte * len(ipc_handles)).from_buffer_copy(bytes(ipc_handles))
            self._ck(self.L.b2c_mgpu_p2p_connect(self.h, buf, None))

    def mgpu_p2p_export_halo(self):
        self._ck(self.L.b2c_mgpu_p2p_export_halo(self.h))

    def mgpu_p2p_import_halo(self):
        self._ck(self.L.b2c_mgpu_p2p_import_halo(self.h))

    def mgpu_p2p_export_departed(self):
        self._ck(self.L.b2c_mgpu_p2p_export_departed(self.h))

    def mgpu_p2p_import_arrivals(self):
        self._ck(self.L.b2c_mgpu_p2p_import_arrivals(self.h))

    def mgpu_broadphase(self):
        self._ck(self.L.b2c_mgpu_broadphase(self.h))

    def mgpu_export_departed(self, keys_ptr, hdr_ptr, pts_ptr, cap):
        n = C.c_int32()
        self._ck(self.L.b2c_mgpu_export_departed(self.h, C.c_void_p(keys_ptr), C.c_void_p(hdr_ptr), C.c_void_p(pts_ptr), cap, C.byref(n)))
        return n.value

    def mgpu_import_arrivals(self, keys_ptr, hdr_ptr, pts_ptr, count):
        self._ck(self.L.b2c_mgpu_import_arrivals(self.h, C.c_void_p(keys_ptr), C.c_void_p(hdr_ptr), C.c_void_p(pts_ptr), count))

    def mgpu_slot_bytes(self, cap):
        return int(self.L.b2c_mgpu_slot_bytes(cap))

    def mgpu_export_departed_slot(self, slot_ptr, cap):
        self._ck(self.L.b2c_mgpu_export_departed_slot(self.h, C.c_void_p(slot_ptr), cap))

    def mgpu_import_arrival_slots(self, slots_ptr, nslots, cap):
        self._ck(self.L.b2c_mgpu_import_arrival_slots(self.h, C.c_void_p(slots_ptr), nslots, cap))

    def mgpu_narrowphase(self):
        self._ck(self.L.b2c_mgpu_narrowphase(self.h))

    # ---- results ----
    def aabbs(self):
        out = np.zeros((self.num_bodies, 6), dtype=np.float32)
        if self.num_bodies:
            self._ck(self.L.b2c_get_aabbs(self.h, _vp(out), self.num_bodies))
        return out

    def pairs(self):
        n = C.c_int32()
        self._ck(self.L.b2c_get_pairs(self.h, None, 0, C.byref(n)))
        out = np.zeros((n.value, 2), dtype=np.int32)
        if n.value:
            self._ck(self.L.b2c_get_pairs(self.h, _vp(out), n.value, C.byref(n)))
        return out

    def manifolds(self, only_touching=False):
        n = C.c_int32()
        self._ck(self.L.b2c_get_manifolds(self.h, None, 0, int(only_touching), C.byref(n)))
        out = np.zeros(n.value, dtype=MANIFOLD_DTYPE)
        if n.value:
            self._ck(self.L.b2c_get_manifolds(self.h, _vp(out), n.value, int(only_touching), C.byref(n)))
        return out

    def pair_deltas(self):
        """(added, removed) pairs of the last pair calculation as sorted (n,2) arrays — the events the reference's pair cache
        sends to its ghost pair callback (bp/HashedOverlappingPairCache.java:135-137, 323-325)."""
        na, nr = C.c_int32(), C.c_int32()
        self._ck(self.L.b2c_get_pair_deltas(self.h, None, 0, None, 0, C.byref(na), C.byref(nr)))
        a = np.zeros((na.value, 2), dtype=np.int32)
        r = np.zeros((nr.value, 2), dtype=np.int32)
        self._ck(self.L.b2c_get_pair_deltas(self.h, _vp(a) if na.value else None, na.value, _vp(r) if nr.value else None, nr.value,
                                            C.byref(na), C.byref(nr)))
        a = a[np.lexsort((a[:, 1], a[:, 0]))] if len(a) else a
        r = r[np.lexsort((r[:, 1], r[:, 0]))] if len(r) else r
        return a, r

    def islands(self):
        """Island tag per body (index = uid-1; -1 = static) and the island count
        (disp/SimulationIslandManager.java:57-110)."""
        t = np.zeros(max(self.num_bodies, 1), dtype=np.int32)
        n = C.c_int32()
        self._ck(self.L.b2c_compute_islands(self.h, _vp(t), self.num_bodies, C.byref(n)))
        return t[: self.num_bodies], n.value

    def raw_contacts(self):
        n = C.c_int32()
        self._ck(self.L.b2c_get_raw_contacts(self.h, None, 0, C.byref(n)))
        out = np.zeros(n.value, dtype=RAW_DTYPE)
        if n.value:
            self._ck(self.L.b2c_get_raw_contacts(self.h, _vp(out), n.value, C.byref(n)))
        return out

    def contacts(self):
        """Compact contact stream: (headers, points) of the touching manifolds (order unspecified)."""
        nh, npt = C.c_int32(), C.c_int32()
        self._ck(self.L.b2c_get_contacts(self.h, None, 0, None, 0, C.byref(nh), C.byref(npt)))
        hdr = np.zeros(nh.value, dtype=_lib.CONTACT_HEADER_DTYPE)
        pts = np.zeros(npt.value, dtype=_lib.MANIFOLD_DTYPE["points"].base)
        if nh.value:
            self._ck(self.L.b2c_get_contacts(self.h, _vp(hdr), nh.value, _vp(pts), npt.value, C.byref(nh), C.byref(npt)))
        return hdr, pts

    def solver_contacts(self):
        """The contact stream with 64-byte solver points (include/b2c.h b2c_solver_point)."""
        nh, npt = C.c_int32(), C.c_int32()
        self._ck(self.L.b2c_get_solver_contacts(self.h, None, 0, None, 0, C.byref(nh), C.byref(npt)))
        hdr = np.zeros(nh.value, dtype=_lib.CONTACT_HEADER_DTYPE)
        pts = np.zeros(npt.value, dtype=_lib.SOLVER_POINT_DTYPE)
        if nh.value:
            self._ck(self.L.b2c_get_solver_contacts(self.h, _vp(hdr), nh.value, _vp(pts), npt.value, C.byref(nh), C.byref(npt)))
        return hdr, pts

    def packed_contacts(self):
        """The contact stream with 16-byte headers and 48-byte points (include/b2c.h b2c_packed_header / b2c_packed_point)."""
        nh, npt = C.c_int32(), C.c_int32()
        self._ck(self.L.b2c_get_packed_contacts(self.h, None, 0, None, 0, C.byref(nh), C.byref(npt)))
        hdr = np.zeros(nh.value, dtype=_lib.PACKED_HEADER_DTYPE)
        pts = np.zeros(npt.value, dtype=_lib.PACKED_POINT_DTYPE)
        if nh.value:
            self._ck(self.L.b2c_get_packed_contacts(self.h, _vp(hdr), nh.value, _vp(pts), npt.value, C.byref(nh), C.byref(npt)))
        return hdr, pts

    def packed_contacts_uid(self):
        """The packed stream keyed by uids (include/b2c.h b2c_packed_uid_header): needs no pair list on the host."""
        nh, npt = C.c_int32(), C.c_int32()
        self._ck(self.L.b2c_get_packed_contacts_uid(self.h, None, 0, None, 0, C.byref(nh), C.byref(npt)))
        hdr = np.zeros(nh.value, dtype=_lib.PACKED_UID_HEADER_DTYPE)
        pts = np.zeros(npt.value, dtype=_lib.PACKED_POINT_DTYPE)
        if nh.value:
            self._ck(self.L.b2c_get_packed_contacts_uid(self.h, _vp(hdr), nh.value, _vp(pts), npt.value, C.byref(nh), C.byref(npt)))
        return hdr, pts

    def set_pair_delta_prefetch(self, on=True):
        """Compute the pair-cache add / remove events inside every pair calculation (b2c_get_pair_deltas then only copies)."""
        self._ck(self.L.b2c_set_pair_delta_prefetch(self.h, int(on)))

    def set_contact_prefetch(self, fmt):
        """Compact the contact stream (0 full / 1 solver / 2 packed / 3 packed uid-keyed points, -1 off) behind every dispatch."""
        self._ck(self.L.b2c_set_contact_prefetch(self.h, int(fmt)))

    def set_profiling(self, on=True):
        self._ck(self.L.b2c_set_profiling(self.h, int(on)))

    def stage_times(self):
        ms = np.zeros(_lib.NUM_STAGES, dtype=np.float32)
        self._ck(self.L.b2c_get_stage_times(self.h, _vp(ms)))
        return {self.L.b2c_stage_name(k).decode(): float(ms[k]) for k in range(_lib.NUM_STAGES)}

    def gjk_kernel_ms(self):
        """Device time of k_gjk alone in the last profiled step (its own CUDA-event pair)."""
        ms = C.c_float()
        self._ck(self.L.b2c_get_gjk_kernel_time(self.h, C.byref(ms)))
        return float(ms.value)

    def stats(self):
        s = Stats()
        self._ck(self.L.b2c_get_stats(self.h, C.byref(s)))
        return {f[0]: getattr(s, f[0]) for f in Stats._fields_ if f[0] != "pad"}

    def stream(self):
        return self.L.b2c_stream(self.h)


class GpuPairCache:
    """bp/OverlappingPairCache.java:34-54 view over the device pair list."""

    def __init__(self, world):
        self.w = world
        self.num = 0

    def getOverlappingPairArray(self):
        return self.w.pairs()

    def getNumOverlappingPairs(self):
        return self.num


class GpuBroadphase:
    """bp/BroadphaseInterface.java:33-52."""

    def __init__(self, world):
        self.w = world
        self.cache = GpuPairCache(world)

    def createProxy(self, aabbMin, aabbMax, shape, transform12, group, mask, dispatcher=None, multiSapProxy=None, static=False):
        # the reference passes the shape's AABB; the device recomputes the identical box from shape + transform
        return self.w.addCollisionObject(shape, transform12, group, mask, static)

    def destroyProxy(self, uid, dispatcher=None):
        self.w.removeCollisionObject(uid)

    def setAabb(self, uid, aabbMin, aabbMax, dispatcher=None):
        mm = np.concatenate([np.asarray(aabbMin, np.float32), np.asarray(aabbMax, np.float32)]).reshape(6, 1)
        u = np.asarray([uid], dtype=np.int32)
        self.w._ck(self.w.L.b2c_set_aabbs(self.w.h, 1, _vp(u), _vp(np.ascontiguousarray(mm))))

    def setAabbs(self, uids, mins, maxs):
        mm = np.ascontiguousarray(np.concatenate([np.asarray(mins, np.float32).reshape(-1, 3).T,
                                                  np.asarray(maxs, np.float32).reshape(-1, 3).T], axis=0))
        u = None if uids is None else np.ascontiguousarray(uids, dtype=np.int32)
        self.w._ck(self.w.L.b2c_set_aabbs(self.w.h, mm.shape[1], _vp(u), _vp(mm)))

    def calculateOverlappingPairs(self, dispatcher=None):
        n = C.c_int32()
        self.w._ck(self.w.L.b2c_calculate_overlapping_pairs(self.w.h, C.byref(n)))
        self.cache.num = n.value
        return n.value

    def getOverlappingPairCache(self):
        return self.cache

    def getBroadphaseAabb(self):
        mn, mx = np.zeros(3, np.float32), np.zeros(3, np.float32)
        self.w._ck(self.w.L.b2c_get_broadphase_aabb(self.w.h, _vp(mn), _vp(mx)))
        return mn, mx


class GpuDispatcher:
    """bp/Dispatcher.java:38-68 (default near callback + default collision configuration)."""

    def __init__(self, world):
        self.w = world
        self.num_manifolds = 0
        self.num_contacts_added = 0
        self._manifolds = None

    def dispatchAllCollisionPairs(self, pairCache=None, dispatchInfo=None, dispatcher=None):
        a, b = C.c_int32(), C.c_int32()
        self.w._ck(self.w.L.b2c_dispatch_all_pairs(self.w.h, C.byref(a), C.byref(b)))
        self.num_manifolds, self.num_contacts_added = a.value, b.value
        self._manifolds = None

    def getNumManifolds(self):
        return self.num_manifolds

    def getManifoldByIndexInternal(self, i):
        if self._manifolds is None:
            self._manifolds = self.w.manifolds()
        return self._manifolds[i]
