"""ctypes loader for libb2c.so (the C ABI declared in include/b2c.h).

The library is built in-tree by build.py (nvcc, sm_100a).  There is no CPU fallback anywhere in this
package: if the shared library is missing or no sm_100 device is visible, the calls fail loudly.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libb2c.so")


class B2CError(RuntimeError):
    def __init__(self, code, msg=""):
        super().__init__(f"b2c error {code}: {msg}")
        self.code = code


class Config(C.Structure):
    _fields_ = [
        ("device", C.c_int32), ("broadphase_mode", C.c_int32), ("max_bodies", C.c_int32), ("max_pairs", C.c_int32),
        ("max_shapes", C.c_int32), ("max_hull_points", C.c_int32), ("max_mesh_items", C.c_int32), ("num_worlds", C.c_int32),
        ("contact_breaking_threshold", C.c_float), ("dbvt_margin", C.c_float), ("dbvt_predicted_frames", C.c_float),
        ("max_compound_items", C.c_int32), ("reserved", C.c_int32 * 4),
    ]


class IndexedMesh(C.Structure):
    """b2c_indexed_mesh (sh/IndexedMesh.java:35-47 + its ScalarType)."""
    _fields_ = [("vertex_base", C.c_void_p), ("num_vertices", C.c_int32), ("vertex_stride", C.c_int32),
                ("index_base", C.c_void_p), ("num_triangles", C.c_int32), ("index_stride", C.c_int32), ("index_type", C.c_int32)]


class Stats(C.Structure):
    _fields_ = [
        ("num_pairs", C.c_int32), ("num_manifolds", C.c_int32), ("num_contacts_added", C.c_int32),
        ("gjk_checks", C.c_int32), ("deep_penetration_checks", C.c_int32), ("epa_failed", C.c_int32),
        ("mesh_items", C.c_int32), ("large_proxies", C.c_int32), ("kernel_launches", C.c_int32), ("grid_rows", C.c_int32),
        ("ms_aabb", C.c_float), ("ms_broadphase", C.c_float), ("ms_narrowphase", C.c_float), ("ms_total", C.c_float),
        ("epa_retries", C.c_int32), ("pad", C.c_int32 * 1),
    ]


MANIFOLD_POINT_DTYPE = np.dtype([
    ("local_a", np.float32, 3), ("local_b", np.float32, 3), ("world_a", np.float32, 3), ("world_b", np.float32, 3),
    ("normal_on_b", np.float32, 3), ("distance", np.float32), ("combined_friction", np.float32),
    ("combined_restitution", np.float32), ("life_time", np.int32), ("src_slot", np.int32), ("part_id1", np.int32),
    ("index1", np.int32), ("pad", np.int32, 2),
])
MANIFOLD_DTYPE = np.dtype([
    ("pair_uid0", np.int32), ("pair_uid1", np.int32), ("body0", np.int32), ("body1", np.int32), ("num_contacts", np.int32),
    ("algorithm", np.int32), ("child0", np.int32), ("child1", np.int32), ("points", MANIFOLD_POINT_DTYPE, 4),
])
RAW_DTYPE = np.dtype([
    ("uid0", np.int32), ("uid1", np.int32), ("tri", np.int32), ("has_contact", np.int32), ("normal", np.float32, 3),
    ("point", np.float32, 3), ("depth", np.float32), ("method", np.int32), ("iters", np.int32), ("pad", np.int32, 1),
])
assert MANIFOLD_POINT_DTYPE.itemsize == 96 and MANIFOLD_DTYPE.itemsize == 416 and RAW_DTYPE.itemsize == 56

# every symbol include/b2c.h declares (checked by the CPU test-suite against the built library)
EXPORTS = [
    "b2c_default_config", "b2c_create", "b2c_destroy", "b2c_last_error_string", "b2c_device_count",
    "b2c_shape_register_box", "b2c_shape_register_sphere", "b2c_shape_register_hull", "b2c_shape_register_plane",
    "b2c_shape_register_mesh", "b2c_mesh_get_bvh", "b2c_proxy_create", "b2c_proxy_create_batch", "b2c_proxy_destroy",
    "b2c_proxy_set_material", "b2c_set_transforms", "b2c_set_activation", "b2c_set_aabbs", "b2c_update_aabbs",
    "b2c_calculate_overlapping_pairs", "b2c_get_pairs", "b2c_dispatch_all_pairs", "b2c_step", "b2c_get_manifolds",
    "b2c_get_raw_contacts", "b2c_get_aabbs", "b2c_get_broadphase_aabb", "b2c_get_stats", "b2c_stream",
    "b2c_device_transforms", "b2c_transforms_written", "b2c_step_device", "b2c_sync_counts", "b2c_get_contacts",
    "b2c_set_profiling", "b2c_get_stage_times", "b2c_get_gjk_kernel_time", "b2c_mgpu_p2p_init", "b2c_mgpu_p2p_connect", "b2c_mgpu_p2p_export_halo", "b2c_mgpu_p2p_import_halo", "b2c_mgpu_p2p_export_departed", "b2c_mgpu_p2p_import_arrivals", "b2c_stage_name", "b2c_set_transforms_device", "b2c_set_partition", "b2c_mgpu_broadphase", "b2c_mgpu_export_departed",
    "b2c_mgpu_import_arrivals", "b2c_mgpu_narrowphase", "b2c_mgpu_slot_bytes", "b2c_mgpu_export_departed_slot",
    "b2c_mgpu_import_arrival_slots", "b2c_get_pair_deltas", "b2c_compute_islands", "b2c_get_solver_contacts", "b2c_set_world_aabb", "b2c_set_no_collide_pairs", "b2c_ray_test_closest", "b2c_convex_sweep_closest", "b2c_ccd_sweep_not_me", "b2c_set_raw_records",
    "b2c_shape_register_compound", "b2c_get_packed_contacts", "b2c_set_contact_prefetch", "b2c_begin_contact_download",
    "b2c_get_packed_contacts_uid", "b2c_set_pair_delta_prefetch", "b2c_set_partition_slabs", "b2c_get_partition",
    "b2c_mgpu_halo_slot_bytes", "b2c_mgpu_update_export_halo", "b2c_mgpu_import_halo", "b2c_shape_register_mesh_parts",
]
NUM_STAGES = 12
CONTACT_HEADER_DTYPE = np.dtype([
    ("pair_uid0", np.int32), ("pair_uid1", np.int32), ("body0", np.int32), ("body1", np.int32), ("num_contacts", np.int32),
    ("algorithm", np.int32), ("first_point", np.int32), ("pair_index", np.int32),
])
assert CONTACT_HEADER_DTYPE.itemsize == 32
SOLVER_POINT_DTYPE = np.dtype([
    ("world_a", np.float32, 3), ("world_b", np.float32, 3), ("normal_on_b", np.float32, 3), ("distance", np.float32),
    ("combined_friction", np.float32), ("combined_restitution", np.float32), ("life_time", np.int32), ("src_slot", np.int32),
    ("part_id1", np.int32), ("index1", np.int32),
])
assert SOLVER_POINT_DTYPE.itemsize == 64
PACKED_HEADER_DTYPE = np.dtype([("pair_index", np.int32), ("first_point", np.int32), ("info", np.int32), ("children", np.int32)])
PACKED_POINT_DTYPE = np.dtype([
    ("world_a", np.float32, 3), ("world_b", np.float32, 3), ("normal_on_b", np.float32, 3), ("distance", np.float32),
    ("life_src", np.int32), ("index1", np.int32),
])
PACKED_UID_HEADER_DTYPE = np.dtype([("pair_uid0", np.int32), ("pair_uid1", np.int32), ("info", np.int32), ("children", np.int32)])
assert PACKED_HEADER_DTYPE.itemsize == 16 and PACKED_POINT_DTYPE.itemsize == 48 and PACKED_UID_HEADER_DTYPE.itemsize == 16

_lib = None


def load():
    """Load libb2c.so.  Raises if it has not been built (run __graft_entry__.build() or build.py)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise FileNotFoundError(f"{LIB_PATH} is missing: build it with `python {os.path.join(HERE, 'build.py')}` "
                                "(there is no CPU fallback)")
    L = C.CDLL(LIB_PATH)
    vp, i32, f32 = C.c_void_p, C.c_int32, C.c_float
    pi32 = C.POINTER(C.c_int32)
    L.b2c_default_config.argtypes = [C.POINTER(Config)]
    L.b2c_default_config.restype = None
    L.b2c_create.argtypes = [C.POINTER(Config), C.POINTER(vp)]
    L.b2c_destroy.argtypes = [vp]
    L.b2c_destroy.restype = None
    L.b2c_last_error_string.argtypes = [vp]
    L.b2c_last_error_string.restype = C.c_char_p
    L.b2c_device_count.argtypes = []
    L.b2c_shape_register_box.argtypes = [vp, vp, f32, pi32]
    L.b2c_shape_register_sphere.argtypes = [vp, f32, pi32]
    L.b2c_shape_register_hull.argtypes = [vp, vp, i32, f32, pi32]
    L.b2c_shape_register_plane.argtypes = [vp, vp, f32, pi32]
    L.b2c_shape_register_mesh.argtypes = [vp, vp, i32, i32, vp, i32, i32, vp, pi32]
    L.b2c_shape_register_mesh_parts.argtypes = [vp, C.POINTER(IndexedMesh), i32, vp, pi32]
    L.b2c_mesh_get_bvh.argtypes = [vp, i32, vp, i32, pi32, vp]
    L.b2c_shape_register_compound.argtypes = [vp, i32, vp, vp, pi32]
    L.b2c_proxy_create.argtypes = [vp, i32, vp, C.c_int16, C.c_int16, i32, i32, pi32]
    L.b2c_proxy_create_batch.argtypes = [vp, i32, vp, vp, vp, vp, vp, vp, pi32]
    L.b2c_proxy_destroy.argtypes = [vp, i32]
    L.b2c_proxy_set_material.argtypes = [vp, i32, f32, f32]
    L.b2c_set_transforms.argtypes = [vp, i32, vp, vp]
    L.b2c_set_activation.argtypes = [vp, i32, vp, vp]
    L.b2c_set_aabbs.argtypes = [vp, i32, vp, vp]
    L.b2c_update_aabbs.argtypes = [vp]
    L.b2c_calculate_overlapping_pairs.argtypes = [vp, pi32]
    L.b2c_get_pairs.argtypes = [vp, vp, i32, pi32]
    L.b2c_dispatch_all_pairs.argtypes = [vp, pi32, pi32]
    L.b2c_step.argtypes = [vp, i32, vp, pi32, pi32, pi32]
    L.b2c_get_manifolds.argtypes = [vp, vp, i32, i32, pi32]
    L.b2c_get_raw_contacts.argtypes = [vp, vp, i32, pi32]
    L.b2c_get_aabbs.argtypes = [vp, vp, i32]
    L.b2c_get_broadphase_aabb.argtypes = [vp, vp, vp]
    L.b2c_get_stats.argtypes = [vp, C.POINTER(Stats)]
    L.b2c_stream.argtypes = [vp]
    L.b2c_stream.restype = vp
    L.b2c_device_transforms.argtypes = [vp]
    L.b2c_device_transforms.restype = vp
    L.b2c_transforms_written.argtypes = [vp, i32]
    L.b2c_step_device.argtypes = [vp]
    L.b2c_sync_counts.argtypes = [vp, pi32, pi32, pi32]
    L.b2c_set_transforms_device.argtypes = [vp, i32, vp]
    L.b2c_set_partition.argtypes = [vp, i32, i32]
    L.b2c_set_partition_slabs.argtypes = [vp, i32, i32, i32, vp]
    L.b2c_get_partition.argtypes = [vp, pi32, vp, vp, i32]
    L.b2c_mgpu_halo_slot_bytes.argtypes = [i32]
    L.b2c_mgpu_halo_slot_bytes.restype = C.c_int64
    L.b2c_mgpu_update_export_halo.argtypes = [vp, vp, i32]
    L.b2c_mgpu_import_halo.argtypes = [vp, vp, i32, i32]
    L.b2c_mgpu_broadphase.argtypes = [vp]
    L.b2c_mgpu_export_departed.argtypes = [vp, vp, vp, vp, i32, pi32]
    L.b2c_mgpu_import_arrivals.argtypes = [vp, vp, vp, vp, i32]
    L.b2c_mgpu_narrowphase.argtypes = [vp]
    L.b2c_set_world_aabb.argtypes = [vp, vp, vp]
    L.b2c_ray_test_closest.argtypes = [vp, i32, vp, vp, C.c_int16, C.c_int16, vp, vp, vp, vp]
    L.b2c_convex_sweep_closest.argtypes = [vp, i32, vp, vp, vp, vp, C.c_int16, C.c_int16, C.c_float, vp, vp, vp, vp]
    L.b2c_set_raw_records.argtypes = [vp, i32]
    L.b2c_ccd_sweep_not_me.argtypes = [vp, i32, vp, vp, vp, C.c_float, vp, vp, vp, vp]
    L.b2c_set_no_collide_pairs.argtypes = [vp, i32, vp]
    L.b2c_get_pair_deltas.argtypes = [vp, vp, i32, vp, i32, pi32, pi32]
    L.b2c_compute_islands.argtypes = [vp, vp, i32, pi32]
    L.b2c_mgpu_slot_bytes.argtypes = [i32]
    L.b2c_mgpu_slot_bytes.restype = C.c_int64
    L.b2c_mgpu_export_departed_slot.argtypes = [vp, vp, i32]
    L.b2c_mgpu_import_arrival_slots.argtypes = [vp, vp, i32, i32]
    L.b2c_get_contacts.argtypes = [vp, vp, i32, vp, i32, pi32, pi32]
    L.b2c_get_solver_contacts.argtypes = [vp, vp, i32, vp, i32, pi32, pi32]
    L.b2c_get_packed_contacts.argtypes = [vp, vp, i32, vp, i32, pi32, pi32]
    L.b2c_get_packed_contacts_uid.argtypes = [vp, vp, i32, vp, i32, pi32, pi32]
    L.b2c_set_pair_delta_prefetch.argtypes = [vp, i32]
    L.b2c_set_contact_prefetch.argtypes = [vp, i32]
    L.b2c_begin_contact_download.argtypes = [vp, vp, i32, vp, i32]
    L.b2c_set_profiling.argtypes = [vp, i32]
    L.b2c_get_stage_times.argtypes = [vp, vp]
    L.b2c_get_gjk_kernel_time.argtypes = [vp, vp]
    L.b2c_mgpu_p2p_init.argtypes = [vp, i32, i32, vp, vp]
    L.b2c_mgpu_p2p_export_departed.argtypes = [vp]
    L.b2c_mgpu_p2p_import_arrivals.argtypes = [vp]
    L.b2c_mgpu_p2p_connect.argtypes = [vp, vp, vp]
    L.b2c_mgpu_p2p_export_halo.argtypes = [vp]
    L.b2c_mgpu_p2p_import_halo.argtypes = [vp]
    L.b2c_stage_name.argtypes = [i32]
    L.b2c_stage_name.restype = C.c_char_p
    _lib = L
    return L
