package com.b200.jbullet;

import java.lang.foreign.*;
import java.lang.invoke.MethodHandle;

import static java.lang.foreign.ValueLayout.*;

/**
 * Panama FFM (JDK 22+) binding of include/b2c.h — the thin C-ABI layer of the B200 collision path.
 * NOT COMPILED IN THIS REPOSITORY'S IMAGE (no JDK): authored against include/b2c.h, see INTEGRATION.md.
 * Every handle below binds exactly one exported symbol of libb2c.so.
 */
final class B2C {
    static final Linker LINKER = Linker.nativeLinker();
    static final SymbolLookup LIB = SymbolLookup.libraryLookup(System.getProperty("b2c.library", "libb2c.so"), Arena.global());

    private static MethodHandle h(String name, FunctionDescriptor fd) {
        return LINKER.downcallHandle(LIB.find(name).orElseThrow(() -> new UnsatisfiedLinkError(name)), fd);
    }

    // b2c_config: 8 x int32, 3 x float, max_compound_items, 4 x int32 reserved (64 bytes)
    static final StructLayout CONFIG = MemoryLayout.structLayout(
        JAVA_INT.withName("device"), JAVA_INT.withName("broadphase_mode"), JAVA_INT.withName("max_bodies"),
        JAVA_INT.withName("max_pairs"), JAVA_INT.withName("max_shapes"), JAVA_INT.withName("max_hull_points"),
        JAVA_INT.withName("max_mesh_items"), JAVA_INT.withName("num_worlds"),
        JAVA_FLOAT.withName("contact_breaking_threshold"), JAVA_FLOAT.withName("dbvt_margin"),
        JAVA_FLOAT.withName("dbvt_predicted_frames"), JAVA_INT.withName("max_compound_items"),
        MemoryLayout.sequenceLayout(4, JAVA_INT).withName("reserved"));

    // b2c_indexed_mesh: one sh/IndexedMesh.java part (40 bytes with the trailing pad of the 8-byte aligned struct)
    static final StructLayout INDEXED_MESH = MemoryLayout.structLayout(
        ADDRESS.withName("vertex_base"), JAVA_INT.withName("num_vertices"), JAVA_INT.withName("vertex_stride"),
        ADDRESS.withName("index_base"), JAVA_INT.withName("num_triangles"), JAVA_INT.withName("index_stride"),
        JAVA_INT.withName("index_type"), MemoryLayout.paddingLayout(4));

    static final MethodHandle defaultConfig = h("b2c_default_config", FunctionDescriptor.ofVoid(ADDRESS));
    static final MethodHandle create = h("b2c_create", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS));
    static final MethodHandle destroy = h("b2c_destroy", FunctionDescriptor.ofVoid(ADDRESS));
    static final MethodHandle lastError = h("b2c_last_error_string", FunctionDescriptor.of(ADDRESS, ADDRESS));
    static final MethodHandle shapeBox = h("b2c_shape_register_box", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, JAVA_FLOAT, ADDRESS));
    static final MethodHandle shapeSphere = h("b2c_shape_register_sphere", FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_FLOAT, ADDRESS));
    static final MethodHandle shapeHull = h("b2c_shape_register_hull", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, JAVA_INT, JAVA_FLOAT, ADDRESS));
    static final MethodHandle shapePlane = h("b2c_shape_register_plane", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, JAVA_FLOAT, ADDRESS));
    static final MethodHandle shapeMesh = h("b2c_shape_register_mesh",
        FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, JAVA_INT, JAVA_INT, ADDRESS, JAVA_INT, JAVA_INT, ADDRESS, ADDRESS));
    // TriangleIndexVertexArray with several IndexedMesh parts / 16-bit indices
    static final MethodHandle shapeMeshParts = h("b2c_shape_register_mesh_parts",
        FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, JAVA_INT, ADDRESS, ADDRESS));
    // new CompoundShape() + addChildShape(localTransform_i, child_i): child shape ids + n x 12 floats
    static final MethodHandle shapeCompound = h("b2c_shape_register_compound",
        FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_INT, ADDRESS, ADDRESS, ADDRESS));
    static final MethodHandle proxyCreate = h("b2c_proxy_create",
        FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_INT, ADDRESS, JAVA_SHORT, JAVA_SHORT, JAVA_INT, JAVA_INT, ADDRESS));
    static final MethodHandle proxyDestroy = h("b2c_proxy_destroy", FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_INT));
    static final MethodHandle proxySetMaterial = h("b2c_proxy_set_material", FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_INT, JAVA_FLOAT, JAVA_FLOAT));
    static final MethodHandle setTransforms = h("b2c_set_transforms", FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_INT, ADDRESS, ADDRESS));
    static final MethodHandle setActivation = h("b2c_set_activation", FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_INT, ADDRESS, ADDRESS));
    static final MethodHandle setAabbs = h("b2c_set_aabbs", FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_INT, ADDRESS, ADDRESS));
    static final MethodHandle updateAabbs = h("b2c_update_aabbs", FunctionDescriptor.of(JAVA_INT, ADDRESS));
    static final MethodHandle calculateOverlappingPairs = h("b2c_calculate_overlapping_pairs", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS));
    static final MethodHandle getPairs = h("b2c_get_pairs", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, JAVA_INT, ADDRESS));
    static final MethodHandle dispatchAllPairs = h("b2c_dispatch_all_pairs", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, ADDRESS));
    static final MethodHandle step = h("b2c_step", FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_INT, ADDRESS, ADDRESS, ADDRESS, ADDRESS));
    static final MethodHandle getContacts = h("b2c_get_contacts",
        FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, JAVA_INT, ADDRESS, JAVA_INT, ADDRESS, ADDRESS));
    // 16-byte headers + 48-byte points: the contact stream with nothing the host can derive itself
    static final MethodHandle getPackedContacts = h("b2c_get_packed_contacts",
        FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, JAVA_INT, ADDRESS, JAVA_INT, ADDRESS, ADDRESS));
    static final MethodHandle setContactPrefetch = h("b2c_set_contact_prefetch", FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_INT));
    static final MethodHandle beginContactDownload = h("b2c_begin_contact_download",
        FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, JAVA_INT, ADDRESS, JAVA_INT));
    static final MethodHandle getManifolds = h("b2c_get_manifolds", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, JAVA_INT, JAVA_INT, ADDRESS));

    // AxisSweep3(worldAabbMin, worldAabbMax): world box of the SAP broadphase modes
    static final MethodHandle setWorldAabb = h("b2c_set_world_aabb", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, ADDRESS));
    // RigidBody.checkCollideWithOverride: constraint-linked body pairs are not dispatched
    static final MethodHandle setNoCollidePairs = h("b2c_set_no_collide_pairs", FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_INT, ADDRESS));
    // CollisionWorld.rayTest + ClosestRayResultCallback, batched
    static final MethodHandle rayTestClosest = h("b2c_ray_test_closest",
        FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_INT, ADDRESS, ADDRESS, JAVA_SHORT, JAVA_SHORT, ADDRESS, ADDRESS, ADDRESS, ADDRESS));
    // CollisionWorld.convexSweepTest + ClosestConvexResultCallback, batched translational sweeps
    static final MethodHandle convexSweepClosest = h("b2c_convex_sweep_closest",
        FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_INT, ADDRESS, ADDRESS, ADDRESS, ADDRESS, JAVA_SHORT, JAVA_SHORT, JAVA_FLOAT, ADDRESS, ADDRESS, ADDRESS, ADDRESS));
    // DiscreteDynamicsWorld.integrateTransforms' CCD motion clamping sweeps (ClosestNotMeConvexResultCallback), batched
    static final MethodHandle ccdSweepNotMe = h("b2c_ccd_sweep_not_me",
        FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_INT, ADDRESS, ADDRESS, ADDRESS, JAVA_FLOAT, ADDRESS, ADDRESS, ADDRESS, ADDRESS));
    // inspection channel: the raw detector record of every dispatched pair (off by default)
    static final MethodHandle setRawRecords = h("b2c_set_raw_records", FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_INT));
    // device-resident stepping: enqueue, download the pair list while the narrowphase runs, then wait for the counts
    static final MethodHandle stepDevice = h("b2c_step_device", FunctionDescriptor.of(JAVA_INT, ADDRESS));
    static final MethodHandle syncCounts = h("b2c_sync_counts", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, ADDRESS, ADDRESS));
    // OverlappingPairCallback / GhostPairCallback events (bp/HashedOverlappingPairCache.java:135-137,323-325)
    static final MethodHandle getPairDeltas = h("b2c_get_pair_deltas",
        FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, JAVA_INT, ADDRESS, JAVA_INT, ADDRESS, ADDRESS));
    // SimulationIslandManager.updateActivationState + storeIslandActivationState (disp/SimulationIslandManager.java:57-110)
    static final MethodHandle computeIslands = h("b2c_compute_islands", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, JAVA_INT, ADDRESS));

    static final MethodHandle setPairDeltaPrefetch = h("b2c_set_pair_delta_prefetch", FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_INT));
    static final MethodHandle getBroadphaseAabb = h("b2c_get_broadphase_aabb", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, ADDRESS));
    // uid-keyed packed contact stream: a host that follows the pair cache through its deltas needs no pair list
    static final MethodHandle getPackedContactsUid = h("b2c_get_packed_contacts_uid",
        FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, JAVA_INT, ADDRESS, JAVA_INT, ADDRESS, ADDRESS));
    // one world over several GPUs (slabs with a halo): the exchanges between these calls are NCCL all-gathers on b2c_stream
    static final MethodHandle setPartition = h("b2c_set_partition", FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_INT, JAVA_INT));
    static final MethodHandle mgpuUpdateExportHalo = h("b2c_mgpu_update_export_halo", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, JAVA_INT));
    static final MethodHandle mgpuImportHalo = h("b2c_mgpu_import_halo", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, JAVA_INT, JAVA_INT));
    static final MethodHandle mgpuBroadphase = h("b2c_mgpu_broadphase", FunctionDescriptor.of(JAVA_INT, ADDRESS));
    static final MethodHandle mgpuExportDepartedSlot = h("b2c_mgpu_export_departed_slot", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, JAVA_INT));
    static final MethodHandle mgpuImportArrivalSlots = h("b2c_mgpu_import_arrival_slots", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, JAVA_INT, JAVA_INT));
    static final MethodHandle mgpuNarrowphase = h("b2c_mgpu_narrowphase", FunctionDescriptor.of(JAVA_INT, ADDRESS));
    // the same exchange as peer-to-peer stores (no collective): inboxes mapped once through CUDA IPC handles
    static final MethodHandle mgpuP2pInit = h("b2c_mgpu_p2p_init", FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_INT, JAVA_INT, ADDRESS, ADDRESS));
    static final MethodHandle mgpuP2pConnect = h("b2c_mgpu_p2p_connect", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, ADDRESS));
    static final MethodHandle mgpuP2pExportHalo = h("b2c_mgpu_p2p_export_halo", FunctionDescriptor.of(JAVA_INT, ADDRESS));
    static final MethodHandle mgpuP2pImportHalo = h("b2c_mgpu_p2p_import_halo", FunctionDescriptor.of(JAVA_INT, ADDRESS));
    static final MethodHandle mgpuP2pExportDeparted = h("b2c_mgpu_p2p_export_departed", FunctionDescriptor.of(JAVA_INT, ADDRESS));
    static final MethodHandle mgpuP2pImportArrivals = h("b2c_mgpu_p2p_import_arrivals", FunctionDescriptor.of(JAVA_INT, ADDRESS));
    static final MethodHandle stream = h("b2c_stream", FunctionDescriptor.of(ADDRESS, ADDRESS));

    static void check(int rc, MemorySegment ctx) {
        if (rc != 0) {
            String msg;
            try {
                msg = ((MemorySegment) lastError.invokeExact(ctx)).reinterpret(256).getString(0);
            } catch (Throwable t) {
                msg = "?";
            }
            throw new IllegalStateException("b2c error " + rc + ": " + msg);
        }
    }

    private B2C() {
    }
}
