/* b2c.h — C ABI of the B200 collision path (broadphase overlap pairs + narrowphase contacts).
 *
 * Drop-in boundary for libgdx-jbullet's per-step collision phase
 *   CollisionWorld.performDiscreteCollisionDetection
 *   (reference: src/com/bulletphysics/collision/dispatch/CollisionWorld.java:123-151)
 * Each entry point below names the reference interface it replaces ("bp/" =
 * collision/broadphase/, "disp/" = collision/dispatch/, "np/" = collision/narrowphase/,
 * "sh/" = collision/shapes/).  A Java host binds these with Panama FFM (or a JNI shim);
 * see INTEGRATION.md.  Plain pointers and sizes only; the caller owns every host buffer and
 * the library borrows it for the duration of the call.  All functions return 0 on success or
 * a negative b2c_status; no exception crosses this boundary.  A ctx is thread-confined (one
 * ctx per world per thread, like the reference's thread-local pools) and owns one CUDA stream.
 *
 * There is no CPU fallback: b2c_create fails with B2C_ERR_CUDA when no sm_100 device is usable.
 */
#ifndef B2C_H
#define B2C_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct b2c_ctx b2c_ctx;

typedef enum {
    B2C_OK = 0,
    B2C_ERR_BAD_ARG = -1,
    B2C_ERR_BAD_HANDLE = -2,
    B2C_ERR_CAPACITY = -3, /* a fixed-capacity buffer overflowed; see b2c_last_error_string */
    B2C_ERR_CUDA = -4,
    B2C_ERR_STATE = -5
} b2c_status;

/* Which reference broadphase defines the exact pair set. */
typedef enum {
    B2C_BP_TIGHT = 0, /* bp/SimpleBroadphase.java:81-110: overlaps of the AABBs last passed to setAabb */
    B2C_BP_DBVT = 1,  /* bp/DbvtBroadphase.java:89-228: overlaps of the per-proxy effective (possibly fattened) AABB */
    /* bp/AxisSweep3.java (16-bit) / bp/AxisSweep3_32.java (31-bit): overlaps of the AABBs QUANTISED over the world box as
     * bp/AxisSweep3Internal.java:201-216 does (min edges even, max edges odd).  The incremental edge sort ends every step
     * with exactly that set (oracle/sap_literal.h + tests/test_oracle_sap.py), so no edge lists are kept.  Set the world box
     * with b2c_set_world_aabb before creating proxies (default +-1000).  Reference uids are handle indices, which it
     * reuses after a removal; here uids are never reused. */
    B2C_BP_SAP16 = 2,
    B2C_BP_SAP32 = 3
} b2c_broadphase_mode;

/* Shape kinds (subset of bp/BroadphaseNativeType.java on the hot path). */
typedef enum { B2C_SHAPE_BOX = 0, B2C_SHAPE_SPHERE = 1, B2C_SHAPE_HULL = 2, B2C_SHAPE_PLANE = 4, B2C_SHAPE_MESH = 5,
               B2C_SHAPE_COMPOUND = 6 } b2c_shape_kind;

typedef struct {
    int32_t device;                    /* CUDA device ordinal */
    int32_t broadphase_mode;           /* b2c_broadphase_mode */
    int32_t max_bodies;                /* LIVE proxy capacity (uids 1..max_bodies; see b2c_proxy_destroy for slot reuse) */
    int32_t max_pairs;                 /* overlapping-pair capacity */
    int32_t max_shapes;                /* shape table capacity */
    int32_t max_hull_points;           /* total hull vertices over all hull shapes */
    int32_t max_mesh_items;            /* (pair, triangle) work items per step for convex-vs-mesh */
    int32_t num_worlds;                /* >1: batched independent worlds (world id per body) */
    float contact_breaking_threshold;  /* BulletGlobals.java:63, default 0.02 */
    float dbvt_margin;                 /* bp/DbvtBroadphase.java:35 DBVT_BP_MARGIN, default 0.05 */
    float dbvt_predicted_frames;       /* bp/DbvtBroadphase.java:73, default 2 */
    int32_t max_compound_items;        /* child work items per step over all CompoundShape pairs (0 = max(65536, max_pairs)) */
    int32_t reserved[4];
} b2c_config;

/* Fills cfg with the reference's defaults (thresholds above; capacities for ~128k bodies). */
void b2c_default_config(b2c_config* cfg);

int32_t b2c_create(const b2c_config* cfg, b2c_ctx** out);
void b2c_destroy(b2c_ctx* ctx);
const char* b2c_last_error_string(const b2c_ctx* ctx);
/* AxisSweep3(worldAabbMin, worldAabbMax, ...) bounds (bp/AxisSweep3.java:43-58) for the SAP modes; before any proxy. */
int32_t b2c_set_world_aabb(b2c_ctx* ctx, const float world_min[3], const float world_max[3]);
/* Library-level check usable without a ctx: number of visible sm_100 devices (0 on a CPU box). */
int32_t b2c_device_count(void);

/* ---- shapes: replace the shape constructors + per-shape parameters the kernels read -------- */
/* sh/BoxShape.java:46-50  new BoxShape(halfExtents); margin = CONVEX_DISTANCE_MARGIN unless >= 0 given */
int32_t b2c_shape_register_box(b2c_ctx*, const float half_extents[3], float margin_or_neg, int32_t* shape_out);
/* sh/SphereShape.java:38-41 */
int32_t b2c_shape_register_sphere(b2c_ctx*, float radius, int32_t* shape_out);
/* sh/ConvexHullShape.java:45-53 (points xyz, tightly packed); localScaling (1,1,1) */
int32_t b2c_shape_register_hull(b2c_ctx*, const float* points_xyz, int32_t num_points, float margin_or_neg, int32_t* shape_out);
/* sh/StaticPlaneShape.java:45-48 */
int32_t b2c_shape_register_plane(b2c_ctx*, const float normal[3], float constant, int32_t* shape_out);
/* sh/BvhTriangleMeshShape.java:68-90 with useQuantizedAabbCompression=true over one
 * sh/IndexedMesh.java part: vertices with a byte stride, 32-bit indices with a byte stride per
 * triangle (sh/TriangleIndexVertexArray.java:51-94; use b2c_shape_register_mesh_parts for 16-bit indices / several parts).  The data is copied; the BVH is built on the host
 * exactly as sh/OptimizedBvh.java:283-342 and uploaded. */
int32_t b2c_shape_register_mesh(b2c_ctx*, const void* vertex_base, int32_t num_vertices, int32_t vertex_stride,
                                const void* index_base, int32_t num_triangles, int32_t index_stride,
                                const float scaling[3], int32_t* shape_out);
/* The same over a sh/TriangleIndexVertexArray.java:72-100 with several parts, each an sh/IndexedMesh.java:35-47 with its own
 * index type (addIndexedMesh(mesh, ScalarType.SHORT | INTEGER); sh/ByteBufferVertexData.java:75-84 reads SHORT indices as
 * unsigned).  index_stride is IndexedMesh.triangleIndexStride (bytes per TRIANGLE; the reference steps stride / 3 per index).
 * Contact points on the mesh report partId1 = the part and index1 = the triangle inside it (b2c_manifold_point); the raw
 * detector records carry partId << 21 | index in `tri` (sh/OptimizedBvh.java:65,278: at most 1024 parts of 2^21 triangles). */
typedef enum { B2C_INDEX_INT16 = 2, B2C_INDEX_INT32 = 4 } b2c_index_type;
typedef struct {
    const void* vertex_base;  int32_t num_vertices;  int32_t vertex_stride;  /* 3 floats per vertex, stride in bytes */
    const void* index_base;   int32_t num_triangles; int32_t index_stride;   /* bytes per triangle */
    int32_t index_type;                                                     /* b2c_index_type */
} b2c_indexed_mesh;
int32_t b2c_shape_register_mesh_parts(b2c_ctx*, const b2c_indexed_mesh* parts, int32_t num_parts, const float scaling[3],
                                      int32_t* shape_out);
/* sh/CompoundShape.java:50-82: new CompoundShape() followed by addChildShape(localTransform_i, child_i) for i = 0..n-1.
 * child_shapes = ids of box / sphere / hull shapes — or of COMPOUND shapes — registered before; child_transforms12 = n x
 * (9 row-major basis floats + origin).  The local AABB is the running Math.min / Math.max of the children's AABBs (:60-80;
 * a nested compound counts with CompoundShape.getAabb of its own local box), collisionMargin stays 0 (:49).
 * Pairs with a compound on either side run disp/CompoundCollisionAlgorithm.java:83-129: one child algorithm and one
 * PersistentManifold per child (per child x child for two compounds), reported by b2c_get_manifolds / b2c_get_contacts with
 * the child indices.  A child that is itself a compound gets the reference's nested CompoundCollisionAlgorithm: its leaves
 * are visited depth first with world transforms composed level by level, ((orgTrans * childTrans) * grandChildTrans) ...,
 * and the child index reported is the leaf's position in that depth-first order.  At most 4 compound levels above a leaf,
 * at most 32767 leaves.  The other object may be a box, sphere, hull, static plane, triangle mesh (ConvexConcave per leaf)
 * or another compound.  Not built: concave children, compounds in a partitioned world. */
int32_t b2c_shape_register_compound(b2c_ctx*, int32_t num_children, const int32_t* child_shapes, const float* child_transforms12,
                                    int32_t* shape_out);
/* Debug/inspection: copy the quantized BVH (16-byte nodes, sh/QuantizedBvhNodes.java:34-48) */
int32_t b2c_mesh_get_bvh(b2c_ctx*, int32_t shape, void* nodes16_out, int32_t cap_nodes, int32_t* num_nodes, float quant9_out[9]);

/* ---- proxies: bp/BroadphaseInterface.java:35-39 ------------------------------------------- */
/* createProxy + CollisionWorld.addCollisionObject (disp/CollisionWorld.java:102-121): the initial
 * AABB is the shape's AABB under `transform12` (9 row-major basis floats + origin), uid = ++gid
 * (bp/DbvtBroadphase.java:179).  flags: bit0 = static object.  world = world id for batched worlds. */
int32_t b2c_proxy_create(b2c_ctx*, int32_t shape, const float transform12[12], int16_t group, int16_t mask,
                         int32_t flags, int32_t world, int32_t* uid_out);
/* The same for `n` proxies in one call (a host shim batches World.addRigidBody loops): shapes[n],
 * transforms as 12 SoA planes of n floats, groups[n], masks[n], flags[n], worlds[n] (NULL = world 0).
 * uids are assigned consecutively; the first one is returned. */
int32_t b2c_proxy_create_batch(b2c_ctx*, int32_t n, const int32_t* shapes, const float* planes12, const int16_t* groups,
                               const int16_t* masks, const int32_t* flags, const int32_t* worlds, int32_t* first_uid_out);
/* destroyProxy (bp/BroadphaseInterface.java:37): pairs containing it disappear at the next calculate.  uids are handed out
 * increasingly (DbvtBroadphase's ++gid) until max_bodies have been created; from then on b2c_proxy_create reuses the slot and
 * uid of a proxy destroyed BEFORE the last pair calculation, lowest first (as AxisSweep3 reuses handles,
 * bp/AxisSweep3Internal.java:380-395), so max_bodies bounds the live proxies, not the creations.  b2c_proxy_create_batch
 * only appends. */
int32_t b2c_proxy_destroy(b2c_ctx*, int32_t uid);
/* CollisionObject friction / restitution (disp/CollisionObject.java:95,71) used by ManifoldResult */
int32_t b2c_proxy_set_material(b2c_ctx*, int32_t uid, float friction, float restitution);

/* ---- per-step inputs ---------------------------------------------------------------------- */
/* World transforms as 12 SoA planes of `n` floats each (m00 m01 m02 m10 ... m22 ox oy oz), plane
 * stride = n floats; uids NULL means bodies 1..n.  One H2D copy. */
int32_t b2c_set_transforms(b2c_ctx*, int32_t n, const int32_t* uids, const float* planes12);
/* CollisionObject.isActive() per body (disp/CollisionObject.java:178-180); 1 = active */
int32_t b2c_set_activation(b2c_ctx*, int32_t n, const int32_t* uids, const uint8_t* active);
/* CollisionObject.checkCollideWith / RigidBody.checkCollideWithOverride (dynamics/RigidBody.java:624-639): bodies linked by
 * a constraint with disableCollisionsBetweenLinkedBodies stay in the pair cache but are not dispatched
 * (disp/CollisionDispatcher.java:216-218).  uid_pairs = 2*n uids, any order; replaces the previous list; n = 0 clears. */
int32_t b2c_set_no_collide_pairs(b2c_ctx*, int32_t n, const int32_t* uid_pairs);
/* BroadphaseInterface.setAabb (bp/BroadphaseInterface.java:39) in bulk, for hosts that compute
 * AABBs themselves: minmax6 = 6 SoA planes of n floats (minx miny minz maxx maxy maxz). */
int32_t b2c_set_aabbs(b2c_ctx*, int32_t n, const int32_t* uids, const float* minmax6);

/* ---- the path ----------------------------------------------------------------------------- */
/* CollisionWorld.updateAabbs (disp/CollisionWorld.java:231-245): shape AABBs on the device from the
 * uploaded transforms, +-threshold, overflow guard, then the broadphase's setAabb state machine. */
int32_t b2c_update_aabbs(b2c_ctx*);
/* BroadphaseInterface.calculateOverlappingPairs (bp/BroadphaseInterface.java:42) */
int32_t b2c_calculate_overlapping_pairs(b2c_ctx*, int32_t* num_pairs_out);
/* OverlappingPairCache.getOverlappingPairArray (bp/OverlappingPairCache.java:36): (uid0<uid1) pairs,
 * sorted lexicographically; 2 int32 per pair. */
int32_t b2c_get_pairs(b2c_ctx*, int32_t* pairs_out, int32_t cap_pairs, int32_t* num_pairs_out);
/* Dispatcher.dispatchAllCollisionPairs (bp/Dispatcher.java:58) with the default near callback and
 * the default collision configuration's algorithm table. */
int32_t b2c_dispatch_all_pairs(b2c_ctx*, int32_t* num_manifolds_out, int32_t* num_contacts_added_out);
/* The whole of performDiscreteCollisionDetection in one call: transforms in (as b2c_set_transforms,
 * planes12 may be NULL to reuse the resident ones), counts out. */
int32_t b2c_step(b2c_ctx*, int32_t n, const float* planes12, int32_t* num_pairs_out, int32_t* num_manifolds_out,
                 int32_t* num_contacts_added_out);

/* ---- results ------------------------------------------------------------------------------ */
typedef struct {
    float local_a[3], local_b[3];     /* np/ManifoldPoint.java:38-39 */
    float world_a[3], world_b[3];     /* positionWorldOnA / positionWorldOnB */
    float normal_on_b[3];             /* normalWorldOnB */
    float distance;                   /* distance1 */
    float combined_friction, combined_restitution;
    int32_t life_time;
    int32_t src_slot;                 /* slot of this manifold at the start of the step this point continues; -1 = new */
    int32_t part_id1, index1;         /* triangle ids for mesh pairs (partId0/index0 are -1 there), else 0 */
    int32_t pad[2];
} b2c_manifold_point; /* 96 bytes */

typedef struct {
    int32_t pair_uid0, pair_uid1;     /* the broadphase pair (uid0 < uid1) */
    int32_t body0, body1;             /* PersistentManifold.getBody0/1 (np/PersistentManifold.java:161-167) */
    int32_t num_contacts;             /* getNumContacts */
    int32_t algorithm;                /* 1 sphere-sphere, 2 convex-plane, 3 convex-convex, 4 convex-concave */
    int32_t child0, child1;           /* child manifold of a compound pair: index of the child in body0's / body1's CompoundShape
                                         (-1 = that object is not a compound); -1, -1 for every other manifold */
    b2c_manifold_point points[4];
} b2c_manifold; /* 416 bytes */

/* Dispatcher.getNumManifolds / getManifoldByIndexInternal (bp/Dispatcher.java:62-64): manifolds in pair
 * order; a compound pair contributes one manifold per child algorithm, in the order
 * disp/CompoundCollisionAlgorithm.java:100-125 runs them.  only_touching != 0 skips manifolds with zero contacts. */
int32_t b2c_get_manifolds(b2c_ctx*, b2c_manifold* out, int32_t cap, int32_t only_touching, int32_t* num_out);

/* Compact contact stream for the solver: only manifolds with >= 1 contact, each a 32-byte header plus its
 * live points (device-side compaction, two exact-size D2H copies).  Order is unspecified; match by uids. */
typedef struct {
    int32_t pair_uid0, pair_uid1, body0, body1;
    int32_t num_contacts, algorithm;
    int32_t first_point;              /* index of this manifold's first point in the point array */
    int32_t pair_index;               /* >= 0: index of the pair in the sorted pair list.  < 0: child manifold of a compound pair;
                                         v = -1 - pair_index, child0 = (v & 0x7fff) - 1, child1 = (v >> 15) - 1 (the child's index in
                                         body0's / body1's CompoundShape, -1 = that object is not a compound) */
} b2c_contact_header; /* 32 bytes */
int32_t b2c_get_contacts(b2c_ctx*, b2c_contact_header* headers_out, int32_t cap_headers, b2c_manifold_point* points_out,
                         int32_t cap_points, int32_t* num_headers_out, int32_t* num_points_out);

/* The same stream with 64-byte points that carry only what the constraint solver reads (world points, normal, distance,
 * combined friction / restitution, lifetime, src_slot, triangle ids) — the manifold cache itself (local points, refresh)
 * stays on the device.  One third fewer bytes over PCIe than b2c_get_contacts. */
typedef struct {
    float world_a[3], world_b[3], normal_on_b[3];
    float distance, combined_friction, combined_restitution;
    int32_t life_time, src_slot, part_id1, index1;
} b2c_solver_point; /* 64 bytes */
int32_t b2c_get_solver_contacts(b2c_ctx*, b2c_contact_header* headers_out, int32_t cap_headers, b2c_solver_point* points_out,
                                int32_t cap_points, int32_t* num_headers_out, int32_t* num_points_out);

/* The same stream with nothing in it that the host can derive itself (bench.py's `e2e` uses this one): 16-byte headers
 * and 48-byte points, one quarter fewer bytes over PCIe than b2c_get_solver_contacts.  The host already holds the pair list
 * (b2c_get_pairs: uid0, uid1 = pairs[pair_index]), the bodies' materials (combined friction = clamp(f0 * f1, +-10),
 * restitution = r0 * r1, disp/ManifoldResult.java:160-175) and positionWorldOnA/B are both kept (the solver reads both). */
typedef struct {
    int32_t pair_index;               /* index of the pair in the sorted pair list (also for child manifolds of compound pairs) */
    int32_t first_point;              /* index of this manifold's first point in the point array */
    int32_t info;                     /* num_contacts | algorithm << 8 | swapped << 16 (swapped: manifold body0 is pairs[pair_index].uid1) */
    int32_t children;                 /* compound child manifold: (uint16)child0 | child1 << 16 (int16 each, -1 = not a compound); else -1 */
} b2c_packed_header; /* 16 bytes */
typedef struct {
    float world_a[3], world_b[3], normal_on_b[3];
    float distance;
    int32_t life_src;                 /* life_time << 8 | (src_slot + 1)  (life_time saturates at 2^24 - 1) */
    int32_t index1;                   /* mesh pairs: partId1 << 21 | index1 (= the triangle index for a one-part mesh), else 0 */
} b2c_packed_point; /* 48 bytes */
int32_t b2c_get_packed_contacts(b2c_ctx*, b2c_packed_header* headers_out, int32_t cap_headers, b2c_packed_point* points_out,
                                int32_t cap_points, int32_t* num_headers_out, int32_t* num_points_out);

/* The packed stream keyed by uids instead of pair indices, for hosts that follow the pair cache through its add / remove
 * events (b2c_get_pair_deltas) and never download the pair list: same 48-byte points; the points of consecutive headers are
 * consecutive, so a manifold's first point is the running sum of num_contacts over the headers before it. */
typedef struct {
    int32_t pair_uid0, pair_uid1;     /* the broadphase pair (uid0 < uid1) */
    int32_t info;                     /* num_contacts | algorithm << 8 | swapped << 16 (swapped: manifold body0 is pair_uid1) */
    int32_t children;                 /* compound child manifold: (uint16)child0 | child1 << 16; else -1 */
} b2c_packed_uid_header; /* 16 bytes */
int32_t b2c_get_packed_contacts_uid(b2c_ctx*, b2c_packed_uid_header* headers_out, int32_t cap_headers, b2c_packed_point* points_out,
                                    int32_t cap_points, int32_t* num_headers_out, int32_t* num_points_out);

/* Compact the contact stream behind every dispatch instead of inside its getter: format 0 = b2c_get_contacts, 1 =
 * b2c_get_solver_contacts, 2 = b2c_get_packed_contacts, 3 = b2c_get_packed_contacts_uid, -1 = off (default).  The compaction then rides in the step's CUDA
 * graph, its counts come back with b2c_sync_counts, and the getter of that format only issues the two exact-size copies. */
int32_t b2c_set_contact_prefetch(b2c_ctx*, int32_t format);
/* With prefetch format 2 or 3: start the download of the packed stream while the step is still running.  Everything outside the
 * penetration bin (and the mesh bin) is final before the EPA kernels end, is compacted there, and this call — issued after
 * b2c_step_device — waits for that point and starts copying that part into the caller's (pinned) buffers on the copy
 * stream.  The following b2c_get_packed_contacts with the SAME buffers only adds what the end of the dispatch appended and
 * waits for both copies.  Optional: without it the getter copies everything. */
int32_t b2c_begin_contact_download(b2c_ctx*, void* headers_out /* b2c_packed_header* or b2c_packed_uid_header* */, int32_t cap_headers,
                                   b2c_packed_point* points_out, int32_t cap_points);

/* Raw detector output per processed pair (or per (pair, triangle)), before ManifoldResult: what
 * DiscreteCollisionDetectorInterface.Result.addContactPoint received. */
typedef struct {
    int32_t uid0, uid1, tri, has_contact; /* tri: triangle index for mesh pairs; -2 - k for child algorithm k of a compound
                                             pair (-2 - (k << 21 | triangle) when the other object is a mesh); else -1 */
    float normal[3], point[3], depth;
    int32_t method; /* GjkPairDetector.lastUsedMethod (np/GjkPairDetector.java:54); 10 sphere-sphere; 11 convex-plane */
    int32_t iters;
    int32_t pad[1];
} b2c_raw_contact; /* 56 bytes */
int32_t b2c_get_raw_contacts(b2c_ctx*, b2c_raw_contact* out, int32_t cap, int32_t* num_out);
/* The records of pairs whose result nobody on the device reads (no contact; closed-form algorithms that update their manifold
 * in place) are only written when this inspection channel is on (default off: 64 bytes per dispatched pair per step).
 * b2c_get_raw_contacts returns B2C_ERR_STATE while it is off.  Takes effect with the next dispatch. */
int32_t b2c_set_raw_records(b2c_ctx*, int32_t on);

/* Effective broadphase AABBs (DbvtProxy.aabb / SimpleBroadphaseProxy min,max) for bodies 1..n: n x 6 floats */
int32_t b2c_get_aabbs(b2c_ctx*, float* minmax_out, int32_t n);
/* BroadphaseInterface.getBroadphaseAabb (bp/BroadphaseInterface.java:48) */
int32_t b2c_get_broadphase_aabb(b2c_ctx*, float min_out[3], float max_out[3]);

/* Counters of the last step (mirrors BulletStats.java:41-56 and adds device timings). */
typedef struct {
    int32_t num_pairs, num_manifolds, num_contacts_added;
    int32_t gjk_checks, deep_penetration_checks; /* BulletStats.gNumGjkChecks / gNumDeepPenetrationChecks */
    int32_t epa_failed, mesh_items, large_proxies;
    int32_t kernel_launches;          /* kernels launched by the last b2c_step */
    int32_t grid_rows;
    float ms_aabb, ms_broadphase, ms_narrowphase, ms_total; /* CUDA-event times of the last b2c_step */
    int32_t epa_retries;              /* penetration-bin items redone with the large pools */
    int32_t pad[1];
} b2c_stats;
int32_t b2c_get_stats(b2c_ctx*, b2c_stats* out);

/* Per-stage device times of the last b2c_step/b2c_step_device (CUDA events on the ctx stream around each
 * kernel group).  Enable with b2c_set_profiling(ctx, 1).  Stage order: see b2c_stage_name. */
#define B2C_NUM_STAGES 12
int32_t b2c_set_profiling(b2c_ctx*, int32_t on);
int32_t b2c_get_stage_times(b2c_ctx*, float ms_out[B2C_NUM_STAGES]);
const char* b2c_stage_name(int32_t stage);
/* Device time of the step's dominant kernel alone (k_gjk: the GJK iterations of all convex-convex pairs that survive the
 * prefilter), from its own CUDA-event pair inside the gjk_mesh stage of the last profiled step. */
int32_t b2c_get_gjk_kernel_time(b2c_ctx*, float* ms_out);

/* ---- device-resident stepping for measurement and host-free pipelines ---------------------- */
/* The ctx's CUDA stream (cudaStream_t) so a caller can time with CUDA events on the right stream. */
void* b2c_stream(b2c_ctx*);
/* Device pointer to the 12 transform planes (plane stride = max_bodies floats): a caller with its own
 * device-side integrator writes here and calls b2c_step_device. */
float* b2c_device_transforms(b2c_ctx*);
/* b2c_set_transforms with a DEVICE pointer (12 planes of n floats, plane stride n): one D2D copy on the ctx
 * stream, for hosts whose integrator already lives on the GPU. */
int32_t b2c_set_transforms_device(b2c_ctx*, int32_t n, const float* device_planes12);
/* Tell the ctx that planes for bodies 1..n were written on the device through b2c_device_transforms. */
int32_t b2c_transforms_written(b2c_ctx*, int32_t n);
/* Enqueue one full collision step on the ctx stream using the resident transforms; does not
 * synchronise and does not copy results to the host.  Counters are read later with b2c_sync_counts. */
int32_t b2c_step_device(b2c_ctx*);
int32_t b2c_sync_counts(b2c_ctx*, int32_t* num_pairs_out, int32_t* num_manifolds_out, int32_t* num_contacts_added_out);

/* ---- consumers directly behind the pair list ------------------------------------------------- */
/* Pairs that entered / left the pair cache in the last b2c_calculate_overlapping_pairs (or b2c_step): what
 * bp/HashedOverlappingPairCache.java:323-325 and :135-137 report to OverlappingPairCallback / GhostPairCallback
 * (disp/GhostPairCallback.java:40-68).  Rows of (uid0 < uid1); order unspecified.  Either output may be NULL to only
 * count.  Valid until the next pair calculation. */
int32_t b2c_get_pair_deltas(b2c_ctx*, int32_t* added_out, int32_t cap_added, int32_t* removed_out, int32_t cap_removed,
                            int32_t* num_added_out, int32_t* num_removed_out);
/* Compute the deltas inside every pair calculation (two more kernels behind the pair ordering): b2c_get_pair_deltas then
 * only waits for the broadphase and copies, i.e. after b2c_step_device its download overlaps the narrowphase. */
int32_t b2c_set_pair_delta_prefetch(b2c_ctx*, int32_t on);
/* SimulationIslandManager.updateActivationState + storeIslandActivationState (disp/SimulationIslandManager.java:57-110):
 * union-find over every broadphase pair whose two objects merge islands (non-static).  tags_out[i] = island tag of body
 * uid i+1 (the smallest body index of its island), -1 for static bodies; n = number of bodies to report. */
int32_t b2c_compute_islands(b2c_ctx*, int32_t* tags_out, int32_t n, int32_t* num_islands_out);

/* CollisionWorld.rayTest with a ClosestRayResultCallback per ray (disp/CollisionWorld.java:553-590, 697-729), batched:
 * from_xyz / to_xyz = 3 floats per ray (host); group / mask = the callback's collisionFilterGroup / collisionFilterMask
 * (disp/CollisionWorld.java:655-670, defaults DEFAULT_FILTER = 1 and ALL_FILTER = -1).  Outputs per ray: uid of the closest
 * body hit (0 = none), closestHitFraction (1 = none), hitNormalWorld, hitPointWorld.  Uses the transforms currently
 * resident.  Every branch of rayTestSingle (disp/CollisionWorld.java:260-356) is covered: convex bodies (box, sphere, hull)
 * are cast with the reference's SubsimplexConvexCast, triangle meshes with the quantised-BVH ray walk
 * (sh/OptimizedBvh.java:817-931) + np/TriangleRaycastCallback.java:46-117, static planes with the two triangles
 * sh/StaticPlaneShape.java:60-122 generates, compounds child by child.  For meshes and planes the normal is the reference's
 * unnormalised triangle normal rotated into world space.  No limit on the number of bodies a ray may meet: beyond 1024
 * candidate boxes a ray is evaluated in body-index tiles (exact, but without spatial culling). */
int32_t b2c_ray_test_closest(b2c_ctx*, int32_t n, const float* from_xyz, const float* to_xyz, int16_t group, int16_t mask,
                             int32_t* uid_out, float* fraction_out, float* normal_out, float* point_out);

/* CollisionWorld.convexSweepTest with a ClosestConvexResultCallback per sweep (disp/CollisionWorld.java:596-651, 765-800),
 * batched, for TRANSLATIONAL sweeps: sweep i moves the registered convex shape cast_shape_ids[i] (box, sphere or hull) with
 * the fixed basis basis9[9*i..] (row-major) from from_xyz[3*i..] to to_xyz[3*i..].  group / mask = the callback's
 * collisionFilterGroup / collisionFilterMask (:752-756); allowed_ccd_penetration = DispatcherInfo.allowedCcdPenetration
 * (bp/DispatcherInfo.java:44, reference default 0.04).  Outputs per sweep: uid of hitCollisionObject (0 = none),
 * closestHitFraction (1 = none), hitNormalWorld, hitPointWorld.  Per target objectQuerySingle (:392-551): convex bodies by
 * np/GjkConvexCast.java:66-196, triangle meshes by the BVH box-cast walk (sh/OptimizedBvh.java:1017-1036) +
 * np/TriangleConvexcastCallback.java:53-88 (SubsimplexConvexCast per triangle), compounds child by child.  The reference's
 * static-plane branch dereferences a null caster (:470-477) and throws: a sweep whose expanded segment meets a static plane
 * that passes the filter reports uid -1 (exclude planes with `mask`).  The cast shape's culling box is
 * calculateTemporalAabb with zero angular velocity (what equal bases mean up to the reference's quaternion rounding).
 * Uses the transforms currently resident.  A sweep may have any number of candidate bodies (the reference's culling box
 * includes the whole motion, so long sweeps through dense scenes have thousands): they are taken in rounds on the device. */
int32_t b2c_convex_sweep_closest(b2c_ctx*, int32_t n, const int32_t* cast_shape_ids, const float* basis9, const float* from_xyz,
                                 const float* to_xyz, int16_t group, int16_t mask, float allowed_ccd_penetration, int32_t* uid_out,
                                 float* fraction_out, float* normal_out, float* point_out);

/* The "CCD motion clamping" query of DiscreteDynamicsWorld.integrateTransforms (dyn/DiscreteDynamicsWorld.java:700-729),
 * batched: for body_uids[i] a SphereShape(ccd_radius[i]) (the body's ccdSweptSphereRadius) is swept from the body's RESIDENT
 * world transform to predicted_xyz[i] (predictedTrans.origin) with a ClosestNotMeConvexResultCallback (:1129-1199) whose
 * filter group / mask are the body's own: the body itself is skipped, so is every object it already has contact points
 * with (any manifold of their pair's algorithm in the pair cache of the last step, :1181-1196), and a result whose normal
 * does not oppose the motion (:1157).  Outputs as b2c_convex_sweep_closest; the host applies the reference's clamp
 * `hasHit && closestHitFraction > 0.0001` (:721) and re-integrates with timeStep * fraction.  The caller lists the bodies
 * whose squared motion exceeds their ccdSquareMotionThreshold and whose shape is convex (:700-703).  The predicted ROTATION
 * enters the reference only through the angular term of the swept sphere's culling box (|w| r sqrt 3), taken as 0 here. */
int32_t b2c_ccd_sweep_not_me(b2c_ctx*, int32_t n, const int32_t* body_uids, const float* ccd_radius, const float* predicted_xyz,
                             float allowed_ccd_penetration, int32_t* hit_uid_out, float* fraction_out, float* normal_out,
                             float* point_out);

/* ---- one world partitioned over several GPUs -------------------------------------
 * Space is cut into `nranks` slabs by planes along one axis; rank r owns slab r.  A PROXY is owned by the rank whose slab
 * held its origin when the partition was set: only that rank runs updateAabbs / setAabb for it.  A PAIR is owned by the slab
 * that holds max(min_a, min_b) along the axis (a coordinate inside both boxes).  Each step every rank needs the boxes of all
 * proxies that touch its slab: its own plus the HALO — proxies owned elsewhere whose box reaches into the slab.  Owners publish
 * their boundary proxies (box not entirely inside the home slab) as 80-byte records (box, proxy index, flags, transform) in
 * a fixed-size slot, ONE all-gather over NVLink (e.g. ncclAllGather on the ctx stream) hands every rank all slots, and each
 * rank sorts and sweeps its local list.  The union of the ranks' pair lists is the single-GPU list.  A pair near a plane can
 * change owner between steps, so its manifold migrates with a second, small all-gather:
 *   b2c_mgpu_update_export_halo -> [all-gather halo slots] -> b2c_mgpu_import_halo -> b2c_mgpu_broadphase ->
 *   b2c_mgpu_export_departed_slot -> [all-gather migration slots] -> b2c_mgpu_import_arrival_slots -> b2c_mgpu_narrowphase.
 * (or, with the halo exchange as peer-to-peer stores — no collective, see b2c_mgpu_p2p_*:
 *   b2c_mgpu_p2p_export_halo -> b2c_mgpu_p2p_import_halo -> b2c_mgpu_broadphase ->
 *   b2c_mgpu_p2p_export_departed -> b2c_mgpu_p2p_import_arrivals -> b2c_mgpu_narrowphase)
 * Every call only enqueues on the ctx stream (no host synchronisation); overflowing slots are reported by the next
 * b2c_sync_counts as B2C_ERR_CAPACITY.  All ranks create the same proxies in the same order (static attributes are
 * replicated); per-step transforms are only needed for the proxies a rank owns (b2c_get_partition tells which).
 * Slot buffers are DEVICE pointers owned by the caller, so NCCL can use them directly. */
/* Slabs along `axis` (0/1/2) cut at planes[0 .. nranks-2] (ascending).  Call on every rank after the proxies exist and while
 * all ranks hold the same transforms; nranks = 1 removes the partition.  At most 16 ranks. */
int32_t b2c_set_partition_slabs(b2c_ctx*, int32_t rank, int32_t nranks, int32_t axis, const float* planes);
/* The same with planes chosen by the library: the axis the origins spread most over, cut at equal-count quantiles. */
int32_t b2c_set_partition(b2c_ctx*, int32_t rank, int32_t nranks);
/* The partition in force: axis, nranks-1 planes, and the owning rank of proxies 1..n (any output may be NULL). */
int32_t b2c_get_partition(b2c_ctx*, int32_t* axis_out, float* planes_out, uint8_t* owner_out, int32_t n);
/* Halo slot = { uint32 count, uint32 pad[3], 80-byte records[cap] }. */
int64_t b2c_mgpu_halo_slot_bytes(int32_t cap);
/* CollisionWorld.updateAabbs for the proxies this rank owns (disp/CollisionWorld.java:231-245), then its boundary proxies
 * into `slot_dev`. */
int32_t b2c_mgpu_update_export_halo(b2c_ctx*, void* slot_dev, int32_t cap);
/* After the all-gather: `slots_dev` = nranks slots in rank order.  Adopts the records that touch this rank's slab. */
int32_t b2c_mgpu_import_halo(b2c_ctx*, const void* slots_dev, int32_t nslots, int32_t cap);
/* The same exchange WITHOUT a collective, fused into the export kernel: every rank owns an inbox (2 step parities x nranks
 * slots) in its own HBM; the export kernel stores each boundary record straight into the inbox of exactly the ranks whose slab
 * the box reaches (peer memory over NVLink / NVSwitch), then publishes {count, epoch} per destination with one release store
 * at system scope; the import waits (acquire, system scope) for this step's epoch from every source and adopts the records.
 * A rank receives only what touches it and nothing is padded to a slot size; no NCCL call sits between the two kernels.
 *   b2c_mgpu_p2p_init    allocates this rank's inboxes (halo: `cap` records per source; manifold migration: `migrate_cap`
 *                        manifolds per source; all ranks must pass the same capacities) in ONE allocation; returns its 64-byte
 *                        CUDA IPC handle (cudaIpcMemHandle_t) and / or its device pointer (either output may be NULL)
 *   b2c_mgpu_p2p_connect maps the peers' inboxes: ipc_handles = nranks x 64 bytes in rank order (ranks in other processes on
 *                        the same node; exchange them once, e.g. with an all-gather), or inbox_ptrs = nranks device pointers
 *                        (ranks that share this process).  The own entry is ignored.
 *   b2c_mgpu_p2p_export_halo = updateAabbs of the owned proxies + the stores into the peers (replaces update_export_halo +
 *                        the all-gather);  b2c_mgpu_p2p_import_halo = wait + adoption (replaces import_halo).
 *   b2c_mgpu_p2p_export_departed / b2c_mgpu_p2p_import_arrivals: the manifolds of pairs that changed owner, pushed to every
 *                        other rank the same way (replace export_departed_slot + all-gather + import_arrival_slots).
 * With both, a partitioned step contains no collective at all — only this library's kernels over peer memory.
 * A source that never publishes (a dead peer) ends the wait after ~10 s; b2c_sync_counts then returns B2C_ERR_STATE. */
int32_t b2c_mgpu_p2p_init(b2c_ctx*, int32_t cap, int32_t migrate_cap, void* ipc_handle_out /* 64 bytes */, void** inbox_dev_out);
int32_t b2c_mgpu_p2p_connect(b2c_ctx*, const void* ipc_handles, void* const* inbox_ptrs);
int32_t b2c_mgpu_p2p_export_halo(b2c_ctx*);
int32_t b2c_mgpu_p2p_import_halo(b2c_ctx*);
int32_t b2c_mgpu_p2p_export_departed(b2c_ctx*);
int32_t b2c_mgpu_p2p_import_arrivals(b2c_ctx*);
/* BroadphaseInterface.calculateOverlappingPairs over the rank's local list; keeps the pairs this rank owns. */
int32_t b2c_mgpu_broadphase(b2c_ctx*);
/* Manifold migration with a host round trip (tests): keys: uint64[cap]; headers: 32-byte records [cap]; points:
 * b2c_manifold_point[4*cap]. */
int32_t b2c_mgpu_export_departed(b2c_ctx*, uint64_t* keys_dev, void* headers_dev, b2c_manifold_point* points_dev, int32_t cap,
                                 int32_t* count_out);
int32_t b2c_mgpu_import_arrivals(b2c_ctx*, const uint64_t* keys_dev, const void* headers_dev, const b2c_manifold_point* points_dev,
                                 int32_t count);
int32_t b2c_mgpu_narrowphase(b2c_ctx*);
/* The same exchange without any host synchronisation, for a fixed-size all-gather: every rank packs its departed manifolds
 * into one SLOT { uint32 count, uint32 pad[3], uint64 keys[cap], 32-byte headers[cap], b2c_manifold_point points[4*cap] }
 * of b2c_mgpu_slot_bytes(cap) bytes; after the all-gather each rank scans all `nslots` slots on the device. */
int64_t b2c_mgpu_slot_bytes(int32_t cap);
int32_t b2c_mgpu_export_departed_slot(b2c_ctx*, void* slot_dev, int32_t cap);
int32_t b2c_mgpu_import_arrival_slots(b2c_ctx*, const void* slots_dev, int32_t nslots, int32_t cap);

#ifdef __cplusplus
}
#endif
#endif /* B2C_H */
