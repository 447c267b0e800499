/* b2c_jni.c — JNI forwarders for com.b200.jbullet.B2CJni (the fallback binding for JDKs without java.lang.foreign).
 *
 * The C ABI (include/b2c.h) already takes plain pointers and sizes, so every function here only unwraps its arguments
 * (direct ByteBuffers -> addresses, long -> b2c_ctx*) and forwards.  Build on a box with a JDK:
 *   gcc -shared -fPIC -I$JAVA_HOME/include -I$JAVA_HOME/include/linux -Iinclude java/jni/b2c_jni.c \
 *       -Llibgdx-jbullet_b200 -lb2c -o libb2cjni.so
 * This image has no JDK; tests/test_abi.py compiles the file against tests/jni_stub/jni.h (declarations only).
 */
#include <jni.h>
#include <stdint.h>
#include <stdlib.h>

#include "b2c.h"

#define CTX(h) ((b2c_ctx*)(intptr_t)(h))
#define FN(name) JNIEXPORT JNICALL Java_com_b200_jbullet_B2CJni_##name

static void* addr(JNIEnv* env, jobject buf) { return buf ? (*env)->GetDirectBufferAddress(env, buf) : NULL; }
static void put_int(JNIEnv* env, jintArray arr, int at, int32_t v) {
    if (arr) { jint j = (jint)v; (*env)->SetIntArrayRegion(env, arr, at, 1, &j); }
}

void FN(defaultConfig)(JNIEnv* env, jclass c, jobject cfg) { (void)c; b2c_default_config((b2c_config*)addr(env, cfg)); }

jint FN(create)(JNIEnv* env, jclass c, jobject cfg, jlongArray out) {
    (void)c;
    b2c_ctx* ctx = NULL;
    int32_t rc = b2c_create((const b2c_config*)addr(env, cfg), &ctx);
    jlong h = (jlong)(intptr_t)ctx;
    (*env)->SetLongArrayRegion(env, out, 0, 1, &h);
    return rc;
}
void FN(destroy)(JNIEnv* env, jclass c, jlong h) { (void)env; (void)c; b2c_destroy(CTX(h)); }
jstring FN(lastError)(JNIEnv* env, jclass c, jlong h) { (void)c; return (*env)->NewStringUTF(env, b2c_last_error_string(CTX(h))); }

jint FN(shapeBox)(JNIEnv* env, jclass c, jlong h, jfloat hx, jfloat hy, jfloat hz, jfloat margin, jintArray out) {
    (void)c;
    const float he[3] = {hx, hy, hz};
    int32_t id = -1, rc = b2c_shape_register_box(CTX(h), he, margin, &id);
    put_int(env, out, 0, id);
    return rc;
}
jint FN(shapeSphere)(JNIEnv* env, jclass c, jlong h, jfloat r, jintArray out) {
    (void)c;
    int32_t id = -1, rc = b2c_shape_register_sphere(CTX(h), r, &id);
    put_int(env, out, 0, id);
    return rc;
}
jint FN(shapeHull)(JNIEnv* env, jclass c, jlong h, jobject pts, jint n, jfloat margin, jintArray out) {
    (void)c;
    int32_t id = -1, rc = b2c_shape_register_hull(CTX(h), (const float*)addr(env, pts), n, margin, &id);
    put_int(env, out, 0, id);
    return rc;
}
jint FN(shapePlane)(JNIEnv* env, jclass c, jlong h, jfloat nx, jfloat ny, jfloat nz, jfloat k, jintArray out) {
    (void)c;
    const float n[3] = {nx, ny, nz};
    int32_t id = -1, rc = b2c_shape_register_plane(CTX(h), n, k, &id);
    put_int(env, out, 0, id);
    return rc;
}
jint FN(shapeMeshParts)(JNIEnv* env, jclass c, jlong h, jobjectArray vbase, jintArray nv, jintArray vstride, jobjectArray ibase,
                        jintArray nt, jintArray istride, jintArray itype, jfloat sx, jfloat sy, jfloat sz, jintArray out) {
    (void)c;
    const jsize parts = (*env)->GetArrayLength(env, vbase);
    b2c_indexed_mesh* m = (b2c_indexed_mesh*)calloc((size_t)(parts > 0 ? parts : 1), sizeof(b2c_indexed_mesh));
    if (!m) return B2C_ERR_BAD_ARG;
    for (jsize p = 0; p < parts; p++) {
        jint v;
        m[p].vertex_base = addr(env, (*env)->GetObjectArrayElement(env, vbase, p));
        m[p].index_base = addr(env, (*env)->GetObjectArrayElement(env, ibase, p));
        (*env)->GetIntArrayRegion(env, nv, p, 1, &v); m[p].num_vertices = v;
        (*env)->GetIntArrayRegion(env, vstride, p, 1, &v); m[p].vertex_stride = v;
        (*env)->GetIntArrayRegion(env, nt, p, 1, &v); m[p].num_triangles = v;
        (*env)->GetIntArrayRegion(env, istride, p, 1, &v); m[p].index_stride = v;
        (*env)->GetIntArrayRegion(env, itype, p, 1, &v); m[p].index_type = v;
    }
    const float s[3] = {sx, sy, sz};
    int32_t id = -1, rc = b2c_shape_register_mesh_parts(CTX(h), m, (int32_t)parts, s, &id);
    free(m);
    put_int(env, out, 0, id);
    return rc;
}
jint FN(shapeCompound)(JNIEnv* env, jclass c, jlong h, jint n, jobject shapes, jobject xf, jintArray out) {
    (void)c;
    int32_t id = -1, rc = b2c_shape_register_compound(CTX(h), n, (const int32_t*)addr(env, shapes), (const float*)addr(env, xf), &id);
    put_int(env, out, 0, id);
    return rc;
}
jint FN(proxyCreate)(JNIEnv* env, jclass c, jlong h, jint shape, jobject xf, jshort group, jshort mask, jint flags, jint world,
                     jintArray out) {
    (void)c;
    int32_t uid = 0, rc = b2c_proxy_create(CTX(h), shape, (const float*)addr(env, xf), group, mask, flags, world, &uid);
    put_int(env, out, 0, uid);
    return rc;
}
jint FN(proxyDestroy)(JNIEnv* env, jclass c, jlong h, jint uid) { (void)env; (void)c; return b2c_proxy_destroy(CTX(h), uid); }
jint FN(proxySetMaterial)(JNIEnv* env, jclass c, jlong h, jint uid, jfloat f, jfloat r) {
    (void)env; (void)c;
    return b2c_proxy_set_material(CTX(h), uid, f, r);
}
jint FN(setTransforms)(JNIEnv* env, jclass c, jlong h, jint n, jobject uids, jobject planes) {
    (void)c;
    return b2c_set_transforms(CTX(h), n, (const int32_t*)addr(env, uids), (const float*)addr(env, planes));
}
jint FN(setActivation)(JNIEnv* env, jclass c, jlong h, jint n, jobject uids, jobject active) {
    (void)c;
    return b2c_set_activation(CTX(h), n, (const int32_t*)addr(env, uids), (const uint8_t*)addr(env, active));
}
jint FN(setNoCollidePairs)(JNIEnv* env, jclass c, jlong h, jint n, jobject pairs) {
    (void)c;
    return b2c_set_no_collide_pairs(CTX(h), n, (const int32_t*)addr(env, pairs));
}
jint FN(setAabbs)(JNIEnv* env, jclass c, jlong h, jint n, jobject uids, jobject mm) {
    (void)c;
    return b2c_set_aabbs(CTX(h), n, (const int32_t*)addr(env, uids), (const float*)addr(env, mm));
}
jint FN(updateAabbs)(JNIEnv* env, jclass c, jlong h) { (void)env; (void)c; return b2c_update_aabbs(CTX(h)); }
jint FN(calculateOverlappingPairs)(JNIEnv* env, jclass c, jlong h, jintArray out) {
    (void)c;
    int32_t n = 0, rc = b2c_calculate_overlapping_pairs(CTX(h), &n);
    put_int(env, out, 0, n);
    return rc;
}
jint FN(getPairs)(JNIEnv* env, jclass c, jlong h, jobject pairs, jint cap, jintArray out) {
    (void)c;
    int32_t n = 0, rc = b2c_get_pairs(CTX(h), (int32_t*)addr(env, pairs), cap, &n);
    put_int(env, out, 0, n);
    return rc;
}
jint FN(getPairDeltas)(JNIEnv* env, jclass c, jlong h, jobject added, jint capA, jobject removed, jint capR, jintArray counts) {
    (void)c;
    int32_t na = 0, nr = 0;
    int32_t rc = b2c_get_pair_deltas(CTX(h), (int32_t*)addr(env, added), capA, (int32_t*)addr(env, removed), capR, &na, &nr);
    put_int(env, counts, 0, na);
    put_int(env, counts, 1, nr);
    return rc;
}
jint FN(dispatchAllPairs)(JNIEnv* env, jclass c, jlong h, jintArray out) {
    (void)c;
    int32_t m = 0, k = 0, rc = b2c_dispatch_all_pairs(CTX(h), &m, &k);
    put_int(env, out, 0, m);
    put_int(env, out, 1, k);
    return rc;
}
jint FN(getContacts)(JNIEnv* env, jclass c, jlong h, jobject hdr, jint capH, jobject pts, jint capP, jintArray counts) {
    (void)c;
    int32_t nh = 0, np = 0;
    int32_t rc = b2c_get_contacts(CTX(h), (b2c_contact_header*)addr(env, hdr), capH, (b2c_manifold_point*)addr(env, pts), capP, &nh, &np);
    put_int(env, counts, 0, nh);
    put_int(env, counts, 1, np);
    return rc;
}
jint FN(step)(JNIEnv* env, jclass c, jlong h, jint n, jobject planes, jintArray counts) {
    (void)c;
    int32_t p = 0, m = 0, k = 0, rc = b2c_step(CTX(h), n, (const float*)addr(env, planes), &p, &m, &k);
    put_int(env, counts, 0, p);
    put_int(env, counts, 1, m);
    put_int(env, counts, 2, k);
    return rc;
}
jint FN(getBroadphaseAabb)(JNIEnv* env, jclass c, jlong h, jfloatArray mn, jfloatArray mx) {
    (void)c;
    float a[3], b[3];
    int32_t rc = b2c_get_broadphase_aabb(CTX(h), a, b);
    (*env)->SetFloatArrayRegion(env, mn, 0, 3, a);
    (*env)->SetFloatArrayRegion(env, mx, 0, 3, b);
    return rc;
}
jint FN(setPairDeltaPrefetch)(JNIEnv* env, jclass c, jlong h, jint on) { (void)env; (void)c; return b2c_set_pair_delta_prefetch(CTX(h), on); }
jint FN(computeIslands)(JNIEnv* env, jclass c, jlong h, jobject tags, jint n, jintArray out) {
    (void)c;
    int32_t k = 0, rc = b2c_compute_islands(CTX(h), (int32_t*)addr(env, tags), n, &k);
    put_int(env, out, 0, k);
    return rc;
}
