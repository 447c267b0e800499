package com.b200.jbullet;

import java.nio.ByteBuffer;

/**
 * JNI fallback for JDKs without java.lang.foreign (the reference targets Java 6-8): the same exported symbols of libb2c.so
 * behind {@code native} methods; java/jni/b2c_jni.c holds the one-line forwarders (libb2cjni.so, linked against libb2c.so).
 * NOT COMPILED IN THIS REPOSITORY'S IMAGE (no JDK); the C side is compile-checked against a stub jni.h by
 * tests/test_abi.py::test_jni_forwarders_compile.
 *
 * Every buffer is a DIRECT ByteBuffer in native byte order (off-heap SoA storage, like the FFM MemorySegments of the
 * primary binding); {@code ctx} is the b2c_ctx* as a long.  Return values are b2c_status codes (0 = ok).
 */
final class B2CJni {
    static { System.loadLibrary("b2cjni"); }

    static native void defaultConfig(ByteBuffer config64);                                         // b2c_default_config
    static native int create(ByteBuffer config64, long[] ctxOut);                                  // b2c_create
    static native void destroy(long ctx);                                                          // b2c_destroy
    static native String lastError(long ctx);                                                      // b2c_last_error_string
    static native int shapeBox(long ctx, float hx, float hy, float hz, float marginOrNeg, int[] shapeOut);
    static native int shapeSphere(long ctx, float radius, int[] shapeOut);
    static native int shapeHull(long ctx, ByteBuffer pointsXyz, int numPoints, float marginOrNeg, int[] shapeOut);
    static native int shapePlane(long ctx, float nx, float ny, float nz, float constant, int[] shapeOut);
    static native int shapeMeshParts(long ctx, ByteBuffer[] vertexBase, int[] numVertices, int[] vertexStride, ByteBuffer[] indexBase,
                                     int[] numTriangles, int[] indexStride, int[] indexType, float sx, float sy, float sz, int[] shapeOut);
    static native int shapeCompound(long ctx, int numChildren, ByteBuffer childShapes, ByteBuffer childTransforms12, int[] shapeOut);
    static native int proxyCreate(long ctx, int shape, ByteBuffer transform12, short group, short mask, int flags, int world, int[] uidOut);
    static native int proxyDestroy(long ctx, int uid);
    static native int proxySetMaterial(long ctx, int uid, float friction, float restitution);
    static native int setTransforms(long ctx, int n, ByteBuffer uidsOrNull, ByteBuffer planes12);
    static native int setActivation(long ctx, int n, ByteBuffer uidsOrNull, ByteBuffer active);
    static native int setNoCollidePairs(long ctx, int n, ByteBuffer uidPairs);
    static native int setAabbs(long ctx, int n, ByteBuffer uidsOrNull, ByteBuffer minmax6);
    static native int updateAabbs(long ctx);
    static native int calculateOverlappingPairs(long ctx, int[] numPairsOut);
    static native int getPairs(long ctx, ByteBuffer pairsOut, int capPairs, int[] numPairsOut);
    static native int getPairDeltas(long ctx, ByteBuffer addedOut, int capAdded, ByteBuffer removedOut, int capRemoved, int[] counts2);
    static native int dispatchAllPairs(long ctx, int[] manifoldsAndContactsOut2);
    static native int getContacts(long ctx, ByteBuffer headersOut, int capHeaders, ByteBuffer pointsOut, int capPoints, int[] counts2);
    static native int step(long ctx, int n, ByteBuffer planes12OrNull, int[] counts3);
    static native int getBroadphaseAabb(long ctx, float[] min3, float[] max3);
    static native int setPairDeltaPrefetch(long ctx, int on);
    static native int computeIslands(long ctx, ByteBuffer tagsOut, int n, int[] numIslandsOut);

    private B2CJni() { }
}
