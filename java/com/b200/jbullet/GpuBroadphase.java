package com.b200.jbullet;

import java.lang.foreign.Arena;
import java.lang.foreign.MemorySegment;

import com.badlogic.gdx.math.Vector3;
import com.bulletphysics.collision.broadphase.*;
import com.bulletphysics.util.ObjectArrayList;

import static java.lang.foreign.ValueLayout.*;

/**
 * Drop-in for {@code DbvtBroadphase} behind {@link BroadphaseInterface} (bp/BroadphaseInterface.java:33-52).
 * NOT COMPILED IN THIS REPOSITORY'S IMAGE (no JDK / libgdx jar).  Host state lives in off-heap SoA
 * MemorySegments: per-step setAabb calls are gathered into six float planes and flushed with ONE
 * b2c_set_aabbs + b2c_calculate_overlapping_pairs; the pair list comes back with one b2c_get_pairs.
 */
public class GpuBroadphase extends BroadphaseInterface {
    final Arena arena = Arena.ofConfined();   // one ctx per world per thread, like the reference's thread-local pools
    final MemorySegment ctx;
    final int maxBodies;
    final MemorySegment aabbPlanes;           // 6 planes x maxBodies floats (minx miny minz maxx maxy maxz)
    final MemorySegment uids;                 // uids touched this step
    int touched = 0;
    final MemorySegment pairBuf;              // 2 x int32 per pair
    final MemorySegment addedBuf, removedBuf; // the step's pair-cache events (b2c_get_pair_deltas), 2 x int32 per pair
    final MemorySegment scratchInt;
    final ObjectArrayList<GpuProxy> proxies = new ObjectArrayList<GpuProxy>();
    final GpuPairCache pairCache = new GpuPairCache(this);

    public static final class GpuProxy extends BroadphaseProxy {
        int uid;
        GpuProxy(Object client, short group, short mask) { super(client, group, mask); }
        @Override public int getUid() { return uid; }
    }

    public GpuBroadphase(MemorySegment ctx, int maxBodies, int maxPairs) {
        this.ctx = ctx;
        this.maxBodies = maxBodies;
        this.aabbPlanes = arena.allocate(JAVA_FLOAT, 6L * maxBodies);
        this.uids = arena.allocate(JAVA_INT, maxBodies);
        this.pairBuf = arena.allocate(JAVA_INT, 2L * maxPairs);
        this.addedBuf = arena.allocate(JAVA_INT, 2L * maxPairs);
        this.removedBuf = arena.allocate(JAVA_INT, 2L * maxPairs);
        this.scratchInt = arena.allocate(JAVA_INT, 4);
        try {
            B2C.check((int) B2C.setPairDeltaPrefetch.invokeExact(ctx, 1), ctx);   // the deltas come out of the pair calculation itself
        } catch (Throwable t) { throw new RuntimeException(t); }
    }

    /** bp/BroadphaseInterface.java:35, called by the reference's own CollisionWorld.addCollisionObject
     *  (disp/CollisionWorld.java:102-121) with {@code userPtr} = the CollisionObject: its shape is registered on first use and
     *  the proxy created from shape + world transform by GpuShapes.createProxyFor -> b2c_proxy_create (uid = ++gid like
     *  bp/DbvtBroadphase.java:179; the device recomputes the identical AABB the caller passes in). */
    @Override
    public BroadphaseProxy createProxy(Vector3 aabbMin, Vector3 aabbMax, BroadphaseNativeType shapeType, Object userPtr,
                                       short group, short mask, Dispatcher dispatcher, Object multiSapProxy) {
        GpuProxy p = new GpuProxy(userPtr, group, mask);
        p.uid = GpuShapes.createProxyFor(ctx, userPtr, group, mask, scratchInt);   // -> b2c_proxy_create
        proxies.add(p);
        return p;
    }

    @Override
    public void destroyProxy(BroadphaseProxy proxy, Dispatcher dispatcher) {
        try {
            B2C.check((int) B2C.proxyDestroy.invokeExact(ctx, ((GpuProxy) proxy).uid), ctx);
        } catch (Throwable t) { throw new RuntimeException(t); }
    }

    /** bp/BroadphaseInterface.java:39 — batched; flushed in calculateOverlappingPairs. */
    @Override
    public void setAabb(BroadphaseProxy proxy, Vector3 mn, Vector3 mx, Dispatcher dispatcher) {
        int k = touched++;
        uids.setAtIndex(JAVA_INT, k, ((GpuProxy) proxy).uid);
        // planes have stride `touched` at flush time; stage row-wise here and transpose in flush()
        long o = 6L * k;
        aabbPlanes.setAtIndex(JAVA_FLOAT, o, mn.x); aabbPlanes.setAtIndex(JAVA_FLOAT, o + 1, mn.y); aabbPlanes.setAtIndex(JAVA_FLOAT, o + 2, mn.z);
        aabbPlanes.setAtIndex(JAVA_FLOAT, o + 3, mx.x); aabbPlanes.setAtIndex(JAVA_FLOAT, o + 4, mx.y); aabbPlanes.setAtIndex(JAVA_FLOAT, o + 5, mx.z);
    }

    /** bp/BroadphaseInterface.java:42 */
    @Override
    public void calculateOverlappingPairs(Dispatcher dispatcher) {
        try (Arena a = Arena.ofConfined()) {
            if (touched > 0) {
                MemorySegment planes = a.allocate(JAVA_FLOAT, 6L * touched);
                for (int k = 0; k < touched; k++)
                    for (int c = 0; c < 6; c++)
                        planes.setAtIndex(JAVA_FLOAT, (long) c * touched + k, aabbPlanes.getAtIndex(JAVA_FLOAT, 6L * k + c));
                B2C.check((int) B2C.setAabbs.invokeExact(ctx, touched, uids, planes), ctx);
                touched = 0;
            }
            B2C.check((int) B2C.calculateOverlappingPairs.invokeExact(ctx, scratchInt), ctx);
            int n = scratchInt.get(JAVA_INT, 0);
            B2C.check((int) B2C.getPairs.invokeExact(ctx, pairBuf, (int) (pairBuf.byteSize() / 8), scratchInt), ctx);
            int cap = (int) (addedBuf.byteSize() / 8);
            B2C.check((int) B2C.getPairDeltas.invokeExact(ctx, addedBuf, cap, removedBuf, cap, scratchInt.asSlice(4), scratchInt.asSlice(8)), ctx);
            // keeps the BroadphasePair objects of surviving pairs, creates / drops the others, fires the ghost add / remove events
            pairCache.refresh(pairBuf, n, addedBuf, scratchInt.get(JAVA_INT, 4), removedBuf, scratchInt.get(JAVA_INT, 8), proxies, dispatcher);
        } catch (Throwable t) { throw new RuntimeException(t); }
    }

    @Override public OverlappingPairCache getOverlappingPairCache() { return pairCache; }
    /** bp/BroadphaseInterface.java:48: unbounded for Dbvt / Simple, the world box for the AxisSweep3 modes. */
    @Override public void getBroadphaseAabb(Vector3 mn, Vector3 mx) {
        try (Arena a = Arena.ofConfined()) {
            MemorySegment lo = a.allocate(JAVA_FLOAT, 3), hi = a.allocate(JAVA_FLOAT, 3);
            B2C.check((int) B2C.getBroadphaseAabb.invokeExact(ctx, lo, hi), ctx);
            mn.set(lo.getAtIndex(JAVA_FLOAT, 0), lo.getAtIndex(JAVA_FLOAT, 1), lo.getAtIndex(JAVA_FLOAT, 2));
            mx.set(hi.getAtIndex(JAVA_FLOAT, 0), hi.getAtIndex(JAVA_FLOAT, 1), hi.getAtIndex(JAVA_FLOAT, 2));
        } catch (Throwable t) { throw new RuntimeException(t); }
    }
    @Override public void printStats() { }
}
