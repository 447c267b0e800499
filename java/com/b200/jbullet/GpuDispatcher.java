package com.b200.jbullet;

import java.lang.foreign.Arena;
import java.lang.foreign.MemorySegment;

import com.bulletphysics.collision.broadphase.*;
import com.bulletphysics.collision.dispatch.CollisionObject;
import com.bulletphysics.collision.narrowphase.ManifoldPoint;
import com.bulletphysics.collision.narrowphase.PersistentManifold;
import com.bulletphysics.util.ObjectArrayList;

import static java.lang.foreign.ValueLayout.*;

/**
 * Drop-in for {@code CollisionDispatcher} behind {@link Dispatcher} (bp/Dispatcher.java:38-68).
 * NOT COMPILED IN THIS REPOSITORY'S IMAGE.  dispatchAllCollisionPairs runs the whole narrowphase on the device
 * (b2c_dispatch_all_pairs) and pulls back only the touching manifolds (b2c_get_contacts; a solver-only host would take
 * b2c_get_packed_contacts: 16-byte headers, 48-byte points); the Java
 * PersistentManifold / ManifoldPoint objects the island manager and the solver read are refreshed from that
 * stream.  {@code src_slot} says which slot of the same manifold a point continues, so the solver's warm-start
 * fields (appliedImpulse, appliedImpulseLateral1/2, lateralFrictionInitialized, userPersistentData) stay attached
 * exactly as np/PersistentManifold.java:280-305 (replaceContactPoint) and :259-278 (removeContactPoint) keep them.
 */
public class GpuDispatcher extends Dispatcher {
    final MemorySegment ctx;
    final Arena arena = Arena.ofConfined();
    final MemorySegment headers;   // b2c_contact_header[maxPairs]  (32 B)
    final MemorySegment points;    // b2c_manifold_point[4*maxPairs] (96 B)
    final MemorySegment out = arena.allocate(JAVA_INT, 4);
    final ObjectArrayList<PersistentManifold> manifolds = new ObjectArrayList<PersistentManifold>();
    /** (pair, child code): a plain pair owns one manifold, a compound pair one per child algorithm. */
    record ManifoldKey(int uid0, int uid1, int childCode) { }
    final java.util.HashMap<ManifoldKey, PersistentManifold> byPair = new java.util.HashMap<ManifoldKey, PersistentManifold>();
    final GpuBroadphase broadphase;

    public GpuDispatcher(MemorySegment ctx, GpuBroadphase bp, int maxPairs) {
        this.ctx = ctx;
        this.broadphase = bp;
        this.headers = arena.allocate(32L * maxPairs);
        this.points = arena.allocate(96L * 4 * maxPairs);   // a manifold holds up to 4 points
    }

    /** bp/Dispatcher.java:58 */
    @Override
    public void dispatchAllCollisionPairs(OverlappingPairCache pairCache, DispatcherInfo info, Dispatcher dispatcher) {
        try {
            B2C.check((int) B2C.dispatchAllPairs.invokeExact(ctx, out, out.asSlice(4)), ctx);
            B2C.check((int) B2C.getContacts.invokeExact(ctx, headers, (int) (headers.byteSize() / 32), points,
                                                        (int) (points.byteSize() / 96), out.asSlice(8), out.asSlice(12)), ctx);
        } catch (Throwable t) { throw new RuntimeException(t); }
        int nh = out.get(JAVA_INT, 8);
        manifolds.clear();
        java.util.HashMap<ManifoldKey, PersistentManifold> next = new java.util.HashMap<ManifoldKey, PersistentManifold>();
        for (int h = 0; h < nh; h++) {
            long o = 32L * h;
            int uid0 = headers.get(JAVA_INT, o), uid1 = headers.get(JAVA_INT, o + 4);
            int body0 = headers.get(JAVA_INT, o + 8), body1 = headers.get(JAVA_INT, o + 12);
            int n = headers.get(JAVA_INT, o + 16), algorithm = headers.get(JAVA_INT, o + 20), first = headers.get(JAVA_INT, o + 24);
            // pair_index < 0: child manifold of a compound pair (disp/CompoundCollisionAlgorithm.java:49-75 keeps one child
            // algorithm, hence one manifold, per child): v = -1 - pair_index carries (child0 + 1) | (child1 + 1) << 15
            int pairIndex = headers.get(JAVA_INT, o + 28);
            ManifoldKey key = new ManifoldKey(uid0, uid1, pairIndex < 0 ? -1 - pairIndex : 0);   // 0 for every plain pair
            PersistentManifold m = byPair.get(key);
            GpuManifolds.SolverState[] old = null;
            if (m == null) {
                m = new PersistentManifold();
                m.init(broadphase.proxies.getQuick(body0 - 1).clientObject, broadphase.proxies.getQuick(body1 - 1).clientObject, 0);
            } else {
                old = GpuManifolds.snapshotSolverState(m);       // copies of the 4 points' warm-start fields
            }
            GpuManifolds.fill(m, points, first, n, old, algorithm);   // geometry from the device, solver state via src_slot
            next.put(key, m);
            manifolds.add(m);
        }
        byPair.clear();
        byPair.putAll(next);
    }

    @Override public int getNumManifolds() { return manifolds.size(); }                               // bp/Dispatcher.java:62
    @Override public PersistentManifold getManifoldByIndexInternal(int i) { return manifolds.getQuick(i); }  // :64
    @Override public ObjectArrayList<PersistentManifold> getInternalManifoldPointer() { return manifolds; }
    /** disp/CollisionDispatcher.java:198-223.  On the device the same rule runs in k_classify: activity from
     *  b2c_set_activation, checkCollideWith from b2c_set_no_collide_pairs (the constraint-linked pairs the shim mirrors). */
    @Override public boolean needsCollision(CollisionObject a, CollisionObject b) {
        if (!a.isActive() && !b.isActive()) return false;
        return a.checkCollideWith(b);
    }
    @Override public boolean needsResponse(CollisionObject a, CollisionObject b) {
        return a.hasContactResponse() && b.hasContactResponse() && (!a.isStaticOrKinematicObject() || !b.isStaticOrKinematicObject());
    }
    // per-pair algorithm objects do not exist on this path
    @Override public CollisionAlgorithm findAlgorithm(CollisionObject a, CollisionObject b, PersistentManifold shared) { throw new UnsupportedOperationException("device narrowphase"); }
    @Override public PersistentManifold getNewManifold(Object a, Object b) { throw new UnsupportedOperationException("device narrowphase"); }
    @Override public void releaseManifold(PersistentManifold m) { }
    @Override public void clearManifold(PersistentManifold m) { m.clearManifold(); }
    @Override public void freeCollisionAlgorithm(CollisionAlgorithm algo) { }
}
