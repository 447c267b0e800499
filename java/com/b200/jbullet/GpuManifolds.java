package com.b200.jbullet;

import java.lang.foreign.MemorySegment;

import com.bulletphysics.collision.narrowphase.ManifoldPoint;
import com.bulletphysics.collision.narrowphase.PersistentManifold;

import static java.lang.foreign.ValueLayout.*;

/**
 * Refreshes the Java {@link PersistentManifold}s the island manager and the solver read from the device's contact stream
 * (b2c_get_contacts: 96-byte b2c_manifold_point records).  NOT COMPILED IN THIS REPOSITORY'S IMAGE (no JDK / libgdx jar).
 *
 * Geometry comes from the device; the solver's warm-start state stays on the host and follows a point through
 * {@code src_slot} — the slot of the SAME manifold the point occupied at the start of the step (-1: new point), which is how
 * np/PersistentManifold.java:280-305 (replaceContactPoint keeps appliedImpulse* and userPersistentData) and :235-257
 * (addManifoldPoint) treat them.
 */
final class GpuManifolds {
    /** what the solver leaves in a point between steps (np/ManifoldPoint.java:53-62) */
    static final class SolverState {
        Object userPersistentData;
        float appliedImpulse, appliedImpulseLateral1, appliedImpulseLateral2;
        boolean lateralFrictionInitialized;
        final com.badlogic.gdx.math.Vector3 dir1 = new com.badlogic.gdx.math.Vector3(), dir2 = new com.badlogic.gdx.math.Vector3();
    }

    static SolverState[] snapshotSolverState(PersistentManifold m) {
        SolverState[] s = new SolverState[PersistentManifold.MANIFOLD_CACHE_SIZE];
        for (int k = 0; k < m.getNumContacts(); k++) {
            ManifoldPoint p = m.getContactPoint(k);
            SolverState st = new SolverState();
            st.userPersistentData = p.userPersistentData;
            st.appliedImpulse = p.appliedImpulse;
            st.appliedImpulseLateral1 = p.appliedImpulseLateral1;
            st.appliedImpulseLateral2 = p.appliedImpulseLateral2;
            st.lateralFrictionInitialized = p.lateralFrictionInitialized;
            st.dir1.set(p.lateralFrictionDir1);
            st.dir2.set(p.lateralFrictionDir2);
            s[k] = st;
        }
        return s;
    }

    private static float f(MemorySegment pts, long rec, int word) { return pts.get(JAVA_FLOAT, rec + 4L * word); }
    private static int i(MemorySegment pts, long rec, int word) { return pts.get(JAVA_INT, rec + 4L * word); }

    /**
     * Rebuild the manifold's point cache from {@code n} records starting at record {@code first}.  Record words
     * (include/b2c.h b2c_manifold_point): 0-2 localA, 3-5 localB, 6-8 worldA, 9-11 worldB, 12-14 normalOnB, 15 distance,
     * 16 friction, 17 restitution, 18 lifeTime, 19 src_slot, 20 partId1, 21 index1.  {@code algorithm} is the header's
     * algorithm code (4 = convex vs triangle mesh: partId0 / index0 are -1 there).
     */
    static void fill(PersistentManifold m, MemorySegment pts, int first, int n, SolverState[] old, int algorithm) {
        m.clearManifold();   // np/PersistentManifold.java:374-380: clears the user cache of every point, then cachedPoints = 0
        ManifoldPoint p = new ManifoldPoint();
        for (int k = 0; k < n; k++) {
            long rec = 96L * (first + k);
            p.localPointA.set(f(pts, rec, 0), f(pts, rec, 1), f(pts, rec, 2));
            p.localPointB.set(f(pts, rec, 3), f(pts, rec, 4), f(pts, rec, 5));
            p.positionWorldOnA.set(f(pts, rec, 6), f(pts, rec, 7), f(pts, rec, 8));
            p.positionWorldOnB.set(f(pts, rec, 9), f(pts, rec, 10), f(pts, rec, 11));
            p.normalWorldOnB.set(f(pts, rec, 12), f(pts, rec, 13), f(pts, rec, 14));
            p.distance1 = f(pts, rec, 15);
            p.combinedFriction = f(pts, rec, 16);
            p.combinedRestitution = f(pts, rec, 17);
            p.lifeTime = i(pts, rec, 18);
            int src = i(pts, rec, 19);
            p.partId1 = i(pts, rec, 20);
            p.index1 = i(pts, rec, 21);
            boolean mesh = algorithm == 4;   // b2c_manifold.algorithm: 4 = convex-concave
            p.partId0 = mesh ? -1 : 0;   // disp/ConvexTriangleCallback.java:164 setShapeIdentifiers(-1, -1, partId, triangleIndex)
            p.index0 = mesh ? -1 : 0;
            SolverState st = (old != null && src >= 0 && src < old.length) ? old[src] : null;
            p.userPersistentData = st != null ? st.userPersistentData : null;
            p.appliedImpulse = st != null ? st.appliedImpulse : 0f;
            p.appliedImpulseLateral1 = st != null ? st.appliedImpulseLateral1 : 0f;
            p.appliedImpulseLateral2 = st != null ? st.appliedImpulseLateral2 : 0f;
            p.lateralFrictionInitialized = st != null && st.lateralFrictionInitialized;
            if (st != null) { p.lateralFrictionDir1.set(st.dir1); p.lateralFrictionDir2.set(st.dir2); }
            m.addManifoldPoint(p);   // copies p into the next free slot (np/PersistentManifold.java:235-257); n <= 4, so no reduction runs
        }
    }

    private GpuManifolds() { }
}
