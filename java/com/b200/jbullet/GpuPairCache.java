package com.b200.jbullet;

import java.lang.foreign.MemorySegment;
import java.util.HashMap;

import com.bulletphysics.collision.broadphase.*;
import com.bulletphysics.util.ObjectArrayList;

import static java.lang.foreign.ValueLayout.JAVA_INT;

/**
 * {@link OverlappingPairCache} (bp/OverlappingPairCache.java:34-54) view over the device pair list.
 * NOT COMPILED IN THIS REPOSITORY'S IMAGE (no JDK / libgdx jar).
 *
 * The device rebuilds the pair set every step (include/b2c.h b2c_calculate_overlapping_pairs); this class keeps the Java
 * {@link BroadphasePair} objects the island manager iterates (disp/SimulationIslandManager.java:57-77) alive across steps —
 * a pair that stays in the cache keeps its object (and its {@code userInfo}), exactly as
 * bp/HashedOverlappingPairCache.java:291-330 keeps them — and replays the add / remove events of the step into the ghost pair
 * callback (bp/HashedOverlappingPairCache.java:135-137, 323-325; disp/GhostPairCallback.java:40-68).
 */
public class GpuPairCache extends OverlappingPairCache {
    private final GpuBroadphase broadphase;
    private final ObjectArrayList<BroadphasePair> pairs = new ObjectArrayList<BroadphasePair>();
    private final HashMap<Long, BroadphasePair> byKey = new HashMap<Long, BroadphasePair>();
    private OverlappingPairCallback ghostPairCallback;
    private OverlapFilterCallback overlapFilterCallback;

    GpuPairCache(GpuBroadphase broadphase) { this.broadphase = broadphase; }

    private static long key(int uid0, int uid1) { return ((long) uid0 << 32) | (uid1 & 0xffffffffL); }

    /**
     * Called by {@link GpuBroadphase#calculateOverlappingPairs}: {@code pairBuf} holds {@code n} rows (uid0 &lt; uid1) sorted
     * lexicographically (b2c_get_pairs); {@code added} / {@code removed} hold the step's deltas (b2c_get_pair_deltas).
     */
    void refresh(MemorySegment pairBuf, int n, MemorySegment added, int nAdded, MemorySegment removed, int nRemoved,
                 ObjectArrayList<GpuBroadphase.GpuProxy> proxies, Dispatcher dispatcher) {
        // pairs that left: drop the object, tell the ghosts (the device has already dropped their manifolds)
        for (int k = 0; k < nRemoved; k++) {
            int u0 = removed.getAtIndex(JAVA_INT, 2L * k), u1 = removed.getAtIndex(JAVA_INT, 2L * k + 1);
            byKey.remove(key(u0, u1));
            if (ghostPairCallback != null)
                ghostPairCallback.removeOverlappingPair(proxies.getQuick(u0 - 1), proxies.getQuick(u1 - 1), dispatcher);
        }
        for (int k = 0; k < nAdded; k++) {
            int u0 = added.getAtIndex(JAVA_INT, 2L * k), u1 = added.getAtIndex(JAVA_INT, 2L * k + 1);
            GpuBroadphase.GpuProxy p0 = proxies.getQuick(u0 - 1), p1 = proxies.getQuick(u1 - 1);
            // a user OverlapFilterCallback is Java code: it post-filters what the device's group/mask test let through
            // (bp/HashedOverlappingPairCache.java:179-188); a rejected pair is simply never listed
            if (overlapFilterCallback != null && !overlapFilterCallback.needBroadphaseCollision(p0, p1)) continue;
            byKey.put(key(u0, u1), new BroadphasePair(p0, p1));   // pProxy0.getUid() < pProxy1.getUid(), :292-296
            if (ghostPairCallback != null) ghostPairCallback.addOverlappingPair(p0, p1);
        }
        // the array in the device's (sorted) order: what getOverlappingPairArray returns
        pairs.clear();
        for (int k = 0; k < n; k++) {
            BroadphasePair p = byKey.get(key(pairBuf.getAtIndex(JAVA_INT, 2L * k), pairBuf.getAtIndex(JAVA_INT, 2L * k + 1)));
            if (p != null) pairs.add(p);
        }
    }

    @Override public ObjectArrayList<BroadphasePair> getOverlappingPairArray() { return pairs; }     // :36
    @Override public int getNumOverlappingPairs() { return pairs.size(); }                          // :40

    @Override
    public BroadphasePair findPair(BroadphaseProxy proxy0, BroadphaseProxy proxy1) {                // :48
        int a = proxy0.getUid(), b = proxy1.getUid();
        return byKey.get(key(Math.min(a, b), Math.max(a, b)));
    }

    @Override
    public void processAllOverlappingPairs(OverlapCallback callback, Dispatcher dispatcher) {       // :46
        // the reference removes the pairs the callback rejects (bp/HashedOverlappingPairCache.java:190-211); here they only
        // disappear from the Java view — the device list is rebuilt from the AABBs at the next calculateOverlappingPairs
        for (int i = 0; i < pairs.size();) {
            BroadphasePair pair = pairs.getQuick(i);
            if (callback.processOverlap(pair)) {
                byKey.remove(key(pair.pProxy0.getUid(), pair.pProxy1.getUid()));
                pairs.removeQuick(i);
            } else {
                i++;
            }
        }
    }

    // The device owns pair lifetime (pairs are a pure function of the AABBs and filters) and the per-pair algorithms /
    // manifolds (cleaned when a pair leaves, bp/HashedOverlappingPairCache.java:129-174), so the mutators below have nothing
    // to do on the host side.  addOverlappingPair / removeOverlappingPair are what a CPU broadphase would call INTO the
    // cache; nothing on this path does.
    @Override public BroadphasePair addOverlappingPair(BroadphaseProxy p0, BroadphaseProxy p1) { return findPair(p0, p1); }
    @Override public Object removeOverlappingPair(BroadphaseProxy p0, BroadphaseProxy p1, Dispatcher d) { return null; }
    @Override public void removeOverlappingPairsContainingProxy(BroadphaseProxy proxy, Dispatcher dispatcher) { }  // b2c_proxy_destroy does it
    @Override public void cleanOverlappingPair(BroadphasePair pair, Dispatcher dispatcher) { }
    @Override public void cleanProxyFromPairs(BroadphaseProxy proxy, Dispatcher dispatcher) { }
    @Override public boolean hasDeferredRemoval() { return false; }                                                  // :50
    @Override public void setOverlapFilterCallback(OverlapFilterCallback cb) { overlapFilterCallback = cb; }         // :44
    @Override public void setInternalGhostPairCallback(OverlappingPairCallback cb) { ghostPairCallback = cb; }       // :52
}
