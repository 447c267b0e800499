import java.io.*;
import java.nio.ByteBuffer;
import java.nio.ByteOrder;
import java.util.*;
import java.util.zip.*;

import com.badlogic.gdx.math.Matrix3;
import com.badlogic.gdx.math.Vector3;
import com.bulletphysics.collision.broadphase.*;
import com.bulletphysics.collision.dispatch.*;
import com.bulletphysics.collision.narrowphase.ManifoldPoint;
import com.bulletphysics.collision.narrowphase.PersistentManifold;
import com.bulletphysics.collision.shapes.*;
import com.bulletphysics.linearmath.Transform;
import com.bulletphysics.util.ObjectArrayList;

/**
 * Golden-vector driver for the REAL reference (vbousquet/libgdx-jbullet): runs
 * CollisionWorld.performDiscreteCollisionDetection (collision/dispatch/CollisionWorld.java:123-151) over the scenes
 * tools/javaref/export_scenes.py wrote and dumps, per step, what tests/golden/make_golden.py dumps from the oracle:
 *   aabb<k>    (N,6)  float32  broadphase AABB of every proxy (DbvtProxy.aabb / SimpleBroadphaseProxy min,max)
 *   pairs<k>   (P,2)  int32    overlapping pairs as (uid0 < uid1), sorted; uid = global body index + 1
 *   mf_hdr<k>  (M,7)  int32    per manifold: pair uid0, uid1, body0 uid, body1 uid, numContacts, -1, -1
 *   mf_pts<k>  (M,4,18) float32 localA, localB, worldA, worldB, normalOnB, distance, friction, restitution
 *   mf_int<k>  (M,4,6) int32   lifeTime, -1 (src_slot is not a reference field), partId0, partId1, index0, index1
 * into tests/golden/java_<case>.npz, which tests/test_golden.py consumes when present.
 *
 * STAGED, NOT RUN: this image has neither a JDK nor the libgdx jar the reference needs (com.badlogicgames.gdx:gdx).  On a box
 * that has both:   bash tools/javaref/run.sh /path/to/gdx.jar
 * Batched worlds (num_worlds > 1) are one CollisionWorld each, as in the reference; uids are made global here.
 */
public class DumpGolden {
    // ---------------------------------------------------------------- minimal .npz (zip of .npy v1.0, C order, '<f4' / '<i4')
    static final class Arr {
        int[] shape; float[] f; int[] i;
        int size() { int n = 1; for (int s : shape) n *= s; return n; }
        int at(int... idx) { int o = 0; for (int k = 0; k < idx.length; k++) o = o * shape[k] + idx[k]; return o; }
    }

    static Map<String, Arr> readNpz(File file) throws IOException {
        Map<String, Arr> out = new HashMap<String, Arr>();
        ZipInputStream z = new ZipInputStream(new FileInputStream(file));
        for (ZipEntry e; (e = z.getNextEntry()) != null;) {
            ByteArrayOutputStream bo = new ByteArrayOutputStream();
            byte[] buf = new byte[1 << 16];
            for (int n; (n = z.read(buf)) > 0;) bo.write(buf, 0, n);
            ByteBuffer b = ByteBuffer.wrap(bo.toByteArray()).order(ByteOrder.LITTLE_ENDIAN);
            int hlen = b.getShort(8) & 0xffff;
            String hdr = new String(b.array(), 10, hlen, "US-ASCII");
            Arr a = new Arr();
            String sh = hdr.substring(hdr.indexOf("'shape':") + 8);
            sh = sh.substring(sh.indexOf('(') + 1, sh.indexOf(')'));
            List<Integer> dims = new ArrayList<Integer>();
            for (String t : sh.split(",")) if (!t.trim().isEmpty()) dims.add(Integer.parseInt(t.trim()));
            a.shape = new int[dims.size()];
            for (int k = 0; k < dims.size(); k++) a.shape[k] = dims.get(k);
            b.position(10 + hlen);
            int n = a.size();
            if (hdr.contains("<f4")) { a.f = new float[n]; for (int k = 0; k < n; k++) a.f[k] = b.getFloat(); }
            else if (hdr.contains("<i4")) { a.i = new int[n]; for (int k = 0; k < n; k++) a.i[k] = b.getInt(); }
            else throw new IOException("unsupported dtype in " + e.getName() + ": " + hdr);
            out.put(e.getName().replace(".npy", ""), a);
        }
        z.close();
        return out;
    }

    static void writeNpy(ZipOutputStream z, String name, int[] shape, float[] f, int[] i) throws IOException {
        StringBuilder sh = new StringBuilder("(");
        for (int k = 0; k < shape.length; k++) sh.append(shape[k]).append(shape.length == 1 || k + 1 < shape.length ? "," : "");
        sh.append(")");
        String hdr = "{'descr': '" + (f != null ? "<f4" : "<i4") + "', 'fortran_order': False, 'shape': " + sh + ", }";
        int pad = 64 - ((10 + hdr.length() + 1) % 64);
        StringBuilder hb = new StringBuilder(hdr);
        for (int k = 0; k < pad % 64; k++) hb.append(' ');
        hb.append('\n');
        int n = f != null ? f.length : i.length;
        ByteBuffer b = ByteBuffer.allocate(10 + hb.length() + 4 * n).order(ByteOrder.LITTLE_ENDIAN);
        b.put((byte) 0x93).put("NUMPY".getBytes("US-ASCII")).put((byte) 1).put((byte) 0).putShort((short) hb.length());
        b.put(hb.toString().getBytes("US-ASCII"));
        for (int k = 0; k < n; k++) { if (f != null) b.putFloat(f[k]); else b.putInt(i[k]); }
        z.putNextEntry(new ZipEntry(name + ".npy"));
        z.write(b.array());
        z.closeEntry();
    }

    // ---------------------------------------------------------------- scene -> reference objects
    static Transform xfOf(float[] a, int o) {
        Transform t = new Transform();
        float[] m = t.basis.val;   // libgdx Matrix3 is column-major
        m[Matrix3.M00] = a[o]; m[Matrix3.M01] = a[o + 1]; m[Matrix3.M02] = a[o + 2];
        m[Matrix3.M10] = a[o + 3]; m[Matrix3.M11] = a[o + 4]; m[Matrix3.M12] = a[o + 5];
        m[Matrix3.M20] = a[o + 6]; m[Matrix3.M21] = a[o + 7]; m[Matrix3.M22] = a[o + 8];
        t.origin.set(a[o + 9], a[o + 10], a[o + 11]);
        return t;
    }

    static CollisionShape[] buildShapes(Map<String, Arr> s) {
        int[] kind = s.get("shape_kind").i;
        float[] par = s.get("shape_params").f;
        CollisionShape[] out = new CollisionShape[kind.length];
        for (int k = 0; k < kind.length; k++) {
            switch (kind[k]) {
            case 0: out[k] = new BoxShape(new Vector3(par[4 * k], par[4 * k + 1], par[4 * k + 2])); break;
            case 1: out[k] = new SphereShape(par[4 * k]); break;
            case 2: {
                ObjectArrayList<Vector3> pts = new ObjectArrayList<Vector3>();
                float[] hp = s.get("hull_pts").f;
                for (int q = s.get("hull_off").i[k]; q < s.get("hull_off").i[k + 1]; q++) pts.add(new Vector3(hp[3 * q], hp[3 * q + 1], hp[3 * q + 2]));
                out[k] = new ConvexHullShape(pts);
                break;
            }
            case 4: out[k] = new StaticPlaneShape(new Vector3(par[4 * k], par[4 * k + 1], par[4 * k + 2]), par[4 * k + 3]); break;
            case 5: {
                int v0 = s.get("mesh_off").i[k], v1 = s.get("mesh_off").i[k + 1], t0 = s.get("mesh_toff").i[k], t1 = s.get("mesh_toff").i[k + 1];
                ByteBuffer vb = ByteBuffer.allocateDirect(12 * (v1 - v0)).order(ByteOrder.nativeOrder());
                ByteBuffer ib = ByteBuffer.allocateDirect(12 * (t1 - t0)).order(ByteOrder.nativeOrder());
                float[] mv = s.get("mesh_verts").f;
                int[] mt = s.get("mesh_tris").i;
                for (int q = 3 * v0; q < 3 * v1; q++) vb.putFloat(mv[q]);
                for (int q = 3 * t0; q < 3 * t1; q++) ib.putInt(mt[q]);
                vb.flip(); ib.flip();
                TriangleIndexVertexArray tiva = new TriangleIndexVertexArray(t1 - t0, ib, 12, v1 - v0, vb, 12);
                out[k] = new BvhTriangleMeshShape(tiva, true);   // useQuantizedAabbCompression
                break;
            }
            case 6: {
                CompoundShape cs = new CompoundShape();
                float[] cx = s.get("comp_xf").f;
                int[] cc = s.get("comp_child").i;
                for (int q = s.get("comp_off").i[k]; q < s.get("comp_off").i[k + 1]; q++) cs.addChildShape(xfOf(cx, 12 * q), out[cc[q]]);
                out[k] = cs;
                break;
            }
            default: throw new IllegalArgumentException("shape kind " + kind[k]);
            }
        }
        return out;
    }

    public static void main(String[] args) throws Exception {
        File in = new File(args[0]), outFile = new File(args[1]);
        Map<String, Arr> s = readNpz(in);
        int mode = s.get("mode").i[0], steps = s.get("steps").i[0], numWorlds = s.get("num_worlds").i[0];
        CollisionShape[] shapes = buildShapes(s);
        int[] bshape = s.get("body_shape").i, bstatic = s.get("body_static").i, bgroup = s.get("body_group").i, bmask = s.get("body_mask").i,
              bworld = s.get("body_world").i;
        int N = bshape.length;
        CollisionWorld[] worlds = new CollisionWorld[numWorlds];
        for (int w = 0; w < numWorlds; w++) {
            DefaultCollisionConfiguration cfg = new DefaultCollisionConfiguration();
            BroadphaseInterface bp = mode == 0 ? new SimpleBroadphase() : new DbvtBroadphase();
            worlds[w] = new CollisionWorld(new CollisionDispatcher(cfg), bp, cfg);
        }
        CollisionObject[] objs = new CollisionObject[N];
        final IdentityHashMap<Object, Integer> uidOf = new IdentityHashMap<Object, Integer>();
        float[] xf0 = s.get("xf0").f;
        for (int b = 0; b < N; b++) {
            CollisionObject o = new CollisionObject();
            o.setCollisionShape(shapes[bshape[b]]);
            o.setWorldTransform(xfOf(xf0, 12 * b));
            if (bstatic[b] != 0) o.setCollisionFlags(o.getCollisionFlags() | CollisionFlags.STATIC_OBJECT);
            worlds[bworld[b]].addCollisionObject(o, (short) bgroup[b], (short) bmask[b]);
            objs[b] = o;
            uidOf.put(o, b + 1);
        }
        ZipOutputStream z = new ZipOutputStream(new FileOutputStream(outFile));
        Vector3 mn = new Vector3(), mx = new Vector3();
        for (int step = 0; step < steps; step++) {
            float[] xf = s.get("xf" + step).f;
            for (int b = 0; b < N; b++) objs[b].setWorldTransform(xfOf(xf, 12 * b));
            for (CollisionWorld w : worlds) w.performDiscreteCollisionDetection();
            float[] aabb = new float[6 * N];
            for (int b = 0; b < N; b++) {
                BroadphaseProxy p = objs[b].getBroadphaseHandle();
                if (p instanceof DbvtProxy) { mn.set(((DbvtProxy) p).aabb.Mins()); mx.set(((DbvtProxy) p).aabb.Maxs()); }
                else SimpleAabb.get(p, mn, mx);
                aabb[6 * b] = mn.x; aabb[6 * b + 1] = mn.y; aabb[6 * b + 2] = mn.z; aabb[6 * b + 3] = mx.x; aabb[6 * b + 4] = mx.y; aabb[6 * b + 5] = mx.z;
            }
            writeNpy(z, "aabb" + step, new int[] {N, 6}, aabb, null);
            List<long[]> pairs = new ArrayList<long[]>();
            for (CollisionWorld w : worlds) {
                ObjectArrayList<BroadphasePair> arr = w.getBroadphase().getOverlappingPairCache().getOverlappingPairArray();
                for (int k = 0; k < arr.size(); k++) {
                    int a = uidOf.get(arr.getQuick(k).pProxy0.clientObject), b = uidOf.get(arr.getQuick(k).pProxy1.clientObject);
                    pairs.add(new long[] {Math.min(a, b), Math.max(a, b)});
                }
            }
            Collections.sort(pairs, new Comparator<long[]>() {
                public int compare(long[] a, long[] b) { return a[0] != b[0] ? Long.compare(a[0], b[0]) : Long.compare(a[1], b[1]); }
            });
            int[] pi = new int[2 * pairs.size()];
            for (int k = 0; k < pairs.size(); k++) { pi[2 * k] = (int) pairs.get(k)[0]; pi[2 * k + 1] = (int) pairs.get(k)[1]; }
            writeNpy(z, "pairs" + step, new int[] {pairs.size(), 2}, null, pi);
            // manifolds, keyed and ordered by their pair (the dispatcher's own order depends on hash-table history)
            List<PersistentManifold> ms = new ArrayList<PersistentManifold>();
            for (CollisionWorld w : worlds)
                for (int k = 0; k < w.getDispatcher().getNumManifolds(); k++) ms.add(w.getDispatcher().getManifoldByIndexInternal(k));
            Collections.sort(ms, new Comparator<PersistentManifold>() {
                long key(PersistentManifold m) {
                    int a = uidOf.get(m.getBody0()), b = uidOf.get(m.getBody1());
                    return ((long) Math.min(a, b) << 32) | Math.max(a, b);
                }
                public int compare(PersistentManifold a, PersistentManifold b) { return Long.compare(key(a), key(b)); }
            });
            int M = ms.size();
            int[] hdr = new int[7 * M], mi = new int[M * 4 * 6];
            float[] mp = new float[M * 4 * 18];
            for (int k = 0; k < M; k++) {
                PersistentManifold m = ms.get(k);
                int a = uidOf.get(m.getBody0()), b = uidOf.get(m.getBody1());
                hdr[7 * k] = Math.min(a, b); hdr[7 * k + 1] = Math.max(a, b); hdr[7 * k + 2] = a; hdr[7 * k + 3] = b;
                hdr[7 * k + 4] = m.getNumContacts(); hdr[7 * k + 5] = -1; hdr[7 * k + 6] = -1;
                for (int q = 0; q < m.getNumContacts(); q++) {
                    ManifoldPoint p = m.getContactPoint(q);
                    int o = (k * 4 + q) * 18;
                    Vector3[] v = {p.localPointA, p.localPointB, p.positionWorldOnA, p.positionWorldOnB, p.normalWorldOnB};
                    for (int c = 0; c < 5; c++) { mp[o + 3 * c] = v[c].x; mp[o + 3 * c + 1] = v[c].y; mp[o + 3 * c + 2] = v[c].z; }
                    mp[o + 15] = p.distance1; mp[o + 16] = p.combinedFriction; mp[o + 17] = p.combinedRestitution;
                    int oi = (k * 4 + q) * 6;
                    mi[oi] = p.lifeTime; mi[oi + 1] = -1; mi[oi + 2] = p.partId0; mi[oi + 3] = p.partId1; mi[oi + 4] = p.index0; mi[oi + 5] = p.index1;
                }
            }
            writeNpy(z, "mf_hdr" + step, new int[] {M, 7}, null, hdr);
            writeNpy(z, "mf_pts" + step, new int[] {M, 4, 18}, mp, null);
            writeNpy(z, "mf_int" + step, new int[] {M, 4, 6}, null, mi);
        }
        z.close();
        System.out.println("wrote " + outFile);
    }

    /** SimpleBroadphaseProxy keeps min / max protected (bp/SimpleBroadphaseProxy.java:34-35): read them reflectively. */
    static final class SimpleAabb {
        static void get(BroadphaseProxy p, Vector3 mn, Vector3 mx) throws Exception {
            java.lang.reflect.Field a = p.getClass().getDeclaredField("min"), b = p.getClass().getDeclaredField("max");
            a.setAccessible(true); b.setAccessible(true);
            mn.set((Vector3) a.get(p)); mx.set((Vector3) b.get(p));
        }
    }
}
