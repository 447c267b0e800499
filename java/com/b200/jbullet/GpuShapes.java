package com.b200.jbullet;

import java.lang.foreign.Arena;
import java.lang.foreign.MemorySegment;
import java.nio.ByteBuffer;
import java.util.IdentityHashMap;

import com.badlogic.gdx.math.Matrix3;
import com.badlogic.gdx.math.Vector3;
import com.bulletphysics.collision.broadphase.BroadphaseNativeType;
import com.bulletphysics.collision.dispatch.CollisionObject;
import com.bulletphysics.collision.shapes.*;
import com.bulletphysics.linearmath.Transform;

import static java.lang.foreign.ValueLayout.*;

/**
 * Registers the reference's {@link CollisionShape} objects with the device (include/b2c.h "shapes") and creates proxies.
 * NOT COMPILED IN THIS REPOSITORY'S IMAGE (no JDK / libgdx jar).  One shape object = one device shape id, shared between
 * the bodies that share the object (the reference shares shapes the same way).
 */
final class GpuShapes {
    private static final IdentityHashMap<CollisionShape, Integer> IDS = new IdentityHashMap<CollisionShape, Integer>();

    /** row-major basis + origin, the 12 floats b2c_proxy_create / b2c_shape_register_compound take */
    static void putTransform(MemorySegment dst, long index, Transform t) {
        float[] m = t.basis.val;   // libgdx Matrix3 is column-major: M01 is row 0, column 1
        dst.setAtIndex(JAVA_FLOAT, index, m[Matrix3.M00]); dst.setAtIndex(JAVA_FLOAT, index + 1, m[Matrix3.M01]); dst.setAtIndex(JAVA_FLOAT, index + 2, m[Matrix3.M02]);
        dst.setAtIndex(JAVA_FLOAT, index + 3, m[Matrix3.M10]); dst.setAtIndex(JAVA_FLOAT, index + 4, m[Matrix3.M11]); dst.setAtIndex(JAVA_FLOAT, index + 5, m[Matrix3.M12]);
        dst.setAtIndex(JAVA_FLOAT, index + 6, m[Matrix3.M20]); dst.setAtIndex(JAVA_FLOAT, index + 7, m[Matrix3.M21]); dst.setAtIndex(JAVA_FLOAT, index + 8, m[Matrix3.M22]);
        dst.setAtIndex(JAVA_FLOAT, index + 9, t.origin.x); dst.setAtIndex(JAVA_FLOAT, index + 10, t.origin.y); dst.setAtIndex(JAVA_FLOAT, index + 11, t.origin.z);
    }

    /** The device id of a shape, registering it on first use. */
    static int idOf(MemorySegment ctx, CollisionShape shape) throws Throwable {
        Integer known = IDS.get(shape);
        if (known != null) return known;
        try (Arena a = Arena.ofConfined()) {
            MemorySegment out = a.allocate(JAVA_INT);
            Vector3 tmp = new Vector3();
            BroadphaseNativeType type = shape.getShapeType();
            if (type == BroadphaseNativeType.BOX_SHAPE_PROXYTYPE) {
                // sh/BoxShape.java:46-50 stores halfExtents * scaling - margin; hand back the extents WITH margin and the margin
                ((BoxShape) shape).getHalfExtentsWithMargin(tmp);
                MemorySegment he = a.allocateFrom(JAVA_FLOAT, tmp.x, tmp.y, tmp.z);
                B2C.check((int) B2C.shapeBox.invokeExact(ctx, he, shape.getMargin(), out), ctx);
            } else if (type == BroadphaseNativeType.SPHERE_SHAPE_PROXYTYPE) {
                B2C.check((int) B2C.shapeSphere.invokeExact(ctx, ((SphereShape) shape).getRadius(), out), ctx);
            } else if (type == BroadphaseNativeType.CONVEX_HULL_SHAPE_PROXYTYPE) {
                ConvexHullShape hull = (ConvexHullShape) shape;
                int n = hull.getNumPoints();
                MemorySegment pts = a.allocate(JAVA_FLOAT, 3L * n);
                for (int i = 0; i < n; i++) {
                    Vector3 p = hull.getPoints().getQuick(i);
                    pts.setAtIndex(JAVA_FLOAT, 3L * i, p.x); pts.setAtIndex(JAVA_FLOAT, 3L * i + 1, p.y); pts.setAtIndex(JAVA_FLOAT, 3L * i + 2, p.z);
                }
                B2C.check((int) B2C.shapeHull.invokeExact(ctx, pts, n, shape.getMargin(), out), ctx);
            } else if (type == BroadphaseNativeType.STATIC_PLANE_PROXYTYPE) {
                StaticPlaneShape pl = (StaticPlaneShape) shape;
                pl.getPlaneNormal(tmp);
                MemorySegment nrm = a.allocateFrom(JAVA_FLOAT, tmp.x, tmp.y, tmp.z);
                B2C.check((int) B2C.shapePlane.invokeExact(ctx, nrm, pl.getPlaneConstant(), out), ctx);
            } else if (type == BroadphaseNativeType.TRIANGLE_MESH_SHAPE_PROXYTYPE) {
                registerMesh(ctx, (BvhTriangleMeshShape) shape, a, out);
            } else if (type == BroadphaseNativeType.COMPOUND_SHAPE_PROXYTYPE) {
                CompoundShape cs = (CompoundShape) shape;
                int n = cs.getNumChildShapes();
                MemorySegment ids = a.allocate(JAVA_INT, n), xf = a.allocate(JAVA_FLOAT, 12L * n);
                Transform t = new Transform();
                for (int i = 0; i < n; i++) {   // sh/CompoundShape.java:50-82: children in addChildShape order
                    ids.setAtIndex(JAVA_INT, i, idOf(ctx, cs.getChildShape(i)));
                    putTransform(xf, 12L * i, cs.getChildTransform(i, t));
                }
                B2C.check((int) B2C.shapeCompound.invokeExact(ctx, n, ids, xf, out), ctx);
            } else {
                throw new UnsupportedOperationException("shape type not on the device path: " + type);
            }
            int id = out.get(JAVA_INT, 0);
            IDS.put(shape, id);
            return id;
        }
    }

    /** sh/TriangleIndexVertexArray.java:72-100: every IndexedMesh part with its own index type (SHORT or INTEGER). */
    private static void registerMesh(MemorySegment ctx, BvhTriangleMeshShape shape, Arena a, MemorySegment out) throws Throwable {
        StridingMeshInterface mi = shape.getMeshInterface();
        int parts = mi.getNumSubParts();
        MemorySegment descs = a.allocate(B2C.INDEXED_MESH, parts);
        for (int p = 0; p < parts; p++) {
            ByteBufferVertexData d = (ByteBufferVertexData) mi.getLockedReadOnlyVertexIndexBase(p);
            // direct buffers are passed by address, heap buffers copied
            MemorySegment v = MemorySegment.ofBuffer(d.vertexData), i = MemorySegment.ofBuffer(d.indexData);
            if (!v.isNative()) v = a.allocate(v.byteSize()).copyFrom(v);
            if (!i.isNative()) i = a.allocate(i.byteSize()).copyFrom(i);
            MemorySegment m = descs.asSlice(p * B2C.INDEXED_MESH.byteSize(), B2C.INDEXED_MESH.byteSize());
            m.set(ADDRESS, 0, v); m.set(JAVA_INT, 8, d.vertexCount); m.set(JAVA_INT, 12, d.vertexStride);
            m.set(ADDRESS, 16, i); m.set(JAVA_INT, 24, d.indexCount / 3); m.set(JAVA_INT, 28, d.indexStride * 3);
            m.set(JAVA_INT, 32, d.indexType == ScalarType.SHORT ? 2 : 4);
            mi.unLockReadOnlyVertexBase(p);
        }
        Vector3 s = mi.getScaling(new Vector3());
        MemorySegment sc = a.allocateFrom(JAVA_FLOAT, s.x, s.y, s.z);
        B2C.check((int) B2C.shapeMeshParts.invokeExact(ctx, descs, parts, sc, out), ctx);
    }

    /** CollisionWorld.addCollisionObject -> BroadphaseInterface.createProxy (disp/CollisionWorld.java:102-121). */
    static int createProxyFor(MemorySegment ctx, Object userPtr, short group, short mask, MemorySegment scratchInt) {
        CollisionObject co = (CollisionObject) userPtr;
        try (Arena a = Arena.ofConfined()) {
            int shape = idOf(ctx, co.getCollisionShape());
            MemorySegment xf = a.allocate(JAVA_FLOAT, 12);
            putTransform(xf, 0, co.getWorldTransform(new Transform()));
            B2C.check((int) B2C.proxyCreate.invokeExact(ctx, shape, xf, group, mask, co.isStaticObject() ? 1 : 0, 0, scratchInt), ctx);
            int uid = scratchInt.get(JAVA_INT, 0);
            B2C.check((int) B2C.proxySetMaterial.invokeExact(ctx, uid, co.getFriction(), co.getRestitution()), ctx);
            return uid;
        } catch (Throwable t) { throw new RuntimeException(t); }
    }

    private GpuShapes() { }
}
