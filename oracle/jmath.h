// ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the product path.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
// legs may build, load or call anything under oracle/.
//
// PARITY UNPINNED: the reference (vbousquet/libgdx-jbullet) ships no tests, golden
// vectors or fixtures for this path, and no JVM exists in the build container, so this
// restatement is pinned only by hand-derived known-answer tests and differential
// checks (tests/test_oracle_*.py), not by reference-produced vectors.
//
// jmath.h — restatement of the libgdx math the reference calls.
// Third-party dependency: com.badlogicgames.gdx:gdx (com.badlogic.gdx.math.Vector3,
// Matrix3, Quaternion), version NOT pinned by the reference (no build file; README.md:4,32).
// The semantics below restate libgdx 1.x's published source for the methods used on the
// collision path.  All arithmetic is IEEE-754 binary32, evaluated left to right, one
// rounding per operation (Java float semantics): compile with -ffp-contract=off and no
// fast-math.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>

namespace orc {

static inline uint32_t floatToIntBits(float f) {  // java.lang.Float.floatToIntBits (canonical NaN)
    if (f != f) return 0x7fc00000u;
    uint32_t u;
    std::memcpy(&u, &f, 4);
    return u;
}
static inline float jsqrt(float x) { return (float)std::sqrt((double)x); }  // (float)Math.sqrt(x)
static inline int jf2i(float f) {  // Java (int) cast of a float: saturating, NaN -> 0
    if (f != f) return 0;
    if (f >= 2147483648.0f) return 2147483647;
    if (f <= -2147483648.0f) return (int)0x80000000;
    return (int)f;
}
static inline float jmaxf(float a, float b) {  // Math.max(float,float)
    if (a != a) return a;
    if (b != b) return b;
    if (a == 0.0f && b == 0.0f) return (std::signbit(a) && std::signbit(b)) ? -0.0f : 0.0f;
    return a >= b ? a : b;
}
static inline float jminf(float a, float b) {  // Math.min(float,float)
    if (a != a) return a;
    if (b != b) return b;
    if (a == 0.0f && b == 0.0f) return (std::signbit(a) || std::signbit(b)) ? -0.0f : 0.0f;
    return a <= b ? a : b;
}
static inline float jabsf(float a) { return std::fabs(a); }

// com.badlogic.gdx.math.Vector3
struct V3 {
    float x, y, z;
    V3() : x(0), y(0), z(0) {}
    V3(float x_, float y_, float z_) : x(x_), y(y_), z(z_) {}
    V3& set(float a, float b, float c) { x = a; y = b; z = c; return *this; }
    V3& set(const V3& v) { x = v.x; y = v.y; z = v.z; return *this; }
    V3& add(const V3& v) { x = x + v.x; y = y + v.y; z = z + v.z; return *this; }
    V3& sub(const V3& v) { x = x - v.x; y = y - v.y; z = z - v.z; return *this; }
    V3& scl(float s) { x = x * s; y = y * s; z = z * s; return *this; }
    float dot(const V3& v) const { return x * v.x + y * v.y + z * v.z; }
    float len2() const { return x * x + y * y + z * z; }
    float len() const { return jsqrt(x * x + y * y + z * z); }
    // crs: this = this x v
    V3& crs(const V3& v) {
        float a = y * v.z - z * v.y, b = z * v.x - x * v.z, c = x * v.y - y * v.x;
        return set(a, b, c);
    }
    V3& nor() {
        float l2 = len2();
        if (l2 == 0.0f || l2 == 1.0f) return *this;
        return scl(1.0f / jsqrt(l2));
    }
    bool equals(const V3& o) const {
        return floatToIntBits(x) == floatToIntBits(o.x) && floatToIntBits(y) == floatToIntBits(o.y) &&
               floatToIntBits(z) == floatToIntBits(o.z);
    }
    float get(int i) const { return i == 0 ? x : (i == 1 ? y : z); }     // VectorUtil.getCoord
    void setc(int i, float v) { if (i == 0) x = v; else if (i == 1) y = v; else z = v; }  // VectorUtil.setCoord
};

// com.badlogic.gdx.math.Matrix3, held row-major here: m[r][c] == val[M<r><c>].
struct M3 {
    float m[3][3];
    M3() { idt(); }
    M3& idt() {
        for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) m[r][c] = (r == c) ? 1.0f : 0.0f;
        return *this;
    }
    M3& set(const M3& o) { std::memcpy(m, o.m, sizeof(m)); return *this; }
    M3& transpose() {
        float t;
        t = m[0][1]; m[0][1] = m[1][0]; m[1][0] = t;
        t = m[0][2]; m[0][2] = m[2][0]; m[2][0] = t;
        t = m[1][2]; m[1][2] = m[2][1]; m[2][1] = t;
        return *this;
    }
    // Matrix3.mul(Matrix3 b): this = this * b
    M3& mul(const M3& b) {
        float r[3][3];
        for (int i = 0; i < 3; i++)
            for (int j = 0; j < 3; j++) r[i][j] = m[i][0] * b.m[0][j] + m[i][1] * b.m[1][j] + m[i][2] * b.m[2][j];
        std::memcpy(m, r, sizeof(m));
        return *this;
    }
};

// Vector3.mul(Matrix3): v = M * v, rows (M00,M01,M02)...
static inline V3& v3mul(V3& v, const M3& a) {
    float x = v.x * a.m[0][0] + v.y * a.m[0][1] + v.z * a.m[0][2];
    float y = v.x * a.m[1][0] + v.y * a.m[1][1] + v.z * a.m[1][2];
    float z = v.x * a.m[2][0] + v.y * a.m[2][1] + v.z * a.m[2][2];
    return v.set(x, y, z);
}

struct Quat { float x, y, z, w; };

// linearmath/MatrixUtil.java:297-316 transposeTransform: dest = mat^T * vec
static inline void transposeTransform(V3& dest, const V3& vec, const M3& mat) {
    float x = mat.m[0][0] * vec.x + mat.m[1][0] * vec.y + mat.m[2][0] * vec.z;
    float y = mat.m[0][1] * vec.x + mat.m[1][1] * vec.y + mat.m[2][1] * vec.z;
    float z = mat.m[0][2] * vec.x + mat.m[1][2] * vec.y + mat.m[2][2] * vec.z;
    dest.x = x; dest.y = y; dest.z = z;
}
// linearmath/MatrixUtil.java:318-335 setRotation(Matrix3, Quaternion)
static inline void setRotation(M3& dest, const Quat& q) {
    float d = q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w;
    float s = 2.0f / d;
    float xs = q.x * s, ys = q.y * s, zs = q.z * s;
    float wx = q.w * xs, wy = q.w * ys, wz = q.w * zs;
    float xx = q.x * xs, xy = q.x * ys, xz = q.x * zs;
    float yy = q.y * ys, yz = q.y * zs, zz = q.z * zs;
    dest.m[0][0] = 1.0f - (yy + zz); dest.m[0][1] = xy - wz; dest.m[0][2] = xz + wy;
    dest.m[1][0] = xy + wz; dest.m[1][1] = 1.0f - (xx + zz); dest.m[1][2] = yz - wx;
    dest.m[2][0] = xz - wy; dest.m[2][1] = yz + wx; dest.m[2][2] = 1.0f - (xx + yy);
}
// linearmath/QuaternionUtil.java:41-46 setRotation(q, axis, angle)
static inline void quatSetRotation(Quat& q, const V3& axis, float angle) {
    float d = axis.len();
    float s = (float)std::sin((double)(angle * 0.5f)) / d;
    q.x = axis.x * s; q.y = axis.y * s; q.z = axis.z * s;
    q.w = (float)std::cos((double)(angle * 0.5f));
}

// linearmath/Transform.java
struct Xf {
    M3 basis;
    V3 origin;
    void set(const Xf& t) { basis.set(t.basis); origin.set(t.origin); }
    void transform(V3& v) const { v3mul(v, basis); v.add(origin); }         // :91-94
    void inverse() { basis.transpose(); origin.scl(-1.0f); v3mul(origin, basis); }  // :101-105
    void mul(const Xf& tr) {                                                    // :112-120  this = this * tr
        V3 vec = tr.origin;
        transform(vec);
        basis.mul(tr.basis);
        origin.set(vec);
    }
    void invXform(const V3& in, V3& out) const {                                // :133-140
        out.set(in).sub(origin);
        M3 mat; mat.set(basis); mat.transpose();
        v3mul(out, mat);
    }
};

static const float CONVEX_DISTANCE_MARGIN = 0.04f;   // BulletGlobals.java:39
static const float FLT_EPSILON_ = 1.19209290e-07f;   // BulletGlobals.java:40
static const float SIMD_INFINITY_ = 3.4028234663852886e38f;  // Float.MAX_VALUE, BulletGlobals.java:48
static const float SIMD_2_PI_ = 6.283185307179586232f;       // BulletGlobals.java:43

}  // namespace orc
