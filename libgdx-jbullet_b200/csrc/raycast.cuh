// raycast.cuh — batched closest-hit ray tests against the convex bodies of the world.
//
// Replaces, per ray: CollisionWorld.rayTest with a ClosestRayResultCallback (disp/CollisionWorld.java:553-590, 697-729):
//   objects in world order, needsCollision filter (:664-670), AabbUtil2.rayAabb (lm/AabbUtil2.java:40-108) with the running
//   closestHitFraction as the exit bound, rayTestSingle's convex branch (:260-300) = SubsimplexConvexCast of a zero-radius,
//   zero-margin sphere (np/SubsimplexConvexCast.java:63-190) on the Voronoi simplex solver of gjk.cuh.
// One block per ray.  The reference loop is sequential (every accepted hit lowers the bound the next rayAabb sees), but
// an object's cast result does not depend on that bound, so: (1) all threads test the ray against every body's AABB with
// the widest bound and collect the candidates, (2) the candidates are cast in parallel, (3) one thread replays the
// reference's loop over the candidates in body order with the stored cast results — the same accept/reject decisions, in
// the same order.  Concave shapes (planes, meshes) are not cast (the ray passes through them); that path is next.
#pragma once
#include "broadphase.cuh"
#include "epa.cuh"
#include "gjk.cuh"

namespace b2c {

constexpr int RAY_THREADS = 128;
constexpr int RAY_MAX_CAND = 1024;

__device__ __forceinline__ int rayOutcode(f3 p, f3 h) {  // lm/AabbUtil2.java:40-43
    return (p.x < -h.x ? 0x01 : 0) | (p.x > h.x ? 0x08 : 0) | (p.y < -h.y ? 0x02 : 0) | (p.y > h.y ? 0x10 : 0) |
           (p.z < -h.z ? 0x04 : 0) | (p.z > h.z ? 0x20 : 0);
}
__device__ __forceinline__ float f3get(f3 v, int i) { return i == 0 ? v.x : (i == 1 ? v.y : v.z); }

// lm/AabbUtil2.java:45-108 (the hit normal of the box is not used by the ray test and is not produced)
__device__ __forceinline__ bool rayAabb(f3 rayFrom, f3 rayTo, f3 aabbMin, f3 aabbMax, float param) {
    const f3 he = scl3(sub3(aabbMax, aabbMin), 0.5f);
    const f3 ce = scl3(add3(aabbMax, aabbMin), 0.5f);
    const f3 source = sub3(rayFrom, ce), target = sub3(rayTo, ce);
    const int so = rayOutcode(source, he), to = rayOutcode(target, he);
    if ((so & to) != 0) return false;
    float lambdaEnter = 0.f, lambdaExit = param;
    const f3 r = sub3(target, source);
    float normSign = 1.f;
    int bit = 1;
    for (int j = 0; j < 2; j++) {
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            if (so & bit) {
                float lambda = (-f3get(source, i) - f3get(he, i) * normSign) / f3get(r, i);
                if (lambdaEnter <= lambda) lambdaEnter = lambda;
            } else if (to & bit) {
                float lambda = (-f3get(source, i) - f3get(he, i) * normSign) / f3get(r, i);
                lambdaExit = jminf(lambdaExit, lambda);
            }
            bit <<= 1;
        }
        normSign = -1.f;
    }
    return lambdaEnter <= lambdaExit;
}

// np/SubsimplexConvexCast.java:63-190 with convexA = zero sphere (support (0,0,0)), A's basis = identity, B fixed.
// Returns true when a time of impact is reported; fraction / normal as CastResult.
__device__ __forceinline__ bool rayConvexCast(f3 rayFrom, f3 rayTo, const AnyS& B, const Xf& xfB, float& fractionOut, f3& normalOut) {
    Simplex S;
    S.reset();
    const f3 linVelA = sub3(rayTo, rayFrom);
    const f3 linVelB = sub3(xfB.o, xfB.o);
    float lambda = 0.f;
    f3 originA = rayFrom;
    Xf interpB = xfB;
    const f3 r = sub3(linVelA, linVelB);
    const f3 zero = mk3(0.f, 0.f, 0.f);
    // I * (0,0,0) + origin, evaluated as Transform.transform does
    f3 supA = add3(mk3(zero.x * 1.f + zero.y * 0.f + zero.z * 0.f, zero.x * 0.f + zero.y * 1.f + zero.z * 0.f,
                       zero.x * 0.f + zero.y * 0.f + zero.z * 1.f), rayFrom);
    f3 supB = xfPoint(xfB, B.supportMargin(mulMtV(xfB.m, r)));
    f3 v = sub3(supA, supB);
    int maxIter = 32;
    f3 n = zero;
    float dist2 = len2_3(v);
    const float epsilon = 0.0001f;
    while ((dist2 > epsilon) && (maxIter--) != 0) {
        supA = add3(mk3(zero.x * 1.f + zero.y * 0.f + zero.z * 0.f, zero.x * 0.f + zero.y * 1.f + zero.z * 0.f,
                        zero.x * 0.f + zero.y * 0.f + zero.z * 1.f), originA);
        supB = xfPoint(interpB, B.supportMargin(mulMtV(interpB.m, v)));
        f3 w = sub3(supA, supB);
        const float VdotW = dot3(v, w);
        if (lambda > 1.f) return false;
        if (VdotW > 0.f) {
            const float VdotR = dot3(v, r);
            if (VdotR >= -(B2C_FLT_EPSILON * B2C_FLT_EPSILON)) return false;
            lambda = lambda - VdotW / VdotR;
            const float s = 1.f - lambda;  // lm/VectorUtil.java:137-141 setInterpolate3
            originA = mk3(s * rayFrom.x + lambda * rayTo.x, s * rayFrom.y + lambda * rayTo.y, s * rayFrom.z + lambda * rayTo.z);
            interpB.o = mk3(s * xfB.o.x + lambda * xfB.o.x, s * xfB.o.y + lambda * xfB.o.y, s * xfB.o.z + lambda * xfB.o.z);
            w = sub3(supA, supB);
            n = v;
        }
        S.addVertex(w, supA, supB);
        const bool ok = S.update();
        v = S.cachedV;
        dist2 = ok ? len2_3(v) : 0.f;
    }
    fractionOut = lambda;
    f3 nn = zero;
    if (len2_3(n) >= B2C_FLT_EPSILON * B2C_FLT_EPSILON) nn = nor3(n);
    normalOut = nn;
    if (dot3(nn, r) >= -0.f) return false;  // allowedPenetration = 0
    return true;
}

// tight shape AABBs (shape.getAabb(worldTransform), no contact threshold) of the bodies a ray can hit; others get an empty box
__global__ void __launch_bounds__(256)
k_ray_aabbs(BodyArrays B, const ShapeDev* __restrict__ shapes, int n, float4* __restrict__ rmin, float4* __restrict__ rmax) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float4 mn = make_float4(1.f, 1.f, 1.f, 0.f), mx = make_float4(-1.f, -1.f, -1.f, 0.f);  // w = 0: not a candidate
    const uint8_t fl = B.flags[i];
    if (fl & BF_ALIVE) {
        const ShapeDev s = shapes[B.shape[i]];
        if (s.type == SH_BOX || s.type == SH_SPHERE || s.type == SH_HULL) {
            Xf t = loadXf(B.xf4, i);
            f3 a, b;
            shapeAabb(s, t, a, b);
            mn = make_float4(a.x, a.y, a.z, 1.f);
            mx = make_float4(b.x, b.y, b.z, 0.f);
        }
    }
    rmin[i] = mn;
    rmax[i] = mx;
}

struct RayOut {
    int uid;          // 0 = no hit
    float fraction;   // closestHitFraction (1 = no hit)
    float normal[3];  // hitNormalWorld
    float point[3];   // hitPointWorld
};

__global__ void __launch_bounds__(RAY_THREADS)
k_ray_test(BodyArrays B, const ShapeDev* __restrict__ shapes, const float4* __restrict__ hullPts, int n,
           const float4* __restrict__ rmin, const float4* __restrict__ rmax, const float* __restrict__ rayFrom,
           const float* __restrict__ rayTo, int numRays, uint32_t cbFilter /* group | mask << 16 */, RayOut* __restrict__ out,
           uint32_t* __restrict__ overflow) {
    __shared__ uint32_t sCount;
    __shared__ uint32_t sCand[RAY_MAX_CAND];
    __shared__ float sFrac[RAY_MAX_CAND];
    __shared__ float sNrm[RAY_MAX_CAND][3];
    __shared__ uint8_t sValid[RAY_MAX_CAND];
    for (int ray = blockIdx.x; ray < numRays; ray += gridDim.x) {
        const f3 from = mk3(rayFrom[3 * ray], rayFrom[3 * ray + 1], rayFrom[3 * ray + 2]);
        const f3 to = mk3(rayTo[3 * ray], rayTo[3 * ray + 1], rayTo[3 * ray + 2]);
        __syncthreads();
        if (threadIdx.x == 0) sCount = 0;
        __syncthreads();
        // (1) every body whose AABB the ray meets under the widest bound (closestHitFraction = 1)
        for (int i = threadIdx.x; i < n; i += RAY_THREADS) {
            const float4 mn = __ldg(rmin + i);
            if (mn.w == 0.f) continue;
            if (!filterPass(cbFilter, B.filt[i])) continue;  // RayResultCallback.needsCollision (disp/CollisionWorld.java:664-670)
            const float4 mx = __ldg(rmax + i);
            if (rayAabb(from, to, mk3(mn.x, mn.y, mn.z), mk3(mx.x, mx.y, mx.z), 1.f)) {
                uint32_t k = atomicAdd(&sCount, 1u);
                if (k < RAY_MAX_CAND) sCand[k] = (uint32_t)i;
            }
        }
        __syncthreads();
        uint32_t cnt = sCount;
        if (cnt > RAY_MAX_CAND) {
            if (threadIdx.x == 0) atomicMax(overflow, cnt);
            cnt = RAY_MAX_CAND;
        }
        // (2) the casts, one candidate per thread
        for (uint32_t k = threadIdx.x; k < cnt; k += RAY_THREADS) {
            const int i = (int)sCand[k];
            const ShapeDev s = shapes[B.shape[i]];
            AnyS shp;
            shp.type = s.type;
            shp.h = mk3(s.dims[0], s.dims[1], s.dims[2]);
            shp.ta = shp.tb = shp.tc = mk3(0.f, 0.f, 0.f);
            shp.pts = hullPts + s.pointOffset;
            shp.n = s.numPoints;
            shp.margin = s.margin;
            const Xf t = loadXf(B.xf4, i);
            float fr = 1.f;
            f3 nn = mk3(0.f, 0.f, 0.f);
            const bool hit = rayConvexCast(from, to, shp, t, fr, nn);
            sValid[k] = (hit && len2_3(nn) > 0.0001f) ? 1 : 0;
            sFrac[k] = fr;
            sNrm[k][0] = nn.x; sNrm[k][1] = nn.y; sNrm[k][2] = nn.z;
        }
        __syncthreads();
        // (3) the reference's loop over the candidates in body order
        if (threadIdx.x == 0) {
            float closest = 1.f;
            int hitBody = -1;
            f3 hn = mk3(0.f, 0.f, 0.f);
            uint32_t done = 0;
            int lastIdx = -1;
            while (done < cnt) {
                // next candidate in ascending body index (selection: the candidate lists are short)
                int best = 0x7fffffff;
                uint32_t bk = 0;
                for (uint32_t k = 0; k < cnt; k++) {
                    const int i = (int)sCand[k];
                    if (i > lastIdx && i < best) { best = i; bk = k; }
                }
                lastIdx = best;
                done++;
                if (closest == 0.f) break;
                const float4 mn = __ldg(rmin + best), mx = __ldg(rmax + best);
                if (!rayAabb(from, to, mk3(mn.x, mn.y, mn.z), mk3(mx.x, mx.y, mx.z), closest)) continue;
                if (sValid[bk] && sFrac[bk] < closest) {
                    closest = sFrac[bk];
                    hitBody = best;
                    // castResult.normal.mul(rayFromTrans.basis) with the identity basis, then nor()
                    f3 c = mk3(sNrm[bk][0], sNrm[bk][1], sNrm[bk][2]);
                    hn = nor3(mk3(c.x * 1.f + c.y * 0.f + c.z * 0.f, c.x * 0.f + c.y * 1.f + c.z * 0.f, c.x * 0.f + c.y * 0.f + c.z * 1.f));
                }
            }
            RayOut o;
            o.uid = hitBody + 1;
            o.fraction = closest;
            o.normal[0] = hn.x; o.normal[1] = hn.y; o.normal[2] = hn.z;
            const float s = 1.f - closest;
            f3 pt = hitBody >= 0 ? mk3(s * from.x + closest * to.x, s * from.y + closest * to.y, s * from.z + closest * to.z) : mk3(0.f, 0.f, 0.f);
            o.point[0] = pt.x; o.point[1] = pt.y; o.point[2] = pt.z;
            out[ray] = o;
        }
    }
}

}  // namespace b2c
