// TEST INFRASTRUCTURE ONLY: host shim of the CUDA constructs the kernels use, for tests/emu (one emulated lane per warp:
// ballots see one lane, shuffles return the caller's value, atomics are plain read-modify-writes).
#pragma once
#include <stdint.h>
#include <string.h>
#include <math.h>
#include <algorithm>
#define __device__
#define __host__
#define __global__ static inline
#define __forceinline__ inline
#define __noinline__
#define __restrict__
#define __shared__
#define __launch_bounds__(...)
#define __align__(x)
using std::min; using std::max;
struct float2 { float x, y; };
struct alignas(16) float4 { float x, y, z, w; };
struct int2 { int x, y; };
struct alignas(16) int4 { int x, y, z, w; };
struct uint3 { unsigned x, y, z; };
static inline float2 make_float2(float x, float y) { return {x, y}; }
static inline float4 make_float4(float x, float y, float z, float w) { return {x, y, z, w}; }
static inline int2 make_int2(int x, int y) { return {x, y}; }
static inline int4 make_int4(int x, int y, int z, int w) { return {x, y, z, w}; }
static uint3 threadIdx, blockIdx, blockDim, gridDim;
template <class T> static inline T __ldg(const T* p) { return *p; }
static inline float __fsqrt_rn(float x) { return sqrtf(x); }
static inline uint32_t __float_as_uint(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static inline float __uint_as_float(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
static inline int __float_as_int(float f) { int u; memcpy(&u, &f, 4); return u; }
static inline float __int_as_float(int u) { float f; memcpy(&f, &u, 4); return f; }
static inline int __float2int_rz(float v) { if (v != v) return 0; if (v >= 2147483648.f) return 2147483647; if (v <= -2147483648.f) return -2147483647 - 1; return (int)v; }
static inline float __uint2float_rn(uint32_t u) { return (float)u; }
template <class T> static inline T atomicAdd(T* p, T v) { T o = *p; *p = o + v; return o; }
template <class T> static inline T atomicMax(T* p, T v) { T o = *p; if (v > o) *p = v; return o; }
template <class T> static inline T atomicMin(T* p, T v) { T o = *p; if (v < o) *p = v; return o; }
template <class T> static inline T atomicCAS(T* p, T c, T v) { T o = *p; if (o == c) *p = v; return o; }
template <class T> static inline T atomicOr(T* p, T v) { T o = *p; *p = o | v; return o; }
template <class T> static inline T atomicExch(T* p, T v) { T o = *p; *p = v; return o; }
static inline void __syncthreads() {}
static inline void __syncwarp(unsigned = 0xffffffffu) {}
static inline void __threadfence() {}
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline int __ffs(unsigned v) { return __builtin_ffs((int)v); }
static inline unsigned __activemask() { return 1u; }
static inline unsigned __ballot_sync(unsigned, bool p) { return p ? 1u : 0u; }
static inline bool __any_sync(unsigned, bool p) { return p; }
template <class T> static inline T __shfl_sync(unsigned, T v, int) { return v; }
template <class T> static inline T __shfl_up_sync(unsigned, T v, int) { return v; }
template <class T> static inline T __shfl_xor_sync(unsigned, T v, int) { return v; }
template <class T> static inline unsigned __match_any_sync(unsigned, T) { return 1u; }
typedef int cudaError_t;
typedef void* cudaStream_t;
enum { cudaSuccess = 0, cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
