// common.cuh -- shared definitions for the diffrp_b200 CUDA library (sm_100a).
//
// All per-element logic is written as DRP_HD functions so that the same source can be compiled (a) by nvcc for
// the device and (b) by g++ with -DDRP_HOSTSIM into a serial host simulator used ONLY by tests/ to debug kernel
// logic in a container without a GPU (tests/hostsim/).  The product never runs the host build.
#pragma once
#include <stdint.h>
#include <string.h>
#include <math.h>
#include <float.h>

#ifdef DRP_HOSTSIM
#include <cmath>
#include <cstdlib>
#define DRP_HD inline
#define DRP_D inline
struct float2 { float x, y; };
struct float3 { float x, y, z; };
struct alignas(16) float4 { float x, y, z, w; };
struct alignas(16) int4 { int x, y, z, w; };
struct alignas(8) int2 { int x, y; };
struct alignas(8) uint2 { unsigned x, y; };
struct alignas(16) uint4 { unsigned x, y, z, w; };
static inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
static inline float3 make_float3(float x, float y, float z) { return float3{x, y, z}; }
static inline float2 make_float2(float x, float y) { return float2{x, y}; }
#else
#include <cuda_runtime.h>
#define DRP_HD __host__ __device__ __forceinline__
#define DRP_D __device__ __forceinline__
#endif

// ---- portable bit casts / integer helpers (device intrinsics on the GPU, plain C++ on the host) ---------------
DRP_HD int f2i(float f) {
#ifdef __CUDA_ARCH__
    return __float_as_int(f);
#else
    int i; memcpy(&i, &f, 4); return i;
#endif
}
DRP_HD float i2f(int i) {
#ifdef __CUDA_ARCH__
    return __int_as_float(i);
#else
    float f; memcpy(&f, &i, 4); return f;
#endif
}
DRP_HD uint32_t f2u(float f) { return (uint32_t)f2i(f); }
DRP_HD float u2f(uint32_t u) { return i2f((int)u); }
DRP_HD int clz32(uint32_t x) {
#ifdef __CUDA_ARCH__
    return __clz((int)x);
#else
    return x == 0 ? 32 : __builtin_clz(x);
#endif
}
DRP_HD int clz64(uint64_t x) {
#ifdef __CUDA_ARCH__
    return __clzll((long long)x);
#else
    return x == 0 ? 64 : __builtin_clzll((unsigned long long)x);
#endif
}
DRP_HD uint32_t umulhi32(uint32_t a, uint32_t b) {
#ifdef __CUDA_ARCH__
    return __umulhi(a, b);
#else
    return (uint32_t)(((uint64_t)a * b) >> 32);
#endif
}
template <typename T>
DRP_HD T ldg(const T* p) {
#ifdef __CUDA_ARCH__
    return __ldg(p);
#else
    return *p;
#endif
}

// ---- exactly-rounded arithmetic (never contracted): the triangle test must be bit-identical to the oracle -----
#if defined(__CUDA_ARCH__)
DRP_HD float x_mul(float a, float b) { return __fmul_rn(a, b); }
DRP_HD float x_add(float a, float b) { return __fadd_rn(a, b); }
DRP_HD float x_sub(float a, float b) { return __fsub_rn(a, b); }
DRP_HD float x_fma(float a, float b, float c) { return __fmaf_rn(a, b, c); }
DRP_HD float x_rcp(float a) { return __frcp_rn(a); }
#else
// host build must be compiled with -ffp-contract=off
DRP_HD float x_mul(float a, float b) { return a * b; }
DRP_HD float x_add(float a, float b) { return a + b; }
DRP_HD float x_sub(float a, float b) { return a - b; }
DRP_HD float x_fma(float a, float b, float c) { return fmaf(a, b, c); }
DRP_HD float x_rcp(float a) { return 1.0f / a; }
#endif

// sin / cos of 2*pi*x for x in [0,1): the device uses sincospif (exact range reduction, compact code);
// the host simulator evaluates the same quantity with libm.
DRP_HD void sincos_2pi(float x, float* s, float* c) {
#ifdef __CUDA_ARCH__
    sincospif(2.0f * x, s, c);
#else
    sincosf(x * 6.283185307179586f, s, c);
#endif
}

struct Vec3 {
    float x, y, z;
};
DRP_HD Vec3 v3(float x, float y, float z) { Vec3 r; r.x = x; r.y = y; r.z = z; return r; }
DRP_HD Vec3 operator+(Vec3 a, Vec3 b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
DRP_HD Vec3 operator-(Vec3 a, Vec3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
DRP_HD Vec3 operator*(Vec3 a, float s) { return v3(a.x * s, a.y * s, a.z * s); }
DRP_HD Vec3 operator*(Vec3 a, Vec3 b) { return v3(a.x * b.x, a.y * b.y, a.z * b.z); }
DRP_HD float dot(Vec3 a, Vec3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
DRP_HD Vec3 cross(Vec3 a, Vec3 b) { return v3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
// F.normalize semantics: x / max(||x||, 1e-12)
DRP_HD Vec3 normalize_ref(Vec3 a) {
    float l = fmaxf(sqrtf(dot(a, a)), 1e-12f);
    return v3(a.x / l, a.y / l, a.z / l);
}

// exact variants used by the triangle test (orc_raycast.c: v_sub / v_cross / v_dot)
DRP_HD Vec3 x_sub3(Vec3 a, Vec3 b) { return v3(x_sub(a.x, b.x), x_sub(a.y, b.y), x_sub(a.z, b.z)); }
DRP_HD Vec3 x_cross3(Vec3 a, Vec3 b) {
    return v3(x_fma(a.y, b.z, -x_mul(a.z, b.y)), x_fma(a.z, b.x, -x_mul(a.x, b.z)), x_fma(a.x, b.y, -x_mul(a.y, b.x)));
}
DRP_HD float x_dot3(Vec3 a, Vec3 b) { return x_add(x_add(x_mul(a.x, b.x), x_mul(a.y, b.y)), x_mul(a.z, b.z)); }

// Moller-Trumbore exactly as the reference's _ray_tri_intersect (diffrp/utils/raycaster.py:58-79), which unbinds the
// triangle's vertices (0,1,2) as (v1, v2, v0).  Returns true and t on a hit.  `eps` is the |det| threshold
// (PathTracingSessionOptions.raycaster_epsilon; 0 disables it).
// Edge form: e1 = A - C and e2 = B - C are rounded exactly as below, so the traversal layouts store (e1, e2, C) per triangle
// (DRP_TRI_EDGES) and skip the six subtractions per test without changing a bit of the result.
DRP_HD bool tri_test_edges(Vec3 o, Vec3 d, Vec3 e1, Vec3 e2, Vec3 C, float eps, float& t_out) {
    Vec3 cr = x_cross3(d, e2);
    float det = x_dot3(e1, cr);
    float inv_det = x_rcp(det);
    Vec3 s = x_sub3(o, C);
    float u = x_mul(inv_det, x_dot3(s, cr));
    Vec3 sc = x_cross3(s, e1);
    float v = x_mul(inv_det, x_dot3(d, sc));
    float t = x_mul(inv_det, x_dot3(e2, sc));
    t_out = t;
    return (fabsf(det) > eps) & (u >= 0.0f) & (v >= 0.0f) & (x_add(u, v) <= 1.0f) & (t > 0.0f);
}
#ifndef DRP_TRI_EDGES
#define DRP_TRI_EDGES 1
#endif
// 48-byte triangle record of the traversal layouts: (e1, e2, C, primitive id) -- or (A, B, C, id) with DRP_TRI_EDGES=0
DRP_HD void pack_triangle(float4* o, Vec3 A, Vec3 B, Vec3 C, int prim) {
#if DRP_TRI_EDGES
    const Vec3 e1 = x_sub3(A, C), e2 = x_sub3(B, C);
    o[0] = make_float4(e1.x, e1.y, e1.z, e2.x);
    o[1] = make_float4(e2.y, e2.z, C.x, C.y);
#else
    o[0] = make_float4(A.x, A.y, A.z, B.x);
    o[1] = make_float4(B.y, B.z, C.x, C.y);
#endif
    o[2] = make_float4(C.z, i2f(prim), 0.0f, 0.0f);
}
DRP_HD bool tri_test_mt(Vec3 o, Vec3 d, Vec3 A, Vec3 B, Vec3 C, float eps, float& t_out) {
    Vec3 e1 = x_sub3(A, C);
    Vec3 e2 = x_sub3(B, C);
    Vec3 cr = x_cross3(d, e2);
    float det = x_dot3(e1, cr);
    float inv_det = x_rcp(det);
    Vec3 s = x_sub3(o, C);
    float u = x_mul(inv_det, x_dot3(s, cr));
    Vec3 sc = x_cross3(s, e1);
    float v = x_mul(inv_det, x_dot3(d, sc));
    float t = x_mul(inv_det, x_dot3(e2, sc));
    t_out = t;
    return (fabsf(det) > eps) & (u >= 0.0f) & (v >= 0.0f) & (x_add(u, v) <= 1.0f) & (t > 0.0f);
}

// ---- Philox4x32-10 (native RNG; definition shared with oracle/orc_shade.c) -------------------------------------
DRP_HD void philox4x32_10(uint32_t c[4], uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint32_t hi0 = umulhi32(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
        uint32_t hi1 = umulhi32(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
        uint32_t n0 = hi1 ^ c[1] ^ k0, n2 = hi0 ^ c[3] ^ k1;
        c[0] = n0; c[1] = lo1; c[2] = n2; c[3] = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
}
DRP_HD void philox_uniform6(uint64_t seed, uint32_t pixel, uint32_t sample, uint32_t bounce, float u[6]) {
    uint32_t a[4] = {pixel, sample, bounce, 0u}, b[4] = {pixel, sample, bounce, 1u};
    philox4x32_10(a, (uint32_t)seed, (uint32_t)(seed >> 32));
    philox4x32_10(b, (uint32_t)seed, (uint32_t)(seed >> 32));
    const float s = 5.9604644775390625e-08f;  // 2^-24
    u[0] = (float)(a[0] >> 8) * s; u[1] = (float)(a[1] >> 8) * s; u[2] = (float)(a[2] >> 8) * s;
    u[3] = (float)(a[3] >> 8) * s; u[4] = (float)(b[0] >> 8) * s; u[5] = (float)(b[1] >> 8) * s;
}
