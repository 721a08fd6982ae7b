// trace_kernels.cuh -- closest-hit kernels over the LBVH ("extend" stage of the wavefront and the standalone
// Raycaster.query entry point).
#pragma once
#include "internal.h"
#include "common.cuh"
#include "traverse.cuh"
#include "cwbvh.cuh"

template <bool WIDE>
__device__ __forceinline__ RayHit trace_any(const float4* __restrict__ nodes, const float4* __restrict__ tris, Vec3 o, Vec3 d, float t_far, float eps,
                                            bool& overflow) {
    if (WIDE) return cw_trace_one(nodes, tris, o, d, t_far, eps, overflow);
    return trace_one(nodes, tris, o, d, t_far, eps, overflow);
}

// Standalone query: rays as two (R,3) fp32 arrays (the Raycaster.query layout, raycaster.py:20-24).
template <bool WIDE>
__global__ void __launch_bounds__(128) k_trace_aos(const float4* __restrict__ nodes, const float4* __restrict__ tris,
                                                   const float* __restrict__ ro, const float* __restrict__ rd, float* __restrict__ out_t,
                                                   int32_t* __restrict__ out_i, float t_far, float eps, int64_t n, int* __restrict__ flags) {
    int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    Vec3 o = v3(ro[3 * r], ro[3 * r + 1], ro[3 * r + 2]);
    Vec3 d = v3(rd[3 * r], rd[3 * r + 1], rd[3 * r + 2]);
    bool overflow = false;
    RayHit h = trace_any<WIDE>(nodes, tris, o, d, t_far, eps, overflow);
    out_t[r] = h.t;
    out_i[r] = h.id;
    if (overflow) atomicAdd(&flags[0], 1);
}

static inline cudaError_t launch_trace_aos(BvhHandle* h, const float* ro, const float* rd, float* out_t, int32_t* out_i, float t_far,
                                           int64_t n, cudaStream_t s) {
    const int T = 128;
    if (h->wide) k_trace_aos<true><<<(unsigned)((n + T - 1) / T), T, 0, s>>>(h->nodes, h->packed, ro, rd, out_t, out_i, t_far, h->eps, n, h->dev_flags);
    else k_trace_aos<false><<<(unsigned)((n + T - 1) / T), T, 0, s>>>(h->nodes, h->packed, ro, rd, out_t, out_i, t_far, h->eps, n, h->dev_flags);
    return cudaGetLastError();
}
