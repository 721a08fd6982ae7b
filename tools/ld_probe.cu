// ld_probe.cu -- divergent record fetch probe for the BVH node layout: every lane follows its own chain of pseudo-random
// records (the next index depends on the data just loaded, like a traversal), with the record fetched as
//   A: 80-byte records, 5 x LDG.128        (node format 1/2)
//   B: 96-byte records, 3 x LDG.256        (ld.global.nc.v8.f32, sm_100+)
//   C: 96-byte records, 6 x LDG.128
//   D: 64-byte records, 2 x LDG.256
//   E: 64-byte records, 4 x LDG.128
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/_bin/ld_probe tools/ld_probe.cu
#include <cstdio>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>

struct F8 { float v[8]; };
__device__ __forceinline__ F8 ld256(const void* p) {
    F8 r;
    asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
        : "=f"(r.v[0]), "=f"(r.v[1]), "=f"(r.v[2]), "=f"(r.v[3]), "=f"(r.v[4]), "=f"(r.v[5]), "=f"(r.v[6]), "=f"(r.v[7]) : "l"(p));
    return r;
}
__device__ __forceinline__ uint32_t mix(uint32_t x) { x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x; }

template <int MODE>
__global__ void __launch_bounds__(128, 9) k_probe(const float4* __restrict__ recs, uint32_t n_recs, int iters, float* __restrict__ out) {
    uint32_t idx = mix(blockIdx.x * blockDim.x + threadIdx.x) % n_recs;
    float acc = 0.0f;
    for (int it = 0; it < iters; ++it) {
        uint32_t h;
        if (MODE == 0) {
            const float4* p = recs + 5 * (size_t)idx;
            float4 a = __ldg(p), b = __ldg(p + 1), c = __ldg(p + 2), d = __ldg(p + 3), e = __ldg(p + 4);
            acc += a.x + b.y + c.z + d.w + e.x + e.w;
            h = __float_as_uint(a.w) ^ __float_as_uint(c.x) ^ __float_as_uint(e.y);
        } else if (MODE == 1) {
            const float4* p = recs + 6 * (size_t)idx;
            F8 a = ld256(p), b = ld256(p + 2), c = ld256(p + 4);
            acc += a.v[0] + a.v[5] + b.v[2] + b.v[7] + c.v[0] + c.v[7];
            h = __float_as_uint(a.v[3]) ^ __float_as_uint(b.v[0]) ^ __float_as_uint(c.v[5]);
        } else if (MODE == 2) {
            const float4* p = recs + 6 * (size_t)idx;
            float4 a = __ldg(p), b = __ldg(p + 1), c = __ldg(p + 2), d = __ldg(p + 3), e = __ldg(p + 4), f = __ldg(p + 5);
            acc += a.x + b.y + c.z + d.w + e.x + f.w;
            h = __float_as_uint(a.w) ^ __float_as_uint(c.x) ^ __float_as_uint(f.y);
        } else if (MODE == 3) {
            const float4* p = recs + 4 * (size_t)idx;
            F8 a = ld256(p), b = ld256(p + 2);
            acc += a.v[0] + a.v[5] + b.v[2] + b.v[7];
            h = __float_as_uint(a.v[3]) ^ __float_as_uint(b.v[0]);
        } else {
            const float4* p = recs + 4 * (size_t)idx;
            float4 a = __ldg(p), b = __ldg(p + 1), c = __ldg(p + 2), d = __ldg(p + 3);
            acc += a.x + b.y + c.z + d.w;
            h = __float_as_uint(a.w) ^ __float_as_uint(c.x);
        }
        // ~200 dependent ALU/FMA instructions would sit here in the traversal; keep a few so that the loads are not back to back
#pragma unroll
        for (int j = 0; j < 16; ++j) acc = fmaf(acc, 1.0001f, 0.5f);
        idx = mix(h + it) % n_recs;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <int MODE>
float run(const float4* recs, uint32_t n_recs, int iters, float* out, int grid) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k_probe<MODE><<<grid, 128>>>(recs, n_recs, iters, out);
    cudaEventRecord(e0);
    k_probe<MODE><<<grid, 128>>>(recs, n_recs, iters, out);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    return ms;
}

int main() {
    const int grid = 148 * 9, iters = 256;
    const size_t max_bytes = 64u << 20;
    float4* recs;
    float* out;
    cudaMalloc(&recs, max_bytes);
    cudaMalloc(&out, sizeof(float) * grid * 128);
    std::vector<uint32_t> h(max_bytes / 4);
    uint32_t s = 12345;
    for (auto& x : h) { s = s * 1664525u + 1013904223u; x = (s >> 9) | 0x3f000000u; }
    cudaMemcpy(recs, h.data(), max_bytes, cudaMemcpyHostToDevice);
    const char* names[5] = {"A 80B 5xLDG.128", "B 96B 3xLDG.256", "C 96B 6xLDG.128", "D 64B 2xLDG.256", "E 64B 4xLDG.128"};
    const int rec_bytes[5] = {80, 96, 96, 64, 64};
    const uint32_t counts[3] = {1000, 60000, 240000};  // L1-resident, ~5 MB, ~19-23 MB (the 2M-triangle scene has 240k nodes)
    printf("{\"grid\": %d, \"block\": 128, \"iters\": %d, \"results\": [\n", grid, iters);
    for (int ci = 0; ci < 3; ++ci)
        for (int m = 0; m < 5; ++m) {
            float ms = m == 0 ? run<0>(recs, counts[ci], iters, out, grid) : m == 1 ? run<1>(recs, counts[ci], iters, out, grid)
                     : m == 2 ? run<2>(recs, counts[ci], iters, out, grid) : m == 3 ? run<3>(recs, counts[ci], iters, out, grid)
                                                                                    : run<4>(recs, counts[ci], iters, out, grid);
            double visits = (double)grid * 128 * iters;
            printf("  {\"mode\": \"%s\", \"records\": %u, \"footprint_mb\": %.2f, \"ms\": %.4f, \"gvisits_per_s\": %.3f}%s\n", names[m], counts[ci],
                   counts[ci] * (double)rec_bytes[m] / 1e6, ms, visits / ms / 1e6, (ci == 2 && m == 4) ? "" : ",");
        }
    printf("]}\n");
    cudaError_t err = cudaDeviceSynchronize();
    if (err != cudaSuccess) { fprintf(stderr, "CUDA error: %s\n", cudaGetErrorString(err)); return 1; }
    return 0;
}
