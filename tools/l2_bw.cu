// l2_bw.cu -- measures the L2-resident read bandwidth of the GPU ("L2 peak is not in MEASURED_PEAKS.json
// -- measure once with a 64 MB-resident read kernel and record it next to the HBM number").  Stand-alone diagnostic, not part
// of the library.  Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/_bin/l2_bw tools/l2_bw.cu
// Output: one JSON line {"sizes_mb": [...], "gbs": [...]}.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_)); exit(1); } } while (0)

__global__ void __launch_bounds__(256) k_read(const float4* __restrict__ p, size_t n_vec, int passes, float* __restrict__ sink) {
    float acc = 0.0f;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (int it = 0; it < passes; ++it) {
        size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
        // 4 independent 128-bit loads in flight per thread
        for (; i + 3 * stride < n_vec; i += 4 * stride) {
            float4 a = __ldcg(p + i), b = __ldcg(p + i + stride), c = __ldcg(p + i + 2 * stride), d = __ldcg(p + i + 3 * stride);
            acc += a.x + b.y + c.z + d.w;
        }
        for (; i < n_vec; i += stride) acc += __ldcg(p + i).x;
    }
    if (acc == 123.456f) *sink = acc;  // keeps the loads alive
}

int main() {
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    const int grid = prop.multiProcessorCount * 8;
    const int sizes_mb[] = {8, 16, 32, 48, 64, 96, 128, 512};
    float* sink; CK(cudaMalloc(&sink, 4));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    printf("{\"device\": \"%s\", \"l2_bytes\": %d, \"sizes_mb\": [", prop.name, prop.l2CacheSize);
    std::vector<double> res;
    for (size_t k = 0; k < sizeof(sizes_mb) / sizeof(int); ++k) {
        const size_t bytes = (size_t)sizes_mb[k] << 20, n_vec = bytes / 16;
        float4* buf; CK(cudaMalloc(&buf, bytes)); CK(cudaMemset(buf, 0, bytes));
        const int passes = sizes_mb[k] <= 128 ? 64 : 8;
        k_read<<<grid, 256>>>(buf, n_vec, 4, sink);  // warm the cache
        double best = 0;
        for (int rep = 0; rep < 5; ++rep) {
            CK(cudaEventRecord(e0));
            k_read<<<grid, 256>>>(buf, n_vec, passes, sink);
            CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
            float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
            const double gbs = (double)bytes * passes / (ms * 1e-3) / 1e9;
            if (gbs > best) best = gbs;
        }
        res.push_back(best);
        printf("%s%d", k ? ", " : "", sizes_mb[k]);
        CK(cudaFree(buf));
    }
    printf("], \"read_gbs\": [");
    for (size_t k = 0; k < res.size(); ++k) printf("%s%.1f", k ? ", " : "", res[k]);
    printf("]}\n");
    return 0;
}
