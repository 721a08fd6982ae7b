// api.cu -- C ABI of libdiffrp_b200.so: handle table, LBVH build, closest-hit trace.
// See include/diffrp_b200.h for the contract and the reference call sites each entry point replaces.
#include <cub/device/device_radix_sort.cuh>
#include <mutex>
#include <unordered_map>
#include <vector>
#include <cstdio>
#include "internal.h"
#include "common.cuh"
#include "lbvh.cuh"
#include "traverse.cuh"
#include "cwbvh.cuh"
#include <cstdlib>
#include <chrono>

// ---- error / handle plumbing ------------------------------------------------------------------------------------
static thread_local std::string g_last_error;
int g_drp_log_level = 0;
static std::mutex g_mutex;
static std::unordered_map<uint64_t, BvhHandle*> g_handles;
static uint64_t g_next_handle = 1;

void drp_set_error(const std::string& msg) {
    g_last_error = msg;
    if (g_drp_log_level >= 1) fprintf(stderr, "[diffrp_b200] error: %s\n", msg.c_str());
}
BvhHandle* drp_lookup(uint64_t handle) {
    std::lock_guard<std::mutex> lk(g_mutex);
    auto it = g_handles.find(handle);
    return it == g_handles.end() ? nullptr : it->second;
}

extern "C" int drp_abi_version(void) { return DRP_ABI_VERSION; }
extern "C" const char* drp_last_error(void) { return g_last_error.c_str(); }
extern "C" int drp_set_log_level(int level) {
    g_drp_log_level = level;
    return DRP_OK;
}

// ---- build kernels ----------------------------------------------------------------------------------------------
__global__ void k_init_bounds(uint32_t* b) {
    int i = threadIdx.x;
    if (i < 12) b[i] = ((i / 3) % 2 == 0) ? 0xffffffffu : 0u;
}

// grid-stride: every thread folds its triangles into 12 running min / max, then warp shuffles, one shared-memory step per block and ONE set of
// 12 atomics per block (a per-warp atomic set on 12 addresses serialised 65k warps at 2 M triangles: 0.5 ms instead of a streaming pass)
__global__ void __launch_bounds__(256) k_prim_bounds(LbvhBuild b) {
    const float inf = i2f(0x7f800000);
    float v[12] = {inf, inf, inf, -inf, -inf, -inf, inf, inf, inf, -inf, -inf, -inf};
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < b.n; i += gridDim.x * blockDim.x) {
        Vec3 lo, hi;
        lbvh_prim_bounds(b, i, lo, hi);
        const float cx = 0.5f * (lo.x + hi.x), cy = 0.5f * (lo.y + hi.y), cz = 0.5f * (lo.z + hi.z);
        v[0] = fminf(v[0], lo.x); v[1] = fminf(v[1], lo.y); v[2] = fminf(v[2], lo.z);
        v[3] = fmaxf(v[3], hi.x); v[4] = fmaxf(v[4], hi.y); v[5] = fmaxf(v[5], hi.z);
        v[6] = fminf(v[6], cx); v[7] = fminf(v[7], cy); v[8] = fminf(v[8], cz);
        v[9] = fmaxf(v[9], cx); v[10] = fmaxf(v[10], cy); v[11] = fmaxf(v[11], cz);
    }
    __shared__ float part[8][12];
#pragma unroll
    for (int k = 0; k < 12; ++k) {
        const bool is_min = (k / 3) % 2 == 0;
#pragma unroll
        for (int s = 16; s >= 1; s >>= 1) {
            float o = __shfl_xor_sync(0xffffffffu, v[k], s);
            v[k] = is_min ? fminf(v[k], o) : fmaxf(v[k], o);
        }
        if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5][k] = v[k];
    }
    __syncthreads();
    if (threadIdx.x < 12) {
        const int k = threadIdx.x;
        const bool is_min = (k / 3) % 2 == 0;
        float r = part[0][k];
#pragma unroll
        for (int w = 1; w < 8; ++w) r = is_min ? fminf(r, part[w][k]) : fmaxf(r, part[w][k]);
        if (is_min) atomicMin(&b.bounds[k], f2ord(r));
        else atomicMax(&b.bounds[k], f2ord(r));
    }
}
__global__ void __launch_bounds__(256) k_morton(LbvhBuild b) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < b.n) lbvh_morton(b, i);
}
__global__ void __launch_bounds__(256) k_karras(LbvhBuild b) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < b.n - 1) lbvh_karras(b, i);
}
__global__ void __launch_bounds__(256) k_refit(LbvhBuild b) {
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < b.n) lbvh_refit(b, j, [](int* p) { return atomicAdd(p, 1); }, []() { __threadfence(); });
}
__global__ void k_cw_init(CwBuild cw) {
    cw.counters[0] = 1;  // root allocated
    cw.counters[1] = 0;
    cw.counters[8] = 0;
    cw.counters[9] = 1;
    cw.work[0] = 0;
}
__global__ void __launch_bounds__(128) k_cw_collapse(CwBuild cw, int level, float* sah) {
    const int begin = cw.counters[8 + level], end = cw.counters[9 + level];
    if (cw.b.n < 2) {
        if (blockIdx.x == 0 && threadIdx.x == 0 && level == 0) { cw_emit_tiny(cw); *sah = 1.0f; }
        return;
    }
    for (int ni = begin + blockIdx.x * blockDim.x + threadIdx.x; ni < end; ni += gridDim.x * blockDim.x)
        cw_collapse_node(cw, ni, [](int* p, int v) { return atomicAdd(p, v); });
    if (level == 0 && blockIdx.x == 0 && threadIdx.x == 0) *sah = cw.b.box_lo[0].w / fmaxf(cw.b.box_hi[0].w, 1e-30f);
}
__global__ void k_cw_advance(CwBuild cw, int level) { cw.counters[10 + level] = cw.counters[0]; }

template <typename T>
static cudaError_t alloc_async(T** p, size_t count, cudaStream_t s) {
    return cudaMallocAsync((void**)p, sizeof(T) * (count ? count : 1), s);
}

extern "C" int drp_build(const float* verts, const int32_t* tris, int64_t n_verts, int64_t n_tris, int device, void* stream,
                         uint64_t* out_handle) {
    if (!out_handle || n_tris < 0 || n_verts < 0 || (n_tris > 0 && (!verts || !tris))) {
        drp_set_error("drp_build: invalid argument");
        return DRP_ERR_INVALID;
    }
    if (n_tris >= (int64_t(1) << 27)) {
        drp_set_error("drp_build: more than 2^27 triangles are not supported by the leaf encoding");
        return DRP_ERR_INVALID;
    }
    DeviceGuard guard(device);
    if (!guard.ok) { drp_set_error("drp_build: cannot select device"); return DRP_ERR_CUDA; }
    cudaStream_t s = (cudaStream_t)stream;
    const int n = (int)n_tris;
    {   // build scratch comes from the device's default stream-ordered pool; keep freed blocks cached across builds instead of
        // returning them to the driver at every synchronisation (sessions are single-use: one build per frame)
        static std::mutex pool_mutex;
        static bool pool_ready[64] = {false};
        std::lock_guard<std::mutex> lk(pool_mutex);
        if (device >= 0 && device < 64 && !pool_ready[device]) {
            cudaMemPool_t pool;
            if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
                uint64_t threshold = UINT64_MAX;
                cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &threshold);
            }
            pool_ready[device] = true;
        }
    }
    const bool timing = g_drp_log_level >= 4;
    auto now = []() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    const double t_begin = now();
    BvhHandle* h = new BvhHandle();
    h->device = device;
    h->n_tris = n_tris;
    h->n_nodes = n > 1 ? n - 1 : 1;
    h->stack_cap = CW_STACK;

    LbvhBuild b;
    memset(&b, 0, sizeof(b));
    b.verts = verts; b.tris = tris; b.n = n; b.max_leaf = CW_MAX_LEAF;
    uint64_t* keys_in = nullptr; uint32_t* vals_in = nullptr; void* sort_tmp = nullptr;
    size_t sort_bytes = 0;
    const size_t nn = (size_t)(n > 0 ? n : 1);
    DRP_CUDA_CHECK(cudaMallocAsync((void**)&h->nodes, sizeof(float4) * CW_NODE_F4 * (size_t)h->n_nodes, s));
    DRP_CUDA_CHECK(cudaMallocAsync((void**)&h->packed, sizeof(float4) * 3 * nn, s));
    DRP_CUDA_CHECK(cudaMalloc((void**)&h->bounds, sizeof(uint32_t) * 12));
    DRP_CUDA_CHECK(cudaMalloc((void**)&h->sah, sizeof(float)));
    DRP_CUDA_CHECK(cudaHostAlloc((void**)&h->sticky_host, sizeof(int), cudaHostAllocMapped));
    *h->sticky_host = 0;
    DRP_CUDA_CHECK(cudaHostGetDevicePointer((void**)&h->sticky_dev, h->sticky_host, 0));
    b.bounds = h->bounds; b.nodes = h->nodes; b.packed = h->packed;
    DRP_CUDA_CHECK(alloc_async(&b.prim_lo, nn, s));
    DRP_CUDA_CHECK(alloc_async(&b.prim_hi, nn, s));
    DRP_CUDA_CHECK(alloc_async(&keys_in, nn, s));
    DRP_CUDA_CHECK(alloc_async(&vals_in, nn, s));
    DRP_CUDA_CHECK(alloc_async(&b.keys, nn, s));
    DRP_CUDA_CHECK(alloc_async(&b.vals, nn, s));
    DRP_CUDA_CHECK(alloc_async(&b.left, nn, s));
    DRP_CUDA_CHECK(alloc_async(&b.right, nn, s));
    DRP_CUDA_CHECK(alloc_async(&b.parent, 2 * nn, s));
    DRP_CUDA_CHECK(alloc_async(&b.range_first, nn, s));
    DRP_CUDA_CHECK(alloc_async(&b.range_last, nn, s));
    DRP_CUDA_CHECK(alloc_async(&b.box_lo, 2 * nn, s));
    DRP_CUDA_CHECK(alloc_async(&b.box_hi, 2 * nn, s));
    DRP_CUDA_CHECK(alloc_async(&b.arrive, nn, s));
    DRP_CUDA_CHECK(alloc_async(&b.collapsed, nn, s));
    DRP_CUDA_CHECK(alloc_async(&b.count, 2 * nn, s));
    // optimal (DP) collapse tables; the greedy collapse measured 2 % slower traversal and 40 % more nodes (profiles/README.md)
    DRP_CUDA_CHECK(alloc_async(&b.dp_cost, 16 * nn, s));
    DRP_CUDA_CHECK(alloc_async(&b.dp_dec, 16 * nn, s));
    DRP_CUDA_CHECK(cudaMemsetAsync(b.arrive, 0, sizeof(int) * nn, s));
    DRP_CUDA_CHECK(cudaMemsetAsync(b.collapsed, 0, nn, s));

    const double t_alloc = now();
    const int T = 256;
    const int G = (int)((nn + T - 1) / T);
    k_init_bounds<<<1, 32, 0, s>>>(h->bounds);
    if (n > 0) {
        k_prim_bounds<<<std::min(G, 148 * 8), T, 0, s>>>(b);
        // phase 2 writes unsorted keys/vals into the *_in buffers; phase 3 sorts into b.keys / b.vals
        LbvhBuild b2 = b;
        b2.keys = keys_in; b2.vals = vals_in;
        k_morton<<<G, T, 0, s>>>(b2);
        DRP_CUDA_CHECK(cub::DeviceRadixSort::SortPairs(nullptr, sort_bytes, keys_in, b.keys, vals_in, b.vals, n, 0, 63, s));
        DRP_CUDA_CHECK(cudaMallocAsync(&sort_tmp, sort_bytes ? sort_bytes : 1, s));
        DRP_CUDA_CHECK(cub::DeviceRadixSort::SortPairs(sort_tmp, sort_bytes, keys_in, b.keys, vals_in, b.vals, n, 0, 63, s));
        if (n > 1) {
            k_karras<<<G, T, 0, s>>>(b);
            k_refit<<<G, T, 0, s>>>(b);
        }
    }
    int* cw_work = nullptr; int* cw_counters = nullptr;
    {
        CwBuild cw;
        cw.b = b;
        cw.cw_nodes = h->nodes; cw.cw_tris = h->packed; cw.capacity = (int)h->n_nodes;
        DRP_CUDA_CHECK(alloc_async(&cw_work, (size_t)h->n_nodes, s));
        DRP_CUDA_CHECK(alloc_async(&cw_counters, 160, s));
        DRP_CUDA_CHECK(cudaMemsetAsync(cw_counters, 0, sizeof(int) * 160, s));
        cw.work = cw_work; cw.counters = cw_counters;
        k_cw_init<<<1, 1, 0, s>>>(cw);
        // level-synchronous collapse: level L reads its node range from device counters written by level L-1.  The number
        // of levels is bounded by the binary depth; ranges of exhausted levels are empty, so extra launches are no-ops.
        const int grid = 148 * 8;
        int levels_done = 0;
        for (;;) {
            for (int L = levels_done; L < levels_done + 24 && L < 128; ++L) {
                k_cw_collapse<<<(n < 2 || L > 6) ? grid : (L < 3 ? 1 : 64), 128, 0, s>>>(cw, L, h->sah);
                k_cw_advance<<<1, 1, 0, s>>>(cw, L);
            }
            levels_done += 24;
            int hc[160];
            DRP_CUDA_CHECK(cudaMemcpyAsync(hc, cw_counters, sizeof(hc), cudaMemcpyDeviceToHost, s));
            DRP_CUDA_CHECK(cudaStreamSynchronize(s));
            h->n_nodes_used = n < 2 ? 1 : hc[0];
            if (n < 2 || levels_done >= 128 || hc[8 + levels_done] >= hc[9 + levels_done]) break;
        }
    }
    DRP_CUDA_CHECK(cudaGetLastError());
    const double t_launch = now();
    void* temps[] = {b.prim_lo, b.prim_hi, keys_in, vals_in, b.keys, b.vals, b.left, b.right, b.parent, b.range_first,
                     b.range_last, b.box_lo, b.box_hi, b.arrive, b.collapsed, sort_tmp, cw_work, cw_counters, b.count, b.dp_cost, b.dp_dec};
    for (void* p : temps)
        if (p) DRP_CUDA_CHECK(cudaFreeAsync(p, s));
    {
        std::lock_guard<std::mutex> lk(g_mutex);
        *out_handle = g_next_handle++;
        g_handles[*out_handle] = h;
    }
    if (timing) fprintf(stderr, "[diffrp_b200] drp_build n=%d: alloc %.2f ms, launch+collapse-sync %.2f ms, free+register %.2f ms\n", n, t_alloc - t_begin,
                        t_launch - t_alloc, now() - t_launch);
    if (g_drp_log_level >= 4) fprintf(stderr, "[diffrp_b200] built LBVH over %lld triangles (handle %llu)\n", (long long)n_tris, (unsigned long long)*out_handle);
    return DRP_OK;
}

extern "C" int drp_release(uint64_t handle) {
    BvhHandle* h = nullptr;
    {
        std::lock_guard<std::mutex> lk(g_mutex);
        auto it = g_handles.find(handle);
        if (it == g_handles.end()) { drp_set_error("drp_release: unknown handle"); return DRP_ERR_HANDLE; }
        h = it->second;
        g_handles.erase(it);
    }
    DeviceGuard guard(h->device);
    cudaDeviceSynchronize();
    drp_free_workspace(h);
    cudaFreeAsync(h->nodes, 0); cudaFreeAsync(h->packed, 0); cudaFree(h->bounds); cudaFree(h->sah); cudaFreeHost(h->sticky_host);
    delete h;
    return DRP_OK;
}

extern "C" int drp_set_epsilon(uint64_t handle, float epsilon) {
    BvhHandle* h = drp_lookup(handle);
    if (!h) { drp_set_error("drp_set_epsilon: unknown handle"); return DRP_ERR_HANDLE; }
    h->eps = epsilon;
    return DRP_OK;
}

extern "C" int drp_bvh_stats(uint64_t handle, drp_bvh_stats_t* out) {
    BvhHandle* h = drp_lookup(handle);
    if (!h || !out) { drp_set_error("drp_bvh_stats: unknown handle"); return DRP_ERR_HANDLE; }
    DeviceGuard guard(h->device);
    DRP_CUDA_CHECK(cudaDeviceSynchronize());
    memset(out, 0, sizeof(*out));
    out->n_tris = h->n_tris;
    out->n_nodes = h->n_nodes;
    out->node_bytes = h->n_nodes * 64;
    out->tri_bytes = h->n_tris * 48;
    uint32_t ob[12];
    DRP_CUDA_CHECK(cudaMemcpy(ob, h->bounds, sizeof(ob), cudaMemcpyDeviceToHost));
    for (int k = 0; k < 6; ++k) out->bounds[k] = ord2f(ob[k]);
    DRP_CUDA_CHECK(cudaMemcpy(&out->sah_cost, h->sah, sizeof(float), cudaMemcpyDeviceToHost));
    {
        const int64_t used = h->n_nodes_used;
        std::vector<float4> nodes((size_t)used * CW_NODE_F4);
        DRP_CUDA_CHECK(cudaMemcpy(nodes.data(), h->nodes, sizeof(float4) * nodes.size(), cudaMemcpyDeviceToHost));
        out->n_nodes = used;
        out->node_bytes = used * 16 * CW_NODE_F4;
        int64_t leaves = 0;
        // depth by walking child links
        std::vector<std::pair<int, int>> st;
        st.push_back({0, 1});
        int md = 0;
        while (!st.empty()) {
            auto [ni, depth] = st.back();
            st.pop_back();
            md = depth > md ? depth : md;
            const float4* p = nodes.data() + (size_t)ni * CW_NODE_F4;
            uint32_t imask = f2u(p[0].w) >> 24;
            int base = f2i(p[1].x), rank = 0;
            for (int sl = 0; sl < 8; ++sl) {
                const bool leaf = cw_slot_kind(p, sl) > 0;
                if (imask & (1u << sl)) st.push_back({base + rank++, depth + 1});
                else if (leaf) ++leaves;
            }
        }
        out->n_leaves = leaves;
        out->max_depth = md;
    }
    return DRP_OK;
}

// ---- trace ------------------------------------------------------------------------------------------------------
extern "C" int drp_trace(uint64_t handle, const float* rays_o, const float* rays_d, float* out_t, int32_t* out_i, float t_far,
                         int64_t n_rays, void* stream) {
    BvhHandle* h = drp_lookup(handle);
    if (!h) { drp_set_error("drp_trace: unknown handle"); return DRP_ERR_HANDLE; }
    if (n_rays < 0 || (n_rays > 0 && (!rays_o || !rays_d || !out_t || !out_i))) {
        drp_set_error("drp_trace: invalid argument");
        return DRP_ERR_INVALID;
    }
    if (n_rays == 0) return DRP_OK;
    DeviceGuard guard(h->device);
    if (!guard.ok) { drp_set_error("drp_trace: cannot select device"); return DRP_ERR_CUDA; }
    if (int rc = drp_check_sticky(h, "drp_trace")) return rc;
    return drp_trace_wide_persistent(h, rays_o, rays_d, out_t, out_i, t_far, n_rays, (cudaStream_t)stream);
}

__global__ void __launch_bounds__(128) k_bruteforce(const float* __restrict__ verts, const int32_t* __restrict__ tris, int64_t n_tris,
                                                    const float* __restrict__ ro, const float* __restrict__ rd, float* __restrict__ out_t,
                                                    int32_t* __restrict__ out_i, float t_far, float eps, int64_t n) {
    int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    RayHit h = bruteforce_one(verts, tris, n_tris, v3(ro[3 * r], ro[3 * r + 1], ro[3 * r + 2]), v3(rd[3 * r], rd[3 * r + 1], rd[3 * r + 2]), t_far, eps);
    out_t[r] = h.t;
    out_i[r] = h.id;
}

extern "C" int drp_trace_bruteforce(const float* verts, const int32_t* tris, int64_t n_tris, const float* rays_o, const float* rays_d,
                                    float* out_t, int32_t* out_i, float t_far, float epsilon, int64_t n_rays, void* stream) {
    if (n_rays <= 0) return DRP_OK;
    k_bruteforce<<<(unsigned)((n_rays + 127) / 128), 128, 0, (cudaStream_t)stream>>>(verts, tris, n_tris, rays_o, rays_d, out_t, out_i, t_far, epsilon, n_rays);
    DRP_CUDA_CHECK(cudaGetLastError());
    return DRP_OK;
}
