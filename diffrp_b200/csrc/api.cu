// api.cu -- C ABI of libdiffrp_b200.so: handle table, LBVH build, closest-hit trace.
// See include/diffrp_b200.h for the contract and the reference call sites each entry point replaces.
#include <cub/device/device_radix_sort.cuh>
#include <mutex>
#include <unordered_map>
#include <vector>
#include <algorithm>
#include <cstdio>
#include "internal.h"
#include "common.cuh"
#include "lbvh.cuh"
#include "traverse.cuh"
#include "cwbvh.cuh"
#include "instance_level.h"
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

// Build scratch: stream-ordered allocations that are freed (stream-ordered as well) when the builder returns -- on the error paths too, so a
// failed build (out of memory half way through the list) leaves nothing behind in the pool.
struct BuildScratch {
    cudaStream_t s;
    std::vector<void*> blocks;
    explicit BuildScratch(cudaStream_t stream) : s(stream) {}
    template <typename T>
    cudaError_t get(T** p, size_t count) { return get_bytes((void**)p, sizeof(T) * (count ? count : 1)); }
    cudaError_t get_bytes(void** p, size_t bytes) {
        const cudaError_t e = cudaMallocAsync(p, bytes ? bytes : 1, s);
        if (e == cudaSuccess) blocks.push_back(*p);
        return e;
    }
    cudaError_t release() {
        cudaError_t first = cudaSuccess;
        for (void* p : blocks) { const cudaError_t e = cudaFreeAsync(p, s); if (first == cudaSuccess) first = e; }
        blocks.clear();
        return first;
    }
    ~BuildScratch() { release(); }
};

// Build the wide hierarchy over (verts, tris[0 .. n_tris)) into *h (not registered): device, eps and stack_cap are set by the caller.
int drp_build_structure(BvhHandle* h, const float* verts, const int32_t* tris, int64_t n_tris, cudaStream_t s) {
    const int device = h->device;
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
    h->n_tris = n_tris;
    h->n_nodes = n > 1 ? n - 1 : 1;
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
    BuildScratch scratch(s);
    DRP_CUDA_CHECK(scratch.get(&b.prim_lo, nn));
    DRP_CUDA_CHECK(scratch.get(&b.prim_hi, nn));
    DRP_CUDA_CHECK(scratch.get(&keys_in, nn));
    DRP_CUDA_CHECK(scratch.get(&vals_in, nn));
    DRP_CUDA_CHECK(scratch.get(&b.keys, nn));
    DRP_CUDA_CHECK(scratch.get(&b.vals, nn));
    DRP_CUDA_CHECK(scratch.get(&b.left, nn));
    DRP_CUDA_CHECK(scratch.get(&b.right, nn));
    DRP_CUDA_CHECK(scratch.get(&b.parent, 2 * nn));
    DRP_CUDA_CHECK(scratch.get(&b.range_first, nn));
    DRP_CUDA_CHECK(scratch.get(&b.range_last, nn));
    DRP_CUDA_CHECK(scratch.get(&b.box_lo, 2 * nn));
    DRP_CUDA_CHECK(scratch.get(&b.box_hi, 2 * nn));
    DRP_CUDA_CHECK(scratch.get(&b.arrive, nn));
    DRP_CUDA_CHECK(scratch.get(&b.collapsed, nn));
    DRP_CUDA_CHECK(scratch.get(&b.count, 2 * nn));
    // optimal (DP) collapse tables; the greedy collapse measured 2 % slower traversal and 40 % more nodes (profiles/README.md)
    DRP_CUDA_CHECK(scratch.get(&b.dp_cost, 16 * nn));
    DRP_CUDA_CHECK(scratch.get(&b.dp_dec, 16 * nn));
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
        DRP_CUDA_CHECK(scratch.get_bytes(&sort_tmp, sort_bytes));
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
        DRP_CUDA_CHECK(scratch.get(&cw_work, (size_t)h->n_nodes));
        DRP_CUDA_CHECK(scratch.get(&cw_counters, 160));
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
            if (n < 2 || levels_done >= 128 || hc[8 + levels_done] >= hc[9 + levels_done]) {
                // level L occupies nodes [hc[8 + L], hc[9 + L]) (breadth-first allocation): kept for the bottom-up refit
                h->level_begin.clear();
                if (n < 2) { h->level_begin = {0, 1}; }
                else {
                    for (int L = 0; L < 150 && hc[8 + L] < hc[9 + L]; ++L) h->level_begin.push_back(hc[8 + L]);
                    h->level_begin.push_back(hc[0]);
                }
                break;
            }
        }
    }
    DRP_CUDA_CHECK(cudaGetLastError());
    const double t_launch = now();
    DRP_CUDA_CHECK(scratch.release());
    if (timing) fprintf(stderr, "[diffrp_b200] build n=%d: alloc %.2f ms, launch+collapse-sync %.2f ms, free %.2f ms\n", n, t_alloc - t_begin,
                        t_launch - t_alloc, now() - t_launch);
    return DRP_OK;
}

void drp_register_handle(BvhHandle* h, uint64_t* out_handle) {
    std::lock_guard<std::mutex> lk(g_mutex);
    *out_handle = g_next_handle++;
    g_handles[*out_handle] = h;
}

BvhHandle* drp_new_handle(int device) {
    BvhHandle* h = new BvhHandle();
    h->device = device;
    h->stack_cap = CW_STACK;
    return h;
}

void drp_destroy_handle(BvhHandle* h) {   // device selected by the caller
    drp_free_workspace(h);
    cudaFreeAsync(h->nodes, 0); cudaFreeAsync(h->packed, 0); cudaFreeAsync(h->node_box, 0);
    cudaFree(h->bounds); cudaFree(h->sah); cudaFreeHost(h->sticky_host);
    delete h;
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
    BvhHandle* h = drp_new_handle(device);
    const int rc = drp_build_structure(h, verts, tris, n_tris, (cudaStream_t)stream);
    if (rc != DRP_OK) { drp_destroy_handle(h); return rc; }
    drp_register_handle(h, out_handle);
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
    drp_destroy_handle(h);
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

// ---- refit and instanced assembly --------------------------------------------------------------------------------------
// The collapse allocates wide nodes breadth-first, so a bottom-up pass is one launch per level, deepest first (BvhHandle::level_begin).
struct RefitJob {
    const float4* src_nodes;   // template hierarchy (== dst for an in-place refit)
    const float4* src_tris;
    float4* dst_nodes;
    float4* dst_tris;
    float4* node_box;
    const float* verts;
    const int32_t* tris;
    const uint32_t* bounds;    // scene bounds (abs_pad)
    // instances sharing this template: instance q uses block offsets inst_node_off[q] / inst_tri_off[q] / inst_prim_off[q]; null = one in-place job
    const int* inst_node_off;
    const int* inst_tri_off;
    const int* inst_prim_off;
    int n_inst;
};
__global__ void __launch_bounds__(128) k_cw_refit_level(RefitJob j, int begin, int end) {
    const int per = end - begin;
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (int64_t)per * j.n_inst) return;
    const int q = (int)(t / per), ni = begin + (int)(t - (int64_t)q * per);
    LbvhBuild b;
    b.bounds = const_cast<uint32_t*>(j.bounds);
    const float abs_pad = lbvh_abs_pad(b);
    const int node_off = j.inst_node_off ? j.inst_node_off[q] : 0, tri_off = j.inst_tri_off ? j.inst_tri_off[q] : 0,
              prim_off = j.inst_prim_off ? j.inst_prim_off[q] : 0;
    cw_refit_node(j.src_nodes + CW_NODE_F4 * (int64_t)ni, j.src_tris, j.dst_nodes, j.dst_tris, j.node_box, ni + node_off, node_off, tri_off, prim_off,
                  j.verts, j.tris, abs_pad);
}

static int scene_bounds(BvhHandle* h, const float* verts, const int32_t* tris, int64_t n_tris, cudaStream_t s) {
    LbvhBuild b;
    memset(&b, 0, sizeof(b));
    b.verts = verts; b.tris = tris; b.n = (int)n_tris; b.bounds = h->bounds;   // prim_lo == null: bounds only
    k_init_bounds<<<1, 32, 0, s>>>(h->bounds);
    if (n_tris > 0) k_prim_bounds<<<(int)std::min<int64_t>((n_tris + 255) / 256, 148 * 8), 256, 0, s>>>(b);
    DRP_CUDA_CHECK(cudaGetLastError());
    return DRP_OK;
}

static int ensure_node_box(BvhHandle* h, cudaStream_t s) {
    if (!h->node_box) DRP_CUDA_CHECK(cudaMallocAsync((void**)&h->node_box, sizeof(float4) * 2 * (size_t)std::max<int64_t>(h->n_nodes_used, 1), s));
    return DRP_OK;
}

extern "C" int drp_refit(uint64_t handle, const float* verts, const int32_t* tris, int64_t n_verts, int64_t n_tris, void* stream) {
    BvhHandle* h = drp_lookup(handle);
    if (!h) { drp_set_error("drp_refit: unknown handle"); return DRP_ERR_HANDLE; }
    if (n_tris != h->n_tris || n_verts < 0 || (n_tris > 0 && (!verts || !tris))) {
        drp_set_error("drp_refit: the triangle count must match the built structure (refit keeps the topology: same index array, new vertex positions)");
        return DRP_ERR_INVALID;
    }
    if (int rc = drp_check_sticky(h, "drp_refit")) return rc;
    if (n_tris < 2 || h->level_begin.size() < 2) { drp_set_error("drp_refit: structures over fewer than 2 triangles are rebuilt, not refitted"); return DRP_ERR_INVALID; }
    DeviceGuard guard(h->device);
    if (!guard.ok) { drp_set_error("drp_refit: cannot select device"); return DRP_ERR_CUDA; }
    cudaStream_t s = (cudaStream_t)stream;
    if (int rc = ensure_node_box(h, s)) return rc;
    if (int rc = scene_bounds(h, verts, tris, n_tris, s)) return rc;
    RefitJob j;
    memset(&j, 0, sizeof(j));
    j.src_nodes = h->nodes; j.src_tris = h->packed; j.dst_nodes = h->nodes; j.dst_tris = h->packed; j.node_box = h->node_box;
    j.verts = verts; j.tris = tris; j.bounds = h->bounds; j.n_inst = 1;
    for (int L = (int)h->level_begin.size() - 2; L >= 0; --L) {
        const int begin = h->level_begin[L], end = h->level_begin[L + 1];
        if (end > begin) k_cw_refit_level<<<(end - begin + 127) / 128, 128, 0, s>>>(j, begin, end);
    }
    DRP_CUDA_CHECK(cudaGetLastError());
    if (h->ws) drp_invalidate_scene_box(h);
    return DRP_OK;
}

__global__ void k_copy_roots(float4* nodes, const int2* copies, const int* inst_node_off, int n) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n * CW_NODE_F4) return;
    const int c = t / CW_NODE_F4, f = t - c * CW_NODE_F4;
    nodes[CW_NODE_F4 * (int64_t)copies[c].y + f] = nodes[CW_NODE_F4 * (int64_t)inst_node_off[copies[c].x] + f];
}
__global__ void k_gather_root_boxes(const float4* node_box, const int* inst_node_off, int n, float4* out) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n) return;
    out[2 * q] = node_box[2 * (int64_t)inst_node_off[q]];
    out[2 * q + 1] = node_box[2 * (int64_t)inst_node_off[q] + 1];
}

extern "C" int drp_build_instanced(const float* verts, const int32_t* tris, int64_t n_verts, int64_t n_tris, const int64_t* inst_first_tri,
                                   const int32_t* inst_mesh, int64_t n_inst, int device, void* stream, uint64_t* out_handle) {
    if (!out_handle || !verts || !tris || !inst_first_tri || !inst_mesh || n_inst < 1 || n_tris < 2 || n_verts < 0) {
        drp_set_error("drp_build_instanced: invalid argument");
        return DRP_ERR_INVALID;
    }
    if (n_tris >= (int64_t(1) << 27)) { drp_set_error("drp_build_instanced: more than 2^27 triangles are not supported"); return DRP_ERR_INVALID; }
    if (inst_first_tri[0] != 0 || inst_first_tri[n_inst] != n_tris) { drp_set_error("drp_build_instanced: instance ranges must tile [0, n_tris)"); return DRP_ERR_INVALID; }
    // representative (first) instance of every mesh; all instances of a mesh must have its triangle count
    std::unordered_map<int, int> rep;
    for (int64_t q = 0; q < n_inst; ++q) {
        const int64_t cnt = inst_first_tri[q + 1] - inst_first_tri[q];
        if (cnt < 2) { drp_set_error("drp_build_instanced: every instance needs at least 2 triangles"); return DRP_ERR_INVALID; }
        auto it = rep.find(inst_mesh[q]);
        if (it == rep.end()) rep[inst_mesh[q]] = (int)q;
        else if (inst_first_tri[it->second + 1] - inst_first_tri[it->second] != cnt) {
            drp_set_error("drp_build_instanced: instances of one mesh must have the same triangle count (and connectivity)");
            return DRP_ERR_INVALID;
        }
    }
    DeviceGuard guard(device);
    if (!guard.ok) { drp_set_error("drp_build_instanced: cannot select device"); return DRP_ERR_CUDA; }
    cudaStream_t s = (cudaStream_t)stream;
    // 1. one template hierarchy per mesh, built over its representative instance's world-space triangles (local primitive ids)
    std::unordered_map<int, BvhHandle*> blas;
    auto cleanup = [&]() { for (auto& kv : blas) drp_destroy_handle(kv.second); };
    for (auto& kv : rep) {
        BvhHandle* t = drp_new_handle(device);
        const int64_t first = inst_first_tri[kv.second], cnt = inst_first_tri[kv.second + 1] - first;
        const int rc = drp_build_structure(t, verts, tris + 3 * first, cnt, s);
        blas[kv.first] = t;
        if (rc != DRP_OK) { cleanup(); return rc; }
    }
    // 2. layout: [instance level (<= 2 n_inst nodes)] [block of instance 0] [block of instance 1] ...; triangles in instance order
    BvhHandle* h = drp_new_handle(device);
    std::vector<int> node_off(n_inst), tri_off(n_inst), prim_off(n_inst);
    const int64_t tlas_cap = 2 * n_inst + 8;
    int64_t nodes_total = tlas_cap;
    for (int64_t q = 0; q < n_inst; ++q) {
        node_off[q] = (int)nodes_total; tri_off[q] = (int)inst_first_tri[q]; prim_off[q] = (int)inst_first_tri[q];
        nodes_total += blas[inst_mesh[q]]->n_nodes_used;
    }
    h->n_tris = n_tris; h->n_nodes = nodes_total; h->n_nodes_used = nodes_total;
    int *d_node_off = nullptr, *d_tri_off = nullptr, *d_prim_off = nullptr;
    float4* d_root_boxes = nullptr;
    int2* d_copies = nullptr;
    auto fail = [&](int rc) { cleanup(); drp_destroy_handle(h); cudaFreeAsync(d_node_off, s); cudaFreeAsync(d_tri_off, s); cudaFreeAsync(d_prim_off, s);
                              cudaFreeAsync(d_root_boxes, s); cudaFreeAsync(d_copies, s); return rc; };
#define DRP_TRY(expr) do { cudaError_t _e = (expr); if (_e != cudaSuccess) { drp_set_error(std::string(#expr) + ": " + cudaGetErrorString(_e)); return fail(DRP_ERR_CUDA); } } while (0)
    DRP_TRY(cudaMallocAsync((void**)&h->nodes, sizeof(float4) * CW_NODE_F4 * (size_t)nodes_total, s));
    DRP_TRY(cudaMallocAsync((void**)&h->packed, sizeof(float4) * 3 * (size_t)n_tris, s));
    DRP_TRY(cudaMallocAsync((void**)&h->node_box, sizeof(float4) * 2 * (size_t)nodes_total, s));
    DRP_TRY(cudaMalloc((void**)&h->bounds, sizeof(uint32_t) * 12));
    DRP_TRY(cudaMalloc((void**)&h->sah, sizeof(float)));
    DRP_TRY(cudaMemsetAsync(h->sah, 0, sizeof(float), s));
    DRP_TRY(cudaHostAlloc((void**)&h->sticky_host, sizeof(int), cudaHostAllocMapped));
    *h->sticky_host = 0;
    DRP_TRY(cudaHostGetDevicePointer((void**)&h->sticky_dev, h->sticky_host, 0));
    DRP_TRY(cudaMemsetAsync(h->nodes, 0, sizeof(float4) * CW_NODE_F4 * (size_t)tlas_cap, s));
    if (scene_bounds(h, verts, tris, n_tris, s) != DRP_OK) return fail(DRP_ERR_CUDA);
    DRP_TRY(cudaMallocAsync((void**)&d_node_off, sizeof(int) * n_inst, s));
    DRP_TRY(cudaMallocAsync((void**)&d_tri_off, sizeof(int) * n_inst, s));
    DRP_TRY(cudaMallocAsync((void**)&d_prim_off, sizeof(int) * n_inst, s));
    DRP_TRY(cudaMallocAsync((void**)&d_root_boxes, sizeof(float4) * 2 * n_inst, s));
    // 3. replicate + refit, mesh by mesh, deepest level first: one thread per (instance, node).  Instances of a mesh are processed together,
    //    so the offset tables are uploaded grouped by mesh.
    std::vector<int> order;   // instances grouped by mesh
    std::vector<std::pair<int, std::pair<int, int>>> groups;   // (mesh, [first, count) in `order`)
    for (auto& kv : rep) {
        const int first = (int)order.size();
        for (int64_t q = 0; q < n_inst; ++q) if (inst_mesh[q] == kv.first) order.push_back((int)q);
        groups.push_back({kv.first, {first, (int)order.size() - first}});
    }
    std::vector<int> g_node(n_inst), g_tri(n_inst), g_prim(n_inst);
    for (int64_t k = 0; k < n_inst; ++k) {
        // block offsets are relative to the template's own indices: template node j -> node_off + j
        g_node[k] = node_off[order[k]]; g_tri[k] = tri_off[order[k]]; g_prim[k] = prim_off[order[k]];
    }
    DRP_TRY(cudaMemcpyAsync(d_node_off, g_node.data(), sizeof(int) * n_inst, cudaMemcpyHostToDevice, s));
    DRP_TRY(cudaMemcpyAsync(d_tri_off, g_tri.data(), sizeof(int) * n_inst, cudaMemcpyHostToDevice, s));
    DRP_TRY(cudaMemcpyAsync(d_prim_off, g_prim.data(), sizeof(int) * n_inst, cudaMemcpyHostToDevice, s));
    for (auto& grp : groups) {
        BvhHandle* t = blas[grp.first];
        RefitJob j;
        memset(&j, 0, sizeof(j));
        j.src_nodes = t->nodes; j.src_tris = t->packed; j.dst_nodes = h->nodes; j.dst_tris = h->packed; j.node_box = h->node_box;
        j.verts = verts; j.tris = tris; j.bounds = h->bounds;
        j.inst_node_off = d_node_off + grp.second.first; j.inst_tri_off = d_tri_off + grp.second.first; j.inst_prim_off = d_prim_off + grp.second.first;
        j.n_inst = grp.second.second;
        for (int L = (int)t->level_begin.size() - 2; L >= 0; --L) {
            const int begin = t->level_begin[L], end = t->level_begin[L + 1];
            const int64_t threads = (int64_t)(end - begin) * j.n_inst;
            if (threads > 0) k_cw_refit_level<<<(unsigned)((threads + 127) / 128), 128, 0, s>>>(j, begin, end);
        }
    }
    // 4. instance level on the host over the refitted root boxes (grouped order -> instance order)
    k_gather_root_boxes<<<(unsigned)((n_inst + 127) / 128), 128, 0, s>>>(h->node_box, d_node_off, (int)n_inst, d_root_boxes);
    std::vector<float4> hb(2 * (size_t)n_inst);
    DRP_TRY(cudaMemcpyAsync(hb.data(), d_root_boxes, sizeof(float4) * 2 * n_inst, cudaMemcpyDeviceToHost, s));
    DRP_TRY(cudaStreamSynchronize(s));
    std::vector<float> lo(3 * (size_t)n_inst), hi(3 * (size_t)n_inst);
    for (int64_t k = 0; k < n_inst; ++k) {
        const int q = order[k];
        lo[3 * (size_t)q] = hb[2 * k].x; lo[3 * (size_t)q + 1] = hb[2 * k].y; lo[3 * (size_t)q + 2] = hb[2 * k].z;
        hi[3 * (size_t)q] = hb[2 * k + 1].x; hi[3 * (size_t)q + 1] = hb[2 * k + 1].y; hi[3 * (size_t)q + 2] = hb[2 * k + 1].z;
    }
    TlasBuilder tb(lo, hi);
    tb.nodes.resize(CW_NODE_F4, make_float4(0, 0, 0, 0));
    std::vector<int> ids(n_inst);
    for (int64_t q = 0; q < n_inst; ++q) ids[q] = (int)q;
    if (n_inst == 1) {   // the root of the instance level has the single instance as its only child
        tlas_single(tb, lo.data(), hi.data());
    } else {
        tb.build(0, ids.data(), (int)n_inst);
    }
    if (tb.allocated > tlas_cap) { drp_set_error("drp_build_instanced: instance level larger than its reservation"); return fail(DRP_ERR_INVALID); }
    tb.nodes.resize((size_t)tb.allocated * CW_NODE_F4, make_float4(0, 0, 0, 0));
    DRP_TRY(cudaMemcpyAsync(h->nodes, tb.nodes.data(), sizeof(float4) * tb.nodes.size(), cudaMemcpyHostToDevice, s));
    std::vector<int2> copies(tb.copies.size());
    for (size_t k = 0; k < copies.size(); ++k) copies[k] = make_int2(tb.copies[k].first, tb.copies[k].second);
    DRP_TRY(cudaMallocAsync((void**)&d_copies, sizeof(int2) * copies.size(), s));
    DRP_TRY(cudaMemcpyAsync(d_copies, copies.data(), sizeof(int2) * copies.size(), cudaMemcpyHostToDevice, s));
    // the copy kernel indexes the offset table by instance id: upload it in instance order
    DRP_TRY(cudaMemcpyAsync(d_node_off, node_off.data(), sizeof(int) * n_inst, cudaMemcpyHostToDevice, s));
    k_copy_roots<<<(unsigned)((tb.copies.size() * CW_NODE_F4 + 127) / 128), 128, 0, s>>>(h->nodes, d_copies, d_node_off, (int)tb.copies.size());
    DRP_TRY(cudaGetLastError());
    DRP_TRY(cudaStreamSynchronize(s));   // host vectors above are pageable sources of the async copies
#undef DRP_TRY
    cudaFreeAsync(d_node_off, s); cudaFreeAsync(d_tri_off, s); cudaFreeAsync(d_prim_off, s); cudaFreeAsync(d_root_boxes, s); cudaFreeAsync(d_copies, s);
    cleanup();
    cudaFreeAsync(h->node_box, s);   // only the assembly needed the exact boxes (instanced structures are rebuilt, not refitted)
    h->node_box = nullptr;
    h->level_begin.clear();   // a mixed structure: refit goes through a rebuild
    drp_register_handle(h, out_handle);
    if (g_drp_log_level >= 4) fprintf(stderr, "[diffrp_b200] instanced build: %lld instances of %zu meshes, %lld triangles, %lld nodes (instance level %d)\n",
                                      (long long)n_inst, rep.size(), (long long)n_tris, (long long)nodes_total, tb.allocated);
    return DRP_OK;
}
