// wavefront.cu -- fused wavefront path tracer: drp_render / drp_finalize / drp_render_stats / profiling.
//
// Replaces the Python section x bounce loop of PathTracingSession.trace_rays with the built-in sampler_brdf
// (diffrp/rendering/path_tracing.py:250-279, 310-352).  Per batch of samples and per bounce, two persistent kernels:
//   k_extend_cw : closest hit for every live ray over the compressed 8-wide BVH (cwbvh.cuh): one ray per lane, dynamic ray
//                 fetch by warp ballot, triangle postponing; bounce 0 generates the primary ray from the ray index instead
//                 of reading it; writes (t, id).  (k_extend: the simple one-warp-batch kernel, kept for the binary layout.)
//   k_shade     : surface attributes + env lookup + BRDF sample + fp32 accumulation (RED.ADD.F32x4) + next ray, appended to
//                 the output queue through a warp-aggregated atomic (stream compaction fused into the shade kernel)
// Ray state lives in HBM as float4 SoA queues (coalesced 128-bit accesses):
//   q_a[k] = (o.x, o.y, o.z, d.x)   q_b[k] = (d.y, d.z, bits(ray index), 0)   q_t[k] = (T.r, T.g, T.b, 0)   hit[k] = (t, bits(id))
// Compile-time / environment switches exist for every design alternative that was measured (see profiles/README.md and
// drp_build_config()); the defaults are the fastest measured configuration.
#include <vector>
#include <algorithm>
#include <string>
#include <cstring>
#include <cstdio>
#include <cstdlib>
#include "internal.h"
#include "common.cuh"
#include "lbvh.cuh"
#include "traverse.cuh"
#include "cwbvh.cuh"
#include "shade.cuh"

#define WF_BLOCK 128
#ifndef WF_FETCH
#define WF_FETCH 128  // rays fetched per warp per atomic
#endif

struct RenderWorkspace {
    int64_t capacity = 0;  // rays
    float4 *qa[2] = {nullptr, nullptr}, *qb[2] = {nullptr, nullptr}, *qt[2] = {nullptr, nullptr};
    float2* hit = nullptr;
    int* hit_list = nullptr;   // queue slots of the rays that hit / missed in the last extend (partitioned shading)
    int* miss_list = nullptr;
    int* counters = nullptr;  // [0..D] live counts per bounce, [64..64+2D) fetch cursors
    int n_counters = 0;
    drp_material_t* d_mats = nullptr;
    int mats_capacity = 0;
    std::vector<unsigned char> mats_host_copy;
    unsigned long long* d_traced = nullptr;  // device: live rays traced by the last call (sum over batches and bounces)
    int sm_count = 0;
    int grid_extend[2] = {0, 0}, grid_shade[2] = {0, 0};
    bool have_box = false;
    float box_lo[3], box_hi[3];
    cudaStream_t streams[2] = {nullptr, nullptr};  // DRP_OVERLAP: one internal stream per ray group
    cudaEvent_t sync_events[4] = {nullptr, nullptr, nullptr, nullptr};
    // optional per-kernel timing (drp_set_profiling)
    bool profiling = false;
    struct Span { cudaEvent_t a, b; int kind; int bounce; unsigned long long* traced_slot; };
    std::vector<Span> spans;
    std::vector<cudaEvent_t> event_pool;
    unsigned long long* d_live = nullptr;  // device: per-launch live-ray counts while profiling
    int live_capacity = 0, live_used = 0;
};

static cudaEvent_t wf_event(RenderWorkspace* ws) {
    if (!ws->event_pool.empty()) { cudaEvent_t e = ws->event_pool.back(); ws->event_pool.pop_back(); return e; }
    cudaEvent_t e;
    cudaEventCreate(&e);
    return e;
}

__global__ void k_record_live(const int* __restrict__ count_ptr, unsigned long long fixed, unsigned long long* __restrict__ slot) {
    *slot = count_ptr ? (unsigned long long)*count_ptr : fixed;
}

struct WfConst {
    drp_scene_t scene;        // device pointers (materials -> device copy)
    drp_render_params_t p;
    const float4* nodes;
    const float4* tris;
    float box_lo[3], box_hi[3];
    float eps;
    int HW;             // pixels rendered per sample (the tile's pixel count when tile sharding)
    int tx0, ty0, tw;   // tile origin and width (tw == frame width, origin 0 for a full frame)
    int tiled8x4;       // primary rays enumerated in 8x4 pixel blocks (rectangle width % 8 == 0 and height % 4 == 0)
    int sample_minor;   // > 0: primary rays enumerated pixel-major / sample-minor with this many samples per batch (see load_ray)
    int64_t R;          // rays in this batch
    int64_t R_total;    // rays of the whole call (replay indexing)
    int64_t ray_base;   // index (within the call) of this launch group's first ray: s_local * HW + pixel of queue slot 0 at bounce 0
    float* accum;
    int* flags;
    uint32_t cw_bias;   // 0x47000000 (cwbvh.cuh: cw_byte_biased), passed through the constant bank
};

template <bool PRIMARY>
__device__ __forceinline__ void load_ray(const WfConst& c, const float4* __restrict__ qa, const float4* __restrict__ qb, int k, Vec3& o,
                                         Vec3& d, int& ray_index) {
    if (PRIMARY) {
        // Queue slot -> (sample, pixel).  Scanline order by default; DRP_PRIMARY_ORDER=tiled enumerates 8x4-pixel blocks so that a
        // warp starts on a compact screen tile instead of a 32x1 strip (measured: no gain, see drp_render).
        // The ray index (RNG key, replay index, accumulator row) is the reference's s * HW + y * W + x either way.
        const int kk = k + (int)c.ray_base;
        int s = kk / c.HW;
        int pix = kk - s * c.HW;
        if (c.sample_minor > 0) {  // the samples of one pixel sit next to each other in the queue: they hit the same triangles / texels
            const int j = kk - (int)c.ray_base;               // ray_base is a multiple of HW in this mode
            pix = j / c.sample_minor;
            s = (int)(c.ray_base / c.HW) + (j - pix * c.sample_minor);
        }
        if (c.tiled8x4) {
            const int blocks_x = c.tw >> 3;
            const int blk = pix >> 5, in = pix & 31;
            const int by = blk / blocks_x, bx = blk - by * blocks_x;
            pix = ((by << 2) + (in >> 3)) * c.tw + (bx << 3) + (in & 7);
        }
        ray_index = s * c.HW + pix;
        int y = pix / c.tw, x = pix - y * c.tw;
        x += c.tx0; y += c.ty0;
        float gx = __ldg(c.p.ndc_x + x) + __ldg(c.p.jitter_x + s);
        float gy = __ldg(c.p.ndc_y + y) + __ldg(c.p.jitter_y + s);
        gen_primary_ray(c.p.inv_vp, c.p.cam_pos, c.p.t_near, gx, gy, o, d);
    } else {
        float4 a = __ldg(qa + k), b = __ldg(qb + k);
        o = v3(a.x, a.y, a.z);
        d = v3(a.w, b.x, b.y);
        ray_index = __float_as_int(b.z);
    }
}

template <bool PRIMARY, bool WIDE>
__global__ void __launch_bounds__(WF_BLOCK) k_extend(const __grid_constant__ WfConst c, const float4* __restrict__ qa, const float4* __restrict__ qb,
                                                     float2* __restrict__ hit, const int* __restrict__ count_ptr, int* __restrict__ cursor) {
    const int count = PRIMARY ? (int)c.R : *count_ptr;
    const int lane = threadIdx.x & 31;
    for (;;) {
        int base = 0;
        if (lane == 0) base = atomicAdd(cursor, WF_FETCH);
        base = __shfl_sync(0xffffffffu, base, 0);
        if (base >= count) break;
#pragma unroll 1
        for (int j = 0; j < WF_FETCH; j += 32) {
            int k = base + j + lane;
            if (k < count) {
                Vec3 o, d;
                int ri;
                load_ray<PRIMARY>(c, qa, qb, k, o, d, ri);
                bool overflow = false;
                RayHit h = WIDE ? cw_trace_one(c.nodes, c.tris, o, d, c.p.t_far, c.eps, overflow) : trace_one(c.nodes, c.tris, o, d, c.p.t_far, c.eps, overflow);
                hit[k] = make_float2(h.t, __int_as_float(h.id));
                if (overflow) atomicAdd(&c.flags[0], 1);
            }
        }
    }
}

// ---- persistent-thread extend over the wide layout -----------------------------------------------------------------
// Every lane owns one ray and walks the 8-wide hierarchy with the (node group, triangle group) state of cwbvh.cuh.
//  * dynamic ray fetch (Aila & Laine 2009 / Ylitie et al. 2017): a lane whose ray terminated does not wait for the slowest
//    ray of its warp -- when the accumulated number of idle lane-iterations exceeds CWK_NW the warp leaves the traversal
//    loop, votes (ballot) which lanes need work and refills them from a warp-private chunk of the queue (one atomic per
//    CWK_CHUNK rays);
//  * triangle postponing: a triangle group is pushed back on the stack when fewer than CWK_POSTPONE of the warp's active
//    lanes have triangles to test, so that the warp stays in the node phase.
// Scheduling only: the closest hit found is the exhaustive one whatever the order (min t, then min id).
#ifndef CWK_SMEM_STACK
#define CWK_SMEM_STACK 0
#endif
#ifndef CWK_PREFETCH
#define CWK_PREFETCH 0
#endif
#ifndef CWK_CHUNK
#define CWK_CHUNK 64   // B200 sweep (profiles/README.md): 32-128 within 1 %, 256 -4 %, 1024 -35 %
#endif
#ifndef CWK_ND
#define CWK_ND 2
#endif
#ifndef CWK_NW
#define CWK_NW 8
#endif
#ifndef CWK_LD256
#define CWK_LD256 1   // node fetch by 256-bit loads (needs the 96-byte node stride, DRP_CW_NODE96)
#endif
__device__ __forceinline__ void cw_ld256(const float4* p, float4& a, float4& b) {
    asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w) : "l"(p));
}
#ifndef CWK_TRI_PIPE
#define CWK_TRI_PIPE 0   // triangle records requested one test ahead: written after the round's GPU budget was spent, not measured yet
#endif
#ifndef CWK_LUT
#define CWK_LUT 1     // node format 2: hit-byte expansion by shared-memory tables (1), arithmetic (0), tables for primary rays only (2)
#endif
#ifndef CWK_POSTPONE
#define CWK_POSTPONE 0.2f
#ifndef CWK_SHARE
#define CWK_SHARE 0   // warp-shared triangle tests (see the triangle phase of k_extend_cw).  A/B on B200 (config 3, extend ms per step):
                      // per-lane loop 5.31, shared through shuffles + __fns 7.00, shared through shared memory 5.90 (56 registers, spills) /
                      // 5.54 vs 5.38 at 64 registers: the triangle phase does run at ~6 of 32 lanes (22 % of the issue slots), but dealing
                      // the tests out costs as much as it saves
#endif
#endif

#define SRC_QUEUE 0    // rays from the float4 queues, result to hit[]
#define SRC_PRIMARY 1  // rays generated from the ray index, result to hit[]
#define SRC_AOS 2      // standalone query: rays from two (R,3) arrays, results to separate t / id arrays (Raycaster.query)
struct AosRays {
    const float* ro;
    const float* rd;
    float* out_t;
    int32_t* out_i;
};
// optional hit / miss partition of the finished rays (queue slots), consumed by k_shade<., SHADE_HITS / SHADE_MISSES>
struct Partition {
    int* hit_list;
    int* miss_list;
    int* hit_count;
    int* miss_count;
};

#ifndef DRP_EXTEND_MINBLOCKS
#define DRP_EXTEND_MINBLOCKS 9
#endif
#ifndef DRP_SHADE_CHUNK_PARTITION
#define DRP_SHADE_CHUNK_PARTITION 0   // not measured yet (written after the GPU budget of the round was spent): see k_shade
#endif
#ifndef DRP_SHADE_PIPELINE
#define DRP_SHADE_PIPELINE 0   // software-pipelined hit / index loads: B200 A/B shade 2.07 -> 2.22 ms per step (slower: the kernel is bound by DRAM throughput on random sectors, not by the per-ray latency chain)
#endif
#ifndef DRP_SHADE_MINBLOCKS
#define DRP_SHADE_MINBLOCKS 6   // B200 A/B with the interleaved texels (shade ms per step): 4 -> 2.30, 5 -> 2.12, 6 -> 2.04 (80 registers)
#endif
template <int SRC, bool PART>
__global__ void __launch_bounds__(WF_BLOCK, DRP_EXTEND_MINBLOCKS) k_extend_cw(const __grid_constant__ WfConst c, const float4* __restrict__ qa, const float4* __restrict__ qb,
                                                        float2* __restrict__ hit, const int* __restrict__ count_ptr, int* __restrict__ cursor, AosRays aos,
                                                        Partition part) {
    const int count = SRC == SRC_QUEUE ? *count_ptr : (int)c.R;
    const int lane = threadIdx.x & 31;
    const unsigned lt_mask = (1u << lane) - 1u;
    int chunk_next = 0, chunk_end = 0;  // warp-uniform: private range of queue slots
    bool exhausted = false;             // warp-uniform: the queue has no more chunks
    int k = -1;                         // queue slot of this lane's ray, -1 = idle
    CwRay r;
    float t_best = 0.0f;
    int id_best = 0, sp = 0;
    uint32_t ng_x = 0, ng_y = 0, tg_x = 0, tg_y = 0;
    // Traversal stack.  CWK_SMEM_STACK > 0 keeps the first entries of every thread in shared memory ([entry][thread],
    // conflict-free 64-bit accesses) and spills deeper ones to local memory.  A/B on B200 (profiles/README.md): the
    // all-local stack (L1-resident, no extra branch per push/pop) is ~5 % FASTER than 8 shared entries, so the default is 0.
#if CWK_SMEM_STACK > 0
    __shared__ uint2 s_stack[CWK_SMEM_STACK][WF_BLOCK];
    const int tid = threadIdx.x;
#endif
    uint32_t st_x[CW_STACK - CWK_SMEM_STACK], st_y[CW_STACK - CWK_SMEM_STACK];
    bool overflow = false;
#if DRP_CW_V2
#if CWK_SHARE
#error "CWK_SHARE is implemented for node format 1 only (DRP_CW_V2=0)"
#endif
    // node format 2: the two expansions of the per-slot hit byte as tables (cwbvh.cuh: cw_perm8, cw_spread3x7)
    __shared__ uint8_t s_perm[8 * 256];
    __shared__ uint32_t s_spread[256];
    for (int e = threadIdx.x; e < 8 * 256; e += WF_BLOCK) s_perm[e] = (uint8_t)cw_perm8((uint32_t)e & 0xffu, (uint32_t)e >> 8);
    for (int e = threadIdx.x; e < 256; e += WF_BLOCK) s_spread[e] = cw_spread3x7((uint32_t)e);
    __syncthreads();
    uint32_t tri_base = 0, tri_valid = 0;  // of the node the current triangle group belongs to
    uint32_t oct_row = 0;                  // octinv << 8: this ray's row of s_perm
#endif
#if CWK_SHARE
    __shared__ uint32_t s_item[WF_BLOCK / 32][32];   // work items of the shared triangle phase: (packed triangle index << 5) | owner lane
    __shared__ float2 s_res[WF_BLOCK / 32][32];      // (t, id) posted by the helper of each item
    __shared__ float4 s_ray[WF_BLOCK / 32][32][2];   // origin / direction of every lane's current ray
    const int wid = threadIdx.x >> 5;
#endif
#if CWK_SMEM_STACK == 0
#define CWK_PUSH(X, Y)                                                 \
    do {                                                               \
        if (sp < CW_STACK) { st_x[sp] = (X); st_y[sp] = (Y); ++sp; }   \
        else overflow = true;                                          \
    } while (0)
#define CWK_POP(X, Y) do { --sp; (X) = st_x[sp]; (Y) = st_y[sp]; } while (0)
#else
#define CWK_PUSH(X, Y)                                                                   \
    do {                                                                                 \
        if (sp < CWK_SMEM_STACK) s_stack[sp][tid] = make_uint2((X), (Y));                \
        else if (sp < CW_STACK) { st_x[sp - CWK_SMEM_STACK] = (X); st_y[sp - CWK_SMEM_STACK] = (Y); } \
        else overflow = true;                                                            \
        if (sp < CW_STACK) ++sp;                                                         \
    } while (0)
#define CWK_POP(X, Y)                                                                    \
    do {                                                                                 \
        --sp;                                                                            \
        if (sp < CWK_SMEM_STACK) { uint2 _v = s_stack[sp][tid]; (X) = _v.x; (Y) = _v.y; } \
        else { (X) = st_x[sp - CWK_SMEM_STACK]; (Y) = st_y[sp - CWK_SMEM_STACK]; }       \
    } while (0)
#endif
    for (;;) {
        // ---- refill idle lanes --------------------------------------------------------------------------------
        const unsigned need = __ballot_sync(0xffffffffu, k < 0);
        if (need && !exhausted) {
            int want = __popc(need);
            if (chunk_end - chunk_next < want) {  // take a new chunk (rays left in the old one, at most 31, come first)
                int base = 0;
                if (lane == 0) base = atomicAdd(cursor, CWK_CHUNK);
                base = __shfl_sync(0xffffffffu, base, 0);
                // hand out the tail of the old chunk, then continue in the new one
                const int left = chunk_end - chunk_next;
                const int rank = __popc(need & lt_mask);
                if (k < 0) {
                    int kk = rank < left ? chunk_next + rank : base + (rank - left);
                    k = kk < count ? kk : -1;
                }
                chunk_next = base + (want - left);
                chunk_end = base + CWK_CHUNK;
                if (base >= count) exhausted = true;
            } else {
                if (k < 0) k = chunk_next + __popc(need & lt_mask);
                chunk_next += want;
                if (k >= count) k = -1;
            }
            if (k >= 0 && (need >> lane & 1u)) {  // freshly assigned: load / generate the ray, reset the traversal state
                Vec3 o, d;
                int ri;
                if (SRC == SRC_AOS) {
                    o = v3(__ldg(aos.ro + 3 * (int64_t)k), __ldg(aos.ro + 3 * (int64_t)k + 1), __ldg(aos.ro + 3 * (int64_t)k + 2));
                    d = v3(__ldg(aos.rd + 3 * (int64_t)k), __ldg(aos.rd + 3 * (int64_t)k + 1), __ldg(aos.rd + 3 * (int64_t)k + 2));
                } else {
                    load_ray<SRC == SRC_PRIMARY>(c, qa, qb, k, o, d, ri);
                }
                r = cw_make_ray(o, d, c.cw_bias);
#if DRP_CW_V2
                oct_row = (r.octinv4 & 0xffu) << 8;
#endif
#if CWK_SHARE
                s_ray[wid][lane][0] = make_float4(o.x, o.y, o.z, 0.0f);
                s_ray[wid][lane][1] = make_float4(d.x, d.y, d.z, 0.0f);
#endif
                t_best = c.p.t_far;
                id_best = 0x7fffffff;
                sp = 0;
                ng_x = 0; ng_y = 0x80000000u;
                tg_x = 0; tg_y = 0;
            }
        }
        if (__all_sync(0xffffffffu, k < 0)) break;  // nothing left anywhere in this warp
        // ---- traverse ---------------------------------------------------------------------------------------------
        if (k >= 0) {
            int lost = 0;
            for (;;) {
                if (ng_y > 0x00ffffffu) {
                    const uint32_t hits = ng_y;
                    const int child_bit = 31 - __clz(hits);
                    const uint32_t base = ng_x;
                    ng_y &= ~(1u << child_bit);
                    if (ng_y > 0x00ffffffu) CWK_PUSH(ng_x, ng_y);
                    const uint32_t slot = (uint32_t)(child_bit - 24) ^ (r.octinv4 & 0xffu);
                    const uint32_t rel = __popc(hits & ~(0xffffffffu << slot));
                    const float4* p = c.nodes + CW_NODE_F4 * (int64_t)(base + rel);
#if DRP_CW_NODE96 && CWK_LD256
                    float4 n0, n1, n2, n3, n4;
                    if (CWK_LD256 == 1 || SRC != SRC_PRIMARY) {  // CWK_LD256 == 2: incoherent rays only
                        // three 256-bit loads (LDG.E.ENL2.256): 3 instead of 5 requests per lane through the L1 data pipe
                        float4 pad;
                        cw_ld256(p, n0, n1);
                        cw_ld256(p + 2, n2, n3);
                        cw_ld256(p + 4, n4, pad);
                    } else {
                        n0 = __ldg(p); n1 = __ldg(p + 1); n2 = __ldg(p + 2); n3 = __ldg(p + 3); n4 = __ldg(p + 4);
                    }
#else
                    const float4 n0 = __ldg(p), n1 = __ldg(p + 1), n2 = __ldg(p + 2), n3 = __ldg(p + 3), n4 = __ldg(p + 4);
#endif
#if DRP_CW_V2
                    const uint32_t hit8 = cw_node_hits(r, n0, n2, n3, n4, t_best * DRP_T_GROW);
                    const uint32_t imask = __float_as_uint(n0.w) >> 24;
                    ng_x = __float_as_uint(n1.x);
                    const bool use_lut = CWK_LUT == 1 || (CWK_LUT == 2 && SRC == SRC_PRIMARY);
                    ng_y = ((use_lut ? (uint32_t)s_perm[oct_row + (hit8 & imask)] : cw_perm8(hit8 & imask, oct_row >> 8)) << 24) | imask;
                    tg_x = base + rel;  // a triangle group is (node index, pending bits): tri_base / V are re-read when it comes off the stack
                    tri_base = __float_as_uint(n1.y);
                    tri_valid = __float_as_uint(n1.z);
                    tg_y = (use_lut ? s_spread[hit8 & ~imask] : cw_spread3x7(hit8 & ~imask)) & tri_valid;
#else
                    const uint32_t hitmask = cw_node_hits(r, n0, n1, n2, n3, n4, t_best * DRP_T_GROW);
                    ng_x = __float_as_uint(n1.x);
                    ng_y = (hitmask & 0xff000000u) | (__float_as_uint(n0.w) >> 24);
                    tg_x = __float_as_uint(n1.y);
                    tg_y = hitmask & 0x00ffffffu;
#endif
#if CWK_PREFETCH
                    if (ng_y > 0x00ffffffu) {  // the child visited next is already known: pull its two cache lines towards L1 while triangles are tested
                        const uint32_t nslot = (uint32_t)(31 - __clz(ng_y) - 24) ^ (r.octinv4 & 0xffu);
                        const float4* np = c.nodes + CW_NODE_F4 * (int64_t)(ng_x + __popc(ng_y & ~(0xffffffffu << nslot)));
                        asm volatile("prefetch.global.L1 [%0];" ::"l"(np));
                        asm volatile("prefetch.global.L1 [%0];" ::"l"(np + 4));
                    }
#endif
                } else {
                    tg_x = ng_x; tg_y = ng_y;  // a postponed triangle group came off the stack
                    ng_x = 0; ng_y = 0;
#if DRP_CW_V2
                    const float4 m1 = __ldg(c.nodes + CW_NODE_F4 * (int64_t)tg_x + 1);
                    tri_base = __float_as_uint(m1.y);
                    tri_valid = __float_as_uint(m1.z);
#endif
                }
#if CWK_SHARE
                {   // ---- triangle phase, work-shared across the warp --------------------------------------------------------------
                    // Only ~1/3 of the traversing lanes hold pending triangles at any time and a lane holds up to several (every triangle of
                    // each hit leaf child), so a per-lane loop runs its Moller-Trumbore tests at ~10 of 32 lanes for several rounds.  Instead
                    // the pending (ray, triangle) pairs of all lanes are dealt out as work items through shared memory: up to three per owner
                    // and round, item i to the i-th traversing lane; a helper reads the owner's ray (staged in shared memory when the ray was
                    // fetched), tests one triangle and posts (t, id); the owner merges its results with the (t, id) order of leaf_intersect,
                    // so the outcome does not depend on who tested what.
                    const unsigned act = __activemask();
                    const int nact = __popc(act);
                    const int hr = __popc(act & lt_mask);                            // my rank among the traversing lanes = my item slot
                    for (;;) {
                        unsigned have = __ballot_sync(act, tg_y != 0);
                        if (!have) break;
                        if ((float)__popc(have) < CWK_POSTPONE * (float)nact) {   // too few lanes have triangles: postpone (lanes with stack room)
                            if (tg_y != 0 && sp < CW_STACK) { CWK_PUSH(tg_x, tg_y); tg_y = 0; }
                            have = __ballot_sync(act, tg_y != 0);
                            if (!have) break;
                        }
                        uint32_t m = tg_y;
                        int t0 = -1, t1 = -1, t2 = -1;
                        if (m) { t0 = 31 - __clz(m); m &= ~(1u << t0); }
                        if (m) { t1 = 31 - __clz(m); m &= ~(1u << t1); }
                        if (m) { t2 = 31 - __clz(m); }
                        unsigned b1 = __ballot_sync(act, t1 >= 0), b2 = __ballot_sync(act, t2 >= 0);
                        const int n0 = __popc(have);
                        int n1 = __popc(b1), n2 = __popc(b2);
                        if (n0 + n1 + n2 > nact) { b2 = 0; n2 = 0; if (n0 + n1 > nact) { b1 = 0; n1 = 0; } }   // more items than helpers: next round
                        const bool i0 = t0 >= 0, i1 = (b1 >> lane) & 1u, i2 = (b2 >> lane) & 1u;
                        const int p0 = __popc(have & lt_mask), p1 = n0 + __popc(b1 & lt_mask), p2 = n0 + n1 + __popc(b2 & lt_mask);
                        if (i0) s_item[wid][p0] = ((tg_x + (uint32_t)t0) << 5) | (uint32_t)lane;
                        if (i1) s_item[wid][p1] = ((tg_x + (uint32_t)t1) << 5) | (uint32_t)lane;
                        if (i2) s_item[wid][p2] = ((tg_x + (uint32_t)t2) << 5) | (uint32_t)lane;
                        __syncwarp(act);
                        if (hr < n0 + n1 + n2) {
                            const uint32_t it = s_item[wid][hr];
                            const float4 ro = s_ray[wid][it & 31u][0], rd = s_ray[wid][it & 31u][1];
                            float ht = c.p.t_far;
                            int hid = 0x7fffffff;
                            leaf_intersect(c.tris, (int)(it >> 5), 1, v3(ro.x, ro.y, ro.z), v3(rd.x, rd.y, rd.z), c.eps, ht, hid);
                            s_res[wid][hr] = make_float2(ht, __int_as_float(hid));
                        }
                        __syncwarp(act);
                        if (i0) { const float2 q = s_res[wid][p0]; const int qi = __float_as_int(q.y); if (q.x < t_best || (q.x == t_best && qi < id_best)) { t_best = q.x; id_best = qi; } tg_y &= ~(1u << t0); }
                        if (i1) { const float2 q = s_res[wid][p1]; const int qi = __float_as_int(q.y); if (q.x < t_best || (q.x == t_best && qi < id_best)) { t_best = q.x; id_best = qi; } tg_y &= ~(1u << t1); }
                        if (i2) { const float2 q = s_res[wid][p2]; const int qi = __float_as_int(q.y); if (q.x < t_best || (q.x == t_best && qi < id_best)) { t_best = q.x; id_best = qi; } tg_y &= ~(1u << t2); }
                    }
                }
#elif CWK_TRI_PIPE && DRP_CW_V2
                // Experiment for the next measurement round (profiles/README.md section 2, per-instruction view: 8.8 % of the stall samples of the
                // secondary-bounce kernel wait for the triangle record): the record of the next pending triangle is requested before the
                // current one is tested.  Same tests, same (t, id) order, same postponing rule.
                {
                    const int total_active = __popc(__activemask());
                    if (tg_y != 0) {
                        if ((float)__popc(__activemask()) < CWK_POSTPONE * (float)total_active && sp < CW_STACK) {
                            CWK_PUSH(tg_x, tg_y);  // postpone: too few lanes have triangles
                        } else {
                            int ti = 31 - __clz(tg_y);
                            tg_y &= ~(1u << ti);
                            const float4* tp = c.tris + 3 * (int64_t)cw_tri_index(tri_base, tri_valid, ti);
                            float4 ta = __ldg(tp), tb = __ldg(tp + 1), tc = __ldg(tp + 2);
                            for (;;) {
                                bool more = tg_y != 0;
                                float4 na = ta, nb = tb, nc = tc;
                                if (more) {
                                    if ((float)__popc(__activemask()) < CWK_POSTPONE * (float)total_active && sp < CW_STACK) {
                                        CWK_PUSH(tg_x, tg_y);
                                        more = false;
                                    } else {
                                        ti = 31 - __clz(tg_y);
                                        tg_y &= ~(1u << ti);
                                        tp = c.tris + 3 * (int64_t)cw_tri_index(tri_base, tri_valid, ti);
                                        na = __ldg(tp); nb = __ldg(tp + 1); nc = __ldg(tp + 2);
                                    }
                                }
                                leaf_update(ta, tb, tc, r.o, r.d, c.eps, t_best, id_best);
                                if (!more) break;
                                ta = na; tb = nb; tc = nc;
                            }
                        }
                    }
                }
#else
                const int total_active = __popc(__activemask());
                while (tg_y != 0) {
                    if ((float)__popc(__activemask()) < CWK_POSTPONE * (float)total_active && sp < CW_STACK) {
                        CWK_PUSH(tg_x, tg_y);  // postpone: too few lanes have triangles
                        break;
                    }
                    const int ti = 31 - __clz(tg_y);
                    tg_y &= ~(1u << ti);
#if DRP_CW_V2
                    leaf_intersect(c.tris, cw_tri_index(tri_base, tri_valid, ti), 1, r.o, r.d, c.eps, t_best, id_best);
#else
                    leaf_intersect(c.tris, (int)tg_x + ti, 1, r.o, r.d, c.eps, t_best, id_best);
#endif
                }
#endif
                if (ng_y <= 0x00ffffffu) {
                    if (sp > 0) {
                        CWK_POP(ng_x, ng_y);
                    } else {  // ray finished
                        const bool is_hit = t_best < c.p.t_far;
                        if (SRC == SRC_AOS) {
                            aos.out_t[k] = is_hit ? t_best : c.p.t_far;
                            aos.out_i[k] = is_hit ? id_best : 0;
                        } else {
                            hit[k] = make_float2(is_hit ? t_best : c.p.t_far, __int_as_float(is_hit ? id_best : 0));
                            if (PART) {  // warp-aggregated append of this iteration's finished rays to the hit / miss lists
                                const unsigned fin = __activemask();
                                const unsigned hm = __ballot_sync(fin, is_hit), mm = fin & ~hm;
                                const unsigned mine = is_hit ? hm : mm;
                                const int leader = __ffs(mine) - 1;
                                int pos = 0;
                                if (lane == leader) pos = atomicAdd(is_hit ? part.hit_count : part.miss_count, __popc(mine));
                                pos = __shfl_sync(mine, pos, leader);
                                (is_hit ? part.hit_list : part.miss_list)[pos + __popc(mine & lt_mask)] = k;
                            }
                        }
                        k = -1;
                        break;
                    }
                }
                if (!exhausted) {  // dynamic fetch: leave when enough lane-iterations were lost to idle lanes
                    lost += 32 - __popc(__activemask()) - CWK_ND;
                    if (lost >= CWK_NW) break;
                }
            }
        }
    }
    if (overflow) atomicAdd(&c.flags[0], 1);
}

__device__ __forceinline__ void accum_add4(float* p, float a, float b, float c, float d) {
    atomicAdd(reinterpret_cast<float4*>(p), make_float4(a, b, c, d));  // RED.E.ADD.F32x4 (sm_90+)
}

// MODE: SHADE_ALL walks the queue itself (hits and misses mixed in one warp); SHADE_HITS / SHADE_MISSES walk the index lists the
// extend kernel partitions finished rays into, so a warp runs either the surface + BRDF path or the environment path, not both
// serialised (ncu: 14-16 of 32 lanes active on bounces >= 1 with the mixed kernel).
#define SHADE_ALL 0
#define SHADE_HITS 1
#define SHADE_MISSES 2
template <bool PRIMARY, int MODE>
__global__ void __launch_bounds__(WF_BLOCK, DRP_SHADE_MINBLOCKS) k_shade(const __grid_constant__ WfConst c, int bounce, const float4* __restrict__ qa,
                                                    const float4* __restrict__ qb, const float4* __restrict__ qt, const float2* __restrict__ hit,
                                                    float4* __restrict__ oa, float4* __restrict__ ob, float4* __restrict__ ot,
                                                    const int* __restrict__ count_ptr, int* __restrict__ out_count, int* __restrict__ cursor,
                                                    const int* __restrict__ index_list) {
    const int count = (PRIMARY && MODE == SHADE_ALL) ? (int)c.R : *count_ptr;
    const int lane = threadIdx.x & 31;
    const bool last = bounce == c.p.ray_depth - 1;
    const bool always_sky = last && c.p.last_bounce_skybox;  // path_tracing.py:260
    for (;;) {
        int base = 0;
        if (lane == 0) base = atomicAdd(cursor, WF_FETCH);
        base = __shfl_sync(0xffffffffu, base, 0);
        if (base >= count) break;
#if DRP_SHADE_PIPELINE == 1
        // Software pipeline over the chunk (SHADE_ALL): the kernel is bound by the dependent DRAM round trips of one ray
        // (ray + hit record -> vertex indices / material id -> vertex records -> texels; ncu: long scoreboard ~10 warps per issue at
        // 0.3 IPC).  The hit record is loaded two iterations and the index quadruple one iteration before use, and the next ray's
        // queue entries are pulled towards L2, so that an iteration starts at the vertex fetch.
        const float2 h_miss = make_float2(c.p.t_far, 0.0f);
        float2 h_cur = h_miss, h_nxt = h_miss;
        TriIdx idx_cur = {0, 0, 0, 0};
        if (MODE == SHADE_ALL) {
            if (base + lane < count) h_cur = __ldg(hit + base + lane);
            if (base + 32 + lane < count) h_nxt = __ldg(hit + base + 32 + lane);
            if (h_cur.x < c.p.t_far) idx_cur = load_tri_idx(c.scene, __float_as_int(h_cur.y));
        }
#endif
#if DRP_SHADE_CHUNK_PARTITION
#if DRP_SHADE_PIPELINE
#error "DRP_SHADE_CHUNK_PARTITION reorders the rays of a chunk and cannot be combined with DRP_SHADE_PIPELINE"
#endif
        // Experiment for the next measurement round (ncu, profiles/README.md section 3: the secondary-bounce shade kernel executes at 13 of 32
        // lanes because hits -- vertex / texel fetches + BRDF -- and misses -- environment lookup -- share warps).  The chunk of WF_FETCH rays a
        // warp owns is classified with one ballot per 32 rays and then walked class by class: lane l of iteration i takes the (32 i + l)-th
        // hit, later the (32 i' + l)-th miss, so that a warp runs one of the two code paths with (nearly) all lanes.  Rays stay inside their
        // chunk (no global lists, no indirection through memory: the extend-side partition of DRP_PARTITION lost to its finish-ordered reads).
        constexpr int NQ = WF_FETCH / 32;
        const bool part_chunk = !PRIMARY && MODE == SHADE_ALL;
        unsigned cls_hit[NQ], cls_miss[NQ];
        int n_hit = 0, n_miss = 0;
        if (part_chunk) {
#pragma unroll
            for (int q = 0; q < NQ; ++q) {
                const int sl = base + 32 * q + lane;
                const bool valid = sl < count;
                const bool ish = valid && __ldg(hit + sl).x < c.p.t_far;
                cls_hit[q] = __ballot_sync(0xffffffffu, ish);
                cls_miss[q] = __ballot_sync(0xffffffffu, valid && !ish);
                n_hit += __popc(cls_hit[q]);
                n_miss += __popc(cls_miss[q]);
            }
        }
        const int it_hit = (n_hit + 31) >> 5, it_all = part_chunk ? it_hit + ((n_miss + 31) >> 5) : NQ;
#pragma unroll 1
        for (int it = 0; it < it_all; ++it) {
            const int j = 32 * it;
            int slot = base + j + lane;
            if (part_chunk) {
                const bool hit_phase = it < it_hit;
                int rank = (hit_phase ? it : it - it_hit) * 32 + lane;
                slot = count;  // no ray for this lane
                if (rank < (hit_phase ? n_hit : n_miss)) {
#pragma unroll
                    for (int q = 0; q < NQ; ++q) {
                        const unsigned m = hit_phase ? cls_hit[q] : cls_miss[q];
                        const int cq = __popc(m);
                        if (rank >= 0 && rank < cq) { slot = base + 32 * q + (int)__fns(m, 0, rank + 1); rank = -1; }
                        else if (rank >= 0) rank -= cq;
                    }
                }
            }
#else
#pragma unroll 1
        for (int j = 0; j < WF_FETCH; j += 32) {
            const int slot = base + j + lane;
#endif
            bool alive = false;
            Vec3 no = v3(0, 0, 0), nd = v3(0, 0, 0), T = v3(1, 1, 1);
            int ri = 0;
#if DRP_SHADE_PIPELINE == 1
            float2 h_n2 = h_miss;
            TriIdx idx_nxt = {0, 0, 0, 0};
            if (MODE == SHADE_ALL) {
                if (j + 64 < WF_FETCH && slot + 64 < count) h_n2 = __ldg(hit + slot + 64);
                if (j + 32 < WF_FETCH && h_nxt.x < c.p.t_far) idx_nxt = load_tri_idx(c.scene, __float_as_int(h_nxt.y));
                if (!PRIMARY && j + 32 < WF_FETCH && slot + 32 < count) {
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(qa + slot + 32));
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(qb + slot + 32));
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(qt + slot + 32));
                }
            }
#endif
#if DRP_SHADE_PIPELINE == 2
            // prefetch-only variant: no carried registers, the next ray's index lines and queue entries are pulled towards L2
            if (MODE == SHADE_ALL && j + 32 < WF_FETCH && slot + 32 < count) {
                const float2 hn = __ldg(hit + slot + 32);
                if (hn.x < c.p.t_far) {
                    const int idn = __float_as_int(hn.y);
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(c.scene.tris + 3 * (int64_t)idn));
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(c.scene.tri_material + idn));
                }
                if (!PRIMARY) {
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(qa + slot + 32));
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(qb + slot + 32));
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(qt + slot + 32));
                }
            }
#endif
            if (slot < count) {
                const int k = MODE == SHADE_ALL ? slot : __ldg(index_list + slot);  // queue slot of the ray
                Vec3 o, d;
                load_ray<PRIMARY>(c, qa, qb, k, o, d, ri);
                if (!PRIMARY) { float4 t4 = __ldg(qt + k); T = v3(t4.x, t4.y, t4.z); }
#if DRP_SHADE_PIPELINE == 1
                float2 h = MODE == SHADE_MISSES ? make_float2(c.p.t_far, 0.0f) : (MODE == SHADE_ALL ? h_cur : __ldg(hit + k));
                const TriIdx* pre = MODE == SHADE_ALL ? &idx_cur : nullptr;
#else
                float2 h = MODE == SHADE_MISSES ? make_float2(c.p.t_far, 0.0f) : __ldg(hit + k);
                const TriIdx* pre = nullptr;
#endif
                const float t = h.x;
                const bool is_hit = MODE == SHADE_HITS ? true : (MODE == SHADE_MISSES ? false : t < c.p.t_far);
                SurfaceAttrs s;
                if (is_hit) {
                    // last bounce of a path (not the first): only emission and alpha reach the outputs, skip the rest of the material
                    s = surface_attrs(c.scene, c.scene.materials, o + d * t, __float_as_int(h.y), !PRIMARY && last, pre);
                } else {
                    s.albedo = s.normal = s.emission = v3(0, 0, 0);
                    s.metal = s.smooth = s.alpha = 0.0f;
                }
                Vec3 env = (always_sky || !is_hit) ? env_fetch(c.scene.env, d) : v3(0, 0, 0);  // path_tracing.py:267-269
                float u[6] = {0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f};
                const int lpix = ri % c.HW, ly = lpix / c.tw;
                const int pix = (c.ty0 + ly) * c.p.width + c.tx0 + (lpix - ly * c.tw);  // global pixel: accumulator row and RNG key
                BounceOut r;
                if (!PRIMARY && last) {  // nothing is sampled after the last bounce: radiance only (shade.cuh: brdf_sample line 1-2)
                    r.hit_pos = o + d * t;
                    r.radiance = s.emission + env;
                    r.next_d = d; r.transfer = v3(0.0f, 0.0f, 0.0f);
                } else {
                    if (c.p.rng_mode == DRP_RNG_REPLAY) {
#pragma unroll
                        for (int q = 0; q < 6; ++q) u[q] = __ldg(c.p.replay_u + ((int64_t)bounce * 6 + q) * c.R_total + ri);
                    } else {
                        philox_uniform6(c.p.seed, (uint32_t)pix, (uint32_t)__ldg(c.p.sample_ids + ri / c.HW), (uint32_t)bounce, u);
                    }
                    r = brdf_sample(s, t, o, d, env, u);
                }
                float* acc = c.accum + (int64_t)DRP_ACCUM_CHANNELS * pix;
                accum_add4(acc, T.x * r.radiance.x, T.y * r.radiance.y, T.z * r.radiance.z, s.alpha);  // path_tracing.py:336-337
                if (PRIMARY) {  // extras on the first hit, path_tracing.py:340-347
                    accum_add4(acc + 4, s.albedo.x, s.albedo.y, s.albedo.z, s.emission.x);
                    accum_add4(acc + 8, s.emission.y, s.emission.z, s.normal.x, s.normal.y);
                    accum_add4(acc + 12, s.normal.z, r.hit_pos.x, r.hit_pos.y, r.hit_pos.z);
                }
                if (!last) {
                    T = is_hit ? T * r.transfer : v3(0, 0, 0);                 // path_tracing.py:338 with :275
                    nd = r.next_d;
                    no = r.hit_pos + nd * c.p.step_epsilon;                    // path_tracing.py:276
                    alive = true;
                    if (c.p.compaction && !is_hit) alive = ray_may_reach_box(no, nd, c.box_lo, c.box_hi);
                }
            }
#if DRP_SHADE_PIPELINE == 1
            h_cur = h_nxt; h_nxt = h_n2; idx_cur = idx_nxt;
#endif
            // warp-aggregated append to the output queue
            unsigned m = __ballot_sync(0xffffffffu, alive);
            if (m) {
                int pos = 0;
                if (lane == __ffs(m) - 1) pos = atomicAdd(out_count, __popc(m));
                pos = __shfl_sync(0xffffffffu, pos, __ffs(m) - 1);
                if (alive) {
                    int w = pos + __popc(m & ((1u << lane) - 1u));
                    oa[w] = make_float4(no.x, no.y, no.z, nd.x);
                    ob[w] = make_float4(nd.y, nd.z, __int_as_float(ri), 0.0f);
                    ot[w] = make_float4(T.x, T.y, T.z, 0.0f);
                }
            }
        }
    }
}

__global__ void k_count_traced(const int* __restrict__ counts, int depth, unsigned long long R, unsigned long long* __restrict__ total) {
    unsigned long long t = R;  // bounce 0 traces every ray of the batch
    for (int b = 1; b < depth; ++b) t += (unsigned long long)counts[b];
    atomicAdd(total, t);
}

__global__ void k_finalize(const float* __restrict__ accum, int H, int W, float spp, float* __restrict__ radiance, float* __restrict__ alpha,
                           float* __restrict__ albedo, float* __restrict__ emission, float* __restrict__ normal, float* __restrict__ position) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= H * W) return;
    int y = i / W, x = i - y * W;
    int src = (H - 1 - y) * W + x;  // flipud, path_tracing.py:348-352
    const float4* a = reinterpret_cast<const float4*>(accum) + 4 * (int64_t)src;
    float4 a0 = a[0], a1 = a[1], a2 = a[2], a3 = a[3];
    if (radiance) { radiance[3 * i] = a0.x / spp; radiance[3 * i + 1] = a0.y / spp; radiance[3 * i + 2] = a0.z / spp; }
    if (alpha) alpha[i] = fminf(fmaxf(a0.w / spp, 0.0f), 1.0f);
    if (albedo) { albedo[3 * i] = a1.x / spp; albedo[3 * i + 1] = a1.y / spp; albedo[3 * i + 2] = a1.z / spp; }
    if (emission) { emission[3 * i] = a1.w / spp; emission[3 * i + 1] = a2.x / spp; emission[3 * i + 2] = a2.y / spp; }
    if (normal) { normal[3 * i] = a2.z / spp; normal[3 * i + 1] = a2.w / spp; normal[3 * i + 2] = a3.x / spp; }
    if (position) { position[3 * i] = a3.y / spp; position[3 * i + 1] = a3.z / spp; position[3 * i + 2] = a3.w / spp; }
}

// ---- host side --------------------------------------------------------------------------------------------------
// Workspaces (ray queues etc.) outlive handles: sessions are single-use (one handle per frame), so a released handle
// parks its workspace in a per-process pool and the next handle on the same device adopts it instead of re-allocating.
#include <mutex>
static std::mutex g_ws_mutex;
static std::vector<std::pair<int, RenderWorkspace*>> g_ws_pool;

void drp_free_workspace(BvhHandle* h) {
    RenderWorkspace* ws = h->ws;
    if (!ws) return;
    for (auto& sp : ws->spans) { ws->event_pool.push_back(sp.a); ws->event_pool.push_back(sp.b); }
    ws->spans.clear();
    ws->live_used = 0;
    ws->profiling = false;
    ws->have_box = false;
    {
        std::lock_guard<std::mutex> lk(g_ws_mutex);
        g_ws_pool.push_back({h->device, ws});
    }
    h->ws = nullptr;
}

static int ensure_workspace(BvhHandle* h, int64_t rays, int n_mats) {
    if (!h->ws) {
        std::lock_guard<std::mutex> lk(g_ws_mutex);
        for (size_t k = 0; k < g_ws_pool.size(); ++k)
            if (g_ws_pool[k].first == h->device) {
                h->ws = g_ws_pool[k].second;
                g_ws_pool.erase(g_ws_pool.begin() + k);
                break;
            }
    }
    if (!h->ws) {
        h->ws = new RenderWorkspace();
        cudaDeviceProp prop;
        DRP_CUDA_CHECK(cudaGetDeviceProperties(&prop, h->device));
        h->ws->sm_count = prop.multiProcessorCount;
        h->ws->n_counters = 1024;  // two launch groups x (64 live counts, 128 fetch cursors, 4 x 64 partition counters / cursors)
        DRP_CUDA_CHECK(cudaMalloc((void**)&h->ws->counters, sizeof(int) * 1024));
        DRP_CUDA_CHECK(cudaMalloc((void**)&h->ws->d_traced, sizeof(unsigned long long)));
        int nb = 0;
        DRP_CUDA_CHECK(h->wide ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_extend_cw<SRC_PRIMARY, false>, WF_BLOCK, 0) : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_extend<true, false>, WF_BLOCK, 0));
        h->ws->grid_extend[1] = std::max(1, nb) * h->ws->sm_count;
        DRP_CUDA_CHECK(h->wide ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_extend_cw<SRC_QUEUE, false>, WF_BLOCK, 0) : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_extend<false, false>, WF_BLOCK, 0));
        h->ws->grid_extend[0] = std::max(1, nb) * h->ws->sm_count;
        DRP_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_shade<true, SHADE_ALL>, WF_BLOCK, 0));
        h->ws->grid_shade[1] = std::max(1, nb) * h->ws->sm_count;
        DRP_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_shade<false, SHADE_ALL>, WF_BLOCK, 0));
        h->ws->grid_shade[0] = std::max(1, nb) * h->ws->sm_count;
    }
    RenderWorkspace* ws = h->ws;
    if (rays > 0 && rays > ws->capacity) {
        for (int k = 0; k < 2; ++k) { cudaFree(ws->qa[k]); cudaFree(ws->qb[k]); cudaFree(ws->qt[k]); }
        cudaFree(ws->hit); cudaFree(ws->hit_list); cudaFree(ws->miss_list);
        ws->capacity = 0;
        for (int k = 0; k < 2; ++k) {
            DRP_CUDA_CHECK(cudaMalloc((void**)&ws->qa[k], sizeof(float4) * rays));
            DRP_CUDA_CHECK(cudaMalloc((void**)&ws->qb[k], sizeof(float4) * rays));
            DRP_CUDA_CHECK(cudaMalloc((void**)&ws->qt[k], sizeof(float4) * rays));
        }
        DRP_CUDA_CHECK(cudaMalloc((void**)&ws->hit, sizeof(float2) * rays));
        DRP_CUDA_CHECK(cudaMalloc((void**)&ws->hit_list, sizeof(int) * rays));
        DRP_CUDA_CHECK(cudaMalloc((void**)&ws->miss_list, sizeof(int) * rays));
        ws->capacity = rays;
    }
    if (n_mats > ws->mats_capacity) {
        cudaFree(ws->d_mats);
        DRP_CUDA_CHECK(cudaMalloc((void**)&ws->d_mats, sizeof(drp_material_t) * n_mats));
        ws->mats_capacity = n_mats;
        ws->mats_host_copy.clear();
    }
    return DRP_OK;
}

#define WF_MAX_BATCH_RAYS (int64_t(1) << 24)  // 16M rays per internal batch (~1.9 GB of queues)
#define WF_MAX_DEPTH 60

extern "C" int drp_render(uint64_t handle, const drp_scene_t* scene, const drp_render_params_t* params, float* accum, void* stream) {
    BvhHandle* h = drp_lookup(handle);
    if (!h) { drp_set_error("drp_render: unknown handle"); return DRP_ERR_HANDLE; }
    if (!scene || !params || !accum) { drp_set_error("drp_render: null argument"); return DRP_ERR_INVALID; }
    const drp_render_params_t& p = *params;
    if (p.height <= 0 || p.width <= 0 || p.ray_depth <= 0 || p.ray_depth > WF_MAX_DEPTH || p.n_samples < 0) {
        drp_set_error("drp_render: bad resolution / depth / sample count");
        return DRP_ERR_INVALID;
    }
    if (scene->n_tris != h->n_tris) { drp_set_error("drp_render: scene does not match the built structure"); return DRP_ERR_INVALID; }
    if (p.rng_mode == DRP_RNG_REPLAY && !p.replay_u) { drp_set_error("drp_render: replay mode without replay_u"); return DRP_ERR_INVALID; }
    if (!p.ndc_x || !p.ndc_y || !p.jitter_x || !p.jitter_y || !p.sample_ids) { drp_set_error("drp_render: missing raygen tables"); return DRP_ERR_INVALID; }
    if (scene->n_materials <= 0 || !scene->materials) { drp_set_error("drp_render: scene has no materials"); return DRP_ERR_INVALID; }
    {   // the shade kernel fetches texels with one 128-bit load: every texture must be RGBA (diffrp_b200.flatten.pad_rgba)
        auto ok = [](const drp_texture_t& t) { return t.data == nullptr || (t.c == 4 && t.h > 0 && t.w > 0); };
        bool all = ok(scene->env);
        for (int k = 0; k < scene->n_materials; ++k) {
            const drp_material_t& m = scene->materials[k];
            all = all && ok(m.base_color_tex) && ok(m.mr_tex) && ok(m.normal_tex) && ok(m.emissive_tex);
        }
        if (!all) { drp_set_error("drp_render: textures must be 4-channel (RGBA-padded) fp32 images"); return DRP_ERR_INVALID; }
    }
    const bool tiled = p.tile_w > 0 && p.tile_h > 0;
    if (tiled && (p.tile_x0 < 0 || p.tile_y0 < 0 || p.tile_x0 + p.tile_w > p.width || p.tile_y0 + p.tile_h > p.height)) {
        drp_set_error("drp_render: tile outside the frame");
        return DRP_ERR_INVALID;
    }
    const int64_t HW = tiled ? (int64_t)p.tile_w * p.tile_h : (int64_t)p.height * p.width;
    if (HW > WF_MAX_BATCH_RAYS) { drp_set_error("drp_render: more than 2^24 pixels per frame not supported"); return DRP_ERR_INVALID; }
    if (p.n_samples == 0) return DRP_OK;
    if (HW * p.n_samples >= (int64_t(1) << 31)) { drp_set_error("drp_render: more than 2^31 rays per call; render the samples in several calls"); return DRP_ERR_INVALID; }
    DeviceGuard guard(h->device);
    if (!guard.ok) { drp_set_error("drp_render: cannot select device"); return DRP_ERR_CUDA; }
    cudaStream_t s = (cudaStream_t)stream;
    // Experiment (DRP_L2_PERSIST_MB=<n>): pin the wide-node array in L2 (persisting access-policy window on the render stream) so that the
    // texel / queue streams of k_shade do not evict the hierarchy between two extend launches.
    static const int persist_mb = getenv("DRP_L2_PERSIST_MB") ? atoi(getenv("DRP_L2_PERSIST_MB")) : 0;
    if (persist_mb > 0 && h->wide) {
        static bool limit_set = false;
        if (!limit_set) { cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, (size_t)persist_mb << 20); limit_set = true; }
        cudaStreamAttrValue av;
        memset(&av, 0, sizeof(av));
        const size_t node_bytes = (size_t)h->n_nodes_used * 16 * CW_NODE_F4;
        av.accessPolicyWindow.base_ptr = (void*)h->nodes;
        av.accessPolicyWindow.num_bytes = node_bytes;
        av.accessPolicyWindow.hitRatio = std::min(1.0f, (float)((double)((size_t)persist_mb << 20) / (double)std::max<size_t>(node_bytes, 1)));
        av.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
        av.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
        cudaError_t e = cudaStreamSetAttribute(s, cudaStreamAttributeAccessPolicyWindow, &av);
        static bool reported = false;
        if (!reported) { fprintf(stderr, "[diffrp_b200] L2 persisting window: %zu bytes of nodes, %d MB set aside: %s\n", node_bytes, persist_mb, cudaGetErrorString(e)); reported = true; }
        (void)cudaGetLastError();
    }
    const int spb = p.reproducible ? 1 : (int)std::max<int64_t>(1, std::min<int64_t>(p.n_samples, WF_MAX_BATCH_RAYS / HW));  // samples per batch
    int rc = ensure_workspace(h, (((int64_t)spb * HW + 1) / 2) * 2 + 64, scene->n_materials);
    if (rc != DRP_OK) return rc;
    RenderWorkspace* ws = h->ws;
    // material table -> device (skipped when unchanged since the previous call)
    const size_t mbytes = sizeof(drp_material_t) * scene->n_materials;
    if (ws->mats_host_copy.size() != mbytes || memcmp(ws->mats_host_copy.data(), scene->materials, mbytes) != 0) {
        ws->mats_host_copy.assign((const unsigned char*)scene->materials, (const unsigned char*)scene->materials + mbytes);
        DRP_CUDA_CHECK(cudaMemcpyAsync(ws->d_mats, ws->mats_host_copy.data(), mbytes, cudaMemcpyHostToDevice, s));
    }
    WfConst c;
    memset(&c, 0, sizeof(c));
    c.scene = *scene;
    c.scene.materials = ws->d_mats;
    c.p = p;
    c.nodes = h->nodes;
    c.tris = h->packed;
    c.eps = h->eps;
    c.HW = (int)HW;
    c.tx0 = tiled ? p.tile_x0 : 0; c.ty0 = tiled ? p.tile_y0 : 0; c.tw = tiled ? p.tile_w : p.width;
    {
        const int rect_h = tiled ? p.tile_h : p.height;
        // A/B on B200: 8x4 blocks leave extend unchanged and make shade 5 % slower (accumulator RED rows less contiguous) -> off by default
        static const bool blocks = getenv("DRP_PRIMARY_ORDER") && (strcmp(getenv("DRP_PRIMARY_ORDER"), "tiled") == 0 || strcmp(getenv("DRP_PRIMARY_ORDER"), "sample_tiled") == 0);
        c.tiled8x4 = (blocks && c.tw % 8 == 0 && rect_h % 4 == 0) ? 1 : 0;
    }
    c.R_total = HW * p.n_samples;
    c.accum = accum;
    c.flags = h->dev_flags;
    c.cw_bias = 0x47000000u;
    if (!ws->have_box) {  // padded scene box for the exact compaction rule: one host read per handle, then cached
        uint32_t ob[12];
        DRP_CUDA_CHECK(cudaMemcpyAsync(ob, h->bounds, sizeof(ob), cudaMemcpyDeviceToHost, s));
        DRP_CUDA_CHECK(cudaStreamSynchronize(s));
        for (int a = 0; a < 3; ++a) {
            float lo = ord2f(ob[a]), hi = ord2f(ob[3 + a]);
            float pad = 1e-5f * fmaxf(fmaxf(fabsf(lo), fabsf(hi)), 1e-30f) + 1e-5f * fmaxf(hi - lo, 0.0f);
            ws->box_lo[a] = lo - pad; ws->box_hi[a] = hi + pad;
        }
        if (h->n_tris == 0) for (int a = 0; a < 3; ++a) { ws->box_lo[a] = 1.0f; ws->box_hi[a] = -1.0f; }
        ws->have_box = true;
    }
    for (int a = 0; a < 3; ++a) { c.box_lo[a] = ws->box_lo[a]; c.box_hi[a] = ws->box_hi[a]; }
    const int D = p.ray_depth;
    static const bool simple_extend = getenv("DRP_EXTEND") && strcmp(getenv("DRP_EXTEND"), "simple") == 0;  // A/B profiling switch
    int64_t launches = 0;
    h->last_render.rays_nominal = HW * p.n_samples * D;
    DRP_CUDA_CHECK(cudaMemsetAsync(ws->d_traced, 0, sizeof(unsigned long long), s));
    // Overlap (DRP_OVERLAP=1, experimental): every batch is split into two ray groups that run as independent kernel chains on two
    // internal streams, the second one bounce-phase behind the first, with grids sized so that an (ALU-bound) extend kernel
    // of one group and a (DRAM-latency-bound) shade kernel of the other are co-resident on every SM.
    static const bool overlap_env = getenv("DRP_OVERLAP") && atoi(getenv("DRP_OVERLAP")) != 0;
    static const int ovl_e = getenv("DRP_OVL_E") ? atoi(getenv("DRP_OVL_E")) : 5, ovl_s = getenv("DRP_OVL_S") ? atoi(getenv("DRP_OVL_S")) : 2;
    const bool overlap = overlap_env && h->wide && !simple_extend;
    // A/B on B200: partitioned shading is 5 % SLOWER on config 3 (shade 3.36 vs 3.04 ms / step: the indirect, finish-ordered ray reads cost
    // more than the hit/miss divergence they remove) -> off unless DRP_PARTITION=1
    static const bool partition_env = getenv("DRP_PARTITION") && atoi(getenv("DRP_PARTITION")) != 0;
    const bool partition = partition_env && h->wide && !simple_extend;
    if (overlap && !ws->streams[0]) {
        for (int g = 0; g < 2; ++g) DRP_CUDA_CHECK(cudaStreamCreateWithFlags(&ws->streams[g], cudaStreamNonBlocking));
        for (int g = 0; g < 4; ++g) DRP_CUDA_CHECK(cudaEventCreateWithFlags(&ws->sync_events[g], cudaEventDisableTiming));
    }
    for (int s0 = 0; s0 < p.n_samples; s0 += spb) {
        const int ns = std::min(spb, p.n_samples - s0);
        const int64_t R_batch = (int64_t)ns * HW;
        const int G = (overlap && R_batch >= (1 << 18)) ? 2 : 1;
        DRP_CUDA_CHECK(cudaMemsetAsync(ws->counters, 0, sizeof(int) * ws->n_counters, s));  // (a memset node, not counted as a kernel launch)
        if (G == 2) {  // fork: both internal streams start after everything already queued on the caller's stream
            DRP_CUDA_CHECK(cudaEventRecord(ws->sync_events[0], s));
            for (int g = 0; g < 2; ++g) DRP_CUDA_CHECK(cudaStreamWaitEvent(ws->streams[g], ws->sync_events[0], 0));
        }
        for (int b = 0; b < D; ++b) {
            for (int g = 0; g < G; ++g) {
                cudaStream_t sg = G == 2 ? ws->streams[g] : s;
                const int64_t r0 = R_batch * g / G, r1 = R_batch * (g + 1) / G;
                const int64_t qoff = (int64_t)g * (ws->capacity / 2);  // each group owns one half of every queue
                c.R = r1 - r0;
                c.ray_base = (int64_t)s0 * HW + r0;
                // default: pixel-major / sample-minor (A/B on config 3: 8.41 -> 7.57 ms per step); DRP_PRIMARY_ORDER=scan restores sample-major
                static const bool sample_order = !getenv("DRP_PRIMARY_ORDER") || strcmp(getenv("DRP_PRIMARY_ORDER"), "sample") == 0 ||
                                                 strcmp(getenv("DRP_PRIMARY_ORDER"), "sample_tiled") == 0;
                c.sample_minor = (sample_order && G == 1 && ns > 1) ? ns : 0;
                int* counts = ws->counters + 512 * g;        // [0..D]
                int* cursors = ws->counters + 512 * g + 64;  // [0..2D)
                int* pc = ws->counters + 512 * g + 192;      // partition: hit counts [0..63], miss counts [64..127], hit cursors [128..191], miss cursors [192..255]
                Partition part = {nullptr, nullptr, nullptr, nullptr};
                if (partition) part = Partition{ws->hit_list + qoff, ws->miss_list + qoff, pc + b, pc + 64 + b};
                float2* hit = ws->hit + qoff;
                const int in = b & 1, out = in ^ 1;
                float4 *qa_in = ws->qa[in] + qoff, *qb_in = ws->qb[in] + qoff, *qt_in = ws->qt[in] + qoff;
                float4 *qa_out = ws->qa[out] + qoff, *qb_out = ws->qb[out] + qoff, *qt_out = ws->qt[out] + qoff;
                const int ge = G == 2 ? ws->sm_count * ovl_e : ws->grid_extend[b == 0], gs = G == 2 ? ws->sm_count * ovl_s : ws->grid_shade[b == 0];
                if (G == 2 && g == 1 && b == 0) DRP_CUDA_CHECK(cudaStreamWaitEvent(sg, ws->sync_events[1], 0));  // phase offset: after group 0's first extend
                auto span_begin = [&](int kind, const int* count_ptr) -> int {
                    if (!ws->profiling || ws->live_used >= ws->live_capacity) return -1;
                    RenderWorkspace::Span sp;
                    sp.a = wf_event(ws); sp.b = wf_event(ws); sp.kind = kind; sp.bounce = b;
                    sp.traced_slot = ws->d_live + ws->live_used++;
                    k_record_live<<<1, 1, 0, sg>>>(count_ptr, (unsigned long long)c.R, sp.traced_slot);
                    ++launches;
                    cudaEventRecord(sp.a, sg);
                    ws->spans.push_back(sp);
                    return (int)ws->spans.size() - 1;
                };
                auto span_end = [&](int idx) { if (idx >= 0) cudaEventRecord(ws->spans[idx].b, sg); };
                if (b == 0) {
                    int sp = span_begin(0, nullptr);
                    if (h->wide && !simple_extend) { if (partition) k_extend_cw<SRC_PRIMARY, true><<<ge, WF_BLOCK, 0, sg>>>(c, nullptr, nullptr, hit, nullptr, cursors + 0, AosRays(), part);
                      else k_extend_cw<SRC_PRIMARY, false><<<ge, WF_BLOCK, 0, sg>>>(c, nullptr, nullptr, hit, nullptr, cursors + 0, AosRays(), part); }
                    else if (h->wide) k_extend<true, true><<<ge, WF_BLOCK, 0, sg>>>(c, nullptr, nullptr, hit, nullptr, cursors + 0);
                    else k_extend<true, false><<<ge, WF_BLOCK, 0, sg>>>(c, nullptr, nullptr, hit, nullptr, cursors + 0);
                    span_end(sp);
                    if (G == 2 && g == 0) DRP_CUDA_CHECK(cudaEventRecord(ws->sync_events[1], sg));
                    sp = span_begin(1, nullptr);
                    if (partition) {
                        k_shade<true, SHADE_HITS><<<gs, WF_BLOCK, 0, sg>>>(c, b, nullptr, nullptr, nullptr, hit, qa_out, qb_out, qt_out, pc + b, counts + 1, pc + 128 + b, part.hit_list);
                        k_shade<true, SHADE_MISSES><<<gs, WF_BLOCK, 0, sg>>>(c, b, nullptr, nullptr, nullptr, hit, qa_out, qb_out, qt_out, pc + 64 + b, counts + 1, pc + 192 + b, part.miss_list);
                        ++launches;
                    } else {
                        k_shade<true, SHADE_ALL><<<gs, WF_BLOCK, 0, sg>>>(c, b, nullptr, nullptr, nullptr, hit, qa_out, qb_out, qt_out, nullptr, counts + 1, cursors + 1, nullptr);
                    }
                    span_end(sp);
                } else {
                    int sp = span_begin(0, counts + b);
                    if (h->wide && !simple_extend) { if (partition) k_extend_cw<SRC_QUEUE, true><<<ge, WF_BLOCK, 0, sg>>>(c, qa_in, qb_in, hit, counts + b, cursors + 2 * b, AosRays(), part);
                      else k_extend_cw<SRC_QUEUE, false><<<ge, WF_BLOCK, 0, sg>>>(c, qa_in, qb_in, hit, counts + b, cursors + 2 * b, AosRays(), part); }
                    else if (h->wide) k_extend<false, true><<<ge, WF_BLOCK, 0, sg>>>(c, qa_in, qb_in, hit, counts + b, cursors + 2 * b);
                    else k_extend<false, false><<<ge, WF_BLOCK, 0, sg>>>(c, qa_in, qb_in, hit, counts + b, cursors + 2 * b);
                    span_end(sp);
                    sp = span_begin(1, counts + b);
                    if (partition) {
                        k_shade<false, SHADE_HITS><<<gs, WF_BLOCK, 0, sg>>>(c, b, qa_in, qb_in, qt_in, hit, qa_out, qb_out, qt_out, pc + b, counts + b + 1, pc + 128 + b, part.hit_list);
                        k_shade<false, SHADE_MISSES><<<gs, WF_BLOCK, 0, sg>>>(c, b, qa_in, qb_in, qt_in, hit, qa_out, qb_out, qt_out, pc + 64 + b, counts + b + 1, pc + 192 + b, part.miss_list);
                        ++launches;
                    } else {
                        k_shade<false, SHADE_ALL><<<gs, WF_BLOCK, 0, sg>>>(c, b, qa_in, qb_in, qt_in, hit, qa_out, qb_out, qt_out, counts + b, counts + b + 1, cursors + 2 * b + 1, nullptr);
                    }
                    span_end(sp);
                }
                launches += 2;
                if (b == D - 1) {
                    k_count_traced<<<1, 1, 0, sg>>>(counts, D, (unsigned long long)c.R, ws->d_traced);  // read lazily by drp_render_stats
                    ++launches;
                }
            }
        }
        if (G == 2) {  // join: the caller's stream continues after both chains
            for (int g = 0; g < 2; ++g) {
                DRP_CUDA_CHECK(cudaEventRecord(ws->sync_events[2 + g], ws->streams[g]));
                DRP_CUDA_CHECK(cudaStreamWaitEvent(s, ws->sync_events[2 + g], 0));
            }
        }
    }
    DRP_CUDA_CHECK(cudaGetLastError());
    h->last_render.kernel_launches = launches;
    return DRP_OK;
}

// Standalone Raycaster.query over the wide layout with the persistent kernel (called by drp_trace).
int drp_trace_wide_persistent(BvhHandle* h, const float* ro, const float* rd, float* out_t, int32_t* out_i, float t_far, int64_t n, cudaStream_t s) {
    int rc = ensure_workspace(h, 0, 0);
    if (rc != DRP_OK) return rc;
    RenderWorkspace* ws = h->ws;
    if (n > 0x7fffff00) { drp_set_error("drp_trace: more than 2^31 rays per call"); return DRP_ERR_INVALID; }
    WfConst c;
    memset(&c, 0, sizeof(c));
    c.nodes = h->nodes; c.tris = h->packed; c.eps = h->eps; c.R = n; c.p.t_far = t_far; c.flags = h->dev_flags; c.cw_bias = 0x47000000u;
    int* cursor = ws->counters + 1020;
    DRP_CUDA_CHECK(cudaMemsetAsync(cursor, 0, sizeof(int), s));
    AosRays aos = {ro, rd, out_t, out_i};
    int grid = ws->grid_extend[1];
    k_extend_cw<SRC_AOS, false><<<grid, WF_BLOCK, 0, s>>>(c, nullptr, nullptr, nullptr, nullptr, cursor, aos, Partition{nullptr, nullptr, nullptr, nullptr});
    DRP_CUDA_CHECK(cudaGetLastError());
    return DRP_OK;
}

// ---- drp_surface_attrs: the material layer for arbitrary ray batches (custom samplers) -------------------------------------
// layer_material_rays + _super_collector + g-buffer collect (path_tracing.py:158-187, mixin.py:115-155, interpolator.py:32-48) for a
// batch of (ray, t, primitive id): one thread per ray, attrs (R,12) = [albedo3 | normal3 | metal | smooth | alpha | emission3], zeros on a
// miss -- the same function (shade.cuh: surface_attrs) the fused k_shade calls.
__global__ void __launch_bounds__(128) k_surface_attrs(const drp_scene_t sc, const float* __restrict__ ro, const float* __restrict__ rd,
                                                       const float* __restrict__ t, const int32_t* __restrict__ tri, float t_far, int64_t n,
                                                       float* __restrict__ attrs) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    float4* out = reinterpret_cast<float4*>(attrs + 12 * r);
    const float tt = __ldg(t + r);
    const int id = __ldg(tri + r);
    if (!(tt < t_far) || id < 0 || id >= sc.n_tris) {
        out[0] = out[1] = out[2] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        return;
    }
    const Vec3 o = v3(ro[3 * r], ro[3 * r + 1], ro[3 * r + 2]), d = v3(rd[3 * r], rd[3 * r + 1], rd[3 * r + 2]);
    const SurfaceAttrs s = surface_attrs(sc, sc.materials, o + d * tt, id);
    out[0] = make_float4(s.albedo.x, s.albedo.y, s.albedo.z, s.normal.x);
    out[1] = make_float4(s.normal.y, s.normal.z, s.metal, s.smooth);
    out[2] = make_float4(s.alpha, s.emission.x, s.emission.y, s.emission.z);
}

extern "C" int drp_surface_attrs(uint64_t handle, const drp_scene_t* scene, const float* rays_o, const float* rays_d, const float* t,
                                 const int32_t* tri, float t_far, int64_t n_rays, float* attrs, void* stream) {
    BvhHandle* h = drp_lookup(handle);
    if (!h) { drp_set_error("drp_surface_attrs: invalid handle"); return DRP_ERR_INVALID; }
    if (!scene || n_rays < 0) { drp_set_error("drp_surface_attrs: invalid argument"); return DRP_ERR_INVALID; }
    if (n_rays == 0) return DRP_OK;
    if (!rays_o || !rays_d || !t || !tri || !attrs) { drp_set_error("drp_surface_attrs: NULL array"); return DRP_ERR_INVALID; }
    if (scene->n_materials <= 0 || !scene->materials) { drp_set_error("drp_surface_attrs: scene has no materials"); return DRP_ERR_INVALID; }
    {
        auto ok = [](const drp_texture_t& x) { return x.data == nullptr || (x.c == 4 && x.h > 0 && x.w > 0); };
        bool all = true;
        for (int k = 0; k < scene->n_materials; ++k) {
            const drp_material_t& m = scene->materials[k];
            all = all && ok(m.base_color_tex) && ok(m.mr_tex) && ok(m.normal_tex) && ok(m.emissive_tex);
        }
        if (!all) { drp_set_error("drp_surface_attrs: textures must be 4-channel (RGBA-padded) fp32 images"); return DRP_ERR_INVALID; }
    }
    DeviceGuard guard(h->device);
    if (!guard.ok) { drp_set_error("drp_surface_attrs: cannot select device"); return DRP_ERR_CUDA; }
    cudaStream_t s = (cudaStream_t)stream;
    int rc = ensure_workspace(h, 0, scene->n_materials);  // material table only, no ray queues
    if (rc != DRP_OK) return rc;
    RenderWorkspace* ws = h->ws;
    const size_t mbytes = sizeof(drp_material_t) * scene->n_materials;
    if (ws->mats_host_copy.size() != mbytes || memcmp(ws->mats_host_copy.data(), scene->materials, mbytes) != 0) {
        ws->mats_host_copy.assign((const unsigned char*)scene->materials, (const unsigned char*)scene->materials + mbytes);
        DRP_CUDA_CHECK(cudaMemcpyAsync(ws->d_mats, ws->mats_host_copy.data(), mbytes, cudaMemcpyHostToDevice, s));
    }
    drp_scene_t sc = *scene;
    sc.materials = ws->d_mats;
    k_surface_attrs<<<(unsigned)((n_rays + 127) / 128), 128, 0, s>>>(sc, rays_o, rays_d, t, tri, t_far, n_rays, attrs);
    DRP_CUDA_CHECK(cudaGetLastError());
    return DRP_OK;
}

extern "C" int drp_render_stats(uint64_t handle, drp_render_stats_t* out) {
    BvhHandle* h = drp_lookup(handle);
    if (!h || !out) { drp_set_error("drp_render_stats: unknown handle"); return DRP_ERR_HANDLE; }
    DeviceGuard guard(h->device);
    DRP_CUDA_CHECK(cudaDeviceSynchronize());
    *out = h->last_render;
    if (h->ws) {
        unsigned long long traced = 0;
        DRP_CUDA_CHECK(cudaMemcpy(&traced, h->ws->d_traced, sizeof(traced), cudaMemcpyDeviceToHost));
        out->rays_traced = (int64_t)traced;
    }
    int flags[4] = {0, 0, 0, 0};
    DRP_CUDA_CHECK(cudaMemcpy(flags, h->dev_flags, sizeof(flags), cudaMemcpyDeviceToHost));
    if (flags[0] != 0) {
        drp_set_error("traversal stack overflow on " + std::to_string(flags[0]) + " rays: results are invalid");
        return DRP_ERR_INVALID;
    }
    return DRP_OK;
}

#define DRP_STR2(x) #x
#define DRP_STR(x) DRP_STR2(x)
extern "C" const char* drp_build_config(void) {
    return "compiled " __DATE__ " " __TIME__ "; DRP_CW_HALFSKIP=" DRP_STR(DRP_CW_HALFSKIP) " DRP_EXTEND_MINBLOCKS=" DRP_STR(DRP_EXTEND_MINBLOCKS)
           " DRP_SHADE_MINBLOCKS=" DRP_STR(DRP_SHADE_MINBLOCKS) " CWK_CHUNK=" DRP_STR(CWK_CHUNK) " CWK_ND=" DRP_STR(CWK_ND) " CWK_NW=" DRP_STR(CWK_NW)
           " CWK_POSTPONE=" DRP_STR(CWK_POSTPONE) " CWK_SMEM_STACK=" DRP_STR(CWK_SMEM_STACK) " CWK_PREFETCH=" DRP_STR(CWK_PREFETCH)
           " DRP_CW_V2=" DRP_STR(DRP_CW_V2) " DRP_CW_NODE96=" DRP_STR(DRP_CW_NODE96) " CWK_LD256=" DRP_STR(CWK_LD256) " CWK_LUT=" DRP_STR(CWK_LUT) " DRP_SHADE_PIPELINE=" DRP_STR(DRP_SHADE_PIPELINE) " DRP_SHADE_CHUNK_PARTITION=" DRP_STR(DRP_SHADE_CHUNK_PARTITION) " CWK_TRI_PIPE=" DRP_STR(CWK_TRI_PIPE);
}

extern "C" int drp_set_profiling(uint64_t handle, int enable) {
    BvhHandle* h = drp_lookup(handle);
    if (!h) { drp_set_error("drp_set_profiling: unknown handle"); return DRP_ERR_HANDLE; }
    DeviceGuard guard(h->device);
    int rc = ensure_workspace(h, 0, 0);
    if (rc != DRP_OK) return rc;
    RenderWorkspace* ws = h->ws;
    if (enable && !ws->d_live) {
        ws->live_capacity = 1 << 16;
        DRP_CUDA_CHECK(cudaMalloc((void**)&ws->d_live, sizeof(unsigned long long) * ws->live_capacity));
    }
    ws->profiling = enable != 0;
    return DRP_OK;
}

extern "C" int drp_get_profile(uint64_t handle, drp_profile_t* out) {
    BvhHandle* h = drp_lookup(handle);
    if (!h || !out) { drp_set_error("drp_get_profile: unknown handle"); return DRP_ERR_HANDLE; }
    memset(out, 0, sizeof(*out));
    if (!h->ws) return DRP_OK;
    DeviceGuard guard(h->device);
    RenderWorkspace* ws = h->ws;
    DRP_CUDA_CHECK(cudaDeviceSynchronize());
    std::vector<unsigned long long> live(ws->live_used > 0 ? ws->live_used : 1);
    if (ws->live_used > 0) DRP_CUDA_CHECK(cudaMemcpy(live.data(), ws->d_live, sizeof(unsigned long long) * ws->live_used, cudaMemcpyDeviceToHost));
    for (auto& sp : ws->spans) {
        float ms = 0.0f;
        cudaEventElapsedTime(&ms, sp.a, sp.b);
        unsigned long long rays = live[sp.traced_slot - ws->d_live];
        if (sp.kind == 0) { out->extend_ms += ms; out->extend_launches++; out->extend_rays += (int64_t)rays; }
        else { out->shade_ms += ms; out->shade_launches++; out->shade_rays += (int64_t)rays; }
        ws->event_pool.push_back(sp.a);
        ws->event_pool.push_back(sp.b);
    }
    ws->spans.clear();
    ws->live_used = 0;
    return DRP_OK;
}

extern "C" int drp_finalize(const float* accum, int32_t height, int32_t width, int32_t spp_total, float* radiance, float* alpha, float* albedo,
                            float* emission, float* world_normal, float* world_position, void* stream) {
    if (!accum || height <= 0 || width <= 0 || spp_total <= 0) { drp_set_error("drp_finalize: invalid argument"); return DRP_ERR_INVALID; }
    const int n = height * width;
    k_finalize<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(accum, height, width, (float)spp_total, radiance, alpha, albedo, emission,
                                                                 world_normal, world_position);
    DRP_CUDA_CHECK(cudaGetLastError());
    return DRP_OK;
}
