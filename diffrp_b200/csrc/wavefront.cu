// wavefront.cu -- fused wavefront path tracer: drp_render / drp_finalize / drp_render_stats / profiling.
//
// Replaces the Python section x bounce loop of PathTracingSession.trace_rays with the built-in sampler_brdf
// (diffrp/rendering/path_tracing.py:250-279, 310-352).  Per batch of samples and per bounce, two persistent kernels:
//   k_extend_cw : closest hit for every live ray over the compressed 8-wide BVH (cwbvh.cuh): one ray per lane, dynamic ray
//                 fetch by warp ballot, triangle postponing; bounce 0 generates the primary ray from the ray index instead
//                 of reading it; writes (t, id).  k_extend_fixup re-traces the (rare) rays that outgrew the per-thread stack.
//   k_shade     : surface attributes + env lookup + BRDF sample + fp32 accumulation (RED.ADD.F32x4) + next ray, appended to
//                 the output queue through a warp-aggregated atomic (stream compaction fused into the shade kernel)
// Ray state lives in HBM as float4 SoA queues (coalesced 128-bit accesses):
//   q_a[k] = (o.x, o.y, o.z, d.x)   q_b[k] = (d.y, d.z, bits(ray index), 0)   q_t[k] = (T.r, T.g, T.b, 0)   hit[k] = (t, bits(id))
// No environment variable changes what these kernels do; the design alternatives that were measured and lost are recorded in
// profiles/README.md and no longer exist in the source.  Scheduling constants (CWK_*) can be overridden at compile time only.
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
#include <nvtx3/nvToolsExt.h>

#define WF_BLOCK 128
#ifndef WF_FETCH
#define WF_FETCH 128  // rays fetched per warp per atomic
#endif

struct RenderWorkspace {
    int64_t capacity = 0;  // rays
    float4 *qa[2] = {nullptr, nullptr}, *qb[2] = {nullptr, nullptr}, *qt[2] = {nullptr, nullptr};
    float2* hit = nullptr;
    int* ovf_list = nullptr;   // queue slots of the rays whose traversal outgrew the per-thread stack (k_extend_fixup)
    uint2* deep_stack = nullptr;
    int* counters = nullptr;  // [0..D] live counts per bounce, [64..64+2D) fetch cursors, [192..192+D) flagged-ray counts
    int n_counters = 0;
    drp_material_t* d_mats = nullptr;
    int mats_capacity = 0;
    std::vector<unsigned char> mats_host_copy;
    unsigned long long* d_traced = nullptr;  // device: live rays traced by the last call (sum over batches and bounces)
    int sm_count = 0;
    int grid_extend[2] = {0, 0}, grid_shade[2] = {0, 0};  // [1]: bounce 0 (primary) kernels
    bool have_box = false;
    float box_lo[3], box_hi[3];
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
    int sample_minor;   // > 0: primary rays enumerated pixel-major / sample-minor with this many samples per batch (see load_ray)
    int64_t R;          // rays in this batch
    int64_t R_total;    // rays of the whole call (replay indexing)
    int64_t ray_base;   // index (within the call) of this launch group's first ray: s_local * HW + pixel of queue slot 0 at bounce 0
    float* accum;
    uint32_t cw_bias;   // 0x47000000 (cwbvh.cuh: cw_byte_biased), passed through the constant bank
};

template <bool PRIMARY>
__device__ __forceinline__ void load_ray(const WfConst& c, const float4* __restrict__ qa, const float4* __restrict__ qb, int k, Vec3& o,
                                         Vec3& d, int& ray_index) {
    if (PRIMARY) {
        // Queue slot -> (sample, pixel): sample-major, or pixel-major / sample-minor within the batch (sample_minor > 0).
        // The ray index (RNG key, replay index, accumulator row) is the reference's s * HW + y * W + x either way.
        const int kk = k + (int)c.ray_base;
        int s = kk / c.HW;
        int pix = kk - s * c.HW;
        if (c.sample_minor > 0) {  // the samples of one pixel sit next to each other in the queue: they hit the same triangles / texels
            pix = k / c.sample_minor;                          // ray_base is a multiple of HW in this mode
            s = (int)(c.ray_base / c.HW) + (k - pix * c.sample_minor);
        }
        ray_index = s * c.HW + pix;
        int y = pix / c.tw, x = pix - y * c.tw;
        x += c.tx0; y += c.ty0;
        float gx = __ldg(c.p.ndc_x + x) + __ldg(c.p.jitter_x + s);
        float gy = __ldg(c.p.ndc_y + y) + __ldg(c.p.jitter_y + s);
        gen_primary_ray(c.p.inv_vp, c.p.cam_pos, c.p.t_near, gx, gy, o, d);
    } else {
        // queue entries are read once: streaming loads (evict-first) keep them from displacing hierarchy nodes in L1 / L2
        // (profiles/r2/ab_cache_policy_r2.json: extend -0.5 %)
        float4 a = __ldcs(qa + k), b = __ldcs(qb + k);
        o = v3(a.x, a.y, a.z);
        d = v3(a.w, b.x, b.y);
        ray_index = __float_as_int(b.z);
    }
}

// ---- persistent-thread extend over the wide layout -----------------------------------------------------------------
// Every lane owns one ray and walks the 8-wide hierarchy with the (node group, triangle group) state of cwbvh.cuh.
//  * dynamic ray fetch (Aila & Laine 2009 / Ylitie et al. 2017): a lane whose ray terminated does not wait for the slowest
//    ray of its warp -- when the accumulated number of idle lane-iterations exceeds CWK_NW the warp leaves the traversal
//    loop, votes (ballot) which lanes need work and refills them from a warp-private chunk of the queue (one atomic per
//    CWK_CHUNK rays);
//  * triangle postponing: a triangle group is pushed back on the stack when fewer than 1 / CWK_POSTPONE_DIV of the warp's active
//    lanes have triangles to test, so that the warp stays in the node phase.
// Scheduling only: the closest hit found is the exhaustive one whatever the order (min t, then min id).
// Every alternative that was measured and lost (shared-memory stack, L1 prefetch of the next child, warp-shared triangle tests, triangle
// lookahead, 96-byte nodes / 256-bit loads, extend-side hit/miss partition, two-stream overlap) is recorded in profiles/README.md and
// was removed from the source.
#ifndef CWK_CHUNK
#define CWK_CHUNK 64   // B200 sweep (profiles/README.md): 32-128 within 1 %, 256 -4 %, 1024 -35 %
#endif
#ifndef CWK_ND
#define CWK_ND 2
#endif
#ifndef CWK_NW
#define CWK_NW 8
#endif
#ifndef CWK_POSTPONE_DIV
#define CWK_POSTPONE_DIV 5   // postpone a triangle group when fewer than 1/5 of the warp's active lanes have triangles (integer compare:
#endif                       // the float form of the same 0.2 threshold cost two conversions per triangle iteration, -0.6 % step time)

#define SRC_QUEUE 0    // rays from the float4 queues, result to hit[]
#define SRC_PRIMARY 1  // rays generated from the ray index, result to hit[]
#define SRC_AOS 2      // standalone query: rays from two (R,3) arrays, results to separate t / id arrays (Raycaster.query)
struct AosRays {
    const float* ro;
    const float* rd;
    float* out_t;
    int32_t* out_i;
};
// Rays whose traversal needed more than `stack_cap` stack entries (deep, degenerate hierarchies): their queue slots are appended to `list`
// (at most WF_OVF_CAP of them; *count keeps counting) and their result is written with the sentinel id WF_OVF_ID; k_extend_fixup, launched
// right behind every extend, re-traces exactly those rays with CW_DEEP_STACK entries each and overwrites the result.
#define WF_OVF_CAP (1 << 20)
#define WF_OVF_ID (-2)
struct Overflow {
    int* count;
    int* list;
    int stack_cap;   // <= CW_STACK; smaller values only through drp_debug_set_stack_limit (tests)
};

#ifndef DRP_EXTEND_MINBLOCKS
#define DRP_EXTEND_MINBLOCKS 9
#endif
#ifndef DRP_EXTEND_MINBLOCKS_QUEUE   // secondary bounces: 8 CTAs / 64 registers, no spills (profiles/r2/ab_extend_occupancy_r2.json: 9 -> +0.5 %,
#define DRP_EXTEND_MINBLOCKS_QUEUE 8 // 10 -> +6 % (spills), 7 -> +2 % (too few warps for the node-fetch latency)); bounce 0 (ALU-bound) stays at 9
#endif
#ifndef DRP_SHADE_MINBLOCKS
#define DRP_SHADE_MINBLOCKS 6   // B200 A/B with the interleaved texels (shade ms per step): 4 -> 2.30, 5 -> 2.12, 6 -> 2.04 (80 registers)
#endif

template <int SRC>
__device__ __forceinline__ void fetch_ray(const WfConst& c, const float4* __restrict__ qa, const float4* __restrict__ qb, const AosRays& aos, int k,
                                          Vec3& o, Vec3& d) {
    if (SRC == SRC_AOS) {
        o = v3(__ldg(aos.ro + 3 * (int64_t)k), __ldg(aos.ro + 3 * (int64_t)k + 1), __ldg(aos.ro + 3 * (int64_t)k + 2));
        d = v3(__ldg(aos.rd + 3 * (int64_t)k), __ldg(aos.rd + 3 * (int64_t)k + 1), __ldg(aos.rd + 3 * (int64_t)k + 2));
    } else {
        int ri;
        load_ray<SRC == SRC_PRIMARY>(c, qa, qb, k, o, d, ri);
    }
}
template <int SRC>
__device__ __forceinline__ void store_hit(float2* __restrict__ hit, const AosRays& aos, int k, float t, int id) {
    if (SRC == SRC_AOS) {
        aos.out_t[k] = t;
        aos.out_i[k] = id;
    } else {
        __stcs(hit + k, make_float2(t, __int_as_float(id)));
    }
}

template <int SRC>
__global__ void __launch_bounds__(WF_BLOCK, SRC == SRC_QUEUE ? DRP_EXTEND_MINBLOCKS_QUEUE : DRP_EXTEND_MINBLOCKS) k_extend_cw(const __grid_constant__ WfConst c, const float4* __restrict__ qa, const float4* __restrict__ qb,
                                                        float2* __restrict__ hit, const int* __restrict__ count_ptr, int* __restrict__ cursor, AosRays aos,
                                                        Overflow ovf) {
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
    // Traversal stack: two 32-bit arrays in local memory (L1-resident).  Measured and removed: 8 entries per thread in shared memory (-5 %,
    // profiles/README.md), 64-bit entries (-1.7 %), top entry mirrored in registers (-4 ... -6 %) -- profiles/r2/ab_stack_r2b.json
    uint32_t st_x[CW_STACK], st_y[CW_STACK];
    bool overflow = false;              // this lane's current ray dropped a stack entry
    // the two expansions of the per-slot hit byte as tables (cwbvh.cuh: cw_perm8, cw_spread3x7); arithmetic instead: +4 %
    __shared__ uint8_t s_perm[8 * 256];
    __shared__ uint32_t s_spread[256];
    for (int e = threadIdx.x; e < 8 * 256; e += WF_BLOCK) s_perm[e] = (uint8_t)cw_perm8((uint32_t)e & 0xffu, (uint32_t)e >> 8);
    for (int e = threadIdx.x; e < 256; e += WF_BLOCK) s_spread[e] = cw_spread3x7((uint32_t)e);
    __syncthreads();
    uint32_t tri_base = 0, tri_valid = 0;  // of the node the current triangle group belongs to
    uint32_t oct_row = 0;                  // octinv << 8: this ray's row of s_perm
#define CWK_PUSH(X, Y)                                                     \
    do {                                                                   \
        if (sp < ovf.stack_cap) { st_x[sp] = (X); st_y[sp] = (Y); ++sp; }  \
        else overflow = true;                                              \
    } while (0)
#define CWK_POP(X, Y) do { --sp; (X) = st_x[sp]; (Y) = st_y[sp]; } while (0)
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
                fetch_ray<SRC>(c, qa, qb, aos, k, o, d);
                r = cw_make_ray(o, d, c.cw_bias);
                oct_row = (r.octinv4 & 0xffu) << 8;
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
                    const float4 n0 = __ldg(p), n1 = __ldg(p + 1), n2 = __ldg(p + 2), n3 = __ldg(p + 3), n4 = __ldg(p + 4);
                    const uint32_t hit8 = cw_node_hits(r, n0, n2, n3, n4, t_best * DRP_T_GROW);
                    const uint32_t imask = __float_as_uint(n0.w) >> 24;
                    ng_x = __float_as_uint(n1.x);
                    ng_y = ((uint32_t)s_perm[oct_row + (hit8 & imask)] << 24) | imask;
                    tg_x = base + rel;  // a triangle group is (node index, pending bits): tri_base / V are re-read when it comes off the stack
                    tri_base = __float_as_uint(n1.y);
                    tri_valid = __float_as_uint(n1.z);
                    tg_y = s_spread[hit8 & ~imask] & tri_valid;
                } else {
                    tg_x = ng_x; tg_y = ng_y;  // a postponed triangle group came off the stack
                    ng_x = 0; ng_y = 0;
                    const float4 m1 = __ldg(c.nodes + CW_NODE_F4 * (int64_t)tg_x + 1);
                    tri_base = __float_as_uint(m1.y);
                    tri_valid = __float_as_uint(m1.z);
                }
                const int total_active = __popc(__activemask());
                while (tg_y != 0) {
                    if (CWK_POSTPONE_DIV * __popc(__activemask()) < total_active && sp < ovf.stack_cap) {
                        CWK_PUSH(tg_x, tg_y);  // postpone: too few lanes have triangles
                        break;
                    }
                    const int ti = 31 - __clz(tg_y);
                    tg_y &= ~(1u << ti);
                    leaf_intersect(c.tris, cw_tri_index(tri_base, tri_valid, ti), 1, r.o, r.d, c.eps, t_best, id_best);
                }
                if (ng_y <= 0x00ffffffu) {
                    if (sp > 0) {
                        CWK_POP(ng_x, ng_y);
                    } else {  // ray finished
                        const bool is_hit = t_best < c.p.t_far;
                        if (overflow) {  // (rare) an entry was dropped: hand the ray to k_extend_fixup
                            const int pos = atomicAdd(ovf.count, 1);
                            if (pos < WF_OVF_CAP) ovf.list[pos] = k;
                            store_hit<SRC>(hit, aos, k, c.p.t_far, WF_OVF_ID);
                            overflow = false;
                        } else {
                            store_hit<SRC>(hit, aos, k, is_hit ? t_best : c.p.t_far, is_hit ? id_best : 0);
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
#undef CWK_PUSH
#undef CWK_POP
}

// Re-trace of the rays k_extend_cw flagged (see Overflow): one ray per thread, stack of CW_DEEP_STACK entries per thread in global memory
// ([entry][thread], coalesced).  The wide hierarchy is at most as deep as the binary one (<= 63 Morton + 27 index bits), a ray stacks at most one
// node group per level, and nothing is postponed here, so CW_DEEP_STACK = 256 entries cannot run out; should it happen regardless, the handle's
// sticky error flag is raised (host-mapped: every later call on the handle fails loudly).
// When more than WF_OVF_CAP rays were flagged, the list is incomplete and the kernel scans all results for the sentinel instead.
#define WF_FIXUP_BLOCKS 148
template <int SRC>
__global__ void __launch_bounds__(WF_BLOCK) k_extend_fixup(const __grid_constant__ WfConst c, const float4* __restrict__ qa, const float4* __restrict__ qb,
                                                           float2* __restrict__ hit, const int* __restrict__ count_ptr, AosRays aos, Overflow ovf,
                                                           uint2* __restrict__ deep_stack, int* __restrict__ sticky) {
    const int n_ovf = *ovf.count;
    if (n_ovf == 0) return;
    const int count = SRC == SRC_QUEUE ? *count_ptr : (int)c.R;
    const bool scan = n_ovf > WF_OVF_CAP;
    const int n_items = scan ? count : n_ovf;
    const int tid = blockIdx.x * blockDim.x + threadIdx.x, nthreads = gridDim.x * blockDim.x;
    uint2* st = deep_stack + tid;
    for (int item = tid; item < n_items; item += nthreads) {
        int k = item;
        if (scan) {
            const int id = SRC == SRC_AOS ? aos.out_i[k] : __float_as_int(hit[k].y);
            if (id != WF_OVF_ID) continue;
        } else {
            k = ovf.list[item];
        }
        Vec3 o, d;
        fetch_ray<SRC>(c, qa, qb, aos, k, o, d);
        bool overflow = false;
        const RayHit h = cw_trace_one_stack(c.nodes, c.tris, o, d, c.p.t_far, c.eps, CwStridedStack{st, (int64_t)nthreads}, CW_DEEP_STACK, overflow);
        store_hit<SRC>(hit, aos, k, h.t, h.id);
        if (overflow) atomicAdd(sticky, 1);
    }
}

__device__ __forceinline__ void accum_add4(float* p, float a, float b, float c, float d) {
    atomicAdd(reinterpret_cast<float4*>(p), make_float4(a, b, c, d));  // RED.E.ADD.F32x4 (sm_90+)
}

// One kernel per bounce walks the ray queue: hits (surface + BRDF) and misses (environment) share warps.  Measured alternatives that lost
// and were removed (profiles/README.md section 3, profiles/r2/ab_tripipe_chunkpart_r2a.json): separate hit / miss kernels over index lists written
// by the extend kernel (-5 %), hit-first / miss-second walk inside a warp's chunk (+-0), software-pipelined or prefetched loads (-3 ... -7 %).
template <bool PRIMARY>
__global__ void __launch_bounds__(WF_BLOCK, DRP_SHADE_MINBLOCKS) k_shade(const __grid_constant__ WfConst c, int bounce, const float4* __restrict__ qa,
                                                    const float4* __restrict__ qb, const float4* __restrict__ qt, const float2* __restrict__ hit,
                                                    float4* __restrict__ oa, float4* __restrict__ ob, float4* __restrict__ ot,
                                                    const int* __restrict__ count_ptr, int* __restrict__ out_count, int* __restrict__ cursor) {
    const int count = PRIMARY ? (int)c.R : *count_ptr;
    const int lane = threadIdx.x & 31;
    const bool last = bounce == c.p.ray_depth - 1;
    const bool always_sky = last && c.p.last_bounce_skybox;  // path_tracing.py:260
    for (;;) {
        int base = 0;
        if (lane == 0) base = atomicAdd(cursor, WF_FETCH);
        base = __shfl_sync(0xffffffffu, base, 0);
        if (base >= count) break;
#pragma unroll 1
        for (int j = 0; j < WF_FETCH; j += 32) {
            const int k = base + j + lane;  // queue slot of the ray
            bool alive = false;
            Vec3 no = v3(0, 0, 0), nd = v3(0, 0, 0), T = v3(1, 1, 1);
            int ri = 0;
            if (k < count) {
                Vec3 o, d;
                load_ray<PRIMARY>(c, qa, qb, k, o, d, ri);
                if (!PRIMARY) { float4 t4 = __ldcs(qt + k); T = v3(t4.x, t4.y, t4.z); }   // queue / hit entries: touched once, streaming
                const float2 h = __ldcs(hit + k);
                const float t = h.x;
                const bool is_hit = t < c.p.t_far;
                SurfaceAttrs s;
                if (is_hit) {
                    // last bounce of a path (not the first): only emission and alpha reach the outputs, skip the rest of the material
                    s = surface_attrs(c.scene, c.scene.materials, o + d * t, __float_as_int(h.y), !PRIMARY && last, nullptr);
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
            // warp-aggregated append to the output queue
            unsigned m = __ballot_sync(0xffffffffu, alive);
            if (m) {
                int pos = 0;
                if (lane == __ffs(m) - 1) pos = atomicAdd(out_count, __popc(m));
                pos = __shfl_sync(0xffffffffu, pos, __ffs(m) - 1);
                if (alive) {
                    int w = pos + __popc(m & ((1u << lane) - 1u));
                    __stcs(oa + w, make_float4(no.x, no.y, no.z, nd.x));
                    __stcs(ob + w, make_float4(nd.y, nd.z, __int_as_float(ri), 0.0f));
                    __stcs(ot + w, make_float4(T.x, T.y, T.z, 0.0f));
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

void drp_invalidate_scene_box(BvhHandle* h) { if (h->ws) h->ws->have_box = false; }

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
        h->ws->n_counters = 1024;  // [0,64) live counts per bounce, [64,192) fetch cursors, [192,256) flagged-ray counts per bounce, [1020] / [1021] drp_trace
        DRP_CUDA_CHECK(cudaMalloc((void**)&h->ws->counters, sizeof(int) * 1024));
        DRP_CUDA_CHECK(cudaMalloc((void**)&h->ws->d_traced, sizeof(unsigned long long)));
        int nb = 0;
        DRP_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_extend_cw<SRC_PRIMARY>, WF_BLOCK, 0));
        h->ws->grid_extend[1] = std::max(1, nb) * h->ws->sm_count;
        DRP_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_extend_cw<SRC_QUEUE>, WF_BLOCK, 0));
        h->ws->grid_extend[0] = std::max(1, nb) * h->ws->sm_count;
        DRP_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_shade<true>, WF_BLOCK, 0));
        h->ws->grid_shade[1] = std::max(1, nb) * h->ws->sm_count;
        DRP_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_shade<false>, WF_BLOCK, 0));
        h->ws->grid_shade[0] = std::max(1, nb) * h->ws->sm_count;
        // deep-traversal fix-up (k_extend_fixup): list of flagged rays + one CW_DEEP_STACK-entry stack per fix-up thread (39 MB)
        DRP_CUDA_CHECK(cudaMalloc((void**)&h->ws->ovf_list, sizeof(int) * WF_OVF_CAP));
        DRP_CUDA_CHECK(cudaMalloc((void**)&h->ws->deep_stack, sizeof(uint2) * (size_t)CW_DEEP_STACK * WF_FIXUP_BLOCKS * WF_BLOCK));
    }
    RenderWorkspace* ws = h->ws;
    if (rays > 0 && rays > ws->capacity) {
        for (int k = 0; k < 2; ++k) { cudaFree(ws->qa[k]); cudaFree(ws->qb[k]); cudaFree(ws->qt[k]); }
        cudaFree(ws->hit);
        ws->capacity = 0;
        for (int k = 0; k < 2; ++k) {
            DRP_CUDA_CHECK(cudaMalloc((void**)&ws->qa[k], sizeof(float4) * rays));
            DRP_CUDA_CHECK(cudaMalloc((void**)&ws->qb[k], sizeof(float4) * rays));
            DRP_CUDA_CHECK(cudaMalloc((void**)&ws->qt[k], sizeof(float4) * rays));
        }
        DRP_CUDA_CHECK(cudaMalloc((void**)&ws->hit, sizeof(float2) * rays));
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

// Sticky failure flag of a handle (host-mapped, written by k_extend_fixup when even the deep stack ran out): checked, without a
// synchronisation, at the start of every call on the handle and by drp_status().
int drp_check_sticky(BvhHandle* h, const char* who) {
    if (h->sticky_host && *(volatile int*)h->sticky_host != 0) {
        drp_set_error(std::string(who) + ": an earlier traversal on this handle ran out of its " + std::to_string(CW_DEEP_STACK) +
                      "-entry deep stack (" + std::to_string(*(volatile int*)h->sticky_host) + " rays): results of that call are invalid");
        return DRP_ERR_INVALID;
    }
    return DRP_OK;
}

struct NvtxRange {   // host-side ranges for nsys / ncu --nvtx: drp_render > batch > bounce (free when no tool is attached)
    explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
};

extern "C" int drp_render(uint64_t handle, const drp_scene_t* scene, const drp_render_params_t* params, float* accum, void* stream) {
    BvhHandle* h = drp_lookup(handle);
    if (!h) { drp_set_error("drp_render: unknown handle"); return DRP_ERR_HANDLE; }
    if (!scene || !params || !accum) { drp_set_error("drp_render: null argument"); return DRP_ERR_INVALID; }
    if (int rc = drp_check_sticky(h, "drp_render")) return rc;
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
    if (HW * p.n_samples >= (int64_t(1) << 31)) {
        // ray indices are 32-bit: render the samples in several passes (the native RNG is keyed by pixel and GLOBAL sample id, the jitter
        // tables are per sample, so the split changes nothing but the accumulation order); the reference sections the same way
        // (path_tracing.py:318-325)
        if (p.rng_mode == DRP_RNG_REPLAY) { drp_set_error("drp_render: more than 2^31 rays per call in replay mode; render the samples in several calls"); return DRP_ERR_INVALID; }
        const int per_pass = (int)std::max<int64_t>(1, ((int64_t(1) << 31) - 1) / HW);
        int64_t launches = 0, nominal = 0;
        for (int s0 = 0; s0 < p.n_samples; s0 += per_pass) {
            drp_render_params_t q = p;
            q.n_samples = std::min(per_pass, p.n_samples - s0);
            q.jitter_x = p.jitter_x + s0; q.jitter_y = p.jitter_y + s0; q.sample_ids = p.sample_ids + s0;
            if (int rc = drp_render(handle, scene, &q, accum, stream)) return rc;
            launches += h->last_render.kernel_launches; nominal += h->last_render.rays_nominal;
        }
        h->last_render.kernel_launches = launches; h->last_render.rays_nominal = nominal;  // (rays_traced: last pass only)
        return DRP_OK;
    }
    DeviceGuard guard(h->device);
    if (!guard.ok) { drp_set_error("drp_render: cannot select device"); return DRP_ERR_CUDA; }
    cudaStream_t s = (cudaStream_t)stream;
    NvtxRange range_call("drp_render");
    const int spb = p.reproducible ? 1 : (int)std::max<int64_t>(1, std::min<int64_t>(p.n_samples, WF_MAX_BATCH_RAYS / HW));  // samples per batch
    int rc = ensure_workspace(h, (int64_t)spb * HW + 64, scene->n_materials);
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
    c.R_total = HW * p.n_samples;
    c.accum = accum;
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
    int64_t launches = 0;
    h->last_render.rays_nominal = HW * p.n_samples * D;
    DRP_CUDA_CHECK(cudaMemsetAsync(ws->d_traced, 0, sizeof(unsigned long long), s));
    for (int s0 = 0; s0 < p.n_samples; s0 += spb) {
        const int ns = std::min(spb, p.n_samples - s0);
        NvtxRange range_batch("batch");
        DRP_CUDA_CHECK(cudaMemsetAsync(ws->counters, 0, sizeof(int) * ws->n_counters, s));  // (a memset node, not counted as a kernel launch)
        c.R = (int64_t)ns * HW;
        c.ray_base = (int64_t)s0 * HW;
        // primary queue pixel-major / sample-minor: the samples of a pixel sit in adjacent lanes (A/B on config 3: 8.41 -> 7.57 ms per step;
        // 8x4-pixel blocks on top of it: extend unchanged, shade 5 % slower -- removed)
        c.sample_minor = ns > 1 ? ns : 0;
        int* counts = ws->counters;         // [0..D]
        int* cursors = ws->counters + 64;   // [0..2D)
        float2* hit = ws->hit;
        for (int b = 0; b < D; ++b) {
            NvtxRange range_bounce(b == 0 ? "bounce 0" : (b == 1 ? "bounce 1" : (b == 2 ? "bounce 2" : "bounce 3+")));
            const Overflow ovf = {ws->counters + 192 + b, ws->ovf_list, h->stack_cap};
            const int in = b & 1, out = in ^ 1;
            float4 *qa_in = ws->qa[in], *qb_in = ws->qb[in], *qt_in = ws->qt[in];
            float4 *qa_out = ws->qa[out], *qb_out = ws->qb[out], *qt_out = ws->qt[out];
            const int ge = ws->grid_extend[b == 0], gs = ws->grid_shade[b == 0];
            auto span_begin = [&](int kind, const int* count_ptr) -> int {
                if (!ws->profiling || ws->live_used >= ws->live_capacity) return -1;
                RenderWorkspace::Span sp;
                sp.a = wf_event(ws); sp.b = wf_event(ws); sp.kind = kind; sp.bounce = b;
                sp.traced_slot = ws->d_live + ws->live_used++;
                k_record_live<<<1, 1, 0, s>>>(count_ptr, (unsigned long long)c.R, sp.traced_slot);
                ++launches;
                cudaEventRecord(sp.a, s);
                ws->spans.push_back(sp);
                return (int)ws->spans.size() - 1;
            };
            auto span_end = [&](int idx) { if (idx >= 0) cudaEventRecord(ws->spans[idx].b, s); };
            if (b == 0) {
                int sp = span_begin(0, nullptr);
                k_extend_cw<SRC_PRIMARY><<<ge, WF_BLOCK, 0, s>>>(c, nullptr, nullptr, hit, nullptr, cursors + 0, AosRays(), ovf);
                k_extend_fixup<SRC_PRIMARY><<<WF_FIXUP_BLOCKS, WF_BLOCK, 0, s>>>(c, nullptr, nullptr, hit, nullptr, AosRays(), ovf, ws->deep_stack, h->sticky_dev);
                span_end(sp);
                if (p.shade_wait_event && s0 == 0) DRP_CUDA_CHECK(cudaStreamWaitEvent(s, (cudaEvent_t)p.shade_wait_event, 0));
                sp = span_begin(1, nullptr);
                k_shade<true><<<gs, WF_BLOCK, 0, s>>>(c, b, nullptr, nullptr, nullptr, hit, qa_out, qb_out, qt_out, nullptr, counts + 1, cursors + 1);
                span_end(sp);
            } else {
                int sp = span_begin(0, counts + b);
                k_extend_cw<SRC_QUEUE><<<ge, WF_BLOCK, 0, s>>>(c, qa_in, qb_in, hit, counts + b, cursors + 2 * b, AosRays(), ovf);
                k_extend_fixup<SRC_QUEUE><<<WF_FIXUP_BLOCKS, WF_BLOCK, 0, s>>>(c, qa_in, qb_in, hit, counts + b, AosRays(), ovf, ws->deep_stack, h->sticky_dev);
                span_end(sp);
                sp = span_begin(1, counts + b);
                k_shade<false><<<gs, WF_BLOCK, 0, s>>>(c, b, qa_in, qb_in, qt_in, hit, qa_out, qb_out, qt_out, counts + b, counts + b + 1, cursors + 2 * b + 1);
                span_end(sp);
            }
            launches += 3;
        }
        k_count_traced<<<1, 1, 0, s>>>(counts, D, (unsigned long long)c.R, ws->d_traced);  // read lazily by drp_render_stats
        ++launches;
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
    c.nodes = h->nodes; c.tris = h->packed; c.eps = h->eps; c.R = n; c.p.t_far = t_far; c.cw_bias = 0x47000000u;
    int* cursor = ws->counters + 1020;
    DRP_CUDA_CHECK(cudaMemsetAsync(cursor, 0, 2 * sizeof(int), s));
    const AosRays aos = {ro, rd, out_t, out_i};
    const Overflow ovf = {ws->counters + 1021, ws->ovf_list, h->stack_cap};
    k_extend_cw<SRC_AOS><<<ws->grid_extend[1], WF_BLOCK, 0, s>>>(c, nullptr, nullptr, nullptr, nullptr, cursor, aos, ovf);
    k_extend_fixup<SRC_AOS><<<WF_FIXUP_BLOCKS, WF_BLOCK, 0, s>>>(c, nullptr, nullptr, nullptr, nullptr, aos, ovf, ws->deep_stack, h->sticky_dev);
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
    if (int rc = drp_check_sticky(h, "drp_surface_attrs")) return rc;
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
    if (int rc = drp_check_sticky(h, "drp_render_stats")) return rc;
    return DRP_OK;
}

#define DRP_STR2(x) #x
#define DRP_STR(x) DRP_STR2(x)
extern "C" const char* drp_build_config(void) {
    return "compiled " __DATE__ " " __TIME__ "; DRP_EXTEND_MINBLOCKS=" DRP_STR(DRP_EXTEND_MINBLOCKS) " DRP_EXTEND_MINBLOCKS_QUEUE=" DRP_STR(DRP_EXTEND_MINBLOCKS_QUEUE) " DRP_SHADE_MINBLOCKS=" DRP_STR(DRP_SHADE_MINBLOCKS)
           " CWK_CHUNK=" DRP_STR(CWK_CHUNK) " CWK_ND=" DRP_STR(CWK_ND) " CWK_NW=" DRP_STR(CWK_NW) " CWK_POSTPONE_DIV=" DRP_STR(CWK_POSTPONE_DIV)
           " CW_STACK=" DRP_STR(CW_STACK) " CW_DEEP_STACK=" DRP_STR(CW_DEEP_STACK);
}

extern "C" int drp_status(uint64_t handle) {
    BvhHandle* h = drp_lookup(handle);
    if (!h) { drp_set_error("drp_status: unknown handle"); return DRP_ERR_HANDLE; }
    return drp_check_sticky(h, "drp_status");
}

extern "C" int drp_debug_set_stack_limit(uint64_t handle, int entries) {
    BvhHandle* h = drp_lookup(handle);
    if (!h) { drp_set_error("drp_debug_set_stack_limit: unknown handle"); return DRP_ERR_HANDLE; }
    if (entries < 0 || entries > CW_STACK) { drp_set_error("drp_debug_set_stack_limit: entries must be in [0, " DRP_STR(CW_STACK) "]"); return DRP_ERR_INVALID; }
    h->stack_cap = entries;
    return DRP_OK;
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
