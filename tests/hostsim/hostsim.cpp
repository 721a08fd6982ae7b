// hostsim.cpp -- serial host build of the kernel logic in diffrp_b200/csrc/*.cuh (compiled with -DDRP_HOSTSIM).
// TEST INFRASTRUCTURE ONLY: lets the `-m "not gpu"` tests exercise the exact per-thread code of the CUDA kernels
// (LBVH phases, traversal, shading) in a container without a GPU.  It is never loaded by the product package.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <numeric>
#include <omp.h>
#include "../../diffrp_b200/csrc/common.cuh"
#include "../../diffrp_b200/csrc/lbvh.cuh"
#include "../../diffrp_b200/csrc/traverse.cuh"
#include "../../diffrp_b200/csrc/cwbvh.cuh"
#include "../../diffrp_b200/csrc/instance_level.h"
#include "../../diffrp_b200/csrc/tonemap.cuh"

struct HsBvh {
    int n;
    std::vector<float4> nodes, packed;
    std::vector<float4> cw_nodes, cw_tris;  // wide layout (cwbvh.cuh)
    int cw_count = 0, cw_levels = 0;
    std::vector<int> level_begin;   // wide nodes of level L: [level_begin[L], level_begin[L + 1])
    std::vector<uint32_t> bounds_ord;
    float sah;
    float bounds[6];
};

static HsBvh* hs_build_impl(const float* verts, const int32_t* tris, int64_t n_tris, int max_leaf, bool wide);
extern "C" HsBvh* hs_build(const float* verts, const int32_t* tris, int64_t n_verts, int64_t n_tris) {
    (void)n_verts;
    return hs_build_impl(verts, tris, n_tris, DRP_MAX_LEAF, false);
}
extern "C" HsBvh* hs_build_wide(const float* verts, const int32_t* tris, int64_t n_verts, int64_t n_tris) {
    (void)n_verts;
    return hs_build_impl(verts, tris, n_tris, CW_MAX_LEAF, true);
}
static HsBvh* hs_build_impl(const float* verts, const int32_t* tris, int64_t n_tris, int max_leaf, bool wide) {
    HsBvh* h = new HsBvh();
    const int n = (int)n_tris;
    const size_t nn = n > 0 ? n : 1;
    h->n = n;
    h->nodes.resize(4 * (n > 1 ? n - 1 : 1));
    h->packed.resize(3 * nn);
    std::vector<uint32_t> bounds(12);
    for (int i = 0; i < 12; ++i) bounds[i] = ((i / 3) % 2 == 0) ? 0xffffffffu : 0u;
    std::vector<float4> prim_lo(nn), prim_hi(nn), box_lo(2 * nn), box_hi(2 * nn);
    std::vector<uint64_t> keys(nn);
    std::vector<uint32_t> vals(nn);
    std::vector<int> left(nn), right(nn), parent(2 * nn), rf(nn), rl(nn), arrive(nn, 0);
    std::vector<uint8_t> collapsed(nn, 0);
    LbvhBuild b;
    memset(&b, 0, sizeof(b));
    b.verts = verts; b.tris = tris; b.n = n; b.max_leaf = max_leaf; b.bounds = bounds.data();
    b.prim_lo = prim_lo.data(); b.prim_hi = prim_hi.data(); b.keys = keys.data(); b.vals = vals.data();
    b.left = left.data(); b.right = right.data(); b.parent = parent.data(); b.range_first = rf.data(); b.range_last = rl.data();
    b.box_lo = box_lo.data(); b.box_hi = box_hi.data(); b.arrive = arrive.data(); b.collapsed = collapsed.data();
    b.nodes = h->nodes.data(); b.packed = h->packed.data();
    std::vector<float> dp_cost;
    std::vector<uint8_t> dp_dec;
    if (wide && getenv("HS_GREEDY") == nullptr) {
        dp_cost.assign(16 * nn, 0.0f); dp_dec.assign(16 * nn, 0);
        b.dp_cost = dp_cost.data(); b.dp_dec = dp_dec.data();
    }
    for (int i = 0; i < n; ++i) {
        Vec3 lo, hi;
        lbvh_prim_bounds(b, i, lo, hi);
        float v[12] = {lo.x, lo.y, lo.z, hi.x, hi.y, hi.z, 0.5f * (lo.x + hi.x), 0.5f * (lo.y + hi.y), 0.5f * (lo.z + hi.z), 0, 0, 0};
        v[9] = v[6]; v[10] = v[7]; v[11] = v[8];
        for (int k = 0; k < 12; ++k) {
            bool is_min = (k / 3) % 2 == 0;
            uint32_t o = f2ord(v[k]);
            bounds[k] = is_min ? std::min(bounds[k], o) : std::max(bounds[k], o);
        }
    }
    for (int i = 0; i < n; ++i) lbvh_morton(b, i);
    {   // phase 3: stable sort by key (cub radix sort is stable as well)
        std::vector<uint32_t> order(n);
        std::iota(order.begin(), order.end(), 0u);
        std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t c) { return keys[a] < keys[c]; });
        std::vector<uint64_t> k2(nn);
        std::vector<uint32_t> v2(nn);
        for (int i = 0; i < n; ++i) { k2[i] = keys[order[i]]; v2[i] = vals[order[i]]; }
        keys.swap(k2); vals.swap(v2);
        b.keys = keys.data(); b.vals = vals.data();
    }
    std::vector<int> count(2 * nn, 0);
    b.count = count.data();
    if (n > 1) {
        for (int i = 0; i < n - 1; ++i) lbvh_karras(b, i);
        for (int j = 0; j < n; ++j) lbvh_refit(b, j, [](int* p) { int o = *p; *p = o + 1; return o; }, []() {});
        if (!wide) for (int i = 0; i < n - 1; ++i) lbvh_emit(b, i);
        h->sah = box_lo[0].w / std::max(box_hi[0].w, 1e-30f);
    } else {
        lbvh_emit_tiny(b);
        h->sah = 1.0f;
    }
    if (!wide) for (int j = 0; j < n; ++j) lbvh_pack_tri(b, j);
    for (int k = 0; k < 6; ++k) h->bounds[k] = ord2f(bounds[k]);
    h->bounds_ord = bounds;
    if (wide) {
        CwBuild cw;
        cw.b = b;
        cw.capacity = n > 1 ? n - 1 : 1;
        h->cw_nodes.resize(CW_NODE_F4 * (size_t)cw.capacity);
        h->cw_tris.resize(3 * nn);
        std::vector<int> work(cw.capacity, 0), counters(128, 0);
        cw.cw_nodes = h->cw_nodes.data(); cw.cw_tris = h->cw_tris.data(); cw.work = work.data(); cw.counters = counters.data();
        if (n < 2) {
            cw_emit_tiny(cw);
            h->cw_count = 1;
        } else {
            counters[0] = 1;
            int begin = 0, end = 1, levels = 0;
            while (begin < end) {
                h->level_begin.push_back(begin);
                for (int ni = begin; ni < end; ++ni) cw_collapse_node(cw, ni, [](int* p, int v) { int o = *p; *p = o + v; return o; });
                begin = end; end = counters[0]; ++levels;
            }
            h->level_begin.push_back(counters[0]);
            h->cw_count = counters[0];
            h->cw_levels = levels;
        }
    }
    return h;
}

extern "C" int64_t hs_trace_wide(const HsBvh* h, const float* ro, const float* rd, int64_t n, float t_far, float eps, float* out_t, int32_t* out_i) {
    int64_t overflow = 0;
#pragma omp parallel for schedule(dynamic, 256) reduction(+ : overflow)
    for (int64_t r = 0; r < n; ++r) {
        bool of = false;
        RayHit hit = cw_trace_one(h->cw_nodes.data(), h->cw_tris.data(), v3(ro[3 * r], ro[3 * r + 1], ro[3 * r + 2]),
                                  v3(rd[3 * r], rd[3 * r + 1], rd[3 * r + 2]), t_far, eps, of);
        out_t[r] = hit.t;
        out_i[r] = hit.id;
        overflow += of;
    }
    return overflow;
}

// The walk of k_extend_fixup (wavefront.cu): the same traversal over a caller-supplied stack of `cap` entries laid out [entry][thread] like the
// kernel's global-memory deep stack (CwStridedStack), here with one "thread" per OpenMP thread.  out_overflow[r] = 1 when ray r had to drop an
// entry (its result is then not trustworthy -- what the kernels hand to the next stage / report through the sticky flag); *max_depth = the
// deepest stack use over all rays, measured with an instrumented stack of the same interface.
struct HsCountingStack {
    uint2* base;
    int64_t stride;
    int* deepest;
    void put(int i, uint32_t a, uint32_t b) const { uint2 v; v.x = a; v.y = b; base[i * stride] = v; if (i + 1 > *deepest) *deepest = i + 1; }
    void get(int i, uint32_t& a, uint32_t& b) const { const uint2 v = base[i * stride]; a = v.x; b = v.y; }
};
extern "C" int64_t hs_trace_wide_strided(const HsBvh* h, const float* ro, const float* rd, int64_t n, float t_far, float eps, int cap, float* out_t,
                                         int32_t* out_i, uint8_t* out_overflow, int* max_depth) {
    int64_t overflow = 0;
    int deepest_all = 0;
    const int64_t nthreads = 8;
    std::vector<uint2> stack((size_t)(cap > 0 ? cap : 1) * nthreads);
#pragma omp parallel for schedule(dynamic, 256) num_threads(8) reduction(+ : overflow) reduction(max : deepest_all)
    for (int64_t r = 0; r < n; ++r) {
        bool of = false;
        int deepest = 0;
        const int tid = omp_get_thread_num();
        const RayHit hit = cw_trace_one_stack(h->cw_nodes.data(), h->cw_tris.data(), v3(ro[3 * r], ro[3 * r + 1], ro[3 * r + 2]),
                                              v3(rd[3 * r], rd[3 * r + 1], rd[3 * r + 2]), t_far, eps, HsCountingStack{stack.data() + tid, nthreads, &deepest},
                                              cap, of);
        out_t[r] = hit.t;
        out_i[r] = hit.id;
        if (out_overflow) out_overflow[r] = of ? 1 : 0;
        overflow += of;
        if (deepest > deepest_all) deepest_all = deepest;
    }
    if (max_depth) *max_depth = deepest_all;
    return overflow;
}

#if DRP_CW_V2
// node format 2 helpers exactly as the kernels build their shared-memory tables: perm[octinv * 256 + byte], spread[byte], and the
// triangle index behind bit j of a node's triangle mask
extern "C" void hs_cw_tables(uint8_t* perm, uint32_t* spread) {
    for (int e = 0; e < 8 * 256; ++e) perm[e] = (uint8_t)cw_perm8((uint32_t)e & 0xffu, (uint32_t)e >> 8);
    for (int e = 0; e < 256; ++e) spread[e] = cw_spread3x7((uint32_t)e);
}
extern "C" int hs_cw_tri_index(uint32_t tri_base, uint32_t vmask, int j) { return cw_tri_index(tri_base, vmask, j); }
#endif

// wide-layout statistics: [nodes, levels, inner children, leaf children, triangles referenced, max tris per node]
extern "C" void hs_stats_wide(const HsBvh* h, int64_t* out) {
    int64_t inner = 0, leaf = 0, tris = 0, maxt = 0;
    for (int ni = 0; ni < h->cw_count; ++ni) {
        const float4* p = h->cw_nodes.data() + CW_NODE_F4 * (size_t)ni;
        int64_t nt = 0;
#if DRP_CW_V2
        for (int s = 0; s < 8; ++s) {
            const int kind = cw_slot_kind(p, s);
            if (kind < 0) ++inner;
            else if (kind > 0) { ++leaf; nt += kind; }
        }
#else
        uint32_t m[2] = {f2u(p[1].z), f2u(p[1].w)};
        for (int s = 0; s < 8; ++s) {
            uint32_t meta = (m[s / 4] >> (8 * (s % 4))) & 0xffu;
            if (meta == 0) continue;
            if ((meta & 0x18u) == 0x18u && (meta >> 5) == 1u) ++inner;
            else { ++leaf; nt += cw_popc(meta >> 5); }
        }
#endif
        tris += nt;
        maxt = std::max(maxt, nt);
    }
    out[0] = h->cw_count; out[1] = h->cw_levels; out[2] = inner; out[3] = leaf; out[4] = tris; out[5] = maxt;
}

extern "C" void hs_free(HsBvh* h) { delete h; }

extern "C" int64_t hs_trace(const HsBvh* h, const float* ro, const float* rd, int64_t n, float t_far, float eps, float* out_t, int32_t* out_i) {
    int64_t overflow = 0;
#pragma omp parallel for schedule(dynamic, 256) reduction(+ : overflow)
    for (int64_t r = 0; r < n; ++r) {
        bool of = false;
        RayHit hit = trace_one(h->nodes.data(), h->packed.data(), v3(ro[3 * r], ro[3 * r + 1], ro[3 * r + 2]),
                               v3(rd[3 * r], rd[3 * r + 1], rd[3 * r + 2]), t_far, eps, of);
        out_t[r] = hit.t;
        out_i[r] = hit.id;
        overflow += of;
    }
    return overflow;
}

// structure statistics: live nodes, leaves, max depth, triangles covered exactly once
extern "C" void hs_stats(const HsBvh* h, int64_t* out /* [n_nodes, n_leaves, max_depth, tris_in_leaves, max_leaf] */, float* sah) {
    std::vector<std::pair<int, int>> stack{{0, 1}};
    int64_t nodes = 0, leaves = 0, tris = 0;
    int md = 0, ml = 0;
    std::vector<uint8_t> seen(h->n > 0 ? h->n : 1, 0);
    bool dup = false;
    while (!stack.empty()) {
        auto [node, depth] = stack.back();
        stack.pop_back();
        md = std::max(md, depth);
        if (node < 0) {
            int first, count;
            leaf_decode(node, first, count);
            if (count > 0) ++leaves;
            tris += count;
            ml = std::max(ml, count);
            for (int k = 0; k < count; ++k) { if (seen[first + k]) dup = true; seen[first + k] = 1; }
            continue;
        }
        ++nodes;
        float4 n3 = h->nodes[(size_t)node * 4 + 3];
        stack.push_back({f2i(n3.x), depth + 1});
        stack.push_back({f2i(n3.y), depth + 1});
    }
    out[0] = nodes; out[1] = leaves; out[2] = md; out[3] = dup ? -1 : tris; out[4] = ml;
    *sah = h->sah;
}

// ---- shading: the per-ray body of k_extend + k_shade run serially, without compaction -------------------------
#include "../../diffrp_b200/csrc/shade.cuh"

extern "C" int64_t hs_render(const HsBvh* h, const drp_scene_t* scene, const drp_render_params_t* pp, float eps, float* accum) {
    const drp_render_params_t& p = *pp;
    const bool tiled = p.tile_w > 0 && p.tile_h > 0;
    const int tw = tiled ? p.tile_w : p.width, th = tiled ? p.tile_h : p.height, tx0 = tiled ? p.tile_x0 : 0, ty0 = tiled ? p.tile_y0 : 0;
    const int HW = tw * th;
    const int64_t R_total = (int64_t)HW * p.n_samples;
    int64_t traced = 0;
#pragma omp parallel for schedule(dynamic, 64) reduction(+ : traced)
    for (int lpix = 0; lpix < HW; ++lpix) {
        for (int s = 0; s < p.n_samples; ++s) {
            const int ri = s * HW + lpix;
            int y = ty0 + lpix / tw, x = tx0 + lpix % tw;
            const int pix = y * p.width + x;
            Vec3 o, d, T = v3(1, 1, 1);
            gen_primary_ray(p.inv_vp, p.cam_pos, p.t_near, p.ndc_x[x] + p.jitter_x[s], p.ndc_y[y] + p.jitter_y[s], o, d);
            for (int b = 0; b < p.ray_depth; ++b) {
                bool of = false;
                RayHit hit = trace_one(h->nodes.data(), h->packed.data(), o, d, p.t_far, eps, of);
                ++traced;
                const bool is_hit = hit.t < p.t_far;
                const bool last = b == p.ray_depth - 1;
                const bool always_sky = last && p.last_bounce_skybox;
                SurfaceAttrs sa;
                if (is_hit) sa = surface_attrs(*scene, scene->materials, o + d * hit.t, hit.id);
                else { sa.albedo = sa.normal = sa.emission = v3(0, 0, 0); sa.metal = sa.smooth = sa.alpha = 0.0f; }
                Vec3 env = (always_sky || !is_hit) ? env_fetch(scene->env, d) : v3(0, 0, 0);
                float u[6];
                if (p.rng_mode == DRP_RNG_REPLAY) for (int q = 0; q < 6; ++q) u[q] = p.replay_u[((int64_t)b * 6 + q) * R_total + ri];
                else philox_uniform6(p.seed, (uint32_t)pix, (uint32_t)p.sample_ids[s], (uint32_t)b, u);
                BounceOut r = brdf_sample(sa, hit.t, o, d, env, u);
                float* acc = accum + (int64_t)DRP_ACCUM_CHANNELS * pix;
                acc[0] += T.x * r.radiance.x; acc[1] += T.y * r.radiance.y; acc[2] += T.z * r.radiance.z; acc[3] += sa.alpha;
                if (b == 0) {
                    acc[4] += sa.albedo.x; acc[5] += sa.albedo.y; acc[6] += sa.albedo.z;
                    acc[7] += sa.emission.x; acc[8] += sa.emission.y; acc[9] += sa.emission.z;
                    acc[10] += sa.normal.x; acc[11] += sa.normal.y; acc[12] += sa.normal.z;
                    acc[13] += r.hit_pos.x; acc[14] += r.hit_pos.y; acc[15] += r.hit_pos.z;
                }
                T = is_hit ? T * r.transfer : v3(0, 0, 0);
                d = r.next_d;
                o = r.hit_pos + d * p.step_epsilon;
            }
        }
    }
    return traced;
}

// mean nodes fetched / triangles tested per ray (tree-quality probe; serial)
extern "C" void hs_trav_stats(const HsBvh* h, int wide, const float* ro, const float* rd, int64_t n, float t_far, float eps, double* out) {
    g_trav.nodes = g_trav.tris = 0;
    for (int64_t r = 0; r < n; ++r) {
        bool of = false;
        Vec3 o = v3(ro[3 * r], ro[3 * r + 1], ro[3 * r + 2]), d = v3(rd[3 * r], rd[3 * r + 1], rd[3 * r + 2]);
        if (wide) cw_trace_one(h->cw_nodes.data(), h->cw_tris.data(), o, d, t_far, eps, of);
        else trace_one(h->nodes.data(), h->packed.data(), o, d, t_far, eps, of);
    }
    out[0] = (double)g_trav.nodes / (double)n;
    out[1] = (double)g_trav.tris / (double)n;
}

// colour epilogue: the per-pixel function of k_tonemap (csrc/tonemap.cuh), same output contract as drp_tonemap
extern "C" void hs_tonemap(const float* src, int64_t height, int64_t width, const drp_tonemap_params_t* pp, uint8_t* out_u8, float* out_f32) {
    const drp_tonemap_params_t p = *pp;
    const int C = p.alpha_offset >= 0 ? 4 : 3;
    for (int64_t i = 0; i < height * width; ++i) {
        const int64_t row = i / width, col = i - row * width;
        const int64_t o = (p.flip_rows ? height - 1 - row : row) * width + col;
        float v[4];
        tm_pixel(src + i * p.in_stride, p, v);
        for (int ch = 0; ch < C; ++ch) {
            if (out_f32) out_f32[C * o + ch] = v[ch];
            if (out_u8) out_u8[C * o + ch] = tm_byte(v[ch]);
        }
    }
}


// ---- refit / instanced assembly (api.cu: drp_refit, drp_build_instanced), run serially with the same per-node functions ------------------
static void hs_scene_bounds(const float* verts, const int32_t* tris, int64_t n, std::vector<uint32_t>& bounds) {
    bounds.assign(12, 0u);
    for (int i = 0; i < 12; ++i) bounds[i] = ((i / 3) % 2 == 0) ? 0xffffffffu : 0u;
    LbvhBuild b;
    memset(&b, 0, sizeof(b));
    b.verts = verts; b.tris = tris; b.n = (int)n;
    for (int64_t i = 0; i < n; ++i) {
        Vec3 lo, hi;
        lbvh_prim_bounds(b, (int)i, lo, hi);
        const float v[6] = {lo.x, lo.y, lo.z, hi.x, hi.y, hi.z};
        for (int k = 0; k < 6; ++k) { const uint32_t o = f2ord(v[k]); bounds[k] = k < 3 ? std::min(bounds[k], o) : std::max(bounds[k], o); }
    }
}
extern "C" int hs_refit_wide(HsBvh* h, const float* verts, const int32_t* tris) {
    if (h->level_begin.size() < 2) return -1;
    hs_scene_bounds(verts, tris, h->n, h->bounds_ord);
    LbvhBuild b;
    memset(&b, 0, sizeof(b));
    b.bounds = h->bounds_ord.data();
    const float abs_pad = lbvh_abs_pad(b);
    std::vector<float4> node_box(2 * (size_t)h->cw_count);
    for (int L = (int)h->level_begin.size() - 2; L >= 0; --L)
        for (int ni = h->level_begin[L]; ni < h->level_begin[L + 1]; ++ni)
            cw_refit_node(h->cw_nodes.data() + CW_NODE_F4 * (size_t)ni, h->cw_tris.data(), h->cw_nodes.data(), h->cw_tris.data(), node_box.data(), ni, 0, 0, 0,
                          verts, tris, abs_pad);
    for (int k = 0; k < 6; ++k) h->bounds[k] = ord2f(h->bounds_ord[k]);
    return 0;
}
extern "C" HsBvh* hs_build_instanced(const float* verts, const int32_t* tris, int64_t n_tris, const int64_t* first, const int32_t* mesh, int64_t n_inst) {
    std::vector<int> rep_of;      // mesh id -> representative instance
    std::vector<HsBvh*> blas;
    auto find = [&](int m) { for (size_t k = 0; k < rep_of.size(); ++k) if (mesh[rep_of[k]] == m) return (int)k; return -1; };
    for (int64_t q = 0; q < n_inst; ++q)
        if (find(mesh[q]) < 0) {
            rep_of.push_back((int)q);
            blas.push_back(hs_build_impl(verts, tris + 3 * first[q], first[q + 1] - first[q], CW_MAX_LEAF, true));
        }
    HsBvh* h = new HsBvh();
    h->n = (int)n_tris;
    const int64_t tlas_cap = 2 * n_inst + 8;
    std::vector<int> node_off(n_inst);
    int64_t total = tlas_cap;
    for (int64_t q = 0; q < n_inst; ++q) { node_off[q] = (int)total; total += blas[find(mesh[q])]->cw_count; }
    h->cw_nodes.assign(CW_NODE_F4 * (size_t)total, make_float4(0, 0, 0, 0));
    h->cw_tris.resize(3 * (size_t)n_tris);
    h->cw_count = (int)total;
    hs_scene_bounds(verts, tris, n_tris, h->bounds_ord);
    LbvhBuild b;
    memset(&b, 0, sizeof(b));
    b.bounds = h->bounds_ord.data();
    const float abs_pad = lbvh_abs_pad(b);
    std::vector<float4> node_box(2 * (size_t)total);
    std::vector<float> lo(3 * (size_t)n_inst), hi(3 * (size_t)n_inst);
    for (int64_t q = 0; q < n_inst; ++q) {
        const HsBvh* t = blas[find(mesh[q])];
        for (int L = (int)t->level_begin.size() - 2; L >= 0; --L)
            for (int ni = t->level_begin[L]; ni < t->level_begin[L + 1]; ++ni)
                cw_refit_node(t->cw_nodes.data() + CW_NODE_F4 * (size_t)ni, t->cw_tris.data(), h->cw_nodes.data(), h->cw_tris.data(), node_box.data(),
                              ni + node_off[q], node_off[q], (int)first[q], (int)first[q], verts, tris, abs_pad);
        const float4 l = node_box[2 * (size_t)node_off[q]], u = node_box[2 * (size_t)node_off[q] + 1];
        lo[3 * q] = l.x; lo[3 * q + 1] = l.y; lo[3 * q + 2] = l.z; hi[3 * q] = u.x; hi[3 * q + 1] = u.y; hi[3 * q + 2] = u.z;
    }
    TlasBuilder tb(lo, hi);
    tb.nodes.resize(CW_NODE_F4, make_float4(0, 0, 0, 0));
    std::vector<int> ids(n_inst);
    std::iota(ids.begin(), ids.end(), 0);
    if (n_inst == 1) tlas_single(tb, lo.data(), hi.data());
    else tb.build(0, ids.data(), (int)n_inst);
    if (tb.allocated > tlas_cap) { delete h; h = nullptr; }
    else {
        for (size_t k = 0; k < (size_t)tb.allocated * CW_NODE_F4; ++k) h->cw_nodes[k] = tb.nodes[k];
        for (auto& c : tb.copies)
            for (int f = 0; f < CW_NODE_F4; ++f) h->cw_nodes[CW_NODE_F4 * (size_t)c.second + f] = h->cw_nodes[CW_NODE_F4 * (size_t)node_off[c.first] + f];
        for (int k = 0; k < 6; ++k) h->bounds[k] = ord2f(h->bounds_ord[k]);
    }
    for (HsBvh* t : blas) delete t;
    return h;
}
