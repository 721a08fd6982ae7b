// traverse.cuh -- closest-hit traversal of the 64-byte two-child node layout produced by lbvh.cuh.
// Replaces NaivePBBVH.query (diffrp/utils/raycaster.py:226-260) / torchoptix.trace_rays.
//
// Closest-hit contract (== the reference's BruteForceRaycaster, raycaster.py:86-97, bit for bit):
//   t = min over triangles of the Moller-Trumbore t (tri_test_mt), ties broken by the smaller primitive id,
//   t == t_far and id == 0 on a miss.
// The result is independent of the hierarchy because the slab test is conservative: boxes are padded at build
// time (lbvh.cuh: pad_box) and the entry/exit distances are widened by 2^-20 relative before comparing.
#pragma once
#include "common.cuh"
#include "lbvh.cuh"

#define DRP_STACK_SIZE 64
#define DRP_T_SHRINK 0.99999905f  // 1 - 2^-20
#define DRP_T_GROW 1.00000095f    // 1 + 2^-20

// traversal statistics, compiled in only for the host simulator (tests/hostsim): nodes fetched / triangles tested
#ifdef DRP_HOSTSIM
struct TravStats { long long nodes, tris; };
static thread_local TravStats g_trav = {0, 0};
#define DRP_COUNT_NODE() (++g_trav.nodes)
#define DRP_COUNT_TRI() (++g_trav.tris)
#else
#define DRP_COUNT_NODE()
#define DRP_COUNT_TRI()
#endif

struct RayHit {
    float t;
    int id;
};

// slab test against one child box; returns entry distance in tmin
DRP_HD bool slab_test(float lox, float loy, float loz, float hix, float hiy, float hiz, Vec3 o, Vec3 idir, float t_best,
                      float& tmin) {
    float t1x = (lox - o.x) * idir.x, t2x = (hix - o.x) * idir.x;
    float t1y = (loy - o.y) * idir.y, t2y = (hiy - o.y) * idir.y;
    float t1z = (loz - o.z) * idir.z, t2z = (hiz - o.z) * idir.z;
    float tn = fmaxf(fmaxf(fminf(t1x, t2x), fminf(t1y, t2y)), fmaxf(fminf(t1z, t2z), 0.0f));
    float tf = fminf(fminf(fmaxf(t1x, t2x), fmaxf(t1y, t2y)), fminf(fmaxf(t1z, t2z), t_best));
    tmin = tn;
    return tn * DRP_T_SHRINK <= tf * DRP_T_GROW;
}

// one packed triangle record against the ray; updates the closest hit with the (t, id) order of the contract
DRP_HD void leaf_update(float4 a, float4 b, float4 c, Vec3 o, Vec3 d, float eps, float& t_best, int& id_best) {
    DRP_COUNT_TRI();
    float t;
#if DRP_TRI_EDGES
    const bool is_hit = tri_test_edges(o, d, v3(a.x, a.y, a.z), v3(a.w, b.x, b.y), v3(b.z, b.w, c.x), eps, t);
#else
    const bool is_hit = tri_test_mt(o, d, v3(a.x, a.y, a.z), v3(a.w, b.x, b.y), v3(b.z, b.w, c.x), eps, t);
#endif
    if (is_hit) {
        int id = f2i(c.y);
        if (t < t_best || (t == t_best && id < id_best)) { t_best = t; id_best = id; }
    }
}
DRP_HD void leaf_intersect(const float4* __restrict__ tris, int first, int count, Vec3 o, Vec3 d, float eps, float& t_best,
                           int& id_best) {
    for (int k = 0; k < count; ++k) {
        const float4* p = tris + 3 * (int64_t)(first + k);
        float4 a = ldg(p), b = ldg(p + 1), c = ldg(p + 2);
        leaf_update(a, b, c, o, d, eps, t_best, id_best);
    }
}

// One ray, private stack.  `nodes` = (n_nodes, 4) float4, `tris` = (n_tris, 3) float4.
DRP_HD RayHit trace_one(const float4* __restrict__ nodes, const float4* __restrict__ tris, Vec3 o, Vec3 d, float t_far,
                        float eps, bool& overflow) {
    Vec3 idir = v3(1.0f / d.x, 1.0f / d.y, 1.0f / d.z);
    float t_best = t_far;
    int id_best = 0x7fffffff;
    int st_ref[DRP_STACK_SIZE];
    float st_t[DRP_STACK_SIZE];
    int sp = 0;
    int node = 0;
    for (;;) {
        if (node >= 0) {
            const float4* p = nodes + 4 * (int64_t)node;
            float4 n0 = ldg(p), n1 = ldg(p + 1), n2 = ldg(p + 2), n3 = ldg(p + 3);
            DRP_COUNT_NODE();
            float tl, tr;
            bool hl = slab_test(n0.x, n0.y, n0.z, n0.w, n1.x, n1.y, o, idir, t_best, tl);
            bool hr = slab_test(n1.z, n1.w, n2.x, n2.y, n2.z, n2.w, o, idir, t_best, tr);
            int cl = f2i(n3.x), cr = f2i(n3.y);
            if (hl & hr) {
                bool swap = tr < tl;
                int nearc = swap ? cr : cl, farc = swap ? cl : cr;
                float tfar_c = swap ? tl : tr;
                if (sp < DRP_STACK_SIZE) { st_ref[sp] = farc; st_t[sp] = tfar_c; ++sp; }
                else overflow = true;  // reported through the handle's device flag; never silent
                node = nearc;
                continue;
            } else if (hl) { node = cl; continue; }
            else if (hr) { node = cr; continue; }
        } else {
            int first, count;
            leaf_decode(node, first, count);
            leaf_intersect(tris, first, count, o, d, eps, t_best, id_best);
        }
        bool done = false;
        for (;;) {  // pop, skipping entries that can no longer contain a closer-or-equal hit
            if (sp == 0) { done = true; break; }
            --sp;
            node = st_ref[sp];
            if (st_t[sp] * DRP_T_SHRINK <= t_best) break;
        }
        if (done) break;
    }
    RayHit h;
    bool hit = t_best < t_far;
    h.t = hit ? t_best : t_far;
    h.id = hit ? id_best : 0;
    return h;
}

// exhaustive closest hit over raw (verts, tris): BruteForceRaycaster semantics (validation aid)
DRP_HD RayHit bruteforce_one(const float* __restrict__ verts, const int32_t* __restrict__ tris, int64_t n_tris, Vec3 o, Vec3 d,
                             float t_far, float eps) {
    float t_best = t_far;
    int id_best = 0;
    for (int64_t k = 0; k < n_tris; ++k) {
        float t;
        if (tri_test_mt(o, d, load_vert(verts, tris[3 * k]), load_vert(verts, tris[3 * k + 1]), load_vert(verts, tris[3 * k + 2]), eps, t)) {
            if (t < t_best) { t_best = t; id_best = (int)k; }
        }
    }
    RayHit h;
    h.t = t_best;
    h.id = t_best < t_far ? id_best : 0;
    return h;
}
