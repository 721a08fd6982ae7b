// lbvh.cuh -- on-GPU LBVH build, phase by phase (replaces NaivePBBVH.build, diffrp/utils/raycaster.py:122-187, and
// torchoptix.build).  Pipeline:
//   1 prim_bounds   per-triangle AABB + scene/centroid bounds (atomic min/max on order-preserving uint keys)
//   2 morton        63-bit Morton code of the AABB centre (21 bits / axis)
//   3 radix sort    (key = Morton, value = triangle id)           -- cub::DeviceRadixSort in api.cu
//   4 karras        binary radix tree over the sorted keys (Karras 2012), one thread per internal node
//   5 refit         bottom-up AABBs with per-node arrival counters; same walk evaluates the SAH and marks
//                   subtrees that are cheaper as one multi-triangle leaf ("SAH-refit leaves")
//   6 emit          traversal nodes: 64 B = both child boxes (padded, conservative) + two child references
//   7 pack_tris     triangles in leaf order as 3 x float4 (A, B, C, original id) for 128-bit loads
// Every phase is a DRP_HD per-element function so tests/hostsim can run the identical logic serially on the CPU.
#pragma once
#include "common.cuh"

#define DRP_MAX_LEAF 4          // max triangles per collapsed leaf (must be < 16, see leaf_ref)
#define DRP_SAH_CI 1.0f         // cost of visiting an internal node (two box tests, one 64 B fetch)
#define DRP_SAH_CT 1.0f         // cost of one triangle test (48 B fetch + Moller-Trumbore)

// child reference: >= 0 internal node index; < 0 leaf: -1 - ((first << 4) | count), count in [0, 15]
DRP_HD int leaf_ref(int first, int count) { return -1 - ((first << 4) | count); }
DRP_HD void leaf_decode(int ref, int& first, int& count) { int c = -1 - ref; first = c >> 4; count = c & 15; }

// order-preserving float <-> uint mapping for atomicMin/atomicMax
DRP_HD uint32_t f2ord(float f) { uint32_t u = f2u(f); return (u & 0x80000000u) ? ~u : (u | 0x80000000u); }
DRP_HD float ord2f(uint32_t u) { return u2f((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u); }

struct LbvhBuild {
    // inputs
    const float* verts;     // (V,3)
    const int32_t* tris;    // (F,3)
    int n;                  // F
    int max_leaf;           // max triangles per collapsed leaf (<= 15); 4 for the binary layout, 3 for the wide layout
    // scene bounds as ordered uints: [0:3] min, [3:6] max of triangle boxes; [6:9] min, [9:12] max of centres
    uint32_t* bounds;
    // per primitive (original order)
    float4* prim_lo;        // xyz = box min
    float4* prim_hi;        // xyz = box max
    // sort
    uint64_t* keys;         // sorted Morton codes (after phase 3)
    uint32_t* vals;         // sorted triangle ids
    // hierarchy; node ids: internal i in [0, n-1), leaf j -> (n-1) + j
    int* left;              // (n-1)
    int* right;             // (n-1)
    int* parent;            // (2n-1)
    int* range_first;       // (n-1)
    int* range_last;        // (n-1)
    float4* box_lo;         // (2n-1) xyz = min, w = SAH cost of the subtree
    float4* box_hi;         // (2n-1) xyz = max, w = surface area
    int* arrive;            // (n-1) arrival counters, zero-initialised
    uint8_t* collapsed;     // (n-1)
    int* count;             // (2n-1) primitives under each node
    // wide-layout collapse by dynamic programming (cwbvh.cuh); null for the binary layout
    float* dp_cost;         // (2n-1, 8): [i] = optimal SAH cost of the subtree as a forest of at most i wide-node children, i = 1..7
    uint8_t* dp_dec;        // (2n-1, 8): [0] = left share of the 8-way split; [1] = 0 leaf / 1 internal; [i>=2] = 0 "same as i-1" or left share
    // outputs
    float4* nodes;          // (max(n-1,1), 4)
    float4* packed;         // (n, 3)
};

DRP_HD Vec3 load_vert(const float* verts, int i) { return v3(verts[3 * (int64_t)i], verts[3 * (int64_t)i + 1], verts[3 * (int64_t)i + 2]); }

// ---- phase 1 --------------------------------------------------------------------------------------------------
DRP_HD void lbvh_prim_bounds(const LbvhBuild& b, int i, Vec3& lo, Vec3& hi) {
    Vec3 A = load_vert(b.verts, b.tris[3 * (int64_t)i]);
    Vec3 B = load_vert(b.verts, b.tris[3 * (int64_t)i + 1]);
    Vec3 C = load_vert(b.verts, b.tris[3 * (int64_t)i + 2]);
    lo = v3(fminf(fminf(A.x, B.x), C.x), fminf(fminf(A.y, B.y), C.y), fminf(fminf(A.z, B.z), C.z));
    hi = v3(fmaxf(fmaxf(A.x, B.x), C.x), fmaxf(fmaxf(A.y, B.y), C.y), fmaxf(fmaxf(A.z, B.z), C.z));
    if (b.prim_lo) {   // (null: scene bounds only -- refit / instanced assembly)
        b.prim_lo[i] = make_float4(lo.x, lo.y, lo.z, 0.0f);
        b.prim_hi[i] = make_float4(hi.x, hi.y, hi.z, 0.0f);
    }
}

// ---- phase 2 --------------------------------------------------------------------------------------------------
DRP_HD uint64_t spread21(uint32_t v) {  // 21 bits -> every third bit of 63
    uint64_t x = v & 0x1fffffu;
    x = (x | x << 32) & 0x1f00000000ffffull;
    x = (x | x << 16) & 0x1f0000ff0000ffull;
    x = (x | x << 8) & 0x100f00f00f00f00full;
    x = (x | x << 4) & 0x10c30c30c30c30c3ull;
    x = (x | x << 2) & 0x1249249249249249ull;
    return x;
}
DRP_HD void lbvh_morton(const LbvhBuild& b, int i) {
    float4 lo = b.prim_lo[i], hi = b.prim_hi[i];
    float c[3] = {0.5f * (lo.x + hi.x), 0.5f * (lo.y + hi.y), 0.5f * (lo.z + hi.z)};
    uint32_t q[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        float mn = ord2f(b.bounds[6 + a]), mx = ord2f(b.bounds[9 + a]);
        float ext = mx - mn;
        float x = ext > 0.0f ? (c[a] - mn) / ext : 0.0f;
        x = fminf(fmaxf(x * 2097152.0f, 0.0f), 2097151.0f);
        q[a] = (uint32_t)x;
    }
    b.keys[i] = spread21(q[0]) | (spread21(q[1]) << 1) | (spread21(q[2]) << 2);
    b.vals[i] = (uint32_t)i;
}

// ---- phase 4: Karras 2012 -----------------------------------------------------------------------------------
DRP_HD int lbvh_delta(const uint64_t* keys, int n, int i, int j) {
    if (j < 0 || j >= n) return -1;
    uint64_t a = keys[i], c = keys[j];
    if (a == c) return 64 + clz32((uint32_t)(i ^ j));
    return clz64(a ^ c);
}
DRP_HD void lbvh_karras(const LbvhBuild& b, int i) {
    const uint64_t* k = b.keys;
    const int n = b.n;
    int d = (lbvh_delta(k, n, i, i + 1) - lbvh_delta(k, n, i, i - 1)) >= 0 ? 1 : -1;
    int dmin = lbvh_delta(k, n, i, i - d);
    int lmax = 2;
    while (lbvh_delta(k, n, i, i + lmax * d) > dmin) lmax <<= 1;
    int l = 0;
    for (int t = lmax >> 1; t >= 1; t >>= 1)
        if (lbvh_delta(k, n, i, i + (l + t) * d) > dmin) l += t;
    int j = i + l * d;
    int dnode = lbvh_delta(k, n, i, j);
    int s = 0, t = l;
    do {
        t = (t + 1) >> 1;
        if (lbvh_delta(k, n, i, i + (s + t) * d) > dnode) s += t;
    } while (t > 1);
    int gamma = i + s * d + (d < 0 ? d : 0);
    int lo = i < j ? i : j, hi = i < j ? j : i;
    int lc = (lo == gamma) ? (n - 1 + gamma) : gamma;
    int rc = (hi == gamma + 1) ? (n - 1 + gamma + 1) : gamma + 1;
    b.left[i] = lc;
    b.right[i] = rc;
    b.parent[lc] = i;
    b.parent[rc] = i;
    b.range_first[i] = lo;
    b.range_last[i] = hi;
    if (i == 0) b.parent[0] = -1;
}

// ---- phase 5: bottom-up refit + SAH leaf decision ------------------------------------------------------------
DRP_HD float box_area(float4 lo, float4 hi) {
    float dx = hi.x - lo.x, dy = hi.y - lo.y, dz = hi.z - lo.z;
    return 2.0f * (dx * dy + dy * dz + dz * dx);
}
// Optimal collapse into 8-wide nodes (Ylitie, Karras, Laine 2017, section 4.2), evaluated bottom-up inside the refit walk.
// C(n,i): minimum SAH cost of subtree n when it may occupy at most i child slots of a wide node.
//   C(n,1) = min(leaf: A P Cprim if P <= max_leaf, internal: A Cnode + D(n,8));  C(n,i) = min(D(n,i), C(n,i-1))
//   D(n,j) = min over 0<k<j of C(left,k) + C(right,j-k)
// Returns whether n taken as ONE slot is a leaf.
#ifndef DRP_DP_CNODE
#define DRP_DP_CNODE 1.0f
#endif
#ifndef DRP_DP_CPRIM
#define DRP_DP_CPRIM 1.0f   // B200 sweep on config 3 (extend ms per step): 0.15 -> 5.37, 0.3 -> 5.31, 0.6 -> 5.25, 1.0 -> 5.24, 1.5 -> 5.25
#endif
DRP_HD bool lbvh_dp_node(const LbvhBuild& b, int p, int lc, int rc, float area, int count) {
    float cl[8], cr[8];
    for (int i = 1; i < 8; ++i) {
#ifdef __CUDA_ARCH__
        cl[i] = __ldcg(&b.dp_cost[8 * (int64_t)lc + i]);
        cr[i] = __ldcg(&b.dp_cost[8 * (int64_t)rc + i]);
#else
        cl[i] = b.dp_cost[8 * (int64_t)lc + i];
        cr[i] = b.dp_cost[8 * (int64_t)rc + i];
#endif
    }
    float dist[9];
    uint8_t argk[9];
    for (int j = 2; j <= 8; ++j) {
        float best = 3e38f;
        int bk = 1;
        for (int k = 1; k < j; ++k) {
            if (k > 7 || j - k > 7) continue;
            float c = cl[k] + cr[j - k];
            if (c < best) { best = c; bk = k; }
        }
        dist[j] = best;
        argk[j] = (uint8_t)bk;
    }
    float* cost = b.dp_cost + 8 * (int64_t)p;
    uint8_t* dec = b.dp_dec + 8 * (int64_t)p;
    const float c_leaf = count <= b.max_leaf ? area * (float)count * DRP_DP_CPRIM : 3e38f;
    const float c_int = dist[8] + area * DRP_DP_CNODE;
    const bool leaf = (c_leaf <= c_int) && (p != 0);
    cost[1] = leaf ? c_leaf : c_int;
    dec[0] = argk[8];
    dec[1] = leaf ? 0 : 1;
    for (int i = 2; i < 8; ++i) {
        if (dist[i] < cost[i - 1]) { cost[i] = dist[i]; dec[i] = argk[i]; }
        else { cost[i] = cost[i - 1]; dec[i] = 0; }
    }
    return leaf;
}

// `atomic_inc(ptr)` returns the previous value; `fence()` orders the box stores before the counter update.
template <typename AtomicInc, typename Fence>
DRP_HD void lbvh_refit(const LbvhBuild& b, int j, AtomicInc atomic_inc, Fence fence) {
    const int n = b.n;
    uint32_t prim = b.vals[j];
    float4 lo = b.prim_lo[prim], hi = b.prim_hi[prim];
    float area = box_area(lo, hi);
    lo.w = DRP_SAH_CT * area;  // cost
    hi.w = area;
    int node = n - 1 + j;
    b.box_lo[node] = lo;
    b.box_hi[node] = hi;
    if (b.count) b.count[node] = 1;
    if (b.dp_cost) {
        for (int i = 1; i < 8; ++i) { b.dp_cost[8 * (int64_t)node + i] = DRP_DP_CPRIM * area; b.dp_dec[8 * (int64_t)node + i] = 0; }
        b.dp_dec[8 * (int64_t)node] = 0;
    }
    int p = b.parent[node];
    while (p >= 0) {
        fence();
        if (atomic_inc(&b.arrive[p]) == 0) return;  // first child to arrive: the sibling's thread continues
        fence();
        int lc = b.left[p], rc = b.right[p];
#ifdef __CUDA_ARCH__
        // children were written by other threads: bypass L1
        float4 llo = __ldcg(&b.box_lo[lc]), lhi = __ldcg(&b.box_hi[lc]), rlo = __ldcg(&b.box_lo[rc]), rhi = __ldcg(&b.box_hi[rc]);
#else
        float4 llo = b.box_lo[lc], lhi = b.box_hi[lc], rlo = b.box_lo[rc], rhi = b.box_hi[rc];
#endif
        float4 plo = make_float4(fminf(llo.x, rlo.x), fminf(llo.y, rlo.y), fminf(llo.z, rlo.z), 0.0f);
        float4 phi = make_float4(fmaxf(lhi.x, rhi.x), fmaxf(lhi.y, rhi.y), fmaxf(lhi.z, rhi.z), 0.0f);
        float a = box_area(plo, phi);
        int count = b.range_last[p] - b.range_first[p] + 1;
        if (b.count) b.count[p] = count;
        float c_split = DRP_SAH_CI * a + llo.w + rlo.w;
        float c_leaf = DRP_SAH_CT * a * (float)count;
        bool make_leaf = (count <= b.max_leaf) && (c_leaf <= c_split) && (p != 0);
        if (b.dp_cost) make_leaf = lbvh_dp_node(b, p, lc, rc, a, count);
        b.collapsed[p] = make_leaf ? 1 : 0;
        plo.w = make_leaf ? c_leaf : c_split;
        phi.w = a;
        b.box_lo[p] = plo;
        b.box_hi[p] = phi;
        p = b.parent[p];
    }
}

// ---- phase 6: emit traversal nodes ----------------------------------------------------------------------------
// A box is padded so that the slab test can never reject a triangle the fp32 triangle test accepts.
DRP_HD void pad_box(float4& lo, float4& hi, float abs_pad) {
    const float rel = 9.5367431640625e-07f;  // 2^-20
    float px = rel * fmaxf(fabsf(lo.x), fabsf(hi.x)) + abs_pad;
    float py = rel * fmaxf(fabsf(lo.y), fabsf(hi.y)) + abs_pad;
    float pz = rel * fmaxf(fabsf(lo.z), fabsf(hi.z)) + abs_pad;
    lo.x -= px; lo.y -= py; lo.z -= pz;
    hi.x += px; hi.y += py; hi.z += pz;
}
DRP_HD float lbvh_abs_pad(const LbvhBuild& b) {
    float dx = ord2f(b.bounds[3]) - ord2f(b.bounds[0]), dy = ord2f(b.bounds[4]) - ord2f(b.bounds[1]), dz = ord2f(b.bounds[5]) - ord2f(b.bounds[2]);
    return 2.384185791015625e-07f * fmaxf(fmaxf(dx, dy), fmaxf(dz, 1e-30f));  // 2^-22 * extent
}
DRP_HD int lbvh_child_ref(const LbvhBuild& b, int c) {
    const int n = b.n;
    if (c >= n - 1) return leaf_ref(c - (n - 1), 1);
    if (b.collapsed[c]) return leaf_ref(b.range_first[c], b.range_last[c] - b.range_first[c] + 1);
    return c;
}
DRP_HD void lbvh_emit(const LbvhBuild& b, int i) {
    float pad = lbvh_abs_pad(b);
    int lc = b.left[i], rc = b.right[i];
    float4 llo = b.box_lo[lc], lhi = b.box_hi[lc], rlo = b.box_lo[rc], rhi = b.box_hi[rc];
    pad_box(llo, lhi, pad);
    pad_box(rlo, rhi, pad);
    float4* o = b.nodes + 4 * (int64_t)i;
    o[0] = make_float4(llo.x, llo.y, llo.z, lhi.x);
    o[1] = make_float4(lhi.y, lhi.z, rlo.x, rlo.y);
    o[2] = make_float4(rlo.z, rhi.x, rhi.y, rhi.z);
    o[3] = make_float4(i2f(lbvh_child_ref(b, lc)), i2f(lbvh_child_ref(b, rc)), 0.0f, 0.0f);
}
// degenerate scenes: n == 1 (single leaf) and n == 0 (nothing): one root whose children are leaves with count 0/1
DRP_HD void lbvh_emit_tiny(const LbvhBuild& b) {
    float4 lo = make_float4(0, 0, 0, 0), hi = make_float4(0, 0, 0, 0);
    if (b.n == 1) { lo = b.prim_lo[0]; hi = b.prim_hi[0]; pad_box(lo, hi, lbvh_abs_pad(b)); }
    float4* o = b.nodes;
    o[0] = make_float4(lo.x, lo.y, lo.z, hi.x);
    o[1] = make_float4(hi.y, hi.z, 0.0f, 0.0f);
    o[2] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    o[3] = make_float4(i2f(leaf_ref(0, b.n == 1 ? 1 : 0)), i2f(leaf_ref(0, 0)), 0.0f, 0.0f);
}

// ---- phase 7 --------------------------------------------------------------------------------------------------
DRP_HD void lbvh_pack_tri(const LbvhBuild& b, int j) {
    int prim = (int)b.vals[j];
    Vec3 A = load_vert(b.verts, b.tris[3 * (int64_t)prim]);
    Vec3 B = load_vert(b.verts, b.tris[3 * (int64_t)prim + 1]);
    Vec3 C = load_vert(b.verts, b.tris[3 * (int64_t)prim + 2]);
    pack_triangle(b.packed + 3 * (int64_t)j, A, B, C, prim);
}
