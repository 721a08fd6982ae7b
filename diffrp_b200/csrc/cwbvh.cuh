// cwbvh.cuh -- compressed 8-wide BVH ("CWBVH", after Ylitie, Karras, Laine 2017) built on the GPU by collapsing the
// LBVH of lbvh.cuh, and its closest-hit traversal.
//
// Why (ncu evidence, profiles/README.md): the 64-byte two-child layout is issue/L1-bound at 6-10 active lanes per
// instruction and ~45 node visits per ray.  An 8-wide node with 8-bit quantised child boxes is 80 bytes for 8 children
// (10 B/child instead of 32 B/child), is tested with one FMA per slab plane, and the per-thread traversal state
// (node group + triangle group bitmasks) postpones triangle tests so that lanes of a warp stay in the same phase.
//
// Node = 5 x float4 (80 B):
//   n0 = (p.x, p.y, p.z, bits{ex, ey, ez, imask})          origin of the local grid, per-axis exponent bytes (IEEE-biased, see
//                                                           cw_pack_exponents), mask of the slots that hold internal children
//   n1 = (child_base, tri_base, V, 0)                       first child node / first triangle; V = valid-triangle mask, bit 3*slot+k set
//                                                           iff leaf slot `slot` holds a k-th triangle
//   n2 = (qlo_x[0..3], qlo_x[4..7], qlo_y[0..3], qlo_y[4..7])
//   n3 = (qlo_z[0..3], qlo_z[4..7], qhi_x[0..3], qhi_x[4..7])
//   n4 = (qhi_y[0..3], qhi_y[4..7], qhi_z[0..3], qhi_z[4..7])
// Internal children of a node are consecutive nodes from child_base in slot order; the triangles of its leaf slots are consecutive
// 48-byte records from tri_base in slot order (common.cuh: pack_triangle).  Child boxes are
// lo = p + qlo * 2^e, hi = p + qhi * 2^e (rounded outwards), and contain the padded boxes of lbvh.cuh, so the traversal is
// conservative and the closest hit equals the exhaustive one (traverse.cuh contract).
#pragma once
#include "common.cuh"
#include "lbvh.cuh"
#include "traverse.cuh"

// Node format ("format 2" in profiles/README.md; the per-slot meta-byte format 1, the 96-byte / 256-bit-load variant and the half-skip
// packing all lost their A/Bs on B200 and were removed): n1 = (child_base, tri_base, V, 0) with V = the valid-triangle mask of the
// node.  The box test leaves one hit bit per slot (sign bit of cmax - cmin funnel-shifted into a byte: 1 ALU-pipe op per slot), and the
// per-node expansion of that byte -- internal children permuted into traversal priority (slot ^ octant), leaf children spread to 3 bits
// per slot and masked with V -- is two table lookups (shared memory in k_extend_cw, arithmetic elsewhere).  Triangles of a node are
// stored compactly in slot order, so triangle bit j is triangle tri_base + popc(V & ((1 << j) - 1)).
#define DRP_CW_V2 1   // (tests/hostsim keys on it)
#define CW_NODE_F4 5
#define CW_MAX_LEAF 3
#define CW_STACK 48       // per-thread entries of the fast path; deeper rays are re-traced by k_extend_fixup with CW_DEEP_STACK entries
#define CW_DEEP_STACK 256
#define CW_SLACK 4.76837158203125e-07f  // 2^-21: relative widening of every slab distance
#define CW_GRID_SLACK 0.0079f           // + this many grid steps: 256 * 2^-21 (FMA cancellation) + 2^-8 (bias folding, x2 margin)

struct CwBuild {
    LbvhBuild b;       // the binary hierarchy (left/right, boxes with area in box_hi.w, collapsed flags, ranges, vals)
    float4* cw_nodes;  // (capacity, CW_NODE_F4)
    float4* cw_tris;   // (n, 3) triangles in node order: A, B, C, original id
    int* work;         // (capacity) binary node collapsed into each wide node
    int* counters;     // [0] wide nodes allocated, [1] triangles placed, [8 + L] first node of level L
    int capacity;
};

DRP_HD bool cw_is_leaf_child(const LbvhBuild& b, int c) { return c >= b.n - 1 || b.collapsed[c]; }
DRP_HD int cw_leaf_count(const LbvhBuild& b, int c) { return b.count[c]; }
// sorted positions of the (<= CW_MAX_LEAF) primitives under binary node c, by walking the subtree (the PLOC tree does not
// keep primitives of a subtree contiguous in Morton order)
DRP_HD int cw_leaf_gather(const LbvhBuild& b, int c, int* out) {
    int stack[8], sp = 0, k = 0;
    stack[sp++] = c;
    while (sp > 0) {
        int nd = stack[--sp];
        if (nd >= b.n - 1) { if (k < CW_MAX_LEAF) out[k] = nd - (b.n - 1); ++k; }
        else { stack[sp++] = b.right[nd]; stack[sp++] = b.left[nd]; }
    }
    return k;
}

// Per-axis grid exponents (|e| <= 100) are stored as IEEE-biased bytes (e + 127) in n0.w bits 0..23, so that the scale 2^e of an
// axis is one shift and one mask away.
DRP_HD uint32_t cw_pack_exponents(int ex, int ey, int ez) {
    return ((uint32_t)(ex + 127) & 0xffu) | (((uint32_t)(ey + 127) & 0xffu) << 8) | (((uint32_t)(ez + 127) & 0xffu) << 16);
}
DRP_HD float cw_axis_scale(uint32_t ebits, int axis) {  // 2^e of axis 0..2
    return u2f(axis == 0 ? (ebits << 23) & 0x7f800000u : (axis == 1 ? (ebits << 15) & 0x7f800000u : (ebits << 7) & 0x7f800000u));
}
DRP_HD uint32_t cw_pack4(const uint32_t* v) { return v[0] | (v[1] << 8) | (v[2] << 16) | (v[3] << 24); }

// Encode one wide node from the (padded, conservative) boxes of its occupied slots: node box = their union, per-axis power-of-two grid
// over it, child boxes quantised outwards to 8 bits.  Slot s is occupied iff it is internal (imask bit s) or holds triangles (vmask bits
// 3s..3s+2).  Used by the collapse, by the refit of an existing topology (cw_refit_node) and by the host-built instance level.
DRP_HD void cw_encode_node(float4* o, const float slo[8][3], const float shi[8][3], uint32_t imask, uint32_t vmask, int child_base, int tri_base,
                           float nlo[3], float nhi[3]) {
    for (int a = 0; a < 3; ++a) { nlo[a] = 3e38f; nhi[a] = -3e38f; }
    uint32_t used = imask;
    for (int s = 0; s < 8; ++s) {
        if ((vmask >> (3 * s)) & 7u) used |= 1u << s;
        if (used & (1u << s))
            for (int a = 0; a < 3; ++a) { nlo[a] = fminf(nlo[a], slo[s][a]); nhi[a] = fmaxf(nhi[a], shi[s][a]); }
    }
    if (used == 0) for (int a = 0; a < 3; ++a) { nlo[a] = 0.0f; nhi[a] = 0.0f; }
    // per-axis exponent of the 8-bit grid
    int e[3];
    float inv_scale[3];
    for (int a = 0; a < 3; ++a) {
        float ext = nhi[a] - nlo[a];
        int ea = (int)ceilf(log2f(fmaxf(ext, 1e-30f) / 255.0f));
        ea = ea < -100 ? -100 : (ea > 100 ? 100 : ea);
        while (ext > 255.0f * i2f((ea + 127) << 23) && ea < 100) ++ea;
        e[a] = ea;
        inv_scale[a] = i2f((127 - ea) << 23);
    }
    uint32_t q[6][8];
    for (int s = 0; s < 8; ++s) {
        if (!(used & (1u << s))) {
            for (int a = 0; a < 3; ++a) { q[a][s] = 255; q[3 + a][s] = 0; }   // empty slot: lo 255 > hi 0 never hits
            continue;
        }
        for (int a = 0; a < 3; ++a) {
            float ql = floorf((slo[s][a] - nlo[a]) * inv_scale[a]);
            float qh = ceilf((shi[s][a] - nlo[a]) * inv_scale[a]);
            q[a][s] = (uint32_t)fminf(fmaxf(ql, 0.0f), 255.0f);
            q[3 + a][s] = (uint32_t)fminf(fmaxf(qh, 0.0f), 255.0f);
        }
    }
    uint32_t ebits = cw_pack_exponents(e[0], e[1], e[2]) | (imask << 24);
    o[0] = make_float4(nlo[0], nlo[1], nlo[2], u2f(ebits));
    o[1] = make_float4(i2f(child_base), i2f(tri_base), u2f(vmask), u2f(0u));
    o[2] = make_float4(u2f(cw_pack4(q[0])), u2f(cw_pack4(q[0] + 4)), u2f(cw_pack4(q[1])), u2f(cw_pack4(q[1] + 4)));
    o[3] = make_float4(u2f(cw_pack4(q[2])), u2f(cw_pack4(q[2] + 4)), u2f(cw_pack4(q[3])), u2f(cw_pack4(q[3] + 4)));
    o[4] = make_float4(u2f(cw_pack4(q[4])), u2f(cw_pack4(q[4] + 4)), u2f(cw_pack4(q[5])), u2f(cw_pack4(q[5] + 4)));
}

// Slot s stands for the octant direction ((s&4)?+:-, (s&2)?+:-, (s&1)?+:-); greedily give each slot the child that lies furthest in
// that direction (centre relative to the centre of the node box), so that (slot ^ octant) orders children front to back for any ray.
// lo / hi: k <= 8 child boxes; nlo / nhi: their union.  slot_child[s] = child index or -1.
DRP_HD void cw_assign_slots(const float lo[8][3], const float hi[8][3], int k, const float nlo[3], const float nhi[3], int slot_child[8]) {
    float cost[8][8];
    for (int j = 0; j < k; ++j) {
        float cx = 0.5f * (lo[j][0] + hi[j][0]) - 0.5f * (nlo[0] + nhi[0]);
        float cy = 0.5f * (lo[j][1] + hi[j][1]) - 0.5f * (nlo[1] + nhi[1]);
        float cz = 0.5f * (lo[j][2] + hi[j][2]) - 0.5f * (nlo[2] + nhi[2]);
        for (int s = 0; s < 8; ++s) cost[j][s] = ((s & 4) ? cx : -cx) + ((s & 2) ? cy : -cy) + ((s & 1) ? cz : -cz);
    }
    bool child_done[8] = {false, false, false, false, false, false, false, false};
    for (int s = 0; s < 8; ++s) slot_child[s] = -1;
    for (int it = 0; it < k; ++it) {
        float bc = -3e38f;
        int bj = -1, bs = -1;
        for (int j = 0; j < k; ++j) {
            if (child_done[j]) continue;
            for (int s = 0; s < 8; ++s)
                if (slot_child[s] < 0 && cost[j][s] > bc) { bc = cost[j][s]; bj = j; bs = s; }
        }
        slot_child[bs] = bj;
        child_done[bj] = true;
    }
}

// Collapse the binary subtree rooted at work[ni] into wide node ni.  `atomic_add(ptr, v)` returns the old value.
template <typename AtomicAdd>
DRP_HD void cw_collapse_node(const CwBuild& cw, int ni, AtomicAdd atomic_add) {
    const LbvhBuild& b = cw.b;
    const int root = cw.work[ni];
    int child[8];
    int k = 0;
    if (b.dp_dec) {
        // children chosen by the optimal collapse (lbvh_dp_node): split the 8 slots between the two binary children as
        // decided bottom-up, then resolve each share recursively
        int st_node[16], st_i[16], sp = 0;
        const int kl = b.dp_dec[8 * (int64_t)root];
        st_node[sp] = b.right[root]; st_i[sp] = 8 - kl; ++sp;
        st_node[sp] = b.left[root]; st_i[sp] = kl; ++sp;
        while (sp > 0) {
            --sp;
            int nd = st_node[sp], i = st_i[sp];
            if (nd >= b.n - 1) { child[k++] = nd; continue; }  // a single primitive
            const uint8_t* dec = b.dp_dec + 8 * (int64_t)nd;
            while (i > 1 && dec[i] == 0) --i;                  // "same as i-1"
            if (i == 1) { child[k++] = nd; continue; }         // one slot: leaf or internal wide node (collapsed flag)
            const int kk = dec[i];
            st_node[sp] = b.right[nd]; st_i[sp] = i - kk; ++sp;
            st_node[sp] = b.left[nd]; st_i[sp] = kk; ++sp;
        }
    } else {
        k = 2;
        child[0] = b.left[root];
        child[1] = b.right[root];
        while (k < 8) {  // greedy: open the child with the largest surface area until 8 children or nothing left to open
            int best = -1;
            float best_area = -1.0f;
            for (int j = 0; j < k; ++j)
                if (!cw_is_leaf_child(b, child[j])) {
                    float a = b.box_hi[child[j]].w;
                    if (a > best_area) { best_area = a; best = j; }
                }
            if (best < 0) break;
            int c = child[best];
            child[best] = b.left[c];
            child[k++] = b.right[c];
        }
    }
    // padded child boxes and the node box
    float lo[8][3], hi[8][3], nlo[3] = {3e38f, 3e38f, 3e38f}, nhi[3] = {-3e38f, -3e38f, -3e38f};
    const float abs_pad = lbvh_abs_pad(b);
    for (int j = 0; j < k; ++j) {
        float4 l = b.box_lo[child[j]], h = b.box_hi[child[j]];
        pad_box(l, h, abs_pad);
        lo[j][0] = l.x; lo[j][1] = l.y; lo[j][2] = l.z;
        hi[j][0] = h.x; hi[j][1] = h.y; hi[j][2] = h.z;
        for (int a = 0; a < 3; ++a) { nlo[a] = fminf(nlo[a], lo[j][a]); nhi[a] = fmaxf(nhi[a], hi[j][a]); }
    }
    int slot_child[8];
    cw_assign_slots(lo, hi, k, nlo, nhi, slot_child);
    // which slots hold internal children / how many triangles the leaf slots hold
    uint32_t imask = 0, vmask = 0;
    int n_inner = 0, n_tris = 0;
    float slo[8][3], shi[8][3];
    for (int s = 0; s < 8; ++s) {
        int j = slot_child[s];
        if (j < 0) continue;
        for (int a = 0; a < 3; ++a) { slo[s][a] = lo[j][a]; shi[s][a] = hi[j][a]; }
        int c = child[j];
        if (cw_is_leaf_child(b, c)) {
            int cnt = cw_leaf_count(b, c);
            vmask |= ((1u << cnt) - 1u) << (3 * s);
            n_tris += cnt;
        } else {
            imask |= 1u << s;
            ++n_inner;
        }
    }
    const int child_base = n_inner ? atomic_add(&cw.counters[0], n_inner) : 0;
    const int tri_base = n_tris ? atomic_add(&cw.counters[1], n_tris) : 0;
    int rank = 0, toff = 0;
    for (int s = 0; s < 8; ++s) {
        int j = slot_child[s];
        if (j < 0) continue;
        int c = child[j];
        if (imask & (1u << s)) {
            if (child_base + rank < cw.capacity) cw.work[child_base + rank] = c;
            ++rank;
        } else {
            int pos[CW_MAX_LEAF];
            const int cnt = cw_leaf_gather(b, c, pos);
            for (int t = 0; t < cnt; ++t) {
                int prim = (int)b.vals[pos[t]];
                Vec3 A = load_vert(b.verts, b.tris[3 * (int64_t)prim]);
                Vec3 B = load_vert(b.verts, b.tris[3 * (int64_t)prim + 1]);
                Vec3 C = load_vert(b.verts, b.tris[3 * (int64_t)prim + 2]);
                pack_triangle(cw.cw_tris + 3 * (int64_t)(tri_base + toff + t), A, B, C, prim);
            }
            toff += cnt;
        }
    }
    float box_lo[3], box_hi[3];
    cw_encode_node(cw.cw_nodes + CW_NODE_F4 * (int64_t)ni, slo, shi, imask, vmask, child_base, tri_base, box_lo, box_hi);
}

// n < 2: one node whose slot 0 is a leaf with n triangles
DRP_HD void cw_emit_tiny(const CwBuild& cw) {
    const LbvhBuild& b = cw.b;
    float4* o = cw.cw_nodes;
    float4 lo = make_float4(0, 0, 0, 0), hi = make_float4(0, 0, 0, 0);
    uint32_t meta0 = 0;
    if (b.n == 1) {
        lo = b.prim_lo[0]; hi = b.prim_hi[0];
        pad_box(lo, hi, lbvh_abs_pad(b));
        meta0 = 1u << 5;
        Vec3 A = load_vert(b.verts, b.tris[0]), B = load_vert(b.verts, b.tris[1]), C = load_vert(b.verts, b.tris[2]);
        pack_triangle(cw.cw_tris, A, B, C, 0);
    }
    int e[3];
    for (int a = 0; a < 3; ++a) {
        float ext = a == 0 ? hi.x - lo.x : (a == 1 ? hi.y - lo.y : hi.z - lo.z);
        int ea = (int)ceilf(log2f(fmaxf(ext, 1e-30f) / 255.0f));
        ea = ea < -100 ? -100 : (ea > 100 ? 100 : ea);
        while (ext > 255.0f * i2f((ea + 127) << 23) && ea < 100) ++ea;
        e[a] = ea;
    }
    uint32_t ebits = cw_pack_exponents(e[0], e[1], e[2]);
    o[0] = make_float4(lo.x, lo.y, lo.z, u2f(ebits));
    o[1] = make_float4(i2f(0), i2f(0), u2f(meta0 ? 1u : 0u), u2f(0u));
    const uint32_t lo_q = 0xffffff00u, hi_q = 0x000000ffu;  // slot 0 spans the grid, slots 1..7 are empty (lo 255 > hi 0)
    o[2] = make_float4(u2f(lo_q), u2f(0xffffffffu), u2f(lo_q), u2f(0xffffffffu));
    o[3] = make_float4(u2f(lo_q), u2f(0xffffffffu), u2f(hi_q), u2f(0u));
    o[4] = make_float4(u2f(hi_q), u2f(0u), u2f(hi_q), u2f(0u));
}

// ---- traversal ----------------------------------------------------------------------------------------------------
// Byte i of x as the float 32768 + byte, without an I2F (quarter-rate XU pipe; 48 per node otherwise): one PRMT drops the
// byte into mantissa bits 8..15 of 2^15.  The 32768 is folded into the FMA's addend (cw_node_hits), whose rounding then
// costs at most 2^-9 of a grid step -- covered by the conservative slack.
#define CW_BIAS 32768.0f
// `bias` (= 0x47000000) is passed in as a runtime value so that PRMT takes the *selector* as its immediate operand and the
// bias from a register / constant bank; with both compile-time constants ptxas spends an extra move per PRMT on the selector.
DRP_HD float cw_byte_biased(uint32_t x, int i, uint32_t bias) {
#ifdef __CUDA_ARCH__
    return __uint_as_float(__byte_perm(x, bias, 0x7504u + ((uint32_t)i << 4)));
#else
    (void)bias;
    return u2f(0x47000000u | (((x >> (8 * i)) & 0xffu) << 8));
#endif
}
DRP_HD int cw_bfind(uint32_t x) { return 31 - clz32(x); }
DRP_HD int cw_popc(uint32_t x) {
#ifdef __CUDA_ARCH__
    return __popc(x);
#else
    return __builtin_popcount(x);
#endif
}

struct CwRay {
    Vec3 o, d, idir;
    uint32_t octinv4;
    uint32_t bias;  // 0x47000000, see cw_byte_biased
};
DRP_HD float cw_safe_rcp(float d) { return 1.0f / (fabsf(d) > 1e-20f ? d : copysignf(1e-20f, d)); }
DRP_HD CwRay cw_make_ray(Vec3 o, Vec3 d, uint32_t bias = 0x47000000u) {
    CwRay r;
    r.bias = bias;
    r.o = o; r.d = d;
    r.idir = v3(cw_safe_rcp(d.x), cw_safe_rcp(d.y), cw_safe_rcp(d.z));
    uint32_t oct = (r.idir.x < 0.0f ? 4u : 0u) | (r.idir.y < 0.0f ? 2u : 0u) | (r.idir.z < 0.0f ? 1u : 0u);
    r.octinv4 = (7u - oct) * 0x01010101u;
    return r;
}

// slot s of node p: -1 = internal child, 0 = empty, n > 0 = leaf with n triangles (statistics / tests)
DRP_HD int cw_slot_kind(const float4* p, int s) {
    if ((f2u(p[0].w) >> 24) & (1u << s)) return -1;
    return cw_popc((f2u(p[1].z) >> (3 * s)) & 7u);
}
// 8 bits (one per slot) -> the same bits at position (slot ^ octinv): internal children in traversal priority
DRP_HD uint32_t cw_perm8(uint32_t x, uint32_t octinv) {
    if (octinv & 4u) x = ((x & 0x0fu) << 4) | ((x >> 4) & 0x0fu);
    if (octinv & 2u) x = ((x & 0x33u) << 2) | ((x >> 2) & 0x33u);
    if (octinv & 1u) x = ((x & 0x55u) << 1) | ((x >> 1) & 0x55u);
    return x;
}
// 8 bits (one per slot) -> bits 3*slot .. 3*slot+2 all set: every triangle position of the hit leaf slots
DRP_HD uint32_t cw_spread3x7(uint32_t x) {
    x = (x | (x << 8)) & 0x0000f00fu;
    x = (x | (x << 4)) & 0x000c30c3u;
    x = (x | (x << 2)) & 0x00249249u;
    return x * 7u;
}
// index of the triangle behind bit j of a node's triangle mask (triangles are stored compactly in slot order)
DRP_HD int cw_tri_index(uint32_t tri_base, uint32_t vmask, int j) { return (int)(tri_base + (uint32_t)cw_popc(vmask & ((1u << j) - 1u))); }

// intersect the 8 quantised child boxes of one node; returns one hit bit per slot (bit s = slot s)
DRP_HD uint32_t cw_node_hits(const CwRay& r, float4 n0, float4 n2, float4 n3, float4 n4, float t_cull) {
    const uint32_t ebits = f2u(n0.w);
    const float ax = cw_axis_scale(ebits, 0) * r.idir.x, ay = cw_axis_scale(ebits, 1) * r.idir.y, az = cw_axis_scale(ebits, 2) * r.idir.z;
    const float ox = (n0.x - r.o.x) * r.idir.x, oy = (n0.y - r.o.y) * r.idir.y, oz = (n0.z - r.o.z) * r.idir.z;
    // conservative widening: the FMA form cancels, its error is relative to |origin term| + |grid term|
    const float sx = CW_SLACK * fabsf(ox) + CW_GRID_SLACK * fabsf(ax), sy = CW_SLACK * fabsf(oy) + CW_GRID_SLACK * fabsf(ay),
                sz = CW_SLACK * fabsf(oz) + CW_GRID_SLACK * fabsf(az);
    // addends with the byte bias folded in: t = (32768 + q) * a + (o -+ slack - 32768 a)
    const float bx = ox - CW_BIAS * ax, by = oy - CW_BIAS * ay, bz = oz - CW_BIAS * az;
    const float oxl = bx - sx, oxh = bx + sx, oyl = by - sy, oyh = by + sy, ozl = bz - sz, ozh = bz + sz;
    uint32_t miss = 0;  // slot 7 is tested first and ends up in bit 7: every test shifts the sign of (cmax - cmin) in from the right
#pragma unroll
    for (int half = 1; half >= 0; --half) {
        const uint32_t qlx = f2u(half ? n2.y : n2.x), qly = f2u(half ? n2.w : n2.z), qlz = f2u(half ? n3.y : n3.x);
        const uint32_t qhx = f2u(half ? n3.w : n3.z), qhy = f2u(half ? n4.y : n4.x), qhz = f2u(half ? n4.w : n4.z);
        const uint32_t nx = r.idir.x < 0.0f ? qhx : qlx, fx = r.idir.x < 0.0f ? qlx : qhx;
        const uint32_t ny = r.idir.y < 0.0f ? qhy : qly, fy = r.idir.y < 0.0f ? qly : qhy;
        const uint32_t nz = r.idir.z < 0.0f ? qhz : qlz, fz = r.idir.z < 0.0f ? qlz : qhz;
#pragma unroll
        for (int i = 3; i >= 0; --i) {
            float tnx = cw_byte_biased(nx, i, r.bias) * ax + oxl, tny = cw_byte_biased(ny, i, r.bias) * ay + oyl, tnz = cw_byte_biased(nz, i, r.bias) * az + ozl;
            float tfx = cw_byte_biased(fx, i, r.bias) * ax + oxh, tfy = cw_byte_biased(fy, i, r.bias) * ay + oyh, tfz = cw_byte_biased(fz, i, r.bias) * az + ozh;
            float cmin = fmaxf(fmaxf(tnx, tny), fmaxf(tnz, 0.0f));
            float cmax = fminf(fminf(tfx, tfy), fminf(tfz, t_cull));
            // hit iff cmin <= cmax.  cmin >= 0 and cmax <= t_cull are never NaN (FMNMX drops NaN operands), so the sign of the
            // difference decides; an empty slot (lo 255 > hi 0) always misses.  (-0) - (+0) counts as a miss: the widened box then
            // ends at the ray origin and cannot hold a hit with t > 0.
#ifdef __CUDA_ARCH__
            miss = __funnelshift_l(__float_as_uint(__fsub_rn(cmax, cmin)), miss, 1);   // FADD (FMA pipe) + SHF (ALU pipe)
#else
            miss = (miss << 1) | (f2u(cmax - cmin) >> 31);
#endif
        }
    }
    return ~miss & 0xffu;
}
// One ray over the wide layout with a caller-supplied stack of `cap` (node group | postponed triangle group) entries.  The closest hit
// does not depend on the visiting order; `overflow` is set when an entry had to be dropped (the result is then not trustworthy).
struct CwPairStack {     // two private arrays (registers / local memory)
    uint32_t *x, *y;
    DRP_HD void put(int i, uint32_t a, uint32_t b) const { x[i] = a; y[i] = b; }
    DRP_HD void get(int i, uint32_t& a, uint32_t& b) const { a = x[i]; b = y[i]; }
};
struct CwStridedStack {  // entry i of this thread at base[i * stride]: [entry][thread] in global memory, coalesced across a warp
    uint2* base;
    int64_t stride;
    DRP_HD void put(int i, uint32_t a, uint32_t b) const { uint2 v; v.x = a; v.y = b; base[i * stride] = v; }
    DRP_HD void get(int i, uint32_t& a, uint32_t& b) const { const uint2 v = base[i * stride]; a = v.x; b = v.y; }
};
template <typename Stack>
DRP_HD RayHit cw_trace_one_stack(const float4* __restrict__ nodes, const float4* __restrict__ tris, Vec3 o, Vec3 d, float t_far, float eps,
                                 const Stack st, int cap, bool& overflow) {
    const CwRay r = cw_make_ray(o, d);
    const uint32_t octinv = r.octinv4 & 0xffu;
    int sp = 0;
    float t_best = t_far;
    int id_best = 0x7fffffff;
    uint32_t ng_x = 0, ng_y = 0x80000000u;  // node group: (child base, hits in bits 24..31 | imask in bits 0..7); starts at the root
    uint32_t tg_x = 0, tg_y = 0;            // triangle group: (node index, pending triangle bits 3*slot+k)
    for (;;) {
        if (ng_y > 0x00ffffffu) {
            const uint32_t hits = ng_y;
            const int child_bit = cw_bfind(hits);
            const uint32_t base = ng_x;
            ng_y &= ~(1u << child_bit);
            if (ng_y > 0x00ffffffu) {
                if (sp < cap) { st.put(sp, ng_x, ng_y); ++sp; }
                else overflow = true;
            }
            const uint32_t slot = (uint32_t)(child_bit - 24) ^ octinv;
            const uint32_t rel = (uint32_t)cw_popc(hits & ~(0xffffffffu << slot));
            const float4* p = nodes + CW_NODE_F4 * (int64_t)(base + rel);
            const float4 n0 = ldg(p), n1 = ldg(p + 1), n2 = ldg(p + 2), n3 = ldg(p + 3), n4 = ldg(p + 4);
            DRP_COUNT_NODE();
            const uint32_t hit8 = cw_node_hits(r, n0, n2, n3, n4, t_best * DRP_T_GROW);
            const uint32_t imask = f2u(n0.w) >> 24;
            ng_x = f2u(n1.x);
            ng_y = (cw_perm8(hit8 & imask, octinv) << 24) | imask;
            tg_x = base + rel;
            tg_y = cw_spread3x7(hit8 & ~imask) & f2u(n1.z);
        } else {
            tg_x = ng_x; tg_y = ng_y;
            ng_x = 0; ng_y = 0;
        }
        if (tg_y != 0) {
            const float4 n1 = ldg(nodes + CW_NODE_F4 * (int64_t)tg_x + 1);
            while (tg_y != 0) {
                const int ti = cw_bfind(tg_y);
                tg_y &= ~(1u << ti);
                leaf_intersect(tris, cw_tri_index(f2u(n1.y), f2u(n1.z), ti), 1, r.o, r.d, eps, t_best, id_best);
            }
        }
        if (ng_y <= 0x00ffffffu) {
            if (sp == 0) break;
            --sp;
            st.get(sp, ng_x, ng_y);
        }
    }
    RayHit h;
    const bool hit = t_best < t_far;
    h.t = hit ? t_best : t_far;
    h.id = hit ? id_best : 0;
    return h;
}
DRP_HD RayHit cw_trace_one(const float4* __restrict__ nodes, const float4* __restrict__ tris, Vec3 o, Vec3 d, float t_far, float eps,
                           bool& overflow) {
    uint32_t st_x[CW_STACK], st_y[CW_STACK];
    return cw_trace_one_stack(nodes, tris, o, d, t_far, eps, CwPairStack{st_x, st_y}, CW_STACK, overflow);
}

// ---- refit: new boxes and triangle records for an existing topology ------------------------------------------------
// Re-derive wide node `src` (5 float4 of a template hierarchy) for new vertex positions and write it as node `dst_ni` of the output
// hierarchy.  The tree is kept -- which children a node has, which triangles a leaf child holds -- while child / triangle bases are shifted
// by `node_off` / `tri_off` and primitive ids by `prim_off` (0 for an in-place refit; block offsets when the template is replicated for
// one instance of a shared mesh).  Internal children must already have been refitted (deeper levels first): their exact boxes are read
// from node_box.  Triangles of leaf children are re-packed from (verts, tris) -- rounded exactly like the builder's -- and their padded
// boxes enter the node, so the slab tests stay conservative and closest hits stay the exhaustive ones, whatever the hierarchy.
// The SLOTS are re-assigned for the new boxes (cw_assign_slots: slot = octant direction, which is what orders the traversal front to back):
// a rotated instance or a deformed mesh would otherwise be walked in an arbitrary order.  Re-slotting permutes the node's children, so the
// (already refitted) child records and their boxes are permuted inside the node's contiguous child run and the triangle records are written
// in the new slot order; nothing outside the node's own children / triangles moves.
DRP_HD void cw_refit_node(const float4* src, const float4* src_tris, float4* out_nodes, float4* out_tris,   // (may alias: in-place refit)
                          float4* node_box, int dst_ni, int node_off, int tri_off, int prim_off, const float* __restrict__ verts,
                          const int32_t* __restrict__ tris, float abs_pad) {
    const uint32_t imask = f2u(src[0].w) >> 24, vmask = f2u(src[1].z);
    const int cb = f2i(src[1].x) + node_off, tb = f2i(src[1].y);
    // children in old slot order: box, kind, payload (internal: old position in the child run; leaf: primitive ids)
    float lo[8][3], hi[8][3], nlo[3] = {3e38f, 3e38f, 3e38f}, nhi[3] = {-3e38f, -3e38f, -3e38f};
    int inner_pos[8], prim[8][3], cnt[8];
    int k = 0, rank = 0;
    for (int s = 0; s < 8; ++s) {
        if (imask & (1u << s)) {
            const int child = cb + rank;
            const float4 l = node_box[2 * (int64_t)child], h = node_box[2 * (int64_t)child + 1];
            lo[k][0] = l.x; lo[k][1] = l.y; lo[k][2] = l.z;
            hi[k][0] = h.x; hi[k][1] = h.y; hi[k][2] = h.z;
            inner_pos[k] = rank++;
            cnt[k] = 0;
        } else {
            const uint32_t bits = (vmask >> (3 * s)) & 7u;
            if (!bits) continue;
            float4 l = make_float4(3e38f, 3e38f, 3e38f, 0.0f), h = make_float4(-3e38f, -3e38f, -3e38f, 0.0f);
            int c = 0;
            for (int q = 0; q < 3; ++q) {
                if (!(bits & (1u << q))) continue;
                const int j = cw_tri_index((uint32_t)tb, vmask, 3 * s + q);
                const int pr = f2i(src_tris[3 * (int64_t)j + 2].y) + prim_off;
                prim[k][c++] = pr;
                const Vec3 A = load_vert(verts, tris[3 * (int64_t)pr]), B = load_vert(verts, tris[3 * (int64_t)pr + 1]), C = load_vert(verts, tris[3 * (int64_t)pr + 2]);
                l.x = fminf(l.x, fminf(fminf(A.x, B.x), C.x)); l.y = fminf(l.y, fminf(fminf(A.y, B.y), C.y)); l.z = fminf(l.z, fminf(fminf(A.z, B.z), C.z));
                h.x = fmaxf(h.x, fmaxf(fmaxf(A.x, B.x), C.x)); h.y = fmaxf(h.y, fmaxf(fmaxf(A.y, B.y), C.y)); h.z = fmaxf(h.z, fmaxf(fmaxf(A.z, B.z), C.z));
            }
            pad_box(l, h, abs_pad);
            lo[k][0] = l.x; lo[k][1] = l.y; lo[k][2] = l.z;
            hi[k][0] = h.x; hi[k][1] = h.y; hi[k][2] = h.z;
            inner_pos[k] = -1;
            cnt[k] = c;
        }
        for (int a = 0; a < 3; ++a) { nlo[a] = fminf(nlo[a], lo[k][a]); nhi[a] = fmaxf(nhi[a], hi[k][a]); }
        ++k;
    }
    int slot_child[8];
    cw_assign_slots(lo, hi, k, nlo, nhi, slot_child);
    // new masks, triangles in the new slot order, permutation of the internal children
    uint32_t new_imask = 0, new_vmask = 0;
    float slo[8][3], shi[8][3];
    int new_of_old[8], n_inner = 0, toff = 0;
    for (int s = 0; s < 8; ++s) {
        const int j = slot_child[s];
        if (j < 0) continue;
        for (int a = 0; a < 3; ++a) { slo[s][a] = lo[j][a]; shi[s][a] = hi[j][a]; }
        if (inner_pos[j] >= 0) {
            new_imask |= 1u << s;
            new_of_old[inner_pos[j]] = n_inner++;
        } else {
            new_vmask |= ((1u << cnt[j]) - 1u) << (3 * s);
            for (int c = 0; c < cnt[j]; ++c) {
                const int pr = prim[j][c];
                const Vec3 A = load_vert(verts, tris[3 * (int64_t)pr]), B = load_vert(verts, tris[3 * (int64_t)pr + 1]), C = load_vert(verts, tris[3 * (int64_t)pr + 2]);
                pack_triangle(out_tris + 3 * (int64_t)(tb + tri_off + toff + c), A, B, C, pr);
            }
            toff += cnt[j];
        }
    }
    bool moved = false;
    for (int r = 0; r < n_inner; ++r) moved = moved || new_of_old[r] != r;
    if (moved) {   // permute the child records (and their boxes) inside this node's child run
        float4 rec[8][CW_NODE_F4], box[8][2];
        for (int r = 0; r < n_inner; ++r) {
            for (int f = 0; f < CW_NODE_F4; ++f) rec[r][f] = out_nodes[CW_NODE_F4 * (int64_t)(cb + r) + f];
            box[r][0] = node_box[2 * (int64_t)(cb + r)]; box[r][1] = node_box[2 * (int64_t)(cb + r) + 1];
        }
        for (int r = 0; r < n_inner; ++r) {
            const int d = cb + new_of_old[r];
            for (int f = 0; f < CW_NODE_F4; ++f) out_nodes[CW_NODE_F4 * (int64_t)d + f] = rec[r][f];
            node_box[2 * (int64_t)d] = box[r][0]; node_box[2 * (int64_t)d + 1] = box[r][1];
        }
    }
    float elo[3], ehi[3];
    cw_encode_node(out_nodes + CW_NODE_F4 * (int64_t)dst_ni, slo, shi, new_imask, new_vmask, cb, tb + tri_off, elo, ehi);
    node_box[2 * (int64_t)dst_ni] = make_float4(elo[0], elo[1], elo[2], 0.0f);
    node_box[2 * (int64_t)dst_ni + 1] = make_float4(ehi[0], ehi[1], ehi[2], 0.0f);
}
