// instance_level.h -- host-side build of the instance level of drp_build_instanced (plain C++: shared by api.cu and the test-only host simulator).
#pragma once
#include <algorithm>
#include <utility>
#include <vector>
#include "common.cuh"
#include "cwbvh.cuh"

// Host-side instance level: a wide hierarchy over the (few) instance root boxes, top-down by median splits of the longest axis into up to 8
// groups.  Leaves of this level are the instances' own root nodes, copied next to their siblings (children of a wide node are contiguous).
struct TlasBuilder {
    const std::vector<float>& lo;   // (n_inst, 3) root boxes
    const std::vector<float>& hi;
    std::vector<float4> nodes;      // records of the instance level (positions that receive a copied instance root stay zero)
    std::vector<std::pair<int, int>> copies;   // (instance, destination node index)
    int allocated = 1;              // node 0 = root
    TlasBuilder(const std::vector<float>& l, const std::vector<float>& h_) : lo(l), hi(h_) {}

    void group_box(const int* ids, int n, float blo[3], float bhi[3]) const {
        for (int a = 0; a < 3; ++a) { blo[a] = 3e38f; bhi[a] = -3e38f; }
        for (int k = 0; k < n; ++k)
            for (int a = 0; a < 3; ++a) { blo[a] = fminf(blo[a], lo[3 * (size_t)ids[k] + a]); bhi[a] = fmaxf(bhi[a], hi[3 * (size_t)ids[k] + a]); }
    }
    // ids[0..n): instances under node `ni` (n >= 2, or the root with n >= 1)
    void build(int ni, int* ids, int n) {
        // split into <= 8 groups: repeatedly halve the largest group at the median of its longest centroid axis
        struct Group { int first, count; };
        Group g[8];
        int k = 1;
        g[0] = {0, n};
        while (k < 8) {
            int best = -1;
            for (int q = 0; q < k; ++q) if (g[q].count > 1 && (best < 0 || g[q].count > g[best].count)) best = q;
            if (best < 0) break;
            int* p = ids + g[best].first;
            const int cnt = g[best].count;
            float clo[3] = {3e38f, 3e38f, 3e38f}, chi[3] = {-3e38f, -3e38f, -3e38f};
            for (int t = 0; t < cnt; ++t)
                for (int a = 0; a < 3; ++a) {
                    const float c = lo[3 * (size_t)p[t] + a] + hi[3 * (size_t)p[t] + a];
                    clo[a] = fminf(clo[a], c); chi[a] = fmaxf(chi[a], c);
                }
            int ax = 0;
            if (chi[1] - clo[1] > chi[ax] - clo[ax]) ax = 1;
            if (chi[2] - clo[2] > chi[ax] - clo[ax]) ax = 2;
            const int half = cnt / 2;
            std::nth_element(p, p + half, p + cnt, [&](int x, int y) { return lo[3 * (size_t)x + ax] + hi[3 * (size_t)x + ax] < lo[3 * (size_t)y + ax] + hi[3 * (size_t)y + ax]; });
            g[k] = {g[best].first + half, cnt - half};
            g[best].count = half;
            ++k;
        }
        float blo[8][3], bhi[8][3], nlo[3] = {3e38f, 3e38f, 3e38f}, nhi[3] = {-3e38f, -3e38f, -3e38f};
        for (int q = 0; q < k; ++q) {
            group_box(ids + g[q].first, g[q].count, blo[q], bhi[q]);
            for (int a = 0; a < 3; ++a) { nlo[a] = fminf(nlo[a], blo[q][a]); nhi[a] = fmaxf(nhi[a], bhi[q][a]); }
        }
        int slot_child[8];
        cw_assign_slots(blo, bhi, k, nlo, nhi, slot_child);
        float slo[8][3], shi[8][3];
        uint32_t imask = 0;
        for (int s = 0; s < 8; ++s) {
            if (slot_child[s] < 0) continue;
            imask |= 1u << s;
            for (int a = 0; a < 3; ++a) { slo[s][a] = blo[slot_child[s]][a]; shi[s][a] = bhi[slot_child[s]][a]; }
        }
        const int child_base = allocated;
        allocated += k;
        if ((size_t)allocated * CW_NODE_F4 > nodes.size()) nodes.resize((size_t)allocated * CW_NODE_F4, make_float4(0, 0, 0, 0));
        float tlo[3], thi[3];
        cw_encode_node(nodes.data() + (size_t)ni * CW_NODE_F4, slo, shi, imask, 0u, child_base, 0, tlo, thi);
        int rank = 0;
        for (int s = 0; s < 8; ++s) {
            const int q = slot_child[s];
            if (q < 0) continue;
            const int child = child_base + rank++;
            if (g[q].count == 1) copies.push_back({ids[g[q].first], child});
            else build(child, ids + g[q].first, g[q].count);
        }
    }
};


// Root of an instance level over ONE instance: a node whose single internal child (slot 0) is the instance's root, copied to node 1.
inline void tlas_single(TlasBuilder& tb, const float lo[3], const float hi[3]) {
    float slo[8][3], shi[8][3], tlo[3], thi[3];
    for (int a = 0; a < 3; ++a) { slo[0][a] = lo[a]; shi[0][a] = hi[a]; }
    tb.allocated = 2;
    tb.nodes.resize(2 * CW_NODE_F4, make_float4(0, 0, 0, 0));
    cw_encode_node(tb.nodes.data(), slo, shi, 1u, 0u, 1, 0, tlo, thi);
    tb.copies.push_back({0, 1});
}
