// internal.h -- host-side state shared by the translation units of libdiffrp_b200.so
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include <vector>
#include "../../include/diffrp_b200.h"

struct RenderWorkspace;

struct BvhHandle {
    int device = 0;
    int64_t n_tris = 0;
    int64_t n_nodes = 0;       // allocated
    int64_t n_nodes_used = 0;  // wide layout: nodes actually emitted
    float eps = 1e-8f;          // |det| threshold of the triangle test (raycaster_epsilon)
    float4* nodes = nullptr;    // (n_nodes, 5): compressed 8-wide layout (cwbvh.cuh)
    float4* node_box = nullptr; // (n_nodes, 2) exact (lo, hi) box of every node: scratch of the refit / instanced assembly, allocated on first use
    std::vector<int> level_begin;  // nodes of level L: [level_begin[L], level_begin[L + 1]) -- breadth-first allocation order of the collapse
    float4* packed = nullptr;   // (n_tris, 3) triangles in leaf order
    uint32_t* bounds = nullptr; // (12) ordered-uint scene bounds (device)
    float* sah = nullptr;       // device: root cost / root area
    int* sticky_host = nullptr; // host-mapped failure flag (see drp_check_sticky): rays that outgrew even the deep traversal stack
    int* sticky_dev = nullptr;  // device alias of sticky_host
    int stack_cap = 0;          // per-thread stack entries of the fast traversal path (CW_STACK; lowered only by drp_debug_set_stack_limit)
    RenderWorkspace* ws = nullptr;
    drp_render_stats_t last_render = {0, 0, 0};
};

void drp_set_error(const std::string& msg);
BvhHandle* drp_lookup(uint64_t handle);
void drp_free_workspace(BvhHandle* h);
void drp_invalidate_scene_box(BvhHandle* h);   // after a refit: the cached compaction box of the workspace is stale
int drp_check_sticky(BvhHandle* h, const char* who);
int drp_build_structure(BvhHandle* h, const float* verts, const int32_t* tris, int64_t n_tris, cudaStream_t s);
BvhHandle* drp_new_handle(int device);
void drp_register_handle(BvhHandle* h, uint64_t* out_handle);
void drp_destroy_handle(BvhHandle* h);
int drp_trace_wide_persistent(BvhHandle* h, const float* ro, const float* rd, float* out_t, int32_t* out_i, float t_far, int64_t n, cudaStream_t s);
extern int g_drp_log_level;

#define DRP_CUDA_CHECK(expr)                                                                      \
    do {                                                                                          \
        cudaError_t _e = (expr);                                                                  \
        if (_e != cudaSuccess) {                                                                  \
            drp_set_error(std::string(#expr) + ": " + cudaGetErrorString(_e));                    \
            return DRP_ERR_CUDA;                                                                  \
        }                                                                                         \
    } while (0)

struct DeviceGuard {
    int prev = -1;
    bool ok = true;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) { ok = false; return; }
        if (prev != dev && cudaSetDevice(dev) != cudaSuccess) ok = false;
    }
    ~DeviceGuard() {
        int cur = -1;
        if (prev >= 0 && cudaGetDevice(&cur) == cudaSuccess && cur != prev) cudaSetDevice(prev);
    }
};
