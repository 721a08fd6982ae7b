// flatten.cu -- drp_flatten: every object of a Scene concatenated into one set of vertex / index buffers in one pass.
// Replaces the per-object torch ops of RenderSessionMixin.vertex_array_object (diffrp/rendering/mixin.py:74-113) and the
// per-material world transforms of SurfaceInput.interpolate_ex (base_material.py:137-143: 'vectornor', 'vector3norex1').
//
// Source attribute pointers may be device memory or PINNED HOST memory: the kernels read them through unified addressing, so
// for a host-resident scene the read over PCIe *is* the upload -- transform, concatenation, index offsetting, stencil /
// material tagging and the interleaved 64-byte shading records are produced in the same pass, with no staging copies and no
// per-object launches (the torch version issues ~15 small ops per object).
#include <vector>
#include <cstring>
#include "internal.h"
#include "common.cuh"

struct FlattenArgs {
    const drp_object_t* objects;  // device copy of the descriptor table
    const int64_t* vert_offsets;  // (n_objects + 1)
    const int64_t* tri_offsets;   // (n_objects + 1)
    int n_objects;
    float *world_pos, *world_nrm, *color4, *uv, *world_tan, *records, *verts_raw, *normals_raw, *tangents_raw;
    int32_t *tris, *tri_material, *stencils;
};

__device__ __forceinline__ int find_object(const int64_t* __restrict__ offsets, int n, int64_t i) {
    int lo = 0, hi = n - 1;  // largest k with offsets[k] <= i
    while (lo < hi) {
        int mid = (lo + hi + 1) >> 1;
        if (__ldg(offsets + mid) <= i) lo = mid; else hi = mid - 1;
    }
    return lo;
}
__device__ __forceinline__ Vec3 mul3x3(const float* M, Vec3 v) {  // rows of the upper-left 3x3 of a row-major 4x4
    return v3(M[0] * v.x + M[1] * v.y + M[2] * v.z, M[4] * v.x + M[5] * v.y + M[6] * v.z, M[8] * v.x + M[9] * v.y + M[10] * v.z);
}

__global__ void __launch_bounds__(256) k_flatten_verts(FlattenArgs a, int64_t n_verts) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_verts) return;
    const int k = find_object(a.vert_offsets, a.n_objects, i);
    const drp_object_t& o = a.objects[k];
    const int64_t j = i - __ldg(a.vert_offsets + k);
    const Vec3 p = v3(o.verts[3 * j], o.verts[3 * j + 1], o.verts[3 * j + 2]);
    const Vec3 n = v3(o.normals[3 * j], o.normals[3 * j + 1], o.normals[3 * j + 2]);
    const float4 tg = make_float4(o.tangents[4 * j], o.tangents[4 * j + 1], o.tangents[4 * j + 2], o.tangents[4 * j + 3]);
    const float2 uv = make_float2(o.uv[2 * j], o.uv[2 * j + 1]);
    float4 col;
    if (o.color_channels == 4) col = make_float4(o.color[4 * j], o.color[4 * j + 1], o.color[4 * j + 2], o.color[4 * j + 3]);
    else col = make_float4(o.color[3 * j], o.color[3 * j + 1], o.color[3 * j + 2], 1.0f);
    Vec3 wp = mul3x3(o.M, p);                                       // transform_point4x3
    wp = v3(wp.x + o.M[3], wp.y + o.M[7], wp.z + o.M[11]);
    const Vec3 wn = normalize_ref(mul3x3(o.M, n));                   // 'vectornor': M, not the inverse transpose (reference behaviour)
    const Vec3 wt = normalize_ref(mul3x3(o.M, v3(tg.x, tg.y, tg.z)));  // 'vector3norex1'
    a.world_pos[3 * i] = wp.x; a.world_pos[3 * i + 1] = wp.y; a.world_pos[3 * i + 2] = wp.z;
    a.world_nrm[3 * i] = wn.x; a.world_nrm[3 * i + 1] = wn.y; a.world_nrm[3 * i + 2] = wn.z;
    reinterpret_cast<float4*>(a.color4)[i] = col;
    reinterpret_cast<float2*>(a.uv)[i] = uv;
    reinterpret_cast<float4*>(a.world_tan)[i] = make_float4(wt.x, wt.y, wt.z, tg.w);
    if (a.records) {
        float4* r = reinterpret_cast<float4*>(a.records) + 4 * i;
        r[0] = make_float4(wp.x, wp.y, wp.z, wn.x);
        r[1] = make_float4(wn.y, wn.z, uv.x, uv.y);
        r[2] = col;
        r[3] = make_float4(wt.x, wt.y, wt.z, tg.w);
    }
    if (a.verts_raw) { a.verts_raw[3 * i] = p.x; a.verts_raw[3 * i + 1] = p.y; a.verts_raw[3 * i + 2] = p.z; }
    if (a.normals_raw) { a.normals_raw[3 * i] = n.x; a.normals_raw[3 * i + 1] = n.y; a.normals_raw[3 * i + 2] = n.z; }
    if (a.tangents_raw) reinterpret_cast<float4*>(a.tangents_raw)[i] = tg;
}

__global__ void __launch_bounds__(256) k_flatten_tris(FlattenArgs a, int64_t n_tris) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) a.stencils[0] = 0;  // stencil 0 = miss (mixin.py:78)
    if (i >= n_tris) return;
    const int k = find_object(a.tri_offsets, a.n_objects, i);
    const drp_object_t& o = a.objects[k];
    const int64_t j = i - __ldg(a.tri_offsets + k);
    const int32_t off = (int32_t)__ldg(a.vert_offsets + k);
    a.tris[3 * i] = o.tris[3 * j] + off;
    a.tris[3 * i + 1] = o.tris[3 * j + 1] + off;
    a.tris[3 * i + 2] = o.tris[3 * j + 2] + off;
    a.tri_material[i] = k;
    a.stencils[i + 1] = k + 1;
}

extern "C" int drp_flatten(const drp_object_t* objects, int32_t n_objects, float* world_pos, float* world_nrm, float* color4, float* uv,
                           float* world_tan, int32_t* tris, int32_t* tri_material, int32_t* stencils, float* records, float* verts_raw,
                           float* normals_raw, float* tangents_raw, void* stream) {
    if (n_objects < 0 || (n_objects > 0 && !objects)) { drp_set_error("drp_flatten: invalid argument"); return DRP_ERR_INVALID; }
    if (!stencils) { drp_set_error("drp_flatten: stencils is required"); return DRP_ERR_INVALID; }
    cudaStream_t s = (cudaStream_t)stream;
    std::vector<int64_t> voff(n_objects + 1, 0), toff(n_objects + 1, 0);
    for (int k = 0; k < n_objects; ++k) {
        const drp_object_t& o = objects[k];
        if (o.n_verts < 0 || o.n_tris < 0 || (o.color_channels != 3 && o.color_channels != 4)) { drp_set_error("drp_flatten: bad object descriptor"); return DRP_ERR_INVALID; }
        voff[k + 1] = voff[k] + o.n_verts;
        toff[k + 1] = toff[k] + o.n_tris;
    }
    const int64_t V = voff[n_objects], F = toff[n_objects];
    if (V >= (int64_t(1) << 31)) { drp_set_error("drp_flatten: more than 2^31 vertices"); return DRP_ERR_INVALID; }
    if (V > 0 && (!world_pos || !world_nrm || !color4 || !uv || !world_tan)) { drp_set_error("drp_flatten: missing vertex output"); return DRP_ERR_INVALID; }
    if (F > 0 && (!tris || !tri_material)) { drp_set_error("drp_flatten: missing index output"); return DRP_ERR_INVALID; }
    // descriptor table + offsets -> device (stream-ordered scratch)
    const size_t ob = sizeof(drp_object_t) * (size_t)(n_objects > 0 ? n_objects : 1), fb = sizeof(int64_t) * (size_t)(n_objects + 1);
    char* scratch = nullptr;
    DRP_CUDA_CHECK(cudaMallocAsync((void**)&scratch, ob + 2 * fb, s));
    if (n_objects > 0) DRP_CUDA_CHECK(cudaMemcpyAsync(scratch, objects, sizeof(drp_object_t) * (size_t)n_objects, cudaMemcpyHostToDevice, s));
    DRP_CUDA_CHECK(cudaMemcpyAsync(scratch + ob, voff.data(), fb, cudaMemcpyHostToDevice, s));
    DRP_CUDA_CHECK(cudaMemcpyAsync(scratch + ob + fb, toff.data(), fb, cudaMemcpyHostToDevice, s));
    DRP_CUDA_CHECK(cudaStreamSynchronize(s));  // the pageable host vectors above go out of scope when we return
    FlattenArgs a;
    a.objects = (const drp_object_t*)scratch;
    a.vert_offsets = (const int64_t*)(scratch + ob);
    a.tri_offsets = (const int64_t*)(scratch + ob + fb);
    a.n_objects = n_objects;
    a.world_pos = world_pos; a.world_nrm = world_nrm; a.color4 = color4; a.uv = uv; a.world_tan = world_tan; a.records = records;
    a.verts_raw = verts_raw; a.normals_raw = normals_raw; a.tangents_raw = tangents_raw;
    a.tris = tris; a.tri_material = tri_material; a.stencils = stencils;
    if (V > 0) k_flatten_verts<<<(unsigned)((V + 255) / 256), 256, 0, s>>>(a, V);
    k_flatten_tris<<<(unsigned)((std::max<int64_t>(F, 1) + 255) / 256), 256, 0, s>>>(a, F);
    DRP_CUDA_CHECK(cudaGetLastError());
    DRP_CUDA_CHECK(cudaFreeAsync(scratch, s));
    return DRP_OK;
}


// Batched host -> device upload: n stream-ordered copies issued from one C call.  A scene is ~300 small tensors (6 per object, 4 per
// material); issuing them one by one from Python costs more host time than the PCIe transfer takes (tools/e2e_phases.py).
extern "C" int drp_upload_batch(int32_t n, void* const* dst, const void* const* src, const int64_t* bytes, void* stream) {
    if (n < 0 || (n > 0 && (!dst || !src || !bytes))) { drp_set_error("drp_upload_batch: invalid argument"); return DRP_ERR_INVALID; }
    cudaStream_t s = (cudaStream_t)stream;
    for (int32_t k = 0; k < n; ++k) {
        if (bytes[k] < 0 || (bytes[k] > 0 && (!dst[k] || !src[k]))) { drp_set_error("drp_upload_batch: invalid segment"); return DRP_ERR_INVALID; }
        if (bytes[k] > 0) DRP_CUDA_CHECK(cudaMemcpyAsync(dst[k], src[k], (size_t)bytes[k], cudaMemcpyDefault, s));
    }
    return DRP_OK;
}
