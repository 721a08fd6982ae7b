// shade.cuh -- per-ray device logic of the fused wavefront: primary-ray generation, surface attributes
// (DefaultMaterial / GLTFMaterial), environment lookup and the metallic-roughness GGX sampler.
// One DRP_HD function per reference op group (citations into eliphatfs/diffrp v0.2.7):
//   gen_primary_ray   mixin.py:31-39, coordinates.py:6-10, path_tracing.py:329-331
//   tex_fetch         shader_ops.py:198-221 (F.grid_sample, align_corners=False), gltf_material.py:15-22
//   env_fetch         coordinates.py:60-71, path_tracing.py:238-248,267
//   surface_attrs     path_tracing.py:158-187, geometry.py:94-110, interpolator.py:32-48, base_material.py:183-257,
//                     default_material.py:21-24, gltf_material.py:48-67, mixin.py:115-128
//   brdf_sample       path_tracing.py:189-236, light_transport.py:35-43,72-83,179-196, shader_ops.py:313-325
#pragma once
#include "common.cuh"
#include "../../include/diffrp_b200.h"

#define DRP_TAU 6.283185307179586f
#define DRP_PI 3.141592653589793f

struct SurfaceAttrs {  // the (R,12) g-buffer row of _super_collector, path_tracing.py:180-187
    Vec3 albedo;
    Vec3 normal;
    float metal, smooth, alpha;
    Vec3 emission;
};

struct BounceOut {
    Vec3 radiance;  // emission + env
    Vec3 transfer;
    Vec3 hit_pos;   // o + d t   (extras['world_position'])
    Vec3 next_d;
};

DRP_HD float clampf(float x, float lo, float hi) { return fminf(hi, fmaxf(x, lo)); }

// ---- primary rays ---------------------------------------------------------------------------------------------
DRP_HD void gen_primary_ray(const float* __restrict__ inv_vp, const float* __restrict__ cam_pos, float t_near, float gx, float gy,
                            Vec3& o, Vec3& d) {
    // grid = (gx, gy, -1, 1);  p = grid @ inv(VP)^T;  p.xyz / p.w;  d = normalize(p - cam);  o = cam + d * near.
    // p - cam cancels ~5 digits (the near plane is 0.1 from the eye), so the rounding ORDER of the 4-term dot product
    // shows up at 1e-5 in the hit position.  Explicitly rounded, left-to-right, unfused operations reproduce the
    // oracle / CPU reference sequence bit for bit.
    float q[4];
#pragma unroll
    for (int k = 0; k < 4; ++k)
        q[k] = x_add(x_sub(x_add(x_mul(gx, inv_vp[4 * k]), x_mul(gy, inv_vp[4 * k + 1])), inv_vp[4 * k + 2]), inv_vp[4 * k + 3]);
    Vec3 c = v3(cam_pos[0], cam_pos[1], cam_pos[2]);
    Vec3 v = x_sub3(v3(q[0] / q[3], q[1] / q[3], q[2] / q[3]), c);
    float len = fmaxf(sqrtf(x_dot3(v, v)), 1e-12f);
    d = v3(v.x / len, v.y / len, v.z / len);
    o = v3(x_add(c.x, x_mul(d.x, t_near)), x_add(c.y, x_mul(d.y, t_near)), x_add(c.z, x_mul(d.z, t_near)));
}

// ---- textures -------------------------------------------------------------------------------------------------
// F.grid_sample(align_corners=False) semantics.  The device path is split into a branch-free address stage (tex_taps:
// four texel offsets + weights) and a load stage (one 128-bit load per tap on RGBA texels), so that the 16 texel loads of a
// GLTF material are independent instructions the scheduler can keep in flight together (the shade kernel is
// latency-bound: ncu long_scoreboard), and the address code exists once instead of being inlined per texture.
struct TexTaps {
    int off[4];   // texel index y * w + x of the four taps (always in range)
    float wt[4];  // bilinear weight; 0 for taps outside the image
};

DRP_HD float reflect_coord(float in, float twice_low, float twice_high) {
    // reflect `in` into [low, high] given 2*low and 2*high (ATen reflect_coordinates); span > 0 here
    float mn = twice_low * 0.5f, span = (twice_high - twice_low) * 0.5f;
    in = fabsf(in - mn);
    float flips = floorf(in / span);
    float extra = in - flips * span;
    return (((int)flips) & 1) ? span - extra + mn : extra + mn;
}
DRP_HD float grid_coord(float g, int size, bool reflection) {
    float c = ((g + 1.0f) * (float)size - 1.0f) * 0.5f;
    float r = reflect_coord(c, -1.0f, (float)(2 * size - 1));
    c = reflection ? r : c;
    return clampf(c, 0.0f, (float)(size - 1));
}
#ifdef __CUDA_ARCH__
__device__ __noinline__
#else
inline
#endif
TexTaps tex_taps(int h, int w, int wrap, int interp, float u, float v, int tile_log2 = 0) {
    TexTaps t;
    if (wrap == DRP_WRAP_REPEAT) {  // uv.remainder(1.0)
        u = u - floorf(u); v = v - floorf(v);
        u = u >= 1.0f ? 0.0f : u;
        v = v >= 1.0f ? 0.0f : v;
    }
    const bool reflection = wrap != DRP_WRAP_CLAMP;
    float ix = grid_coord(u * 2.0f - 1.0f, w, reflection);
    float iy = grid_coord(-(v * 2.0f - 1.0f), h, reflection);
    if (interp == DRP_INTERP_POINT) { ix = nearbyintf(ix); iy = nearbyintf(iy); }
    const float x0 = floorf(ix), y0 = floorf(iy);
    const float fx = ix - x0, fy = iy - y0, gx = (x0 + 1.0f) - ix, gy = (y0 + 1.0f) - iy;
    const int X = (int)x0, Y = (int)y0;
    const bool xin = X + 1 < w, yin = Y + 1 < h;  // X, Y themselves are in range after the clamp
    const int X1 = xin ? X + 1 : X, Y1 = yin ? Y + 1 : Y;
    if (tile_log2 > 0) {   // tiled texel records (drp_material_t.texel_tile_log2): tile-major, row-major inside a 2^L x 2^L tile
        const int L = tile_log2, mask = (1 << L) - 1, tw = w >> L;
        const int ty0 = (Y >> L) * tw, ty1 = (Y1 >> L) * tw, iy0 = (Y & mask) << L, iy1 = (Y1 & mask) << L;
        const int tx0 = X >> L, tx1 = X1 >> L, ix0 = X & mask, ix1 = X1 & mask;
        t.off[0] = ((ty0 + tx0) << (2 * L)) | iy0 | ix0;
        t.off[1] = ((ty0 + tx1) << (2 * L)) | iy0 | ix1;
        t.off[2] = ((ty1 + tx0) << (2 * L)) | iy1 | ix0;
        t.off[3] = ((ty1 + tx1) << (2 * L)) | iy1 | ix1;
    } else {
        t.off[0] = Y * w + X;
        t.off[1] = Y * w + X1;
        t.off[2] = Y1 * w + X;
        t.off[3] = Y1 * w + X1;
    }
    t.wt[0] = gx * gy;
    t.wt[1] = xin ? fx * gy : 0.0f;
    t.wt[2] = yin ? gx * fy : 0.0f;
    t.wt[3] = (xin && yin) ? fx * fy : 0.0f;
    return t;
}

// generic-channel fetch (host oracle-compatible path; textures with c in {1,3,4})
DRP_HD void tex_fetch(const drp_texture_t& t, float u, float v, float out[4]) {
    out[0] = out[1] = out[2] = out[3] = 0.0f;
    TexTaps tp = tex_taps(t.h, t.w, t.wrap, t.interp, u, v);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        if (t.c == 4) {
            float4 x = ldg(reinterpret_cast<const float4*>(t.data) + tp.off[k]);
            out[0] += x.x * tp.wt[k]; out[1] += x.y * tp.wt[k]; out[2] += x.z * tp.wt[k]; out[3] += x.w * tp.wt[k];
        } else {
            for (int c = 0; c < t.c; ++c) out[c] += ldg(t.data + (int64_t)tp.off[k] * t.c + c) * tp.wt[k];
        }
    }
}
// weighted sum of the four taps (the order of tex_fetch: tap 0..3)
DRP_HD float4 tex_combine(float4 a, float4 b, float4 c, float4 d, const TexTaps& tp) {
    return make_float4(a.x * tp.wt[0] + b.x * tp.wt[1] + c.x * tp.wt[2] + d.x * tp.wt[3],
                       a.y * tp.wt[0] + b.y * tp.wt[1] + c.y * tp.wt[2] + d.y * tp.wt[3],
                       a.z * tp.wt[0] + b.z * tp.wt[1] + c.z * tp.wt[2] + d.z * tp.wt[3],
                       a.w * tp.wt[0] + b.w * tp.wt[1] + c.w * tp.wt[2] + d.w * tp.wt[3]);
}
// part `part` (0..2) of the 48-byte interleaved texels (drp_material_t.texel_records)
DRP_HD float4 tex_fetch_record(const float* __restrict__ records, const TexTaps& tp, int part) {
    const float4* p = reinterpret_cast<const float4*>(records) + part;
    float4 a = ldg(p + 3 * (int64_t)tp.off[0]), b = ldg(p + 3 * (int64_t)tp.off[1]), c = ldg(p + 3 * (int64_t)tp.off[2]), d = ldg(p + 3 * (int64_t)tp.off[3]);
    return tex_combine(a, b, c, d, tp);
}
// RGBA-only fetch from precomputed taps: four independent 128-bit loads
DRP_HD float4 tex_fetch4(const float* __restrict__ data, const TexTaps& tp) {
    const float4* p = reinterpret_cast<const float4*>(data);
    float4 a = ldg(p + tp.off[0]), b = ldg(p + tp.off[1]), c = ldg(p + tp.off[2]), d = ldg(p + tp.off[3]);
    return tex_combine(a, b, c, d, tp);
}

DRP_HD Vec3 env_fetch(const drp_texture_t& env, Vec3 d) {
    if (!env.data) return v3(0.0f, 0.0f, 0.0f);  // black_tex, path_tracing.py:247
    float a = atan2f(d.x, d.z) * (0.5f / DRP_PI);
    a = a - floorf(a);
    if (a >= 1.0f) a = 0.0f;
    float v = (1.0f / DRP_PI) * asinf(clampf(d.y, -0.999999f, 0.999999f)) + 0.5f;
#ifndef __CUDA_ARCH__
    if (env.c != 4) {
        float rgb[4];
        tex_fetch(env, a, v, rgb);  // env.wrap = CLAMP ('border'), env.interp = LINEAR set by the host
        return v3(rgb[0], rgb[1], rgb[2]);
    }
#endif
    float4 r = tex_fetch4(env.data, tex_taps(env.h, env.w, DRP_WRAP_CLAMP, DRP_INTERP_LINEAR, a, v));  // device: RGBA-padded
    return v3(r.x, r.y, r.z);
}

// ---- surface attributes -----------------------------------------------------------------------------------------
DRP_HD Vec3 ld3(const float* __restrict__ p, int i) { return v3(ldg(p + 3 * (int64_t)i), ldg(p + 3 * (int64_t)i + 1), ldg(p + 3 * (int64_t)i + 2)); }
DRP_HD float4 ld4(const float* __restrict__ p, int i) { return ldg(reinterpret_cast<const float4*>(p) + i); }
DRP_HD float2 ld2(const float* __restrict__ p, int i) { return ldg(reinterpret_cast<const float2*>(p) + i); }
// (v1 - v3) * u + ((v2 - v3) * v + v3), interpolator.py:32-48
DRP_HD float lerp3(float a, float b, float c, float u, float v) { return (a - c) * u + ((b - c) * v + c); }
DRP_HD Vec3 lerp3v(Vec3 a, Vec3 b, Vec3 c, float u, float v) { return v3(lerp3(a.x, b.x, c.x, u, v), lerp3(a.y, b.y, c.y, u, v), lerp3(a.z, b.z, c.z, u, v)); }

DRP_HD void hit_barycentric(Vec3 a, Vec3 b, Vec3 c, Vec3 p, float& u, float& v) {  // geometry.py:94-110
    Vec3 v0 = b - a, v1 = c - a, v2 = p - a;
    float d00 = dot(v0, v0), d01 = dot(v0, v1), d11 = dot(v1, v1), d20 = dot(v2, v0), d21 = dot(v2, v1);
    float denom = d00 * d11 - d01 * d01;
    float bv = (d11 * d20 - d01 * d21) / denom, bw = (d00 * d21 - d01 * d20) / denom;
    if (bv != bv) bv = 0.0f;
    if (bw != bw) bw = 0.0f;
    bv = clampf(bv, 0.0f, 1.0f);
    bw = clampf(bw, 0.0f, 1.0f);
    u = 1.0f - bv - bw;
    v = bv;
}

// per-vertex shading inputs, from the interleaved 64-byte record when the scene has one, else from the five arrays
struct VertexIn {
    Vec3 pos, nrm;
    float2 uv;
    float4 color, tan;
};
DRP_HD VertexIn load_vertex(const drp_scene_t& sc, int i) {
    VertexIn v;
    if (sc.vertex_records) {
        const float4* r = reinterpret_cast<const float4*>(sc.vertex_records) + 4 * (int64_t)i;
        const float4 a = ldg(r), b = ldg(r + 1);
        v.pos = v3(a.x, a.y, a.z);
        v.nrm = v3(a.w, b.x, b.y);
        v.uv = make_float2(b.z, b.w);
        v.color = ldg(r + 2);
        v.tan = ldg(r + 3);
    } else {
        v.pos = ld3(sc.world_pos, i);
        v.nrm = ld3(sc.world_nrm, i);
        v.uv = ld2(sc.uv, i);
        v.color = ld4(sc.color, i);
        v.tan = ld4(sc.world_tan, i);
    }
    return v;
}

// radiance_only: the caller uses nothing but `emission` and `alpha` (last bounce of a path: no next ray is sampled and no g-buffer is
// written, path_tracing.py:336-338), so everything that cannot influence those two is skipped -- for an opaque material without emission that
// is the whole evaluation, including the vertex fetches.  The values that are produced are computed by the same code as in the full mode.
// vertex indices + material id of a triangle; k_shade loads them one iteration ahead of their use
struct TriIdx {
    int i0, i1, i2, mat;
};
DRP_HD TriIdx load_tri_idx(const drp_scene_t& sc, int tri_id) {
    TriIdx x;
    x.i0 = ldg(sc.tris + 3 * (int64_t)tri_id); x.i1 = ldg(sc.tris + 3 * (int64_t)tri_id + 1); x.i2 = ldg(sc.tris + 3 * (int64_t)tri_id + 2);
    x.mat = ldg(sc.tri_material + tri_id);
    return x;
}
DRP_HD SurfaceAttrs surface_attrs(const drp_scene_t& sc, const drp_material_t* __restrict__ mats, Vec3 hit_pos, int tri_id, const bool radiance_only = false,
                                  const TriIdx* pre = nullptr) {
    SurfaceAttrs s;
    const int mat_id = pre ? pre->mat : ldg(sc.tri_material + tri_id);
    if (radiance_only) {
        const drp_material_t& m0 = mats[mat_id];
        const bool need_alpha = m0.kind != DRP_MAT_DEFAULT && (m0.alpha_mode == DRP_ALPHA_MASK || m0.alpha_mode == DRP_ALPHA_BLEND);
        const bool need_em = m0.kind != DRP_MAT_DEFAULT && m0.has_emissive && m0.emissive_tex.data;
        if (!need_alpha && !need_em) {
            s.albedo = s.normal = s.emission = v3(0.0f, 0.0f, 0.0f);
            s.metal = 0.0f; s.smooth = 0.5f; s.alpha = 1.0f;
            return s;
        }
    }
    const int i0 = pre ? pre->i0 : ldg(sc.tris + 3 * (int64_t)tri_id), i1 = pre ? pre->i1 : ldg(sc.tris + 3 * (int64_t)tri_id + 1),
              i2 = pre ? pre->i2 : ldg(sc.tris + 3 * (int64_t)tri_id + 2);
    const VertexIn q0 = load_vertex(sc, i0), q1 = load_vertex(sc, i1), q2 = load_vertex(sc, i2);
    float u, v;
    hit_barycentric(q0.pos, q1.pos, q2.pos, hit_pos, u, v);
    const drp_material_t& m = mats[mat_id];
    Vec3 nu = lerp3v(q0.nrm, q1.nrm, q2.nrm, u, v);  // world_normal_unnormalized
    const float4 c0 = q0.color, c1 = q1.color, c2 = q2.color;
    float col[4] = {lerp3(c0.x, c1.x, c2.x, u, v), lerp3(c0.y, c1.y, c2.y, u, v), lerp3(c0.z, c1.z, c2.z, u, v), lerp3(c0.w, c1.w, c2.w, u, v)};
    s.normal = normalize_ref(nu);
    s.metal = 0.0f; s.smooth = 0.5f; s.alpha = 1.0f;  // path_tracing.py:183-186
    s.emission = v3(0.0f, 0.0f, 0.0f);
    if (m.kind == DRP_MAT_DEFAULT) {
        s.albedo = v3(col[0] * m.tint[0], col[1] * m.tint[1], col[2] * m.tint[2]);
        return s;
    }
    const float2 t0 = q0.uv, t1 = q1.uv, t2 = q2.uv;
    float tu = lerp3(t0.x, t1.x, t2.x, u, v), tv = lerp3(t0.y, t1.y, t2.y, u, v);
    float bc[4] = {1.0f, 1.0f, 1.0f, 1.0f}, mr[4] = {0.0f, 0.0f, 0.0f, 0.0f}, nt[4] = {0.0f, 0.0f, 0.0f, 0.0f}, em[4] = {0.0f, 0.0f, 0.0f, 0.0f};
    const bool use_nt = !radiance_only && m.has_normal_tex && m.normal_tex.data, use_em = m.has_emissive && m.emissive_tex.data;
    const bool use_bc = m.base_color_tex.data && !(radiance_only && m.alpha_mode != DRP_ALPHA_MASK && m.alpha_mode != DRP_ALPHA_BLEND);
    const bool use_mr = m.mr_tex.data && !radiance_only;
#ifdef __CUDA_ARCH__
    const bool rgba = true;  // drp_render only accepts RGBA-padded textures (validated on the host side)
#else
    const bool rgba = (!m.base_color_tex.data || m.base_color_tex.c == 4) && (!m.mr_tex.data || m.mr_tex.c == 4) &&
                      (!use_nt || m.normal_tex.c == 4) && (!use_em || m.emissive_tex.c == 4);
#endif
    if (m.texel_records) {  // interleaved texels: one address stage, the taps of all four textures in the same sectors
        const TexTaps tp = tex_taps(m.base_color_tex.h, m.base_color_tex.w, m.base_color_tex.wrap, m.base_color_tex.interp, tu, tv, m.texel_tile_log2);
        float4 r0 = make_float4(1.0f, 1.0f, 1.0f, 1.0f), r1 = make_float4(0.0f, 0.0f, 0.0f, 0.0f), r2 = r1;
        if (use_bc) r0 = tex_fetch_record(m.texel_records, tp, 0);
        if (use_mr || use_nt) r1 = tex_fetch_record(m.texel_records, tp, 1);
        if (use_nt || use_em) r2 = tex_fetch_record(m.texel_records, tp, 2);
        if (use_bc) { bc[0] = r0.x; bc[1] = r0.y; bc[2] = r0.z; bc[3] = r0.w; }
        if (use_mr) { mr[1] = r1.x; mr[2] = r1.y; }
        if (use_nt) { nt[0] = r1.z; nt[1] = r1.w; nt[2] = r2.x; }
        if (use_em) { em[0] = r2.y; em[1] = r2.z; em[2] = r2.w; }
    } else if (rgba) {  // device path: address stage for all textures first, then all loads back to back
        TexTaps ta, tb, tc, td;
        if (use_bc) ta = tex_taps(m.base_color_tex.h, m.base_color_tex.w, m.base_color_tex.wrap, m.base_color_tex.interp, tu, tv);
        if (use_mr) tb = tex_taps(m.mr_tex.h, m.mr_tex.w, m.mr_tex.wrap, m.mr_tex.interp, tu, tv);
        if (use_nt) tc = tex_taps(m.normal_tex.h, m.normal_tex.w, m.normal_tex.wrap, m.normal_tex.interp, tu, tv);
        if (use_em) td = tex_taps(m.emissive_tex.h, m.emissive_tex.w, m.emissive_tex.wrap, m.emissive_tex.interp, tu, tv);
        if (use_bc) { float4 r = tex_fetch4(m.base_color_tex.data, ta); bc[0] = r.x; bc[1] = r.y; bc[2] = r.z; bc[3] = r.w; }
        if (use_mr) { float4 r = tex_fetch4(m.mr_tex.data, tb); mr[1] = r.y; mr[2] = r.z; }
        if (use_nt) { float4 r = tex_fetch4(m.normal_tex.data, tc); nt[0] = r.x; nt[1] = r.y; nt[2] = r.z; }
        if (use_em) { float4 r = tex_fetch4(m.emissive_tex.data, td); em[0] = r.x; em[1] = r.y; em[2] = r.z; }
    }
#ifndef __CUDA_ARCH__
    else {
        if (use_bc) {
            tex_fetch(m.base_color_tex, tu, tv, bc);
            if (m.base_color_tex.c < 4) bc[3] = 1.0f;
        }
        if (use_mr) tex_fetch(m.mr_tex, tu, tv, mr);
        if (use_nt) tex_fetch(m.normal_tex, tu, tv, nt);
        if (use_em) tex_fetch(m.emissive_tex, tu, tv, em);
    }
#endif
    float a = m.base_color_factor[3] * col[3] * bc[3];
    s.albedo = v3(m.base_color_factor[0] * col[0] * bc[0], m.base_color_factor[1] * col[1] * bc[1], m.base_color_factor[2] * col[2] * bc[2]);
    s.metal = m.metallic_factor * mr[2];
    s.smooth = 1.0f + (-m.roughness_factor) * mr[1];
    if (m.alpha_mode == DRP_ALPHA_MASK) s.alpha = a > m.alpha_cutoff ? 1.0f : 0.0f;
    else if (m.alpha_mode == DRP_ALPHA_BLEND) s.alpha = a;
    if (m.has_emissive) s.emission = v3(m.emissive_factor[0] * em[0], m.emissive_factor[1] * em[1], m.emissive_factor[2] * em[2]);
    if (use_nt) {  // tangent-space normal map, mixin.py:118-123
        float nx = 2.0f * nt[0] - 1.0f, ny = 2.0f * nt[1] - 1.0f, nz = 2.0f * nt[2] - 1.0f;
        const float4 g0 = q0.tan, g1 = q1.tan, g2 = q2.tan;
        Vec3 vt = v3(lerp3(g0.x, g1.x, g2.x, u, v), lerp3(g0.y, g1.y, g2.y, u, v), lerp3(g0.z, g1.z, g2.z, u, v));
        float vs = lerp3(g0.w, g1.w, g2.w, u, v);
        Vec3 vb = cross(nu, vt) * vs;
        s.normal = normalize_ref(vt * nx + (vb * ny + nu * nz));
    }
    return s;
}

// ---- BRDF sampler -----------------------------------------------------------------------------------------------
DRP_HD Vec3 tangent_combine(float x, float y, float z, Vec3 n) {  // light_transport.py:35-43
    Vec3 up = n.y < 0.999f ? v3(0.0f, 1.0f, 0.0f) : v3(1.0f, 0.0f, 0.0f);
    Vec3 right = normalize_ref(cross(up, n));
    Vec3 up2 = cross(n, right);
    return right * x + up2 * z + n * y;
}
DRP_HD float g_schlick(float ndv, float rough) {
    float k = (rough * rough) * 0.5f;
    return ndv / (ndv * (1.0f - k) + k);
}

// `s` must be all-zero for a miss (zero-initialised g-buffer, path_tracing.py:261-263)
DRP_HD BounceOut brdf_sample(const SurfaceAttrs& s, float t, Vec3 o, Vec3 d, Vec3 env, const float u[6]) {
    BounceOut r;
    r.hit_pos = o + d * t;
    r.radiance = s.emission + env;
    float diel = 1.0f - s.metal;
    Vec3 dc = s.albedo * diel;
    float dm = fmaxf(fmaxf(dc.x, dc.y), dc.z);
    float p_diff = diel * dm / (0.04f + dm);
    float p_spec = 1.0f - p_diff;
    if (u[0] >= s.alpha) {  // transmit
        r.next_d = d;
        r.transfer = v3(1.0f, 1.0f, 1.0f);
    } else if (u[1] >= p_spec) {  // diffuse, cosine-weighted about n
        float z2 = u[2], xy = sqrtf(1.0f - z2);
        float sn, cs;
        sincos_2pi(u[3], &sn, &cs);
        r.next_d = tangent_combine(xy * cs, sqrtf(z2), xy * sn, s.normal);
        float den = fmaxf(p_diff, 0.0001f);
        r.transfer = v3(dc.x / den, dc.y / den, dc.z / den);
    } else {  // GGX specular
        float rough = fmaxf(1.0f - s.smooth, 1.0f / 512.0f);
        float a = rough * rough;
        float ct = sqrtf((1.0f - u[5]) / (1.0f + (a * a - 1.0f) * u[5]));
        float st = sqrtf(1.0f - ct * ct);
        float sn, cs;
        sincos_2pi(u[4], &sn, &cs);
        Vec3 h = tangent_combine(cs * st, ct, sn * st, s.normal);
        float hd = dot(h, d);
        r.next_d = d + h * (-2.0f * hd);  // reflect
        float vh = -hd;
        float ndv = fmaxf(-dot(s.normal, d), 0.0f), ndl = fmaxf(dot(s.normal, r.next_d), 0.0f);
        float G = g_schlick(ndl, rough) * g_schlick(ndv, rough);
        float w1 = 1.0f - vh, w2 = w1 * w1, w5 = w2 * w2 * w1;
        float geo = G * fmaxf(vh, 1e-6f) / (fmaxf(dot(s.normal, h), 1e-6f) * fmaxf(-dot(s.normal, d), 1e-6f));
        float den = fmaxf(p_spec * s.alpha, 0.0001f);
        float f0x = s.albedo.x * s.metal + 0.04f * diel, f0y = s.albedo.y * s.metal + 0.04f * diel, f0z = s.albedo.z * s.metal + 0.04f * diel;
        r.transfer = v3((fmaxf(s.smooth - f0x, 0.0f) * w5 + f0x) * geo / den, (fmaxf(s.smooth - f0y, 0.0f) * w5 + f0y) * geo / den,
                        (fmaxf(s.smooth - f0z, 0.0f) * w5 + f0z) * geo / den);
    }
    return r;
}

// does the ray (o, d), t in [0, inf), touch the (padded) scene box?  Used by the exact compaction rule: a ray that
// missed everything and whose continuation cannot reach the scene box contributes nothing to any output again.
DRP_HD bool ray_may_reach_box(Vec3 o, Vec3 d, const float* __restrict__ lo, const float* __restrict__ hi) {
    float tn = 0.0f, tf = 3.0e38f;
    const float oo[3] = {o.x, o.y, o.z}, dd[3] = {d.x, d.y, d.z};
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        if (dd[a] == 0.0f) {
            if (oo[a] < lo[a] || oo[a] > hi[a]) return false;
        } else {
            float t1 = (lo[a] - oo[a]) / dd[a], t2 = (hi[a] - oo[a]) / dd[a];
            tn = fmaxf(tn, fminf(t1, t2));
            tf = fminf(tf, fmaxf(t1, t2));
        }
    }
    return tn <= tf;
}
