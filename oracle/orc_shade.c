/*
 * orc_shade.c -- CPU oracle, part 2: ray generation, surface attributes, BRDF sampling, bounce loop.
 * TEST INFRASTRUCTURE (see orc.h).  Restates, function by function, the reference's default path:
 *   orc_raygen          mixin.py:31-39, coordinates.py:6-10, path_tracing.py:329-331
 *   orc_texture_sample  shader_ops.py:198-221 (F.grid_sample semantics), gltf_material.py:15-22
 *   orc_env_lookup      coordinates.py:60-71, path_tracing.py:238-248,267, lights.py:49-50
 *   orc_surface_attrs   path_tracing.py:158-187, geometry.py:94-110, interpolator.py:32-48,
 *                       base_material.py:183-257, default_material.py:21-24, gltf_material.py:48-67,
 *                       mixin.py:115-128
 *   orc_sampler_brdf    path_tracing.py:189-236, light_transport.py:35-43,72-83,179-196, shader_ops.py:313-325
 *   orc_render          path_tracing.py:325-347 (+ :266-279 sampler tail)
 */
#include "orc.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define ORC_TAU 6.283185307179586f
#define ORC_PI 3.141592653589793f

static inline float clampf(float x, float lo, float hi) { return fminf(hi, fmaxf(x, lo)); }
static inline float dot3(const float* a, const float* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
static inline void cross3(const float* a, const float* b, float* r) {
    r[0] = a[1] * b[2] - a[2] * b[1]; r[1] = a[2] * b[0] - a[0] * b[2]; r[2] = a[0] * b[1] - a[1] * b[0];
}
/* F.normalize(x, dim=-1): x / max(||x||, 1e-12) (shader_ops.py:353-360) */
static inline void normalize3(float* v) {
    float l = fmaxf(sqrtf(dot3(v, v)), 1e-12f);
    v[0] /= l; v[1] /= l; v[2] /= l;
}

/* ---- Philox4x32-10 (native RNG mode; the CUDA kernels use the same definition) -------------------------- */
static inline void philox_round(uint32_t c[4], const uint32_t k[2]) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c[0], p1 = (uint64_t)0xCD9E8D57u * c[2];
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k[0], n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k[1], n3 = (uint32_t)p0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
}
static void philox4x32_10(uint32_t c[4], uint32_t k0, uint32_t k1) {
    uint32_t k[2] = {k0, k1};
    for (int r = 0; r < 10; ++r) {
        philox_round(c, k);
        k[0] += 0x9E3779B9u; k[1] += 0xBB67AE85u;
    }
}
void orc_philox_uniform6(uint64_t seed, uint32_t pixel, uint32_t sample, uint32_t bounce, float out6[6]) {
    uint32_t a[4] = {pixel, sample, bounce, 0u}, b[4] = {pixel, sample, bounce, 1u};
    philox4x32_10(a, (uint32_t)seed, (uint32_t)(seed >> 32));
    philox4x32_10(b, (uint32_t)seed, (uint32_t)(seed >> 32));
    const float s = 5.9604644775390625e-08f; /* 2^-24 */
    out6[0] = (float)(a[0] >> 8) * s; out6[1] = (float)(a[1] >> 8) * s; out6[2] = (float)(a[2] >> 8) * s;
    out6[3] = (float)(a[3] >> 8) * s; out6[4] = (float)(b[0] >> 8) * s; out6[5] = (float)(b[1] >> 8) * s;
}

/* ---- textures: F.grid_sample(align_corners=False) -------------------------------------------------------- */
static inline float reflect_coord(float in, int twice_low, int twice_high) {
    if (twice_low == twice_high) return 0.0f;
    float mn = (float)twice_low / 2.0f, span = (float)(twice_high - twice_low) / 2.0f;
    in = fabsf(in - mn);
    float extra = fmodf(in, span);
    int flips = (int)floorf(in / span);
    return (flips % 2 == 0) ? extra + mn : span - extra + mn;
}
static inline float grid_coord(float g, int size, int reflection) {
    float c = ((g + 1.0f) * (float)size - 1.0f) / 2.0f;
    if (reflection) c = reflect_coord(c, -1, 2 * size - 1);
    return clampf(c, 0.0f, (float)(size - 1));
}
static inline float texel(const drp_texture_t* t, int x, int y, int ch) {
    if (x < 0 || y < 0 || x >= t->w || y >= t->h) return 0.0f;
    return t->data[((int64_t)y * t->w + x) * t->c + ch];
}
/* one lookup; out has tex->c channels */
static void tex_fetch(const drp_texture_t* tex, float u, float v, float* out) {
    if (tex->wrap == DRP_WRAP_REPEAT) { /* uv.remainder(1.0) */
        u = u - floorf(u); v = v - floorf(v);
        if (u >= 1.0f) u = 0.0f; if (v >= 1.0f) v = 0.0f; /* remainder() is in [0,1) */
    }
    int reflection = tex->wrap != DRP_WRAP_CLAMP;
    float gx = u * 2.0f - 1.0f, gy = -(v * 2.0f - 1.0f); /* flipper_2d = (1,-1) */
    float ix = grid_coord(gx, tex->w, reflection), iy = grid_coord(gy, tex->h, reflection);
    if (tex->interp == DRP_INTERP_POINT) {
        int x = (int)nearbyintf(ix), y = (int)nearbyintf(iy);
        for (int c = 0; c < tex->c; ++c) out[c] = texel(tex, x, y, c);
        return;
    }
    float x0 = floorf(ix), y0 = floorf(iy), x1 = x0 + 1.0f, y1 = y0 + 1.0f;
    float wnw = (x1 - ix) * (y1 - iy), wne = (ix - x0) * (y1 - iy), wsw = (x1 - ix) * (iy - y0), wse = (ix - x0) * (iy - y0);
    int X0 = (int)x0, Y0 = (int)y0;
    for (int c = 0; c < tex->c; ++c)
        out[c] = texel(tex, X0, Y0, c) * wnw + texel(tex, X0 + 1, Y0, c) * wne + texel(tex, X0, Y0 + 1, c) * wsw +
                 texel(tex, X0 + 1, Y0 + 1, c) * wse;
}
void orc_texture_sample(const drp_texture_t* tex, const float* uv, int64_t n, float* out) {
#pragma omp parallel for
    for (int64_t k = 0; k < n; ++k) tex_fetch(tex, uv[2 * k], uv[2 * k + 1], out + (int64_t)tex->c * k);
}

/* unit direction -> lat-long uv, coordinates.py:60-71 */
static inline void latlong_uv(const float* d, float* u, float* v) {
    float a = atan2f(d[0], d[2]) * (0.5f / ORC_PI);
    a = a - floorf(a); /* python % 1 */
    if (a >= 1.0f) a = 0.0f;
    *u = a;
    *v = (1.0f / ORC_PI) * asinf(clampf(d[1], -0.999999f, 0.999999f)) + 0.5f;
}
static void env_fetch(const drp_texture_t* env, const float* d, float* rgb) {
    rgb[0] = rgb[1] = rgb[2] = 0.0f;
    if (!env->data) return; /* black_tex, path_tracing.py:247 */
    float u, v, tmp[4] = {0, 0, 0, 0};
    latlong_uv(d, &u, &v);
    drp_texture_t e = *env;
    e.wrap = DRP_WRAP_CLAMP; e.interp = DRP_INTERP_LINEAR; /* sample2d defaults: 'border', 'bilinear' */
    tex_fetch(&e, u, v, tmp);
    rgb[0] = tmp[0]; rgb[1] = tmp[1]; rgb[2] = tmp[2];
}
void orc_env_lookup(const drp_texture_t* env, const float* rays_d, int64_t n, float* out_rgb) {
#pragma omp parallel for
    for (int64_t k = 0; k < n; ++k) env_fetch(env, rays_d + 3 * k, out_rgb + 3 * k);
}

/* ---- primary rays -------------------------------------------------------------------------------------- */
void orc_raygen(const drp_render_params_t* p, float jx, float jy, const float* ndc_x, const float* ndc_y, float* o,
                float* d) {
    /* whole frame, or the tile [x0, x0+w) x [y0, y0+h) when tile sharding (rays stored tile-locally) */
    const int tiled = p->tile_w > 0 && p->tile_h > 0;
    const int H = tiled ? p->tile_h : p->height, W = tiled ? p->tile_w : p->width;
    const int x0 = tiled ? p->tile_x0 : 0, y0 = tiled ? p->tile_y0 : 0;
    const float* m = p->inv_vp;
#pragma omp parallel for
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) {
            float g[4] = {ndc_x[x0 + x] + jx, ndc_y[y0 + y] + jy, -1.0f, 1.0f};
            float q[4];
            for (int k = 0; k < 4; ++k) q[k] = g[0] * m[4 * k] + g[1] * m[4 * k + 1] + g[2] * m[4 * k + 2] + g[3] * m[4 * k + 3];
            float dir[3] = {q[0] / q[3] - p->cam_pos[0], q[1] / q[3] - p->cam_pos[1], q[2] / q[3] - p->cam_pos[2]};
            normalize3(dir);
            int64_t r = (int64_t)y * W + x;
            for (int k = 0; k < 3; ++k) { d[3 * r + k] = dir[k]; o[3 * r + k] = p->cam_pos[k] + dir[k] * p->t_near; }
        }
}

/* ---- surface attributes ---------------------------------------------------------------------------------- */
/* geometry.py:94-110: barycentric weights of the *hit point*, nan->0, clipped; returns (u, v) = weights of vertex 0, 1 */
static inline void bary_uv(const float* a, const float* b, const float* c, const float* p, float* u, float* v) {
    float v0[3], v1[3], v2[3];
    for (int k = 0; k < 3; ++k) { v0[k] = b[k] - a[k]; v1[k] = c[k] - a[k]; v2[k] = p[k] - a[k]; }
    float d00 = dot3(v0, v0), d01 = dot3(v0, v1), d11 = dot3(v1, v1), d20 = dot3(v2, v0), d21 = dot3(v2, v1);
    float denom = d00 * d11 - d01 * d01;
    float bv = (d11 * d20 - d01 * d21) / denom, bw = (d00 * d21 - d01 * d20) / denom;
    if (bv != bv) bv = 0.0f; if (bw != bw) bw = 0.0f; /* nan_to_num (inf -> clip below) */
    bv = clampf(bv, 0.0f, 1.0f); bw = clampf(bw, 0.0f, 1.0f);
    *u = 1.0f - bv - bw;
    *v = bv;
}
/* interpolator.py:32-48: (v1 - v3) * u + ((v2 - v3) * v + v3) */
static inline void interp(const float* buf, int C, const int32_t* tri, float u, float v, float* out) {
    const float* a = buf + (int64_t)C * tri[0];
    const float* b = buf + (int64_t)C * tri[1];
    const float* c = buf + (int64_t)C * tri[2];
    for (int k = 0; k < C; ++k) out[k] = (a[k] - c[k]) * u + ((b[k] - c[k]) * v + c[k]);
}

static void surface_one(const drp_scene_t* sc, const float* o, const float* d, float t, int32_t id, float* at) {
    const int32_t* tri = sc->tris + 3 * (int64_t)id;
    float P[3] = {o[0] + d[0] * t, o[1] + d[1] * t, o[2] + d[2] * t};
    float u, v;
    bary_uv(sc->world_pos + 3 * (int64_t)tri[0], sc->world_pos + 3 * (int64_t)tri[1], sc->world_pos + 3 * (int64_t)tri[2], P, &u, &v);
    const drp_material_t* m = sc->materials + sc->tri_material[id];
    float nu[3], col[4];
    interp(sc->world_nrm, 3, tri, u, v, nu); /* world_normal_unnormalized */
    interp(sc->color, 4, tri, u, v, col);
    float N[3] = {nu[0], nu[1], nu[2]};
    normalize3(N);
    float albedo[3], metal = 0.0f, smooth = 0.5f, alpha = 1.0f, emis[3] = {0, 0, 0}; /* path_tracing.py:183-186 */
    if (m->kind == DRP_MAT_DEFAULT) {
        for (int k = 0; k < 3; ++k) albedo[k] = col[k] * m->tint[k];
    } else {
        float uv[2], bc[4] = {1, 1, 1, 1}, mr[4] = {0, 0, 0, 0};
        interp(sc->uv, 2, tri, u, v, uv);
        if (m->base_color_tex.data) tex_fetch(&m->base_color_tex, uv[0], uv[1], bc);
        if (m->base_color_tex.data && m->base_color_tex.c < 4) bc[3] = 1.0f;
        if (m->mr_tex.data) tex_fetch(&m->mr_tex, uv[0], uv[1], mr);
        float rgba[4];
        for (int k = 0; k < 4; ++k) rgba[k] = m->base_color_factor[k] * col[k] * bc[k];
        for (int k = 0; k < 3; ++k) albedo[k] = rgba[k];
        metal = m->metallic_factor * mr[2];
        smooth = 1.0f + (-m->roughness_factor) * mr[1];
        if (m->alpha_mode == DRP_ALPHA_MASK) alpha = rgba[3] > m->alpha_cutoff ? 1.0f : 0.0f;
        else if (m->alpha_mode == DRP_ALPHA_BLEND) alpha = rgba[3];
        if (m->has_emissive) {
            float e[4] = {0, 0, 0, 0};
            if (m->emissive_tex.data) tex_fetch(&m->emissive_tex, uv[0], uv[1], e);
            for (int k = 0; k < 3; ++k) emis[k] = m->emissive_factor[k] * e[k];
        }
        if (m->has_normal_tex && m->normal_tex.data) { /* mixin.py:118-123, tangent space */
            float nt[4] = {0, 0, 0, 0}, tg[4], vb[3];
            tex_fetch(&m->normal_tex, uv[0], uv[1], nt);
            for (int k = 0; k < 3; ++k) nt[k] = 2.0f * nt[k] - 1.0f;
            interp(sc->world_tan, 4, tri, u, v, tg);
            cross3(nu, tg, vb);
            for (int k = 0; k < 3; ++k) vb[k] *= tg[3];
            for (int k = 0; k < 3; ++k) N[k] = nt[0] * tg[k] + (nt[1] * vb[k] + nt[2] * nu[k]);
            normalize3(N);
        }
    }
    at[0] = albedo[0]; at[1] = albedo[1]; at[2] = albedo[2];
    at[3] = N[0]; at[4] = N[1]; at[5] = N[2];
    at[6] = metal; at[7] = smooth; at[8] = alpha;
    at[9] = emis[0]; at[10] = emis[1]; at[11] = emis[2];
}

void orc_surface_attrs(const drp_scene_t* scene, const float* rays_o, const float* rays_d, const float* t,
                       const int32_t* tri, float t_far, int64_t n, float* attrs) {
#pragma omp parallel for schedule(dynamic, 1024)
    for (int64_t r = 0; r < n; ++r) {
        float* at = attrs + 12 * r;
        if (t[r] < t_far) surface_one(scene, rays_o + 3 * r, rays_d + 3 * r, t[r], tri[r], at);
        else memset(at, 0, 48); /* zeros_like_vec(rays_o, 12), path_tracing.py:261-263 */
    }
}

/* ---- BRDF sampler --------------------------------------------------------------------------------------- */
/* light_transport.py:35-43 */
static inline void tangent_combine(float x, float y, float z, const float* n, float* out) {
    float up[3] = {0, 0, 0}, right[3], up2[3];
    if (n[1] < 0.999f) up[1] = 1.0f; else up[0] = 1.0f;
    cross3(up, n, right);
    normalize3(right);
    cross3(n, right, up2);
    for (int k = 0; k < 3; ++k) out[k] = x * right[k] + z * up2[k] + y * n[k];
}
static inline float g_schlick(float ndv, float rough) { /* light_transport.py:179-184 */
    float k = (rough * rough) / 2.0f;
    return ndv / (ndv * (1.0f - k) + k);
}

static void brdf_one(const float* at, float t, const float* o, const float* d, const float* env, const float* u,
                     float* radiance, float* transfer, float* next_o, float* next_d) {
    const float* albedo = at;
    const float* n = at + 3;
    float metal = at[6], smooth = at[7], alpha = at[8];
    const float* emission = at + 9;
    for (int k = 0; k < 3; ++k) next_o[k] = o[k] + d[k] * t;
    float diel = 1.0f - metal;
    float dc[3] = {diel * albedo[0], diel * albedo[1], diel * albedo[2]};
    float dm = fmaxf(fmaxf(dc[0], dc[1]), dc[2]);
    float p_diff = diel * dm / (0.04f + dm);
    float p_spec = 1.0f - p_diff;
    int is_transmit = u[0] >= alpha;
    int is_diffuse = u[1] >= p_spec;
    for (int k = 0; k < 3; ++k) radiance[k] = emission[k] + env[k];
    if (is_transmit) {
        for (int k = 0; k < 3; ++k) { next_d[k] = d[k]; transfer[k] = 1.0f; }
    } else if (is_diffuse) {
        float z2 = u[2], theta = u[3] * ORC_TAU, xy = sqrtf(1.0f - z2);
        tangent_combine(xy * cosf(theta), sqrtf(z2), xy * sinf(theta), n, next_d);
        float den = fmaxf(p_diff, 0.0001f);
        for (int k = 0; k < 3; ++k) transfer[k] = (albedo[k] * diel) / den;
    } else {
        float rough = fmaxf(1.0f - smooth, 1.0f / 512.0f);
        float a = rough * rough; /* light_transport.py:72-83 */
        float phi = ORC_TAU * u[4];
        float ct = sqrtf((1.0f - u[5]) / (1.0f + (a * a - 1.0f) * u[5]));
        float st = sqrtf(1.0f - ct * ct);
        float h[3];
        tangent_combine(cosf(phi) * st, ct, sinf(phi) * st, n, h);
        float hd = dot3(h, d);
        for (int k = 0; k < 3; ++k) next_d[k] = d[k] + (-2.0f) * hd * h[k]; /* reflect, shader_ops.py:313-325 */
        float vh = -hd;
        float md[3] = {-d[0], -d[1], -d[2]};
        float ndv = fmaxf(dot3(n, md), 0.0f), ndl = fmaxf(dot3(n, next_d), 0.0f);
        float G = g_schlick(ndl, rough) * g_schlick(ndv, rough);
        float w = powf(1.0f - vh, 5.0f);
        float geo = G * fmaxf(vh, 1e-6f) / (fmaxf(dot3(n, h), 1e-6f) * fmaxf(-dot3(n, d), 1e-6f));
        float den = fmaxf(p_spec * alpha, 0.0001f);
        for (int k = 0; k < 3; ++k) {
            float f0 = albedo[k] * metal + 0.04f * diel;
            float F = fmaxf(smooth - f0, 0.0f) * w + f0; /* light_transport.py:195-196 */
            transfer[k] = F * geo / den;
        }
    }
}

void orc_sampler_brdf(const float* attrs, const float* t, const float* rays_o, const float* rays_d,
                      const float* env_radiance, const float* u6, int64_t n, float* radiance, float* transfer,
                      float* next_o, float* next_d) {
#pragma omp parallel for schedule(static)
    for (int64_t r = 0; r < n; ++r) {
        float u[6];
        for (int k = 0; k < 6; ++k) u[k] = u6[(int64_t)k * n + r];
        brdf_one(attrs + 12 * r, t[r], rays_o + 3 * r, rays_d + 3 * r, env_radiance + 3 * r, u, radiance + 3 * r,
                 transfer + 3 * r, next_o + 3 * r, next_d + 3 * r);
    }
}

/* ---- bounce loop ------------------------------------------------------------------------------------------ */
int64_t orc_render(const orc_bvh_t* bvh, const drp_scene_t* scene, const drp_render_params_t* p, float* accum) {
    const int tiled = p->tile_w > 0 && p->tile_h > 0;
    const int TW = tiled ? p->tile_w : p->width, TX0 = tiled ? p->tile_x0 : 0, TY0 = tiled ? p->tile_y0 : 0;
    const int64_t HW = tiled ? (int64_t)p->tile_w * p->tile_h : (int64_t)p->height * p->width;
    const int64_t R_total = HW * p->n_samples;
    float* o = (float*)malloc(sizeof(float) * 3 * HW);
    float* d = (float*)malloc(sizeof(float) * 3 * HW);
    float* T = (float*)malloc(sizeof(float) * 3 * HW);
    float* t = (float*)malloc(sizeof(float) * HW);
    int32_t* id = (int32_t*)malloc(sizeof(int32_t) * HW);
    float* attrs = (float*)malloc(sizeof(float) * 12 * HW);
    int64_t traced = 0;
    for (int s = 0; s < p->n_samples; ++s) {
        orc_raygen(p, p->jitter_x[s], p->jitter_y[s], p->ndc_x, p->ndc_y, o, d);
        for (int64_t k = 0; k < 3 * HW; ++k) T[k] = 1.0f; /* path_tracing.py:332 */
        for (int b = 0; b < p->ray_depth; ++b) {
            orc_bvh_query(bvh, o, d, HW, p->t_far, 0.0f, ORC_TRI_MT, 0, t, id);
            traced += HW;
            orc_surface_attrs(scene, o, d, t, id, p->t_far, HW, attrs);
            const int always_sky = (b == p->ray_depth - 1) && p->last_bounce_skybox; /* path_tracing.py:260 */
#pragma omp parallel for schedule(static)
            for (int64_t px = 0; px < HW; ++px) {
                const int hit = t[px] < p->t_far;
                const int64_t gpx = (int64_t)(TY0 + px / TW) * p->width + TX0 + px % TW; /* global pixel (accumulator row, RNG key) */
                float env[3], u[6], rad[3], tr[3], no[3], nd[3];
                env_fetch(&scene->env, d + 3 * px, env);
                if (!always_sky && hit) env[0] = env[1] = env[2] = 0.0f; /* path_tracing.py:268-269 */
                if (p->rng_mode == DRP_RNG_REPLAY) {
                    for (int k = 0; k < 6; ++k) u[k] = p->replay_u[((int64_t)b * 6 + k) * R_total + (int64_t)s * HW + px];
                } else {
                    orc_philox_uniform6(p->seed, (uint32_t)gpx, (uint32_t)p->sample_ids[s], (uint32_t)b, u);
                }
                const float* at = attrs + 12 * px;
                brdf_one(at, t[px], o + 3 * px, d + 3 * px, env, u, rad, tr, no, nd);
                float* acc = accum + DRP_ACCUM_CHANNELS * gpx;
                for (int k = 0; k < 3; ++k) acc[k] += T[3 * px + k] * rad[k]; /* path_tracing.py:336 */
                acc[3] += at[8];                                             /* :337 */
                if (b == 0) {                                                /* :340-347, extras at :278 */
                    for (int k = 0; k < 3; ++k) {
                        acc[4 + k] += at[k]; acc[7 + k] += at[9 + k]; acc[10 + k] += at[3 + k]; acc[13 + k] += no[k];
                    }
                }
                for (int k = 0; k < 3; ++k) {
                    T[3 * px + k] *= hit ? tr[k] : 0.0f;                      /* :338 with :275 */
                    o[3 * px + k] = no[k] + nd[k] * p->step_epsilon;        /* :276 */
                    d[3 * px + k] = nd[k];
                }
            }
        }
    }
    free(o); free(d); free(T); free(t); free(id); free(attrs);
    return traced;
}
