/*
 * orc_raycast.c -- CPU oracle, part 1: closest-hit ray/triangle queries.
 * TEST INFRASTRUCTURE (see orc.h).  Restates diffrp/utils/raycaster.py of the reference:
 *   - orc_tri_mt           : _ray_tri_intersect                      raycaster.py:58-79
 *   - orc_tri_unit(+xform) : _ray_tri_pretransform / _pretransformed raycaster.py:27-55
 *   - orc_bruteforce       : BruteForceRaycaster.query               raycaster.py:86-97
 *   - orc_bvh_build        : NaivePBBVH.build                        raycaster.py:122-187
 *   - orc_bvh_query        : NaivePBBVH.query + helpers              raycaster.py:189-260
 * Build with -ffp-contract=off: every fused multiply-add below is explicit (fmaf) and mirrors what the
 * reference's ATen CPU kernels do (probe: torch.linalg.cross contracts a*b-c*d into fma(a,b,-(c*d));
 * torch.linalg.vecdot sums left-to-right without contraction).
 */
#include "orc.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <omp.h>

int orc_num_threads(void) { return omp_get_max_threads(); }
void orc_set_num_threads(int n) { if (n > 0) omp_set_num_threads(n); }

/* ---- small vector helpers (explicit rounding points) --------------------------------------------- */
static inline void v_sub(const float* a, const float* b, float* r) { r[0] = a[0] - b[0]; r[1] = a[1] - b[1]; r[2] = a[2] - b[2]; }
static inline void v_cross(const float* a, const float* b, float* r) {
    r[0] = fmaf(a[1], b[2], -(a[2] * b[1]));
    r[1] = fmaf(a[2], b[0], -(a[0] * b[2]));
    r[2] = fmaf(a[0], b[1], -(a[1] * b[0]));
}
static inline float v_dot(const float* a, const float* b) { return (a[0] * b[0] + a[1] * b[1]) + a[2] * b[2]; }

/* raycaster.py:58-79.  The reference unbinds (v1, v2, v0) = triangle vertices (0, 1, 2). */
static inline float orc_tri_mt(const float* o, const float* d, const float* A, const float* B, const float* C,
                               float t_far, float eps) {
    float e1[3], e2[3], cr[3], s[3], sc[3];
    v_sub(A, C, e1);
    v_sub(B, C, e2);
    v_cross(d, e2, cr);
    float det = v_dot(e1, cr);
    float inv_det = 1.0f / det;
    v_sub(o, C, s);
    float u = inv_det * v_dot(s, cr);
    v_cross(s, e1, sc);
    float v = inv_det * v_dot(d, sc);
    float t = inv_det * v_dot(e2, sc);
    int hit = (fabsf(det) > eps) & (u >= 0.0f) & (v >= 0.0f) & (u + v <= 1.0f) & (t > 0.0f);
    return hit ? t : t_far;
}

/* raycaster.py:27-39: inverse of the affine frame [B-A, C-A, normalize((B-A)x(C-A)), A].  The reference
 * inverts with LAPACK (torch.linalg.inv_ex); this closed form agrees up to rounding. g = 9 rot + 3 trans. */
static void orc_unit_xform(const float* A, const float* B, const float* C, float* g) {
    float e1[3], e2[3], n[3], r0[3], r1[3], r2[3];
    v_sub(B, A, e1);
    v_sub(C, A, e2);
    v_cross(e1, e2, n);
    float len = sqrtf(v_dot(n, n));
    float il = 1.0f / fmaxf(len, 1e-12f);
    n[0] *= il; n[1] *= il; n[2] *= il;
    v_cross(e2, n, r0);
    v_cross(n, e1, r1);
    v_cross(e1, e2, r2);
    float det = v_dot(e1, r0);
    float id = 1.0f / det;
    for (int k = 0; k < 3; ++k) { g[k] = r0[k] * id; g[3 + k] = r1[k] * id; g[6 + k] = r2[k] * id; }
    g[9] = -v_dot(g, A); g[10] = -v_dot(g + 3, A); g[11] = -v_dot(g + 6, A);
}

/* raycaster.py:42-55 */
static inline float orc_tri_unit(const float* o, const float* d, const float* g, float t_far) {
    float ox = v_dot(g, o) + g[9], oy = v_dot(g + 3, o) + g[10], oz = v_dot(g + 6, o) + g[11];
    float dx = v_dot(g, d), dy = v_dot(g + 3, d), dz = v_dot(g + 6, d);
    float t = -oz / dz;
    float b1 = ox + t * dx;
    float b2 = oy + t * dy;
    int hit = (t > 0.0f) & (b1 >= 0.0f) & (b2 >= 0.0f) & (b1 + b2 <= 1.0f);
    return hit ? t : t_far;
}

/* ---- brute force: raycaster.py:86-97 (argmin = first index of the minimum; 0 when every t == far) -------- */
void orc_bruteforce(const float* verts, const int32_t* tris, int64_t n_tris, const float* rays_o, const float* rays_d,
                    int64_t n_rays, float t_far, float eps, int tri_test, float* out_t, int32_t* out_i) {
    float* g = NULL;
    if (tri_test == ORC_TRI_UNIT) {
        g = (float*)malloc(sizeof(float) * 12 * (size_t)(n_tris > 0 ? n_tris : 1));
        for (int64_t k = 0; k < n_tris; ++k)
            orc_unit_xform(verts + 3 * (int64_t)tris[3 * k], verts + 3 * (int64_t)tris[3 * k + 1],
                           verts + 3 * (int64_t)tris[3 * k + 2], g + 12 * k);
    }
#pragma omp parallel for schedule(dynamic, 64)
    for (int64_t r = 0; r < n_rays; ++r) {
        const float* o = rays_o + 3 * r;
        const float* d = rays_d + 3 * r;
        float best = t_far;
        int32_t bi = 0;
        for (int64_t k = 0; k < n_tris; ++k) {
            float t;
            if (tri_test == ORC_TRI_UNIT) t = orc_tri_unit(o, d, g + 12 * k, t_far);
            else t = orc_tri_mt(o, d, verts + 3 * (int64_t)tris[3 * k], verts + 3 * (int64_t)tris[3 * k + 1],
                                verts + 3 * (int64_t)tris[3 * k + 2], t_far, eps);
            if (t < best) { best = t; bi = (int32_t)k; }
        }
        out_t[r] = best;
        out_i[r] = bi;
    }
    free(g);
}

/* ---- NaivePBBVH: implicit complete binary heap over 2^n (padded) triangles ------------------------------ */
struct orc_bvh {
    int64_t M;     /* real triangle count                         */
    int64_t P;     /* padded count = 2^n                          */
    int n;
    int32_t* rank; /* (P) leaf slot -> original triangle id (already % M), raycaster.py:183 */
    float* bmin;   /* (2P-1, 3) heap order, root first, raycaster.py:169-178 */
    float* bmax;
    float* tri;    /* (P, 9) vertex positions in leaf order      */
    float* g2b;    /* (P, 12) unit-triangle transforms, raycaster.py:187 */
};

typedef struct { float key; int32_t idx; } orc_kv_t;
static inline int kv_less(orc_kv_t a, orc_kv_t b) { return a.key < b.key || (a.key == b.key && a.idx < b.idx); }

/* quickselect so that kv[0..k) are the k smallest: equivalent to "argsort then split in halves" (raycaster.py:148-153) */
static void kv_select(orc_kv_t* kv, int64_t n, int64_t k) {
    int64_t lo = 0, hi = n - 1;
    while (lo < hi) {
        orc_kv_t p = kv[lo + (hi - lo) / 2];
        int64_t i = lo, j = hi;
        while (i <= j) {
            while (kv_less(kv[i], p)) ++i;
            while (kv_less(p, kv[j])) --j;
            if (i <= j) { orc_kv_t t = kv[i]; kv[i] = kv[j]; kv[j] = t; ++i; --j; }
        }
        if (k <= j) hi = j; else if (k >= i) lo = i; else break;
    }
}

static void splitaxis_rec(int32_t* idx, orc_kv_t* tmp, int64_t m, const float* tmin, const float* tmax,
                          const float* cen) {
    if (m <= 1) return;
    float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (int64_t j = 0; j < m; ++j)
        for (int a = 0; a < 3; ++a) {
            lo[a] = fminf(lo[a], tmin[3 * (int64_t)idx[j] + a]);
            hi[a] = fmaxf(hi[a], tmax[3 * (int64_t)idx[j] + a]);
        }
    int axis = 0; /* torch.argmax: first maximum */
    float ext = hi[0] - lo[0];
    for (int a = 1; a < 3; ++a) if (hi[a] - lo[a] > ext) { ext = hi[a] - lo[a]; axis = a; }
    for (int64_t j = 0; j < m; ++j) { tmp[j].key = cen[3 * (int64_t)idx[j] + axis]; tmp[j].idx = idx[j]; }
    kv_select(tmp, m, m / 2);
    for (int64_t j = 0; j < m; ++j) idx[j] = tmp[j].idx;
    if (m >= 8192) {
#pragma omp task
        splitaxis_rec(idx, tmp, m / 2, tmin, tmax, cen);
#pragma omp task
        splitaxis_rec(idx + m / 2, tmp + m / 2, m / 2, tmin, tmax, cen);
#pragma omp taskwait
    } else {
        splitaxis_rec(idx, tmp, m / 2, tmin, tmax, cen);
        splitaxis_rec(idx + m / 2, tmp + m / 2, m / 2, tmin, tmax, cen);
    }
}

static inline uint32_t expand_bits(uint32_t v) { /* raycaster.py:100-107 */
    v = (v * 0x00010001u) & 0xFF0000FFu;
    v = (v * 0x00000101u) & 0x0F00F00Fu;
    v = (v * 0x00000011u) & 0xC30C30C3u;
    v = (v * 0x00000005u) & 0x49249249u;
    return v;
}

static int cmp_u64(const void* a, const void* b) {
    uint64_t x = *(const uint64_t*)a, y = *(const uint64_t*)b;
    return (x > y) - (x < y);
}

orc_bvh_t* orc_bvh_build(const float* verts, const int32_t* tris, int64_t n_tris, int builder) {
    if (n_tris <= 0) return NULL;
    orc_bvh_t* b = (orc_bvh_t*)calloc(1, sizeof(orc_bvh_t));
    int n = 0;
    while (((int64_t)1 << n) < n_tris) ++n; /* (M-1).bit_length() */
    int64_t P = (int64_t)1 << n;
    b->M = n_tris; b->n = n; b->P = P;
    float* T = (float*)malloc(sizeof(float) * 9 * P); /* padded triangles, raycaster.py:131-133 */
    for (int64_t k = 0; k < P; ++k) {
        int64_t src = k < n_tris ? k : k - n_tris;
        for (int c = 0; c < 3; ++c) memcpy(T + 9 * k + 3 * c, verts + 3 * (int64_t)tris[3 * src + c], 12);
    }
    float* tmin = (float*)malloc(sizeof(float) * 3 * P);
    float* tmax = (float*)malloc(sizeof(float) * 3 * P);
    float* cen = (float*)malloc(sizeof(float) * 3 * P);
    for (int64_t k = 0; k < P; ++k)
        for (int a = 0; a < 3; ++a) {
            float x = T[9 * k + a], y = T[9 * k + 3 + a], z = T[9 * k + 6 + a];
            tmin[3 * k + a] = fminf(fminf(x, y), z);
            tmax[3 * k + a] = fmaxf(fmaxf(x, y), z);
            cen[3 * k + a] = ((x + y) + z) / 3.0f; /* triangles.mean(-2) */
        }
    int32_t* idx = (int32_t*)malloc(sizeof(int32_t) * P);
    for (int64_t k = 0; k < P; ++k) idx[k] = (int32_t)k;
    if (builder == ORC_BUILD_SPLITAXIS) {
        orc_kv_t* tmp = (orc_kv_t*)malloc(sizeof(orc_kv_t) * P);
#pragma omp parallel
#pragma omp single
        splitaxis_rec(idx, tmp, P, tmin, tmax, cen);
        free(tmp);
    } else { /* zorder_3d, raycaster.py:110-117 */
        float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
        for (int64_t k = 0; k < P; ++k) for (int a = 0; a < 3; ++a) lo[a] = fminf(lo[a], cen[3 * k + a]);
        for (int64_t k = 0; k < P; ++k) for (int a = 0; a < 3; ++a) hi[a] = fmaxf(hi[a], cen[3 * k + a] - lo[a]);
        uint64_t* keys = (uint64_t*)malloc(sizeof(uint64_t) * P);
        for (int64_t k = 0; k < P; ++k) {
            uint32_t q[3];
            for (int a = 0; a < 3; ++a) {
                float x = (cen[3 * k + a] - lo[a]) / (hi[a] + 1e-8f);
                q[a] = expand_bits((uint32_t)(int32_t)(x * 1023.0f));
            }
            uint32_t code = q[0] | (q[1] << 1) | (q[2] << 2);
            keys[k] = ((uint64_t)code << 32) | (uint32_t)k;
        }
        qsort(keys, (size_t)P, sizeof(uint64_t), cmp_u64);
        for (int64_t k = 0; k < P; ++k) idx[k] = (int32_t)(keys[k] & 0xffffffffu);
        free(keys);
    }
    b->rank = (int32_t*)malloc(sizeof(int32_t) * P);
    b->tri = (float*)malloc(sizeof(float) * 9 * P);
    b->g2b = (float*)malloc(sizeof(float) * 12 * P);
    b->bmin = (float*)malloc(sizeof(float) * 3 * (2 * P - 1));
    b->bmax = (float*)malloc(sizeof(float) * 3 * (2 * P - 1));
    for (int64_t j = 0; j < P; ++j) {
        int64_t k = idx[j];
        b->rank[j] = (int32_t)(k % n_tris);
        memcpy(b->tri + 9 * j, T + 9 * k, 36);
        orc_unit_xform(T + 9 * k, T + 9 * k + 3, T + 9 * k + 6, b->g2b + 12 * j);
        memcpy(b->bmin + 3 * (P - 1 + j), tmin + 3 * k, 12);
        memcpy(b->bmax + 3 * (P - 1 + j), tmax + 3 * k, 12);
    }
    for (int64_t k = P - 2; k >= 0; --k)
        for (int a = 0; a < 3; ++a) {
            b->bmin[3 * k + a] = fminf(b->bmin[3 * (2 * k + 1) + a], b->bmin[3 * (2 * k + 2) + a]);
            b->bmax[3 * k + a] = fmaxf(b->bmax[3 * (2 * k + 1) + a], b->bmax[3 * (2 * k + 2) + a]);
        }
    free(T); free(tmin); free(tmax); free(cen); free(idx);
    return b;
}

void orc_bvh_free(orc_bvh_t* b) {
    if (!b) return;
    free(b->rank); free(b->bmin); free(b->bmax); free(b->tri); free(b->g2b); free(b);
}

/* torch.min / torch.max propagate NaN (reference defect B9 depends on it) */
static inline float nan_min(float a, float b) { return (a != a || b != b) ? NAN : (a < b ? a : b); }
static inline float nan_max(float a, float b) { return (a != a || b != b) ? NAN : (a > b ? a : b); }

/* raycaster.py:198-204 verbatim */
static inline int box_test_reference(const float* lo, const float* hi, const float* o, const float* d, float t_live) {
    float tmn = -INFINITY, tmx = INFINITY;
    for (int a = 0; a < 3; ++a) {
        float t1 = (lo[a] - o[a]) / d[a], t2 = (hi[a] - o[a]) / d[a];
        float mn = nan_min(t1, t2), mx = nan_max(t1, t2);
        tmn = a == 0 ? mn : nan_max(tmn, mn);
        tmx = a == 0 ? mx : nan_min(tmx, mx);
    }
    return (tmn <= t_live) & (tmx > 0.0f) & (tmn <= tmx);
}

/* conservative slab test: never rejects a box that holds a triangle the fp32 triangle test accepts */
static inline int box_test_conservative(const float* lo, const float* hi, const float* o, const float* d, float t_best) {
    float tmn = 0.0f, tmx = t_best;
    for (int a = 0; a < 3; ++a) {
        float pad = 1e-6f * fmaxf(fmaxf(fabsf(lo[a]), fabsf(hi[a])), fabsf(o[a])) + 1e-30f;
        float l = lo[a] - pad, h = hi[a] + pad;
        if (d[a] == 0.0f) {
            if (o[a] < l || o[a] > h) return 0;
            continue;
        }
        float t1 = (l - o[a]) / d[a], t2 = (h - o[a]) / d[a];
        float mn = fminf(t1, t2), mx = fmaxf(t1, t2);
        mn -= fabsf(mn) * 4e-7f; mx += fabsf(mx) * 4e-7f;
        tmn = fmaxf(tmn, mn);
        tmx = fminf(tmx, mx);
    }
    return tmn <= tmx;
}

void orc_bvh_query(const orc_bvh_t* b, const float* rays_o, const float* rays_d, int64_t n_rays, float t_far,
                   float eps, int tri_test, int reference_mode, float* out_t, int32_t* out_i) {
    const int64_t P = b->P;
    const int64_t tri_start = P - 1; /* raycaster.py:235 */
#pragma omp parallel for schedule(dynamic, 256)
    for (int64_t r = 0; r < n_rays; ++r) {
        const float* o = rays_o + 3 * r;
        const float* d = rays_d + 3 * r;
        float t = t_far;
        int64_t slot = -1; /* reference: i = 0 initially (leaf slot 0 -> rank[0]) */
        int32_t best_id = 0;
        int64_t k = 0;
        do {
            int pass = reference_mode ? box_test_reference(b->bmin + 3 * k, b->bmax + 3 * k, o, d, t)
                                      : box_test_conservative(b->bmin + 3 * k, b->bmax + 3 * k, o, d, t);
            if (pass && k >= tri_start) {
                int64_t j = k - tri_start;
                float tt = tri_test == ORC_TRI_UNIT
                               ? orc_tri_unit(o, d, b->g2b + 12 * j, t_far)
                               : orc_tri_mt(o, d, b->tri + 9 * j, b->tri + 9 * j + 3, b->tri + 9 * j + 6, t_far, eps);
                if (reference_mode) { /* raycaster.py:223-224: amin, then `test_t <= t` takes the id */
                    if (tt < t) t = tt;
                    if (tt <= t) slot = j;
                } else if (tt < t_far) {
                    int32_t id = b->rank[j];
                    if (tt < t || (tt == t && id < best_id)) { t = tt; best_id = id; }
                }
            }
            /* scan_next / skip_next, raycaster.py:163-167,250 */
            if (pass && k < tri_start) k = 2 * k + 1;
            else { int64_t q = k + 2; k = (q >> __builtin_ctzll((unsigned long long)q)) - 1; }
        } while (k != 0);
        out_t[r] = t;
        if (reference_mode) out_i[r] = b->rank[slot < 0 ? 0 : slot];
        else out_i[r] = t < t_far ? best_id : 0;
    }
}

/* ---- fp64 referee ------------------------------------------------------------------------------------ */
void orc_referee_f64(const float* verts, const int32_t* tris, int64_t n_tris, const float* rays_o,
                     const float* rays_d, int64_t n_rays, double* best_t, int32_t* best_i, double* second_t,
                     int32_t* second_i, double* best_edge) {
#pragma omp parallel for schedule(dynamic, 16)
    for (int64_t r = 0; r < n_rays; ++r) {
        double o[3] = {rays_o[3 * r], rays_o[3 * r + 1], rays_o[3 * r + 2]};
        double d[3] = {rays_d[3 * r], rays_d[3 * r + 1], rays_d[3 * r + 2]};
        double b1 = INFINITY, b2 = INFINITY, be = 0.0;
        int32_t i1 = -1, i2 = -1;
        for (int64_t k = 0; k < n_tris; ++k) {
            const float* A = verts + 3 * (int64_t)tris[3 * k];
            const float* B = verts + 3 * (int64_t)tris[3 * k + 1];
            const float* C = verts + 3 * (int64_t)tris[3 * k + 2];
            double e1[3], e2[3], cr[3], s[3], sc[3];
            for (int a = 0; a < 3; ++a) { e1[a] = (double)A[a] - C[a]; e2[a] = (double)B[a] - C[a]; s[a] = o[a] - C[a]; }
            cr[0] = d[1] * e2[2] - d[2] * e2[1]; cr[1] = d[2] * e2[0] - d[0] * e2[2]; cr[2] = d[0] * e2[1] - d[1] * e2[0];
            double det = e1[0] * cr[0] + e1[1] * cr[1] + e1[2] * cr[2];
            if (det == 0.0) continue;
            double u = (s[0] * cr[0] + s[1] * cr[1] + s[2] * cr[2]) / det;
            sc[0] = s[1] * e1[2] - s[2] * e1[1]; sc[1] = s[2] * e1[0] - s[0] * e1[2]; sc[2] = s[0] * e1[1] - s[1] * e1[0];
            double v = (d[0] * sc[0] + d[1] * sc[1] + d[2] * sc[2]) / det;
            double t = (e2[0] * sc[0] + e2[1] * sc[1] + e2[2] * sc[2]) / det;
            /* slightly widened acceptance so that fp32-borderline hits are visible to the referee */
            const double tol = 1e-5;
            if (!(u >= -tol && v >= -tol && u + v <= 1.0 + tol && t > 0.0)) continue;
            double edge = fmin(fmin(u, v), 1.0 - u - v);
            if (t < b1) { b2 = b1; i2 = i1; b1 = t; i1 = (int32_t)k; be = edge; }
            else if (t < b2) { b2 = t; i2 = (int32_t)k; }
        }
        best_t[r] = b1; best_i[r] = i1; second_t[r] = b2; second_i[r] = i2; best_edge[r] = be;
    }
}
