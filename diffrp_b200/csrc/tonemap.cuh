// tonemap.cuh -- per-pixel colour epilogue: the reference's linear_to_alexa_logc_ei1000 (utils/colors.py:94-102),
// sample3d border-clamped trilinear LUT lookup (utils/shader_ops.py:262-310 -> F.grid_sample, align_corners=False),
// linear_to_srgb (utils/colors.py:33-42) and to_pil's byte conversion (utils/exchange.py:17), as one function per pixel.
// Arithmetic uses the explicitly rounded helpers of common.cuh so the CUDA kernel and the host build (tests/hostsim) agree;
// only powf / log10f differ between libm and CUDA (<= 2 ulp).
#pragma once
#include "common.cuh"
#include "../../include/diffrp_b200.h"

// fsa(a, x, b) = torch.add(b, x, alpha=a) = b + a*x (shader_ops.py:70-74); ATen's vectorised CPU kernel fuses it
DRP_HD float tm_fsa(float a, float x, float b) { return x_fma(a, x, b); }

DRP_HD float tm_logc(float x) {  // colors.py:94-102
    return x > 0.010591f ? tm_fsa(0.247190f, log10f(tm_fsa(5.555556f, x, 0.052272f)), 0.385537f) : tm_fsa(5.367655f, x, 0.092809f);
}
DRP_HD float tm_srgb(float x) {  // colors.py:42
    return x < 0.0031308f ? x_mul(12.92f, x) : tm_fsa(1.055f, powf(x, (float)(1.0 / 2.4)), -0.055f);
}
// grid_sample source index for a texture coordinate c in [0,1] (flip: the y axis, flipper_3d = (1,-1,1)), border padding
DRP_HD float tm_src_index(float c, int n, bool flip) {
    float g = x_add(x_mul(c, 2.0f), -1.0f);
    if (flip) g = -g;
    float i = x_mul(x_add(x_mul(x_add(g, 1.0f), (float)n), -1.0f), 0.5f);
    return fminf(fmaxf(i, 0.0f), (float)(n - 1));
}
// lut: (n,n,n,3) z y x c.  Weights and summation order follow ATen's grid_sampler_3d (tnw, tne, tsw, tse, bnw, bne, bsw, bse).
DRP_HD void tm_lut3d(const float* __restrict__ lut, int n, float cx, float cy, float cz, float out[3]) {
    const float ix = tm_src_index(cx, n, false), iy = tm_src_index(cy, n, true), iz = tm_src_index(cz, n, false);
    const float fx0 = floorf(ix), fy0 = floorf(iy), fz0 = floorf(iz);
    const int x0 = (int)fx0, y0 = (int)fy0, z0 = (int)fz0;
    const float wx1 = x_add(ix, -fx0), wy1 = x_add(iy, -fy0), wz1 = x_add(iz, -fz0);
    const float wx0 = x_add(x_add(fx0, 1.0f), -ix), wy0 = x_add(x_add(fy0, 1.0f), -iy), wz0 = x_add(x_add(fz0, 1.0f), -iz);
    out[0] = out[1] = out[2] = 0.0f;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const int dx = k & 1, dy = (k >> 1) & 1, dz = (k >> 2) & 1;
        const int x = x0 + dx, y = y0 + dy, z = z0 + dz;
        if (x >= n || y >= n || z >= n) continue;  // within_bounds_3d (the weight is 0 there)
        const float w = x_mul(x_mul(dx ? wx1 : wx0, dy ? wy1 : wy0), dz ? wz1 : wz0);
        const float* p = lut + 3 * (((int64_t)z * n + y) * n + x);
        out[0] = x_add(out[0], x_mul(ldg(p), w));
        out[1] = x_add(out[1], x_mul(ldg(p + 1), w));
        out[2] = x_add(out[2], x_mul(ldg(p + 2), w));
    }
}
DRP_HD uint8_t tm_byte(float v) {  // (saturate(v) * 255).byte(): truncation; NaN -> 0
    if (!(v > 0.0f)) return 0;
    return (uint8_t)(int)x_mul(fminf(v, 1.0f), 255.0f);
}
// one pixel: src -> rgb (+ alpha)
DRP_HD void tm_pixel(const float* __restrict__ src, const drp_tonemap_params_t& p, float out[4]) {
    float v[3] = {x_mul(src[0], p.scale), x_mul(src[1], p.scale), x_mul(src[2], p.scale)};
    if (p.tone == DRP_TONE_AGX) {
        float l[3];
        tm_lut3d(p.lut, p.lut_n, tm_logc(v[0]), tm_logc(v[1]), tm_logc(v[2]), l);
        v[0] = tm_srgb(l[0]); v[1] = tm_srgb(l[1]); v[2] = tm_srgb(l[2]);
    } else if (p.tone == DRP_TONE_SRGB) {
        v[0] = tm_srgb(v[0]); v[1] = tm_srgb(v[1]); v[2] = tm_srgb(v[2]);
    }
    out[0] = v[0]; out[1] = v[1]; out[2] = v[2];
    out[3] = p.alpha_offset >= 0 ? x_mul(src[p.alpha_offset], p.scale) : 1.0f;
}
