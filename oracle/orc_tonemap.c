/*
 * orc_tonemap.c -- CPU oracle of the colour epilogue.  TEST INFRASTRUCTURE (see orc.h).
 *
 * Restates, per pixel:
 *   linear_to_alexa_logc_ei1000   diffrp/utils/colors.py:94-102
 *   sample3d(lut, logc)           diffrp/utils/shader_ops.py:262-310 -> F.grid_sample(5-D, bilinear, padding 'border',
 *                                 align_corners=False) with texcoords*2-1 and the y axis flipped (flipper_3d, :193-195)
 *   linear_to_srgb                diffrp/utils/colors.py:33-42
 *   agx_base_contrast             diffrp/utils/tone_mapping.py:21-35 = srgb(sample3d(lut, logc(rgb)))
 *   to_pil byte conversion        diffrp/utils/exchange.py:17       = (clamp(x,0,1)*255).byte()
 * Pinned against the reference's own outputs in tests/golden/tonemap.npz (tests/test_oracle_golden.py).
 */
#include <math.h>
#include <stdint.h>
#include "orc.h"

static float fsa(float a, float x, float b) { return fmaf(a, x, b); } /* torch.add(b, x, alpha=a) */

float orc_logc(float x) { return x > 0.010591f ? fsa(0.247190f, log10f(fsa(5.555556f, x, 0.052272f)), 0.385537f) : fsa(5.367655f, x, 0.092809f); }
float orc_srgb(float x) { return x < 0.0031308f ? 12.92f * x : fsa(1.055f, powf(x, (float)(1.0 / 2.4)), -0.055f); }

static float src_index(float c, int n, float flip) {
    float g = (c * 2.0f - 1.0f) * flip;
    float i = ((g + 1.0f) * (float)n - 1.0f) / 2.0f;
    if (i < 0.0f) i = 0.0f;
    if (i > (float)(n - 1)) i = (float)(n - 1);
    return i;
}

void orc_lut3d(const float* lut, int n, const float c[3], float out[3]) {
    const float ix = src_index(c[0], n, 1.0f), iy = src_index(c[1], n, -1.0f), iz = src_index(c[2], n, 1.0f);
    const float x0 = floorf(ix), y0 = floorf(iy), z0 = floorf(iz);
    const float x1 = x0 + 1.0f, y1 = y0 + 1.0f, z1 = z0 + 1.0f;
    /* ATen grid_sampler_3d corner weights; t/b = z low/high, n/s = y low/high, w/e = x low/high */
    const float w[8] = {(x1 - ix) * (y1 - iy) * (z1 - iz), (ix - x0) * (y1 - iy) * (z1 - iz), (x1 - ix) * (iy - y0) * (z1 - iz),
                        (ix - x0) * (iy - y0) * (z1 - iz), (x1 - ix) * (y1 - iy) * (iz - z0), (ix - x0) * (y1 - iy) * (iz - z0),
                        (x1 - ix) * (iy - y0) * (iz - z0), (ix - x0) * (iy - y0) * (iz - z0)};
    out[0] = out[1] = out[2] = 0.0f;
    for (int k = 0; k < 8; ++k) {
        const int x = (int)x0 + (k & 1), y = (int)y0 + ((k >> 1) & 1), z = (int)z0 + ((k >> 2) & 1);
        if (x < 0 || y < 0 || z < 0 || x >= n || y >= n || z >= n) continue;
        const float* p = lut + 3 * (((int64_t)z * n + y) * n + x);
        for (int ch = 0; ch < 3; ++ch) out[ch] += p[ch] * w[k];
    }
}

static uint8_t to_byte(float v) {
    if (!(v > 0.0f)) return 0;
    if (v > 1.0f) v = 1.0f;
    return (uint8_t)(int)(v * 255.0f);
}

/* Same contract as drp_tonemap (include/diffrp_b200.h) with host pointers. */
void orc_tonemap(const float* src, int64_t height, int64_t width, const drp_tonemap_params_t* p, uint8_t* out_u8, float* out_f32) {
    const int C = p->alpha_offset >= 0 ? 4 : 3;
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < height * width; ++i) {
        const int64_t row = i / width, col = i - row * width;
        const int64_t o = (p->flip_rows ? height - 1 - row : row) * width + col;
        const float* s = src + i * p->in_stride;
        float v[4] = {s[0] * p->scale, s[1] * p->scale, s[2] * p->scale, 1.0f};
        if (p->tone == DRP_TONE_AGX) {
            const float c[3] = {orc_logc(v[0]), orc_logc(v[1]), orc_logc(v[2])};
            float l[3];
            orc_lut3d(p->lut, p->lut_n, c, l);
            for (int ch = 0; ch < 3; ++ch) v[ch] = orc_srgb(l[ch]);
        } else if (p->tone == DRP_TONE_SRGB) {
            for (int ch = 0; ch < 3; ++ch) v[ch] = orc_srgb(v[ch]);
        }
        if (C == 4) v[3] = s[p->alpha_offset] * p->scale;
        for (int ch = 0; ch < C; ++ch) {
            if (out_f32) out_f32[C * o + ch] = v[ch];
            if (out_u8) out_u8[C * o + ch] = to_byte(v[ch]);
        }
    }
}
