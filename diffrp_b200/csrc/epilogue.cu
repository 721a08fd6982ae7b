// epilogue.cu -- drp_tonemap: normalise + tone-map + sRGB + (optional) 8-bit quantisation of a frame in one streaming pass.
// Replaces agx_base_contrast (diffrp/utils/tone_mapping.py:21-35), linear_to_srgb (utils/colors.py:33-42), the torch.cat with
// alpha and to_pil's saturate*255 -> byte (utils/exchange.py:7-18) of the documented post-pbr() workflow; when the source is the
// path tracer's accumulator it also folds in trace_rays' /spp and flipud (rendering/path_tracing.py:348-351), so the frame
// leaves the GPU as 4 bytes per pixel instead of 64.
// HBM-streaming: one thread per pixel, 16-64 B read (the LUT, <= a few hundred KB, stays in L1/L2), 3-16 B written.
#include "internal.h"
#include "tonemap.cuh"

__global__ void __launch_bounds__(256) k_tonemap(const float* __restrict__ src, int64_t height, int64_t width, drp_tonemap_params_t p,
                                                 uint8_t* __restrict__ out_u8, float* __restrict__ out_f32) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= height * width) return;
    const int64_t row = i / width, col = i - row * width;
    const int64_t o = (p.flip_rows ? height - 1 - row : row) * width + col;
    float v[4];
    const float* s = src + i * p.in_stride;
    if (p.in_stride == 16) {  // accumulator rows: one 128-bit load brings radiance + alpha
        const float4 a = __ldg(reinterpret_cast<const float4*>(s));
        const float t[4] = {a.x, a.y, a.z, a.w};
        tm_pixel(t, p, v);
    } else {
        tm_pixel(s, p, v);
    }
    const bool with_alpha = p.alpha_offset >= 0;
    if (out_f32) {
        if (with_alpha) reinterpret_cast<float4*>(out_f32)[o] = make_float4(v[0], v[1], v[2], v[3]);
        else { out_f32[3 * o] = v[0]; out_f32[3 * o + 1] = v[1]; out_f32[3 * o + 2] = v[2]; }
    }
    if (out_u8) {
        if (with_alpha) reinterpret_cast<uchar4*>(out_u8)[o] = make_uchar4(tm_byte(v[0]), tm_byte(v[1]), tm_byte(v[2]), tm_byte(v[3]));
        else { out_u8[3 * o] = tm_byte(v[0]); out_u8[3 * o + 1] = tm_byte(v[1]); out_u8[3 * o + 2] = tm_byte(v[2]); }
    }
}

extern "C" int drp_tonemap(const float* src, int64_t height, int64_t width, const drp_tonemap_params_t* params, uint8_t* out_u8,
                           float* out_f32, void* stream) {
    if (!params || height < 0 || width < 0) { drp_set_error("drp_tonemap: invalid argument"); return DRP_ERR_INVALID; }
    const drp_tonemap_params_t p = *params;
    if (p.tone < DRP_TONE_LINEAR || p.tone > DRP_TONE_AGX) { drp_set_error("drp_tonemap: unknown tone operator"); return DRP_ERR_INVALID; }
    if (p.tone == DRP_TONE_AGX && (!p.lut || p.lut_n < 2)) { drp_set_error("drp_tonemap: DRP_TONE_AGX needs a (n,n,n,3) LUT with n >= 2"); return DRP_ERR_INVALID; }
    if (p.in_stride < 3 || p.alpha_offset >= p.in_stride) { drp_set_error("drp_tonemap: bad in_stride / alpha_offset"); return DRP_ERR_INVALID; }
    if (p.in_stride == 16 && p.alpha_offset >= 0 && p.alpha_offset != 3) { drp_set_error("drp_tonemap: accumulator rows carry alpha at offset 3"); return DRP_ERR_INVALID; }
    if (!out_u8 && !out_f32) { drp_set_error("drp_tonemap: no output requested"); return DRP_ERR_INVALID; }
    const int64_t n = height * width;
    if (n == 0) return DRP_OK;
    if (!src) { drp_set_error("drp_tonemap: src is NULL"); return DRP_ERR_INVALID; }
    k_tonemap<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(src, height, width, p, out_u8, out_f32);
    DRP_CUDA_CHECK(cudaGetLastError());
    return DRP_OK;
}
