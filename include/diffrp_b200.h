/*
 * diffrp_b200.h -- C ABI of libdiffrp_b200.so, the B200 (sm_100a) drop-in for diffrp's
 * path-tracing hot path.  Plain pointers and sizes only; no torch types.
 *
 * Every entry point returns 0 on success and a non-zero status on failure;
 * drp_last_error() returns a human-readable description of the last failure on the calling
 * thread.  All device pointers are raw CUDA device addresses (tensor.data_ptr()).  All work is
 * enqueued on the caller-supplied cudaStream_t (passed as void*; NULL = legacy default stream)
 * and is asynchronous with respect to the host unless stated otherwise.
 *
 * Which reference interface each entry point replaces (citations into eliphatfs/diffrp v0.2.7):
 *
 *   drp_set_log_level   torchoptix.set_log_level      called at diffrp/utils/raycaster.py:269-270
 *   drp_build           torchoptix.build              called at diffrp/utils/raycaster.py:273-276
 *                       NaivePBBVH.build              diffrp/utils/raycaster.py:122-187
 *   drp_trace           torchoptix.trace_rays         called at diffrp/utils/raycaster.py:284-290
 *                       NaivePBBVH.query              diffrp/utils/raycaster.py:226-260
 *                       BruteForceRaycaster.query     diffrp/utils/raycaster.py:86-97
 *   drp_release         torchoptix.release            called at diffrp/utils/raycaster.py:293-296
 *   drp_set_epsilon     Raycaster.config['epsilon']   diffrp/utils/raycaster.py:91-92, path_tracing.py:144
 *   drp_flatten         RenderSessionMixin.vertex_array_object  diffrp/rendering/mixin.py:74-113
 *   drp_render          PathTracingSession.trace_rays diffrp/rendering/path_tracing.py:310-347
 *                       (section x bounce loop with the built-in sampler_brdf, :250-279)
 *   drp_surface_attrs   layer_material_rays + collectors  diffrp/rendering/path_tracing.py:158-187, mixin.py:115-155
 *   drp_finalize        trace_rays epilogue           diffrp/rendering/path_tracing.py:348-352
 *   drp_tonemap         agx_base_contrast             diffrp/utils/tone_mapping.py:21-35
 *                       linear_to_srgb                diffrp/utils/colors.py:33-42
 *                       linear_to_alexa_logc_ei1000   diffrp/utils/colors.py:94-102
 *                       sample3d (LUT lookup)         diffrp/utils/shader_ops.py:262-310
 *                       to_pil byte conversion        diffrp/utils/exchange.py:7-18
 *   drp_conv3x3         Conv(...) + relu / pool / upsample / concat  diffrp/rendering/denoiser.py:42-66,117-173
 *   drp_denoise_pack / drp_denoise_unpack   run_denoiser pre/post   diffrp/rendering/denoiser.py:24-35
 *   drp_trace_bruteforce (validation aid)             BruteForceRaycaster semantics on the GPU
 */
#ifndef DIFFRP_B200_H
#define DIFFRP_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DRP_ABI_VERSION 6

/* ---- status codes ------------------------------------------------------------------------- */
#define DRP_OK 0
#define DRP_ERR_INVALID 1   /* bad argument                        */
#define DRP_ERR_CUDA 2      /* a CUDA runtime call failed          */
#define DRP_ERR_HANDLE 3    /* unknown / released handle           */
#define DRP_ERR_NOMEM 4     /* device allocation failed            */

/* ---- textures / materials / flattened scene (device pointers) ------------------------------ */

/* wrap modes follow GLTFSampler.wrap_mode (diffrp/materials/gltf_material.py:9-22) */
#define DRP_WRAP_REPEAT 0 /* uv.remainder(1) then grid_sample padding 'reflection' */
#define DRP_WRAP_CLAMP 1  /* grid_sample padding 'border'                          */
#define DRP_WRAP_MIRROR 2 /* grid_sample padding 'reflection'                      */
#define DRP_INTERP_POINT 0
#define DRP_INTERP_LINEAR 1

typedef struct drp_texture {
    const float* data; /* (h, w, c) fp32, row 0 = top (v = 1)      */
    int32_t h, w, c;   /* channels; drp_render requires c == 4 (RGBA-padded; RGB gets alpha 1); data == NULL => absent */
    int32_t wrap;      /* DRP_WRAP_*                               */
    int32_t interp;    /* DRP_INTERP_*                             */
    int32_t _pad;
} drp_texture_t;

#define DRP_MAT_DEFAULT 0 /* DefaultMaterial (diffrp/materials/default_material.py:21-24) */
#define DRP_MAT_GLTF 1    /* GLTFMaterial    (diffrp/materials/gltf_material.py:48-67)    */
#define DRP_ALPHA_OPAQUE 0
#define DRP_ALPHA_MASK 1
#define DRP_ALPHA_BLEND 2

typedef struct drp_material {
    int32_t kind;       /* DRP_MAT_*                                          */
    int32_t alpha_mode; /* DRP_ALPHA_* (GLTF only)                            */
    int32_t has_emissive;
    int32_t has_normal_tex;
    float tint[4];              /* Default: albedo = color.rgb * tint (1,1,1 when None) */
    float base_color_factor[4]; /* GLTF                                                 */
    float emissive_factor[4];   /* GLTF, xyz used                                       */
    float metallic_factor;
    float roughness_factor;
    float alpha_cutoff;
    int32_t texel_tile_log2;    /* layout of texel_records: 0 = row-major (H,W,12); L > 0 = tiles of 2^L x 2^L texels stored contiguously
                                   (tile-major, row-major inside a tile; H and W multiples of 2^L): the four bilinear taps of a hit then
                                   mostly share one 768-byte (L = 2) block instead of two rows W * 48 bytes apart                         */
    drp_texture_t base_color_tex;
    drp_texture_t mr_tex;
    drp_texture_t normal_tex;
    drp_texture_t emissive_tex;
    /* Optional (H,W,12) fp32 interleaved copy of the four textures above, 48 B per texel:
     * [base r g b a | mr.g mr.b normal.x normal.y | normal.z emissive.r emissive.g emissive.b].
     * Only valid when all four textures are present with the same h, w, wrap and interp (those are read from
     * base_color_tex).  drp_render / drp_surface_attrs prefer it: one address computation and 12 instead of 16
     * 128-bit loads per hit, and the taps of the four textures share their 32-byte sectors.  NULL = not provided. */
    const float* texel_records;
} drp_material_t;

/* The scene flattened the way RenderSessionMixin.vertex_array_object does
 * (diffrp/rendering/mixin.py:74-113), with per-vertex world-space normals / tangents already baked
 * ('vectornor' / 'vector3norex1' of SurfaceInput.interpolate_ex, base_material.py:137-143). */
typedef struct drp_scene {
    const float* world_pos;      /* (V,3)                                              */
    const float* world_nrm;      /* (V,3) normalize(M3x3 n)                            */
    const float* color;          /* (V,4)                                              */
    const float* uv;             /* (V,2)                                              */
    const float* world_tan;      /* (V,4) (normalize(M3x3 t.xyz), w)                   */
    const int32_t* tris;         /* (F,3)                                              */
    const int32_t* tri_material; /* (F,)  index into materials (= stencil - 1)         */
    const drp_material_t* materials; /* HOST pointer, n_materials entries (copied)      */
    const float* vertex_records; /* optional (V,16): [pos3 nrm3 uv2 | color4 tan4] interleaved copy of the five arrays above;
                                    64 B per vertex = 2 sectors instead of 5 scattered ones (drp_render prefers it)          */
    int64_t n_verts;
    int64_t n_tris;
    int32_t n_materials;
    int32_t _pad;
    drp_texture_t env; /* ImageEnvironmentLight.image_rh() (lights.py:49-50); NULL data => black */
} drp_scene_t;

#define DRP_RNG_NATIVE 0 /* Philox4x32-10 keyed by (seed; pixel, global sample, bounce) */
#define DRP_RNG_REPLAY 1 /* consume caller-supplied uniforms in the reference's draw order */

typedef struct drp_render_params {
    int32_t height, width;
    int32_t ray_depth;          /* PathTracingSessionOptions.ray_depth                      */
    int32_t n_samples;          /* number of samples rendered by this call                  */
    int32_t last_bounce_skybox; /* pbr_ray_last_bounce == 'skybox'                          */
    int32_t rng_mode;           /* DRP_RNG_*                                                */
    int32_t compaction;         /* 1: drop rays that provably contribute nothing (default)  */
    int32_t reproducible;       /* 1: one sample per internal batch, so every accumulator row receives its fp32 adds in a fixed
                                      (kernel) order and the image is bit-identical from run to run (slower: smaller launches)   */
    int32_t tile_x0, tile_y0;   /* tile sharding: render only the pixel rectangle [x0, x0+w) x [y0, y0+h) (y counted  */
    int32_t tile_w, tile_h;     /* from the bottom row, like ndc_y); w == 0 or h == 0 means the whole frame            */
    float step_epsilon;         /* pbr_ray_step_epsilon                                     */
    float t_far;                /* camera_far()  (mixin.py:41-44)                           */
    float t_near;               /* P[2,3]/(P[2,2]-1) (mixin.py:38)                          */
    float _padf;
    float cam_pos[4];           /* inv(V)[:3,3]                                             */
    float inv_vp[16];           /* inv(P V), row-major                                      */
    uint64_t seed;              /* native RNG key                                           */
    const float* ndc_x;         /* (W,) pixel-centre NDC x  (coordinates.py:6-10)   device  */
    const float* ndc_y;         /* (H,) pixel-centre NDC y, row 0 = bottom          device  */
    const float* jitter_x;      /* (n_samples,) (q_x-0.5)*(2/W) (path_tracing.py:329) device */
    const float* jitter_y;      /* (n_samples,)                                     device  */
    const int32_t* sample_ids;  /* (n_samples,) global sample index (RNG counter)   device  */
    const float* replay_u;      /* replay: (ray_depth, 6, n_samples*H*W)            device  */
    void* shade_wait_event;     /* optional cudaEvent_t (ABI v6): `stream` waits for it right before the first shade launch of the call --
                                   textures / environment still being uploaded on another stream; the first extend does not need them    */
} drp_render_params_t;

/* Accumulator layout: (H*W, 16) fp32, un-normalised sums, row 0 = bottom pixel row:
 *   [0:3] radiance  [3] alpha  [4:7] albedo  [7:10] emission  [10:13] world_normal  [13:16] world_position */
#define DRP_ACCUM_CHANNELS 16

/* One MeshObject as drp_flatten reads it (diffrp/scene/objects.py:11-69, after preprocess()).  The attribute pointers may be
 * device memory or PINNED host memory (read through unified addressing: for host scenes the read is the upload). */
typedef struct drp_object {
    const float* verts;     /* (n_verts, 3)               */
    const float* normals;   /* (n_verts, 3)               */
    const float* color;     /* (n_verts, color_channels)  */
    const float* uv;        /* (n_verts, 2)               */
    const float* tangents;  /* (n_verts, 4)               */
    const int32_t* tris;    /* (n_tris, 3)                */
    float M[16];            /* model matrix, row-major    */
    int64_t n_verts, n_tris;
    int32_t color_channels; /* 3 or 4                     */
    int32_t _pad;
} drp_object_t;

/* ---- entry points -------------------------------------------------------------------------- */

int drp_abi_version(void);
/* build configuration string (compile-time tuning macros, compile date) -- for logs and A/B experiments */
const char* drp_build_config(void);
const char* drp_last_error(void);
int drp_set_log_level(int level);

/* Build the on-GPU LBVH over (verts, tris).  Copies what it needs: no lifetime coupling with the
 * caller's buffers after the call returns (stream-ordered).  *out_handle identifies the structure. */
int drp_build(const float* verts, const int32_t* tris, int64_t n_verts, int64_t n_tris, int device, void* stream,
              uint64_t* out_handle);

/* Closest hit for n rays.  out_t[k] == t_far exactly on a miss; out_i[k] is the 0-based index into
 * tris (int32), 0 on a miss.  Tie rule: minimum t, then minimum primitive index. */
int drp_trace(uint64_t handle, const float* rays_o, const float* rays_d, float* out_t, int32_t* out_i, float t_far,
              int64_t n_rays, void* stream);

/* O(R*F) exhaustive closest hit with the same triangle test and tie rule (validation aid). */
int drp_trace_bruteforce(const float* verts, const int32_t* tris, int64_t n_tris, const float* rays_o,
                         const float* rays_d, float* out_t, int32_t* out_i, float t_far, float epsilon,
                         int64_t n_rays, void* stream);

int drp_release(uint64_t handle);

/* |det| rejection threshold of the Moller-Trumbore test: the `epsilon` entry of the Raycaster config
 * (diffrp/utils/raycaster.py:91-92; PathTracingSessionOptions.raycaster_epsilon, default 1e-8). */
int drp_set_epsilon(uint64_t handle, float epsilon);

/* Statistics about a built structure (host-synchronous; for tests / bench reporting). */
typedef struct drp_bvh_stats {
    int64_t n_tris;
    int64_t n_nodes;
    int64_t n_leaves;
    int64_t node_bytes;
    int64_t tri_bytes;
    float sah_cost;
    float bounds[6];
    int32_t max_depth;
} drp_bvh_stats_t;
int drp_bvh_stats(uint64_t handle, drp_bvh_stats_t* out);

/* Scene flattening in one pass (RenderSessionMixin.vertex_array_object, diffrp/rendering/mixin.py:74-113, + the per-material
 * world transforms of base_material.py:137-143): world positions, normalised world normals / tangents, RGBA colours, uv, offset
 * indices, per-triangle material index, stencils (F+1, leading 0), optional interleaved (V,16) shading records and optional raw
 * concatenations.  `objects` is a HOST array; outputs are device buffers sized from the descriptor totals. */
int drp_flatten(const drp_object_t* objects, int32_t n_objects, float* world_pos, float* world_nrm, float* color4, float* uv,
                float* world_tan, int32_t* tris, int32_t* tri_material, int32_t* stencils, float* records, float* verts_raw,
                float* normals_raw, float* tangents_raw, void* stream);

/* n stream-ordered memcpys (host or device sources, cudaMemcpyDefault) from one call: the upload of a host-resident scene -- the `.to(device)`
 * of every MeshObject / texture tensor that precedes mixin.py:74-113 in the reference -- without one host round trip per tensor. */
int drp_upload_batch(int32_t n, void* const* dst, const void* const* src, const int64_t* bytes, void* stream);

/* Fused wavefront: raygen -> [extend -> shade/sample/accumulate/compact] x ray_depth over
 * n_samples samples of every pixel, added into accum (H*W, 16).  `handle` must have been built over
 * scene->world_pos / scene->tris.  workspace may be NULL (internally allocated and cached on the handle). */
int drp_render(uint64_t handle, const drp_scene_t* scene, const drp_render_params_t* params, float* accum,
               void* stream);

/* The material layer for arbitrary ray batches (custom samplers keep the trace_rays(sampler) protocol,
 * path_tracing.py:281-309, and get attribute interpolation + material evaluation as one kernel).  Replaces, for Default / GLTF
 * materials, layer_material_rays + _super_collector + the g-buffer collect (path_tracing.py:158-187, mixin.py:115-155,
 * interpolator.py:32-48, base_material.py:79-276): attrs (R,12) = [albedo3 | normal3 | metal | smooth | alpha | emission3] at the hit
 * o + d*t of primitive tri[r]; zeros where t >= t_far.  Same per-hit function as the fused shade kernel. */
int drp_surface_attrs(uint64_t handle, const drp_scene_t* scene, const float* rays_o, const float* rays_d, const float* t,
                      const int32_t* tri, float t_far, int64_t n_rays, float* attrs, void* stream);

/* accum (H*W,16) sums -> radiance (H,W,3), alpha (H,W,1) saturated, albedo/emission/world_normal/
 * world_position (H,W,3) each; all divided by spp_total and flipped vertically (row 0 = top). */
int drp_finalize(const float* accum, int32_t height, int32_t width, int32_t spp_total, float* radiance, float* alpha,
                 float* albedo, float* emission, float* world_normal, float* world_position, void* stream);

/* Colour epilogue, one streaming pass instead of the reference's ~12 full-frame torch ops + fp32 D2H:
 *   v   = src[pixel*in_stride + 0..2] * scale                     (scale = 1/spp when src is the accumulator)
 *   rgb = tone == DRP_TONE_AGX  ? linear_to_srgb(sample3d(lut, linear_to_alexa_logc_ei1000(v)))   tone_mapping.py:21-35
 *       : tone == DRP_TONE_SRGB ? linear_to_srgb(v)                                               colors.py:33-42
 *       :                         v
 *   a   = alpha_offset >= 0 ? src[pixel*in_stride + alpha_offset] * scale : absent
 *   out_f32 (H,W,C) = (rgb, a) unclamped (what the reference's functions return); C = 3, or 4 with alpha
 *   out_u8  (H,W,C) = trunc(clamp((rgb, a), 0, 1) * 255)          to_pil, exchange.py:17
 * lut: (n,n,n,3) fp32, laid out as AgxLutLoader.load returns it (z y x 3, fliplr already applied); border-clamped
 * trilinear lookup with grid_sample(align_corners=False) arithmetic.  flip_rows != 0 writes row r to H-1-r (trace_rays'
 * flipud).  Either output may be NULL.  All pointers are device addresses. */
#define DRP_TONE_LINEAR 0
#define DRP_TONE_SRGB 1
#define DRP_TONE_AGX 2
typedef struct drp_tonemap_params {
    int32_t tone;
    int32_t lut_n;
    const float* lut;
    int32_t in_stride;    /* floats per source pixel (16 for the accumulator, 3 for an (N,3) tensor) */
    int32_t alpha_offset; /* -1: no alpha channel */
    int32_t flip_rows;
    float scale;
} drp_tonemap_params_t;
int drp_tonemap(const float* src, int64_t height, int64_t width, const drp_tonemap_params_t* params, uint8_t* out_u8,
                float* out_f32, void* stream);

/* ---- denoiser: the 3x3 convolutions of the OIDN-style U-Net the reference runs after pbr() -------------------
 * Replaces nn.Conv2d(cin, cout, 3, padding=1) + F.relu (+ F.max_pool2d(x, 2, 2) | F.interpolate(scale_factor=2, 'nearest') +
 * torch.cat) of diffrp/rendering/denoiser.py:42-66,117-173 with one tcgen05 (TF32, fp32 accumulate in TMEM) implicit-GEMM kernel per
 * layer; cuDNN runs the reference's fp32 convolutions in TF32 as well (torch.backends.cudnn.allow_tf32 defaults to True).
 * Activations are NHWC fp32 with a pixel stride, so a layer can read / write a channel slice of a wider (concatenation) buffer:
 *   in  : pixel (y,x) channel c at in [(y*width + x)*in_stride  + in_offset  + c],  c < cin   (cin  % 16 == 0)
 *   w   : [cout_pad][9*cin] fp32, k = (ky*3 + kx)*cin + c   (cout_pad % 16 == 0, 16..256; rows >= the real cout are zero)
 *   bias: [cout_pad]
 *   out : mode DRP_CONV_PLAIN     (y,x)        -> out[(y*width + x)*out_stride + out_offset + c], c < cout_store
 *         mode DRP_CONV_POOL2     max over 2x2 -> out[((y/2)*(width/2) + x/2)*out_stride + ...]       (height, width even)
 *         mode DRP_CONV_UPSAMPLE2 replicated   -> out[((2y+a)*(2*width) + 2x+b)*out_stride + ...], a,b in {0,1}
 * All strides / offsets / cout_store in floats, multiples of 4 (16 for the input side): the tensor stores move 16-byte units.  relu != 0 applies max(.,0) before pooling / replication. */
#define DRP_CONV_PLAIN 0
#define DRP_CONV_POOL2 1
#define DRP_CONV_UPSAMPLE2 2
typedef struct drp_conv3x3_params {
    const float* in;
    const float* weight;
    const float* bias;
    float* out;
    int32_t height, width;
    int32_t cin, in_stride, in_offset;
    int32_t cout_pad, cout_store, out_stride, out_offset;
    int32_t mode, relu;
    int32_t round_tf32; /* != 0: round the stored activations to TF32 (nearest) -- for outputs that only feed further drp_conv3x3 layers,
                         * so that the tensor core's operand truncation is exact and rounding stays unbiased across the net */
} drp_conv3x3_params_t;
int drp_conv3x3(const drp_conv3x3_params_t* params, void* stream);

/* Network input / output of run_denoiser (diffrp/rendering/denoiser.py:24-35), one kernel each:
 * pack:   cat([linear_to_pu(hdr) / linear_to_pu(65504), albedo_srgb, normal*0.5+0.5]) (utils/colors.py:18-23), reflection-padded
 *         (nn.ReflectionPad2d, dh//2 rows on top, dw//2 columns on the left) to (padded_height, padded_width), written as 16 channels
 *         (9 + 7 zeros) at dst[(y*padded_width + x)*dst_stride + dst_offset + c];
 * unpack: crop of the 3 output channels + pu_to_linear(x * linear_to_pu(65504)) (utils/colors.py:25-30) -> out (height, width, 3). */
int drp_denoise_pack(const float* hdr, const float* albedo_srgb, const float* normal, int32_t height, int32_t width, float* dst,
                     int32_t padded_height, int32_t padded_width, int32_t dst_stride, int32_t dst_offset, void* stream);
int drp_denoise_unpack(const float* src, int32_t padded_height, int32_t padded_width, int32_t src_stride, float* out, int32_t height,
                       int32_t width, void* stream);

/* Counters of the last drp_render call on this handle (host-synchronous). */
typedef struct drp_render_stats {
    int64_t rays_traced;     /* live rays actually traced (sum over bounces)         */
    int64_t rays_nominal;    /* H*W*n_samples*ray_depth                              */
    int64_t kernel_launches; /* kernels launched by the call                         */
} drp_render_stats_t;
int drp_render_stats(uint64_t handle, drp_render_stats_t* out);

/* Dynamic scenes (the reference rebuilds its structure for every session: rendering/path_tracing.py:142-156 builds inside the
 * single-use session, rendering/mixin.py:74-113 re-flattens the scene): keep the hierarchy's topology and recompute every box and triangle
 * record for new vertex positions -- one bottom-up pass, no sort, no collapse.  `tris` must be the index array the structure was built
 * from (n_tris must match); only `verts` may differ.  Boxes stay conservative and triangle records are rounded exactly like the builder's,
 * so closest hits equal those of a fresh drp_build over the same arrays bit for bit; only traversal speed depends on how far the geometry moved. */
int drp_refit(uint64_t handle, const float* verts, const int32_t* tris, int64_t n_verts, int64_t n_tris, void* stream);

/* Instanced scenes (BASELINE configs[4]: 1000 MeshObjects sharing one mesh; the reference flattens them, scene/scene.py:33-75,
 * rendering/mixin.py:74-113): the input is still the flattened world-space geometry (what drp_flatten writes), plus the instance table:
 * instance q owns triangles [inst_first_tri[q], inst_first_tri[q+1]) and is a copy (same connectivity, any transform) of mesh inst_mesh[q].
 * One hierarchy is built per mesh; its topology is replicated for every instance and refitted to that instance's world-space triangles, and a
 * small instance level (built on the host over the instance root boxes) joins them into ONE wide hierarchy over world-space triangles.
 * Traversal, primitive ids and hits are those of drp_build over the same arrays (bit for bit); the sort / Karras / collapse cost is paid per
 * mesh instead of per triangle.  B200's 180 GB of HBM holds the replicated nodes (80 B per ~2.6 triangles), so no per-ray transform is needed. */
int drp_build_instanced(const float* verts, const int32_t* tris, int64_t n_verts, int64_t n_tris, const int64_t* inst_first_tri,
                        const int32_t* inst_mesh, int64_t n_inst, int device, void* stream, uint64_t* out_handle);

/* Health of a handle, without synchronising.  A ray whose traversal outgrows the per-thread stack of the fast kernel is re-traced by a fix-up
 * kernel with a 256-entry stack (deeper than any hierarchy the builder can emit: <= 63 Morton + 27 index levels), so results stay exact on
 * degenerate scenes.  Should even that stack run out, the kernel raises a sticky host-mapped flag: drp_status and EVERY later call on the handle
 * (drp_trace, drp_render, drp_surface_attrs, drp_render_stats) return DRP_ERR_INVALID -- never a silent wrong hit.  (The reference's
 * NaivePBBVH.query is stackless, utils/raycaster.py:226-260; torchoptix's trace has no failure mode to mirror.) */
int drp_status(uint64_t handle);
/* Test hook: lower the per-thread stack depth of the fast traversal path (0..48 entries) so that ordinary scenes exercise the fix-up path. */
int drp_debug_set_stack_limit(uint64_t handle, int32_t entries);

/* Optional per-kernel timing (measurement aid, no reference equivalent): when enabled, drp_render brackets every
 * extend / shade launch with CUDA events on the launch stream; drp_get_profile synchronises, sums the elapsed times
 * since the last call, and resets.  rays_* are the live rays those launches processed. */
typedef struct drp_profile {
    double extend_ms;
    double shade_ms;
    int64_t extend_launches;
    int64_t shade_launches;
    int64_t extend_rays;
    int64_t shade_rays;
} drp_profile_t;
int drp_set_profiling(uint64_t handle, int enable);
int drp_get_profile(uint64_t handle, drp_profile_t* out);

#ifdef __cplusplus
}
#endif
#endif /* DIFFRP_B200_H */
