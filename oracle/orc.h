/*
 * orc.h -- CPU oracle for the diffrp path-tracing hot path.
 *
 * TEST INFRASTRUCTURE.  This is a plain-C restatement of the *reference's* algorithm
 * (eliphatfs/diffrp v0.2.7), used only as the checker by tests/, __graft_entry__.smoke() and by
 * bench.py's cpu_baseline / --impl reference leg.  The product (diffrp_b200/) never links, imports
 * or calls anything in this directory and fails loudly when its CUDA library is missing.
 *
 * Parity pinning: the reference ships no tests / golden vectors for this path,
 * so the oracle is pinned against outputs of the reference itself, generated in the build container
 * by tests/golden/make_golden.py (imports /root/reference through tests/golden/refharness.py) and
 * committed under tests/golden/ (npz files); tests/test_oracle_golden.py checks every one of them.
 */
#ifndef ORC_H
#define ORC_H
#include <stdint.h>
#include "../include/diffrp_b200.h" /* scene / material / render-param structs (host pointers here) */

#ifdef __cplusplus
extern "C" {
#endif

/* triangle tests */
#define ORC_TRI_MT 0   /* Moller-Trumbore, raycaster.py:58-79 (BruteForceRaycaster)            */
#define ORC_TRI_UNIT 1 /* pre-transformed unit-triangle test, raycaster.py:27-55 (NaivePBBVH)  */
/* closest-hit selection */
#define ORC_TIE_MIN_ID 0    /* min t, then min primitive id == torch.argmin over triangles (raycaster.py:94-97) */
#define ORC_TIE_REFERENCE 1 /* NaivePBBVH: `test_t <= t` in visit order, later visit wins (raycaster.py:223-224) */
/* builders */
#define ORC_BUILD_SPLITAXIS 0 /* raycaster.py:137-157 */
#define ORC_BUILD_MORTON 1    /* raycaster.py:110-117,135-136 */

typedef struct orc_bvh orc_bvh_t;

int orc_num_threads(void);
void orc_set_num_threads(int n);

void orc_bruteforce(const float* verts, const int32_t* tris, int64_t n_tris, const float* rays_o, const float* rays_d,
                    int64_t n_rays, float t_far, float eps, int tri_test, float* out_t, int32_t* out_i);

orc_bvh_t* orc_bvh_build(const float* verts, const int32_t* tris, int64_t n_tris, int builder);
void orc_bvh_free(orc_bvh_t* b);
/* reference_mode = 0: conservative (NaN-safe, padded) slab test, result == orc_bruteforce(ORC_TIE_MIN_ID);
 * reference_mode = 1: NaivePBBVH.query verbatim semantics (NaN-propagating slab test, visit-order ties). */
void orc_bvh_query(const orc_bvh_t* b, const float* rays_o, const float* rays_d, int64_t n_rays, float t_far,
                   float eps, int tri_test, int reference_mode, float* out_t, int32_t* out_i);

/* fp64 referee: exhaustive Moller-Trumbore in double. Per ray: best t, best id, second-best t, second id,
 * and min(u, v, 1-u-v) of the best hit (distance to the nearest edge in barycentric units). t = +inf when none. */
void orc_referee_f64(const float* verts, const int32_t* tris, int64_t n_tris, const float* rays_o,
                     const float* rays_d, int64_t n_rays, double* best_t, int32_t* best_i, double* second_t,
                     int32_t* second_i, double* best_edge);

/* ---- shading / bounce loop -------------------------------------------------------------------- */

/* _sampler_brdf_impl, path_tracing.py:189-236.  attrs (R,12), u (6,R) -> outputs.  */
void orc_sampler_brdf(const float* attrs, const float* t, const float* rays_o, const float* rays_d,
                      const float* env_radiance, const float* u6, int64_t n, float* radiance, float* transfer,
                      float* next_o, float* next_d);

/* material + attribute evaluation for a batch of hits: layer_material_rays / _super_collector
 * (path_tracing.py:158-187).  Writes attrs (R,12) = [albedo3|normal3|metal|smooth|alpha|emission3]; zeros on miss. */
void orc_surface_attrs(const drp_scene_t* scene, const float* rays_o, const float* rays_d, const float* t,
                       const int32_t* tri, float t_far, int64_t n, float* attrs);

/* env lookup, path_tracing.py:267 (sample2d of image_rh at the lat-long uv of d), (R,3) */
void orc_env_lookup(const drp_texture_t* env, const float* rays_d, int64_t n, float* out_rgb);

/* texture fetch with grid_sample semantics, shader_ops.py:198-221 + gltf_material.py:15-22 */
void orc_texture_sample(const drp_texture_t* tex, const float* uv, int64_t n, float* out);

/* primary rays for one sample, mixin.py:31-39 + path_tracing.py:329-331; out (H*W,3) each */
void orc_raygen(const drp_render_params_t* p, float jx, float jy, const float* ndc_x, const float* ndc_y, float* o,
                float* d);

/* trace_rays over params->n_samples samples (path_tracing.py:325-347) with the built-in sampler_brdf;
 * adds un-normalised sums into accum (H*W,16).  All pointers are HOST pointers.  The closest-hit search uses
 * `bvh` (conservative mode, ORC_TRI_MT, ORC_TIE_MIN_ID).  Returns number of ray-bounces traced. */
int64_t orc_render(const orc_bvh_t* bvh, const drp_scene_t* scene, const drp_render_params_t* params, float* accum);

/* native RNG (shared definition with the CUDA kernels): Philox4x32-10 */
void orc_philox_uniform6(uint64_t seed, uint32_t pixel, uint32_t sample, uint32_t bounce, float out6[6]);

/* ---- colour epilogue (orc_tonemap.c): drp_tonemap's contract with host pointers ------------------------------ */
float orc_logc(float x);
float orc_srgb(float x);
void orc_lut3d(const float* lut, int n, const float c[3], float out[3]);
void orc_tonemap(const float* src, int64_t height, int64_t width, const drp_tonemap_params_t* p, uint8_t* out_u8, float* out_f32);

#ifdef __cplusplus
}
#endif
#endif
