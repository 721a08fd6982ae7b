"""
CPU oracle for the diffrp path-tracing hot path -- numpy front-end of oracle/liborc.so.

TEST INFRASTRUCTURE ONLY (see oracle/orc.h).  Imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / ``--impl reference`` leg; never by the product package.
"""
import os
import ctypes as C
import subprocess
import numpy as np

from diffrp_b200 import _abi

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

TRI_MT, TRI_UNIT = 0, 1
BUILD_SPLITAXIS, BUILD_MORTON = 0, 1
_TRI = {'mt': TRI_MT, 'unit': TRI_UNIT}
_BUILDER = {'splitaxis': BUILD_SPLITAXIS, 'morton': BUILD_MORTON}


def build(force: bool = False) -> str:
    """Compile oracle/liborc.so with the committed Makefile (gcc, OpenMP)."""
    so = os.path.join(_HERE, "liborc.so")
    srcs = [os.path.join(_HERE, f) for f in ("orc_raycast.c", "orc_shade.c", "orc_tonemap.c", "orc.h")]
    srcs.append(os.path.join(_HERE, "..", "include", "diffrp_b200.h"))
    stale = (not os.path.exists(so)) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs)
    if force or stale:
        subprocess.run(["make", "-C", _HERE] + (["-B"] if force else []), check=True, capture_output=True)
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        vp, i64, f32, i32 = C.c_void_p, C.c_int64, C.c_float, C.c_int
        L.orc_num_threads.restype = C.c_int
        L.orc_set_num_threads.argtypes = [C.c_int]
        L.orc_bruteforce.argtypes = [vp, vp, i64, vp, vp, i64, f32, f32, i32, vp, vp]
        L.orc_bvh_build.argtypes = [vp, vp, i64, i32]
        L.orc_bvh_build.restype = vp
        L.orc_bvh_free.argtypes = [vp]
        L.orc_bvh_query.argtypes = [vp, vp, vp, i64, f32, f32, i32, i32, vp, vp]
        L.orc_referee_f64.argtypes = [vp, vp, i64, vp, vp, i64, vp, vp, vp, vp, vp]
        L.orc_sampler_brdf.argtypes = [vp, vp, vp, vp, vp, vp, i64, vp, vp, vp, vp]
        L.orc_surface_attrs.argtypes = [C.POINTER(_abi.Scene), vp, vp, vp, vp, f32, i64, vp]
        L.orc_env_lookup.argtypes = [C.POINTER(_abi.Texture), vp, i64, vp]
        L.orc_texture_sample.argtypes = [C.POINTER(_abi.Texture), vp, i64, vp]
        L.orc_raygen.argtypes = [C.POINTER(_abi.RenderParams), f32, f32, vp, vp, vp, vp]
        L.orc_render.argtypes = [vp, C.POINTER(_abi.Scene), C.POINTER(_abi.RenderParams), vp]
        L.orc_render.restype = i64
        L.orc_philox_uniform6.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, vp]
        L.orc_tonemap.argtypes = [vp, i64, i64, C.POINTER(_abi.TonemapParams), vp, vp]
        _LIB = L
    return _LIB


def _f32(a):
    return np.ascontiguousarray(np.asarray(a), dtype=np.float32)


def _i32(a):
    return np.ascontiguousarray(np.asarray(a), dtype=np.int32)


def _p(a):
    return a.ctypes.data


def num_threads():
    return lib().orc_num_threads()


def set_num_threads(n):
    lib().orc_set_num_threads(int(n))


def bruteforce(verts, tris, rays_o, rays_d, far, eps=1e-8, tri_test='mt'):
    """BruteForceRaycaster.query (raycaster.py:86-97) -> (t f32 (R,), i int32 (R,))."""
    verts, tris, o, d = _f32(verts), _i32(tris), _f32(rays_o), _f32(rays_d)
    t = np.empty(len(o), np.float32)
    i = np.empty(len(o), np.int32)
    lib().orc_bruteforce(_p(verts), _p(tris), len(tris), _p(o), _p(d), len(o), far, eps, _TRI[tri_test], _p(t), _p(i))
    return t, i


class BVH:
    """NaivePBBVH (raycaster.py:120-260) restated: implicit heap, stackless traversal."""

    def __init__(self, verts, tris, builder='splitaxis'):
        self.verts, self.tris = _f32(verts), _i32(tris)
        self.handle = lib().orc_bvh_build(_p(self.verts), _p(self.tris), len(self.tris), _BUILDER[builder])

    def query(self, rays_o, rays_d, far, eps=1e-8, tri_test='mt', reference_mode=False):
        o, d = _f32(rays_o), _f32(rays_d)
        t = np.empty(len(o), np.float32)
        i = np.empty(len(o), np.int32)
        lib().orc_bvh_query(self.handle, _p(o), _p(d), len(o), far, eps, _TRI[tri_test], int(reference_mode), _p(t), _p(i))
        return t, i

    def __del__(self):
        if getattr(self, 'handle', None) and _LIB is not None:
            _LIB.orc_bvh_free(self.handle)
            self.handle = None


def referee(verts, tris, rays_o, rays_d):
    """fp64 exhaustive referee -> dict(best_t, best_i, second_t, second_i, best_edge)."""
    verts, tris, o, d = _f32(verts), _i32(tris), _f32(rays_o), _f32(rays_d)
    n = len(o)
    out = dict(best_t=np.empty(n), best_i=np.empty(n, np.int32), second_t=np.empty(n),
               second_i=np.empty(n, np.int32), best_edge=np.empty(n))
    lib().orc_referee_f64(_p(verts), _p(tris), len(tris), _p(o), _p(d), n, _p(out['best_t']), _p(out['best_i']),
                          _p(out['second_t']), _p(out['second_i']), _p(out['best_edge']))
    return out


def philox_uniform6(seed, pixel, sample, bounce):
    out = np.empty(6, np.float32)
    lib().orc_philox_uniform6(seed, pixel, sample, bounce, _p(out))
    return out


# ---- scene-level ---------------------------------------------------------------------------------------------

class HostScene:
    """A flattened scene (see drp_scene_t) living in numpy arrays."""

    def __init__(self, world_pos, world_nrm, color, uv, world_tan, tris, tri_material, materials, env=None):
        self.arrays = dict(world_pos=_f32(world_pos), world_nrm=_f32(world_nrm), color=_f32(color), uv=_f32(uv),
                           world_tan=_f32(world_tan), tris=_i32(tris), tri_material=_i32(tri_material))
        mats = []
        for m in materials:
            m = dict(m)
            for k in ('base_color_tex', 'mr_tex', 'normal_tex', 'emissive_tex'):
                if m.get(k) is not None:
                    m[k] = dict(m[k], image=_f32(m[k]['image']))
            mats.append(m)
        env = None if env is None else dict(image=_f32(env['image'] if isinstance(env, dict) else env))
        self.struct, self._keep = _abi.pack_scene(self.arrays, mats, env, _p)


def texture_sample(image, uv, wrap='repeat', interp='linear'):
    keep = []
    tex = _abi.pack_texture(dict(image=_f32(image), wrap=wrap, interp=interp), _p, keep)
    uv = _f32(uv)
    out = np.empty((len(uv), tex.c), np.float32)
    lib().orc_texture_sample(C.byref(tex), _p(uv), len(uv), _p(out))
    return out


def env_lookup(image_rh, rays_d):
    keep = []
    d = _f32(rays_d)
    tex = _abi.pack_texture(None if image_rh is None else dict(image=_f32(image_rh)), _p, keep)
    out = np.empty((len(d), 3), np.float32)
    lib().orc_env_lookup(C.byref(tex), _p(d), len(d), _p(out))
    return out


def surface_attrs(scene: HostScene, rays_o, rays_d, t, tri, far):
    o, d, t, tri = _f32(rays_o), _f32(rays_d), _f32(t), _i32(tri)
    out = np.empty((len(o), 12), np.float32)
    lib().orc_surface_attrs(C.byref(scene.struct), _p(o), _p(d), _p(t), _p(tri), far, len(o), _p(out))
    return out


def sampler_brdf(attrs, t, rays_o, rays_d, env_radiance, u6):
    """_sampler_brdf_impl (path_tracing.py:189-236); u6 is (6, R). Returns radiance, transfer, next_o, next_d."""
    a, t, o, d, e, u = _f32(attrs), _f32(t), _f32(rays_o), _f32(rays_d), _f32(env_radiance), _f32(u6)
    n = len(o)
    outs = [np.empty((n, 3), np.float32) for _ in range(4)]
    lib().orc_sampler_brdf(_p(a), _p(t), _p(o), _p(d), _p(e), _p(u), n, *[_p(x) for x in outs])
    return tuple(outs)


def make_params(height, width, ray_depth, t_far, t_near, cam_pos, inv_vp, ndc_x, ndc_y, jitter_x, jitter_y,
                sample_ids=None, step_epsilon=1e-3, last_bounce_skybox=False, seed=0, replay_u=None, tile=None):
    """Host-pointer drp_render_params_t + keepalive."""
    p = _abi.RenderParams()
    keep = dict(ndc_x=_f32(ndc_x), ndc_y=_f32(ndc_y), jitter_x=_f32(jitter_x), jitter_y=_f32(jitter_y))
    n = len(keep['jitter_x'])
    keep['sample_ids'] = _i32(np.arange(n) if sample_ids is None else sample_ids)
    p.height, p.width, p.ray_depth, p.n_samples = height, width, ray_depth, n
    p.last_bounce_skybox = int(last_bounce_skybox)
    p.compaction = 0
    p.step_epsilon, p.t_far, p.t_near = step_epsilon, t_far, t_near
    p.cam_pos[:3] = [float(x) for x in cam_pos]
    p.inv_vp[:] = [float(x) for x in np.asarray(inv_vp, np.float32).reshape(-1)]
    p.seed = seed
    if tile is not None:
        p.tile_x0, p.tile_y0, p.tile_w, p.tile_h = [int(x) for x in tile]
    for k in ('ndc_x', 'ndc_y', 'jitter_x', 'jitter_y', 'sample_ids'):
        setattr(p, k, _p(keep[k]))
    if replay_u is not None:
        keep['replay_u'] = _f32(replay_u)
        assert keep['replay_u'].size == ray_depth * 6 * n * (height * width if tile is None else tile[2] * tile[3])
        p.replay_u = _p(keep['replay_u'])
        p.rng_mode = _abi.RNG_REPLAY
    else:
        p.rng_mode = _abi.RNG_NATIVE
    return p, keep


def raygen(params, jx, jy, keep):
    H, W = params.height, params.width
    o = np.empty((H * W, 3), np.float32)
    d = np.empty((H * W, 3), np.float32)
    lib().orc_raygen(C.byref(params), jx, jy, _p(keep['ndc_x']), _p(keep['ndc_y']), _p(o), _p(d))
    return o, d


def render(bvh: BVH, scene: HostScene, params, accum=None):
    """orc_render: adds n_samples samples into accum (H*W,16); returns (accum, rays_traced)."""
    if accum is None:
        accum = np.zeros((params.height * params.width, _abi.ACCUM_CHANNELS), np.float32)
    n = lib().orc_render(bvh.handle, C.byref(scene.struct), C.byref(params), _p(accum))
    return accum, n


def finalize(accum, height, width, spp_total):
    """trace_rays epilogue (path_tracing.py:348-352): /spp, saturate(alpha), flipud."""
    a = accum.reshape(height, width, _abi.ACCUM_CHANNELS)[::-1] / np.float32(spp_total)
    return dict(radiance=a[..., 0:3].copy(), alpha=np.clip(a[..., 3:4], 0.0, 1.0), albedo=a[..., 4:7].copy(),
                emission=a[..., 7:10].copy(), world_normal=a[..., 10:13].copy(), world_position=a[..., 13:16].copy())


_TONES = {None: _abi.TONE_LINEAR, 'linear': _abi.TONE_LINEAR, 'srgb': _abi.TONE_SRGB, 'agx': _abi.TONE_AGX}


def tonemap(src, tone='agx', lut=None, scale=1.0, alpha_offset=-1, flip_rows=False):
    """orc_tonemap: src (H,W,S) or (N,S) fp32 -> (f32 (..,C), u8 (..,C)), C = 3 or 4 (with alpha).  tone_mapping.py:21-35,
    colors.py:33-42, exchange.py:17."""
    src = _f32(src)
    shape = src.shape[:-1]
    h, w = (shape if len(shape) == 2 else (1, int(np.prod(shape))))
    lut = None if lut is None else _f32(lut)
    p = _abi.TonemapParams(_TONES[tone], 0 if lut is None else lut.shape[0], None if lut is None else lut.ctypes.data, src.shape[-1],
                           alpha_offset, int(flip_rows), scale)
    c = 4 if alpha_offset >= 0 else 3
    out_f, out_b = np.empty(shape + (c,), np.float32), np.empty(shape + (c,), np.uint8)
    lib().orc_tonemap(_p(src), h, w, C.byref(p), _p(out_b), _p(out_f))
    return out_f, out_b


def inputs_from_scene(scene, camera, spp, depth, last_bounce='void', step_eps=1e-3, replay_u=None, seed=0, sample_ids=None):
    """Flatten a diffrp_b200 ``Scene`` on the CPU and build the oracle's HostScene + host-pointer render params.
    ``camera.V()`` / ``camera.P()`` must be CPU tensors (diffrp_b200.ops.set_default_device('cpu'))."""
    from diffrp_b200.flatten import flatten_scene, material_descriptions
    from diffrp_b200.path_tracing import raygen_tables
    vao = flatten_scene(scene.objects, 'cpu')
    mats = []
    for d in material_descriptions(scene.objects, 'cpu'):
        d = dict(d)
        for k in ('base_color_tex', 'mr_tex', 'normal_tex', 'emissive_tex'):
            if d.get(k) is not None:
                d[k] = dict(d[k], image=d[k]['image'].numpy())
        mats.append(d)
    env = None
    for l in scene.lights:
        if hasattr(l, 'image_rh'):
            env = l.image_rh().numpy()
    hs = HostScene(vao.world_pos.numpy(), vao.world_nrm.numpy(), vao.color.numpy(), vao.uv.numpy(), vao.world_tan.numpy(),
                   vao.tris.numpy(), vao.tri_material.numpy(), mats, env=env)
    H, W = camera.resolution()
    tab = raygen_tables(camera.V().cpu(), camera.P().cpu(), H, W, spp, True, 'cpu')
    ids = np.arange(spp) if sample_ids is None else np.asarray(sample_ids)
    p, keep = make_params(H, W, depth, tab['t_far'], tab['t_near'], tab['cam_pos'], tab['inv_vp'], tab['ndc_x'].numpy(),
                          tab['ndc_y'].numpy(), tab['jitter_x'].numpy()[ids], tab['jitter_y'].numpy()[ids], sample_ids=ids,
                          step_epsilon=step_eps, last_bounce_skybox=(last_bounce == 'skybox'), seed=seed, replay_u=replay_u)
    keep['jitter_x_all'], keep['jitter_y_all'] = tab['jitter_x'].numpy(), tab['jitter_y'].numpy()
    return vao, hs, p, keep
