"""
ctypes mirror of ``include/diffrp_b200.h`` (structs and constants only, no library loading).

Kept separate from :mod:`diffrp_b200._lib` so that struct layouts can be inspected without the CUDA library.
"""
import ctypes as C

ABI_VERSION = 6

WRAP_REPEAT, WRAP_CLAMP, WRAP_MIRROR = 0, 1, 2
INTERP_POINT, INTERP_LINEAR = 0, 1
MAT_DEFAULT, MAT_GLTF = 0, 1
ALPHA_OPAQUE, ALPHA_MASK, ALPHA_BLEND = 0, 1, 2
RNG_NATIVE, RNG_REPLAY = 0, 1
ACCUM_CHANNELS = 16

WRAP_MODES = {'repeat': WRAP_REPEAT, 'clamp': WRAP_CLAMP, 'mirror': WRAP_MIRROR}
INTERP_MODES = {'point': INTERP_POINT, 'linear': INTERP_LINEAR}
ALPHA_MODES = {'OPAQUE': ALPHA_OPAQUE, 'MASK': ALPHA_MASK, 'BLEND': ALPHA_BLEND}


class Texture(C.Structure):
    _fields_ = [
        ("data", C.c_void_p),
        ("h", C.c_int32), ("w", C.c_int32), ("c", C.c_int32),
        ("wrap", C.c_int32), ("interp", C.c_int32), ("_pad", C.c_int32),
    ]


class Material(C.Structure):
    _fields_ = [
        ("kind", C.c_int32), ("alpha_mode", C.c_int32), ("has_emissive", C.c_int32), ("has_normal_tex", C.c_int32),
        ("tint", C.c_float * 4),
        ("base_color_factor", C.c_float * 4),
        ("emissive_factor", C.c_float * 4),
        ("metallic_factor", C.c_float), ("roughness_factor", C.c_float), ("alpha_cutoff", C.c_float), ("texel_tile_log2", C.c_int32),
        ("base_color_tex", Texture), ("mr_tex", Texture), ("normal_tex", Texture), ("emissive_tex", Texture),
        ("texel_records", C.c_void_p),
    ]


class Scene(C.Structure):
    _fields_ = [
        ("world_pos", C.c_void_p), ("world_nrm", C.c_void_p), ("color", C.c_void_p), ("uv", C.c_void_p),
        ("world_tan", C.c_void_p), ("tris", C.c_void_p), ("tri_material", C.c_void_p),
        ("materials", C.POINTER(Material)),
        ("vertex_records", C.c_void_p),
        ("n_verts", C.c_int64), ("n_tris", C.c_int64),
        ("n_materials", C.c_int32), ("_pad", C.c_int32),
        ("env", Texture),
    ]


class RenderParams(C.Structure):
    _fields_ = [
        ("height", C.c_int32), ("width", C.c_int32),
        ("ray_depth", C.c_int32), ("n_samples", C.c_int32),
        ("last_bounce_skybox", C.c_int32), ("rng_mode", C.c_int32),
        ("compaction", C.c_int32), ("reproducible", C.c_int32),
        ("tile_x0", C.c_int32), ("tile_y0", C.c_int32), ("tile_w", C.c_int32), ("tile_h", C.c_int32),
        ("step_epsilon", C.c_float), ("t_far", C.c_float), ("t_near", C.c_float), ("_padf", C.c_float),
        ("cam_pos", C.c_float * 4),
        ("inv_vp", C.c_float * 16),
        ("seed", C.c_uint64),
        ("ndc_x", C.c_void_p), ("ndc_y", C.c_void_p),
        ("jitter_x", C.c_void_p), ("jitter_y", C.c_void_p),
        ("sample_ids", C.c_void_p),
        ("replay_u", C.c_void_p),
        ("shade_wait_event", C.c_void_p),
    ]


class BVHStats(C.Structure):
    _fields_ = [
        ("n_tris", C.c_int64), ("n_nodes", C.c_int64), ("n_leaves", C.c_int64),
        ("node_bytes", C.c_int64), ("tri_bytes", C.c_int64),
        ("sah_cost", C.c_float), ("bounds", C.c_float * 6), ("max_depth", C.c_int32),
    ]


class RenderStats(C.Structure):
    _fields_ = [("rays_traced", C.c_int64), ("rays_nominal", C.c_int64), ("kernel_launches", C.c_int64)]


class Object(C.Structure):
    _fields_ = [
        ("verts", C.c_void_p), ("normals", C.c_void_p), ("color", C.c_void_p), ("uv", C.c_void_p), ("tangents", C.c_void_p),
        ("tris", C.c_void_p), ("M", C.c_float * 16), ("n_verts", C.c_int64), ("n_tris", C.c_int64),
        ("color_channels", C.c_int32), ("_pad", C.c_int32),
    ]


class TonemapParams(C.Structure):
    _fields_ = [("tone", C.c_int32), ("lut_n", C.c_int32), ("lut", C.c_void_p), ("in_stride", C.c_int32), ("alpha_offset", C.c_int32),
                ("flip_rows", C.c_int32), ("scale", C.c_float)]


TONE_LINEAR, TONE_SRGB, TONE_AGX = 0, 1, 2


class Conv3x3Params(C.Structure):
    _fields_ = [("inp", C.c_void_p), ("weight", C.c_void_p), ("bias", C.c_void_p), ("out", C.c_void_p), ("height", C.c_int32), ("width", C.c_int32),
                ("cin", C.c_int32), ("in_stride", C.c_int32), ("in_offset", C.c_int32), ("cout_pad", C.c_int32), ("cout_store", C.c_int32),
                ("out_stride", C.c_int32), ("out_offset", C.c_int32), ("mode", C.c_int32), ("relu", C.c_int32), ("round_tf32", C.c_int32)]


CONV_PLAIN, CONV_POOL2, CONV_UPSAMPLE2 = 0, 1, 2


class Profile(C.Structure):
    _fields_ = [("extend_ms", C.c_double), ("shade_ms", C.c_double), ("extend_launches", C.c_int64), ("shade_launches", C.c_int64),
                ("extend_rays", C.c_int64), ("shade_rays", C.c_int64)]


#: every symbol declared in include/diffrp_b200.h (checked by tests/test_abi.py against the built library)
EXPORTED_SYMBOLS = (
    "drp_abi_version", "drp_build_config", "drp_last_error", "drp_set_log_level", "drp_build", "drp_trace", "drp_trace_bruteforce",
    "drp_release", "drp_set_epsilon", "drp_bvh_stats", "drp_flatten", "drp_render", "drp_finalize", "drp_render_stats", "drp_set_profiling", "drp_get_profile",
    "drp_tonemap", "drp_conv3x3", "drp_denoise_pack", "drp_denoise_unpack", "drp_surface_attrs", "drp_status", "drp_debug_set_stack_limit", "drp_refit", "drp_build_instanced", "drp_upload_batch",
)


# ---- packing helpers (host-neutral: `ptr_of` maps an array object to its raw address) -------------------------

def pack_texture(desc, ptr_of, keep) -> Texture:
    """``desc`` is None or a dict {'image': (H,W,C) fp32 contiguous array, 'wrap': str, 'interp': str}."""
    t = Texture()
    if desc is None:
        return t
    img = desc['image']
    assert len(img.shape) == 3 and img.shape[-1] in (1, 3, 4), "textures must be (H, W, C) with C in {1,3,4}"
    keep.append(img)
    t.data = ptr_of(img)
    t.h, t.w, t.c = int(img.shape[0]), int(img.shape[1]), int(img.shape[2])
    t.wrap = WRAP_MODES[desc.get('wrap', 'repeat')]
    t.interp = INTERP_MODES[desc.get('interp', 'linear')]
    return t


def pack_material(desc, ptr_of, keep) -> Material:
    m = Material()
    m.tint[:] = [1.0, 1.0, 1.0, 1.0]
    m.base_color_factor[:] = [1.0, 1.0, 1.0, 1.0]
    if desc['kind'] == 'default':
        m.kind = MAT_DEFAULT
        tint = desc.get('tint')
        if tint is not None:
            m.tint[:3] = [float(x) for x in tint][:3]
        return m
    assert desc['kind'] == 'gltf', desc['kind']
    m.kind = MAT_GLTF
    m.alpha_mode = ALPHA_MODES[desc.get('alpha_mode', 'OPAQUE')]
    m.base_color_factor[:] = [float(x) for x in desc['base_color_factor']]
    m.metallic_factor = float(desc['metallic_factor'])
    m.roughness_factor = float(desc['roughness_factor'])
    m.alpha_cutoff = float(desc.get('alpha_cutoff', 0.5))
    ef = desc.get('emissive_factor')
    m.has_emissive = int(ef is not None)
    if ef is not None:
        m.emissive_factor[:3] = [float(x) for x in ef]
    m.base_color_tex = pack_texture(desc.get('base_color_tex'), ptr_of, keep)
    m.mr_tex = pack_texture(desc.get('mr_tex'), ptr_of, keep)
    m.normal_tex = pack_texture(desc.get('normal_tex'), ptr_of, keep)
    m.has_normal_tex = int(desc.get('normal_tex') is not None)
    m.emissive_tex = pack_texture(desc.get('emissive_tex'), ptr_of, keep)
    rec = desc.get('texel_records')
    if rec is not None:  # (H,W,12) interleaved copy of the four textures (flatten.texel_records)
        assert int(rec.numel()) == m.base_color_tex.h * m.base_color_tex.w * 12
        keep.append(rec)
        m.texel_records = ptr_of(rec)
        m.texel_tile_log2 = int(desc.get('texel_tile_log2', 0))
    return m


def pack_scene(arrays, materials, env, ptr_of):
    """
    arrays: dict with world_pos (V,3) f32, world_nrm (V,3), color (V,4), uv (V,2), world_tan (V,4),
            tris (F,3) i32, tri_material (F,) i32 -- all contiguous, all on the same side (host or device).
    materials: list of material description dicts; env: texture description dict or None.
    Returns (Scene, keepalive).
    """
    keep = []
    s = Scene()
    for k in ("world_pos", "world_nrm", "color", "uv", "world_tan", "tris", "tri_material"):
        a = arrays[k]
        keep.append(a)
        setattr(s, k, ptr_of(a))
    if arrays.get("vertex_records") is not None:
        keep.append(arrays["vertex_records"])
        s.vertex_records = ptr_of(arrays["vertex_records"])
    s.n_verts = int(arrays["world_pos"].shape[0])
    s.n_tris = int(arrays["tris"].shape[0])
    mats = (Material * max(1, len(materials)))()
    for i, d in enumerate(materials):
        mats[i] = pack_material(d, ptr_of, keep)
    keep.append(mats)
    s.materials = C.cast(mats, C.POINTER(Material))
    s.n_materials = len(materials)
    s.env = pack_texture(env, ptr_of, keep)
    if env is not None:
        s.env.wrap, s.env.interp = WRAP_CLAMP, INTERP_LINEAR
    return s, keep
