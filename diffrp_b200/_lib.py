"""
ctypes binding of ``libdiffrp_b200.so`` (the C ABI in ``include/diffrp_b200.h``).

There is deliberately no CPU fallback: if the CUDA library cannot be built or loaded, every entry point raises.
"""
import os
import ctypes as C
from . import _abi
from .build import build_library, LIB_PATH

_LIB = None


class DiffrpB200Error(RuntimeError):
    pass


def _declare(L):
    vp, i64, i32, f32, u64 = C.c_void_p, C.c_int64, C.c_int32, C.c_float, C.c_uint64
    L.drp_abi_version.restype = C.c_int
    L.drp_last_error.restype = C.c_char_p
    L.drp_build_config.restype = C.c_char_p
    L.drp_set_log_level.argtypes = [C.c_int]
    L.drp_build.argtypes = [vp, vp, i64, i64, C.c_int, vp, C.POINTER(u64)]
    L.drp_trace.argtypes = [u64, vp, vp, vp, vp, f32, i64, vp]
    L.drp_trace_bruteforce.argtypes = [vp, vp, i64, vp, vp, vp, vp, f32, f32, i64, vp]
    L.drp_release.argtypes = [u64]
    L.drp_set_epsilon.argtypes = [u64, f32]
    L.drp_bvh_stats.argtypes = [u64, C.POINTER(_abi.BVHStats)]
    L.drp_render.argtypes = [u64, C.POINTER(_abi.Scene), C.POINTER(_abi.RenderParams), vp, vp]
    L.drp_finalize.argtypes = [vp, i32, i32, i32, vp, vp, vp, vp, vp, vp, vp]
    L.drp_render_stats.argtypes = [u64, C.POINTER(_abi.RenderStats)]
    L.drp_upload_batch.argtypes = [i32, vp, vp, vp, vp]
    L.drp_status.argtypes = [u64]
    L.drp_refit.argtypes = [u64, vp, vp, i64, i64, vp]
    L.drp_build_instanced.argtypes = [vp, vp, i64, i64, vp, vp, i64, C.c_int, vp, C.POINTER(C.c_uint64)]
    L.drp_debug_set_stack_limit.argtypes = [u64, i32]
    L.drp_flatten.argtypes = [C.POINTER(_abi.Object), i32] + [vp] * 13
    L.drp_set_profiling.argtypes = [u64, C.c_int]
    L.drp_get_profile.argtypes = [u64, C.POINTER(_abi.Profile)]
    L.drp_tonemap.argtypes = [vp, i64, i64, C.POINTER(_abi.TonemapParams), vp, vp, vp]
    L.drp_surface_attrs.argtypes = [u64, C.POINTER(_abi.Scene), vp, vp, vp, vp, f32, i64, vp, vp]
    L.drp_conv3x3.argtypes = [C.POINTER(_abi.Conv3x3Params), vp]
    L.drp_denoise_pack.argtypes = [vp, vp, vp, i32, i32, vp, i32, i32, i32, i32, vp]
    L.drp_denoise_unpack.argtypes = [vp, i32, i32, i32, vp, i32, i32, vp]
    for name in _abi.EXPORTED_SYMBOLS:
        fn = getattr(L, name)
        if name not in ("drp_last_error", "drp_build_config"):
            fn.restype = C.c_int


def lib():
    """Load (building first if the sources are newer) the CUDA library; raises if that is impossible."""
    global _LIB
    if _LIB is None:
        path = build_library() if os.environ.get("DIFFRP_B200_NO_BUILD") != "1" else LIB_PATH
        if os.environ.get("DIFFRP_B200_LIB"):  # A/B experiments: a library built with other -D switches (tools/build_variants.sh)
            path = os.environ["DIFFRP_B200_LIB"]
        if not os.path.exists(path):
            raise DiffrpB200Error("libdiffrp_b200.so is missing and could not be built; there is no CPU fallback")
        L = C.CDLL(path)
        _declare(L)
        if L.drp_abi_version() != _abi.ABI_VERSION:
            raise DiffrpB200Error("libdiffrp_b200.so ABI version mismatch")
        _LIB = L
    return _LIB


def check(status: int, what: str = ""):
    if status != 0:
        msg = lib().drp_last_error()
        raise DiffrpB200Error("%s failed (status %d): %s" % (what, status, msg.decode() if msg else "?"))


def build_config() -> str:
    return lib().drp_build_config().decode()


def loaded_path():
    return LIB_PATH if _LIB is not None else None
