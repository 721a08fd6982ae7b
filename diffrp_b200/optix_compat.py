"""
A module with the call surface of the third-party ``torchoptix`` package as diffrp uses it
(diffrp/utils/raycaster.py:267-296), backed by libdiffrp_b200.so.  Registering it under that name makes diffrp's own,
unmodified ``TorchOptiX`` raycaster -- and therefore its default ``raycaster_impl='torchoptix'`` -- run on the B200 kernels:

    import sys, diffrp_b200.optix_compat
    sys.modules['torchoptix'] = diffrp_b200.optix_compat      # before diffrp builds its first raycaster

All pointers are raw CUDA device addresses (``tensor.data_ptr()``), exactly as diffrp passes them.  Work is enqueued on
torch's current stream of the current device, so results are ordered with the torch ops that consume them.
"""
import ctypes as C

import torch

from ._lib import lib, check

_log_level = 0


def set_log_level(level: int) -> None:
    """torchoptix.set_log_level (raycaster.py:269-270)."""
    global _log_level
    _log_level = int(level)
    lib().drp_set_log_level(int(level))


def build(verts_ptr: int, tris_ptr: int, n_verts: int, n_tris: int) -> int:
    """torchoptix.build (raycaster.py:273-276): (V,3) fp32 vertices, (F,3) int32 indices -> opaque handle."""
    dev = torch.cuda.current_device()
    handle = C.c_uint64(0)
    check(lib().drp_build(verts_ptr, tris_ptr, n_verts, n_tris, dev, torch.cuda.current_stream(dev).cuda_stream, C.byref(handle)), "drp_build")
    return handle.value


def trace_rays(handle: int, rays_o_ptr: int, rays_d_ptr: int, out_t_ptr: int, out_i_ptr: int, far: float, n_rays: int) -> None:
    """torchoptix.trace_rays (raycaster.py:284-290): closest hit, out_t fp32 (== far on a miss), out_i int32."""
    dev = torch.cuda.current_device()
    check(lib().drp_trace(handle, rays_o_ptr, rays_d_ptr, out_t_ptr, out_i_ptr, float(far), int(n_rays),
                          torch.cuda.current_stream(dev).cuda_stream), "drp_trace")


def release(handle: int) -> None:
    """torchoptix.release (raycaster.py:293-296)."""
    lib().drp_release(handle)
