"""
The ``Raycaster`` seam of diffrp (reference: diffrp/utils/raycaster.py:13-24) and its B200 implementation.

``B200Raycaster`` is what ``PathTracingSessionOptions(raycaster_impl='b200')`` selects; it has the call shape of the
reference's ``TorchOptiX`` wrapper (raycaster.py:263-296): raw device pointers go to a C-ABI library, the outputs are
allocated by the caller, ``t == far`` on a miss and ``i`` is int32.
"""
import abc
import ctypes as C
from typing import Tuple

import torch

from . import _abi
from ._lib import lib, check


class Raycaster(metaclass=abc.ABCMeta):
    """Same two-method interface as the reference's ``Raycaster`` (raycaster.py:13-24)."""

    def __init__(self, verts: torch.Tensor, tris: torch.IntTensor, config: dict = None) -> None:
        self.config = {} if config is None else config
        self.build(verts, tris, self.config)

    @abc.abstractmethod
    def build(self, verts: torch.Tensor, tris: torch.IntTensor, config: dict) -> None:
        raise NotImplementedError

    @abc.abstractmethod
    def query(self, rays_o: torch.Tensor, rays_d: torch.Tensor, far: float) -> Tuple[torch.Tensor, torch.Tensor]:
        raise NotImplementedError


def _stream_ptr(device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


class B200Raycaster(Raycaster):
    """
    On-GPU LBVH + closest-hit traversal in hand-written sm_100a CUDA.

    Config keys (all optional, same names as the reference passes, path_tracing.py:144-153):
    ``epsilon`` (|det| threshold of the Moller-Trumbore test), ``builder`` (accepted, ignored: there is one builder),
    ``optix_log_level`` (verbosity).
    """

    @torch.no_grad()
    def build(self, verts: torch.Tensor, tris: torch.IntTensor, config: dict) -> None:
        self.handle = None
        self._lib = lib()
        if not verts.is_cuda or not tris.is_cuda:
            raise ValueError("B200Raycaster needs CUDA tensors (there is no CPU path)")
        if verts.ndim != 2 or verts.shape[-1] != 3 or tris.ndim != 2 or tris.shape[-1] != 3:
            raise ValueError("verts must be (V, 3) and tris (F, 3)")
        if len(tris) and (int(tris.min()) < 0 or int(tris.max()) >= len(verts)):
            raise ValueError("triangle indices out of range")
        self._lib.drp_set_log_level(int(config.get('optix_log_level', 0)))
        # the library copies what it needs; these are only kept for introspection (cf. raycaster.py:271-272)
        self.verts = verts.detach().to(torch.float32).contiguous()
        self.tris = tris.detach().to(torch.int32).contiguous()
        self.device = self.verts.device
        dev_index = self.device.index if self.device.index is not None else torch.cuda.current_device()
        handle = C.c_uint64(0)
        inst = config.get('instances')
        if inst is not None:
            # (first_tri (n_inst + 1,) int64, mesh (n_inst,) int32) host arrays: instance q owns triangles [first_tri[q], first_tri[q+1]) and copies mesh[q]
            import numpy as np
            first = np.ascontiguousarray(inst[0], dtype=np.int64)
            mesh = np.ascontiguousarray(inst[1], dtype=np.int32)
            if len(first) != len(mesh) + 1:
                raise ValueError("instances: first_tri must have one more entry than mesh")
            check(self._lib.drp_build_instanced(self.verts.data_ptr(), self.tris.data_ptr(), len(self.verts), len(self.tris), first.ctypes.data,
                                                mesh.ctypes.data, len(mesh), dev_index, _stream_ptr(self.device), C.byref(handle)), "drp_build_instanced")
            self.instanced = True
        else:
            check(self._lib.drp_build(self.verts.data_ptr(), self.tris.data_ptr(), len(self.verts), len(self.tris),
                                      dev_index, _stream_ptr(self.device), C.byref(handle)), "drp_build")
            self.instanced = False
        self.handle = handle.value
        check(self._lib.drp_set_epsilon(self.handle, float(config.get('epsilon', 1e-8))), "drp_set_epsilon")

    def _check_rays(self, rays_o: torch.Tensor, rays_d: torch.Tensor):
        if not (rays_o.is_cuda and rays_d.is_cuda):
            raise ValueError("B200Raycaster.query needs CUDA tensors (there is no CPU path)")
        if rays_o.device != self.device or rays_d.device != self.device:
            raise ValueError("rays live on %s but the structure was built on %s" % (rays_o.device, self.device))
        if rays_o.shape != rays_d.shape or rays_o.ndim != 2 or rays_o.shape[-1] != 3:
            raise ValueError("rays_o and rays_d must both have shape (R, 3)")
        if self.handle is None:
            raise RuntimeError("raycaster has been released")

    @torch.no_grad()
    def query(self, rays_o: torch.Tensor, rays_d: torch.Tensor, far: float) -> Tuple[torch.Tensor, torch.Tensor]:
        self._check_rays(rays_o, rays_d)
        rays_o = rays_o.to(torch.float32).contiguous()
        rays_d = rays_d.to(torch.float32).contiguous()
        n = rays_o.shape[0] if rays_o.ndim > 0 else 0
        out_t = rays_o.new_empty([n])
        out_i = rays_o.new_empty([n], dtype=torch.int32)
        check(self._lib.drp_trace(self.handle, rays_o.data_ptr(), rays_d.data_ptr(), out_t.data_ptr(), out_i.data_ptr(),
                                  float(far), n, _stream_ptr(rays_o.device)), "drp_trace")
        return out_t, out_i

    @torch.no_grad()
    def query_bruteforce(self, rays_o: torch.Tensor, rays_d: torch.Tensor, far: float):
        """Exhaustive O(R*F) query with the same triangle test / tie rule (validation aid)."""
        self._check_rays(rays_o, rays_d)
        rays_o = rays_o.to(torch.float32).contiguous()
        rays_d = rays_d.to(torch.float32).contiguous()
        out_t = rays_o.new_empty([len(rays_o)])
        out_i = rays_o.new_empty([len(rays_o)], dtype=torch.int32)
        check(self._lib.drp_trace_bruteforce(self.verts.data_ptr(), self.tris.data_ptr(), len(self.tris), rays_o.data_ptr(),
                                             rays_d.data_ptr(), out_t.data_ptr(), out_i.data_ptr(), float(far),
                                             float(self.config.get('epsilon', 1e-8)), len(rays_o),
                                             _stream_ptr(rays_o.device)), "drp_trace_bruteforce")
        return out_t, out_i

    @torch.no_grad()
    def refit(self, verts: torch.Tensor) -> None:
        """New vertex positions, same triangles: recompute boxes and triangle records of the existing hierarchy (``drp_refit``; one bottom-up
        pass instead of a rebuild).  Hits are those of a fresh build over the new positions; traversal slows down if the geometry moved far."""
        if self.handle is None:
            raise RuntimeError("raycaster has been released")
        if getattr(self, 'instanced', False):
            raise RuntimeError("instanced structures are rebuilt, not refitted")
        if not verts.is_cuda or verts.device != self.device or tuple(verts.shape) != tuple(self.verts.shape):
            raise ValueError("refit needs a CUDA tensor with the shape of the vertices the structure was built from")
        self.verts = verts.detach().to(torch.float32).contiguous()
        check(self._lib.drp_refit(self.handle, self.verts.data_ptr(), self.tris.data_ptr(), len(self.verts), len(self.tris),
                                  _stream_ptr(self.device)), "drp_refit")

    def check_status(self) -> None:
        """Raise if an earlier traversal on this structure failed on the device (drp_status: non-blocking read of the handle's sticky flag)."""
        check(self._lib.drp_status(self.handle), "drp_status")

    def stats(self) -> dict:
        st = _abi.BVHStats()
        check(self._lib.drp_bvh_stats(self.handle, C.byref(st)), "drp_bvh_stats")
        return dict(n_tris=st.n_tris, n_nodes=st.n_nodes, n_leaves=st.n_leaves, node_bytes=st.node_bytes,
                    tri_bytes=st.tri_bytes, sah_cost=st.sah_cost, bounds=list(st.bounds), max_depth=st.max_depth)

    def release(self):
        if getattr(self, 'handle', None) is not None and getattr(self, '_lib', None) is not None:
            self._lib.drp_release(self.handle)
            self.handle = None

    def __del__(self):
        # guarded for interpreter teardown like the reference (raycaster.py:293-296)
        try:
            self.release()
        except Exception:
            pass
