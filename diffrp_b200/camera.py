"""
Cameras with the interface of the reference's ``diffrp.rendering.camera`` (camera.py:10-137):
``V()`` / ``P()`` return (4,4) fp32 GL view / projection matrices, ``resolution()`` returns (h, w).
Host-side numpy only; nothing here is on the hot path.
"""
import math
from typing import List, Union

import numpy
import torch

from .ops import gpu_f32


def _translation(xyz) -> numpy.ndarray:
    m = numpy.identity(4, dtype=numpy.float64)
    m[:3, 3] = numpy.asarray(xyz, dtype=numpy.float64)[:3]
    return m


def _unit(v) -> numpy.ndarray:
    v = numpy.asarray(v, dtype=numpy.float64)
    return v / numpy.linalg.norm(v)


def gl_perspective(width: int, height: int, cx: float, cy: float, fx: float, fy: float, near: float, far: float):
    """OpenGL clip-space projection from pinhole intrinsics (what the reference obtains from calibur, camera.py:80-88)."""
    return numpy.array([
        [2.0 * fx / width, 0.0, 1.0 - 2.0 * cx / width, 0.0],
        [0.0, 2.0 * fy / height, 2.0 * cy / height - 1.0, 0.0],
        [0.0, 0.0, (far + near) / (near - far), 2.0 * far * near / (near - far)],
        [0.0, 0.0, -1.0, 0.0],
    ], dtype=numpy.float64)


class Camera:
    """Abstract camera: implement ``V()``, ``P()`` and ``resolution()``."""

    def __init__(self) -> None:
        self.t = _translation([0.0, 0.0, 3.2])  # default pose, camera.py:19

    def V(self) -> torch.Tensor:
        """GL view matrix = inverse of the camera pose ``self.t`` (X right, Y up, -Z forward)."""
        return gpu_f32(self.V_host())

    def V_host(self) -> numpy.ndarray:
        """``V()`` as a host fp32 array: the session's set-up reads the matrices on the host, so that no device round trip stalls the
        stream between two single-use sessions (multi-view rendering)."""
        return numpy.linalg.inv(self.t).astype(numpy.float32)

    def P_host(self) -> numpy.ndarray:
        return self.P().detach().cpu().numpy().astype(numpy.float32)

    def P(self) -> torch.Tensor:
        raise NotImplementedError

    def resolution(self):
        raise NotImplementedError


class RawCamera(Camera):
    """Camera driven directly by view / projection tensors (camera.py:49-66)."""

    def __init__(self, h: int, w: int, v: torch.Tensor, p: torch.Tensor) -> None:
        self.h, self.w, self.v, self.p = h, w, v, p

    def V(self):
        return self.v

    def P(self):
        return self.p

    def V_host(self):
        return self.v.detach().cpu().numpy().astype(numpy.float32)

    def resolution(self):
        return self.h, self.w


class PerspectiveCamera(Camera):
    """Perspective camera; angles in degrees (camera.py:68-137)."""

    def __init__(self, fov=30, h=512, w=512, near=0.1, far=10.0) -> None:
        super().__init__()
        self.fov, self.h, self.w, self.near, self.far = fov, h, w, near, far

    def P(self):
        return gpu_f32(self.P_host())

    def P_host(self) -> numpy.ndarray:
        focal = self.h / (2.0 * math.tan(math.radians(self.fov) / 2.0))  # vertical fov -> focal length in pixels
        return gl_perspective(self.w, self.h, self.w / 2, self.h / 2, focal, focal, self.near, self.far).astype(numpy.float32)

    def resolution(self):
        return self.h, self.w

    def set_transform(self, tr: numpy.ndarray):
        self.t = tr

    def lookat(self, point: Union[List[float], numpy.ndarray]):
        point = numpy.asarray(point, dtype=numpy.float64)
        assert list(point.shape) == [3]
        fwd = point - self.t[:3, 3]
        right = numpy.cross(fwd, [0.0, 1.0, 0.0])
        up = numpy.cross(right, fwd)
        self.t[:3, 0], self.t[:3, 1], self.t[:3, 2] = _unit(right), _unit(up), _unit(-fwd)
        return self

    @classmethod
    def from_orbit(cls, h, w, radius, azim, elev, origin, fov=30, near=0.1, far=10.0):
        """Orbit camera: azimuth 0 looks from +Z, elevation 90 looks down (camera.py:109-137)."""
        cam = cls(h=h, w=w, fov=fov, near=near, far=far)
        theta, phi = math.radians(azim), math.radians(elev)
        pos = [radius * math.sin(theta) * math.cos(phi) + origin[0], radius * math.sin(phi) + origin[1],
               radius * math.cos(theta) * math.cos(phi) + origin[2]]
        cam.t = _translation(pos)
        cam.lookat(origin)
        return cam
