"""
Small torch helpers used by the host-side mirror of diffrp's interfaces (subset of the reference's
``diffrp.utils.shader_ops`` that the path-tracing seam needs).  Plumbing only: the hot path is CUDA.
"""
import numpy
import torch
import torch.nn.functional as F

_DEFAULT_DEVICE = None


def default_device() -> torch.device:
    """'cuda' when a GPU is present (the reference hard-codes 'cuda', shader_ops.py:35-49); CPU only for host-logic tests."""
    if _DEFAULT_DEVICE is not None:
        return _DEFAULT_DEVICE
    return torch.device('cuda') if torch.cuda.is_available() else torch.device('cpu')


def set_default_device(device):
    global _DEFAULT_DEVICE
    _DEFAULT_DEVICE = None if device is None else torch.device(device)


def gpu_f32(inputs) -> torch.Tensor:
    if isinstance(inputs, torch.Tensor):
        return inputs.to(dtype=torch.float32, device=default_device())
    if isinstance(inputs, (float, int)):
        return torch.full([], inputs, dtype=torch.float32, device=default_device())
    return torch.tensor(numpy.asarray(inputs), dtype=torch.float32, device=default_device())


def gpu_i32(inputs) -> torch.Tensor:
    if isinstance(inputs, torch.Tensor):
        return inputs.to(dtype=torch.int32, device=default_device())
    return torch.tensor(numpy.asarray(inputs), dtype=torch.int32, device=default_device())


def normalized(x: torch.Tensor) -> torch.Tensor:
    return F.normalize(x, dim=-1)


def transform_point4x3(xyz: torch.Tensor, matrix: torch.Tensor) -> torch.Tensor:
    """Affine transform of points (B,3) by a (4,4) matrix."""
    return torch.addmm(matrix[:-1, -1], xyz, matrix[:-1, :-1].T)


def transform_vector3x3(xyz: torch.Tensor, matrix: torch.Tensor) -> torch.Tensor:
    return torch.matmul(xyz, matrix[:-1, :-1].T)


def zeros_like_vec(x: torch.Tensor, c: int) -> torch.Tensor:
    return x.new_zeros(*x.shape[:-1], c)


def ones_like_vec(x: torch.Tensor, c: int) -> torch.Tensor:
    return x.new_ones(*x.shape[:-1], c)


def full_like_vec(x: torch.Tensor, value, c: int) -> torch.Tensor:
    return x.new_full([*x.shape[:-1], c], value)


def saturate(x: torch.Tensor) -> torch.Tensor:
    return torch.clamp(x, 0.0, 1.0)


def dot(a: torch.Tensor, b: torch.Tensor, keepdim: bool = True) -> torch.Tensor:
    r = torch.linalg.vecdot(a, b)
    return r.unsqueeze(-1) if keepdim else r


def cross(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    return torch.linalg.cross(a, b)


def small_matrix_inverse(x: torch.Tensor) -> torch.Tensor:
    """Inverse of a tiny (batch of) matrix on the host with numpy, like shader_ops.py:578-596 (fp32 LAPACK)."""
    return x.new_tensor(numpy.linalg.inv(x.detach().cpu().numpy()))


def sample2d(texture2d: torch.Tensor, texcoords: torch.Tensor, wrap: str = "border", mode: str = "bilinear") -> torch.Tensor:
    """(H,W,C) texture sampled at (...,2) uv, (0,0) = bottom-left: shader_ops.py:198-255 (non-'cyclic'/'latlong' wraps)."""
    shape = texcoords.shape
    grid = texcoords.reshape(1, 1, -1, 2) * 2 - 1
    grid = grid * grid.new_tensor([1.0, -1.0])
    out = F.grid_sample(texture2d[None].permute(0, 3, 1, 2), grid, padding_mode=wrap, mode=mode, align_corners=False)
    return out.view(texture2d.shape[-1], -1).T.reshape(*shape[:-1], texture2d.shape[-1])
