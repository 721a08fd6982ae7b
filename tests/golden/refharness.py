"""
Harness that makes the *unmodified-in-place* Python reference importable in the build
container so golden vectors can be generated from it (tests/golden/make_golden.py).

TEST INFRASTRUCTURE ONLY.  Nothing in the product (diffrp_b200/), bench.py or the -m gpu
tests imports this file; /root/reference does not exist on the GPU box.

What it does:
  * copies /root/reference/diffrp into a scratch dir under /tmp (never into the repo),
  * applies the two documented patches on that copy:
      A. hard-coded 'cuda' device strings -> 'cpu'   (shader_ops.py:35-49, light_transport.py:201,
         mixin.py:48,78, path_tracing.py:310)  -- only so it runs without a GPU;
      B. `.int()` on the primitive index in layer_material_rays (path_tracing.py:163): as shipped the two
         torch raycasters return int64 ids, which breaks triidx_to_float (interpolator.py:28-29);
  * trims the package __init__ files so loaders / rasterizer (trimesh, nvdiffrast) are not imported,
  * puts ~40 lines of shims on sys.path for third-party modules that are absent in this image
    (torch_redstone.supercat, calibur.{fov_to_focal,projection_gl_persp,normalized},
    trimesh.transformations.{translation_matrix,inverse_matrix,translation_from_matrix},
    empty nvdiffrast.torch / pyexr).
"""
import os
import sys
import shutil
import tempfile
import importlib

REFERENCE_ROOT = "/root/reference"

_SHIMS = {
    "torch_redstone/__init__.py": '''
import torch
def supercat(tensors, dim=0):
    tensors = [t if isinstance(t, torch.Tensor) else torch.as_tensor(t) for t in tensors]
    nd = max(t.ndim for t in tensors)
    tensors = [t.reshape((1,) * (nd - t.ndim) + tuple(t.shape)) for t in tensors]
    d = dim % nd
    shape = [max(t.shape[k] for t in tensors) for k in range(nd)]
    out = []
    for t in tensors:
        s = list(shape); s[d] = t.shape[d]
        out.append(t.expand(*s))
    return torch.cat(out, dim=d)
def torch_to_numpy(t):
    return t.detach().cpu().numpy()
''',
    "calibur/__init__.py": '''
import numpy
def fov_to_focal(fov, size):
    return size / (2.0 * numpy.tan(fov / 2.0))
def projection_gl_persp(width, height, cx, cy, fx, fy, near, far):
    return numpy.array([
        [2.0 * fx / width, 0.0, 1.0 - 2.0 * cx / width, 0.0],
        [0.0, 2.0 * fy / height, 2.0 * cy / height - 1.0, 0.0],
        [0.0, 0.0, (far + near) / (near - far), 2.0 * far * near / (near - far)],
        [0.0, 0.0, -1.0, 0.0]], dtype=numpy.float64)
def normalized(x):
    x = numpy.asarray(x, dtype=numpy.float64)
    return x / numpy.linalg.norm(x, axis=-1, keepdims=True)
''',
    "trimesh/__init__.py": "from . import transformations\n",
    "trimesh/transformations.py": '''
import numpy
def translation_matrix(d):
    m = numpy.identity(4); m[:3, 3] = numpy.asarray(d, dtype=numpy.float64)[:3]; return m
def inverse_matrix(m):
    return numpy.linalg.inv(m)
def translation_from_matrix(m):
    return numpy.array(m, copy=False)[:3, 3].copy()
''',
    "nvdiffrast/__init__.py": "",
    "nvdiffrast/torch.py": "",
    "pyexr/__init__.py": "",
}

_TRIMMED_INIT = {
    "diffrp/__init__.py": (
        "from .rendering import *\nfrom .materials import *\nfrom .scene import *\n"
        "from .utils import *\nfrom .utils.cache import *\nfrom .version import __version__\n"),
    "diffrp/rendering/__init__.py": (
        "from .camera import Camera, RawCamera, PerspectiveCamera\n"
        "from .interpolator import Interpolator, MaskedSparseInterpolator, polyfill_interpolate\n"
        "from .mixin import RenderSessionMixin\n"
        "from .path_tracing import PathTracingSession, PathTracingSessionOptions, RayOutputs\n"),
}

_CUDA_FILES = ["diffrp/utils/shader_ops.py", "diffrp/utils/light_transport.py",
               "diffrp/rendering/mixin.py", "diffrp/rendering/path_tracing.py"]


def _patch(path, old, new, count_min=1):
    with open(path) as fi:
        src = fi.read()
    assert src.count(old) >= count_min, (path, old)
    with open(path, "w") as fo:
        fo.write(src.replace(old, new))


def load_reference(device: str = "cpu"):
    """Returns the imported (patched-copy) `diffrp` module of the reference."""
    if "diffrp" in sys.modules and getattr(sys.modules["diffrp"], "_b200_refharness", False):
        return sys.modules["diffrp"]
    if not os.path.isdir(REFERENCE_ROOT):
        raise RuntimeError("reference tree not present (expected only in the build container)")
    root = tempfile.mkdtemp(prefix="diffrp_ref_")
    shutil.copytree(os.path.join(REFERENCE_ROOT, "diffrp"), os.path.join(root, "diffrp"),
                    ignore=shutil.ignore_patterns("__pycache__", "*.so"))
    for rel, text in _SHIMS.items():
        p = os.path.join(root, "shims", rel)
        os.makedirs(os.path.dirname(p), exist_ok=True)
        with open(p, "w") as fo:
            fo.write(text)
    for rel, text in _TRIMMED_INIT.items():
        with open(os.path.join(root, rel), "w") as fo:
            fo.write(text)
    if device != "cuda":
        for rel in _CUDA_FILES:
            p = os.path.join(root, rel)
            _patch(p, "'cuda'", repr(device))
    # Patch B: int32 ids (reference bug)
    _patch(os.path.join(root, "diffrp/rendering/path_tracing.py"),
           "i = torch.where(t < far, i + 1, 0)", "i = torch.where(t < far, i + 1, 0).int()")
    sys.path.insert(0, os.path.join(root, "shims"))
    sys.path.insert(0, root)
    mod = importlib.import_module("diffrp")
    mod._b200_refharness = True
    mod._b200_root = root
    return mod
