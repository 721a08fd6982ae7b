"""
Import the REFERENCE (eliphatfs/diffrp, pure Python) from ``baseline/_ref`` -- the unmodified files ``pip install --no-deps --target``
put there (baseline/make_ref.py) -- or, in the build container, straight from ``/root/reference``.

TEST / MEASUREMENT INFRASTRUCTURE ONLY: imported by ``tests/`` (drop-in and parity tests), ``tests/golden/make_golden.py`` and the
reference legs of ``bench.py``.  Nothing under ``diffrp_b200/`` imports it.

The files on disk are never edited.  What has to differ for the reference to run in this image is applied IN MEMORY by an import hook
(a ``SourceLoader`` without bytecode cache that rewrites the source text of four modules while it is being imported) and by stub modules:

* third-party modules that are not installed here (no network): ``torch_redstone``, ``calibur``, ``trimesh``, ``nvdiffrast``, ``pyexr`` ->
  ``baseline/shims`` (own code, appended to ``sys.path`` so that a real installation wins when there is one);
* ``diffrp/loaders/__init__.py`` (glTF import, needs the real trimesh; not on the path-tracing path) is replaced, in memory, by the
  first-party imports of ``loaders/gltf_loader.py:10-14`` alone, so that the package's (circular) import order stays the shipped one;
* **patch A**, only for ``device='cpu'``: the hard-coded ``'cuda'`` device strings (utils/shader_ops.py:35-49, utils/light_transport.py:201,
  rendering/mixin.py:48,78, rendering/path_tracing.py:310) become ``'cpu'``;
* **patch B**, unless ``patch_int32=False``: ``.int()`` on the 1-based primitive id in ``layer_material_rays`` (rendering/path_tracing.py:163).
  As shipped, the two torch raycasters return int64 ids and ``triidx_to_float`` (rendering/interpolator.py:28-29) then fails its float4
  assertion: without the patch only ``raycaster_impl='torchoptix'`` runs.  With ``torchoptix`` provided by ``diffrp_b200.optix_compat`` the
  reference runs with ``patch_int32=False``, i.e. with no source change at all.
"""
import importlib
import importlib.abc
import importlib.machinery
import importlib.util
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
INSTALLED_ROOT = os.path.join(HERE, "_ref")
CONTAINER_ROOT = "/root/reference"
SHIMS = os.path.join(HERE, "shims")

_CUDA_FILES = ("diffrp/utils/shader_ops.py", "diffrp/utils/light_transport.py", "diffrp/rendering/mixin.py", "diffrp/rendering/path_tracing.py")
_INT32_FILE = "diffrp/rendering/path_tracing.py"
_LOADERS_FILE = "diffrp/loaders/__init__.py"
_LOADERS_STUB = ("from ..utils import colors\nfrom ..utils.shader_ops import *\nfrom ..scene import Scene, MeshObject\n"
                 "from ..rendering.camera import PerspectiveCamera\nfrom ..materials.gltf_material import GLTFMaterial, GLTFSampler\n__all__ = []\n")
_INT32_OLD, _INT32_NEW = "i = torch.where(t < far, i + 1, 0)", "i = torch.where(t < far, i + 1, 0).int()"

_state = {"root": None, "device": None, "patch_int32": None, "finder": None}


def reference_root(prefer_installed: bool = True):
    """Directory that contains the reference's ``diffrp`` package, or None."""
    cands = [INSTALLED_ROOT, CONTAINER_ROOT] if prefer_installed else [CONTAINER_ROOT, INSTALLED_ROOT]
    for c in cands:
        if os.path.isfile(os.path.join(c, "diffrp", "rendering", "path_tracing.py")):
            return c
    return None


class _PatchingLoader(importlib.abc.SourceLoader):
    """Source loader without a bytecode cache (no ``path_stats``), so the rewritten text is what gets compiled -- and what
    ``inspect.getsource`` / TorchScript see through ``get_source``."""

    def __init__(self, fullname, path, root, device, patch_int32, stub_loaders):
        self.fullname, self.path, self.root, self.device, self.patch_int32 = fullname, path, root, device, patch_int32
        self.stub_loaders = stub_loaders

    def get_filename(self, fullname):
        return self.path

    def get_data(self, path):
        with open(path, "rb") as fi:
            data = fi.read()
        rel = os.path.relpath(path, self.root).replace(os.sep, "/")
        if self.stub_loaders and rel == _LOADERS_FILE:
            return _LOADERS_STUB.encode()
        if self.device != "cuda" and rel in _CUDA_FILES:
            text = data.decode()
            assert "'cuda'" in text, rel
            data = text.replace("'cuda'", repr(self.device)).encode()
        if self.patch_int32 and rel == _INT32_FILE:
            text = data.decode()
            assert text.count(_INT32_OLD) == 1, "patch B does not apply: reference version changed?"
            data = text.replace(_INT32_OLD, _INT32_NEW).encode()
        return data


class _Finder(importlib.abc.MetaPathFinder):
    def __init__(self, root, device, patch_int32):
        self.root, self.device, self.patch_int32 = root, device, patch_int32
        self.stub_loaders = False

    def find_spec(self, fullname, path=None, target=None):
        if fullname != "diffrp" and not fullname.startswith("diffrp."):
            return None
        search = [self.root] if fullname == "diffrp" else path
        spec = importlib.machinery.PathFinder.find_spec(fullname, search)
        if spec is None or not spec.origin or not spec.origin.endswith(".py"):
            return spec
        loader = _PatchingLoader(fullname, spec.origin, self.root, self.device, self.patch_int32, self.stub_loaders)
        return importlib.util.spec_from_file_location(fullname, spec.origin, loader=loader,
                                                      submodule_search_locations=spec.submodule_search_locations)


def _have(module: str) -> bool:
    try:
        return importlib.util.find_spec(module) is not None
    except (ImportError, ValueError):
        return False


def unload_reference():
    """Forget an imported reference (to import it again for another device)."""
    for name in [n for n in sys.modules if n == "diffrp" or n.startswith("diffrp.")]:
        del sys.modules[name]
    if _state["finder"] is not None and _state["finder"] in sys.meta_path:
        sys.meta_path.remove(_state["finder"])
    _state.update(root=None, device=None, patch_int32=None, finder=None)


def load_reference(device: str = "cuda", root: str = None, patch_int32: bool = True):
    """Returns the reference's ``diffrp`` module.  ``device``: 'cuda' (sources as shipped) or 'cpu' (patch A)."""
    root = root or reference_root()
    if root is None:
        raise RuntimeError("reference not found: run `python baseline/make_ref.py` in the build container (installs it into baseline/_ref)")
    if "diffrp" in sys.modules:
        if (_state["root"], _state["device"], _state["patch_int32"]) == (root, device, patch_int32):
            return sys.modules["diffrp"]
        if _state["root"] is None:
            raise RuntimeError("a foreign `diffrp` module is already imported")
        unload_reference()
    if SHIMS not in sys.path:
        sys.path.append(SHIMS)  # appended: a real installation of any of these modules wins
    finder = _Finder(root, device, patch_int32)
    sys.meta_path.insert(0, finder)
    _state.update(root=root, device=device, patch_int32=patch_int32, finder=finder)
    trimesh = importlib.import_module("trimesh")
    finder.stub_loaders = not hasattr(trimesh, "Trimesh")  # the shim: the glTF loader cannot be imported, and is not on the path
    mod = importlib.import_module("diffrp")
    mod._b200_ref_root, mod._b200_ref_device = root, device
    return mod
