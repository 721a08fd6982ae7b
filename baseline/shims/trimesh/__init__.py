"""Shim: trimesh.transformations subset used by rendering/camera.py:19,32,100,135.  The glTF loader (which needs the real trimesh)
is not importable with this shim; baseline/ref_loader.py stubs `diffrp.loaders`."""
from . import transformations  # noqa: F401
