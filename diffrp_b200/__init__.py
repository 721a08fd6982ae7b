"""
diffrp_b200 -- a B200-native (sm_100a CUDA) implementation of diffrp's path-tracing hot path behind diffrp's own
Python API: ``Scene`` / ``MeshObject``, the ``Raycaster`` seam, ``PathTracingSession(Options)`` and the
``(radiance, alpha, extras)`` outputs.  See DESIGN.md and INTEGRATION.md.
"""
from .camera import Camera, RawCamera, PerspectiveCamera
from .scene import Scene, MeshObject, Light, DirectionalLight, PointLight, ImageEnvironmentLight
from .materials import SurfaceMaterial, SurfaceOutputStandard, DefaultMaterial, GLTFMaterial, GLTFSampler
from .raycaster import Raycaster, B200Raycaster
from .path_tracing import PathTracingSession, PathTracingSessionOptions, RayOutputs, hammersley
from .flatten import VertexArrayObject
from .generic import SurfaceInput, SurfaceUniform, MaskedSparseInterpolator
from . import synthetic, tonemap, denoiser
from .tonemap import agx_base_contrast, linear_to_srgb, to_uint8
from .denoiser import get_denoiser, run_denoiser

__version__ = "0.1.0"
