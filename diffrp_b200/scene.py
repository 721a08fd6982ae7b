"""
Scene description with the interface of the reference's ``diffrp.scene`` (scene.py:10-75, objects.py:11-98,
lights.py:6-50): ``Scene``, ``MeshObject`` and the light dataclasses.  These are the *inputs* of the hot path.
"""
import collections
from dataclasses import dataclass
from typing import Any, Dict, List, Optional, Union

import torch

from .ops import normalized, transform_point4x3, transform_vector3x3, zeros_like_vec, ones_like_vec


def face_normals(verts: torch.Tensor, tris: torch.Tensor, normalize: bool = False) -> torch.Tensor:
    """Per-face normals of CCW triangles (geometry.py:26-41)."""
    t = tris.long()
    n = torch.linalg.cross(verts[t[:, 1]] - verts[t[:, 0]], verts[t[:, 2]] - verts[t[:, 0]])
    return normalized(n) if normalize else n


def vertex_normals(verts: torch.Tensor, tris: torch.Tensor) -> torch.Tensor:
    """Area-weighted smooth vertex normals (geometry.py:44-65)."""
    fn = face_normals(verts, tris)
    t = tris.long()
    acc = torch.zeros_like(verts)
    for k in range(3):
        acc = acc.index_add(0, t[:, k], fn)
    return normalized(acc)


@dataclass
class MeshObject:
    """Triangle mesh + material + per-vertex attributes; same fields and defaults as objects.py:11-69."""
    material: Any
    verts: torch.Tensor
    tris: torch.Tensor
    normals: Union[str, torch.Tensor] = 'flat'
    M: Optional[torch.Tensor] = None
    color: Optional[torch.Tensor] = None
    uv: Optional[torch.Tensor] = None
    tangents: Optional[torch.Tensor] = None
    custom_attrs: Optional[Dict[str, torch.Tensor]] = None
    metadata: Optional[Dict[str, Any]] = None

    def to(self, device, memo: Optional[dict] = None):
        """Copy of this object with every tensor on ``device`` (the material object is shared).  ``memo`` (id of the source tensor -> moved
        tensor) keeps tensors that several objects share -- the mesh of instanced objects -- shared after the move."""
        def mv(x):
            if not isinstance(x, torch.Tensor):
                return x
            if memo is None:
                return x.to(device)
            if id(x) not in memo:
                memo[id(x)] = (x, x.to(device))   # the source is kept alive so that its id stays unique
            return memo[id(x)][1]
        return MeshObject(self.material, mv(self.verts), mv(self.tris), mv(self.normals), mv(self.M), mv(self.color), mv(self.uv),
                          mv(self.tangents), None if self.custom_attrs is None else {k: mv(v) for k, v in self.custom_attrs.items()},
                          self.metadata)

    def preprocess(self):
        """Fill defaults; 'flat' normals turn the mesh into a face soup (objects.py:71-98)."""
        v = self.verts
        if self.M is None:
            self.M = torch.eye(4, device=v.device, dtype=v.dtype)
        if self.color is None:
            self.color = ones_like_vec(v, 4)
        if self.uv is None:
            self.uv = zeros_like_vec(v, 2)
        if self.tangents is None:
            self.tangents = zeros_like_vec(v, 4)
        if self.custom_attrs is None:
            self.custom_attrs = {}
        if self.metadata is None:
            self.metadata = {}
        if isinstance(self.normals, str):
            if self.normals == 'smooth':
                self.normals = vertex_normals(self.verts, self.tris)
            elif self.normals == 'flat':
                f = self.tris.long()
                fn = face_normals(self.verts, self.tris, normalize=True)
                self.verts = self.verts[f].reshape(-1, 3)
                self.normals = fn[:, None, :].expand(-1, 3, -1).reshape(-1, 3)
                self.tris = torch.arange(f.numel(), device=f.device, dtype=torch.int32).reshape(f.shape)
                self.color = self.color[f].flatten(0, 1)
                self.uv = self.uv[f].flatten(0, 1)
                self.tangents = self.tangents[f].flatten(0, 1)
                self.custom_attrs = {k: a[f].flatten(0, 1) for k, a in self.custom_attrs.items()}
            else:
                raise ValueError("normals must be 'flat', 'smooth' or a tensor, got %r" % (self.normals,))
        return self


@dataclass
class Light:
    intensity: Union[torch.Tensor, float]
    color: torch.Tensor


@dataclass
class DirectionalLight(Light):
    """Data only (not supported by the path tracer; lights.py:20-25)."""
    direction: torch.Tensor


@dataclass
class PointLight(Light):
    """Data only (not supported by the path tracer; lights.py:28-33)."""
    position: torch.Tensor


@dataclass
class ImageEnvironmentLight(Light):
    """Lat-long (H, 2H, 3) environment image (lights.py:36-50)."""
    image: torch.Tensor
    render_skybox: bool = True

    def image_rh(self) -> torch.Tensor:
        return torch.fliplr(self.image) * (self.intensity * self.color)


class Scene:
    """Container of objects and lights; ``add_*`` return the scene for chaining (scene.py:10-75)."""

    def __init__(self) -> None:
        self.lights: List[Light] = []
        self.objects: List[MeshObject] = []
        self.metadata = {}

    def add_light(self, light: Light):
        self.lights.append(light)
        return self

    def add_mesh_object(self, mesh_obj: MeshObject):
        self.objects.append(mesh_obj.preprocess())
        return self

    def to(self, device):
        """Copy of the scene with geometry and light tensors on ``device`` (objects must already be preprocessed)."""
        out = Scene()
        memo: dict = {}
        out.objects = [o.to(device, memo) for o in self.objects]
        for l in self.lights:
            if isinstance(l, ImageEnvironmentLight):
                out.lights.append(ImageEnvironmentLight(l.intensity, l.color.to(device), l.image.to(device), l.render_skybox))
            else:
                out.lights.append(l)
        out.metadata = dict(self.metadata)
        return out

    def pin_memory(self, pin: bool = True):
        """
        Copy of a host scene whose tensors are views of ONE page-locked arena, laid out in the order the session uploads them (per object:
        verts, normals, color, uv, tangents, tris; then every material's textures; then the lights' images), each 256-byte aligned.
        ``PathTracingSession`` recognises the arena (``flatten.PackedUpload``) and moves the whole scene with a handful of large DMA copies
        instead of one per tensor (a scene is ~300 tensors; issuing 300 copies costs more host time than the transfer takes).
        Tensors shared between objects (instanced meshes) stay shared.  Model matrices and material factors are left where they are.
        """
        import copy
        ALIGN = 256
        order, seen = [], {}

        def want(t):
            if isinstance(t, torch.Tensor) and not t.is_cuda and id(t) not in seen:
                seen[id(t)] = len(order)
                order.append(t)
        tex_fields = ('base_color_texture', 'metallic_roughness_texture', 'normal_texture', 'emissive_texture', 'occlusion_texture')
        for o in self.objects:
            for t in (o.verts, o.normals, o.color, o.uv, o.tangents, o.tris):
                want(t)
        mats = []
        for o in self.objects:
            if id(o.material) not in [id(m) for m in mats]:
                mats.append(o.material)
        for m in mats:
            for f in tex_fields:
                smp = getattr(m, f, None)
                if smp is not None and hasattr(smp, 'image'):
                    want(smp.image)
        for l in self.lights:
            want(getattr(l, 'image', None))
        offs, total = [], 0
        for t in order:
            offs.append(total)
            total += -(-(t.numel() * t.element_size()) // ALIGN) * ALIGN
        arena = torch.empty([max(total, ALIGN)], dtype=torch.uint8)
        if pin:   # (False: the same arena in pageable memory -- layout tests on machines without a GPU)
            arena = arena.pin_memory()
        views = []
        for t, off in zip(order, offs):
            v = arena[off:off + t.numel() * t.element_size()].view(t.dtype).view(t.shape)
            v.copy_(t)
            views.append(v)
        mv = lambda t: views[seen[id(t)]] if isinstance(t, torch.Tensor) and id(t) in seen else t   # noqa: E731
        new_mats = {}
        for m in mats:
            nm = copy.copy(m)
            for f in tex_fields:
                smp = getattr(m, f, None)
                if smp is not None and hasattr(smp, 'image'):
                    ns = copy.copy(smp)
                    ns.image = mv(smp.image)
                    setattr(nm, f, ns)
            new_mats[id(m)] = nm
        out = Scene()
        out.objects = [MeshObject(new_mats[id(o.material)], mv(o.verts), mv(o.tris), mv(o.normals), o.M, mv(o.color), mv(o.uv), mv(o.tangents),
                                  o.custom_attrs, o.metadata) for o in self.objects]
        for l in self.lights:
            if isinstance(l, ImageEnvironmentLight):
                out.lights.append(ImageEnvironmentLight(l.intensity, l.color, mv(l.image), l.render_skybox))
            else:
                out.lights.append(l)
        out.metadata = dict(self.metadata)
        out._arena = arena
        return out

    def static_batching(self):
        """Merge meshes that share a material object into one world-space mesh, in place (scene.py:33-75)."""
        groups = collections.OrderedDict()
        for mesh in self.objects:
            groups.setdefault(id(mesh.material), []).append(mesh)
        merged = []
        for meshes in groups.values():
            if len(meshes) == 1:
                merged.append(meshes[0])
                continue
            tris, offset = [], 0
            for m in meshes:
                tris.append(m.tris + offset)
                offset += len(m.verts)
            last = meshes[-1]
            merged.append(MeshObject(
                last.material,
                torch.cat([transform_point4x3(m.verts, m.M) for m in meshes]),
                torch.cat(tris),
                torch.cat([normalized(transform_vector3x3(m.normals, m.M)) for m in meshes]),
                torch.eye(4, dtype=torch.float32, device=last.verts.device),
                torch.cat([m.color for m in meshes]),
                torch.cat([m.uv for m in meshes]),
                torch.cat([torch.cat([normalized(transform_vector3x3(m.tangents[..., :3], m.M)), m.tangents[..., 3:]], -1) for m in meshes]),
                {k: torch.cat([m.custom_attrs[k] for m in meshes]) for k in last.custom_attrs},
                {},
            ))
        self.objects = merged
        return self
