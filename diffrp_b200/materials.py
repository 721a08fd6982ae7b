"""
Material protocol of diffrp (reference: diffrp/materials/base_material.py:300-382, default_material.py,
gltf_material.py) and the two built-in materials the fused CUDA shade kernel understands.

``SurfaceMaterial.shade(su, si) -> SurfaceOutputStandard`` stays the plugin seam for arbitrary user materials
(evaluated in PyTorch by ``PathTracingSession``'s generic path); ``DefaultMaterial`` and ``GLTFMaterial`` additionally
implement ``fused_description()`` which is what the fused kernel consumes (drp_material_t).
"""
import abc
from dataclasses import dataclass
from typing import Dict, Optional

import torch

from .ops import sample2d, ones_like_vec


@dataclass
class SurfaceOutputStandard:
    """Standard outputs of a material; every field optional (base_material.py:300-363)."""
    albedo: Optional[torch.Tensor] = None      # (...,3) default magenta
    normal: Optional[torch.Tensor] = None      # (...,3) in `normal_space`, default geometry normal
    emission: Optional[torch.Tensor] = None    # (...,3) default 0
    metallic: Optional[torch.Tensor] = None    # (...,1) default 0
    smoothness: Optional[torch.Tensor] = None  # (...,1) default 0.5
    occlusion: Optional[torch.Tensor] = None   # (...,1) default 1
    alpha: Optional[torch.Tensor] = None       # (...,1) default 1
    aovs: Optional[Dict[str, torch.Tensor]] = None
    normal_space: str = 'tangent'


class SurfaceMaterial(metaclass=abc.ABCMeta):
    @abc.abstractmethod
    def shade(self, su, si) -> SurfaceOutputStandard:
        """Batch-trivial shading function (base_material.py:366-382)."""
        raise NotImplementedError

    def fused_description(self) -> Optional[dict]:
        """Description consumed by the fused CUDA kernel, or None when only the Python ``shade`` exists."""
        return None


class DefaultMaterial(SurfaceMaterial):
    """albedo = interpolated vertex colour (x optional tint); everything else default (default_material.py:6-24)."""

    def __init__(self, tint: Optional[torch.Tensor] = None) -> None:
        super().__init__()
        self.tint = tint

    def shade(self, su, si) -> SurfaceOutputStandard:
        rgb = si.color[..., :3]
        return SurfaceOutputStandard(rgb * torch.as_tensor(self.tint).to(rgb) if self.tint is not None else rgb)

    def fused_description(self) -> Optional[dict]:
        if type(self) is not DefaultMaterial:
            return None  # a subclass may override shade()
        tint = None if self.tint is None else [float(x) for x in torch.as_tensor(self.tint).detach().cpu().reshape(-1)[:3]]
        return dict(kind='default', tint=tint)


@dataclass
class GLTFSampler:
    """Texture + sampling state (gltf_material.py:9-22)."""
    image: torch.Tensor  # (H, W, 1|3|4)
    wrap_mode: str = 'repeat'      # 'repeat' | 'clamp' | 'mirror'
    interpolation: str = 'linear'  # 'point' | 'linear'

    def sample(self, uv: torch.Tensor) -> torch.Tensor:
        if self.wrap_mode == 'repeat':
            uv = uv.remainder(1.0)
        return sample2d(self.image.to(uv), uv, wrap='border' if self.wrap_mode == 'clamp' else 'reflection',
                        mode='bilinear' if self.interpolation == 'linear' else 'nearest')

    def fused_description(self) -> dict:
        return dict(image=self.image, wrap=self.wrap_mode, interp=self.interpolation)


@dataclass
class GLTFMaterial(SurfaceMaterial):
    """glTF 2.0 metallic-roughness material (gltf_material.py:25-67)."""
    base_color_factor: torch.Tensor
    base_color_texture: GLTFSampler
    metallic_factor: float
    roughness_factor: float
    metallic_roughness_texture: GLTFSampler
    normal_texture: Optional[GLTFSampler]
    occlusion_texture: Optional[GLTFSampler]
    emissive_factor: Optional[torch.Tensor]
    emissive_texture: GLTFSampler
    alpha_cutoff: float
    alpha_mode: str  # 'OPAQUE' | 'MASK' | 'BLEND'

    def shade(self, su, si) -> SurfaceOutputStandard:
        uv = si.uv
        rgba = torch.as_tensor(self.base_color_factor).to(uv) * si.color * self.base_color_texture.sample(uv)
        mr = self.metallic_roughness_texture.sample(uv)
        if self.alpha_mode == 'OPAQUE':
            alpha = None
        elif self.alpha_mode == 'MASK':
            alpha = (rgba[..., 3:4] > self.alpha_cutoff).float()
        elif self.alpha_mode == 'BLEND':
            alpha = rgba[..., 3:4]
        else:
            raise ValueError('bad GLTFMaterial.alpha_mode', self.alpha_mode)
        return SurfaceOutputStandard(
            albedo=rgba[..., :3],
            normal=torch.add(-1, self.normal_texture.sample(uv), alpha=2) if self.normal_texture is not None else None,
            emission=torch.as_tensor(self.emissive_factor).to(uv) * self.emissive_texture.sample(uv) if self.emissive_factor is not None else None,
            metallic=self.metallic_factor * mr[..., 2:3],
            smoothness=torch.add(1.0, mr[..., 1:2], alpha=-self.roughness_factor),
            occlusion=self.occlusion_texture.sample(uv)[..., 0:1] if self.occlusion_texture is not None else ones_like_vec(uv, 1),
            alpha=alpha,
        )

    def fused_description(self) -> Optional[dict]:
        if type(self) is not GLTFMaterial:
            return None
        if self.alpha_mode not in ('OPAQUE', 'MASK', 'BLEND'):
            raise ValueError('bad GLTFMaterial.alpha_mode', self.alpha_mode)
        f = lambda t: [float(x) for x in torch.as_tensor(t).detach().cpu().reshape(-1)]
        return dict(
            kind='gltf', alpha_mode=self.alpha_mode, alpha_cutoff=float(self.alpha_cutoff),
            base_color_factor=f(self.base_color_factor), metallic_factor=float(self.metallic_factor),
            roughness_factor=float(self.roughness_factor),
            emissive_factor=None if self.emissive_factor is None else f(self.emissive_factor)[:3],
            base_color_tex=self.base_color_texture.fused_description(),
            mr_tex=self.metallic_roughness_texture.fused_description(),
            normal_tex=None if self.normal_texture is None else self.normal_texture.fused_description(),
            emissive_tex=self.emissive_texture.fused_description(),
        )
