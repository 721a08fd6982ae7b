"""
Generic (plugin) path of ``PathTracingSession``: user samplers and arbitrary Python materials.

The reference's protocols are kept (diffrp/rendering/path_tracing.py:281-309 sampler protocol,
diffrp/materials/base_material.py:366-382 material protocol): per bounce the sampler receives
``(rays_o, rays_d, t, i, d)`` and returns ``RayOutputs``; materials receive ``(SurfaceUniform, SurfaceInput)``.
Only the intersection runs in hand-written CUDA here (``B200Raycaster.query``); the per-bounce tensor algebra below is
PyTorch glue for code the fused kernel cannot see.  Built-in materials never come through this file from ``pbr()``.
"""
import math
from collections.abc import Mapping
from typing import Callable, List, Tuple

import torch

from .ops import (normalized, transform_point4x3, transform_vector3x3, zeros_like_vec, ones_like_vec, full_like_vec, saturate,
                  dot, cross, sample2d, small_matrix_inverse)
from .materials import SurfaceOutputStandard


class SurfaceUniform:
    """Per-object uniforms handed to ``shade`` (base_material.py:37-76)."""

    def __init__(self, M: torch.Tensor, V: torch.Tensor, P: torch.Tensor):
        self.M, self.V, self.P = M, V, P
        self._cam = None

    @property
    def camera_matrix(self):
        if self._cam is None:
            self._cam = small_matrix_inverse(self.V)
        return self._cam

    @property
    def camera_position(self):
        return self.camera_matrix[:3, 3]


class MaskedSparseInterpolator:
    """Barycentric attribute interpolation for the rays selected by ``mask`` (interpolator.py:87-97)."""

    def __init__(self, vi_data: torch.Tensor, tris: torch.Tensor, mask: torch.Tensor):
        self.tris = tris
        self.indices = mask.nonzero(as_tuple=True)
        self.vi_data = vi_data[self.indices]  # (u, v, t, 1-based triangle id as float)
        self.tri_idx = tris[(float_to_triidx(self.vi_data[..., -1]) - 1).long()].long()

    def interpolate(self, vertex_buffer: torch.Tensor) -> torch.Tensor:
        corners = vertex_buffer[self.tri_idx]  # (n, 3, C)
        a, b, c = corners[:, 0], corners[:, 1], corners[:, 2]
        u, v = self.vi_data[:, 0:1], self.vi_data[:, 1:2]
        return torch.addcmul(torch.addcmul(c, b - c, v), a - c, u)  # (a-c)u + ((b-c)v + c)


class _CustomAttrs(Mapping):
    def __init__(self, si):
        self.si, self._c = si, {}

    def __getitem__(self, k):
        if k not in self._c:
            self._c[k] = self.si.interpolate_ex(self.si.vertex_buffers.custom_attrs[k])
        return self._c[k]

    def __len__(self):
        return len(self.si.vertex_buffers.custom_attrs)

    def __iter__(self):
        return iter(self.si.vertex_buffers.custom_attrs)


class SurfaceInput:
    """Lazily interpolated per-fragment inputs (base_material.py:79-276); attribute names match the reference."""

    def __init__(self, uniforms: SurfaceUniform, vertex_buffers, interpolator: MaskedSparseInterpolator):
        self.cache = {}
        self.uniforms = uniforms
        self.interpolator = interpolator
        self.vertex_buffers = vertex_buffers
        self.custom_surface_inputs = _CustomAttrs(self)

    def interpolate_ex(self, vertex_buffer: torch.Tensor, world_transform: str = 'none') -> torch.Tensor:
        M = self.uniforms.M
        if world_transform == 'point':
            vertex_buffer = transform_point4x3(vertex_buffer, M)
        elif world_transform == 'vector':
            vertex_buffer = transform_vector3x3(vertex_buffer, M)
        elif world_transform == 'vectornor':
            vertex_buffer = normalized(transform_vector3x3(vertex_buffer, M))
        elif world_transform == 'vector3ex1':
            vertex_buffer = torch.cat([transform_vector3x3(vertex_buffer[..., :3], M), vertex_buffer[..., 3:]], -1)
        elif world_transform == 'vector3norex1':
            vertex_buffer = torch.cat([normalized(transform_vector3x3(vertex_buffer[..., :3], M)), vertex_buffer[..., 3:]], -1)
        elif world_transform != 'none':
            raise ValueError("unknown world_transform: %r" % (world_transform,))
        return self.interpolator.interpolate(vertex_buffer)

    def _memo(self, key, fn):
        if key not in self.cache:
            self.cache[key] = fn()
        return self.cache[key]

    @property
    def world_pos(self):
        return self._memo('world_pos', lambda: self.interpolate_ex(self.vertex_buffers.world_pos))

    @property
    def local_pos(self):
        return self._memo('local_pos', lambda: self.interpolate_ex(self.vertex_buffers.verts))

    @property
    def view_dir(self):
        return self._memo('view_dir', lambda: normalized(self.world_pos - self.uniforms.camera_position))

    @property
    def world_normal_unnormalized(self):
        return self._memo('wnu', lambda: self.interpolate_ex(self.vertex_buffers.normals, 'vectornor'))

    @property
    def world_normal(self):
        return self._memo('wn', lambda: normalized(self.world_normal_unnormalized))

    @property
    def color(self):
        return self._memo('color', lambda: self.interpolate_ex(self.vertex_buffers.color))

    @property
    def uv(self):
        return self._memo('uv', lambda: self.interpolate_ex(self.vertex_buffers.uv))

    @property
    def world_tangent(self):
        return self._memo('wt', lambda: self.interpolate_ex(self.vertex_buffers.tangents, 'vector3norex1'))

    @property
    def custom_attrs(self):
        return self.custom_surface_inputs


def hit_barycentric(a, b, c, p):
    """Weights (u, v, w) of the point p in triangle (a, b, c), projected, nan -> 0, clipped (geometry.py:94-110)."""
    e0, e1, e2 = b - a, c - a, p - a
    d00, d01, d11 = (e0 * e0).sum(-1), (e0 * e1).sum(-1), (e1 * e1).sum(-1)
    d20, d21 = (e2 * e0).sum(-1), (e2 * e1).sum(-1)
    den = d00 * d11 - d01 * d01
    v = torch.clip(torch.nan_to_num((d11 * d20 - d01 * d21) / den), 0, 1)
    w = torch.clip(torch.nan_to_num((d00 * d21 - d01 * d20) / den), 0, 1)
    return torch.stack([1 - v - w, v, w], -1)


def triidx_to_float(ids: torch.Tensor) -> torch.Tensor:
    """1-based triangle id -> float channel of the raster record: exact float up to 2^24, bit-packed into the float's
    mantissa beyond (interpolator.py:28-29), so ids never lose precision in ``vi_data``."""
    ids = ids.int()
    return torch.where(ids <= 0x01000000, ids.float(), (ids + 0x4a800000).view(torch.float32))


def float_to_triidx(x: torch.Tensor) -> torch.Tensor:
    """Inverse of triidx_to_float (interpolator.py:24-25)."""
    return torch.where(x <= 16777216, x.int(), x.view(torch.int32) - 0x4a800000)


def layer_material_rays(sess, rays_o, rays_d, t, i) -> List[Tuple[SurfaceInput, SurfaceOutputStandard]]:
    """Evaluate every object's material on the rays that hit it (path_tracing.py:158-178)."""
    V, P, far = sess.camera_V(), sess.camera_P(), sess.camera_far()
    vao = sess.vertex_array_object()
    ids = torch.where(t < far, i.int() + 1, 0)  # 1-based, 0 = miss; int32 (the reference's int64 ids break here)
    corners = vao.world_pos[vao.tris[(ids - 1).long()].long()]
    bary = hit_barycentric(corners[:, 0], corners[:, 1], corners[:, 2], rays_o + rays_d * t[..., None])
    vi_data = torch.cat([bary[:, :2], t[..., None], triidx_to_float(ids[..., None])], -1)
    stencil = vao.stencils[ids.long()]
    mats = []
    for k, obj in enumerate(sess.scene.objects):
        su = SurfaceUniform(obj.M.to(vi_data.device), V, P)
        si = SurfaceInput(su, vao, MaskedSparseInterpolator(vi_data, vao.tris, stencil == k + 1))
        if len(si.interpolator.indices[0]) == 0:
            mats.append((si, SurfaceOutputStandard()))
        else:
            mats.append((si, obj.material.shade(su, si)))
    return mats


def world_normal_of(si: SurfaceInput, so: SurfaceOutputStandard) -> torch.Tensor:
    """Shading normal in world space from the material's normal output (mixin.py:115-128)."""
    if so.normal is None:
        return si.world_normal
    if so.normal_space == 'tangent':
        n, nt = si.world_normal_unnormalized, so.normal
        tg = si.world_tangent
        bit = tg[..., 3:] * cross(n, tg[..., :3])
        return normalized(nt[..., 0:1] * tg[..., :3] + (nt[..., 1:2] * bit + nt[..., 2:3] * n))
    if so.normal_space == 'object':
        return normalized(transform_vector3x3(so.normal, si.uniforms.M))
    if so.normal_space == 'world':
        return normalized(so.normal)
    raise ValueError("Unknown normal space: " + so.normal_space)


def surface_row(si: SurfaceInput, so: SurfaceOutputStandard) -> torch.Tensor:
    """[albedo3 | normal3 | metal | smooth | alpha | emission3] with the path tracer's defaults (path_tracing.py:180-187)."""
    n = world_normal_of(si, so)
    albedo = so.albedo if so.albedo is not None else n.new_tensor([1.0, 0.0, 1.0]).expand_as(n)
    metal = so.metallic if so.metallic is not None else zeros_like_vec(n, 1)
    smooth = so.smoothness if so.smoothness is not None else full_like_vec(n, 0.5, 1)
    alpha = so.alpha if so.alpha is not None else ones_like_vec(n, 1)
    emission = so.emission if so.emission is not None else zeros_like_vec(n, 3)
    return torch.cat([albedo, n, metal, smooth, alpha, emission], -1)


def collect_gbuffer(mats, operator: Callable, initial: torch.Tensor) -> torch.Tensor:
    """Scatter per-material rows into the (R, C) buffer; rays that hit nothing keep the initial value (mixin.py:130-155)."""
    idx, vals = [], []
    for si, so in mats:
        if len(si.interpolator.indices[0]) == 0:
            continue
        row = operator(si, so)
        if row is not None:
            idx.append(si.interpolator.indices[0])
            vals.append(row)
    if not idx:
        return initial
    return initial.index_put((torch.cat(idx),), torch.cat(vals))


def tangent_frame_combine(x, y, z, n):
    """x*right + z*up' + y*n with the fixed up-vector rule (light_transport.py:35-43)."""
    up = torch.where((n[..., 1:2] < 0.999), n.new_tensor([0.0, 1.0, 0.0]), n.new_tensor([1.0, 0.0, 0.0]))
    right = normalized(cross(up.expand_as(n), n))
    up2 = cross(n, right)
    return x * right + z * up2 + y * n


def brdf_sample_torch(attrs, t, rays_o, rays_d, env, u):
    """Tensor version of the metallic-roughness sampler (path_tracing.py:189-236); ``u`` is a list of six (R,1) uniforms."""
    albedo, n = attrs[..., 0:3], attrs[..., 3:6]
    metal, smooth, alpha, emission = attrs[..., 6:7], attrs[..., 7:8], attrs[..., 8:9], attrs[..., 9:12]
    hit_pos = rays_o + rays_d * t[..., None]
    diel = 1 - metal
    dc = diel * albedo
    dm = dc.max(-1, True).values
    p_diff = diel * dm / (0.04 + dm)
    p_spec = 1 - p_diff
    is_tr, is_di = u[0] >= alpha, u[1] >= p_spec
    # diffuse lobe
    theta = u[3] * math.tau
    xy = torch.sqrt(1.0 - u[2])
    d_di = tangent_frame_combine(xy * torch.cos(theta), u[2].sqrt(), xy * torch.sin(theta), n)
    t_di = dc / torch.clamp_min(p_diff, 0.0001)
    # specular lobe
    rough = (1 - smooth).clamp_min(1 / 512)
    a = rough * rough
    phi = math.tau * u[4]
    ct = torch.sqrt((1 - u[5]) / (1 + (a * a - 1) * u[5]))
    st = torch.sqrt(1 - ct * ct)
    h = tangent_frame_combine(torch.cos(phi) * st, ct, torch.sin(phi) * st, n)
    hd = dot(h, rays_d)
    d_sp = rays_d - 2.0 * hd * h
    vh = -hd
    k = a / 2.0
    g1 = lambda x: x / (x * (1.0 - k) + k)
    G = g1(torch.relu(dot(n, d_sp))) * g1(torch.relu(-dot(n, rays_d)))
    f0 = albedo * metal + 0.04 * diel
    F = torch.relu(smooth - f0) * (1.0 - vh) ** 5.0 + f0
    t_sp = F * G * torch.clamp_min(vh, 1e-6) / (torch.clamp_min(dot(n, h), 1e-6) * torch.clamp_min(-dot(n, rays_d), 1e-6))
    t_sp = t_sp / torch.clamp_min(p_spec * alpha, 0.0001)
    next_d = torch.where(is_tr, rays_d, torch.where(is_di, d_di, d_sp))
    transfer = torch.where(is_tr, torch.ones_like(t_sp), torch.where(is_di, t_di, t_sp))
    return albedo, emission, n, alpha, emission + env, transfer, hit_pos, next_d


def env_radiance(sess, rays_d):
    env = sess._single_env_light()
    if env is None:
        return zeros_like_vec(rays_d, 3)
    u = (torch.atan2(rays_d[..., 0:1], rays_d[..., 2:3]) * (0.5 / math.pi)) % 1
    v = (1 / math.pi) * torch.asin(torch.clamp(rays_d[..., 1:2], -0.999999, 0.999999)) + 0.5
    return sample2d(env, torch.cat([u, v], -1))


def sampler_brdf(sess, rays_o, rays_d, t, i, d: int):
    """The built-in PBR sampler on the generic path (path_tracing.py:250-279)."""
    from .path_tracing import RayOutputs
    far = sess.camera_far()
    mats = layer_material_rays(sess, rays_o, rays_d, t, i)
    always_sky = sess.options.ray_depth - 1 == d and sess.options.pbr_ray_last_bounce == 'skybox'
    attrs = collect_gbuffer(mats, surface_row, zeros_like_vec(rays_o, 12))
    hit = t[..., None] < far
    env = env_radiance(sess, rays_d)
    if not always_sky:
        env = torch.where(hit, torch.zeros_like(env), env)
    u = [torch.rand_like(attrs[..., :1]) for _ in range(6)]
    albedo, emission, n, alpha, radiance, transfer, hit_pos, next_d = brdf_sample_torch(attrs, t, rays_o, rays_d, env, u)
    return RayOutputs(radiance=radiance, transfer=torch.where(hit, transfer, torch.zeros_like(transfer)),
                      next_rays_o=hit_pos + next_d * sess.options.pbr_ray_step_epsilon, next_rays_d=next_d, alpha=alpha,
                      extras=dict(albedo=albedo, emission=emission, world_normal=n, world_position=hit_pos))


def primary_rays(sess, grid):
    """NDC grid points -> (origins on the near plane, unit directions) (mixin.py:31-39)."""
    V, P = sess.camera_V(), sess.camera_P()
    inv = small_matrix_inverse(torch.stack([V, sess.camera_VP()]))
    cam = inv[0, :3, 3]
    pts = torch.matmul(grid, inv[1].T)
    pts = pts[..., :3] / pts[..., 3:]
    d = normalized(pts - cam)
    return cam + d * (P[2, 3] / (P[2, 2] - 1)), d


def ray_may_reach_box(o: torch.Tensor, d: torch.Tensor, lo: torch.Tensor, hi: torch.Tensor) -> torch.Tensor:
    """(R,) bool: can the ray o + t d, t >= 0, touch the box [lo, hi]?  Tensor version of csrc/shade.cuh:ray_may_reach_box (same rule)."""
    par = d == 0
    safe = torch.where(par, torch.ones_like(d), d)
    t1, t2 = (lo - o) / safe, (hi - o) / safe
    tn = torch.where(par, torch.zeros_like(d), torch.minimum(t1, t2)).max(-1).values.clamp_min(0.0)
    tf = torch.where(par, torch.full_like(d, 3.0e38), torch.maximum(t1, t2)).min(-1).values
    outside = (par & ((o < lo) | (o > hi))).any(-1)
    return (tn <= tf) & ~outside


def trace_rays(sess, sampler: Callable, radiance_channels: int = 3, compact: bool = False):
    """Section x bounce loop with a user sampler (path_tracing.py:310-352); intersection = B200Raycaster.query.

    ``compact=True`` (extension): the sampler only sees the rays that can still contribute.  After every bounce a ray is
    dropped iff it hit nothing, its throughput is zero in every channel and its continuation cannot reach the padded scene box -- it would
    miss at every later bounce and add zero radiance.  Exact for any sampler that returns zero ``alpha`` for rays that miss (the built-in
    sampler does); the arrays passed to the sampler then shrink from bounce to bounce (open scenes: about half of the rays after bounce 0)
    and sums are scattered back by pixel."""
    from .path_tracing import hammersley
    opt, dev = sess.options, sess.device
    rc = sess.raycaster()
    far = sess.camera_far()
    H, W = sess.camera.resolution()
    ys = torch.linspace(-1 + 1 / H, 1 - 1 / H, H, dtype=torch.float32, device=dev).view(H, 1, 1).expand(H, W, 1)
    xs = torch.linspace(-1 + 1 / W, 1 - 1 / W, W, dtype=torch.float32, device=dev).view(1, W, 1).expand(H, W, 1)
    pix = torch.cat([xs, ys, -torch.ones_like(xs), torch.ones_like(xs)], -1).reshape(-1, 4)
    qx, qy = hammersley(opt.ray_spp, opt.deterministic, dev)
    sections = max(1, min(opt.ray_spp, math.ceil((H * W * opt.ray_spp) / opt.ray_split_size)))
    radiance = zeros_like_vec(pix, radiance_channels)
    alpha = zeros_like_vec(pix, 1)
    extras_sum, extras_cnt = {}, {}
    if compact:
        b = rc.stats()['bounds']
        lo, hi = torch.tensor(b[:3], dtype=torch.float32, device=dev), torch.tensor(b[3:], dtype=torch.float32, device=dev)
        pad = 1e-5 * torch.maximum(lo.abs(), hi.abs()).clamp_min(1e-30) + 1e-5 * (hi - lo).clamp_min(0)
        lo, hi = lo - pad, hi + pad
    for sx, sy in zip(torch.tensor_split(qx, sections), torch.tensor_split(qy, sections)):
        n = len(sx)
        sx, sy = sx[:, None, None], sy[:, None, None]
        grid = torch.cat([pix[:, 0:1] + (sx - 0.5) * (2 / W), pix[:, 1:2] + (sy - 0.5) * (2 / H), pix[:, 2:].expand(n, -1, -1)], -1).reshape(-1, 4)
        rays_o, rays_d = primary_rays(sess, grid)
        throughput = ones_like_vec(rays_o, radiance_channels)
        owner = torch.arange(H * W, device=dev).repeat(n) if compact else None   # pixel of every live ray
        for d in range(opt.ray_depth):
            t, i = rc.query(rays_o.detach(), rays_d.detach(), far)
            out = sampler(rays_o, rays_d, t, i, d)
            if compact and d > 0:
                radiance = radiance.index_add(0, owner, throughput * out.radiance)
                alpha = alpha.index_add(0, owner, out.alpha.reshape(-1, 1))
            else:
                radiance = radiance + (throughput * out.radiance).view(n, -1, radiance_channels).sum(0)
                alpha = alpha + out.alpha.reshape(n, -1, 1).sum(0)
            throughput = throughput * out.transfer
            rays_o, rays_d = out.next_rays_o, out.next_rays_d
            if d == 0:
                for k, v in out.extras.items():
                    s = v.reshape(n, H, W, v.shape[-1]).sum(0)
                    extras_sum[k] = extras_sum[k] + s if k in extras_sum else s
                    extras_cnt[k] = extras_cnt.get(k, 0) + n
            if compact and d + 1 < opt.ray_depth:
                dead = (t >= far) & (throughput == 0).all(-1) & ~ray_may_reach_box(rays_o.detach(), rays_d.detach(), lo, hi)
                keep = (~dead).nonzero(as_tuple=True)[0]
                rays_o, rays_d, throughput, owner = rays_o[keep], rays_d[keep], throughput[keep], owner[keep]
    radiance = radiance.reshape(H, W, radiance_channels) / opt.ray_spp
    alpha = saturate(alpha.reshape(H, W, 1) / opt.ray_spp)
    extras = {k: torch.flipud(extras_sum[k] / extras_cnt[k]) for k in extras_sum}
    return torch.flipud(radiance), torch.flipud(alpha), extras
