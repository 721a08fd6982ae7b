"""
Feed the REFERENCE (the unmodified ``diffrp`` that baseline/ref_loader.py imports) the same synthetic inputs as ``diffrp_b200``.

TEST / MEASUREMENT INFRASTRUCTURE ONLY (tests/, tests/golden/make_golden.py, the reference legs of bench.py).

* ``to_reference_scene``: a ``diffrp_b200.Scene`` -> the reference's own ``Scene`` / ``MeshObject`` / ``GLTFMaterial`` / ``DefaultMaterial`` /
  ``ImageEnvironmentLight`` objects over the same tensors (moved to ``device``).
* ``window_camera``: the reference's ``RawCamera`` (rendering/camera.py:49-66) for a pixel window of a larger frame -- the bounded sample
  of a full-frame workload that the CPU reference can finish in seconds.  Crop in clip space: x' = (W/w) x + tx w_c, y' likewise;
  the depth rows of P are untouched, so ``camera_far`` (rendering/mixin.py:41-44) is the full frame's.
"""
import torch


def to_reference_scene(diffrp, scene, device=None):
    import diffrp_b200 as drp
    mv = (lambda x: x) if device is None else (lambda x: x.to(device) if isinstance(x, torch.Tensor) else x)
    mat_cache = {}

    def material(m):
        if id(m) in mat_cache:
            return mat_cache[id(m)]
        if isinstance(m, drp.GLTFMaterial):
            smp = lambda s: None if s is None else diffrp.GLTFSampler(mv(s.image), s.wrap_mode, s.interpolation)  # noqa: E731
            rm = diffrp.GLTFMaterial(mv(m.base_color_factor), smp(m.base_color_texture), m.metallic_factor, m.roughness_factor,
                                     smp(m.metallic_roughness_texture), smp(m.normal_texture), smp(m.occlusion_texture),
                                     mv(m.emissive_factor), smp(m.emissive_texture), m.alpha_cutoff, m.alpha_mode)
        elif isinstance(m, drp.DefaultMaterial):
            rm = diffrp.DefaultMaterial(mv(m.tint))
        else:
            raise TypeError("no reference counterpart for material %r" % (type(m),))
        mat_cache[id(m)] = rm
        return rm

    out = diffrp.Scene()
    for o in scene.objects:  # objects of a diffrp_b200.Scene are already preprocessed (defaults filled, flat normals -> face soup)
        col = o.color if o.color.shape[-1] == 4 else torch.cat([o.color, torch.ones_like(o.color[:, :1])], -1)
        out.objects.append(diffrp.MeshObject(material(o.material), mv(o.verts), mv(o.tris), mv(o.normals), mv(o.M), mv(col), mv(o.uv),
                                             mv(o.tangents), {k: mv(v) for k, v in dict(o.custom_attrs).items()}, {}))
    for l in scene.lights:
        out.add_light(diffrp.ImageEnvironmentLight(l.intensity, mv(l.color), mv(l.image), l.render_skybox))
    return out


def window_camera(diffrp, V: torch.Tensor, P: torch.Tensor, H: int, W: int, x0: int, y0: int, w: int, h: int):
    """RawCamera whose (h, w) image is the pixel window [x0, x0+w) x [y0, y0+h) (y counted from the bottom row, like NDC) of the (H, W) frame."""
    C = torch.eye(4, dtype=torch.float32, device=P.device)
    xa, xb = -1.0 + 2.0 * x0 / W, -1.0 + 2.0 * (x0 + w) / W
    ya, yb = -1.0 + 2.0 * y0 / H, -1.0 + 2.0 * (y0 + h) / H
    C[0, 0], C[0, 3] = 2.0 / (xb - xa), -(xa + xb) / (xb - xa)
    C[1, 1], C[1, 3] = 2.0 / (yb - ya), -(ya + yb) / (yb - ya)
    # clip-space w multiplies the translation: P' = (S + T e_w^T) P  with  x_ndc' = sx x_ndc + tx
    Pw = P.clone().to(torch.float32)
    Pw[0] = C[0, 0] * P[0] + C[0, 3] * P[3]
    Pw[1] = C[1, 1] * P[1] + C[1, 3] * P[3]
    return diffrp.RawCamera(h, w, V.to(torch.float32), Pw)
