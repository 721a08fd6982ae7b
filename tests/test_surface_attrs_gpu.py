"""drp_surface_attrs: the material layer for arbitrary ray batches / custom samplers, against the oracle and the torch path."""
import numpy as np
import pytest
import torch

import oracle
import scenes
import diffrp_b200 as drp
from diffrp_b200 import generic
from test_oracle_golden import make_camera

pytestmark = pytest.mark.gpu
ORBIT = dict(h=40, w=56, radius=3.0, azim=25, elev=15, origin=[0.0, -0.1, 0.0], fov=32, near=0.1, far=10.0)


def rays_and_hits(sess, n=20000, seed=0):
    o, d = drp.synthetic.random_rays(n, origin_radius=2.5, target_sigma=0.6, seed=seed)
    o, d = torch.from_numpy(o).cuda(), torch.from_numpy(d).cuda()
    t, i = sess.raycaster().query(o, d, sess.camera_far())
    return o, d, t, i


def test_matches_oracle_and_torch_layer():
    scene = scenes.mixed_scene()
    sess = drp.PathTracingSession(scene, drp.PerspectiveCamera.from_orbit(**ORBIT), drp.PathTracingSessionOptions())
    o, d, t, i = rays_and_hits(sess)
    far = sess.camera_far()
    assert 0.2 < float((t < far).float().mean()) < 0.95
    attrs = sess.surface_attributes(o, d, t, i)
    assert attrs.shape == (len(o), 12)
    # 1. the CPU oracle (orc_surface_attrs) on the same hits.  Stated tolerance (fp32 interpolation + texture filtering): 2e-5 absolute
    cpu_cam = make_camera(None, ORBIT)
    vao, hs, p, keep = scenes.oracle_inputs(scene, cpu_cam, 1, 1)
    want = oracle.surface_attrs(hs, o.cpu().numpy(), d.cpu().numpy(), t.cpu().numpy(), i.cpu().numpy(), far)
    got = attrs.cpu().numpy()
    err = np.abs(got - want).max(-1)
    assert (err > 2e-5).mean() <= 1e-3, (err.max(), (err > 2e-5).mean())   # the rest: hits within rounding of a texel / triangle border
    miss = (t >= far).cpu().numpy()
    assert np.all(got[miss] == 0.0)
    # 2. the torch material layer of the generic path (SurfaceInput / MaskedSparseInterpolator / material.shade)
    mats = generic.layer_material_rays(sess, o, d, t, i)
    ref = generic.collect_gbuffer(mats, generic.surface_row, torch.zeros(len(o), 12, device='cuda'))
    err2 = (attrs - ref).abs().amax(-1)
    assert float((err2 > 2e-5).float().mean()) <= 1e-3, float(err2.max())


def test_custom_sampler_built_on_surface_attributes_equals_builtin():
    """A user sampler for trace_rays() that takes its g-buffer from surface_attributes() reproduces the built-in torch sampler."""
    scene = scenes.mixed_scene()
    cam = drp.PerspectiveCamera.from_orbit(**ORBIT)
    opts = dict(ray_spp=2, ray_depth=3)

    def make_sampler(sess):
        def sampler(rays_o, rays_d, t, i, depth):
            far = sess.camera_far()
            attrs = sess.surface_attributes(rays_o, rays_d, t, i)
            hit = t[..., None] < far
            env = generic.env_radiance(sess, rays_d)
            env = torch.where(hit, torch.zeros_like(env), env)
            u = [torch.rand_like(attrs[..., :1]) for _ in range(6)]
            albedo, emission, n, alpha, radiance, transfer, hit_pos, next_d = generic.brdf_sample_torch(attrs, t, rays_o, rays_d, env, u)
            return drp.RayOutputs(radiance=radiance, transfer=torch.where(hit, transfer, torch.zeros_like(transfer)),
                                  next_rays_o=hit_pos + next_d * sess.options.pbr_ray_step_epsilon, next_rays_d=next_d, alpha=alpha,
                                  extras=dict(albedo=albedo, world_normal=n))
        return sampler
    s1 = drp.PathTracingSession(scene, cam, drp.PathTracingSessionOptions(**opts))
    torch.manual_seed(7)
    rad1, alpha1, ex1 = s1.trace_rays(make_sampler(s1))
    s2 = drp.PathTracingSession(scene, cam, drp.PathTracingSessionOptions(**opts))
    torch.manual_seed(7)
    rad2, alpha2, ex2 = s2.trace_rays(s2.sampler_brdf)
    bad = ((rad1 - rad2).abs().amax(-1) > 1e-3).float().mean().item()
    assert bad <= 0.005, bad
    assert torch.allclose(alpha1, alpha2, atol=1e-6)
    assert ((ex1['albedo'] - ex2['albedo']).abs().amax(-1) > 1e-4).float().mean().item() <= 0.005


def test_errors_and_empty_batch():
    scene = scenes.icosphere_scene()
    sess = drp.PathTracingSession(scene, drp.PerspectiveCamera(h=8, w=8), drp.PathTracingSessionOptions())
    e = torch.empty(0, 3, device='cuda')
    assert sess.surface_attributes(e, e, torch.empty(0, device='cuda'), torch.empty(0, dtype=torch.int32, device='cuda')).shape == (0, 12)
    o = torch.zeros(4, 3, device='cuda'); d = torch.tensor([[0.0, 0.0, -1.0]], device='cuda').expand(4, 3).contiguous()
    t = torch.tensor([1.0, 2.0, 1e9, 1.0], device='cuda')
    i = torch.tensor([0, 5, 3, 10 ** 9], dtype=torch.int32, device='cuda')      # out-of-range id -> treated as a miss, no fault
    a = sess.surface_attributes(o, d, t, i)
    assert torch.isfinite(a).all() and (a[2] == 0).all() and (a[3] == 0).all()


def test_trace_rays_compact_view_equals_the_full_loop():
    """trace_rays(sampler, compact=True) calls a user sampler with the live rays only.  A deterministic custom sampler (mirror
    bounces shaded by the material layer kernel, no RNG) gives the same image either way; the per-bounce batch really shrinks."""
    from diffrp_b200.path_tracing import RayOutputs
    scene = scenes.mixed_scene().to(torch.device('cuda'))
    cam = drp.PerspectiveCamera.from_orbit(h=72, w=96, radius=3.0, azim=25, elev=15, origin=[0.0, -0.1, 0.0], fov=32)
    sizes = {}

    def run(compact):
        sess = drp.PathTracingSession(scene, cam, drp.PathTracingSessionOptions(ray_spp=2, ray_depth=4))
        far = sess.camera_far()
        seen = sizes.setdefault(compact, [])

        def sampler(rays_o, rays_d, t, i, d):
            seen.append(len(rays_o))
            a = sess.surface_attributes(rays_o, rays_d, t, i)
            hit = (t < far)[:, None]
            n = a[:, 3:6]
            pos = rays_o + rays_d * t[:, None]
            refl = rays_d - 2.0 * (rays_d * n).sum(-1, keepdim=True) * n
            sky = 0.5 + 0.5 * rays_d[:, 1:2].expand(-1, 3)
            return RayOutputs(radiance=torch.where(hit, a[:, 9:12], sky), transfer=torch.where(hit, a[:, 0:3] * 0.8, torch.zeros_like(n)),
                              next_rays_o=pos + refl * 1e-3, next_rays_d=torch.where(hit, refl, rays_d), alpha=hit.float() * a[:, 8:9],
                              extras=dict(albedo=a[:, 0:3]))
        return sess.trace_rays(sampler, compact=compact)
    full, comp = run(False), run(True)
    torch.testing.assert_close(comp[0], full[0], rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(comp[1], full[1], rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(comp[2]['albedo'], full[2]['albedo'], rtol=1e-5, atol=1e-6)
    assert sizes[False] == [72 * 96 * 2] * 4
    assert sizes[True][0] == 72 * 96 * 2 and sizes[True][1] < 0.9 * sizes[True][0] and sizes[True][3] <= sizes[True][2] <= sizes[True][1]
