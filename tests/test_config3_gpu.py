"""BASELINE config 3 at FULL size (2,097,152 triangles, 1024 x 1024, 4 bounces) through size-independent properties: the oracle cannot render
this in seconds, so the parity evidence at this size is (i) closest hits of a ray slice bit-identical to the exhaustive GPU kernel (same
triangle test and tie rule as the oracle's brute force), (ii) exact compaction == no compaction, (iii) sample shards and tile shards sum
to the whole frame, (iv) the reproducible mode repeats bit for bit, (v) the g-buffer of the frame agrees with an independent
query + surface_attributes evaluation of the primary rays."""
import numpy as np
import pytest
import torch

import diffrp_b200 as drp
from diffrp_b200 import synthetic as syn, generic

pytestmark = pytest.mark.gpu
RES, DEPTH, SPP = 1024, 4, 4


@pytest.fixture(scope="module")
def setup():
    scene_host, camkw = syn.teaser_scene('cpu', tex=256)
    scene = scene_host.to(torch.device('cuda'))
    cam = drp.PerspectiveCamera.from_orbit(h=RES, w=RES, **camkw)
    return scene, cam


def session(setup, **kw):
    scene, cam = setup
    opt = dict(ray_spp=SPP, ray_depth=DEPTH, rng='native', seed=4)
    opt.update(kw)
    return drp.PathTracingSession(scene, cam, drp.PathTracingSessionOptions(**opt))


def test_scene_is_config3(setup):
    s = session(setup)
    assert s.vertex_array_object().tris.shape[0] == 2_097_152


def test_hits_bit_identical_to_exhaustive_kernel(setup):
    s = session(setup)
    rc, far = s.raycaster(), s.camera_far()
    o, d = syn.random_rays(40_000, origin_radius=1.2, target_sigma=0.25, seed=3)
    o, d = torch.from_numpy(o).cuda(), torch.from_numpy(d).cuda()
    t, i = rc.query(o, d, far)
    tb, ib = rc.query_bruteforce(o, d, far)
    assert 0.2 < float((t < far).float().mean()) < 0.999
    assert torch.equal(t.view(torch.int32), tb.view(torch.int32)) and torch.equal(i, ib)


def test_compaction_sharding_tiling_and_reproducibility(setup):
    a = session(setup, compaction=True)
    acc = a.render_accumulators()
    sa = a.render_stats()
    b = session(setup, compaction=False)
    acc_nc = b.render_accumulators()
    sb = b.render_stats()
    assert sb['rays_traced'] == sb['rays_nominal'] == RES * RES * SPP * DEPTH and sa['rays_traced'] < sb['rays_traced']
    torch.testing.assert_close(acc, acc_nc, rtol=1e-5, atol=1e-5)              # fp32 atomics reorder only
    del acc_nc
    parts = [session(setup, shard_rank=r, shard_world=2).render_accumulators() for r in range(2)]
    torch.testing.assert_close(parts[0] + parts[1], acc, rtol=1e-5, atol=1e-5)
    tiles = [session(setup, shard_rank=r, shard_world=2, shard_mode='tile', tile_size=256).render_accumulators() for r in range(2)]
    assert not ((tiles[0].abs().sum(-1) > 0) & (tiles[1].abs().sum(-1) > 0)).any()
    torch.testing.assert_close(tiles[0] + tiles[1], acc, rtol=1e-5, atol=1e-5)
    del parts, tiles
    r1 = session(setup, ray_spp=2, reproducible=True).render_accumulators()
    r2 = session(setup, ray_spp=2, reproducible=True).render_accumulators()
    assert torch.equal(r1, r2)


def test_gbuffer_agrees_with_independent_primary_ray_evaluation(setup):
    """albedo / world_normal / world_position of a 1-spp frame vs query() + surface_attributes() on the same primary rays."""
    s = session(setup, ray_spp=1, ray_depth=1)
    rad, alpha, extras = s.pbr()
    H = W = RES
    dev = rad.device
    # the frame's single sample: Hammersley point 0 of 1 = (0, 0) -> sub-pixel offset (-0.5, -0.5) pixel (path_tracing.py:329)
    ys = torch.linspace(-1 + 1 / H, 1 - 1 / H, H, device=dev).view(H, 1, 1).expand(H, W, 1)
    xs = torch.linspace(-1 + 1 / W, 1 - 1 / W, W, device=dev).view(1, W, 1).expand(H, W, 1)
    qx, qy = drp.hammersley(1, True, dev)
    grid = torch.cat([xs + (qx[0] - 0.5) * (2 / W), ys + (qy[0] - 0.5) * (2 / H), -torch.ones_like(xs), torch.ones_like(xs)], -1).reshape(-1, 4)
    o, d = generic.primary_rays(s, grid)
    o = o.expand_as(d).contiguous()
    t, i = s.raycaster().query(o, d.contiguous(), s.camera_far())
    attrs = s.surface_attributes(o, d, t, i).view(H, W, 12).flip(0)               # outputs are flipud'd (row 0 = top)
    hit = (t < s.camera_far()).view(H, W, 1).flip(0)
    assert torch.equal(alpha > 0.5, hit)
    bad_albedo = ((extras['albedo'] - attrs[..., 0:3]).abs().amax(-1) > 1e-3).float().mean().item()
    bad_normal = ((extras['world_normal'] - attrs[..., 3:6]).abs().amax(-1) > 1e-3).float().mean().item()
    pos = (o + d * t[:, None]).view(H, W, 3).flip(0)      # also for misses (t == far), like the reference's sampler tail (path_tracing.py:266-279)
    bad_pos = ((extras['world_position'] - pos).abs().amax(-1) > 1e-4 * pos.abs().amax(-1).clamp_min(1.0)).float().mean().item()
    # rays generated in torch vs in the kernel differ in the last bits: allow the few pixels whose ray crosses a triangle / texel border
    assert bad_albedo <= 2e-3 and bad_normal <= 2e-3 and bad_pos <= 2e-3, (bad_albedo, bad_normal, bad_pos)
