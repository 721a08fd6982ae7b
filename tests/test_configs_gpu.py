"""
GPU parity of the BASELINE configurations that round 1 left without an oracle comparison ("what's missing" 1 and "what's weak" 1):

  * config 3 AT THE BENCH SETTINGS (bench.py: teaser scene with 1024^2 textures, rng='native', seed 1, 4 bounces, 1024-sample Hammersley
    sequence): the central 128 x 128 window of the 1024^2 frame, 8 spp, rendered by the fused kernels through the tile entry point and by the
    CPU oracle on the same window -- same RNG keys (global pixel, global sample id), same scene bytes;
  * config 4 (multi-view datagen, radiance + albedo + world_normal AOVs, views sharded) at reduced size vs the oracle, view by view;
  * config 5 (instanced scene: many MeshObjects sharing one mesh, 8 tints, env-lit, tile-sharded) at reduced size: 2-way tile shards sum to the
    whole frame and the whole frame equals the oracle's.

Tolerance (fp32, the one stated in tests/test_render_gpu.py): a pixel is an outlier if any channel differs by more than 1e-3 (rays grazing a
triangle edge: the primary-ray arithmetic differs in the last ulp between the host matmul and the kernel's FMAs); outliers <= 0.5 %, mean abs
error of the remaining pixels <= 2e-5.
"""
import numpy as np
import pytest
import torch

import oracle
import diffrp_b200 as drp
from diffrp_b200 import synthetic as syn, _abi
from test_oracle_golden import make_camera, image_errors

pytestmark = pytest.mark.gpu
OUTLIER_FRAC = 0.005
INLIER_MEAN = 2e-5


def check_images(out, ref, keys, what):
    errs = image_errors(out, ref, keys=keys)
    for k, (emax, emean, frac) in errs.items():
        assert frac <= OUTLIER_FRAC and emean <= INLIER_MEAN, (what, k, errs[k])


def window_of(acc, H, W, tile, spp):
    """(h, w, 16) un-flipped accumulator window / spp (row 0 = bottom row of the tile)."""
    x0, y0, w, h = tile
    a = acc.reshape(H, W, _abi.ACCUM_CHANNELS)[y0:y0 + h, x0:x0 + w] / np.float32(spp)
    return dict(radiance=a[..., 0:3], alpha=a[..., 3:4], albedo=a[..., 4:7], emission=a[..., 7:10], world_normal=a[..., 10:13],
                world_position=a[..., 13:16])


def test_config3_bench_scene_window_matches_oracle():
    """bench.py's scene, RNG mode, seed, texture size and sample sequence; the oracle renders the same 128^2 window."""
    RES, DEPTH, TOTAL_SPP, TEX, SEED = 1024, 4, 1024, 1024, 1
    ids = np.arange(8)
    tile = (RES // 2 - 64, RES // 2 - 64, 128, 128)
    scene_host, camkw = syn.teaser_scene('cpu', tex=TEX)
    cam = drp.PerspectiveCamera.from_orbit(h=RES, w=RES, **camkw)
    sess = drp.PathTracingSession(scene_host.to(torch.device('cuda')), cam,
                                  drp.PathTracingSessionOptions(ray_spp=TOTAL_SPP, ray_depth=DEPTH, rng='native', seed=SEED))
    assert sess.vertex_array_object().tris.shape[0] == 2_097_152
    acc = sess.render_samples(torch.from_numpy(ids).int(), tile=tile).cpu().numpy()
    got = window_of(acc, RES, RES, tile, len(ids))
    outside = acc.reshape(RES, RES, -1).copy()
    outside[tile[1]:tile[1] + tile[3], tile[0]:tile[0] + tile[2]] = 0
    assert not outside.any()                                     # the tile call touches no other accumulator row
    cpu_cam = make_camera(None, dict(h=RES, w=RES, **camkw))
    vao, hs, p, keep = oracle.inputs_from_scene(scene_host, cpu_cam, TOTAL_SPP, DEPTH, seed=SEED, sample_ids=ids)
    p.tile_x0, p.tile_y0, p.tile_w, p.tile_h = tile
    ref_acc, n = oracle.render(oracle.BVH(vao.world_pos.numpy(), vao.tris.numpy()), hs, p)
    assert n == 128 * 128 * len(ids) * DEPTH
    ref = window_of(ref_acc, RES, RES, tile, len(ids))
    assert float(ref['alpha'].mean()) > 0.5 and float(ref['radiance'].mean()) > 1e-3   # the window is not empty sky
    check_images(got, ref, ('radiance', 'alpha', 'albedo', 'emission', 'world_normal', 'world_position'), 'config 3 bench window')


def test_config4_multiview_aovs_match_oracle_and_views_shard():
    """64 cameras x 500k triangles reduced to 3 cameras x 20k triangles, 96^2, 8 spp, 3 bounces; AOVs radiance + albedo + world_normal."""
    scene_host, orbit = syn.datagen_scene('cpu', n_theta=100, n_phi=100, env_res=(32, 64))
    scene = scene_host.to(torch.device('cuda'))
    n_views, res, spp, depth = 3, 96, 8, 3
    hs = bvh = None
    for k in range(n_views):
        kw = orbit(k, n_views, res)
        sess = drp.PathTracingSession(scene, drp.PerspectiveCamera.from_orbit(**kw), drp.PathTracingSessionOptions(ray_spp=spp, ray_depth=depth, seed=k))
        rad, alpha, extras = sess.pbr()
        out = dict(radiance=rad.cpu().numpy(), alpha=alpha.cpu().numpy(), albedo=extras['albedo'].cpu().numpy(),
                   world_normal=extras['world_normal'].cpu().numpy())
        vao, hs, p, keep = oracle.inputs_from_scene(scene_host, make_camera(None, kw), spp, depth, seed=k)
        if bvh is None:
            bvh = oracle.BVH(vao.world_pos.numpy(), vao.tris.numpy())
        ref_acc, _ = oracle.render(bvh, hs, p)
        ref = oracle.finalize(ref_acc, res, res, spp)
        assert 0.2 < float(ref['alpha'].mean()) < 0.95
        check_images(out, ref, ('radiance', 'alpha', 'albedo', 'world_normal'), 'config 4 view %d' % k)
    # one flatten + one build for all views (options.reuse_scene), and the view shards of 8 ranks partition the 64 views
    assert sess.raycaster() is drp.PathTracingSession(scene, drp.PerspectiveCamera.from_orbit(**orbit(0, n_views, res))).raycaster()
    shards = [list(range(64))[r::8] for r in range(8)]
    assert sorted(sum(shards, [])) == list(range(64)) and all(len(s) == 8 for s in shards)


def test_config5_instanced_tile_shards_sum_to_the_whole_and_match_oracle():
    """1000 instances x 10k triangles reduced to 64 instances x 200 triangles, 192 x 128, 64-pixel tiles dealt to 2 ranks."""
    scene_host, camkw = syn.instanced_scene('cpu', n_instances=64, mesh_res=(10, 10), env_res=(32, 64), spread=(0.5, 0.3, 0.3))
    scene = scene_host.to(torch.device('cuda'))
    H, W, spp, depth = 128, 192, 8, 3
    camkw = dict(camkw, radius=2.0)
    opts = dict(ray_spp=spp, ray_depth=depth, seed=2, tile_size=64)
    cam = drp.PerspectiveCamera.from_orbit(h=H, w=W, **camkw)
    whole = drp.PathTracingSession(scene, cam, drp.PathTracingSessionOptions(**opts))
    assert whole.vertex_array_object().tris.shape[0] == 64 * 200
    acc = whole.render_accumulators()
    parts = [drp.PathTracingSession(scene, cam, drp.PathTracingSessionOptions(shard_rank=r, shard_world=2, shard_mode='tile', **opts)).render_accumulators()
             for r in range(2)]
    assert not ((parts[0].abs().sum(-1) > 0) & (parts[1].abs().sum(-1) > 0)).any()   # disjoint supports
    torch.testing.assert_close(parts[0] + parts[1], acc, rtol=1e-5, atol=1e-5)
    rad, alpha, extras = whole.finalize(parts[0] + parts[1])
    out = dict(radiance=rad.cpu().numpy(), alpha=alpha.cpu().numpy(), albedo=extras['albedo'].cpu().numpy(),
               world_normal=extras['world_normal'].cpu().numpy(), world_position=extras['world_position'].cpu().numpy())
    vao, hs, p, keep = oracle.inputs_from_scene(scene_host, make_camera(None, dict(h=H, w=W, **camkw)), spp, depth, seed=2)
    ref_acc, _ = oracle.render(oracle.BVH(vao.world_pos.numpy(), vao.tris.numpy()), hs, p)
    ref = oracle.finalize(ref_acc, H, W, spp)
    assert 0.05 < float(ref['alpha'].mean()) < 0.95
    assert len({tuple(np.round(c, 3)) for c in ref['albedo'].reshape(-1, 3)[ref['alpha'].reshape(-1) > 0.99][::37]}) >= 4  # several tints visible
    check_images(out, ref, ('radiance', 'alpha', 'albedo', 'world_normal', 'world_position'), 'config 5 reduced')
