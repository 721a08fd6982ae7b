"""Edge cases of the session API on the GPU: empty scene, no light, depth 1, odd sizes, spp = 1, options compatibility."""
import numpy as np
import pytest
import torch

import oracle
import scenes
import diffrp_b200 as drp
from test_oracle_golden import make_camera, image_errors

pytestmark = pytest.mark.gpu


def render(scene, cam, **opt):
    s = drp.PathTracingSession(scene, cam, drp.PathTracingSessionOptions(**opt))
    rad, alpha, extras = s.pbr()
    out = {k: v.cpu().numpy() for k, v in extras.items()}
    out['radiance'], out['alpha'] = rad.cpu().numpy(), alpha.cpu().numpy()
    return s, out


def test_empty_scene_renders_the_environment_only():
    scene = drp.Scene().add_light(drp.ImageEnvironmentLight(1.0, torch.ones(3), scenes.T(drp.synthetic.gradient_env())))
    s, out = render(scene, drp.PerspectiveCamera(h=16, w=24), ray_spp=2, ray_depth=2)
    assert out['alpha'].max() == 0.0 and out['radiance'].min() >= 0.0 and out['radiance'].max() > 0.0
    assert np.abs(out['albedo']).max() == 0.0
    assert s.render_stats()['rays_traced'] == 16 * 24 * 2  # every ray is dropped after the first (provable) miss


def test_no_light_gives_black_radiance_and_correct_alpha():
    scene = scenes.icosphere_scene(env=False)
    cam = drp.PerspectiveCamera(h=33, w=47)  # odd, non-square
    s, out = render(scene, cam, ray_spp=3, ray_depth=1)
    assert out['radiance'].shape == (33, 47, 3) and np.abs(out['radiance']).max() == 0.0
    cpu_cam = make_camera(dict(h=33, w=47), None)
    vao, hs, p, keep = scenes.oracle_inputs(scene, cpu_cam, 3, 1)
    acc, n = oracle.render(oracle.BVH(vao.world_pos.numpy(), vao.tris.numpy()), hs, p)
    ref = oracle.finalize(acc, 33, 47, 3)
    errs = image_errors(out, ref)
    for k, (emax, emean, frac) in errs.items():
        assert frac <= 0.005 and emean <= 2e-5, (k, errs[k])


def test_reference_option_names_are_accepted():
    scene = scenes.icosphere_scene()
    cam = drp.PerspectiveCamera(h=16, w=16)
    outs = []
    for impl in ('torchoptix', 'naive-pbbvh', 'brute-force', 'b200'):
        s, out = render(scene, cam, ray_spp=2, ray_depth=2, raycaster_impl=impl, raycaster_builder='morton', optix_log_level=0, seed=1)
        outs.append(out['radiance'])
    for o in outs[1:]:
        np.testing.assert_allclose(o, outs[0], rtol=1e-5, atol=1e-6)
    with pytest.raises(ValueError):
        render(scene, cam, raycaster_impl='embree')


def test_nondeterministic_hammersley_shift_uses_torch_rng():
    scene = scenes.icosphere_scene()
    cam = drp.PerspectiveCamera(h=16, w=16)
    torch.manual_seed(3)
    _, a = render(scene, cam, ray_spp=2, ray_depth=1, deterministic=False)
    torch.manual_seed(3)
    _, b = render(scene, cam, ray_spp=2, ray_depth=1, deterministic=False)
    torch.manual_seed(4)
    _, c = render(scene, cam, ray_spp=2, ray_depth=1, deterministic=False)
    np.testing.assert_allclose(a['world_position'], b['world_position'], rtol=0, atol=1e-6)
    assert np.abs(a['world_position'] - c['world_position']).max() > 1e-4


def test_many_samples_split_into_internal_batches():
    """More rays than one internal batch (2^24): the result must equal the sum of two half renders."""
    scene = scenes.icosphere_scene()
    cam = drp.PerspectiveCamera(h=512, w=512)
    base = dict(ray_spp=96, ray_depth=2, seed=6)  # 512*512*96 = 25.2M rays > 16.8M
    s = drp.PathTracingSession(scene, cam, drp.PathTracingSessionOptions(**base))
    whole = s.render_accumulators()
    assert s.render_stats()['rays_nominal'] == 512 * 512 * 96 * 2
    ids = torch.arange(96, dtype=torch.int32, device='cuda')
    s2 = drp.PathTracingSession(scene, cam, drp.PathTracingSessionOptions(**base))
    parts = s2.render_samples(ids[:40])
    parts = s2.render_samples(ids[40:], parts)
    torch.testing.assert_close(parts, whole, rtol=2e-5, atol=2e-4)


def test_scene_cache_is_shared_between_sessions_and_invalidated_by_in_place_edits():
    scene = scenes.to_device(scenes.mixed_scene(), 'cuda')
    cam = drp.PerspectiveCamera.from_orbit(h=32, w=48, radius=3.0, azim=5, elev=10, origin=[0.0, -0.1, 0.0], fov=32)
    a = drp.PathTracingSession(scene, cam, drp.PathTracingSessionOptions(ray_spp=2, ray_depth=2, seed=1))
    ra = a.pbr()[0]
    b = drp.PathTracingSession(scene, drp.PerspectiveCamera.from_orbit(h=32, w=48, radius=3.0, azim=90, elev=10, origin=[0.0, -0.1, 0.0], fov=32),
                               drp.PathTracingSessionOptions(ray_spp=2, ray_depth=2, seed=1))
    assert b.raycaster() is a.raycaster() and b.vertex_array_object() is a.vertex_array_object()  # built once, used by both views
    rb = b.pbr()[0]
    assert (ra - rb).abs().max() > 1e-3  # a different view, really rendered
    private = drp.PathTracingSession(scene, cam, drp.PathTracingSessionOptions(ray_spp=2, ray_depth=2, seed=1, reuse_scene=False))
    assert private.raycaster() is not a.raycaster()
    torch.testing.assert_close(private.pbr()[0], ra, rtol=1e-5, atol=1e-6)
    scene.objects[1].verts.mul_(0.5)  # in-place edit bumps the tensor version -> the cache entry is stale
    c = drp.PathTracingSession(scene, cam, drp.PathTracingSessionOptions(ray_spp=2, ray_depth=2, seed=1))
    assert c.vertex_array_object() is not a.vertex_array_object()      # re-flattened ...
    assert c.raycaster() is a.raycaster()                              # ... and, the connectivity being unchanged, REFITTED (drp_refit), not rebuilt
    rc_img = c.pbr()[0]
    assert (rc_img - ra).abs().max() > 1e-3
    fresh = drp.PathTracingSession(scene, cam, drp.PathTracingSessionOptions(ray_spp=2, ray_depth=2, seed=1, reuse_scene=False))
    torch.testing.assert_close(fresh.pbr()[0], rc_img, rtol=1e-5, atol=1e-6)   # the refitted structure renders what a rebuilt one renders
    scene.objects[1].verts.mul_(2.0)
    d = drp.PathTracingSession(scene, cam, drp.PathTracingSessionOptions(ray_spp=2, ray_depth=2, seed=1, refit_scene=False))
    assert d.raycaster() is not a.raycaster()                          # opt-out: rebuilt


def test_torchoptix_shaped_module_drives_raw_pointers():
    """diffrp_b200.optix_compat has the call surface diffrp's TorchOptiX wrapper uses (raycaster.py:267-296)."""
    import diffrp_b200.optix_compat as optix
    from diffrp_b200 import synthetic as syn
    v, f = syn.icosphere(3, 0.8)
    o, d = syn.random_rays(30_000, seed=2)
    V, F, O, D = (torch.from_numpy(x).cuda() for x in (v, f, o, d))
    optix.set_log_level(0)
    handle = optix.build(V.data_ptr(), F.data_ptr(), len(V), len(F))
    out_t = O.new_empty([len(O)])
    out_i = O.new_empty([len(O)], dtype=torch.int32)
    optix.trace_rays(handle, O.data_ptr(), D.data_ptr(), out_t.data_ptr(), out_i.data_ptr(), 10.0, len(O))
    torch.cuda.synchronize()
    optix.release(handle)
    ot, oi = oracle.bruteforce(v, f, o, d, 10.0, 1e-8)
    assert np.array_equal(out_t.cpu().numpy().view(np.int32), ot.view(np.int32)) and np.array_equal(out_i.cpu().numpy(), oi)


def test_triangle_id_float_packing_round_trips():
    from diffrp_b200.generic import triidx_to_float, float_to_triidx
    ids = torch.tensor([0, 1, 5, 2 ** 24 - 1, 2 ** 24, 2 ** 24 + 1, 2 ** 24 + 12345, 2 ** 26 + 7], dtype=torch.int32, device='cuda')
    assert torch.equal(float_to_triidx(triidx_to_float(ids)), ids)


def test_raycaster_rejects_bad_inputs():
    from diffrp_b200 import synthetic as syn
    v, f = syn.icosphere(1, 0.8)
    V, F = torch.from_numpy(v).cuda(), torch.from_numpy(f).cuda()
    rc = drp.B200Raycaster(V, F)
    with pytest.raises(ValueError):
        rc.query(torch.zeros(4, 3), torch.zeros(4, 3), 1.0)            # CPU rays
    with pytest.raises(ValueError):
        rc.query(torch.zeros(4, 3).cuda(), torch.zeros(5, 3).cuda(), 1.0)  # shape mismatch
    with pytest.raises(ValueError):
        drp.B200Raycaster(V, (F + 1000))                                # indices out of range
    rc.release()
    with pytest.raises(RuntimeError):
        rc.query(torch.zeros(4, 3).cuda(), torch.zeros(4, 3).cuda(), 1.0)


def test_reproducible_mode_is_bit_identical_run_to_run():
    scene = scenes.mixed_scene()
    cam = drp.PerspectiveCamera.from_orbit(h=64, w=96, radius=3.0, azim=25, elev=15, origin=[0.0, -0.1, 0.0], fov=32)
    opt = dict(ray_spp=6, ray_depth=4, seed=3, reproducible=True)
    a = drp.PathTracingSession(scene, cam, drp.PathTracingSessionOptions(**opt)).render_accumulators()
    b = drp.PathTracingSession(scene, cam, drp.PathTracingSessionOptions(**opt)).render_accumulators()
    assert torch.equal(a, b)
    c = drp.PathTracingSession(scene, cam, drp.PathTracingSessionOptions(**dict(opt, reproducible=False))).render_accumulators()
    torch.testing.assert_close(a, c, rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("where", ["pinned", "device", "pageable"])
def test_cuda_flatten_matches_torch_flatten(where):
    """drp_flatten (one pass, zero-copy reads of pinned host sources) vs the torch twin used by the CPU tests."""
    from diffrp_b200.flatten import flatten_scene, flatten_scene_cuda
    scene = scenes.mixed_scene()
    if where == "device":
        scene = scenes.to_device(scene, 'cuda')
    elif where == "pinned":
        for o in scene.objects:
            for name in ("verts", "tris", "normals", "M", "color", "uv", "tangents"):
                setattr(o, name, getattr(o, name).contiguous().pin_memory())
    a = flatten_scene_cuda(scene.objects, 'cuda')
    b = flatten_scene(scene.objects, 'cuda')
    torch.cuda.synchronize()
    for k in ("verts", "normals", "world_pos", "color", "uv", "tangents", "world_nrm", "world_tan"):
        torch.testing.assert_close(getattr(a, k), getattr(b, k), rtol=1e-6, atol=1e-6, msg=k)
    for k in ("tris", "stencils", "tri_material"):
        assert torch.equal(getattr(a, k), getattr(b, k)), k
    rec = torch.cat([b.world_pos, b.world_nrm, b.uv, b.color, b.world_tan], 1)
    torch.testing.assert_close(a.records, rec, rtol=1e-6, atol=1e-6)


def test_pbr_is_differentiable_like_the_reference():
    """The reference's trace_rays is differentiable through material / sampler code (only raycaster.query is no_grad).  A scene whose
    tensors require grad is routed to the generic path: gradients reach vertex colours; without grad the fused path returns plain tensors."""
    scene = scenes.icosphere_scene()
    obj = scene.objects[0]
    obj.color = obj.color.clone().cuda().requires_grad_(True)
    cam = drp.PerspectiveCamera(h=24, w=24)
    torch.manual_seed(0)
    rad, alpha, extras = drp.PathTracingSession(scene, cam, drp.PathTracingSessionOptions(ray_spp=4, ray_depth=2)).pbr()
    assert rad.requires_grad and extras['albedo'].requires_grad
    (rad.sum() + extras['albedo'].sum()).backward()
    g = obj.color.grad
    assert g is not None and torch.isfinite(g).all() and float(g.abs().sum()) > 0.0
    with torch.no_grad():
        rad2, _, _ = drp.PathTracingSession(scene, cam, drp.PathTracingSessionOptions(ray_spp=4, ray_depth=2)).pbr()
    assert not rad2.requires_grad
    # same estimator on both paths: albedo AOV is noise-free up to the jittered sub-pixel positions (identical Hammersley set)
    obj.color = obj.color.detach()
    _, _, ex3 = drp.PathTracingSession(scene, cam, drp.PathTracingSessionOptions(ray_spp=4, ray_depth=2)).pbr()
    np.testing.assert_allclose(ex3['albedo'].cpu().numpy(), extras['albedo'].detach().cpu().numpy(), atol=2e-5)


def test_fixup_path_renders_the_same_image():
    """Whole renders with the fast traversal stack lowered to 1 entry (every deep ray goes through k_extend_fixup) are bit-identical to the
    normal ones in reproducible mode."""
    from diffrp_b200._lib import lib, check
    scene = scenes.mixed_scene()
    cam = drp.PerspectiveCamera(h=48, w=64)
    opt = dict(ray_spp=4, ray_depth=3, seed=11, reproducible=True, reuse_scene=False)
    s1 = drp.PathTracingSession(scene, cam, drp.PathTracingSessionOptions(**opt))
    a = s1.pbr()
    s2 = drp.PathTracingSession(scene, cam, drp.PathTracingSessionOptions(**opt))
    check(lib().drp_debug_set_stack_limit(s2.raycaster().handle, 1), "drp_debug_set_stack_limit")
    b = s2.pbr()
    assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])
    for k in a[2]:
        assert torch.equal(a[2][k], b[2][k]), k


def test_more_than_2_pow_31_rays_in_one_pbr_call():
    """1024 x 1024 at 2049 spp is > 2^31 primary rays; drp_render sections the samples internally (like the reference's
    ray_split_size sections) instead of refusing.  Depth 1 over an empty scene keeps it cheap: every pixel must see exactly spp samples."""
    scene = drp.Scene().add_light(drp.ImageEnvironmentLight(1.0, torch.ones(3), torch.ones(4, 8, 3)))
    rad, alpha, _ = drp.PathTracingSession(scene, drp.PerspectiveCamera(h=1024, w=1024), drp.PathTracingSessionOptions(ray_spp=2049, ray_depth=1)).pbr()
    np.testing.assert_allclose(rad.cpu().numpy(), 1.0, rtol=1e-4)
    assert float(alpha.max()) == 0.0


def test_unsharded_process_cannot_silently_lose_samples():
    scene = scenes.icosphere_scene()
    with pytest.raises(RuntimeError):
        drp.PathTracingSession(scene, drp.PerspectiveCamera(h=8, w=8), drp.PathTracingSessionOptions(ray_spp=4, shard_world=2)).pbr()


def test_host_scene_paths_render_the_same_image():
    """The session accepts scenes on the device, in pageable host memory and in one page-locked arena (Scene.pin_memory(): packed upload with a
    few large copies, textures on the side stream): same image bit for bit in reproducible mode."""
    host = scenes.mixed_scene()
    cam = drp.PerspectiveCamera(h=48, w=64)
    opt = dict(ray_spp=4, ray_depth=3, seed=13, reproducible=True, reuse_scene=False)
    outs = []
    for sc in (host.to(torch.device('cuda')), host, host.pin_memory()):
        rad, alpha, extras = drp.PathTracingSession(sc, cam, drp.PathTracingSessionOptions(**opt)).pbr()
        outs.append(torch.cat([rad, alpha] + [extras[k] for k in sorted(extras)], -1))
    pinned = host.pin_memory()
    assert pinned._arena.is_pinned() and all(o.verts.is_pinned() for o in pinned.objects)
    assert torch.equal(outs[0], outs[1]) and torch.equal(outs[0], outs[2])
    assert float(outs[0][..., 3].mean()) > 0.1
