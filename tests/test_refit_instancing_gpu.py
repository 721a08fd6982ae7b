"""
SURVEY 8 f2 on the GPU: BVH refit (dynamic scenes) and instanced assembly.  Both keep the closest-hit contract, so the check is bit-equality with
a fresh drp_build over the same world-space arrays -- and with the CPU oracle's exhaustive search.
"""
import numpy as np
import pytest
import torch

import oracle
import scenes
import diffrp_b200 as drp
from diffrp_b200 import synthetic as syn
from diffrp_b200.raycaster import B200Raycaster
from diffrp_b200.path_tracing import scene_instances

pytestmark = pytest.mark.gpu


def _same_hits(a, b):
    return torch.equal(a[0].view(torch.int32), b[0].view(torch.int32)) and torch.equal(a[1], b[1])


@pytest.mark.parametrize("motion", ["small", "large"])
def test_refit_gives_the_hits_of_a_fresh_build(motion):
    v, f = syn.uv_sphere(192, 96, radius=0.8, bump=0.05, noise=0.01, seed=0)
    o, d = syn.random_rays(300_000, seed=3)
    to, td = torch.from_numpy(o).cuda(), torch.from_numpy(d).cuda()
    tv, tf = torch.from_numpy(v).cuda(), torch.from_numpy(f).cuda()
    rc = B200Raycaster(tv, tf, {'epsilon': 1e-8})
    rng = np.random.default_rng(1)
    amp = 0.01 if motion == "small" else 0.6          # "large": a twist that scrambles the Morton order the topology was built for
    ang = amp * v[:, 1:2] * 4.0
    v2 = np.concatenate([v[:, 0:1] * np.cos(ang) - v[:, 2:3] * np.sin(ang), v[:, 1:2] * (1 + amp), v[:, 0:1] * np.sin(ang) + v[:, 2:3] * np.cos(ang)], 1)
    v2 = (v2 + rng.normal(0, amp * 0.05, v2.shape)).astype(np.float32)
    tv2 = torch.from_numpy(v2).cuda()
    rc.refit(tv2)
    got = rc.query(to, td, 10.0)
    fresh = B200Raycaster(tv2, tf, {'epsilon': 1e-8}).query(to, td, 10.0)
    assert _same_hits(got, fresh)
    assert 0.3 < float((got[0] < 10.0).float().mean()) < 0.95
    sl = slice(0, 3000)
    ot, oi = oracle.bruteforce(v2, f, o[sl], d[sl], 10.0, 1e-8)
    assert np.array_equal(got[0][sl].cpu().numpy().view(np.int32), ot.view(np.int32)) and np.array_equal(got[1][sl].cpu().numpy(), oi)
    rc.refit(tv)                                      # and back: the structure is reusable indefinitely
    assert _same_hits(rc.query(to, td, 10.0), B200Raycaster(tv, tf, {'epsilon': 1e-8}).query(to, td, 10.0))
    rc.check_status()
    with pytest.raises(ValueError):
        rc.refit(tv[:-3])


def test_session_refits_when_only_transforms_change():
    """Second session over the same Scene after an in-place change of a model matrix: the cached structure is refitted (same handle), and the
    image equals the one of a scene built from scratch at the new pose, bit for bit in reproducible mode."""
    def make(angle):
        sc = scenes.mixed_scene().to(torch.device('cuda'))
        c, s = np.cos(angle), np.sin(angle)
        R = torch.tensor([[c, 0, s, 0], [0, 1, 0, 0.05 * angle], [-s, 0, c, 0], [0, 0, 0, 1]], dtype=torch.float32, device='cuda')
        sc.objects[1].M = R @ sc.objects[1].M
        return sc
    cam = drp.PerspectiveCamera(h=64, w=80)
    opt = dict(ray_spp=4, ray_depth=3, seed=5, reproducible=True)
    scene = make(0.0)
    s1 = drp.PathTracingSession(scene, cam, drp.PathTracingSessionOptions(**opt))
    s1.pbr()
    h1 = s1.raycaster().handle
    moved = make(0.7)
    scene.objects[1].M = moved.objects[1].M            # only a transform changes; every index tensor is untouched
    s2 = drp.PathTracingSession(scene, cam, drp.PathTracingSessionOptions(**opt))
    a = s2.pbr()
    assert s2.raycaster().handle == h1                 # refitted, not rebuilt
    b = drp.PathTracingSession(moved, cam, drp.PathTracingSessionOptions(**opt)).pbr()
    assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])
    for k in a[2]:
        assert torch.equal(a[2][k], b[2][k]), k
    s3 = drp.PathTracingSession(scene, cam, drp.PathTracingSessionOptions(refit_scene=False, **opt))
    scene.objects[1].M = make(0.2).objects[1].M
    s3 = drp.PathTracingSession(scene, cam, drp.PathTracingSessionOptions(refit_scene=False, **opt))
    s3.pbr()
    assert s3.raycaster().handle != h1                 # opt-out rebuilds


@pytest.mark.parametrize("n_inst,mesh_res", [(1, (12, 8)), (9, (10, 6)), (200, (16, 10))])
def test_instanced_build_gives_the_hits_of_the_flat_build(n_inst, mesh_res):
    scene_host, camkw = syn.instanced_scene('cpu', n_instances=n_inst, mesh_res=mesh_res, env_res=(8, 16), spread=(0.6, 0.4, 0.4))
    scene = scene_host.to(torch.device('cuda'))
    sess = drp.PathTracingSession(scene, drp.PerspectiveCamera.from_orbit(h=8, w=8, **camkw), drp.PathTracingSessionOptions(instancing=False))
    vao = sess.vertex_array_object()
    nt = 2 * mesh_res[0] * mesh_res[1]
    first = np.arange(n_inst + 1, dtype=np.int64) * nt
    mesh = np.zeros(n_inst, dtype=np.int32)
    inst = B200Raycaster(vao.world_pos, vao.tris, {'epsilon': 1e-8, 'instances': (first, mesh)})
    flat = B200Raycaster(vao.world_pos, vao.tris, {'epsilon': 1e-8})
    assert inst.instanced and not flat.instanced
    o, d = syn.random_rays(200_000, seed=8)
    o = (o * 0.8).astype(np.float32)
    to, td = torch.from_numpy(o).cuda(), torch.from_numpy(d).cuda()
    a, b = inst.query(to, td, 10.0), flat.query(to, td, 10.0)
    assert _same_hits(a, b)
    assert float((a[0] < 10.0).float().mean()) > (0.0005 if n_inst == 1 else 0.02)
    st = inst.stats()
    assert st['n_tris'] == n_inst * nt and st['n_leaves'] > 0
    inst.check_status()
    with pytest.raises(RuntimeError):
        inst.refit(vao.world_pos)
    # two different meshes, interleaved
    if n_inst == 9:
        v2, f2 = syn.icosphere(2, 0.07)
        objs = list(scene.objects)
        tv2, tf2 = torch.from_numpy(v2).cuda(), torch.from_numpy(f2).cuda()
        tn2 = torch.nn.functional.normalize(tv2, dim=-1)
        extra = [drp.MeshObject(objs[0].material, tv2, tf2, normals=tn2,
                                M=scenes.rigid(50 + k, 1.0, (0.1 * k - 0.4, 0.2, 0.1)).cuda()) for k in range(9)]
        sc2 = drp.Scene()
        for x, y in zip(objs, extra):
            sc2.add_mesh_object(x).add_mesh_object(y)
        tab = scene_instances(sc2.objects)
        assert tab is not None and len(tab[1]) == 18 and len(set(tab[1].tolist())) == 2
        s2 = drp.PathTracingSession(sc2, drp.PerspectiveCamera.from_orbit(h=8, w=8, **camkw), drp.PathTracingSessionOptions(instancing=True))
        assert s2.raycaster().instanced
        v = s2.vertex_array_object()
        assert _same_hits(s2.raycaster().query(to, td, 10.0), B200Raycaster(v.world_pos, v.tris, {'epsilon': 1e-8}).query(to, td, 10.0))


def test_config5_session_uses_instancing_and_renders_the_same_image():
    scene_host, camkw = syn.instanced_scene('cpu', n_instances=64, mesh_res=(10, 10), env_res=(32, 64), spread=(0.5, 0.3, 0.3))
    scene = scene_host.to(torch.device('cuda'))
    cam = drp.PerspectiveCamera.from_orbit(h=96, w=128, **dict(camkw, radius=2.0))
    opt = dict(ray_spp=4, ray_depth=3, seed=2, reproducible=True, reuse_scene=False)
    a_s = drp.PathTracingSession(scene, cam, drp.PathTracingSessionOptions(instancing=True, **opt))
    b_s = drp.PathTracingSession(scene, cam, drp.PathTracingSessionOptions(instancing=False, **opt))
    a, b = a_s.pbr(), b_s.pbr()
    assert a_s.raycaster().instanced and not b_s.raycaster().instanced
    assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])
    for k in a[2]:
        assert torch.equal(a[2][k], b[2][k]), k
