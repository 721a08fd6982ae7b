"""
GPU parity of the closest-hit path (drp_build + drp_trace through B200Raycaster) against the CPU oracle.
Bar: t bit-identical (fp32 bit pattern), primitive id identical, hit/miss identical -- on every ray.
"""
import numpy as np
import pytest
import torch

import oracle
from diffrp_b200 import synthetic as syn
from diffrp_b200.raycaster import B200Raycaster

pytestmark = pytest.mark.gpu
FAR = 10.0


def _query(v, f, o, d, far=FAR, eps=1e-8):
    rc = B200Raycaster(torch.from_numpy(v).cuda(), torch.from_numpy(f).cuda(), {'epsilon': eps})
    t, i = rc.query(torch.from_numpy(o).cuda(), torch.from_numpy(d).cuda(), far)
    torch.cuda.synchronize()
    return rc, t.cpu().numpy(), i.cpu().numpy()


def _assert_same(t, i, ot, oi):
    assert t.dtype == np.float32 and i.dtype == np.int32
    assert np.array_equal(t.view(np.int32), ot.view(np.int32)), "t differs from the oracle: %d rays" % (t != ot).sum()
    assert np.array_equal(i, oi), "primitive id differs from the oracle: %d rays" % (i != oi).sum()


def test_icosphere_random_rays_vs_bruteforce_oracle():
    v, f = syn.icosphere(3, 0.8)
    o, d = syn.random_rays(200_000)
    rc, t, i = _query(v, f, o, d)
    ot, oi = oracle.bruteforce(v, f, o, d, FAR, 1e-8)
    _assert_same(t, i, ot, oi)
    assert 0.6 < (t < FAR).mean() < 0.9
    assert (t[t >= FAR] == np.float32(FAR)).all() and (i[t >= FAR] == 0).all()
    st = rc.stats()
    assert st['n_tris'] == 1280 and st['max_depth'] < 64


def test_gpu_bruteforce_kernel_matches_oracle():
    v, f = syn.icosphere(2, 0.8)
    o, d = syn.random_rays(50_000, seed=7)
    rc = B200Raycaster(torch.from_numpy(v).cuda(), torch.from_numpy(f).cuda(), {'epsilon': 1e-8})
    t, i = rc.query_bruteforce(torch.from_numpy(o).cuda(), torch.from_numpy(d).cuda(), FAR)
    ot, oi = oracle.bruteforce(v, f, o, d, FAR, 1e-8)
    _assert_same(t.cpu().numpy(), i.cpu().numpy(), ot, oi)


@pytest.mark.parametrize("n_tris", [0, 1, 2, 3, 5])
def test_tiny_and_empty_meshes(n_tris):
    v, f = syn.icosphere(1, 0.8)
    f = f[:n_tris].copy()
    o, d = syn.random_rays(20_000, seed=3)
    rc, t, i = _query(v, f, o, d)
    if n_tris == 0:
        assert (t == np.float32(FAR)).all() and (i == 0).all()
    else:
        ot, oi = oracle.bruteforce(v, f, o, d, FAR, 1e-8)
        _assert_same(t, i, ot, oi)


def test_empty_ray_batch():
    v, f = syn.icosphere(1, 0.8)
    rc, t, i = _query(v, f, np.zeros((0, 3), np.float32), np.zeros((0, 3), np.float32))
    assert t.shape == (0,) and i.shape == (0,)


def test_symmetric_edge_plane_rays_config1_stress():
    """Rays lying exactly in the icosphere's symmetry plane (x = 0) hit shared edges: every hit is a two-triangle
    tie that must resolve to the smaller id, exactly as the oracle / torch.argmin does."""
    v, f = syn.icosphere(3, 0.8)
    ys = np.linspace(-0.9, 0.9, 4001, dtype=np.float32)
    o = np.stack([np.zeros_like(ys), ys, np.full_like(ys, 3.2)], -1)
    d = np.tile(np.array([[0, 0, -1]], np.float32), (len(ys), 1))
    rc, t, i = _query(v, f, o, d)
    ot, oi = oracle.bruteforce(v, f, o, d, FAR, 1e-8)
    _assert_same(t, i, ot, oi)
    gt, gi = rc.query_bruteforce(torch.from_numpy(o).cuda(), torch.from_numpy(d).cuda(), FAR)
    _assert_same(t, i, gt.cpu().numpy(), gi.cpu().numpy())


def test_degenerate_pole_triangles_never_hit():
    v, f = syn.uv_sphere(64, 32)
    o, d = syn.random_rays(100_000, seed=11)
    rc, t, i = _query(v, f, o, d)
    ot, oi = oracle.bruteforce(v, f, o, d, FAR, 1e-8)
    _assert_same(t, i, ot, oi)
    tri = v[f[i[t < FAR]]]
    area = np.linalg.norm(np.cross(tri[:, 1] - tri[:, 0], tri[:, 2] - tri[:, 0]), axis=-1)
    assert (area > 0).all()


def test_million_triangles_vs_bvh_oracle_and_gpu_bruteforce():
    """Config-2 shape at a size the oracle finishes in seconds: 1M triangles, 1M rays vs the BVH oracle, and a
    100k-ray slice vs the exhaustive GPU kernel."""
    v, f = syn.uv_sphere(1024, 512)
    o, d = syn.random_rays(1_000_000)
    rc, t, i = _query(v, f, o, d)
    ot, oi = oracle.BVH(v, f).query(o, d, FAR, 1e-8)
    _assert_same(t, i, ot, oi)
    n = 100_000
    gt, gi = rc.query_bruteforce(torch.from_numpy(o[:n]).cuda(), torch.from_numpy(d[:n]).cuda(), FAR)
    _assert_same(t[:n], i[:n], gt.cpu().numpy(), gi.cpu().numpy())
    st = rc.stats()
    assert st['max_depth'] < 64 and st['n_leaves'] > 0


def test_noncontiguous_inputs_and_far_contract():
    v, f = syn.icosphere(2, 0.8)
    o, d = syn.random_rays(10_000, seed=5)
    rc = B200Raycaster(torch.from_numpy(v).cuda(), torch.from_numpy(f).cuda())
    od = torch.from_numpy(np.concatenate([o, d], -1)).cuda()
    t, i = rc.query(od[:, :3], od[:, 3:], 2.5)  # non-contiguous views, small far
    ot, oi = oracle.bruteforce(v, f, o, d, 2.5, 1e-8)
    _assert_same(t.cpu().numpy(), i.cpu().numpy(), ot, oi)
    assert i.dtype == torch.int32 and t.dtype == torch.float32


@pytest.mark.parametrize("scale,offset", [(1.0, 0.0), (1e-3, 0.0), (1e3, 0.0), (1.0, 1e4), (1e-2, 37.5)])
def test_random_triangle_soups_extreme_coordinates(scale, offset):
    """Same stress as tests/test_hostsim.py on the real kernels: sliver / degenerate / duplicate / flat triangles, axis-parallel
    rays, coordinates scaled by 1e-3..1e3 and translated up to 1e4: t and ids bit-identical to the exhaustive oracle."""
    rng = np.random.default_rng(3)
    n = 2000
    c = rng.uniform(-1, 1, (n, 1, 3))
    tri = c + rng.normal(0, 0.08, (n, 3, 3)) * rng.uniform(0.02, 1.0, (n, 1, 1))
    tri[:20, 2] = tri[:20, 1]
    tri[20:40] = tri[40:60]
    tri[60:80, :, 1] = tri[60:80, :1, 1]
    v = ((tri.reshape(-1, 3) * scale) + offset).astype(np.float32)
    f = np.arange(3 * n, dtype=np.int32).reshape(n, 3)
    m = 200_000
    org = (rng.uniform(-3, 3, (m, 3)) * scale + offset).astype(np.float32)
    tgt = (rng.uniform(-1, 1, (m, 3)) * scale + offset).astype(np.float32)
    d = tgt - org
    d = (d / np.linalg.norm(d, axis=-1, keepdims=True)).astype(np.float32)
    d[:500, 0] = 0.0
    d[500:1000, 1] = 0.0
    d[:1000] /= np.maximum(np.linalg.norm(d[:1000], axis=-1, keepdims=True), 1e-20)
    far = float(np.float32(20.0 * scale))
    rc, t, i = _query(v, f, org, d, far=far, eps=0.0)
    ot, oi = oracle.bruteforce(v, f, org, d, far, 0.0)
    _assert_same(t, i, ot, oi)
