"""BASELINE configs[1] at full size: 16,777,216 random rays vs the 1,048,576-triangle displaced sphere; primary-hit ids and t
bit-exact against the CPU oracle (BVH restatement) on every ray.  Also runs the binary-layout / greedy-collapse A/B variants
of the traversal structure through the same check (results must not depend on the hierarchy)."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

import oracle
from diffrp_b200 import synthetic as syn
from diffrp_b200.raycaster import B200Raycaster

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_config2_full_size_bit_exact():
    v, f = syn.uv_sphere(1024, 512)
    o, d = syn.random_rays(16 * 2 ** 20)
    rc = B200Raycaster(torch.from_numpy(v).cuda(), torch.from_numpy(f).cuda(), {'epsilon': 1e-8})
    t, i = rc.query(torch.from_numpy(o).cuda(), torch.from_numpy(d).cuda(), 10.0)
    t, i = t.cpu().numpy(), i.cpu().numpy()
    ot, oi = oracle.BVH(v, f).query(o, d, 10.0, 1e-8)
    assert np.array_equal(t.view(np.int32), ot.view(np.int32))
    assert np.array_equal(i, oi)
    assert 0.7 < (t < 10.0).mean() < 0.8


@pytest.mark.parametrize("cap", [0, 1, 3, 8])
def test_deep_stack_fixup_gives_identical_hits(cap):
    """Rays whose traversal outgrows the per-thread stack are re-traced by k_extend_fixup (256-entry stack in global memory).  With the fast
    path's stack lowered to `cap` entries (drp_debug_set_stack_limit) most rays of an ordinary scene take that route; hits must stay bit-identical
    to the oracle, and the handle must stay healthy (drp_status)."""
    from diffrp_b200._lib import lib, check
    v, f = syn.uv_sphere(256, 128)
    o, d = syn.random_rays(400_000, seed=9)
    rc = B200Raycaster(torch.from_numpy(v).cuda(), torch.from_numpy(f).cuda(), {'epsilon': 1e-8})
    ref_t, ref_i = rc.query(torch.from_numpy(o).cuda(), torch.from_numpy(d).cuda(), 10.0)
    check(lib().drp_debug_set_stack_limit(rc.handle, cap), "drp_debug_set_stack_limit")
    t, i = rc.query(torch.from_numpy(o).cuda(), torch.from_numpy(d).cuda(), 10.0)
    torch.cuda.synchronize()
    assert torch.equal(t.view(torch.int32), ref_t.view(torch.int32)) and torch.equal(i, ref_i)
    ot, oi = oracle.BVH(v, f).query(o, d, 10.0, 1e-8)
    assert np.array_equal(t.cpu().numpy().view(np.int32), ot.view(np.int32)) and np.array_equal(i.cpu().numpy(), oi)
    assert int(i.min()) >= 0  # no sentinel survives
    check(lib().drp_status(rc.handle), "drp_status")


def test_degenerate_deep_scene_is_exact():
    """A hierarchy as deep as the builder can make it: triangles whose centroids form a geometric progression (one Morton bit splits off one
    triangle at a time) plus thousands of exact duplicates (ties resolved by the index bits).  Closest hits stay the brute-force ones."""
    rng = np.random.default_rng(5)
    tri = np.array([[0.0, 0.0, 0.0], [1.0, 0.0, 0.0], [0.0, 1.0, 0.0]], np.float32) * 1e-3
    offs = [np.full(3, 2.0 ** -k, np.float32) for k in range(1, 22)] * 3
    offs += [np.full(3, 0.3, np.float32)] * 6000
    verts = np.concatenate([tri + o for o in offs]).astype(np.float32)
    faces = np.arange(len(verts), dtype=np.int32).reshape(-1, 3)
    o = rng.uniform(-0.1, 1.1, (200_000, 3)).astype(np.float32)
    tgt = np.stack(offs)[rng.integers(0, len(offs), len(o))] + rng.uniform(0, 1e-3, (len(o), 3)).astype(np.float32) * [1, 1, 0]
    d = (tgt - o).astype(np.float32)
    rc = B200Raycaster(torch.from_numpy(verts).cuda(), torch.from_numpy(faces).cuda(), {'epsilon': 0.0})
    t, i = rc.query(torch.from_numpy(o).cuda(), torch.from_numpy(d).cuda(), 10.0)
    bt, bi = rc.query_bruteforce(torch.from_numpy(o).cuda(), torch.from_numpy(d).cuda(), 10.0)
    assert torch.equal(t.view(torch.int32), bt.view(torch.int32)) and torch.equal(i, bi)
    assert float((t < 10.0).float().mean()) > 0.2
    from diffrp_b200._lib import lib, check
    check(lib().drp_status(rc.handle), "drp_status")
    assert rc.stats()["max_depth"] >= 4


def test_fixup_scan_mode_when_more_rays_overflow_than_the_list_holds():
    """More than 2^20 flagged rays in one launch: the overflow list is incomplete and k_extend_fixup scans the results for the sentinel id
    instead.  3 M rays with the fast path's stack at 0 entries (every ray that would push once is flagged): bit-identical to the normal run."""
    from diffrp_b200._lib import lib, check
    v, f = syn.uv_sphere(96, 48)
    o, d = syn.random_rays(3_000_000, seed=12)
    to, td = torch.from_numpy(o).cuda(), torch.from_numpy(d).cuda()
    rc = B200Raycaster(torch.from_numpy(v).cuda(), torch.from_numpy(f).cuda(), {'epsilon': 1e-8})
    ref_t, ref_i = rc.query(to, td, 10.0)
    check(lib().drp_debug_set_stack_limit(rc.handle, 0), "drp_debug_set_stack_limit")
    t, i = rc.query(to, td, 10.0)
    assert torch.equal(t.view(torch.int32), ref_t.view(torch.int32)) and torch.equal(i, ref_i)
    assert int((ref_t < 10.0).sum()) > (1 << 20)       # more hits than list entries: the scan path really ran
    assert int(i.min()) >= 0
    check(lib().drp_status(rc.handle), "drp_status")
