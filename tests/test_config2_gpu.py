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


@pytest.mark.parametrize("env", [{"DRP_LAYOUT": "bvh2"}, {"DRP_COLLAPSE": "greedy"}, {"DRP_EXTEND": "simple"}])
def test_alternative_structures_give_identical_hits(env):
    """The A/B switches (binary layout, greedy collapse, non-persistent kernel) run in a subprocess (they are read once per
    process) and must reproduce the oracle bit for bit, like the default structure."""
    code = (
        "import numpy as np, torch, oracle\n"
        "from diffrp_b200 import synthetic as syn\n"
        "from diffrp_b200.raycaster import B200Raycaster\n"
        "v, f = syn.uv_sphere(256, 128); o, d = syn.random_rays(400000, seed=9)\n"
        "rc = B200Raycaster(torch.from_numpy(v).cuda(), torch.from_numpy(f).cuda(), {'epsilon': 1e-8})\n"
        "t, i = rc.query(torch.from_numpy(o).cuda(), torch.from_numpy(d).cuda(), 10.0)\n"
        "ot, oi = oracle.BVH(v, f).query(o, d, 10.0, 1e-8)\n"
        "assert np.array_equal(t.cpu().numpy().view(np.int32), ot.view(np.int32)) and np.array_equal(i.cpu().numpy(), oi)\n"
        "print('OK')\n")
    res = subprocess.run([sys.executable, "-c", code], cwd=ROOT, env=dict(os.environ, **env), capture_output=True, text=True)
    assert res.returncode == 0 and "OK" in res.stdout, res.stderr[-2000:]
