"""Colour epilogue (drp_tonemap) on the GPU: against the reference's own outputs (tests/golden/tonemap.npz) and the oracle."""
import os

import numpy as np
import pytest
import torch

import oracle
import scenes
import diffrp_b200 as drp
from diffrp_b200 import tonemap as tm
from test_oracle_golden import load, check_bytes

pytestmark = pytest.mark.gpu
# stated tolerance (fp32; CUDA powf / log10f vs the reference's SLEEF / libm, each <= 2 ulp, through one LUT interpolation):
RTOL, ATOL = 4e-6, 4e-6


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def test_named_functions_match_reference_outputs():
    g = load("tonemap")
    rgb, lut = dev(g['rgb']), dev(g['lut'])
    np.testing.assert_allclose(tm.linear_to_srgb(rgb).cpu().numpy(), g['srgb'], rtol=RTOL, atol=ATOL)
    np.testing.assert_allclose(tm.agx_base_contrast(rgb, lut).cpu().numpy(), g['agx'], rtol=RTOL, atol=ATOL)
    rgba = torch.cat([dev(g['agx']), dev(g['alpha'])], -1)
    assert np.array_equal(tm.to_uint8(rgba).cpu().numpy(), g['byte_agx'])     # same floats in -> same bytes out, exactly
    assert np.array_equal(tm.to_uint8(dev(g['srgb'])).cpu().numpy(), g['byte_srgb'][:, :3])
    # batch shapes are kept
    assert tm.linear_to_srgb(rgb.view(60, 100, 3)).shape == (60, 100, 3)


@pytest.mark.parametrize("tone", ['agx', 'srgb', 'linear'])
def test_fused_accumulator_epilogue_equals_oracle(tone):
    g = load("tonemap")
    H, W, spp = 60, 100, 5
    acc = np.zeros((H, W, 16), np.float32)
    acc[..., :3] = g['rgb'].reshape(H, W, 3) * spp
    acc[..., 3] = g['alpha'].reshape(H, W) * spp
    f, b = tm.tonemap(dev(acc), tone, lut=dev(g['lut']), scale=1.0 / spp, alpha_offset=3, flip_rows=True, want_f32=True)
    fo, bo = oracle.tonemap(acc, tone, lut=g['lut'], scale=1.0 / spp, alpha_offset=3, flip_rows=True)
    np.testing.assert_allclose(f.cpu().numpy(), fo, rtol=RTOL, atol=ATOL)
    if tone == 'linear':
        assert np.array_equal(f.cpu().numpy().view(np.uint32), fo.view(np.uint32)) and np.array_equal(b.cpu().numpy(), bo)
    check_bytes(b.cpu().numpy(), bo)
    assert b.shape == (H, W, 4) and b.dtype == torch.uint8


def test_pbr_image_equals_pbr_then_reference_chain():
    g = load("tonemap")
    scene, cam = scenes.mixed_scene(), drp.PerspectiveCamera(h=40, w=56)
    opts = dict(ray_spp=4, ray_depth=3, seed=9)
    rad, alpha, _ = drp.PathTracingSession(scene, cam, drp.PathTracingSessionOptions(**opts)).pbr()
    img = drp.PathTracingSession(scene, cam, drp.PathTracingSessionOptions(**opts)).pbr_image('agx', lut=dev(g['lut']))
    assert img.shape == (40, 56, 4) and img.dtype == torch.uint8 and img.is_cuda
    _, want = oracle.tonemap(np.concatenate([rad.cpu().numpy(), alpha.cpu().numpy()], -1), 'agx', lut=g['lut'], alpha_offset=3)
    check_bytes(img.cpu().numpy(), want)
    img2 = drp.PathTracingSession(scene, cam, drp.PathTracingSessionOptions(**opts)).pbr_image('srgb')
    _, want2 = oracle.tonemap(np.concatenate([rad.cpu().numpy(), alpha.cpu().numpy()], -1), 'srgb', alpha_offset=3)
    check_bytes(img2.cpu().numpy(), want2)


def test_argument_errors_are_loud():
    x = torch.rand(8, 3, device='cuda')
    with pytest.raises(ValueError):
        tm.tonemap(x, 'agx')                       # no LUT
    with pytest.raises(ValueError):
        tm.tonemap(x, 'filmic')
    with pytest.raises(RuntimeError):
        tm.linear_to_srgb(torch.rand(8, 3))        # CPU tensor: no fallback
    with pytest.raises(TypeError):
        tm.linear_to_srgb(x.double())
    assert tm.linear_to_srgb(torch.empty(0, 3, device='cuda')).shape == (0, 3)
