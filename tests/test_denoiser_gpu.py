"""drp_conv3x3 (tcgen05 TF32 implicit GEMM) against plain fp32 torch convolutions on the GPU."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from diffrp_b200 import _abi, denoiser as dn

pytestmark = pytest.mark.gpu
# stated tolerance: TF32 operands (10-bit mantissa, like cuDNN's default for fp32 convolutions) with fp32 accumulation:
# |err| <= 2e-3 * max|reference| over the layer output
TF32_TOL = 2e-3


def tf32_trunc(x):
    return (x.view(torch.int32) & ~0x1fff).view(torch.float32)


def reference_layer(x_nhwc, w, b, mode, relu):
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    y = F.conv2d(x_nhwc.permute(2, 0, 1)[None], w, b, padding=1)
    if relu:
        y = F.relu(y)
    if mode == _abi.CONV_POOL2:
        y = F.max_pool2d(y, 2, 2)
    elif mode == _abi.CONV_UPSAMPLE2:
        y = F.interpolate(y, scale_factor=2.0, mode="nearest")
    return y[0].permute(1, 2, 0).contiguous()


@pytest.mark.parametrize("H,W,cin,cout,mode,relu", [
    (16, 16, 16, 32, _abi.CONV_PLAIN, True),
    (32, 48, 32, 48, _abi.CONV_POOL2, True),
    (24, 40, 160, 112, _abi.CONV_PLAIN, True),
    (8, 16, 96, 96, _abi.CONV_UPSAMPLE2, True),
    (6, 10, 80, 96, _abi.CONV_PLAIN, True),       # ragged tiles: neither dimension a multiple of the 8 x 16 tile
    (2, 2, 96, 96, _abi.CONV_UPSAMPLE2, True),    # smaller than one tile
    (40, 72, 32, 3, _abi.CONV_PLAIN, False),      # output layer: 3 (+1 zero: stores move 16-byte units) of 16 padded channels stored, no ReLU
])
def test_single_layer_matches_torch(H, W, cin, cout, mode, relu):
    g = torch.Generator(device='cuda').manual_seed(H * 1000 + W + cin + cout)
    in_stride, in_offset = cin + 32, 16                 # read a channel slice of a wider buffer
    buf = torch.randn(H, W, in_stride, device='cuda', generator=g)
    w = torch.randn(cout, cin, 3, 3, device='cuda', generator=g) / math_sqrt(9 * cin)
    b = torch.randn(cout, device='cuda', generator=g) * 0.1
    wm, bm = dn.pack_weight(w, b, list(range(cin)), cin)
    oh, ow = {_abi.CONV_PLAIN: (H, W), _abi.CONV_POOL2: (H // 2, W // 2), _abi.CONV_UPSAMPLE2: (2 * H, 2 * W)}[mode]
    out_stride, out_offset = dn._pad16(cout) + 16, 4 if cout == 3 else 16
    out = torch.full((oh, ow, out_stride), -7.0, device='cuda')
    stored = (cout + 3) // 4 * 4
    dn.conv3x3(buf, in_offset, cin, wm, bm, out, out_offset, stored, mode, relu)
    torch.cuda.synchronize()
    x = buf[..., in_offset:in_offset + cin].contiguous()
    ref = reference_layer(x, w, b, mode, relu)
    got = out[..., out_offset:out_offset + cout]
    scale = ref.abs().max().item()
    err = (got - ref).abs().max().item()
    ref_t = reference_layer(tf32_trunc(x), tf32_trunc(w), b, mode, relu)
    err_t = (got - ref_t).abs().max().item()
    print("max|err| vs fp32 %.3e (scale %.3f), vs tf32-truncated operands %.3e" % (err, scale, err_t))
    assert err <= TF32_TOL * scale, (err, scale)
    # nothing outside the slice was touched
    assert (out[..., :out_offset] == -7.0).all() and (out[..., out_offset + stored:] == -7.0).all()
    assert (out[..., out_offset + cout:out_offset + stored] == 0.0).all()      # padded output channels: zero weights, zero bias


def math_sqrt(v):
    return float(np.sqrt(v))


# ---- whole network ------------------------------------------------------------------------------------------------------------
def torch_unet(sd, x, trunc=False):
    """Plain torch restatement of UNet.forward (denoiser.py:117-173), NCHW fp32; ``trunc`` rounds every conv operand to the nearest TF32,
    which is what the kernels do (weights at pack time, activations in the producing layer's epilogue)."""
    q = dn.round_tf32 if trunc else (lambda v: v)

    def conv(name, v, relu=True):
        v = F.conv2d(q(v.contiguous()), q(sd[name + ".weight"].contiguous()), sd[name + ".bias"], padding=1)
        return F.relu(v) if relu else v
    pool = lambda v: F.max_pool2d(v, 2, 2)                                     # noqa: E731
    up = lambda v: F.interpolate(v, scale_factor=2.0, mode="nearest")         # noqa: E731
    inp = x
    x = conv("enc_conv0", inp)
    x = p1 = pool(conv("enc_conv1", x))
    x = p2 = pool(conv("enc_conv2", x))
    x = p3 = pool(conv("enc_conv3", x))
    x = pool(conv("enc_conv4", x))
    x = conv("enc_conv5b", conv("enc_conv5a", x))
    x = conv("dec_conv4b", conv("dec_conv4a", torch.cat([up(x), p3], 1)))
    x = conv("dec_conv3b", conv("dec_conv3a", torch.cat([up(x), p2], 1)))
    x = conv("dec_conv2b", conv("dec_conv2a", torch.cat([up(x), p1], 1)))
    x = conv("dec_conv1b", conv("dec_conv1a", torch.cat([up(x), inp], 1)))
    return conv("dec_conv0", x, relu=False)


# stated tolerance for the whole net: 16 chained TF32 layers, relative to max|output|
NET_TOL_FP32 = 5e-3
NET_TOL_TF32_EMULATED = 2e-3   # re-rounding after every layer amplifies accumulation-order differences


@pytest.mark.parametrize("H,W", [(16, 16), (64, 96), (128, 80)])
def test_unet_matches_torch(H, W):
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    weights, sd = dn.UNetWeights.random(seed=3)
    sd = {k: v.cuda() for k, v in sd.items()}
    net = dn.UNet(weights)
    g = torch.Generator(device='cuda').manual_seed(H + W)
    x = dn.round_tf32(torch.rand(H, W, 9, device='cuda', generator=g))     # drp_denoise_pack rounds the network input the same way
    buf, off = net.input_slice(H, W)
    buf[..., off:off + 9] = x
    out = net.forward(H, W)[..., :3]
    ref = torch_unet(sd, x.permute(2, 0, 1)[None])[0].permute(1, 2, 0)
    ref_t = torch_unet(sd, x.permute(2, 0, 1)[None], trunc=True)[0].permute(1, 2, 0)
    scale = ref.abs().max().item()
    e32, et = (out - ref).abs().max().item() / scale, (out - ref_t).abs().max().item() / scale
    print("UNet %dx%d: rel err vs fp32 %.3e, vs TF32-emulated %.3e" % (H, W, e32, et))
    assert e32 <= NET_TOL_FP32 and et <= NET_TOL_TF32_EMULATED
    # a second forward on the same buffers (stale upsample / skip slices are fully overwritten) gives the same result
    out2 = net.forward(H, W)[..., :3]
    assert torch.equal(out, out2)


def test_run_denoiser_matches_reference_golden():
    """Against the reference's own run_denoiser + UNet (CPU, fp32; tests/golden/make_golden.py:denoiser_fixture)."""
    import os
    g = dict(np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "denoiser.npz")))
    net = dn.get_denoiser(seed=int(g['seed']))
    cu = lambda a: torch.from_numpy(a).cuda()                                  # noqa: E731
    out = dn.run_denoiser(net, cu(g['hdr']), cu(g['albedo']), cu(g['normal']))
    assert out.shape == (41, 53, 3)

    def to_pu(o):                                                              # compare in the network's output domain
        P = dn._PU
        o = o.double()
        return torch.where(o <= P['Y0'], P['A'] * o, torch.where(o <= P['Y1'], P['B'] * o.clamp_min(1e-30) ** P['C'] + P['D'], P['E'] * torch.log(o.clamp_min(0) + P['F']) + P['G']))
    a, b = to_pu(out.cpu()), to_pu(torch.from_numpy(g['out']))
    rel = ((a - b).abs().max() / b.abs().max()).item()
    print("run_denoiser vs reference golden: rel err in the PU domain %.3e" % rel)
    assert rel <= NET_TOL_FP32


def test_denoiser_argument_errors():
    net = dn.get_denoiser(seed=0)
    x = torch.rand(8, 8, 3, device='cuda')
    with pytest.raises(ValueError):
        dn.run_denoiser(net, x.cpu(), x, x)
    with pytest.raises(ValueError):
        dn.run_denoiser(net, x, x, x, alignment=32)
    with pytest.raises(ValueError):
        net.forward(24, 40)


def test_documented_chain_pbr_denoise_tonemap():
    """docs/source/ptpbr.md:55-58 of the reference, with this package's names: pbr -> run_denoiser(pbr, srgb(albedo), normal) -> agx -> bytes."""
    import scenes
    import diffrp_b200 as drp
    g = dict(np.load(__import__('os').path.join(__import__('os').path.dirname(__import__('os').path.abspath(__file__)), "golden", "tonemap.npz")))
    scene, cam = scenes.mixed_scene(), drp.PerspectiveCamera(h=72, w=100)
    pbr, alpha, extras = drp.PathTracingSession(scene, cam, drp.PathTracingSessionOptions(ray_spp=4, ray_depth=3, seed=2)).pbr()
    denoiser = drp.get_denoiser(seed=5)
    den = drp.run_denoiser(denoiser, pbr, drp.linear_to_srgb(extras['albedo']), extras['world_normal'])
    assert den.shape == pbr.shape and torch.isfinite(den).all()
    img = drp.to_uint8(torch.cat([drp.agx_base_contrast(den, torch.from_numpy(g['lut']).cuda()), alpha], -1))
    assert img.shape == (72, 100, 4) and img.dtype == torch.uint8
    # same inputs, same bytes: the whole chain is deterministic
    den2 = drp.run_denoiser(denoiser, pbr, drp.linear_to_srgb(extras['albedo']), extras['world_normal'])
    assert torch.equal(den, den2)
