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
    (40, 72, 32, 3, _abi.CONV_PLAIN, False),      # output layer: 3 of 16 padded channels stored, no ReLU
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
    dn.conv3x3(buf, in_offset, cin, wm, bm, out, out_offset, cout, mode, relu)
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
    assert (out[..., :out_offset] == -7.0).all() and (out[..., out_offset + cout:] == -7.0).all()


def math_sqrt(v):
    return float(np.sqrt(v))
