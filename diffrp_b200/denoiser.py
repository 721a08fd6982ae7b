"""
OIDN-style U-Net denoiser on the tcgen05 tensor cores -- the step the reference runs right after ``pbr()``
(diffrp/rendering/denoiser.py; docs/source/ptpbr.md:55-58: radiance + albedo + world normal in, radiance out).

Every layer of ``UNet.forward`` (denoiser.py:117-173) is ONE launch of ``drp_conv3x3`` (csrc/conv3x3.cu): 3x3 convolution + bias + ReLU with the
following ``pool`` / ``upsample`` + ``concat`` fused into its epilogue -- pooled encoder outputs and upsampled decoder outputs are written
straight into channel slices of the NHWC concatenation buffers the next layer reads, so no activation is ever copied.

The network weights of the reference (``resources/denoisers/rt_hdr_alb_nrm.pt``, Intel OIDN's) are a data file that is not redistributed
here: construct ``UNetWeights`` from any ``state_dict`` with the reference's parameter names (``get_denoiser().state_dict()`` of an
installed diffrp) or ``UNetWeights.random(seed)`` for tests and benchmarks.  No CPU path.
"""
import ctypes as C
import math

import torch

from . import _abi
from ._lib import lib, check

# (name, cin parts, cout) in forward order; channel counts of the reference's full-size net (denoiser.py:89-101)
_IC, _EC1, _EC2, _EC3, _EC4, _EC5, _DC4, _DC3, _DC2A, _DC2B, _DC1A, _DC1B, _OC = 9, 32, 48, 64, 80, 96, 112, 96, 64, 64, 64, 32, 3
LAYERS = (
    ("enc_conv0", _IC, _EC1), ("enc_conv1", _EC1, _EC1), ("enc_conv2", _EC1, _EC2), ("enc_conv3", _EC2, _EC3), ("enc_conv4", _EC3, _EC4),
    ("enc_conv5a", _EC4, _EC5), ("enc_conv5b", _EC5, _EC5), ("dec_conv4a", _EC5 + _EC3, _DC4), ("dec_conv4b", _DC4, _DC4),
    ("dec_conv3a", _DC4 + _EC2, _DC3), ("dec_conv3b", _DC3, _DC3), ("dec_conv2a", _DC3 + _EC1, _DC2A), ("dec_conv2b", _DC2A, _DC2B),
    ("dec_conv1a", _DC2B + _IC, _DC1A), ("dec_conv1b", _DC1A, _DC1B), ("dec_conv0", _DC1B, _OC),
)


def _pad16(c):
    return (c + 15) // 16 * 16


def _stream_ptr(dev):
    return torch.cuda.current_stream(dev).cuda_stream


def pack_weight(w: torch.Tensor, b: torch.Tensor, channel_map, cin_buf: int):
    """(cout, cin, 3, 3) + (cout,) -> ([cout_pad][9*cin_buf] with k = tap*cin_buf + buffer channel, [cout_pad]); ``channel_map[c]`` is the
    buffer channel that holds the convolution's input channel c (padding channels of the buffer get zero weights)."""
    cout, cin = w.shape[0], w.shape[1]
    cout_pad = _pad16(cout)
    m = torch.zeros(cout_pad, 9, cin_buf, dtype=torch.float32, device=w.device)
    m[:cout, :, torch.as_tensor(channel_map, device=w.device)] = w.float().permute(0, 2, 3, 1).reshape(cout, 9, cin)
    bias = torch.zeros(cout_pad, dtype=torch.float32, device=w.device)
    bias[:cout] = b.float()
    return m.reshape(cout_pad, 9 * cin_buf).contiguous(), bias


def conv3x3(inp: torch.Tensor, in_offset: int, cin: int, weight: torch.Tensor, bias: torch.Tensor, out: torch.Tensor, out_offset: int, cout_store: int,
            mode: int = _abi.CONV_PLAIN, relu: bool = True):
    """One ``drp_conv3x3`` launch.  ``inp`` (H, W, in_stride) and ``out`` (H', W', out_stride) are NHWC fp32 CUDA buffers; the layer reads
    channels [in_offset, in_offset+cin) and writes [out_offset, out_offset+cout_store)."""
    if not (inp.is_cuda and out.is_cuda and weight.is_cuda and bias.is_cuda):
        raise RuntimeError("diffrp_b200.denoiser: CUDA tensors required (there is no CPU fallback)")
    H, W, in_stride = inp.shape
    expect = {_abi.CONV_PLAIN: (H, W), _abi.CONV_POOL2: (H // 2, W // 2), _abi.CONV_UPSAMPLE2: (2 * H, 2 * W)}[mode]
    if tuple(out.shape[:2]) != expect:
        raise ValueError("output buffer has shape %s, expected %s" % (tuple(out.shape[:2]), expect))
    if weight.shape != (bias.shape[0], 9 * cin):
        raise ValueError("weight matrix must be [cout_pad][9*cin]")
    p = _abi.Conv3x3Params(inp.data_ptr(), weight.data_ptr(), bias.data_ptr(), out.data_ptr(), H, W, cin, in_stride, in_offset, bias.shape[0], cout_store,
                           out.shape[2], out_offset, mode, int(relu))
    check(lib().drp_conv3x3(C.byref(p), _stream_ptr(inp.device)), "drp_conv3x3")
