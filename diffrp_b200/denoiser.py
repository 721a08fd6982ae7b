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


def round_tf32(x: torch.Tensor) -> torch.Tensor:
    """fp32 -> nearest TF32 (10-bit mantissa, ties away from zero like cvt.rna.tf32.f32), kept in fp32 storage."""
    return ((x.contiguous().view(torch.int32) + 0x1000) & ~0x1fff).view(torch.float32)


def pack_weight(w: torch.Tensor, b: torch.Tensor, channel_map, cin_buf: int, round_weights: bool = False):
    """(cout, cin, 3, 3) + (cout,) -> ([cout_pad][9*cin_buf] with k = tap*cin_buf + buffer channel, [cout_pad]); ``channel_map[c]`` is the
    buffer channel that holds the convolution's input channel c (padding channels of the buffer get zero weights)."""
    cout, cin = w.shape[0], w.shape[1]
    cout_pad = _pad16(cout)
    m = torch.zeros(cout_pad, 9, cin_buf, dtype=torch.float32, device=w.device)
    wf = round_tf32(w.float()) if round_weights else w.float()
    m[:cout, :, torch.as_tensor(channel_map, device=w.device)] = wf.permute(0, 2, 3, 1).reshape(cout, 9, cin)
    bias = torch.zeros(cout_pad, dtype=torch.float32, device=w.device)
    bias[:cout] = b.float()
    return m.reshape(cout_pad, 9 * cin_buf).contiguous(), bias


def conv3x3(inp: torch.Tensor, in_offset: int, cin: int, weight: torch.Tensor, bias: torch.Tensor, out: torch.Tensor, out_offset: int, cout_store: int,
            mode: int = _abi.CONV_PLAIN, relu: bool = True, round_tf32: bool = False):
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
                           out.shape[2], out_offset, mode, int(relu), int(round_tf32))
    check(lib().drp_conv3x3(C.byref(p), _stream_ptr(inp.device)), "drp_conv3x3")


class UNetWeights:
    """Parameters of the reference's ``UNet(9, 3)`` (denoiser.py:72-115), packed for ``drp_conv3x3``."""

    def __init__(self, state_dict, device='cuda'):
        self.device = torch.device(device)
        self.raw = {}
        for name, cin, cout in LAYERS:
            w, b = state_dict[name + ".weight"], state_dict[name + ".bias"]
            if tuple(w.shape) != (cout, cin, 3, 3) or tuple(b.shape) != (cout,):
                raise ValueError("%s: expected weight (%d, %d, 3, 3)" % (name, cout, cin))
            self.raw[name] = (w.detach().to(self.device, torch.float32), b.detach().to(self.device, torch.float32))
        ident = lambda n: list(range(n))  # noqa: E731
        # buffer channel that holds each input channel of the layer, and the channel count of the slice the layer reads
        maps = {name: (ident(cin), _pad16(cin)) for name, cin, _ in LAYERS}
        maps["dec_conv1a"] = (ident(_DC2B + _IC), _DC2B + _pad16(_IC))      # [upsampled dec_conv2b | input padded to 16]
        # operands are rounded to TF32 once, here and in the layer epilogues, so the tensor core's truncation is exact (unbiased rounding)
        self.packed = {name: pack_weight(*self.raw[name], *maps[name], round_weights=True) for name, _, _ in LAYERS}

    @staticmethod
    def random(seed: int = 0, device='cuda'):
        """Deterministic He-initialised parameters (CPU generator, so the same on every machine) -- stand-in for the OIDN weights file."""
        g = torch.Generator().manual_seed(seed)
        sd = {}
        for name, cin, cout in LAYERS:
            sd[name + ".weight"] = torch.randn(cout, cin, 3, 3, generator=g) * math.sqrt(2.0 / (9 * cin))
            sd[name + ".bias"] = torch.randn(cout, generator=g) * 0.05
        return UNetWeights(sd, device), sd


class UNet:
    """``UNet.forward`` of the reference (denoiser.py:117-173) as 16 ``drp_conv3x3`` launches over preallocated NHWC buffers."""

    def __init__(self, weights: UNetWeights):
        self.w = weights
        self._bufs = {}

    def _buffers(self, H, W):
        key = (H, W)
        if key not in self._bufs:
            dev = self.w.device
            z = lambda h, w, c: torch.zeros(h, w, c, dtype=torch.float32, device=dev)  # noqa: E731
            self._bufs = {key: dict(
                cat1=z(H, W, _DC2B + 16), e0=z(H, W, _EC1), cat2=z(H // 2, W // 2, _DC3 + _EC1), cat3=z(H // 4, W // 4, _DC4 + _EC2),
                cat4=z(H // 8, W // 8, _EC5 + _EC3), p4=z(H // 16, W // 16, _EC4), b5=z(H // 16, W // 16, _EC5), d4=z(H // 8, W // 8, _DC4),
                d3=z(H // 4, W // 4, _DC3), d2=z(H // 2, W // 2, _DC2A), d1a=z(H, W, _DC1A), d1b=z(H, W, _DC1B), out=z(H, W, 4))}
        return self._bufs[key]

    def input_slice(self, H, W):
        """(buffer, channel offset): where the (H, W, 16) network input (9 channels + zero padding) has to be written."""
        return self._buffers(H, W)['cat1'], _DC2B

    def forward(self, H, W):
        """Runs the net on the input previously written to ``input_slice``; returns the (H, W, 4) output buffer (3 channels used)."""
        if H % 16 or W % 16:
            raise ValueError("the U-Net needs height and width to be multiples of 16 (run_denoiser pads)")
        B, P = self._buffers(H, W), self.w.packed
        PL, PO, UP = _abi.CONV_PLAIN, _abi.CONV_POOL2, _abi.CONV_UPSAMPLE2

        def L(name, src, s_off, cin, dst, d_off, cout, mode=PL, relu=True):
            conv3x3(B[src], s_off, cin, *P[name], B[dst], d_off, cout, mode, relu, round_tf32=relu)  # every hidden activation feeds convolutions only
        L("enc_conv0", 'cat1', _DC2B, 16, 'e0', 0, _EC1)
        L("enc_conv1", 'e0', 0, _EC1, 'cat2', _DC3, _EC1, PO)            # pool1 -> skip slice of concat2
        L("enc_conv2", 'cat2', _DC3, _EC1, 'cat3', _DC4, _EC2, PO)       # pool2 -> concat3
        L("enc_conv3", 'cat3', _DC4, _EC2, 'cat4', _EC5, _EC3, PO)       # pool3 -> concat4
        L("enc_conv4", 'cat4', _EC5, _EC3, 'p4', 0, _EC4, PO)
        L("enc_conv5a", 'p4', 0, _EC4, 'b5', 0, _EC5)
        L("enc_conv5b", 'b5', 0, _EC5, 'cat4', 0, _EC5, UP)              # upsample4 -> concat4
        L("dec_conv4a", 'cat4', 0, _EC5 + _EC3, 'd4', 0, _DC4)
        L("dec_conv4b", 'd4', 0, _DC4, 'cat3', 0, _DC4, UP)
        L("dec_conv3a", 'cat3', 0, _DC4 + _EC2, 'd3', 0, _DC3)
        L("dec_conv3b", 'd3', 0, _DC3, 'cat2', 0, _DC3, UP)
        L("dec_conv2a", 'cat2', 0, _DC3 + _EC1, 'd2', 0, _DC2A)
        L("dec_conv2b", 'd2', 0, _DC2A, 'cat1', 0, _DC2B, UP)
        L("dec_conv1a", 'cat1', 0, _DC2B + 16, 'd1a', 0, _DC1A)
        L("dec_conv1b", 'd1a', 0, _DC1A, 'd1b', 0, _DC1B)
        L("dec_conv0", 'd1b', 0, _DC1B, 'out', 0, 4, PL, False)          # 3 channels + one zero (stores move 16-byte units)
        return B['out']


# PU transfer function constants of the reference (utils/colors.py:5-15)
_PU = dict(A=1.41283765e+03, B=1.64593172e+00, C=4.31384981e-01, D=-2.94139609e-03, E=1.92653254e-01, F=6.26026094e-03, G=9.98620152e-01,
           Y0=1.57945760e-06, Y1=3.22087631e-02, X0=2.23151711e-03, X1=3.70974749e-01)


def get_denoiser(state_dict=None, seed: int = 0) -> UNet:
    """``get_denoiser`` (denoiser.py:16-21).  ``state_dict``: the reference's OIDN parameters; None -> seeded random stand-in."""
    return UNet(UNetWeights(state_dict) if state_dict is not None else UNetWeights.random(seed)[0])


def run_denoiser(denoiser: UNet, pbr_hdr: torch.Tensor, albedo_srgb: torch.Tensor, normal: torch.Tensor, alignment: int = 16) -> torch.Tensor:
    """``run_denoiser`` (denoiser.py:24-35): (H, W, 3) HDR radiance + sRGB albedo + world normal -> denoised (H, W, 3) radiance.
    PU-encode + concatenate + reflection-pad in one kernel, 16 tensor-core layers, crop + PU-decode in one kernel."""
    if alignment != 16:
        raise ValueError("alignment is fixed to 16 by the network's four pooling levels")
    h, w = pbr_hdr.shape[:2]
    for t in (pbr_hdr, albedo_srgb, normal):
        if not t.is_cuda or t.dtype != torch.float32 or tuple(t.shape) != (h, w, 3):
            raise ValueError("run_denoiser: three (H, W, 3) fp32 CUDA tensors are required")
    H, W = math.ceil(h / 16) * 16, math.ceil(w / 16) * 16
    buf, off = denoiser.input_slice(H, W)
    check(lib().drp_denoise_pack(pbr_hdr.contiguous().data_ptr(), albedo_srgb.contiguous().data_ptr(), normal.contiguous().data_ptr(), h, w,
                                 buf.data_ptr(), H, W, buf.shape[2], off, _stream_ptr(buf.device)), "drp_denoise_pack")
    out = denoiser.forward(H, W)
    res = torch.empty(h, w, 3, dtype=torch.float32, device=buf.device)
    check(lib().drp_denoise_unpack(out.data_ptr(), H, W, out.shape[2], res.data_ptr(), h, w, _stream_ptr(buf.device)), "drp_denoise_unpack")
    return res
