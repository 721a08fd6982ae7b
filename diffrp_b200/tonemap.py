"""
Colour epilogue after the path: AgX / sRGB tone mapping and 8-bit quantisation as ONE CUDA pass (``drp_tonemap``,
csrc/epilogue.cu) instead of the reference's chain of full-frame torch ops.

Same names, argument meaning and results as the reference functions they replace:

* ``linear_to_srgb(rgb)``            -- diffrp/utils/colors.py:33-42
* ``agx_base_contrast(rgb, lut)``    -- diffrp/utils/tone_mapping.py:21-35.  The reference loads its LUT from a packaged resource
  (``luts/agx-base-contrast.pt``, tone_mapping.py:13-14); that data file is not redistributed here, so the LUT is an argument:
  pass ``diffrp.utils.tone_mapping.agx_lut_loader.load("base-contrast")`` from an installed diffrp (see
  ``diffrp_b200.integration.agx_lut``), or any (n,n,n,3) LUT in the same z-y-x layout.
* ``to_uint8(rgb_or_rgba)``          -- the tensor half of to_pil, diffrp/utils/exchange.py:17 (``(saturate(x)*255).byte()``)

There is no CPU path: inputs must be CUDA tensors.
"""
import ctypes as C

import torch

from . import _abi
from ._lib import lib, check

_TONES = {None: _abi.TONE_LINEAR, 'linear': _abi.TONE_LINEAR, 'srgb': _abi.TONE_SRGB, 'agx': _abi.TONE_AGX}


def _stream_ptr(dev):
    return torch.cuda.current_stream(dev).cuda_stream


def tonemap(src: torch.Tensor, tone='agx', lut: torch.Tensor = None, scale: float = 1.0, alpha_offset: int = -1, flip_rows: bool = False,
            want_u8: bool = True, want_f32: bool = False):
    """``drp_tonemap`` on ``src`` (H, W, S) or (..., S) fp32 CUDA; returns ``(f32 or None, u8 or None)`` of shape (..., C) with
    C = 4 when ``alpha_offset >= 0`` else 3.  ``flip_rows`` needs a (H, W, S) source."""
    if tone not in _TONES:
        raise ValueError("tone must be one of 'agx', 'srgb', 'linear'")
    if not src.is_cuda:
        raise RuntimeError("diffrp_b200.tonemap: CUDA tensor required (there is no CPU fallback)")
    if src.dtype != torch.float32:
        raise TypeError("diffrp_b200.tonemap: fp32 input required")
    src = src.contiguous()
    stride = src.shape[-1]
    lead = tuple(src.shape[:-1])
    if flip_rows and len(lead) != 2:
        raise ValueError("flip_rows needs a (H, W, S) source")
    h, w = lead if len(lead) == 2 else (1, src.numel() // max(stride, 1))
    if _TONES[tone] == _abi.TONE_AGX:
        if lut is None:
            raise ValueError("tone='agx' needs the AgX LUT (n,n,n,3); see diffrp_b200.integration.agx_lut")
        if lut.ndim != 4 or lut.shape[3] != 3 or not (lut.shape[0] == lut.shape[1] == lut.shape[2]):
            raise ValueError("LUT must have shape (n, n, n, 3)")
        lut = lut.to(device=src.device, dtype=torch.float32).contiguous()
    else:
        lut = None
    c = 4 if alpha_offset >= 0 else 3
    out_f = torch.empty(lead + (c,), dtype=torch.float32, device=src.device) if want_f32 else None
    out_b = torch.empty(lead + (c,), dtype=torch.uint8, device=src.device) if want_u8 else None
    if src.numel() == 0:
        return out_f, out_b
    p = _abi.TonemapParams(_TONES[tone], 0 if lut is None else lut.shape[0], None if lut is None else lut.data_ptr(), stride, alpha_offset,
                           int(flip_rows), float(scale))
    check(lib().drp_tonemap(src.data_ptr(), h, w, C.byref(p), None if out_b is None else out_b.data_ptr(),
                            None if out_f is None else out_f.data_ptr(), _stream_ptr(src.device)), "drp_tonemap")
    return out_f, out_b


def linear_to_srgb(rgb: torch.Tensor) -> torch.Tensor:
    """colors.py:33-42; same shape as the input (last dimension 3)."""
    return tonemap(rgb, 'srgb', want_u8=False, want_f32=True)[0]


def agx_base_contrast(rgb: torch.Tensor, lut: torch.Tensor) -> torch.Tensor:
    """tone_mapping.py:21-35; HDR linear RGB (..., 3) -> LDR sRGB (..., 3)."""
    return tonemap(rgb, 'agx', lut=lut, want_u8=False, want_f32=True)[0]


def to_uint8(rgb_or_rgba: torch.Tensor) -> torch.Tensor:
    """exchange.py:17: clamp to [0, 1], scale by 255, truncate to bytes.  (..., 3) or (..., 4)."""
    a = 3 if rgb_or_rgba.shape[-1] == 4 else -1
    return tonemap(rgb_or_rgba, 'linear', alpha_offset=a)[1]
