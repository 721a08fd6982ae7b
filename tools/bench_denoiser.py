"""Times diffrp_b200.denoiser.run_denoiser (16 tcgen05 conv launches) against the same U-Net run by torch/cuDNN (the library baseline the
reference uses: TF32 allowed, its default) on one GPU.  Output: one JSON line.  Usage: python tools/bench_denoiser.py [--res 1024]"""
import argparse
import json
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from diffrp_b200 import denoiser as dn  # noqa: E402


def torch_unet(sd, x):
    def conv(name, v, relu=True):
        v = F.conv2d(v, sd[name + ".weight"], sd[name + ".bias"], padding=1)
        return F.relu(v) if relu else v
    pool = lambda v: F.max_pool2d(v, 2, 2)                                  # noqa: E731
    up = lambda v: F.interpolate(v, scale_factor=2.0, mode="nearest")      # noqa: E731
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


def timed(fn, iters):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--res", type=int, default=1024)
    ap.add_argument("--iters", type=int, default=20)
    args = ap.parse_args()
    R = args.res
    weights, sd = dn.UNetWeights.random(seed=1)
    net = dn.UNet(weights)
    g = torch.Generator(device='cuda').manual_seed(0)
    hdr = torch.exp(torch.randn(R, R, 3, device='cuda', generator=g))
    alb = torch.rand(R, R, 3, device='cuda', generator=g)
    nrm = F.normalize(torch.randn(R, R, 3, device='cuda', generator=g), dim=-1)
    ms_b200 = timed(lambda: dn.run_denoiser(net, hdr, alb, nrm), args.iters)
    ms_net = timed(lambda: net.forward(R, R), args.iters)
    sdc = {k: v.cuda() for k, v in sd.items()}
    x = torch.rand(1, 9, R, R, device='cuda', generator=g)
    out = {}
    for tf32 in (True, False):
        torch.backends.cudnn.allow_tf32 = tf32
        torch.backends.cudnn.benchmark = True
        with torch.no_grad():
            out["cudnn_tf32" if tf32 else "cudnn_fp32"] = timed(lambda: torch_unet(sdc, x), args.iters)
            xc = x.contiguous(memory_format=torch.channels_last)
            sdl = {k: (v.contiguous(memory_format=torch.channels_last) if v.ndim == 4 else v) for k, v in sdc.items()}
            out[("cudnn_tf32" if tf32 else "cudnn_fp32") + "_channels_last"] = timed(lambda: torch_unet(sdl, xc), args.iters)
    macs = 0
    res = {"enc_conv0": 1, "enc_conv1": 1, "enc_conv2": 2, "enc_conv3": 4, "enc_conv4": 8, "enc_conv5a": 16, "enc_conv5b": 16, "dec_conv4a": 8, "dec_conv4b": 8,
           "dec_conv3a": 4, "dec_conv3b": 4, "dec_conv2a": 2, "dec_conv2b": 2, "dec_conv1a": 1, "dec_conv1b": 1, "dec_conv0": 1}
    for name, cin, cout in dn.LAYERS:
        macs += 9 * cin * cout * (R // res[name]) ** 2
    print(json.dumps({"res": R, "run_denoiser_ms": ms_b200, "unet_forward_ms": ms_net, "tflops_tf32": 2 * macs / (ms_net * 1e-3) / 1e12,
                      "gmacs": macs / 1e9, "torch_ms": out, "iters": args.iters}))


if __name__ == "__main__":
    main()
