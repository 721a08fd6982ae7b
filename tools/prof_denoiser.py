"""One run_denoiser call at --res for ncu (tools: ncu --set full -k regex:k_conv3x3 ... python tools/prof_denoiser.py)."""
import os, sys
import torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from diffrp_b200 import denoiser as dn
R = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
net = dn.get_denoiser(seed=1)
x = torch.rand(R, R, 3, device='cuda')
for _ in range(2):
    dn.run_denoiser(net, x, x, x)
torch.cuda.synchronize()
