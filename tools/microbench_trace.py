"""Config-2 microbenchmark: N random rays vs a ~1M-triangle displaced UV sphere. One GPU."""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from diffrp_b200 import synthetic as syn
from diffrp_b200.raycaster import B200Raycaster

ap = argparse.ArgumentParser()
ap.add_argument('--rays', type=int, default=16 * 2 ** 20)
ap.add_argument('--theta', type=int, default=1024)
ap.add_argument('--phi', type=int, default=512)
ap.add_argument('--iters', type=int, default=5)
ap.add_argument('--coherent', action='store_true', help='sort rays by origin/direction cell first')
args = ap.parse_args()
v, f = syn.uv_sphere(args.theta, args.phi)
o, d = syn.random_rays(args.rays)
V, F = torch.from_numpy(v).cuda(), torch.from_numpy(f).cuda()
O, D = torch.from_numpy(o).cuda(), torch.from_numpy(d).cuda()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
torch.cuda.synchronize()
ev[0].record(); rc = B200Raycaster(V, F); ev[1].record(); torch.cuda.synchronize()
build_ms = ev[0].elapsed_time(ev[1])
ev[0].record(); rc2 = B200Raycaster(V, F); ev[1].record(); torch.cuda.synchronize()
build_ms2 = ev[0].elapsed_time(ev[1])
far = 10.0
for _ in range(2):
    t, i = rc.query(O, D, far)
times = []
for _ in range(args.iters):
    ev[0].record(); t, i = rc.query(O, D, far); ev[1].record(); torch.cuda.synchronize()
    times.append(ev[0].elapsed_time(ev[1]))
ms = min(times)
B_query = 32 + 48 * int(np.ceil(np.log2(len(f)))) + 36
print(json.dumps(dict(tris=len(f), rays=args.rays, build_ms_first=build_ms, build_ms=build_ms2, trace_ms=ms, trace_ms_all=times,
                      mrays_s=args.rays / ms / 1e3, hit_frac=float((t < far).float().mean()),
                      algo_GBs=args.rays * B_query / ms / 1e6, stats=rc.stats())))
