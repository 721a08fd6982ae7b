#!/bin/bash
# compute-sanitizer memcheck over a small end-to-end run (LBVH build, collapse, trace, fused wavefront, finalize).
# Usage (GPU box): bash tools/sanitize.sh   -> gpurun_out/sanitizer.log
set -e
mkdir -p gpurun_out
compute-sanitizer --tool memcheck --error-exitcode 3 python - <<'PY' > gpurun_out/sanitizer.log 2>&1
import sys, os, torch, numpy as np
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import scenes, diffrp_b200 as drp
from diffrp_b200 import synthetic as syn
v, f = syn.uv_sphere(64, 32); o, d = syn.random_rays(20000)
rc = drp.B200Raycaster(torch.from_numpy(v).cuda(), torch.from_numpy(f).cuda())
t, i = rc.query(torch.from_numpy(o).cuda(), torch.from_numpy(d).cuda(), 10.0)
cam = drp.PerspectiveCamera.from_orbit(h=48, w=64, radius=3.0, azim=25, elev=15, origin=[0.0, -0.1, 0.0], fov=32)
for opt in (dict(rng='native'), dict(rng='torch', pbr_ray_last_bounce='skybox'), dict(shard_rank=1, shard_world=2, shard_mode='tile', tile_size=32)):
    s = drp.PathTracingSession(scenes.mixed_scene(), cam, drp.PathTracingSessionOptions(ray_spp=2, ray_depth=3, **opt))
    acc = s.render_accumulators(); out = s.finalize(acc)
# colour epilogue + denoiser (tcgen05 / TMA kernels): odd sizes so that clipped tiles and reflection padding are exercised
lut = torch.rand(8, 8, 8, 3, device='cuda')
img = s.pbr_image('agx', lut=lut)
den = drp.get_denoiser(seed=1)
rad, alpha, extras = drp.PathTracingSession(scenes.mixed_scene(), cam, drp.PathTracingSessionOptions(ray_spp=1, ray_depth=2)).pbr()
dn = drp.run_denoiser(den, rad[:41, :53].contiguous(), drp.linear_to_srgb(extras['albedo'][:41, :53].contiguous()), extras['world_normal'][:41, :53].contiguous())
torch.cuda.synchronize()
print("SANITIZER_RUN_COMPLETE", float(t.mean()), float(out[0].mean()), float(dn.mean()), int(img.sum()))
PY
tail -5 gpurun_out/sanitizer.log
