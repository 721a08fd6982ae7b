#!/bin/bash
# compute-sanitizer over a small end-to-end run (LBVH build, collapse, refit, instanced assembly, trace with and without the deep-stack fix-up,
# fused wavefront in every RNG / shard mode, finalize, tone map, tcgen05 denoiser).
# Usage (GPU box): bash tools/sanitize.sh [memcheck|racecheck|synccheck]   -> gpurun_out/sanitizer_<tool>.log
TOOL=${1:-memcheck}
mkdir -p gpurun_out
compute-sanitizer --tool $TOOL --error-exitcode 3 python - <<'PY' > gpurun_out/sanitizer_$TOOL.log 2>&1
import sys, os, torch, numpy as np
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import scenes, diffrp_b200 as drp
from diffrp_b200 import synthetic as syn
from diffrp_b200._lib import lib, check
v, f = syn.uv_sphere(64, 32); o, d = syn.random_rays(20000)
V, F, O, D = torch.from_numpy(v).cuda(), torch.from_numpy(f).cuda(), torch.from_numpy(o).cuda(), torch.from_numpy(d).cuda()
rc = drp.B200Raycaster(V, F)
t, i = rc.query(O, D, 10.0)
check(lib().drp_debug_set_stack_limit(rc.handle, 1), "stack limit")      # most rays through k_extend_fixup
t2, i2 = rc.query(O, D, 10.0)
assert torch.equal(t, t2) and torch.equal(i, i2)
rc.refit((V * 1.1).contiguous())                                          # k_cw_refit_level
t3, _ = rc.query(O, D, 10.0)
isc, camkw = syn.instanced_scene('cpu', n_instances=12, mesh_res=(8, 6), env_res=(8, 16), spread=(0.5, 0.3, 0.3))
isc = isc.to(torch.device('cuda'))
si = drp.PathTracingSession(isc, drp.PerspectiveCamera.from_orbit(h=24, w=32, **camkw), drp.PathTracingSessionOptions(ray_spp=2, ray_depth=2, instancing=True))
ri = si.pbr()[0]                                                          # drp_build_instanced + render
assert si.raycaster().instanced
cam = drp.PerspectiveCamera.from_orbit(h=48, w=64, radius=3.0, azim=25, elev=15, origin=[0.0, -0.1, 0.0], fov=32)
for opt in (dict(rng='native'), dict(rng='torch', pbr_ray_last_bounce='skybox'), dict(shard_rank=1, shard_world=2, shard_mode='tile', tile_size=32)):
    s = drp.PathTracingSession(scenes.mixed_scene(), cam, drp.PathTracingSessionOptions(ray_spp=2, ray_depth=3, **opt))
    acc = s.render_accumulators(); out = s.finalize(acc)
# colour epilogue + denoiser (tcgen05 / TMA kernels): odd sizes so that clipped tiles and reflection padding are exercised
lut = torch.rand(8, 8, 8, 3, device='cuda')
img = drp.PathTracingSession(scenes.mixed_scene(), cam, drp.PathTracingSessionOptions(ray_spp=1, ray_depth=2)).pbr_image('agx', lut=lut)
den = drp.get_denoiser(seed=1)
rad, alpha, extras = drp.PathTracingSession(scenes.mixed_scene(), cam, drp.PathTracingSessionOptions(ray_spp=1, ray_depth=2)).pbr()
dn = drp.run_denoiser(den, rad[:41, :53].contiguous(), drp.linear_to_srgb(extras['albedo'][:41, :53].contiguous()), extras['world_normal'][:41, :53].contiguous())
torch.cuda.synchronize()
print("SANITIZER_RUN_COMPLETE", float(t.mean()), float(t3.mean()), float(ri.mean()), float(out[0].mean()), float(dn.mean()), int(img.sum()))
PY
echo "exit code $?" >> gpurun_out/sanitizer_$TOOL.log
tail -8 gpurun_out/sanitizer_$TOOL.log
