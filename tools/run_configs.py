"""
BASELINE.json configs[3] and configs[4] at full size (parity-test / capability cases, not bench lines):

  --config 4   multi-view datagen: 64 orbit cameras around a 500,000-triangle displaced sphere, 512x512, 64 spp, 3 bounces,
               radiance + albedo + world_normal AOVs; VIEWS sharded across ranks (no reduction, results gathered).
  --config 5   3840x2160, 256 spp, 3 bounces, env-lit, 10,000,000 triangles = 1000 MeshObjects sharing one 10k-triangle mesh
               with seeded rigid transforms and 8 tints; TILE-sharded across ranks with an NCCL all-reduce of the accumulator.

    python tools/run_configs.py --config 5            # one GPU
    torchrun --nproc-per-node 8 tools/run_configs.py --config 5
"""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
import diffrp_b200 as drp
from diffrp_b200 import synthetic as syn

ap = argparse.ArgumentParser()
ap.add_argument("--config", type=int, required=True, choices=[4, 5])
ap.add_argument("--scale", type=float, default=1.0, help="scale spp (and views for config 4) for quick runs")
ap.add_argument("--no-reuse", action="store_true", help="config 4: rebuild the BVH for every view (the reference's behaviour)")
args = ap.parse_args()
world, rank, local = int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
env = syn.torch_noise_texture(256, 512, 3, 7, 0.0, 1.0).to(dev) ** 3 * 5.0 + 0.1


def rigid(rng, scale, t):
    q, _ = np.linalg.qr(rng.standard_normal((3, 3)))
    if np.linalg.det(q) < 0:
        q[:, 0] = -q[:, 0]
    m = np.eye(4, dtype=np.float32); m[:3, :3] = q * scale; m[:3, 3] = t
    return T(m)


ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
if args.config == 4:
    v, f, n, uv, tg = syn.uv_sphere(500, 500, radius=0.8, bump=0.05, noise=0.01, seed=0, with_attrs=True)
    col = torch.rand(len(v), 4, generator=torch.Generator().manual_seed(3)).to(dev) * 0.6 + 0.4
    scene = drp.Scene().add_mesh_object(drp.MeshObject(drp.DefaultMaterial(), T(v), T(f), normals=T(n), color=col, uv=T(uv)))
    scene.add_light(drp.ImageEnvironmentLight(1.0, torch.ones(3, device=dev), env))
    n_views, spp, depth = max(world, int(64 * args.scale)), max(1, int(64 * args.scale)), 3
    views = list(range(n_views))[rank::world]
    torch.cuda.synchronize(); ev0.record()
    outs, traced = [], 0
    for k in views:
        cam = drp.PerspectiveCamera.from_orbit(h=512, w=512, radius=3.0, azim=360.0 * k / n_views, elev=20.0 * np.sin(2 * np.pi * k / n_views), origin=[0, 0, 0])
        sess = drp.PathTracingSession(scene, cam, drp.PathTracingSessionOptions(ray_spp=spp, ray_depth=depth, seed=k, reuse_scene=not args.no_reuse))
        rad, alpha, extras = sess.pbr()  # a session per view (single-use, like the reference); the scene cache keeps flatten + BVH
        outs.append(torch.cat([rad, extras['albedo'], extras['world_normal']], -1))
    ev1.record(); torch.cuda.synchronize()
    ms = torch.tensor([ev0.elapsed_time(ev1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    imgs = torch.stack(outs)
    assert torch.isfinite(imgs).all() and imgs[..., 3:6].max() <= 1.0 + 1e-5
    nominal = n_views * 512 * 512 * spp * depth
    res = dict(config=4, n_gpus=world, views=n_views, triangles=int(len(f)), spp=spp, seconds=ms.item() / 1e3, mrays_s=nominal / ms.item() / 1e3,
               views_per_s=n_views / (ms.item() / 1e3), mean_radiance=float(imgs[..., :3].mean()))
else:
    v, f, n, uv, tg = syn.uv_sphere(100, 50, radius=0.045, bump=0.004, noise=0.001, seed=1, with_attrs=True)
    Vt, Ft, Nt = T(v), T(f), T(n)
    tints = [torch.tensor(c, device=dev) for c in ([1, .3, .3], [.3, 1, .3], [.3, .3, 1], [1, 1, .3], [1, .3, 1], [.3, 1, 1], [.9, .9, .9], [.5, .5, .5])]
    mats = [drp.DefaultMaterial(t) for t in tints]
    rng = np.random.default_rng(0)
    scene = drp.Scene()
    for k in range(1000):
        pos = rng.uniform([-1.6, -0.9, -1.0], [1.6, 0.9, 1.0])
        scene.add_mesh_object(drp.MeshObject(mats[k % 8], Vt, Ft, normals=Nt, M=rigid(rng, 0.6 + rng.random(), pos)))
    scene.add_light(drp.ImageEnvironmentLight(1.0, torch.ones(3, device=dev), env))
    H, W, spp, depth = 2160, 3840, max(1, int(256 * args.scale)), 3
    cam = drp.PerspectiveCamera.from_orbit(h=H, w=W, radius=4.0, azim=20.0, elev=10.0, origin=[0, 0, 0], fov=35)
    sess = drp.PathTracingSession(scene, cam, drp.PathTracingSessionOptions(ray_spp=spp, ray_depth=depth, seed=2, shard_rank=rank, shard_world=world,
                                                                          shard_mode='tile', tile_size=256))
    t0 = time.perf_counter(); sess.raycaster(); torch.cuda.synchronize(); build_s = time.perf_counter() - t0
    torch.cuda.synchronize(); ev0.record()
    rad, alpha, extras = sess.pbr()
    ev1.record(); torch.cuda.synchronize()
    ms = torch.tensor([ev0.elapsed_time(ev1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    assert torch.isfinite(rad).all()
    nominal = H * W * spp * depth
    res = dict(config=5, n_gpus=world, triangles=int(sess.vertex_array_object().tris.shape[0]), resolution=[W, H], spp=spp, flatten_and_build_s=build_s,
               seconds=ms.item() / 1e3, mrays_s=nominal / ms.item() / 1e3, coverage=float((alpha > 0).float().mean()), mean_radiance=float(rad.mean()),
               bvh=sess.raycaster().stats())
if rank == 0:
    print(json.dumps(res))
if world > 1:
    dist.destroy_process_group()
