"""
BASELINE.json configs[3] and configs[4] at full size (parity-test / capability cases, not bench lines; reduced-size parity against the oracle:
tests/test_configs_gpu.py):

  --config 4   multi-view datagen: 64 orbit cameras around a 500,000-triangle displaced sphere, 512x512, 64 spp, 3 bounces,
               radiance + albedo + world_normal AOVs; VIEWS sharded across ranks (no reduction, results stay on the rank that rendered them).
  --config 5   3840x2160, 256 spp, 3 bounces, env-lit, 10,000,000 triangles = 1000 MeshObjects sharing one 10k-triangle mesh
               with seeded rigid transforms and 8 tints; TILE-sharded across ranks; exchange = all-gather of the owned tiles
               (--tile-collective allreduce: sum of whole frames, for the A/B).

    python tools/run_configs.py --config 5            # one GPU
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/run_configs.py --config 5 --out gpurun_out/c5_n8.json

Scenes are the seeded generators of diffrp_b200/synthetic.py (datagen_scene, instanced_scene), built on the host and moved to the GPU before the
timed region (device-resident, like bench.py's `value`).  Roofline: SURVEY 8(d) per-bounce bytes B_bounce = B_query(T) + S_default.
"""
import argparse, json, math, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import diffrp_b200 as drp
from diffrp_b200 import synthetic as syn

ap = argparse.ArgumentParser()
ap.add_argument("--config", type=int, required=True, choices=[4, 5])
ap.add_argument("--scale", type=float, default=1.0, help="scale spp (and views for config 4) for quick runs")
ap.add_argument("--no-reuse", action="store_true", help="config 4: rebuild the BVH for every view (the reference's behaviour)")
ap.add_argument("--tile-collective", default="allreduce", choices=["gather", "allreduce"])
ap.add_argument("--out", default=None)
ap.add_argument("--no-instancing", action="store_true", help="config 5: flat drp_build over the 10 M flattened triangles instead of drp_build_instanced")
ap.add_argument("--profile", action="store_true", help="config 4: cProfile of the view loop on rank 0 (host overhead per session)")
args = ap.parse_args()
world, rank, local = int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    dist.init_process_group("nccl", device_id=dev)


def b_query(n_tris):   # ideal-descent bytes per traced ray (500 k: 980 B, 10 M: 1220 B)
    return 32 + 48 * math.ceil(math.log2(max(2, n_tris))) + 36


def peak_gbs():
    p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
    try:
        j = json.load(open(p))
        for k in ("hbm_gbs", "hbm_copy_gbs", "hbm_gbs_sustained", "hbm_bw_gbs"):
            if k in j:
                return float(j[k]), "MEASURED_PEAKS.json:" + k
        for k, v in j.items():
            if "hbm" in k.lower() and isinstance(v, (int, float)):
                return float(v), "MEASURED_PEAKS.json:" + k
    except Exception:
        pass
    return 7700.0, "fallback (B200_PROFILING.md nominal)"


def max_over_ranks(x):
    t = torch.tensor([x], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
S_DEFAULT = 224   # shading bytes per ray-bounce, DefaultMaterial (B_bounce: config 4 = 1204 B, config 5 = 1444 B)
if args.config == 4:
    scene_host, orbit = syn.datagen_scene('cpu')
    scene = scene_host.to(dev)
    n_views, spp, depth, res = max(world, int(64 * args.scale)), max(1, int(64 * args.scale)), 3, 512
    views = list(range(n_views))[rank::world]
    sess0 = drp.PathTracingSession(scene, drp.PerspectiveCamera.from_orbit(**orbit(0, n_views, res)), drp.PathTracingSessionOptions(ray_spp=spp, ray_depth=depth))
    sess0.raycaster(); sess0._fused_scene()   # scene upload + flatten + build once, outside the timed region (options.reuse_scene)
    n_tris = int(sess0.vertex_array_object().tris.shape[0])
    sess0.pbr()   # one untimed view: ray-queue workspace allocated at its full size, kernels loaded (like the warm-up steps of bench.py)
    if world > 1:
        dist.barrier()
    prof = None
    if args.profile and rank == 0:
        import cProfile
        prof = cProfile.Profile()
        prof.enable()
    torch.cuda.synchronize(); ev[0].record()
    outs, traced = [], 0
    for k in views:
        cam = drp.PerspectiveCamera.from_orbit(**orbit(k, n_views, res))
        sess = drp.PathTracingSession(scene, cam, drp.PathTracingSessionOptions(ray_spp=spp, ray_depth=depth, seed=k, reuse_scene=not args.no_reuse))
        rad, alpha, extras = sess.pbr()  # a session per view (single-use, like the reference); the scene cache keeps flatten + BVH
        outs.append(torch.cat([rad, alpha, extras['albedo'], extras['world_normal']], -1))
    ev[1].record(); torch.cuda.synchronize()
    if prof is not None:
        import pstats
        prof.disable()
        pstats.Stats(prof, stream=sys.stderr).sort_stats("cumulative").print_stats(30)
    ms = max_over_ranks(ev[0].elapsed_time(ev[1]))
    imgs = torch.stack(outs)
    assert torch.isfinite(imgs).all() and imgs[..., 4:7].max() <= 1.0 + 1e-5
    stats = sess.render_stats()
    live = stats['rays_traced'] / max(1, stats['rays_nominal'])
    nominal = n_views * res * res * spp * depth
    peak, peak_src = peak_gbs()
    bb = b_query(n_tris) + S_DEFAULT
    res_d = dict(config=4, n_gpus=world, sharding="views rank::world, no exchange", views=n_views, triangles=n_tris, resolution=[res, res], spp=spp, ray_depth=depth,
                 seconds=ms / 1e3, mrays_s=nominal / ms / 1e3, views_per_s=n_views / (ms / 1e3), live_ray_fraction_last_view=live,
                 aovs=["radiance", "alpha", "albedo", "world_normal"], mean_radiance=float(imgs[..., :3].mean()), coverage=float((imgs[..., 3] > 0).float().mean()),
                 roofline=dict(bound="hbm", B_bounce=bb, achieved=nominal * live * bb / (ms * 1e-3) / 1e9 / world, peak=peak, unit="GB/s per GPU",
                               frac=nominal * live * bb / (ms * 1e-3) / 1e9 / world / peak, peak_source=peak_src,
                               note="algorithmic bytes = live ray-bounces x B_bounce (SURVEY 8d), per GPU"))
else:
    scene_host, camkw = syn.instanced_scene('cpu')
    scene = scene_host.to(dev)
    H, W, spp, depth = 2160, 3840, max(1, int(256 * args.scale)), 3
    cam = drp.PerspectiveCamera.from_orbit(h=H, w=W, **camkw)
    opt = drp.PathTracingSessionOptions(ray_spp=spp, ray_depth=depth, seed=2, shard_rank=rank, shard_world=world, shard_mode='tile', tile_size=256,
                                        tile_collective=args.tile_collective, instancing=not args.no_instancing)
    sess = drp.PathTracingSession(scene, cam, opt)
    torch.cuda.synchronize(); t0 = time.perf_counter(); sess.vertex_array_object(); torch.cuda.synchronize(); flatten_s = time.perf_counter() - t0
    t0 = time.perf_counter(); sess.raycaster(); torch.cuda.synchronize(); structure_s = time.perf_counter() - t0
    t0 = time.perf_counter(); sess._fused_scene(); torch.cuda.synchronize(); build_s = flatten_s + structure_s + time.perf_counter() - t0
    # steady-state structure time (the first build pays the stream-ordered pool's growth)
    v_ = sess.vertex_array_object()
    from diffrp_b200.path_tracing import scene_instances
    cfg_ = {'epsilon': 1e-8}
    if not args.no_instancing:
        cfg_['instances'] = scene_instances(scene.objects)
    torch.cuda.synchronize(); t0 = time.perf_counter(); rc2 = drp.B200Raycaster(v_.world_pos, v_.tris, cfg_); torch.cuda.synchronize(); structure2_s = time.perf_counter() - t0
    rc2.release()
    n_tris = int(sess.vertex_array_object().tris.shape[0])
    warm = sess.new_accumulators()
    sess.render_samples(torch.arange(1, dtype=torch.int32, device=dev), warm, tile=sess.tiles()[rank])   # workspace at its full size, kernels loaded
    sess.exchange_accumulators(warm)
    del warm
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(); ev[0].record()
    acc = sess.render_accumulators()
    ev[1].record()
    if world > 1:
        dist.barrier()   # so that exchange_ms is the exchange, not the wait for the rank with the most expensive tiles (that is in render_ms: max over ranks)
    ev[4].record()
    acc = sess.exchange_accumulators(acc)
    ev[2].record()
    rad, alpha, extras = sess.finalize(acc)
    ev[3].record(); torch.cuda.synchronize()
    ms = max_over_ranks(ev[0].elapsed_time(ev[3]))
    ms_render, ms_exchange = max_over_ranks(ev[0].elapsed_time(ev[1])), max_over_ranks(ev[4].elapsed_time(ev[2]))
    assert torch.isfinite(rad).all()
    nominal = H * W * spp * depth
    peak, peak_src = peak_gbs()
    bb = b_query(n_tris) + S_DEFAULT
    res_d = dict(config=5, n_gpus=world, sharding="256-pixel tiles rank::world, exchange = %s" % args.tile_collective, triangles=n_tris, objects=len(scene.objects),
                 resolution=[W, H], spp=spp, ray_depth=depth, upload_flatten_build_s=build_s, flatten_s=flatten_s, structure_first_s=structure_s, structure_steady_s=structure2_s,
                 structure="drp_build_instanced" if sess.raycaster().instanced else "drp_build", seconds=ms / 1e3, render_ms=ms_render, exchange_ms=ms_exchange,
                 exchange_bytes_per_rank=(H * W * 64 * (world - 1) // world) * (1 if args.tile_collective == 'gather' else 2), mrays_s=nominal / ms / 1e3,
                 coverage=float((alpha > 0).float().mean()), mean_radiance=float(rad.mean()), bvh=sess.raycaster().stats(),
                 roofline=dict(bound="hbm", B_bounce=bb, peak=peak, unit="GB/s per GPU", peak_source=peak_src,
                               note="nominal ray-bounces x B_bounce / time / GPUs (live fraction not tracked across tile calls: upper bound of the achieved figure)",
                               achieved_nominal=nominal * bb / (ms * 1e-3) / 1e9 / world, frac_nominal=nominal * bb / (ms * 1e-3) / 1e9 / world / peak))
if rank == 0:
    txt = json.dumps(res_d)
    print(txt)
    if args.out:
        os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
        open(args.out, "w").write(txt + "\n")
if world > 1:
    dist.destroy_process_group()
