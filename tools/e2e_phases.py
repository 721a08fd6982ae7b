"""
Per-phase, per-rank timing of the end-to-end call of bench.py (PathTracingSession(host-pinned scene).pbr() + D2H) under torchrun:
where does the fixed cost of a call go when N ranks run it at once?  (17 -> 83 ms from 1 to 8 GPUs.)

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tools/e2e_phases.py --steps 16

Every phase is bracketed by a device synchronisation (so the phases do not overlap: the sum is an upper bound of the fused call, which is
also timed, un-instrumented, for comparison).  Rank 0 prints one JSON object: per phase min / max / mean over ranks, in ms.
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import diffrp_b200 as drp
from diffrp_b200 import synthetic as syn
from diffrp_b200.path_tracing import reduce_accumulators


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=16)
    ap.add_argument("--spp-per-step", type=int, default=8)
    ap.add_argument("--tex", type=int, default=1024)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--upload", default="auto", choices=["auto", "direct", "sharded"], help="options.scene_upload")
    ap.add_argument("--strided-d2h", action="store_true", help="read the outputs back into channel slices of one (H,W,16) pinned array (round 1's bench)")
    ap.add_argument("--all-ranks-read", action="store_true", help="all-reduce + finalize + D2H on every rank (round-2a behaviour) instead of reduce to rank 0")
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    world, rank, local = int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=dev)
    scene_host, camkw = syn.teaser_scene('cpu', tex=args.tex)
    scene_host = scene_host.pin_memory()   # one page-locked arena in upload order (Scene.pin_memory)
    cam = drp.PerspectiveCamera.from_orbit(h=1024, w=1024, **camkw)
    out_keys = ("radiance", "alpha", "albedo", "emission", "world_normal", "world_position")
    out_host = {k: torch.empty([1024, 1024, 1 if k == "alpha" else 3], dtype=torch.float32).pin_memory() for k in out_keys}
    out_strided = torch.empty([1024, 1024, 16], dtype=torch.float32).pin_memory()   # --strided-d2h: round 1's layout
    spp = world * args.steps * args.spp_per_step
    extra = {'scene_upload': args.upload}

    def options(seed):
        return drp.PathTracingSessionOptions(ray_spp=spp, ray_depth=4, rng='native', seed=seed, reuse_scene=False, shard_rank=rank, shard_world=world,
                                                result_rank=0 if (world > 1 and not args.all_ranks_read) else None, **extra)

    def d2h(r, a, x):
        if args.strided_d2h:
            out_strided[..., 0:3].copy_(r, non_blocking=True)
            out_strided[..., 3:4].copy_(a, non_blocking=True)
            for q, k in enumerate(("albedo", "emission", "world_normal", "world_position")):
                out_strided[..., 4 + 3 * q:7 + 3 * q].copy_(x[k], non_blocking=True)
            return
        out_host["radiance"].copy_(r, non_blocking=True)
        out_host["alpha"].copy_(a, non_blocking=True)
        for k in ("albedo", "emission", "world_normal", "world_position"):
            out_host[k].copy_(x[k], non_blocking=True)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    phases = ("flatten+H2D geometry", "LBVH build", "H2D textures + records", "render setup", "render (K sections)", "all-reduce", "finalize", "D2H", "release")
    rows, fused = [], []
    for rep in range(args.reps + 1):
        barrier()
        t = []

        def timed(fn):
            torch.cuda.synchronize(); t0 = time.perf_counter(); r = fn(); torch.cuda.synchronize()
            t.append((time.perf_counter() - t0) * 1e3)
            return r
        s = drp.PathTracingSession(scene_host, cam, options(rep))
        timed(s.vertex_array_object)
        timed(s.raycaster)
        timed(s._fused_scene)
        timed(s._render_setup)
        acc = timed(s.render_accumulators)
        acc = timed(lambda: s.exchange_accumulators(acc))
        mine_out = world == 1 or args.all_ranks_read or rank == 0
        out = timed(lambda: s.finalize(acc) if mine_out else None)
        timed(lambda: d2h(*out) if mine_out else None)
        timed(lambda: s.raycaster().release())
        barrier()
        t0 = time.perf_counter()
        s = drp.PathTracingSession(scene_host, cam, options(100 + rep))
        res = s.pbr()
        if res is not None:
            d2h(*res)
        torch.cuda.synchronize()
        t_fused = (time.perf_counter() - t0) * 1e3
        s.raycaster().release()
        if rep > 0:  # the first repetition warms allocator pools and the workspace
            rows.append(t)
            fused.append(t_fused)
    med = lambda xs: sorted(xs)[len(xs) // 2]   # median over the repetitions: one slow repetition (allocator growth, a host hiccup) must not set the figure
    mine = torch.tensor([[med([r[k] for r in rows]) for k in range(len(phases))] + [med(fused)]], device=dev)
    if world > 1:
        allr = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allr, mine)
        allr = torch.cat(allr).cpu()
    else:
        allr = mine.cpu()
    if rank == 0:
        res = {"n_gpus": world, "steps_per_rank": args.steps, "spp_total": spp, "scene_upload": args.upload, "strided_d2h": bool(args.strided_d2h),
               "phases_ms": {p: {"min": float(allr[:, k].min()), "max": float(allr[:, k].max()), "mean": float(allr[:, k].mean())} for k, p in enumerate(phases)},
               "sum_of_phases_ms_max_rank": float(allr[:, :-1].sum(1).max()),
               "fused_call_ms": {"min": float(allr[:, -1].min()), "max": float(allr[:, -1].max())},
               "per_rank_fused_ms": [float(x) for x in allr[:, -1]]}
        txt = json.dumps(res, indent=1)
        print(txt)
        if args.out:
            os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
            open(args.out, "w").write(txt)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
