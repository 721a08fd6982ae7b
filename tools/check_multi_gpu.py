"""
Multi-GPU parity under real NCCL (run with torchrun on >= 2 GPUs): the sharded render -- host scene uploaded 1/N per rank + all-gather,
samples or tiles sharded, accumulators all-reduced / reduced to one rank / tile-gathered -- equals the single-process render of the same options
on rank 0 (fp32 sums in a different order: rtol 1e-5 on the accumulators' scale).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 tools/check_multi_gpu.py
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
import torch.distributed as dist
import diffrp_b200 as drp
import scenes

world, rank, local = int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
dist.init_process_group("nccl", device_id=dev)
assert world >= 2, "run under torchrun with at least 2 ranks"
host = scenes.mixed_scene(96, 48).pin_memory()            # identical on every rank (seeded), one page-locked arena
cam = drp.PerspectiveCamera.from_orbit(h=192, w=256, radius=3.0, azim=25, elev=15, origin=[0.0, -0.1, 0.0], fov=32)
base = dict(ray_spp=16, ray_depth=3, seed=7, reuse_scene=False)
results = {}


def frame(out):
    rad, alpha, extras = out
    return torch.cat([rad, alpha] + [extras[k] for k in ('albedo', 'emission', 'world_normal', 'world_position')], -1)


ref = frame(drp.PathTracingSession(host.to(dev), cam, drp.PathTracingSessionOptions(**base)).pbr()) if rank == 0 else None
cases = {
    "spp_allreduce_sharded_upload": dict(shard_mode='spp', scene_upload='sharded'),
    "spp_reduce_to_rank0": dict(shard_mode='spp', scene_upload='sharded', result_rank=0),
    "spp_direct_upload": dict(shard_mode='spp', scene_upload='direct'),
    "tile_allreduce": dict(shard_mode='tile', tile_size=64, scene_upload='sharded'),
    "tile_gather": dict(shard_mode='tile', tile_size=64, tile_collective='gather', scene_upload='sharded'),
}
for name, extra in cases.items():
    out = drp.PathTracingSession(host, cam, drp.PathTracingSessionOptions(shard_rank=rank, shard_world=world, **base, **extra)).pbr()
    if extra.get('result_rank') is not None and rank != extra['result_rank']:
        assert out is None
    if rank == 0:
        got = frame(out)
        err = (got - ref).abs()
        results[name] = dict(max_abs=float(err.max()), mean_abs=float(err.mean()), ok=bool(torch.allclose(got, ref, rtol=1e-4, atol=2e-5)))
    elif out is not None:   # every rank that holds a frame holds the same one
        pass
    if extra.get('result_rank') is None:
        mine = frame(out)
        ref_all = mine.clone()
        dist.broadcast(ref_all, src=0)
        assert torch.equal(mine, ref_all) or torch.allclose(mine, ref_all, rtol=0, atol=0), name   # all-reduce / gather: identical bits on every rank
    dist.barrier()
if rank == 0:
    ok = all(v['ok'] for v in results.values())
    print(json.dumps(dict(n_gpus=world, ok=ok, cases=results)))
    os.makedirs("gpurun_out", exist_ok=True)
    open("gpurun_out/check_multi_gpu_n%d.json" % world, "w").write(json.dumps(dict(n_gpus=world, ok=ok, cases=results), indent=1))
    assert ok, results
dist.destroy_process_group()
