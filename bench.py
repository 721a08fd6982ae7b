#!/usr/bin/env python
"""
bench.py -- headline benchmark of the diffrp path-tracing hot path on B200 ("Measurement").

Workload (BASELINE.json configs[2], the configuration the north-star metric is quoted on): synthetic 2,097,152-triangle
textured PBR scene (49 objects, 8 GLTF materials with 1024^2 fp32 textures, HDR env), 1024x1024, 4 bounces,
samples-per-pixel sharded across ranks.  A *step* is one section of `--spp-per-step` (default 8) samples of every pixel
through all bounces = 8,388,608 paths = 33,554,432 nominal ray-bounces -- the reference's own section size
(ray_split_size = 8M rays, path_tracing.py:74,318).  The default K = 128 steps is the full 1024-spp frame at N = 1.

    python bench.py --gpus N --steps K --warmup W            # this framework
    python bench.py --impl reference --gpus N --steps K ...  # the reference's algorithm (CPU oracle port) on host cores

Prints ONE JSON line (rank 0).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

RES, DEPTH = 1024, 4
METRIC, UNIT = "nominal_ray_bounces_per_second", "Mrays/s"


def b_query(n_tris):  # ideal-descent bytes per traced ray
    return 32 + 48 * int(np.ceil(np.log2(max(2, n_tris)))) + 36


S_GLTF = 504  # shading bytes per ray-bounce, textured GLTF


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p)).get("hbm_gbs", 6650.0), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.rows, self.proc, self.index, self.first = [], None, index, 0

    def wait_ready(self, timeout=5.0):
        """Block until the sampler delivers rows (nvidia-smi start-up -- fork, NVML init -- can stall driver calls of this process
        for ~100 ms, so it must be over before the timed region starts), then count only rows from here on."""
        t0 = time.time()
        while self.proc and not self.rows and time.time() - t0 < timeout:
            time.sleep(0.01)
        self.first = len(self.rows)

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, pw = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows[self.first:]:
            try:
                sm.append(float(r[0])); mx.append(float(r[1])); pw.append(float(r[2]))
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


def dist_setup(n_gpus):
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return world, rank, local


def scene_bytes(scene):
    n = 0
    seen = set()

    def add(t):
        nonlocal n
        if isinstance(t, torch.Tensor) and t.data_ptr() not in seen:
            seen.add(t.data_ptr())
            n += t.numel() * t.element_size()
    for o in scene.objects:
        for t in (o.verts, o.tris, o.normals, o.M, o.color, o.uv, o.tangents):
            add(t)
        m = o.material
        for k in ("base_color_texture", "metallic_roughness_texture", "normal_texture", "emissive_texture"):
            s = getattr(m, k, None)
            if s is not None:
                add(s.image)
    for l in scene.lights:
        add(l.image)
    return n


def cpu_oracle_setup(tex):
    """Scene flattened on the host + oracle BVH (the reference's algorithm restated in C, oracle/)."""
    import oracle
    from diffrp_b200 import synthetic as syn, ops
    import diffrp_b200 as drp
    host_threads(oracle)
    scene, camkw = syn.teaser_scene('cpu', tex=tex)
    ops.set_default_device('cpu')
    try:
        cam = drp.PerspectiveCamera.from_orbit(h=RES, w=RES, **camkw)
        cam._keep = (cam.V(), cam.P())
        vao, hs, p, keep = oracle.inputs_from_scene(scene, cam, 1024, DEPTH, seed=1)
    finally:
        ops.set_default_device(None)
    t0 = time.perf_counter()
    bvh = oracle.BVH(vao.world_pos.numpy(), vao.tris.numpy(), 'splitaxis')
    build_s = time.perf_counter() - t0
    return oracle, bvh, hs, p, keep, build_s, int(vao.tris.shape[0])


def cpu_oracle_step(oracle, bvh, hs, p, keep, win, sample_ids):
    """One bounded sample of the workload: central win x win window of the 1024^2 frame, given samples, 4 bounces."""
    lo = RES // 2 - win // 2
    p2, k2 = oracle.make_params(win, win, DEPTH, p.t_far, p.t_near, list(p.cam_pos)[:3], np.array(list(p.inv_vp), np.float32),
                                keep['ndc_x'][lo:lo + win].copy(), keep['ndc_y'][lo:lo + win].copy(), keep['jitter_x_all'][sample_ids],
                                keep['jitter_y_all'][sample_ids], sample_ids=sample_ids, seed=1)
    t0 = time.perf_counter()
    acc, n = oracle.render(bvh, hs, p2)
    return time.perf_counter() - t0, n


def gpu_torch_baseline(scene, camkw, dev, args):
    """The reference's torch-level GPU path, timed on this GPU in this run (north_star: "next to the reference timed in the same
    run: its OptiX/torch GPU path"): the reference is a Python package that cannot travel to the box and torchoptix is not in the
    image, so this is a PORT -- oracle/torch_pbbvh.py (NaivePBBVH restated as whole-array torch ops, checked against the reference's
    own outputs in tests/test_oracle_golden.py) for intersection + the per-bounce torch shading ops of diffrp_b200/generic.py (the
    restatement of path_tracing.py:158-352 that the parity tests compare with the fused kernels).  BVH build excluded, like `value`."""
    import diffrp_b200 as drp
    from oracle.torch_pbbvh import TorchPBBVHRaycaster
    res, spp = args.torch_baseline_res, args.torch_baseline_spp
    cam = drp.PerspectiveCamera.from_orbit(h=res, w=res, **camkw)
    holder = {}

    class TorchPathSession(drp.PathTracingSession):
        def _fused_scene(self):
            return None                                  # force the per-bounce torch path

        def raycaster(self):
            if 'rc' not in holder:
                vao = self.vertex_array_object()
                holder['rc'] = TorchPBBVHRaycaster(vao.world_pos, vao.tris, {'builder': 'splitaxis'})
            return holder['rc']

    def one(seed):
        o = drp.PathTracingSessionOptions(ray_spp=spp, ray_depth=DEPTH, rng='native', seed=seed)
        r, a, x = TorchPathSession(scene, cam, o).pbr()
        return r

    t0 = time.perf_counter()
    TorchPathSession(scene, cam, drp.PathTracingSessionOptions(ray_spp=1, ray_depth=1)).raycaster()
    torch.cuda.synchronize()
    build_s = time.perf_counter() - t0
    one(0)                                               # warm-up (jit, allocator)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n, ms = 0, 0.0
    while n < args.torch_baseline_steps or ms < 3000.0:
        e0.record(); r = one(1 + n); e1.record(); torch.cuda.synchronize()
        ms += e0.elapsed_time(e1); n += 1
        if n >= 16:
            break
    assert torch.isfinite(r).all()
    rays = n * res * res * spp * DEPTH
    return {"value": rays / (ms * 1e-3) / 1e6, "unit": UNIT, "kind": "port", "device": "same B200, same run",
            "what": "torch-level NaivePBBVH (splitaxis) + per-bounce torch shading, the reference's fallback GPU path restated (oracle/torch_pbbvh.py + diffrp_b200/generic.py)",
            "sample": "%d frames of %dx%d x %d spp x %d bounces, same scene and camera (%d nominal ray-bounces, %.1f s); torch BVH build %.1f s excluded"
                      % (n, res, res, spp, DEPTH, rays, ms * 1e-3, build_s)}


def host_threads(oracle):
    """All the host threads this process may use: torchrun exports OMP_NUM_THREADS=1 for its workers, which would time the CPU arm on one core."""
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    oracle.set_num_threads(max(1, n))


def run_reference(args, world, rank):
    """--impl reference: the reference's own implementation of the path on the box's host cores -- unmodified diffrp from baseline/_ref
    (kind "reference"); the CPU oracle port only when no reference install travelled with the repo (kind "port")."""
    if rank != 0:
        return
    from baseline import ref_bench
    if ref_bench.reference_available() and not args.ref_port:
        r = ref_bench.run('cpu', args.tex, args.ref_window, args.ref_spp, args.steps, min(args.warmup, args.ref_max_warmup), budget_s=args.ref_budget_s)
        val, cores, kind, sample, ms_step, n_tris = r["value"], r["cores"], r["kind"], r["sample"], r["ms_per_step"], r["n_tris"]
    else:
        oracle, bvh, hs, p, keep, build_s, n_tris = cpu_oracle_setup(args.tex)
        host_threads(oracle)
        cores, kind = oracle.num_threads(), "port"
        win, spp = args.ref_window, args.ref_spp
        for w in range(min(args.warmup, 1)):
            cpu_oracle_step(oracle, bvh, hs, p, keep, win, np.arange(spp))
        tot_t, tot_n = 0.0, 0
        for k in range(args.steps):
            dt, n = cpu_oracle_step(oracle, bvh, hs, p, keep, win, np.arange(k * spp, (k + 1) * spp) % 1024)
            tot_t += dt
            tot_n += n
        val, ms_step = tot_n / tot_t / 1e6, tot_t / max(1, args.steps) * 1e3
        sample = "CPU oracle port (no baseline/_ref): %dx%d central window of the %dx%d frame, %d spp, %d bounces per step (%d ray-bounces), 2,097,152-tri scene, CPU BVH build %.1f s excluded" % (
            win, win, RES, RES, spp, DEPTH, win * win * spp * DEPTH, build_s)
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": workload_config(n_tris, args),
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


def workload_config(n_tris, args):
    return {"workload": "configs[2]: synthetic %d-triangle textured PBR scene (49 objects, 8 GLTF materials, %d^2 fp32 textures, HDR env), "
                        "%dx%d, %d bounces, %d spp per step, spp-sharded" % (n_tris, args.tex, RES, RES, DEPTH, args.spp_per_step),
            "triangles": n_tris, "resolution": [RES, RES], "ray_depth": DEPTH, "spp_per_step": args.spp_per_step,
            "rays_per_step": RES * RES * args.spp_per_step * DEPTH, "l2_policy": "working set (BVH 0.2 GB + attributes + 0.45 GB textures + 0.9 GB ray queues) exceeds the 126 MB L2",
            "parallelism": "spp-shard x%d, scene replicated, NCCL all-reduce of the fp32 accumulators" % args.gpus}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None, help="default: 128 (20 for --impl reference, whose steps are seconds of CPU work each)")
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--spp-per-step", type=int, default=8)
    ap.add_argument("--tex", type=int, default=1024)
    ap.add_argument("--e2e-steps", type=int, default=6)
    ap.add_argument("--ref-window", type=int, default=256, help="--impl reference: window edge of one step (the reference's CPU throughput grows with the batch: "
                    "0.035 / 0.096 / 0.17 Mrays/s at 128^2 / 256^2 / 512^2 x 4 spp on 8 cores; 256 keeps a 20-step run within minutes)")
    ap.add_argument("--cpu-baseline-window", type=int, default=512)
    ap.add_argument("--ref-spp", type=int, default=4)
    ap.add_argument("--ref-port", action="store_true", help="reference legs: time the CPU oracle port instead of the installed reference")
    ap.add_argument("--ref-budget-s", type=float, default=150.0, help="--impl reference: bound on the timed region; the per-step window shrinks when K steps would exceed it")
    ap.add_argument("--ref-max-warmup", type=int, default=2, help="--impl reference: cap on the untimed warm-up steps (each is seconds of CPU work)")
    ap.add_argument("--cpu-baseline-steps", type=int, default=2)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--torch-baseline-res", type=int, default=1024, help="window edge of the reference's GPU leg: 1024 x 8 spp = one real section of the "
                    "reference (ray_split_size = 8M rays, path_tracing.py:74,318) = exactly one bench step")
    ap.add_argument("--torch-baseline-spp", type=int, default=8)
    ap.add_argument("--torch-baseline-steps", type=int, default=1)
    ap.add_argument("--no-torch-baseline", action="store_true")
    ap.add_argument("--texel-tile", type=int, default=None, help="A/B: log2 of the texel-record tile edge (0 = row-major; default: the library's)")
    ap.add_argument("--separate-textures", action="store_true", help="A/B: four separate RGBA textures per material instead of the interleaved texel records")
    ap.add_argument("--no-strong", action="store_true", help="skip the strong-scaling record (one 1024-spp frame split over the ranks)")
    ap.add_argument("--strong-spp", type=int, default=1024)
    args = ap.parse_args()
    if args.steps is None:
        args.steps = 20 if args.impl == "reference" else 128
    world, rank, local = dist_setup(args.gpus)
    if args.impl == "reference":
        return run_reference(args, world, rank)

    import torch.distributed as dist
    import diffrp_b200 as drp
    from diffrp_b200 import synthetic as syn
    from diffrp_b200._lib import loaded_path, build_config
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if args.separate_textures:
        import diffrp_b200.flatten as _flatten_mod
        _flatten_mod.INTERLEAVE_TEXELS = False
    if args.texel_tile is not None:
        import diffrp_b200.flatten as _flatten_mod
        _flatten_mod.TEXEL_TILE_LOG2 = args.texel_tile
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")   # NCCL prints its version banner on stdout; stdout carries exactly one JSON line
        dist.init_process_group("nccl", device_id=dev)
    S, K, W = args.spp_per_step, args.steps, max(args.warmup, 3)
    HW = RES * RES

    # ---- device-resident arm -------------------------------------------------------------------------------------
    scene_host, camkw = syn.teaser_scene('cpu', tex=args.tex)
    scene_host = scene_host.pin_memory()   # one page-locked arena in upload order (Scene.pin_memory)
    scene = scene_host.to(dev)
    cam = drp.PerspectiveCamera.from_orbit(h=RES, w=RES, **camkw)
    total_spp = max(1, world * K * S)
    opts = drp.PathTracingSessionOptions(ray_spp=total_spp, ray_depth=DEPTH, rng='native', seed=1, shard_rank=rank, shard_world=world)
    sess = drp.PathTracingSession(scene, cam, opts)
    ev = lambda: torch.cuda.Event(enable_timing=True)
    e0, e1 = ev(), ev()
    torch.cuda.synchronize()
    vao = sess.vertex_array_object()
    n_tris = int(vao.tris.shape[0])
    # steady-state build time: sessions are single-use, so frame N's structure is released before frame N+1's is built and the
    # stream-ordered pool hands the same blocks back; the first build pays the one-off pool growth and is not the one timed
    warm_rc = drp.B200Raycaster(vao.world_pos, vao.tris)
    warm_rc.release()
    torch.cuda.synchronize()
    e0.record(); rc = sess.raycaster(); e1.record(); torch.cuda.synchronize()
    build_ms = e0.elapsed_time(e1)
    my_ids = torch.arange(total_spp, dtype=torch.int32, device=dev)[rank::world]  # global Hammersley indices of this rank
    step_ids = [my_ids[j * S:(j + 1) * S] for j in range(K)]
    scratch = sess.new_accumulators()
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()  # started before the warm-up: its start-up cost stays out of the timed region
    for j in range(W):
        sess.render_samples(step_ids[j % K], scratch)
    if world > 1:
        dist.all_reduce(scratch)
    sess.finalize(scratch)  # the warm-up runs everything the timed region runs (first use of a kernel / first allocation of the outputs)
    torch.cuda.synchronize()
    if rank == 0:
        clocks.wait_ready()
    for j in range(2):  # the GPU idled while the sampler came up: two more untimed steps so that the timed region starts hot
        sess.render_samples(step_ids[j % K], scratch)
    torch.cuda.synchronize()
    sess.get_profile()
    sess.set_profiling(True)
    launches0 = 0
    accum = sess.new_accumulators()
    import gc
    gc.collect()
    gc.disable()  # no collector pause between the launches of the timed region (re-enabled right after it)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0.record()
    t_host0 = time.perf_counter()
    for j in range(K):
        sess.render_samples(step_ids[j], accum)
    if world > 1:
        dist.all_reduce(accum)  # the path's one exchange step: fp32 accumulators over NVLink
    radiance, alpha, extras = sess.finalize(accum)
    host_issue_ms = (time.perf_counter() - t_host0) * 1e3  # host time to issue the region's launches (diagnostic; with per-kernel event spans the launch queue fills up, so this is mostly back-pressure from the GPU, not host work: tools/host_overhead.py measures 0.09 ms of host time per step without spans)
    e1.record()
    torch.cuda.synchronize()
    gc.enable()
    if world > 1:
        dist.barrier()
    ms = e0.elapsed_time(e1)
    clk = clocks.stop() if rank == 0 else None
    t_ms = torch.tensor([ms], device=dev)
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    ms_max = float(t_ms.item())
    prof = sess.get_profile()
    sess.set_profiling(False)
    stats = sess.render_stats()
    launches_per_step = stats["kernel_launches"]
    rays_nominal = world * K * S * HW * DEPTH
    value = rays_nominal / (ms_max * 1e-3) / 1e6
    assert torch.isfinite(radiance).all()

    # ---- end-to-end arm: the call a user makes for this workload ------------------------------------------------------
    # ONE public-API call renders the job: PathTracingSession(host-pinned scene, camera, options(ray_spp = K*S per rank ...)).pbr()
    # followed by the D2H read of every output.  Inside the timed region: H2D upload of the whole scene (geometry, textures,
    # env) from pinned host memory, flatten, LBVH build, all K sections, the accumulator all-reduce, finalize, D2H.
    # Also timed (reported as `single_section_session`): the same call for ONE section (ray_spp = S), i.e. the scene upload
    # and build are paid again for every 8 spp -- the worst case for a single-use session API.
    # one CONTIGUOUS pinned buffer per output: each read is a single DMA.  (Round 1 copied into channel slices of one (H, W, 16) pinned array;
    # torch stages such strided device->host copies through a CPU-side strided memcpy on all host threads -- 7 ms at N = 1, 49 ms per rank at
    # N = 2 when the ranks' thread pools oversubscribe the cores: tools/e2e_phases.py, profiles/r2/e2e_phases_*.json.)
    out_keys = ("radiance", "alpha", "albedo", "emission", "world_normal", "world_position")
    out_host = {k: torch.empty([RES, RES, 1 if k == "alpha" else 3], dtype=torch.float32).pin_memory() for k in out_keys}
    h2d = scene_bytes(scene_host)
    d2h = sum(v.numel() for v in out_host.values()) * 4

    def e2e_call(spp_total, seed, sharded):
        o = drp.PathTracingSessionOptions(ray_spp=spp_total, ray_depth=DEPTH, rng='native', seed=seed, reuse_scene=False,  # nothing cached between calls
                                          shard_rank=rank if sharded else 0, shard_world=world if sharded else 1,
                                          result_rank=0 if (sharded and world > 1) else None)   # one job, one frame: rank 0 receives the sum and reads it back
        s = drp.PathTracingSession(scene_host, cam, o)          # pinned host tensors: H2D of the whole scene happens inside
        res = s.pbr()                                           # flatten + LBVH build + wavefront (+ reduce to rank 0) + finalize
        if res is not None:
            r, a, x = res
            out_host["radiance"].copy_(r, non_blocking=True)    # D2H of every output
            out_host["alpha"].copy_(a, non_blocking=True)
            for k in ("albedo", "emission", "world_normal", "world_position"):
                out_host[k].copy_(x[k], non_blocking=True)
        torch.cuda.synchronize()
        s.raycaster().release()

    def timed_calls(n_calls, spp_total, sharded):
        e2e_call(min(spp_total, 2 * S * (world if sharded else 1)), 99, sharded)  # warm-up: allocator pools, ray-queue workspace at its full (2^24-ray) size
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0.record()
        for j in range(n_calls):
            e2e_call(spp_total, 100 + j, sharded)
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()) / n_calls

    E = max(1, args.e2e_steps)
    e2e_job_ms = timed_calls(1, world * K * S, True)              # the whole job: world*K*S spp, spp-sharded like the device arm
    e2e_value = world * K * S * HW * DEPTH / (e2e_job_ms * 1e-3) / 1e6
    e2e_sec_ms = timed_calls(E, S, False)                         # one section per session, every rank its own frame
    e2e_sec_value = world * S * HW * DEPTH / (e2e_sec_ms * 1e-3) / 1e6

    # ---- strong scaling of the north-star job: ONE 1024-spp frame (fixed total work) split over the N ranks ----------------------
    # device-resident: this rank's 1024 / N samples in sections of S, all-reduce, finalize; e2e: the same public-API call as above with
    # ray_spp = 1024.  Reported at every N (N = 1 included) so that the ratio between two lines of a scaling run is the strong-scaling speed-up.
    strong = None
    if not args.no_strong:
        FRAME_SPP = args.strong_spp
        s_opts = drp.PathTracingSessionOptions(ray_spp=FRAME_SPP, ray_depth=DEPTH, rng='native', seed=1, shard_rank=rank, shard_world=world)
        s_sess = drp.PathTracingSession(scene, cam, s_opts)
        s_ids = torch.arange(FRAME_SPP, dtype=torch.int32, device=dev)[rank::world]
        s_steps = [s_ids[j:j + S] for j in range(0, len(s_ids), S)]
        s_acc = s_sess.new_accumulators()
        s_sess.render_samples(s_steps[0], s_acc)   # warm-up of this session's tables
        s_acc.zero_()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0.record()
        for ids_j in s_steps:
            s_sess.render_samples(ids_j, s_acc)
        if world > 1:
            dist.all_reduce(s_acc)
        s_out = s_sess.finalize(s_acc)
        e1.record()
        torch.cuda.synchronize()
        t_s = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(t_s, op=dist.ReduceOp.MAX)
        s_dev_ms = float(t_s.item())
        e2e_call(FRAME_SPP, 98, True)                      # warm-up at the job's own size (allocator pools, collectives at this message size)
        s_e2e_runs = [timed_calls(1, FRAME_SPP, True) for _ in range(2)]
        s_e2e_ms = min(s_e2e_runs)
        rays = FRAME_SPP * HW * DEPTH
        strong = {"job": "one %dx%d frame, %d spp, %d bounces (fixed total work), spp-sharded x%d" % (RES, RES, FRAME_SPP, DEPTH, world),
                  "device_ms": s_dev_ms, "e2e_ms": s_e2e_ms, "e2e_ms_runs": s_e2e_runs, "value": rays / (s_dev_ms * 1e-3) / 1e6, "e2e_value": rays / (s_e2e_ms * 1e-3) / 1e6,
                  "unit": UNIT, "scaling": "strong"}
        del s_out, s_acc

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (k_extend, all bounces), from live CUDA events on the launch stream ---------
    peak, peak_src = measured_peaks()
    ext_s = prof["extend_ms"] * 1e-3
    algo_bytes = prof["extend_rays"] * b_query(n_tris)
    achieved = algo_bytes / ext_s / 1e9 if ext_s > 0 else 0.0
    roofline = {"bound": "hbm", "kernel": "k_extend (closest hit, all bounces)", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": None, "peak_source": peak_src,
                "algorithmic_bytes_per_ray": b_query(n_tris), "rays_per_launch_avg": prof["extend_rays"] / max(1, prof["extend_launches"]),
                "launches": prof["extend_launches"], "avg_launch_ms": prof["extend_ms"] / max(1, prof["extend_launches"]),
                "share_of_step": prof["extend_ms"] / ms, "shade_share_of_step": prof["shade_ms"] / ms,
                "whole_path_frac": (stats_bytes(prof, n_tris) / (ms * 1e-3) / 1e9) / peak}
    traffic_file = os.path.join(ROOT, "profiles", "traffic_k_extend.json")
    if os.path.exists(traffic_file):  # one `ncu --set full` capture of the same command, committed under profiles/
        roofline["traffic"] = json.load(open(traffic_file)).get("dram_bytes_per_launch")
        roofline["traffic_unit"] = "bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum, ncu)"
        tj = json.load(open(traffic_file))
        if tj.get("issue"):
            # the limiter of this kernel is neither memory level but warp-instruction issue x SIMT efficiency (ncu, same capture): report that roofline too
            roofline["issue"] = {k: tj["issue"][k] for k in ("ipc_per_smsp", "active_lanes", "frac", "ipc_by_bounce", "active_lanes_by_bounce", "note")}
            roofline["issue"]["source"] = tj.get("source")
        if tj.get("l2_bytes_per_launch") and tj.get("l2_read_peak_gbs") and roofline["avg_launch_ms"] > 0:
            # L2-level view (SURVEY 8d asks for % of the L2 roofline too): ncu L2 sector bytes per launch / live launch time / measured L2 peak
            l2_gbs = tj["l2_bytes_per_launch"] / (roofline["avg_launch_ms"] * 1e-3) / 1e9
            roofline["l2"] = {"achieved": l2_gbs, "peak": tj["l2_read_peak_gbs"], "unit": "GB/s", "frac": l2_gbs / tj["l2_read_peak_gbs"],
                              "bytes_per_launch": tj["l2_bytes_per_launch"], "peak_source": tj.get("l2_peak_source")}
    roofline["algorithmic_bytes_per_launch"] = roofline["rays_per_launch_avg"] * b_query(n_tris)
    roofline["note"] = ("algorithmic bytes = live rays x B_query (SURVEY 8d ideal-descent model, every level re-read from memory); measured DRAM "
                        "traffic is ~10x lower because the upper BVH levels stay in L1/L2 -- the kernel is issue-bound (ncu: math-pipe throttle, "
                        "~0.7 inst/cycle/SMSP, 19-21 of 32 lanes active), see profiles/README.md")

    # ---- the reference timed in the same run (rank 0, N = 1): its torch GPU path on this B200, then its CPU path on the host cores --------
    # Both legs run UNMODIFIED diffrp from baseline/_ref (baseline/ref_bench.py); the restatements (oracle port / torch port) only stand in
    # when no reference install travelled with the repo.
    from baseline import ref_bench, ref_loader
    have_ref = ref_bench.reference_available() and not args.ref_port
    torch_baseline = None
    if world == 1 and not args.no_torch_baseline:
        try:
            if have_ref:
                try:
                    r = ref_bench.run('cuda', args.tex, args.torch_baseline_res, args.torch_baseline_spp, args.torch_baseline_steps, 1)
                except torch.cuda.OutOfMemoryError:  # the reference holds >= 12 full-R temporaries per bounce: retry with a quarter of the rays
                    ref_loader.unload_reference()
                    torch.cuda.empty_cache()
                    r = ref_bench.run('cuda', args.tex, args.torch_baseline_res // 2, args.torch_baseline_spp, args.torch_baseline_steps, 1)
                torch_baseline = {"value": r["value"], "unit": UNIT, "kind": "reference", "device": "same B200, same run", "ms_per_step": r["ms_per_step"],
                                  "what": "the reference's own torch GPU path (NaivePBBVH, splitaxis) + its torch shading", "sample": r["sample"]}
                ref_loader.unload_reference()
                torch.cuda.empty_cache()
            else:
                torch_baseline = gpu_torch_baseline(scene, camkw, dev, args)
        except Exception as exc:  # a baseline must never take the product's line down
            torch_baseline = {"unavailable": "%s: %s" % (type(exc).__name__, str(exc)[:200])}
    cpu_baseline = None
    if world == 1 and not args.no_cpu_baseline:
        if have_ref:
            try:
                r = ref_bench.run('cpu', args.tex, args.cpu_baseline_window, args.ref_spp, args.cpu_baseline_steps, 0)
                cpu_baseline = {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": "reference",
                                "sample": "%d steps (%.1f s): %s" % (args.cpu_baseline_steps, r["seconds"], r["sample"])}
            except Exception as exc:
                cpu_baseline = {"unavailable": "%s: %s" % (type(exc).__name__, str(exc)[:200])}
        if cpu_baseline is None or "unavailable" in cpu_baseline:
            oracle, bvh, hs, p, keep, build_s, _ = cpu_oracle_setup(args.tex)
            host_threads(oracle)
            tt, nn = 0.0, 0
            cpu_oracle_step(oracle, bvh, hs, p, keep, 128, np.arange(8))
            for k in range(8):
                dt, n = cpu_oracle_step(oracle, bvh, hs, p, keep, 128, np.arange(k * 8, (k + 1) * 8))
                tt += dt; nn += n
            cpu_baseline = {"value": nn / tt / 1e6, "unit": UNIT, "cores": oracle.num_threads(), "kind": "port",
                            "sample": "%d steps of a %dx%d central window x %d spp x %d bounces of the same scene (%d ray-bounces, %.1f s); CPU BVH build %.1f s excluded"
                                      % (8, 128, 128, 8, DEPTH, nn, tt, build_s)}

    # ---- the steps after the path, timed on this frame: informational, not part of `value` ----------------
    post = None
    if world == 1:
        from diffrp_b200 import denoiser as dn, tonemap as tm

        def timed(fn, iters=20):
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            e0.record()
            for _ in range(iters):
                fn()
            e1.record()
            torch.cuda.synchronize()
            return e0.elapsed_time(e1) / iters
        lut = torch.rand(32, 32, 32, 3, device=dev)
        net = dn.get_denoiser(seed=0)
        alb = tm.linear_to_srgb(extras['albedo'])
        den_ms = timed(lambda: dn.run_denoiser(net, radiance, alb, extras['world_normal']))
        tone_ms = timed(lambda: tm.tonemap(accum.view(RES, RES, 16), 'agx', lut=lut, scale=1.0 / total_spp, alpha_offset=3, flip_rows=True))
        macs = sum(9 * cin * cout * (RES // r) ** 2 for (_, cin, cout), r in zip(dn.LAYERS, (1, 1, 2, 4, 8, 16, 16, 8, 8, 4, 4, 2, 2, 1, 1, 1)))
        tf32_peak = None
        try:
            tf32_peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["bf16_tflops"] / 2.0
        except Exception:
            pass
        tfl = 2 * macs / (den_ms * 1e-3) / 1e12
        post = {"denoiser": {"ms": den_ms, "what": "run_denoiser: pack + 16 tcgen05 TF32 conv layers + unpack at %dx%d, seeded stand-in weights" % (RES, RES),
                             "roofline": {"bound": "tensor", "achieved": tfl, "unit": "TFLOP/s", "peak": tf32_peak,
                                          "frac": (tfl / tf32_peak) if tf32_peak else None,
                                          "peak_source": "MEASURED_PEAKS.json bf16_tflops / 2 (TF32 issues at half the bf16 rate)"}},
                "tonemap": {"ms": tone_ms, "what": "drp_tonemap: accumulator -> /spp, flipud, AgX LUT, sRGB, alpha, RGBA8",
                            "gbs": RES * RES * (64 + 4) / (tone_ms * 1e-3) / 1e9}}

    print(json.dumps({
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms_max / K,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(n_tris, args),
        "spp_per_second": world * K * S / (ms_max * 1e-3),
        "live_ray_fraction": prof["extend_rays"] / max(1, K * S * HW * DEPTH),
        "bvh_build_ms": build_ms,
        "clocks": clk, "host_issue_ms": host_issue_ms,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d / K, "d2h_bytes_per_step": d2h / K, "steps": K,
                "bytes_note": "job totals over all ranks: the scene crosses PCIe once in total (each rank uploads 1/N of every tensor, one coalesced NVLink "
                              "all-gather completes it: options.scene_upload); the accumulators are reduced to rank 0, which reads the frame back once "
                              "(options.result_rank)",
                "ms_per_step": e2e_job_ms / K, "ms_total": e2e_job_ms, "h2d_bytes_total": h2d, "d2h_bytes_total": d2h,
                "what": "ONE public-API call for the whole job: PathTracingSession(host-pinned scene, camera, options(ray_spp=%d, spp-sharded x%d)).pbr() "
                        "+ D2H of all outputs on rank 0; timed region = H2D scene upload, flatten, LBVH build, %d sections, reduce, finalize, D2H"
                        % (world * K * S, world, K),
                "single_section_session": {"value": e2e_sec_value, "unit": UNIT, "ms_per_call": e2e_sec_ms, "calls": E,
                                           "h2d_bytes_per_call": h2d, "d2h_bytes_per_call": d2h,
                                           "what": "same call with ray_spp=%d: scene upload + build paid per section" % S}},
        "strong_scaling": strong,
        "gpu_launches": int(launches_per_step * K + 1),  # kernels of libdiffrp_b200.so in the timed region: per step 4 x (k_extend_cw + k_shade)
                                                          # + k_count_traced + 8 x k_record_live (profiling spans); + 1 k_finalize
        "roofline": roofline,
        "cpu_baseline": cpu_baseline,
        "gpu_torch_baseline": torch_baseline,
        "post_path": post,
        "native_library": loaded_path(),
        "native_build": build_config(),
    }))
    if world > 1:
        dist.destroy_process_group()


def stats_bytes(prof, n_tris):
    """Algorithmic bytes of the whole path: traced rays x B_query + shaded rays x S_gltf."""
    return prof["extend_rays"] * b_query(n_tris) + prof["shade_rays"] * S_GLTF


if __name__ == "__main__":
    main()
