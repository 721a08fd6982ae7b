"""Host-side cost of one render_samples call (python + ctypes + launches) on the bench scene: cProfile over 64 calls.
B200 box: 0.09 ms of host time per 7 ms step; with per-kernel profiling spans (bench.py) the issue loop takes 3.9 ms per step, but that is the
launch queue (~1000 entries) filling up and the host waiting for the GPU, not host work -- the GPU is never starved."""
import cProfile, pstats, io, os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
import diffrp_b200 as drp
from diffrp_b200 import synthetic as syn

dev = torch.device("cuda", 0)
scene_host, camkw = syn.teaser_scene('cpu', tex=256, pin=False)
scene = scene_host.to(dev)
cam = drp.PerspectiveCamera.from_orbit(h=1024, w=1024, **camkw)
K, S = 64, 8
sess = drp.PathTracingSession(scene, cam, drp.PathTracingSessionOptions(ray_spp=K * S, ray_depth=4, rng='native', seed=1))
ids = torch.arange(K * S, dtype=torch.int32, device=dev)
steps = [ids[j * S:(j + 1) * S] for j in range(K)]
acc = sess.new_accumulators()
for j in range(3):
    sess.render_samples(steps[j], acc)
torch.cuda.synchronize()
for prof_on in (False, True):
    sess.set_profiling(prof_on)
    t0 = time.perf_counter()
    for j in range(K):
        sess.render_samples(steps[j], acc)
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print("profiling spans %s: host issue %.2f ms / step, total %.2f ms / step" % (prof_on, (t1 - t0) * 1e3 / K, (t2 - t0) * 1e3 / K))
sess.set_profiling(False)
pr = cProfile.Profile()
pr.enable()
for j in range(K):
    sess.render_samples(steps[j], acc)
pr.disable()
torch.cuda.synchronize()
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(18)
print(s.getvalue()[:5000])
