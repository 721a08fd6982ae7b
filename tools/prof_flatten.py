import os, sys, time, cProfile, pstats
sys.path.insert(0, os.getcwd())
import torch
import diffrp_b200 as drp
from diffrp_b200 import synthetic as syn
scene, camkw = syn.teaser_scene('cpu', tex=1024)
scene = scene.pin_memory()
cam = drp.PerspectiveCamera.from_orbit(h=1024, w=1024, **camkw)
for it in range(3):
    s = drp.PathTracingSession(scene, cam, drp.PathTracingSessionOptions(ray_spp=8, ray_depth=4, seed=it, reuse_scene=False))
    torch.cuda.synchronize()
    pr = cProfile.Profile(); pr.enable()
    t0 = time.perf_counter(); vao = s.vertex_array_object(); t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    rc = s.raycaster(); t3 = time.perf_counter(); torch.cuda.synchronize(); t4 = time.perf_counter()
    f = s._fused_scene(); t5 = time.perf_counter(); torch.cuda.synchronize(); t6 = time.perf_counter()
    pr.disable()
    print("flatten host %.2f ms (+%.2f to drain) | build host %.2f (+%.2f) | fused scene host %.2f (+%.2f)" % ((t1-t0)*1e3, (t2-t1)*1e3, (t3-t2)*1e3, (t4-t3)*1e3, (t5-t4)*1e3, (t6-t5)*1e3))
    if it == 2:
        pstats.Stats(pr).sort_stats("tottime").print_stats(18)
    s.raycaster().release()
