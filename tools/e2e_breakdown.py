"""Where does an end-to-end pbr() call spend its time?  (host-pinned config-3 scene, one GPU)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import diffrp_b200 as drp
from diffrp_b200 import synthetic as syn

scene, camkw = syn.teaser_scene('cpu', tex=1024)
scene = scene.pin_memory()
cam = drp.PerspectiveCamera.from_orbit(h=1024, w=1024, **camkw)


def timed(label, fn):
    torch.cuda.synchronize(); t0 = time.perf_counter(); r = fn(); torch.cuda.synchronize()
    print("%-28s %8.2f ms" % (label, (time.perf_counter() - t0) * 1e3)); return r


for it in range(3):
    print("--- iteration", it)
    t_all = time.perf_counter()
    s = drp.PathTracingSession(scene, cam, drp.PathTracingSessionOptions(ray_spp=8, ray_depth=4, seed=it, reuse_scene=False))
    vao = timed("flatten (H2D geometry)", s.vertex_array_object)
    timed("raycaster (LBVH build)", s.raycaster)
    timed("fused scene (H2D textures)", s._fused_scene)
    timed("render setup", s._render_setup)
    acc = timed("render_accumulators", s.render_accumulators)
    out = timed("finalize", lambda: s.finalize(acc))
    timed("D2H", lambda: [out[0].cpu(), out[1].cpu()] + [v.cpu() for v in out[2].values()])
    timed("release", lambda: s.raycaster().release())
    torch.cuda.synchronize()
    print("%-28s %8.2f ms" % ("TOTAL", (time.perf_counter() - t_all) * 1e3))
