"""Upper-bound probe for ray reordering: how much faster does the traversal kernel trace the SAME secondary rays when they are sorted
by (direction octant, coarse origin Morton code)?  Bounce-1 rays of the bench scene (config 3) are emulated: primary hits + cosine-
distributed directions, in pixel order.  One JSON line."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import diffrp_b200 as drp
from diffrp_b200 import synthetic as syn, generic

dev = torch.device('cuda')
scene_host, camkw = syn.teaser_scene('cpu', tex=64)
scene = scene_host.to(dev)
R = 1024
cam = drp.PerspectiveCamera.from_orbit(h=R, w=R, **camkw)
sess = drp.PathTracingSession(scene, cam, drp.PathTracingSessionOptions(ray_spp=1, ray_depth=1))
rc, far = sess.raycaster(), sess.camera_far()
ys = torch.linspace(-1 + 1 / R, 1 - 1 / R, R, device=dev).view(R, 1, 1).expand(R, R, 1)
xs = torch.linspace(-1 + 1 / R, 1 - 1 / R, R, device=dev).view(1, R, 1).expand(R, R, 1)
grid = torch.cat([xs, ys, -torch.ones_like(xs), torch.ones_like(xs)], -1).reshape(-1, 4)
o0, d0 = generic.primary_rays(sess, grid)
o0, d0 = o0.expand_as(d0).contiguous().repeat(4, 1), d0.contiguous().repeat(4, 1)      # 4 spp worth of rays
t, i = rc.query(o0, d0, far)
hit = t < far
vao = sess.vertex_array_object()
tri = vao.tris[i[hit].long()].long()
p = vao.world_pos
n = torch.nn.functional.normalize(torch.linalg.cross(p[tri[:, 1]] - p[tri[:, 0]], p[tri[:, 2]] - p[tri[:, 0]]), dim=-1)
n = torch.where((n * d0[hit]).sum(-1, keepdim=True) > 0, -n, n)
g = torch.Generator(device=dev).manual_seed(0)
v = torch.nn.functional.normalize(torch.randn(n.shape, device=dev, generator=g), dim=-1)
d1 = torch.nn.functional.normalize(n + v * 0.999, dim=-1)                                   # cosine-distributed about n
o1 = (o0[hit] + d0[hit] * t[hit, None] + d1 * 1e-3).contiguous()
d1 = d1.contiguous()


def timed(o, d, iters=5):
    for _ in range(2):
        rc.query(o, d, far)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for _ in range(iters):
        e0.record(); rc.query(o, d, far); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


def morton(o, bits):
    lo, hi = o.amin(0), o.amax(0)
    q = ((o - lo) / (hi - lo + 1e-9) * ((1 << bits) - 1)).long()
    code = torch.zeros(len(o), dtype=torch.long, device=dev)
    for b in range(bits):
        for a in range(3):
            code |= ((q[:, a] >> b) & 1) << (3 * b + a)
    return code


octant = ((d1[:, 0] < 0).long() << 2) | ((d1[:, 1] < 0).long() << 1) | (d1[:, 2] < 0).long()
out = {"rays": len(o1), "primary_ms": timed(o0, d0), "primary_rays": len(o0), "pixel_order_ms": timed(o1, d1)}
for name, key in (("octant", octant), ("octant_morton5", (octant << 15) | morton(o1, 5)), ("morton5_octant", (morton(o1, 5) << 3) | octant),
                  ("morton7_octant", (morton(o1, 7) << 3) | octant), ("octant_morton7", (octant << 21) | morton(o1, 7))):
    perm = torch.argsort(key)
    out[name + "_ms"] = timed(o1[perm].contiguous(), d1[perm].contiguous())
perm = torch.randperm(len(o1), device=dev)
out["shuffled_ms"] = timed(o1[perm].contiguous(), d1[perm].contiguous())
print(json.dumps(out))
