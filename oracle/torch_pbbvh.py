"""
torch_pbbvh.py -- torch-level restatement of the reference's fallback raycaster ``NaivePBBVH`` (diffrp/utils/raycaster.py:120-260).

TEST / MEASUREMENT INFRASTRUCTURE (same status as the rest of oracle/): the product never imports it.  It exists so that
``bench.py`` can time, on the same B200 and in the same run, the *kind of program* the reference executes when ``torchoptix`` is not
installed -- a stackless BVH walk written as whole-array torch ops with ``nonzero`` compaction -- because the reference itself is a
Python package that does not travel to the GPU box.  Own code, written from the algorithm (also restated in C in orc_raycast.c):

  build   raycaster.py:122-187  pad the triangle list to P = 2^n by wrapping around, order the leaves (longest-axis median splits, or
                                30-bit Morton codes), implicit complete binary heap of AABBs (node k -> children 2k+1, 2k+2, leaves at
                                P-1 .. 2P-2), `scan_next` (first child, or the skip link for a leaf) / `skip_next` (next subtree in DFS
                                order, 0 = finished), per-triangle inverse frame for the unit-triangle test (:27-39)
  query   raycaster.py:226-260  every live ray holds one heap index; per step: slab test (:189-205), unit-triangle test for rays sitting on
                                a leaf (:42-55, :207-224: scatter-amin into t, `test_t <= t` picks the id), follow scan/skip; after
                                ceil(n/2) steps the finished rays are compacted away; the pruning distance is refreshed only then.
Validated against orc_bvh_query(reference_mode=1) in tests/test_oracle_golden.py.
"""
import torch


def _expand_bits10(v):
    v = (v | (v << 16)) & 0x030000FF
    v = (v | (v << 8)) & 0x0300F00F
    v = (v | (v << 4)) & 0x030C30C3
    v = (v | (v << 2)) & 0x09249249
    return v


def _unit_triangle_frames(tri):
    """(P,3,3) triangles -> rotation (P,3,3) and translation (P,3) taking world space to the frame where the triangle is
    (0,0,0),(1,0,0),(0,1,0) in the z = 0 plane (closed-form inverse of [e1 e2 n | A])."""
    a, e1, e2 = tri[:, 0], tri[:, 1] - tri[:, 0], tri[:, 2] - tri[:, 0]
    n = torch.linalg.cross(e1, e2)
    n = n / n.norm(dim=-1, keepdim=True).clamp_min(1e-12)
    r0, r1, r2 = torch.linalg.cross(e2, n), torch.linalg.cross(n, e1), torch.linalg.cross(e1, e2)
    det = (e1 * r0).sum(-1, keepdim=True)
    rot = torch.stack([r0 / det, r1 / det, r2 / det], 1)
    return rot, -(rot @ a[:, :, None])[:, :, 0]


class TorchPBBVH:
    def __init__(self, verts: torch.Tensor, tris: torch.Tensor, builder: str = 'splitaxis'):
        dev = verts.device
        M = tris.shape[0]
        self.n = n = max(0, (M - 1).bit_length())
        P = 1 << n
        tri = verts[tris.long()]                                    # (M,3,3)
        tri = tri[torch.arange(P, device=dev) % M]                  # wrap-around padding
        lo, hi, cen = tri.amin(1), tri.amax(1), tri.mean(1)
        order = torch.arange(P, device=dev)
        if builder == 'morton':
            c = cen - cen.amin(0)
            q = (c / (c.amax(0) + 1e-8) * 1023.0).long()
            code = _expand_bits10(q[:, 0]) | (_expand_bits10(q[:, 1]) << 1) | (_expand_bits10(q[:, 2]) << 2)
            order = torch.argsort(code, stable=True)
        elif builder == 'splitaxis':
            for level in range(n):                                  # one segmented sort per level
                seg = order.view(1 << level, -1)
                ext = hi[seg].amax(1) - lo[seg].amin(1)             # (segments, 3)
                axis = ext.argmax(-1)
                key = cen[seg].gather(2, axis[:, None, None].expand(-1, seg.shape[1], 1))[..., 0]
                order = seg.gather(1, torch.argsort(key, dim=1, stable=True)).reshape(-1)
        else:
            raise ValueError(builder)
        self.rank = (order % M)
        tri = tri[order]
        bmin = torch.empty(2 * P - 1, 3, device=dev)
        bmax = torch.empty(2 * P - 1, 3, device=dev)
        bmin[P - 1:], bmax[P - 1:] = lo[order], hi[order]
        for level in range(n - 1, -1, -1):
            a, b = (1 << level) - 1, (1 << (level + 1)) - 1         # nodes of this level; their children start at b
            bmin[a:b] = torch.minimum(bmin[b:2 * b + 1:2], bmin[b + 1:2 * b + 2:2])
            bmax[a:b] = torch.maximum(bmax[b:2 * b + 1:2], bmax[b + 1:2 * b + 2:2])
        skip = torch.zeros(2 * P - 1, dtype=torch.int64, device=dev)
        for level in range(1, n + 1):
            k = torch.arange((1 << level) - 1, (1 << (level + 1)) - 1, device=dev)
            skip[k] = torch.where(k % 2 == 1, k + 1, skip[(k - 1) // 2])
        scan = 2 * torch.arange(2 * P - 1, device=dev) + 1
        scan[P - 1:] = skip[P - 1:]
        self.bmin, self.bmax, self.skip_next, self.scan_next = bmin, bmax, skip, scan
        self.rot, self.trans = _unit_triangle_frames(tri)
        self.first_leaf = P - 1

    @torch.no_grad()
    def query(self, rays_o: torch.Tensor, rays_d: torch.Tensor, far: float):
        R, dev = rays_o.shape[0], rays_o.device
        t = torch.full((R,), far, dtype=rays_o.dtype, device=dev)
        hit = torch.zeros(R, dtype=torch.int64, device=dev)
        node = torch.zeros(R, dtype=torch.int64, device=dev)
        ray = torch.arange(R, device=dev)
        o, d = rays_o, rays_d
        steps = (self.n + 1) // 2
        while ray.numel() > 0:
            alive = torch.ones_like(node, dtype=torch.bool)
            t_prune = t[ray]                                         # refreshed once per compaction, as in the reference
            for _ in range(steps):
                t1 = (self.bmin[node] - o) / d
                t2 = (self.bmax[node] - o) / d
                near = torch.minimum(t1, t2).amax(-1)
                farr = torch.maximum(t1, t2).amin(-1)
                inside = (near <= t_prune) & (farr > 0) & (near <= farr)
                on_leaf = (inside & (node >= self.first_leaf)).nonzero().squeeze(-1)
                if on_leaf.numel() > 0:
                    leaf = node[on_leaf] - self.first_leaf
                    lo_ = (self.rot[leaf] @ o[on_leaf, :, None])[:, :, 0] + self.trans[leaf]
                    ld_ = (self.rot[leaf] @ d[on_leaf, :, None])[:, :, 0]
                    tt = -lo_[:, 2] / ld_[:, 2]
                    b1, b2 = lo_[:, 0] + tt * ld_[:, 0], lo_[:, 1] + tt * ld_[:, 1]
                    tt = torch.where((tt > 0) & (b1 >= 0) & (b2 >= 0) & (b1 + b2 <= 1), tt, torch.full_like(tt, far))
                    r = ray[on_leaf]
                    t.scatter_reduce_(0, r, tt, 'amin')
                    hit[r] = torch.where(tt <= t[r], leaf, hit[r])
                node = torch.where(inside, self.scan_next[node], self.skip_next[node])
                alive &= node != 0
            keep = alive.nonzero().squeeze(-1)
            node, ray, o, d = node[keep], ray[keep], o[keep], d[keep]
        return t, self.rank[hit]


class TorchPBBVHRaycaster:
    """Raycaster-shaped wrapper (build in the constructor, ``query`` -> (t, i)) for diffrp_b200's generic torch path."""
    def __init__(self, verts, tris, config=None):
        self.impl = TorchPBBVH(verts, tris, (config or {}).get('builder', 'splitaxis'))
        self.handle = -1

    def query(self, rays_o, rays_d, far):
        t, i = self.impl.query(rays_o, rays_d, far)
        return t, i.int()
