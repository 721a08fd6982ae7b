"""
Scene flattening: every object of a ``Scene`` concatenated into one set of vertex / index buffers, the way the
reference's ``RenderSessionMixin.vertex_array_object`` does it (diffrp/rendering/mixin.py:74-113), plus the per-vertex
world-space normals / tangents that ``SurfaceInput.interpolate_ex`` would compute per material
(base_material.py:137-143: 'vectornor', 'vector3norex1') baked once for the fused kernel.

Device-agnostic torch code (runs once per session; not the hot path).
"""
import os
from dataclasses import dataclass
from typing import Dict, List, Optional

import torch

from .ops import transform_point4x3, transform_vector3x3, normalized


@dataclass
class VertexArrayObject:
    """Flattened scene buffers (base_material.py:12-34) + the baked world-space attributes the fused kernel reads."""
    verts: torch.Tensor
    normals: torch.Tensor
    world_pos: torch.Tensor
    tris: torch.Tensor
    stencils: torch.Tensor
    color: torch.Tensor
    uv: torch.Tensor
    tangents: torch.Tensor
    custom_attrs: Dict[str, torch.Tensor]
    world_nrm: torch.Tensor = None
    world_tan: torch.Tensor = None
    tri_material: torch.Tensor = None
    records: torch.Tensor = None  # (V,16) interleaved [pos3 nrm3 uv2 | color4 tan4] shading records (CUDA path)



def flatten_scene(objs: List, dev) -> VertexArrayObject:
    dev = torch.device(dev)
    f32 = lambda t: t.to(dev, torch.float32, non_blocking=True)
    keys = set().union(*(o.custom_attrs.keys() for o in objs)) if objs else set()
    verts, nrms, wpos, tris, sts, cols, uvs, tans, wn, wt, tm = ([] for _ in range(11))
    customs = {k: [] for k in keys}
    sts.append(torch.zeros([1], dtype=torch.int32, device=dev))  # stencil 0 = miss (mixin.py:78)
    offset = 0
    for s, o in enumerate(objs):
        assert o.tris.shape[-1] == 3, "Expected 3 vertices per triangle, got %d" % o.tris.shape[-1]
        v, n, M = f32(o.verts), f32(o.normals), f32(o.M)
        col, uv, tg = f32(o.color), f32(o.uv), f32(o.tangents)
        for name, a, c in (("normals", n, 3), ("uv", uv, 2), ("tangents", tg, 4), ("color", col, None)):
            assert a.shape[0] == v.shape[0], "attribute length not the same as number of vertices"
            assert c is None or a.shape[-1] == c, "expected %s dims but got %d for vertex attribute %s" % (c, a.shape[-1], name)
        if col.shape[-1] == 3:
            col = torch.cat([col, torch.ones_like(col[:, :1])], -1)
        verts.append(v); nrms.append(n); cols.append(col); uvs.append(uv); tans.append(tg)
        wpos.append(transform_point4x3(v, M))
        wn.append(normalized(transform_vector3x3(n, M)))                                     # 'vectornor'
        wt.append(torch.cat([normalized(transform_vector3x3(tg[:, :3], M)), tg[:, 3:]], -1))  # 'vector3norex1'
        t = o.tris.to(dev, torch.int32, non_blocking=True) + offset
        tris.append(t)
        sts.append(torch.full([len(t)], s + 1, dtype=torch.int32, device=dev))
        tm.append(torch.full([len(t)], s, dtype=torch.int32, device=dev))
        for k in keys:
            if k in o.custom_attrs:
                assert len(o.custom_attrs[k]) == len(o.verts), "Attribute length not the same as number of vertices: %s" % k
                customs[k].append(f32(o.custom_attrs[k]))
            else:
                size = next(x.custom_attrs[k].shape[-1] for x in objs if k in x.custom_attrs)
                customs[k].append(torch.zeros([len(v), size], device=dev))
        offset += len(v)
    cat = lambda xs, shape, dt=torch.float32: (torch.cat(xs).contiguous() if xs else torch.zeros(shape, dtype=dt, device=dev))
    return VertexArrayObject(
        cat(verts, [0, 3]), cat(nrms, [0, 3]), cat(wpos, [0, 3]), cat(tris, [0, 3], torch.int32), torch.cat(sts).contiguous(),
        cat(cols, [0, 4]), cat(uvs, [0, 2]), cat(tans, [0, 4]), {k: torch.cat(v) for k, v in customs.items()},
        cat(wn, [0, 3]), cat(wt, [0, 4]), cat(tm, [0], torch.int32))


#: host tensors smaller than this are uploaded whole by every rank even when sharded upload is on (latency-bound anyway)
SHARD_MIN_BYTES = 1 << 20


def merge_runs(items, arena):
    """items: (buffer offset, source address, bytes) in registration order; arena: (first, last + 1) address of a page-locked arena or None.
    Consecutive items whose sources lie in the arena at the same relative offsets as in the buffer become ONE copy (alignment gaps, which
    are arena bytes too, included)."""
    runs = []
    for off, p, nbytes in items:
        in_arena = arena is not None and arena[0] <= p and p + nbytes <= arena[1]
        if runs and in_arena and runs[-1][3] and p - runs[-1][1] == off - runs[-1][0]:
            runs[-1][2] = off + nbytes - runs[-1][0]
        else:
            runs.append([off, p, nbytes, in_arena])
    return [(r[0], r[1], r[2]) for r in runs]


def clip_runs(runs, lo, hi):
    """The part of every (buffer offset, source address, bytes) run that falls into the byte range [lo, hi) of the packed buffer: what ONE rank
    copies when the upload is sharded (the ranks' ranges tile the buffer, so the union of their segments is every byte exactly once)."""
    out = []
    for off, p, nbytes in runs:
        a, b = max(off, lo), min(off + nbytes, hi)
        if b > a:
            out.append((a, p + (a - off), b - a))
    return out


class PackedUpload:
    """
    Host -> device move of MANY scene tensors as one packed device buffer: ``add()`` registers a tensor and returns a ticket, ``commit()``
    allocates one buffer, issues every copy from ONE C call (``drp_upload_batch``) and returns the device tensors (views into the buffer).
    ``shard=(rank, world)`` (scene replicated over the ranks of an NCCL group -- the precondition of spp / tile sharding): every rank copies
    only its 1/world byte range of the packed layout over its own PCIe link and ONE in-place all-gather over NVLink completes the buffer on
    every GPU: the scene crosses PCIe once in total, at one collective instead of one per tensor.  Collective: every rank must register
    the same tensors in the same order.  Tensors that already live on the device, or need a dtype / layout conversion, are passed through
    ``.to()`` immediately.
    """
    ALIGN = 256

    def __init__(self, dev, shard=None, arena: Optional[torch.Tensor] = None):
        self.dev, self.shard = torch.device(dev), shard if (shard is not None and shard[1] > 1) else None
        self.items, self.ready, self.total = [], {}, 0
        # Scene.pin_memory(): sources that are views of one page-locked arena, laid out like this buffer, merge into a few large copies
        self.arena = None if arena is None else (arena.data_ptr(), arena.data_ptr() + arena.numel())

    def add(self, t: torch.Tensor, dtype) -> int:
        ticket = len(self.items) + len(self.ready)
        if self.dev.type != 'cuda' or t.is_cuda or t.dtype != dtype or not t.is_contiguous() or t.numel() == 0:
            self.ready[ticket] = t.to(self.dev, dtype, non_blocking=True).contiguous()
            return ticket
        nbytes = t.numel() * t.element_size()
        self.items.append((ticket, t, self.total, nbytes))
        self.total += -(-nbytes // self.ALIGN) * self.ALIGN
        return ticket

    def commit(self) -> dict:
        out = dict(self.ready)
        if not self.items:
            return out
        import ctypes as C
        from ._lib import lib, check
        world = self.shard[1] if self.shard else 1
        per = -(-self.total // (world * self.ALIGN)) * self.ALIGN
        buf = torch.empty([per * world], dtype=torch.uint8, device=self.dev)
        lo, hi = (self.shard[0] * per, (self.shard[0] + 1) * per) if self.shard else (0, per)
        base = buf.data_ptr()
        segs = clip_runs(merge_runs([(off, t.data_ptr(), nbytes) for _, t, off, nbytes in self.items], self.arena), lo, hi)
        dst, src, nb = [base + a for a, _, _ in segs], [p for _, p, _ in segs], [n for _, _, n in segs]
        n = len(dst)
        if n:
            check(lib().drp_upload_batch(n, (C.c_void_p * n)(*dst), (C.c_void_p * n)(*src), (C.c_int64 * n)(*nb),
                                         torch.cuda.current_stream(self.dev).cuda_stream), "drp_upload_batch")
        if self.shard:
            import torch.distributed as dist
            dist.all_gather_into_tensor(buf, buf[lo:hi])
        for ticket, t, off, nbytes in self.items:
            out[ticket] = buf[off:off + nbytes].view(t.dtype).view(t.shape)
        self._sources = [t for _, t, _, _ in self.items]   # host sources stay referenced until the caller drops this object
        return out


def upload(t: torch.Tensor, dev, dtype, shard=None) -> torch.Tensor:
    """Host -> device move of ONE scene tensor (e.g. the environment map; whole scenes go through ``PackedUpload``).  ``shard=(rank, world)``:
    every rank copies its 1/world slice and an in-place all-gather completes it (collective: same call on every rank)."""
    if (shard is not None and shard[1] > 1 and not t.is_cuda and t.dtype == dtype and t.is_contiguous()
            and t.numel() * t.element_size() >= SHARD_MIN_BYTES):
        import torch.distributed as dist
        rank, world = shard
        n = t.numel()
        per = -(-n // world)
        out = torch.empty([world * per], dtype=dtype, device=dev)
        lo, hi = min(rank * per, n), min((rank + 1) * per, n)
        if hi > lo:
            out[lo:hi].copy_(t.view(-1)[lo:hi], non_blocking=True)
        dist.all_gather_into_tensor(out, out[rank * per:(rank + 1) * per])
        return out[:n].view(t.shape)
    return t.to(dev, dtype, non_blocking=True)


def flatten_scene_cuda(objs: List, dev, shard=None, arena=None) -> VertexArrayObject:
    """
    ``flatten_scene`` in one CUDA pass (``drp_flatten``): sources that already live on ``dev`` are read in place, host tensors are
    first moved with one asynchronous DMA copy each (measured on B200: letting the kernel read pinned host memory directly works --
    the C entry point accepts such pointers -- but 12-byte strided reads over PCIe reach ~11 GB/s versus ~50 GB/s for the DMA).
    Custom vertex attributes (only visible to Python materials) are still concatenated with torch.
    """
    import ctypes as C
    from . import _abi
    from ._lib import lib, check
    dev = torch.device(dev)
    keep = []
    pack = PackedUpload(dev, shard, arena)
    tickets, local = [], {}
    for o in objs:   # pass 1: register every source tensor (tensors shared between objects -- instanced meshes -- are moved once)
        row = []
        for t, dtype in ((o.verts, torch.float32), (o.normals, torch.float32), (o.color, torch.float32), (o.uv, torch.float32),
                         (o.tangents, torch.float32), (o.tris, torch.int32)):
            if t.dtype == dtype and t.is_contiguous() and t.is_cuda and t.device == dev:
                row.append(t)
            else:
                key = (t.data_ptr(), tuple(t.shape), t.dtype, str(t.device))
                if key not in local:
                    local[key] = pack.add(t, dtype)
                row.append(local[key])
        tickets.append(row)
    moved = pack.commit()
    keep.append(pack)

    descs = (_abi.Object * max(1, len(objs)))()
    V = F = 0
    host_M = [o.M.detach() for o in objs]
    if any(m.is_cuda for m in host_M):   # one read-back for all model matrices instead of one per object
        host_M = list(torch.stack([m.to(dev, torch.float32) for m in host_M]).cpu())
    for k, o in enumerate(objs):
        assert o.tris.shape[-1] == 3, "Expected 3 vertices per triangle, got %d" % o.tris.shape[-1]
        v, n, col, uv, tg, tr = (x if isinstance(x, torch.Tensor) else moved[x] for x in tickets[k])
        keep.extend((v, n, col, uv, tg, tr))
        nv = v.shape[0]
        for name, a, c in (("normals", n, 3), ("uv", uv, 2), ("tangents", tg, 4), ("color", col, None)):
            assert a.shape[0] == nv, "attribute length not the same as number of vertices"
            assert c is None or a.shape[-1] == c, "expected %s dims but got %d for vertex attribute %s" % (c, a.shape[-1], name)
        assert col.shape[-1] in (3, 4), "vertex colours must have 3 or 4 channels"
        d = descs[k]
        d.verts, d.normals, d.color, d.uv, d.tangents, d.tris = (v.data_ptr(), n.data_ptr(), col.data_ptr(), uv.data_ptr(),
                                                                 tg.data_ptr(), tr.data_ptr())
        d.M[:] = host_M[k].to(torch.float32).reshape(-1).tolist()
        d.n_verts, d.n_tris, d.color_channels = nv, tr.shape[0], col.shape[-1]
        V += nv
        F += tr.shape[0]
    f = lambda *shape: torch.empty(shape, dtype=torch.float32, device=dev)
    i = lambda *shape: torch.empty(shape, dtype=torch.int32, device=dev)
    out = dict(world_pos=f(V, 3), world_nrm=f(V, 3), color=f(V, 4), uv=f(V, 2), world_tan=f(V, 4), tris=i(F, 3), tri_material=i(F),
               stencils=i(F + 1), records=f(V, 16), verts=f(V, 3), normals=f(V, 3), tangents=f(V, 4))
    check(lib().drp_flatten(descs, len(objs), out['world_pos'].data_ptr(), out['world_nrm'].data_ptr(), out['color'].data_ptr(),
                            out['uv'].data_ptr(), out['world_tan'].data_ptr(), out['tris'].data_ptr(), out['tri_material'].data_ptr(),
                            out['stencils'].data_ptr(), out['records'].data_ptr(), out['verts'].data_ptr(), out['normals'].data_ptr(),
                            out['tangents'].data_ptr(), torch.cuda.current_stream(dev).cuda_stream), "drp_flatten")
    keys = set().union(*(o.custom_attrs.keys() for o in objs)) if objs else set()
    customs = {}
    for key in keys:
        size = next(x.custom_attrs[key].shape[-1] for x in objs if key in x.custom_attrs)
        parts = []
        for o in objs:
            if key in o.custom_attrs:
                assert len(o.custom_attrs[key]) == len(o.verts), "Attribute length not the same as number of vertices: %s" % key
                parts.append(o.custom_attrs[key].to(dev, torch.float32))
            else:
                parts.append(torch.zeros([len(o.verts), size], device=dev))
        customs[key] = torch.cat(parts)
    vao = VertexArrayObject(out['verts'], out['normals'], out['world_pos'], out['tris'], out['stencils'], out['color'], out['uv'],
                            out['tangents'], customs, out['world_nrm'], out['world_tan'], out['tri_material'])
    vao.records = out['records']
    vao._sources = keep  # pinned / converted sources stay alive until the stream has consumed them
    return vao


def pad_rgba(img: torch.Tensor) -> torch.Tensor:
    """(H,W,1|3|4) -> (H,W,4): the fused kernel fetches every texel with one 128-bit load.  One channel broadcasts
    (what ``factor * color * tex`` does in the reference), RGB gets alpha = 1."""
    c = img.shape[-1]
    if c == 4:
        return img.contiguous()
    if c == 1:
        return img.expand(*img.shape[:-1], 4).contiguous()
    assert c == 3, "textures must have 1, 3 or 4 channels"
    return torch.cat([img, torch.ones_like(img[..., :1])], -1).contiguous()


def texel_records(d: dict) -> Optional[torch.Tensor]:
    """(H,W,12) interleaved copy of a GLTF material's four RGBA-padded textures, [base rgba | mr.g mr.b n.x n.y | n.z e.r e.g e.b]
    (drp_material_t.texel_records), or None when a texture is missing or their sizes / wrap / filter modes differ.  A material without
    an emissive factor still qualifies (its emissive texture is interleaved but never read: has_emissive = 0)."""
    tex = [d.get(k) for k in ('base_color_tex', 'mr_tex', 'normal_tex', 'emissive_tex')]
    if d.get('kind') != 'gltf' or any(t is None for t in tex):
        return None
    first = tex[0]
    for t in tex[1:]:
        if tuple(t['image'].shape[:2]) != tuple(first['image'].shape[:2]) or t.get('wrap', 'repeat') != first.get('wrap', 'repeat') \
                or t.get('interp', 'linear') != first.get('interp', 'linear'):
            return None
    b, m, n, e = (t['image'] for t in tex)
    if any(x.shape[-1] != 4 for x in (b, m, n, e)):
        return None
    rec = torch.cat([b, m[..., 1:3], n[..., :3], e[..., :3]], -1).contiguous()
    T = 1 << TEXEL_TILE_LOG2
    if TEXEL_TILE_LOG2 > 0 and rec.shape[0] % T == 0 and rec.shape[1] % T == 0:   # tile-major: (H/T, W/T, T, T, 12)
        d['texel_tile_log2'] = TEXEL_TILE_LOG2
        return rec.view(rec.shape[0] // T, T, rec.shape[1] // T, T, 12).permute(0, 2, 1, 3, 4).contiguous()
    d['texel_tile_log2'] = 0
    return rec


#: interleave the four textures of a material into 48-byte texel records (layout only; tests flip it to prove bit-equality with separate textures)
INTERLEAVE_TEXELS = True
#: texel records are stored in tiles of 2^L x 2^L texels (0 = row-major); layout only, the taps / weights / sums are the same
TEXEL_TILE_LOG2 = 2


def material_descriptions(objs: List, dev, rgba: bool = False, shard=None, arena=None) -> Optional[List[dict]]:
    """Per-object drp_material_t descriptions (textures moved to ``dev``, RGBA-padded for the CUDA path when ``rgba``),
    or None if any material is Python-only.  Objects sharing a material share its uploaded textures."""
    descs = []
    uploaded = {}
    pack = PackedUpload(dev, shard, arena)
    raw = []
    for o in objs:   # pass 1: register every texture; ONE packed upload (+ one all-gather when sharded) moves them all
        d = o.material.fused_description() if hasattr(o.material, 'fused_description') else None
        if d is None:
            return None
        d = dict(d)
        for k in ('base_color_tex', 'mr_tex', 'normal_tex', 'emissive_tex'):
            if d.get(k) is not None:
                src = d[k]['image']
                key = (src.data_ptr(), tuple(src.shape))
                if key not in uploaded:
                    uploaded[key] = pack.add(src, torch.float32)
                d[k] = dict(d[k], _key=key)
        raw.append(d)
    moved = pack.commit()
    uploaded = {key: moved[ticket] for key, ticket in uploaded.items()}
    padded = {}
    for d in raw:    # pass 2: RGBA padding and texel interleaving on the device
        for k in ('base_color_tex', 'mr_tex', 'normal_tex', 'emissive_tex'):
            if d.get(k) is not None:
                key = d[k].pop('_key')
                if key not in padded:
                    padded[key] = pad_rgba(uploaded[key]) if rgba else uploaded[key].contiguous()
                d[k] = dict(d[k], image=padded[key])
        if rgba and INTERLEAVE_TEXELS:  # CUDA path: interleaved texels, shared between the objects that share the material
            key = ('records',) + tuple(d[k]['image'].data_ptr() if d.get(k) is not None else 0 for k in ('base_color_tex', 'mr_tex', 'normal_tex', 'emissive_tex'))
            if key not in padded:
                rec = texel_records(d)
                padded[key] = (rec, d.get('texel_tile_log2', 0))
            if padded[key][0] is not None:
                d['texel_records'], d['texel_tile_log2'] = padded[key]
        descs.append(d)
    if not descs:
        descs = [dict(kind='default', tint=None)]
    return descs
