"""
``PathTracingSession`` / ``PathTracingSessionOptions`` / ``RayOutputs`` with the API of the reference
(diffrp/rendering/path_tracing.py:23-367), driving the hand-written sm_100a kernels of ``libdiffrp_b200.so``.

Two execution paths behind the same methods:

* **fused** (``pbr()`` when every material is a built-in ``DefaultMaterial`` / ``GLTFMaterial``): the whole
  section x bounce loop of ``trace_rays`` + ``sampler_brdf`` runs as CUDA kernels (``drp_render``); PyTorch only
  allocates tensors and supplies the stream.
* **generic** (``trace_rays(sampler)`` with a user sampler, or scenes with custom Python materials): the reference's
  protocol is kept -- the sampler is called per bounce with ``(rays_o, rays_d, t, i, d)`` -- and only the
  intersection runs in CUDA (``B200Raycaster``).  See ``diffrp_b200.generic``.

There is no CPU fallback: without the CUDA library / a CUDA device the session raises.
"""
import math
import ctypes as C

import numpy
from dataclasses import dataclass
from typing import Callable, Dict, Optional

import torch

from . import _abi
from ._lib import lib, check
from .camera import Camera
from .ops import small_matrix_inverse
from .raycaster import B200Raycaster, _stream_ptr
from .scene import Scene, ImageEnvironmentLight
from .flatten import VertexArrayObject, flatten_scene, flatten_scene_cuda, material_descriptions, pad_rgba


@dataclass
class PathTracingSessionOptions:
    """
    Same fields, defaults and meaning as the reference's options (path_tracing.py:23-84), plus B200 extensions.

    ``raycaster_impl``: ``'b200'`` (default).  The reference's names ``'torchoptix'``, ``'naive-pbbvh'`` and
    ``'brute-force'`` are accepted so existing option objects keep working; all of them select the B200 LBVH
    raycaster, whose closest-hit contract is the brute-force one (min t, then min primitive id).
    ``raycaster_builder`` and ``optix_log_level`` are accepted (one builder exists; the level controls verbosity).

    Extensions:
        rng (str): ``'native'`` -- counter-based Philox4x32-10 keyed by (seed; pixel, sample, bounce), no RNG tensors
            in HBM; ``'torch'`` -- draw the six ``torch.rand((R,1))`` tensors per bounce exactly like the reference's
            sampler (path_tracing.py:205-223) and replay them in the kernel: with the same ``torch.manual_seed`` the image
            reproduces the reference's within fp32 tolerance.
        seed (int): key of the native RNG.  NOTE (difference from the reference, which advances torch's global generator): two ``pbr()``
            calls with the same seed return the same noise; pass ``seed=None`` to draw a fresh key from torch's generator per session
            (``torch.manual_seed`` then makes a sequence of renders reproducible as a whole, and averaging renders reduces variance).
        reproducible (bool): the accumulators are updated with fp32 atomic adds, so the low bits of an image depend on the order in
            which rays of the same pixel retire.  With ``reproducible=True`` every launch batch holds one sample per pixel, which fixes the
            order (bounce by bounce, sample by sample) at the price of smaller launches.
        reuse_scene (bool): sessions are single-use like the reference's, but the flattened buffers, uploaded textures and the BVH
            of a ``Scene`` are kept (one entry per scene, keyed by the identity, shape and in-place version counter of every tensor)
            and adopted by the next session over the same unmodified scene -- multi-view rendering builds once, not per view.
        refit_scene (bool): when a later session renders the same ``Scene`` object and only positions / transforms / attributes changed (every
            object's index tensor is the same, unmodified tensor), the structure of the previous session is refitted -- boxes and triangle
            records recomputed bottom-up over the old topology (``drp_refit``) -- instead of rebuilt.  Hits are exact either way.
        instancing (bool): objects that share their vertex and index tensors are built as instances of one mesh (``drp_build_instanced``): the
            result is still one world-space hierarchy with the flattened primitive ids, the sort / collapse cost is paid once per mesh.
        compaction (bool): drop rays that provably contribute nothing to any output (exact; the reference keeps
            tracing them with zero throughput).
        scene_upload: with ``shard_world > 1`` the scene is replicated; when it still lives in (pinned) host memory, ``'sharded'`` (what ``'auto'``
            picks under an NCCL group) makes each rank upload 1/world of every large tensor and all-gather the rest over NVLink instead of
            ``world`` full PCIe uploads from the same host.  Requires identical scenes on all ranks (already the contract of sharding).
        shard_rank / shard_world: this process renders its share of the frame (scene replicated) and the fp32 accumulators
            are summed with ``torch.distributed.all_reduce`` when a process group exists.  ``shard_mode='spp'``: global
            sample indices ``rank::world`` of every pixel; ``shard_mode='tile'``: every sample of the ``tile_size``^2 tiles
            ``rank::world`` (row-major tile order), the rest of the accumulator stays zero; the exchange is then an all-gather of the owned
            tiles when ``tile_collective='gather'``.  Measured on 8 B200 at 4K (531 MB frames, profiles/r2/c5_n8*.json): all-reduce 1.47 ms
            (NVLS: the switch reduces), gather 1.90 ms (17 pack + 118 unpack copies per rank around a 465 MB all-gather) -- so the default is
            the all-reduce; at 2 GPUs 1.17 vs 2.08 ms.
        result_rank: when only one rank consumes the frame, a ``reduce`` to it replaces the all-reduce and the other ranks skip the epilogue.
    """
    ray_depth: int = 3
    ray_spp: int = 16
    ray_split_size: int = 8 * 1024 * 1024
    deterministic: bool = True
    pbr_ray_step_epsilon: float = 1e-3
    pbr_ray_last_bounce: str = 'void'
    raycaster_impl: str = 'b200'
    raycaster_epsilon: float = 1e-8
    raycaster_builder: str = 'splitaxis'
    optix_log_level: int = 3
    rng: str = 'native'
    seed: Optional[int] = 0
    compaction: bool = True
    shard_rank: int = 0
    shard_world: int = 1
    shard_mode: str = 'spp'   # 'spp': samples rank::world of every pixel; 'tile': all samples of the tiles rank::world
    tile_size: int = 256      # tile edge in pixels for shard_mode='tile'
    tile_collective: str = 'allreduce'  # 'allreduce': sum whole frames (NVSwitch reduces in the fabric); 'gather': all-gather of the owned tiles
    result_rank: Optional[int] = None   # sharded renders: None = every rank gets the frame (all-reduce); r = only rank r does (reduce), the others' pbr() returns None
    refit_scene: bool = True          # a later session over the same Scene with unchanged connectivity refits the structure instead of rebuilding
    instancing: bool = False          # objects sharing vertex + index tensors: one hierarchy per mesh, replicated and refitted per instance
    scene_upload: str = 'auto'        # host scenes under sharding: 'sharded' = 1/world of every tensor per rank over PCIe + all-gather over NVLink
    reuse_scene: bool = True  # share the flattened scene + BVH between sessions over the same, unmodified Scene
    reproducible: bool = False  # bit-identical images run to run: one sample per launch batch, fixed fp32 accumulation order


@dataclass
class RayOutputs:
    """Sampler protocol outputs (path_tracing.py:87-124)."""
    radiance: torch.Tensor
    transfer: torch.Tensor
    next_rays_o: torch.Tensor
    next_rays_d: torch.Tensor
    alpha: torch.Tensor
    extras: Dict[str, torch.Tensor]


def _bit_reverse32(i: torch.Tensor) -> torch.Tensor:
    """Van der Corput radical inverse numerator: reverse the low 32 bits of an int64 tensor (light_transport.py:10-17)."""
    b = i
    for shift, mask in ((16, 0x0000FFFF), (8, 0x00FF00FF), (4, 0x0F0F0F0F), (2, 0x33333333), (1, 0x55555555)):
        b = ((b & mask) << shift) | ((b >> shift) & mask)
    return b


def hammersley(n: int, deterministic: bool = False, device=None):
    """x_i = i/n, y_i = radical inverse of i; one global random shift when not deterministic (light_transport.py:20-32)."""
    i = torch.arange(n, dtype=torch.int64, device=device)
    x = i.float() * (1 / n)
    y = _bit_reverse32(i).float() * 2.3283064365386963e-10
    if not deterministic:
        x = (x + torch.rand(1, dtype=x.dtype, device=x.device)) % 1.0
        y = (y + torch.rand(1, dtype=x.dtype, device=x.device)) % 1.0
    return x, y


def raygen_tables(V: torch.Tensor, P: torch.Tensor, H: int, W: int, spp: int, deterministic: bool, device) -> dict:
    """
    Everything the primary-ray generator needs, computed with the same torch / numpy calls as the reference so the
    values are identical: camera position and inv(VP) (mixin.py:31-39, host numpy inverse), far / near (mixin.py:38,44),
    pixel-centre NDC ramps (coordinates.py:6-10) and the per-sample Hammersley offsets (path_tracing.py:317,329).
    ``V`` / ``P`` may live on the host (the session passes host copies): nothing here then waits for the device -- the ramps and offsets
    are produced on ``device`` by asynchronous kernels -- so the set-up of the next single-use session overlaps the previous one's render.
    """
    V, P = V.to(torch.float32), P.to(torch.float32)
    inv = small_matrix_inverse(torch.stack([V, torch.mm(P, V)]))
    qx, qy = hammersley(spp, deterministic, device)
    return dict(
        cam_pos=[float(x) for x in inv[0, :3, 3].cpu()],
        inv_vp=[float(x) for x in inv[1].reshape(-1).cpu()],
        t_far=(P[2, 3] / (P[2, 2] + 1)).item(),
        t_near=(P[2, 3] / (P[2, 2] - 1)).item(),
        ndc_x=torch.linspace(-1 + 1 / W, 1 - 1 / W, W, dtype=torch.float32, device=device),
        ndc_y=torch.linspace(-1 + 1 / H, 1 - 1 / H, H, dtype=torch.float32, device=device),
        jitter_x=((qx - 0.5) * (2 / W)).contiguous(),
        jitter_y=((qy - 0.5) * (2 / H)).contiguous(),
    )


def shard_sample_ids(spp: int, rank: int, world: int, device=None) -> torch.Tensor:
    """Global Hammersley indices rendered by ``rank`` of ``world``: a strided share, so the union over ranks is exactly
    the single-process sample set and every rank sees the whole [0,1) range of the low-discrepancy sequence."""
    ids = torch.arange(spp, dtype=torch.int32, device=device)
    return ids[rank::world] if world > 1 else ids


def reduce_accumulators(accum: torch.Tensor, world: int, dst: Optional[int] = None) -> torch.Tensor:
    """The path's one exchange step: sum the packed fp32 accumulators over all ranks (NCCL on GPUs, gloo in CPU tests).
    ``dst``: only that rank receives the sum (``reduce``); None: every rank does (``all_reduce``)."""
    if world > 1:
        import torch.distributed as dist
        if not (dist.is_available() and dist.is_initialized()):
            raise RuntimeError("shard_world=%d but torch.distributed is not initialised: the image would hold 1/%d of the samples. "
                               "Initialise a process group, or combine render_accumulators() of the shards yourself." % (world, world))
        if dist.get_world_size() != world:
            raise RuntimeError("shard_world=%d does not match the process group's world size %d" % (world, dist.get_world_size()))
        if dst is None:
            dist.all_reduce(accum, op=dist.ReduceOp.SUM)
        else:
            dist.reduce(accum, dst=dst, op=dist.ReduceOp.SUM)
    return accum


#: pixels one drp_render call may cover (csrc/wavefront.cu: WF_MAX_BATCH_RAYS)
MAX_PIXELS_PER_CALL = 1 << 24


def frame_tiles(H: int, W: int, tile: int):
    """Row-major list of (x0, y0, w, h) tiles of an H x W frame (y counted from the bottom row, like the accumulator)."""
    T = max(1, int(tile))
    return [(x, y, min(T, W - x), min(T, H - y)) for y in range(0, H, T) for x in range(0, W, T)]


def tile_rows(tiles, W: int, device=None) -> torch.Tensor:
    """Accumulator row indices (y * W + x) of the pixels of the given tiles, tile after tile, row-major inside a tile."""
    parts = [(torch.arange(y0, y0 + h, device=device).view(h, 1) * W + torch.arange(x0, x0 + w, device=device).view(1, w)).reshape(-1)
             for (x0, y0, w, h) in tiles]
    return torch.cat(parts) if parts else torch.zeros([0], dtype=torch.int64, device=device)


def gather_tile_accumulators(accum: torch.Tensor, H: int, W: int, tile: int, rank: int, world: int) -> torch.Tensor:
    """
    Exchange step of tile sharding: rank r owns tiles r::world (row-major) and every other accumulator row of its frame is zero, so the sum
    over ranks is a GATHER of disjoint supports.  Each rank packs the rows of its own tiles, one ``all_gather`` moves them, and every rank
    scatters the others' rows into its frame: (world-1)/world x frame bytes received per rank and no arithmetic, instead of an all-reduce
    of the whole frame (2 x (world-1)/world x frame bytes through every link plus the adds; 531 MB per frame at 4K).
    The result equals ``all_reduce(SUM)`` of the same buffers bit for bit (x + 0 = x).
    """
    if world <= 1:
        return accum
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        raise RuntimeError("shard_world=%d but torch.distributed is not initialised" % world)
    if dist.get_world_size() != world:
        raise RuntimeError("shard_world=%d does not match the process group's world size %d" % (world, dist.get_world_size()))
    tiles = frame_tiles(H, W, tile)
    C = accum.shape[-1]
    frame = accum.view(H, W, C)
    counts = [sum(w * h for (_, _, w, h) in tiles[r::world]) for r in range(world)]
    n_max = max(counts)
    recv = accum.new_empty([world, n_max, C])

    def segments(r):   # (tile rectangle, its slice of rank r's packed rows): tiles are rectangles, so packing is one strided copy per tile
        off = 0
        for (x0, y0, w, h) in tiles[r::world]:
            yield (x0, y0, w, h), off
            off += w * h
    for (x0, y0, w, h), off in segments(rank):
        recv[rank, off:off + w * h].view(h, w, C).copy_(frame[y0:y0 + h, x0:x0 + w])
    if dist.get_backend() == 'nccl':
        dist.all_gather_into_tensor(recv.view(-1), recv[rank].reshape(-1))   # in place: the input is this rank's slot of the output
    else:
        _all_gather_generic(recv, recv[rank].clone(), world)
    for r in range(world):
        if r != rank:
            for (x0, y0, w, h), off in segments(r):
                frame[y0:y0 + h, x0:x0 + w].copy_(recv[r, off:off + w * h].view(h, w, C))
    return accum


def _all_gather_generic(recv: torch.Tensor, send: torch.Tensor, world: int):
    import torch.distributed as dist
    parts = [torch.empty_like(send) for _ in range(world)]
    dist.all_gather(parts, send)
    for r in range(world):
        recv[r] = parts[r]


def _scene_signature(scene: Scene, device) -> tuple:
    """Changes whenever an object, a tensor (identity, shape or in-place version), a material or a light of the scene changes."""
    sig = [str(device)]

    def t(x):
        return (x.data_ptr(), tuple(x.shape), x._version, str(x.device)) if isinstance(x, torch.Tensor) else repr(x)
    for o in scene.objects:
        sig.append((id(o), id(o.material), t(o.verts), t(o.tris), t(o.normals), t(o.M), t(o.color), t(o.uv), t(o.tangents),
                    tuple(sorted((k, t(v)) for k, v in (o.custom_attrs or {}).items()))))
        m = o.material
        for name in ('tint', 'base_color_factor', 'emissive_factor', 'metallic_factor', 'roughness_factor', 'alpha_cutoff', 'alpha_mode'):
            if hasattr(m, name):
                sig.append(t(getattr(m, name)))
        for name in ('base_color_texture', 'metallic_roughness_texture', 'normal_texture', 'emissive_texture'):
            smp = getattr(m, name, None)
            if smp is not None:
                sig.append((t(smp.image), smp.wrap_mode, smp.interpolation))
    for l in scene.lights:
        sig.append((id(l), t(getattr(l, 'image', None)), t(l.color), repr(l.intensity) if not isinstance(l.intensity, torch.Tensor) else t(l.intensity)))
    return tuple(sig)


#: id(scene) -> (weakref to the scene, signature, {cache key: value}); one entry per live Scene object
_SCENE_CACHE: Dict[int, tuple] = {}


def _topology_signature(scene: Scene) -> tuple:
    """What a refit keeps: per object the identity / shape / version of its index array and its vertex count."""
    return tuple((o.tris.data_ptr(), tuple(o.tris.shape), o.tris._version, int(o.verts.shape[0])) for o in scene.objects)


def _shared_scene_cache(scene: Scene, device, epsilon: float, refit: bool = True) -> dict:
    import weakref
    sig = _scene_signature(scene, device) + (float(epsilon),)
    topo = _topology_signature(scene)
    key = id(scene)
    ent = _SCENE_CACHE.get(key)
    if ent is not None and ent[0]() is scene and ent[1] == sig:
        return ent[2]
    store: dict = {}
    if refit and ent is not None and ent[0]() is scene and ent[3] == topo and ent[1][0] == sig[0] and ent[1][-1] == sig[-1]:
        # same device, same epsilon, same connectivity -- only positions / transforms / attributes / materials changed: the next session
        # re-flattens and REFITS the previous structure (drp_refit) instead of sorting and collapsing again
        old = ent[2].get('raycaster') or ent[2].get('raycaster_to_refit')
        if old is not None and old.handle is not None and not getattr(old, 'instanced', False) and len(old.tris) >= 2:
            store['raycaster_to_refit'] = old
    _SCENE_CACHE[key] = (weakref.ref(scene, lambda _r, k=key: _SCENE_CACHE.pop(k, None)), sig, store, topo)
    return store


def scene_instances(objects):
    """
    Instance table of a flattened scene, or None when instancing would not pay: objects that share their vertex AND index tensors are
    copies of one mesh (any transform / material).  Returns (first_tri (n + 1,) int64, mesh (n,) int32) as numpy arrays for
    ``drp_build_instanced``.  Used when at least 8 objects share meshes 4 : 1 or better (BASELINE configs[4]: 1000 objects, 1 mesh).
    """
    import numpy as np
    if len(objects) < 8:
        return None
    keys, mesh, first = {}, [], [0]
    for o in objects:
        if o.tris.shape[0] < 2:
            return None
        k = (o.verts.data_ptr(), tuple(o.verts.shape), o.tris.data_ptr(), tuple(o.tris.shape))
        mesh.append(keys.setdefault(k, len(keys)))
        first.append(first[-1] + int(o.tris.shape[0]))
    if 4 * len(keys) > len(objects):
        return None
    return np.asarray(first, dtype=np.int64), np.asarray(mesh, dtype=np.int32)


_SIDE_STREAMS: Dict[int, "torch.cuda.Stream"] = {}


def _side_stream(device) -> "torch.cuda.Stream":
    """One upload stream per device for the texture path of ``_build_fused_scene`` (always entered with ``wait_stream(current)``, so memory
    freed by an earlier session is never reused before the caller's stream has passed that session's kernels)."""
    idx = device.index if device.index is not None else torch.cuda.current_device()
    if idx not in _SIDE_STREAMS:
        _SIDE_STREAMS[idx] = torch.cuda.Stream(device=device)
    return _SIDE_STREAMS[idx]


def _cached(fn):
    name = fn.__name__

    def wrapped(self):
        cache = self.__dict__.setdefault('_cache', {})
        if name not in cache:
            cache[name] = fn(self)
        return cache[name]
    wrapped.__name__ = name
    wrapped.__doc__ = fn.__doc__
    return wrapped


class PathTracingSession:
    """Path tracing render session: single use, like the reference's sessions."""

    def __init__(self, scene: Scene, camera: Camera, options: Optional[PathTracingSessionOptions] = None) -> None:
        self.scene = scene
        self.camera = camera
        self.options = options if options is not None else PathTracingSessionOptions()
        if not torch.cuda.is_available():
            raise RuntimeError("diffrp_b200.PathTracingSession needs a CUDA device: there is no CPU fallback")
        self.device = torch.device('cuda', torch.cuda.current_device())
        #: rng='torch' only: callable(shape) -> uniforms; defaults to torch.rand on the CUDA device (what the reference's
        #: rand_like draws).  Tests inject the CPU generator here to replay golden images made by the CPU reference.
        self.uniform_source = None

    # ---- camera (mixin.py:19-44) ---------------------------------------------------------------------------------
    @_cached
    def _camera_host(self):
        """(V, P) as host fp32 tensors.  Cameras of this package expose them without touching the device; a foreign Camera object
        (only V() / P(), possibly CUDA tensors) costs one read-back."""
        cam = self.camera
        V = cam.V_host() if hasattr(cam, 'V_host') else cam.V().detach().cpu().numpy()
        Pm = cam.P_host() if hasattr(cam, 'P_host') else cam.P().detach().cpu().numpy()
        return torch.from_numpy(numpy.ascontiguousarray(V, dtype=numpy.float32)), torch.from_numpy(numpy.ascontiguousarray(Pm, dtype=numpy.float32))

    @_cached
    def camera_V(self):
        return self._camera_host()[0].pin_memory().to(self.device, non_blocking=True)

    @_cached
    def camera_P(self):
        return self._camera_host()[1].pin_memory().to(self.device, non_blocking=True)

    @_cached
    def camera_VP(self):
        return torch.mm(self.camera_P(), self.camera_V())

    @_cached
    def camera_far(self) -> float:
        p = self._camera_host()[1]      # same fp32 operations as mixin.py:44, on the host copy: no device round trip
        return (p[2, 3] / (p[2, 2] + 1)).item()

    @_cached
    def camera_near(self) -> float:
        p = self._camera_host()[1]
        return (p[2, 3] / (p[2, 2] - 1)).item()

    # ---- scene flattening (mixin.py:74-113) ------------------------------------------------------------------------
    def _scene_store(self) -> dict:
        """Per-scene cache shared between sessions (options.reuse_scene), else private to this session."""
        if '_store' not in self.__dict__:
            self._store = (_shared_scene_cache(self.scene, self.device, self.options.raycaster_epsilon, self.options.refit_scene)
                           if self.options.reuse_scene and not self._wants_grad() else {})
        return self._store

    def _upload_shard(self):
        """(rank, world) when host scene tensors are uploaded 1/world per rank + all-gather (flatten.upload), else None."""
        opt = self.options
        if opt.shard_world <= 1 or opt.scene_upload == 'direct':
            return None
        if opt.scene_upload not in ('auto', 'sharded'):
            raise ValueError("scene_upload must be 'auto', 'direct' or 'sharded'")
        import torch.distributed as dist
        ok = dist.is_available() and dist.is_initialized() and dist.get_backend() == 'nccl' and dist.get_world_size() == opt.shard_world
        if not ok:
            if opt.scene_upload == 'sharded':
                raise RuntimeError("scene_upload='sharded' needs an initialised NCCL process group of shard_world ranks")
            return None
        return (opt.shard_rank, opt.shard_world)

    def vertex_array_object(self) -> VertexArrayObject:
        st = self._scene_store()
        if 'vao' not in st:
            # one drp_flatten pass; its torch twin (differentiable) when some scene tensor requires grad
            if self._wants_grad():
                st['vao'] = flatten_scene(self.scene.objects, self.device)
            else:
                st['vao'] = flatten_scene_cuda(self.scene.objects, self.device, self._upload_shard(), getattr(self.scene, '_arena', None))
        return st['vao']

    # ---- the Raycaster seam (path_tracing.py:142-156) --------------------------------------------------------------
    def raycaster(self):
        impl = self.options.raycaster_impl
        if impl not in ('b200', 'torchoptix', 'naive-pbbvh', 'brute-force'):
            raise ValueError("unknown raycaster_impl: %r" % (impl,))
        st = self._scene_store()
        rc = st.get('raycaster')
        if rc is None or rc.handle is None:
            vao = self.vertex_array_object()
            old = st.pop('raycaster_to_refit', None)
            if old is not None and old.handle is not None and tuple(old.verts.shape) == tuple(vao.world_pos.shape):
                old.refit(vao.world_pos.detach())      # same connectivity as the previous session over this scene: keep the topology
                rc = st['raycaster'] = old
                return rc
            cfg = {'epsilon': self.options.raycaster_epsilon, 'builder': self.options.raycaster_builder,
                   'optix_log_level': self.options.optix_log_level}
            if self.options.instancing:
                inst = scene_instances(self.scene.objects)
                if inst is not None:
                    cfg['instances'] = inst
            rc = st['raycaster'] = B200Raycaster(vao.world_pos.detach(), vao.tris, cfg)
        return rc

    def _single_env_light(self):
        st = self._scene_store()
        if 'env' not in st:
            st['env'] = self._load_env_light()
        self._wait_textures()   # (the fused path loads it on the upload stream)
        return st['env']

    def _load_env_light(self):
        env = None
        for light in self.scene.lights:
            if isinstance(light, ImageEnvironmentLight):
                if env is not None:
                    raise ValueError("Only one environment light is supported in path tracing now.")
                from .flatten import upload
                env = upload(light.image_rh().contiguous(), self.device, torch.float32, self._upload_shard()).contiguous()
        return env  # None == the reference's 16x16 black texture

    # ---- fused path ------------------------------------------------------------------------------------------------
    def _fused_scene(self):
        st = self._scene_store()
        if 'fused' not in st:
            st['fused'] = self._build_fused_scene()
        return st['fused']

    def _build_fused_scene(self):
        """drp_scene_t for the fused kernels, or None when some material only exists as Python code.

        Textures and the environment map are first read by the first ``k_shade`` launch, long after the geometry: they are uploaded (one packed
        DMA), RGBA-padded and interleaved on a SIDE stream that starts where the caller's stream stands, so that the copy engine moves the
        0.45 GB of texels while the SMs flatten the geometry and build the hierarchy on the caller's stream; ``_wait_textures`` joins the two
        before the first kernel that samples them."""
        main = torch.cuda.current_stream(self.device)
        side = _side_stream(self.device)
        side.wait_stream(main)
        vao = self.vertex_array_object()   # geometry first: the copy engine serves requests in issue order, and the SMs need it first
        with torch.cuda.stream(side):
            descs = material_descriptions(self.scene.objects, self.device, rgba=True, shard=self._upload_shard(),
                                          arena=getattr(self.scene, '_arena', None))
            if descs is not None:
                env = self._single_env_light()
                env_desc = None if env is None else dict(image=pad_rgba(env))
            ready = side.record_event()
        if descs is None:
            main.wait_event(ready)
            return None
        self._scene_store()['textures_ready'] = ready
        records = vao.records if vao.records is not None else torch.cat([vao.world_pos, vao.world_nrm, vao.uv, vao.color, vao.world_tan], 1).contiguous()  # (V,16), 64 B / vertex
        arrays = dict(world_pos=vao.world_pos, world_nrm=vao.world_nrm, color=vao.color, uv=vao.uv, world_tan=vao.world_tan,
                      tris=vao.tris, tri_material=vao.tri_material, vertex_records=records)
        return _abi.pack_scene(arrays, descs, env_desc, lambda t: t.data_ptr())

    def _wait_textures(self):
        """Order the current stream after the side-stream texture upload of this scene (no-op once it has been waited for on this stream)."""
        st = self._scene_store()
        ev = st.get('textures_ready')
        if ev is not None:
            torch.cuda.current_stream(self.device).wait_event(ev)

    def _section_sample_ids(self, ids: torch.Tensor):
        """Split like the reference: sections = min(spp, ceil(H*W*spp / ray_split_size)) (path_tracing.py:318-320)."""
        H, W = self.camera.resolution()
        n = len(ids)
        if n == 0:
            return []
        sections = min(n, math.ceil((H * W * n) / self.options.ray_split_size))
        return list(torch.tensor_split(ids, max(1, sections)))

    @_cached
    def _render_setup(self):
        """Per-session constants of the fused path: scene struct, raygen tables, parameter block."""
        opt = self.options
        fused = self._fused_scene()
        if fused is None:
            raise RuntimeError("scene contains custom Python materials: use pbr() / trace_rays() (generic path)")
        H, W = self.camera.resolution()
        Vh, Ph = self._camera_host()
        tab = raygen_tables(Vh, Ph, H, W, opt.ray_spp, opt.deterministic, self.device)
        p = _abi.RenderParams()
        p.height, p.width, p.ray_depth = H, W, opt.ray_depth
        p.last_bounce_skybox = int(opt.pbr_ray_last_bounce == 'skybox')
        p.compaction = int(opt.compaction)
        p.reproducible = int(opt.reproducible)
        p.step_epsilon, p.t_far, p.t_near = opt.pbr_ray_step_epsilon, tab['t_far'], tab['t_near']
        p.cam_pos[:3] = tab['cam_pos']
        p.inv_vp[:] = tab['inv_vp']
        if opt.seed is None:  # fresh key per session, drawn from torch's (seedable) CPU generator
            if '_drawn_seed' not in self.__dict__:
                self._drawn_seed = int(torch.randint(0, 2 ** 62, (1,)).item())
            p.seed = self._drawn_seed
        else:
            p.seed = opt.seed
        p.ndc_x, p.ndc_y = tab['ndc_x'].data_ptr(), tab['ndc_y'].data_ptr()
        return fused[0], tab, p, fused[1]

    def new_accumulators(self) -> torch.Tensor:
        H, W = self.camera.resolution()
        return torch.zeros([H * W, _abi.ACCUM_CHANNELS], dtype=torch.float32, device=self.device)

    def tiles(self):
        """Row-major list of (x0, y0, w, h) tiles of the frame (y counted from the bottom row, like the accumulator)."""
        H, W = self.camera.resolution()
        return frame_tiles(H, W, self.options.tile_size)

    def render_samples(self, sample_ids: torch.Tensor, accum: Optional[torch.Tensor] = None, tile=None) -> torch.Tensor:
        """
        Progressive entry point of the fused path: add the given GLOBAL sample indices (into the n = ``ray_spp``
        Hammersley sequence) of every pixel to ``accum`` (H*W, 16) and return it.  Asynchronous on the current stream.
        ``pbr()`` is ``finalize(all_reduce(render_samples(my share)))``.  ``tile=(x0, y0, w, h)`` restricts the call to a
        pixel rectangle (tile sharding); RNG streams and accumulator rows are keyed by the global pixel either way.
        """
        opt, dev = self.options, self.device
        scene_struct, tab, p, _keep = self._render_setup()
        H, W = self.camera.resolution()
        rc = self.raycaster()
        if accum is None:
            accum = self.new_accumulators()
        ids = sample_ids.to(dev, torch.int32).contiguous()
        if len(ids) == 0:
            return accum
        if opt.rng == 'torch':
            chunks = self._section_sample_ids(ids)
        elif opt.rng == 'native':
            chunks = [ids]
        else:
            raise ValueError("rng must be 'native' or 'torch'")
        L, stream = lib(), _stream_ptr(dev)
        p.tile_x0, p.tile_y0, p.tile_w, p.tile_h = tile if tile is not None else (0, 0, 0, 0)
        if tile is not None and opt.rng == 'torch':
            raise ValueError("rng='torch' (reference replay) renders whole frames only")
        for ids in chunks:
            ids = ids.contiguous()
            idl = ids.long()
            jx, jy = tab['jitter_x'][idl].contiguous(), tab['jitter_y'][idl].contiguous()
            p.n_samples = len(ids)
            p.jitter_x, p.jitter_y, p.sample_ids = jx.data_ptr(), jy.data_ptr(), ids.data_ptr()
            if opt.rng == 'torch':
                R = len(ids) * H * W
                # the reference draws six rand_like((R,1)) per bounce, bounce-major (path_tracing.py:205-223)
                draw = self.uniform_source if self.uniform_source is not None else (lambda shape: torch.rand(shape, dtype=torch.float32, device=dev))
                u = torch.stack([draw([R, 1]).to(dev, torch.float32) for _ in range(opt.ray_depth * 6)])
                u = u.reshape(opt.ray_depth, 6, R).contiguous()
                p.replay_u, p.rng_mode = u.data_ptr(), _abi.RNG_REPLAY
            else:
                p.replay_u, p.rng_mode = None, _abi.RNG_NATIVE
            ev = self._scene_store().get('textures_ready')
            p.shade_wait_event = ev.cuda_event if ev is not None else None   # joined inside drp_render, after the first extend launch
            check(L.drp_render(rc.handle, C.byref(scene_struct), C.byref(p), accum.data_ptr(), stream), "drp_render")
        return accum

    def render_accumulators(self) -> torch.Tensor:
        """Run the fused wavefront for this process' share of the samples; returns the (H*W, 16) fp32 sums."""
        opt = self.options
        if opt.rng == 'torch' and opt.shard_world > 1:
            raise ValueError("rng='torch' (reference replay) is a single-process mode")
        if opt.shard_mode == 'tile' and opt.shard_world > 1:
            accum = self.new_accumulators()
            ids = torch.arange(opt.ray_spp, dtype=torch.int32, device=self.device)
            for tile in self.tiles()[opt.shard_rank::opt.shard_world]:
                self.render_samples(ids, accum, tile=tile)
            return accum
        if opt.shard_mode not in ('spp', 'tile'):
            raise ValueError("shard_mode must be 'spp' or 'tile'")
        ids = shard_sample_ids(opt.ray_spp, opt.shard_rank, opt.shard_world, self.device)
        H, W = self.camera.resolution()
        if H * W > MAX_PIXELS_PER_CALL:
            # one drp_render call covers at most 2^24 pixels (a launch batch holds at least one sample of every pixel of the call): larger frames
            # (8K and up) are rendered tile by tile into the same accumulator -- RNG streams and rows are keyed by the global pixel, so the
            # image does not depend on the tiling (the property tile sharding rests on)
            accum = self.new_accumulators()
            for tile in frame_tiles(H, W, 4096):
                self.render_samples(ids, accum, tile=tile)
            return accum
        return self.render_samples(ids)

    def exchange_accumulators(self, accum: torch.Tensor) -> torch.Tensor:
        """The path's one exchange step between ranks: the packed accumulators are summed -- all-reduce, or reduce to ``options.result_rank``.
        Tile shards have disjoint supports, so ``options.tile_collective='gather'`` moves only the owned tiles instead (same bits; measured
        slower than the in-fabric all-reduce on NVSwitch, which therefore stays the default: profiles/r2/c5_n8*.json)."""
        opt = self.options
        if opt.shard_world > 1 and opt.shard_mode == 'tile' and opt.tile_collective == 'gather' and opt.result_rank is None:
            H, W = self.camera.resolution()
            return gather_tile_accumulators(accum, H, W, opt.tile_size, opt.shard_rank, opt.shard_world)
        return reduce_accumulators(accum, opt.shard_world, opt.result_rank)

    def finalize(self, accum: torch.Tensor):
        """Epilogue of trace_rays (path_tracing.py:348-352): /spp, saturate(alpha), flipud -- one kernel."""
        H, W = self.camera.resolution()
        dev = accum.device
        outs = {k: torch.empty([H, W, 3], dtype=torch.float32, device=dev) for k in
                ('radiance', 'albedo', 'emission', 'world_normal', 'world_position')}
        alpha = torch.empty([H, W, 1], dtype=torch.float32, device=dev)
        check(lib().drp_finalize(accum.data_ptr(), H, W, self.options.ray_spp, outs['radiance'].data_ptr(), alpha.data_ptr(),
                                 outs['albedo'].data_ptr(), outs['emission'].data_ptr(), outs['world_normal'].data_ptr(),
                                 outs['world_position'].data_ptr(), _stream_ptr(dev)), "drp_finalize")
        radiance = outs.pop('radiance')
        return radiance, alpha, outs

    def render_stats(self) -> dict:
        st = _abi.RenderStats()
        check(lib().drp_render_stats(self.raycaster().handle, C.byref(st)), "drp_render_stats")  # also fails on the handle's sticky device-error flag
        return dict(rays_traced=st.rays_traced, rays_nominal=st.rays_nominal, kernel_launches=st.kernel_launches)

    def set_profiling(self, enable: bool = True):
        check(lib().drp_set_profiling(self.raycaster().handle, int(enable)), "drp_set_profiling")

    def get_profile(self) -> dict:
        pr = _abi.Profile()
        check(lib().drp_get_profile(self.raycaster().handle, C.byref(pr)), "drp_get_profile")
        return {k: getattr(pr, k) for k, _ in _abi.Profile._fields_}

    def _wants_grad(self) -> bool:
        """True when autograd is on and some scene tensor (geometry, vertex attribute, material factor / texture, light) requires grad."""
        if not torch.is_grad_enabled():
            return False

        def g(x):
            return isinstance(x, torch.Tensor) and x.requires_grad
        for o in self.scene.objects:
            if any(g(x) for x in (o.verts, o.normals, o.M, o.color, o.uv, o.tangents)) or any(g(v) for v in (o.custom_attrs or {}).values()):
                return True
            m = o.material
            if any(g(v) for v in vars(m).values()) or any(g(getattr(v, 'image', None)) for v in vars(m).values()):
                return True
        return any(g(getattr(l, 'image', None)) or g(l.color) or g(l.intensity) for l in self.scene.lights)

    def pbr(self):
        """
        Path-traced PBR rendering; returns ``(radiance (H,W,3), alpha (H,W,1), extras)`` with extras ``albedo``,
        ``emission``, ``world_normal`` and ``world_position`` (H,W,3) -- same contract as path_tracing.py:354-367.

        Differentiability: the fused kernels have no backward pass.  Like the reference -- whose ``trace_rays`` is differentiable through the
        material / sampler torch code and detached only through ``raycaster.query`` -- a scene with tensors that require grad is rendered on
        the generic path (CUDA intersection + the per-bounce torch algebra of ``generic.py``), so gradients flow to materials, vertex
        attributes and lights; everything else takes the fused path.
        """
        if self._fused_scene() is None or self._wants_grad():
            if self.options.shard_world > 1:
                raise ValueError("sharded rendering needs the fused path (built-in materials, no tensors that require grad)")
            return self.trace_rays(self.sampler_brdf)
        with torch.no_grad():
            accum = self.exchange_accumulators(self.render_accumulators())
            self.raycaster().check_status()
            opt = self.options
            if opt.shard_world > 1 and opt.result_rank is not None and opt.shard_rank != opt.result_rank:
                return None   # the frame exists on options.result_rank only
            return self.finalize(accum)

    @torch.no_grad()
    def pbr_image(self, tone='agx', lut: torch.Tensor = None):
        """
        ``pbr()`` followed by the documented display chain -- ``to_pil(cat([agx_base_contrast(radiance), alpha]))``
        (tone_mapping.py:21-35, exchange.py:7-18) -- with normalisation, flipud, tone mapping, sRGB and quantisation fused into one pass
        over the reduced accumulator (``drp_tonemap``): returns a (H, W, 4) uint8 CUDA tensor, 4 B/pixel to read back instead of 64.
        ``tone``: 'agx' (needs ``lut``, see diffrp_b200.tonemap), 'srgb' or 'linear'.
        """
        from .tonemap import tonemap
        if self._fused_scene() is None:
            radiance, alpha, _ = self.trace_rays(self.sampler_brdf)
            return tonemap(torch.cat([radiance, alpha], -1), tone, lut=lut, alpha_offset=3)[1]
        accum = self.exchange_accumulators(self.render_accumulators())
        self.raycaster().check_status()
        opt = self.options
        if opt.shard_world > 1 and opt.result_rank is not None and opt.shard_rank != opt.result_rank:
            return None   # the frame exists on options.result_rank only
        H, W = self.camera.resolution()
        return tonemap(accum.view(H, W, _abi.ACCUM_CHANNELS), tone, lut=lut, scale=1.0 / self.options.ray_spp, alpha_offset=3, flip_rows=True)[1]

    def surface_attributes(self, rays_o: torch.Tensor, rays_d: torch.Tensor, t: torch.Tensor, i: torch.Tensor) -> torch.Tensor:
        """
        For custom samplers (``trace_rays(sampler)``): the whole material layer of the built-in sampler -- ``layer_material_rays`` + collectors
        (path_tracing.py:158-187) -- as ONE kernel over the ray batch.  Returns (R, 12) ``[albedo3 | normal3 | metallic | smoothness | alpha |
        emission3]`` (world-space shading normal), zeros for rays with ``t >= far``.  Needs Default / GLTF materials (raises otherwise).
        """
        fused = self._fused_scene()
        if fused is None:
            raise RuntimeError("scene contains custom Python materials: use layer_material_rays() (generic path)")
        n = rays_o.shape[0]
        attrs = torch.empty([n, 12], dtype=torch.float32, device=self.device)
        f32 = lambda x: x.to(self.device, torch.float32).contiguous()  # noqa: E731
        o, d, tt, ii = f32(rays_o), f32(rays_d), f32(t), i.to(self.device, torch.int32).contiguous()
        self._wait_textures()
        check(lib().drp_surface_attrs(self.raycaster().handle, C.byref(fused[0]), o.data_ptr(), d.data_ptr(), tt.data_ptr(), ii.data_ptr(),
                                      self.camera_far(), n, attrs.data_ptr(), _stream_ptr(self.device)), "drp_surface_attrs")
        return attrs

    # ---- generic path (user samplers / Python materials): see diffrp_b200/generic.py -------------------------------
    def layer_material_rays(self, rays_o, rays_d, t, i):
        from . import generic
        return generic.layer_material_rays(self, rays_o, rays_d, t, i)

    def sampler_brdf(self, rays_o, rays_d, t, i, d: int) -> RayOutputs:
        from . import generic
        return generic.sampler_brdf(self, rays_o, rays_d, t, i, d)

    def trace_rays(self, sampler: Callable, radiance_channels: int = 3, compact: bool = False):
        """Differentiable through the sampler / material torch code; only ``raycaster.query`` is detached (as in the reference).
        ``compact=True`` (extension): the sampler is only called with the rays that can still contribute (see ``generic.trace_rays``)."""
        from . import generic
        return generic.trace_rays(self, sampler, radiance_channels, compact)

    # the helper names the reference's docstring points sampler authors to (path_tracing.py:296, mixin.py:115-155)
    def _gbuffer_collect_layer_impl_stencil_masked(self, mats, operator, initial):
        from . import generic
        return generic.collect_gbuffer(mats, operator, initial)

    def _super_collector(self, si, so):
        from . import generic
        return generic.surface_row(si, so)

    def _collector_world_normal(self, si, so):
        from . import generic
        return generic.world_normal_of(si, so)
