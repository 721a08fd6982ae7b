"""
Timing legs that run the UNMODIFIED reference (``baseline/_ref``, imported by baseline/ref_loader.py) on bench.py's workload.

MEASUREMENT INFRASTRUCTURE ONLY: called by ``bench.py --impl reference``, by bench.py's ``cpu_baseline`` / ``gpu_torch_baseline`` legs and
by tests.  Nothing under ``diffrp_b200/`` imports it; none of the repo's kernels run inside these functions.

What is timed is the reference's own public API and stock code path: ``diffrp.PathTracingSession(scene, camera, options).pbr()`` with
``raycaster_impl='naive-pbbvh'`` (the torch-level intersection the reference falls back to without torchoptix, utils/raycaster.py:120-260,
selected at rendering/path_tracing.py:142-156).  A *step* is a bounded sample of bench.py's configs[2] workload: a central window of the
1024^2 frame (``window_camera``: the reference's ``RawCamera`` with a cropped projection) x ``spp`` samples x 4 bounces over the same
2,097,152-triangle scene with the same textures.  Sessions are single-use in the reference; the BVH is built once (reported, excluded
-- like ``value`` on the GPU arm) and handed to the following sessions through the reference's own ``@cached`` store
(utils/cache.py:13-27, the mechanism rendering/surface_deferred.py:666-669 uses to inject values).
"""
import os
import time

import torch

from . import ref_loader, ref_scene

RES, DEPTH = 1024, 4


def reference_available() -> bool:
    return ref_loader.reference_root() is not None


def host_threads() -> int:
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


class ReferenceWorkload:
    """bench.py's scene inside the reference, on ``device`` ('cpu' or 'cuda'), with a window camera."""

    def __init__(self, device: str, tex: int, window: int, spp: int, builder: str = 'splitaxis'):
        from diffrp_b200 import synthetic as syn
        self.device, self.window, self.spp = device, window, spp
        if device == 'cpu':
            torch.set_num_threads(host_threads())  # torchrun exports OMP_NUM_THREADS=1 for its workers
        self.diffrp = ref_loader.load_reference(device)
        scene_host, camkw = syn.teaser_scene('cpu', tex=tex)
        self.scene = ref_scene.to_reference_scene(self.diffrp, scene_host, device)
        self._full = self.diffrp.PerspectiveCamera.from_orbit(h=RES, w=RES, **camkw)
        self.set_window(window)
        self.options = self.diffrp.PathTracingSessionOptions(ray_spp=spp, ray_depth=DEPTH, raycaster_impl='naive-pbbvh', raycaster_builder=builder)
        t0 = time.perf_counter()
        first = self.diffrp.PathTracingSession(self.scene, self.camera, self.options)
        first.raycaster()
        self._sync()
        self.build_s = time.perf_counter() - t0
        self.n_tris = int(first.vertex_array_object().tris.shape[0])
        # what later sessions adopt: flattened scene + BVH (keys = the @cached qualnames, utils/cache.py:13-27)
        self.shared = {k: v for k, v in first._cache.items() if k.endswith(('.raycaster', '.vertex_array_object'))}
        assert len(self.shared) == 2, sorted(first._cache)

    def set_window(self, window: int):
        """Central ``window`` x ``window`` crop of the 1024^2 frame (same camera, cropped projection)."""
        self.window = window
        lo = RES // 2 - window // 2
        self.camera = ref_scene.window_camera(self.diffrp, self._full.V(), self._full.P(), RES, RES, lo, lo, window, window)

    def _sync(self):
        if self.device != 'cpu':
            torch.cuda.synchronize()

    def step(self, seed: int):
        """One session over the window: returns (seconds, nominal ray-bounces)."""
        torch.manual_seed(seed)
        sess = self.diffrp.PathTracingSession(self.scene, self.camera, self.options)
        sess._cache = dict(self.shared)
        self._sync()
        t0 = time.perf_counter()
        radiance, alpha, extras = sess.pbr()
        self._sync()
        dt = time.perf_counter() - t0
        assert bool(torch.isfinite(radiance).all())
        return dt, self.window * self.window * self.spp * DEPTH

    def describe(self) -> str:
        return ("unmodified diffrp %s: PathTracingSession(raycaster_impl='naive-pbbvh').pbr() on %s, %dx%d central window of the %dx%d frame x %d spp x "
                "%d bounces per step (%d ray-bounces), %d-triangle scene; BVH build %.1f s excluded"
                % (getattr(self.diffrp, '__version__', ''), self.device, self.window, self.window, RES, RES, self.spp, DEPTH,
                   self.window * self.window * self.spp * DEPTH, self.n_tris, self.build_s))


def run(device: str, tex: int, window: int, spp: int, steps: int, warmup: int, budget_s: float = None):
    """-> dict(value Mrays/s, ms_per_step, cores, sample, kind='reference').
    ``budget_s``: bound on the timed region.  The first (untimed) step is clocked; if ``steps`` such steps would exceed the budget the window is
    shrunk (area ~ budget / projected time, with a margin because the reference's throughput falls with the batch size) -- every step stays
    a sample of the same workload, and the sample actually used is what ``sample`` describes."""
    wl = ReferenceWorkload(device, tex, window, spp)
    shrunk = ""
    if budget_s is not None and steps > 0:
        dt, _ = wl.step(999)
        if dt * steps > budget_s:
            w2 = max(32, int(window * (0.6 * budget_s / (dt * steps)) ** 0.5) // 8 * 8)
            if w2 < window:
                shrunk = " [window shrunk from %d to keep %d steps within %.0f s: one %d^2 step took %.1f s]" % (window, steps, budget_s, window, dt)
                wl.set_window(w2)
    for k in range(warmup):
        wl.step(1000 + k)
    tot_t, tot_n = 0.0, 0
    for k in range(steps):
        dt, n = wl.step(k)
        tot_t += dt
        tot_n += n
    return dict(value=tot_n / tot_t / 1e6, ms_per_step=tot_t / max(1, steps) * 1e3, cores=host_threads() if device == 'cpu' else None,
                kind="reference", sample=wl.describe() + shrunk, n_tris=wl.n_tris, seconds=tot_t)
