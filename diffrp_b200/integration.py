"""
Glue for using the B200 kernels from inside an existing diffrp installation, without editing diffrp
(see INTEGRATION.md).  Nothing here is needed when ``diffrp_b200`` is used standalone.

    import diffrp, diffrp_b200.integration as b200
    b200.install(diffrp)                       # adds raycaster_impl='b200' to diffrp.PathTracingSession
    opts = diffrp.PathTracingSessionOptions(raycaster_impl='b200')
"""
from .raycaster import B200Raycaster


def make_raycaster_class(diffrp_module):
    """A subclass of *diffrp's own* ``Raycaster`` ABC (diffrp/utils/raycaster.py:13-24) backed by libdiffrp_b200.so."""
    from importlib import import_module
    base = import_module(diffrp_module.__name__ + ".utils.raycaster").Raycaster

    class DiffrpB200Raycaster(base):
        def build(self, verts, tris, config):
            self._impl = B200Raycaster(verts, tris, dict(config))

        def query(self, rays_o, rays_d, far):
            return self._impl.query(rays_o, rays_d, far)  # (t fp32, i int32); t == far on a miss

    return DiffrpB200Raycaster


def install(diffrp_module):
    """Teach ``diffrp.PathTracingSession.raycaster()`` the value ``raycaster_impl='b200'`` (path_tracing.py:142-156)."""
    from importlib import import_module
    pt = import_module(diffrp_module.__name__ + ".rendering.path_tracing")
    cls = make_raycaster_class(diffrp_module)
    session = pt.PathTracingSession
    if getattr(session, "_b200_installed", False):
        return cls
    original = session.raycaster
    key = "PathTracingSession.raycaster"  # the @cached key (utils/cache.py:13-27)

    def raycaster(self):
        if self.options.raycaster_impl != 'b200':
            return original(self)
        cache = self.__dict__.setdefault('_cache', {})
        if key not in cache:
            vao = self.vertex_array_object()
            cache[key] = cls(vao.world_pos, vao.tris, {'epsilon': self.options.raycaster_epsilon})
        return cache[key]

    session.raycaster = raycaster
    session._b200_installed = True
    return cls


def agx_lut(diffrp_module, variant: str = "base-contrast"):
    """The AgX LUT of an installed diffrp, as its own loader returns it (tone_mapping.py:9-18: fliplr'd, z y x 3, on the GPU) --
    the ``lut`` argument of ``diffrp_b200.tonemap.agx_base_contrast`` / ``PathTracingSession.pbr_image``."""
    from importlib import import_module
    return import_module(diffrp_module.__name__ + ".utils.tone_mapping").agx_lut_loader.load(variant)
