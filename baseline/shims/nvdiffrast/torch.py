"""Shim: import-only.  nvdiffrast is used by the reference's rasterizer (rendering/surface_deferred.py), which is not on the path-tracing
path; the two context classes exist because surface_deferred.py:209 names them in an annotation evaluated at import."""


class _Unavailable:
    def __init__(self, *a, **k):
        raise RuntimeError("nvdiffrast is not installed in this image (shim): the rasterizer is unavailable")


class RasterizeCudaContext(_Unavailable):
    pass


class RasterizeGLContext(_Unavailable):
    pass
