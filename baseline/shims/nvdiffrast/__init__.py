"""Shim: import-only (rendering/interpolator.py:12, surface_deferred.py:4); the rasterizer is not on the path."""
