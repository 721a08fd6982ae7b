"""Shim: import-only (resources/hdris.py:2); the bundled HDRI is not used by the benchmarks."""
