"""
Makes the Python reference importable for generating golden vectors (tests/golden/make_golden.py).

TEST INFRASTRUCTURE ONLY.  A thin front for ``baseline/ref_loader.py`` (one implementation of the import hook, the shims for absent
third-party modules and the two documented in-memory patches -- see its docstring): in the build container the reference is imported
straight from the read-only ``/root/reference``; elsewhere from ``baseline/_ref`` (``python baseline/make_ref.py``).
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from baseline import ref_loader  # noqa: E402

REFERENCE_ROOT = "/root/reference"


def load_reference(device: str = "cpu"):
    """The reference's ``diffrp`` module (patch A for 'cpu', patch B always)."""
    root = REFERENCE_ROOT if os.path.isdir(os.path.join(REFERENCE_ROOT, "diffrp")) else None
    return ref_loader.load_reference(device, root=root, patch_int32=True)
