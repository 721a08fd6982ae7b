#!/usr/bin/env python
"""
Install the UNMODIFIED reference into ``baseline/_ref`` (git-ignored, NOT gpurun-ignored: it travels to the GPU box), so that the
reference legs of bench.py, the drop-in tests and the GPU parity tests run the reference's own code there.

    python baseline/make_ref.py

Recipe (the offline install the task allows): copy /root/reference to a scratch dir (the tree is read-only and setuptools writes
build/ + egg-info next to setup.py), then
    python -m pip install --no-index --no-build-isolation --find-links /opt/wheelhouse --no-deps --target baseline/_ref <copy>
``--no-deps``: trimesh, calibur, torch_redstone, pyexr, diffrp-nvdiffrast are not in the wheelhouse; baseline/ref_loader.py supplies
stand-ins for the few symbols the path-tracing path needs and documents the two in-memory patches.  If pip fails, the package directory
is copied instead.  Writes baseline/_ref/INSTALL.json with what happened.
"""
import json
import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference"
DST = os.path.join(HERE, "_ref")


def main():
    if not os.path.isdir(os.path.join(SRC, "diffrp")):
        if os.path.isdir(os.path.join(DST, "diffrp")):
            print("reference tree absent (GPU box?); using the shipped", DST)
            return 0
        print("reference tree absent and nothing installed", file=sys.stderr)
        return 1
    shutil.rmtree(DST, ignore_errors=True)
    os.makedirs(DST)
    tmp = tempfile.mkdtemp(prefix="diffrp_src_")
    copy = os.path.join(tmp, "reference")
    shutil.copytree(SRC, copy, ignore=shutil.ignore_patterns(".git", "__pycache__", "docs"))
    cmd = [sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation", "--find-links", "/opt/wheelhouse", "--no-deps",
           "--target", DST, copy]
    res = subprocess.run(cmd, capture_output=True, text=True)
    how = "pip install --no-deps --target (unmodified wheel built from /root/reference)"
    if res.returncode != 0 or not os.path.isdir(os.path.join(DST, "diffrp")):
        how = "copytree of /root/reference/diffrp (pip failed: %s)" % (res.stderr.strip().splitlines() or ["?"])[-1]
        shutil.copytree(os.path.join(SRC, "diffrp"), os.path.join(DST, "diffrp"), ignore=shutil.ignore_patterns("__pycache__"))
    shutil.rmtree(tmp, ignore_errors=True)
    version = {}
    exec(open(os.path.join(DST, "diffrp", "version.py")).read(), version)
    # prove the files are unmodified: same bytes as the reference tree
    same, total = 0, 0
    for base, _, files in os.walk(os.path.join(SRC, "diffrp")):
        for f in files:
            if f.endswith(".py"):
                total += 1
                rel = os.path.relpath(os.path.join(base, f), SRC)
                p = os.path.join(DST, rel)
                same += int(os.path.exists(p) and open(p, "rb").read() == open(os.path.join(base, f), "rb").read())
    info = {"how": how, "version": version.get("__version__"), "python_files": total, "byte_identical_to_reference": same}
    json.dump(info, open(os.path.join(DST, "INSTALL.json"), "w"), indent=1)
    print(json.dumps(info))
    return 0 if same == total else 2


if __name__ == "__main__":
    sys.exit(main())
