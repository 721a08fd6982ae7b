"""
Build ``libdiffrp_b200.so`` in-tree with nvcc for sm_100a (same compile-then-ctypes idiom as the reference's
mikktspace plugin, diffrp/plugins/mikktspace/__init__.py:11-26, with nvcc instead of gcc).
"""
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_PATH = os.path.join(HERE, "libdiffrp_b200.so")
SOURCES = ["api.cu", "wavefront.cu", "flatten.cu", "epilogue.cu", "conv3x3.cu"]
NVCC_FLAGS = [
    "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "--extended-lambda",
    "-Xcompiler", "-fPIC", "-shared", "-Xptxas", "-v",
]


def find_nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libdiffrp_b200.so cannot be built (there is no CPU fallback)")


def needs_build() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh", ".h"))]
    deps.append(os.path.join(HERE, "..", "include", "diffrp_b200.h"))
    return any(os.path.getmtime(d) > t for d in deps)


def build_library(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB_PATH
    extra = os.environ.get("DRP_NVCC_EXTRA", "").split()  # e.g. "-DDRP_SHADE_MINBLOCKS=8" for A/B experiments
    cmd = [find_nvcc()] + NVCC_FLAGS + extra + ["-o", LIB_PATH] + [os.path.join(CSRC, s) for s in SOURCES]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    log = os.path.join(HERE, "csrc", "ptxas.log")
    with open(log, "w") as fo:
        fo.write(res.stderr)
    if verbose:
        print(res.stderr)
    return LIB_PATH


if __name__ == "__main__":
    print(build_library(force=True, verbose=True))
