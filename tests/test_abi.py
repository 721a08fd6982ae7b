"""CPU checks of the C-ABI boundary: the library builds, loads, and exports every symbol include/diffrp_b200.h declares."""
import ctypes
import os
import re

import pytest

from diffrp_b200 import _abi
from diffrp_b200._lib import lib, DiffrpB200Error, check

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "diffrp_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(drp_[a-z0-9_]+)\s*\(", src)))


def test_header_and_binding_agree():
    assert declared_symbols() == sorted(_abi.EXPORTED_SYMBOLS)


def test_library_loads_and_exports_every_declared_symbol():
    L = lib()
    assert L.drp_abi_version() == _abi.ABI_VERSION
    for name in declared_symbols():
        assert getattr(L, name) is not None


def test_struct_layouts_match_the_header():
    # sizes computed from the header's field lists (LP64)
    assert ctypes.sizeof(_abi.Texture) == 32
    assert ctypes.sizeof(_abi.Material) == 16 + 3 * 16 + 16 + 4 * 32 + 8
    assert ctypes.sizeof(_abi.Scene) == 9 * 8 + 16 + 8 + 32
    assert ctypes.sizeof(_abi.RenderParams) == 32 + 16 + 16 + 16 + 64 + 8 + 7 * 8
    assert ctypes.sizeof(_abi.RenderStats) == 24


def test_errors_are_reported_not_swallowed():
    L = lib()
    assert L.drp_release(123456789) != 0
    assert b"unknown handle" in L.drp_last_error()
    with pytest.raises(DiffrpB200Error):
        check(L.drp_set_epsilon(987654321, 0.0), "drp_set_epsilon")


def test_product_has_no_cpu_fallback():
    import torch
    import diffrp_b200 as drp
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(RuntimeError):
        drp.PathTracingSession(drp.Scene(), drp.PerspectiveCamera(h=8, w=8))
    with pytest.raises((ValueError, RuntimeError)):
        drp.B200Raycaster(torch.zeros(3, 3), torch.zeros(1, 3, dtype=torch.int32))
    # the product package never imports the oracle
    for root, _, files in os.walk(os.path.join(ROOT, "diffrp_b200")):
        for f in files:
            if f.endswith(".py"):
                assert "oracle" not in open(os.path.join(root, f)).read().replace("oracle/", ""), f


def test_texel_records_layout_and_eligibility():
    """Host side of drp_material_t.texel_records: channel order of the 48-byte texel, and the cases that must fall back to four textures."""
    import torch
    import scenes
    from diffrp_b200.flatten import material_descriptions, pad_rgba, texel_records
    objs = scenes.mixed_scene().objects
    descs = material_descriptions(objs, 'cpu', rgba=True)
    with_rec = [d for d in descs if 'texel_records' in d]
    assert len(with_rec) == 2  # the OPAQUE and the BLEND sphere (no emissive factor, but all four textures); MASK has no normal map
    assert [d.get('emissive_factor') is None for d in with_rec] == [False, True]
    d = with_rec[0]
    rec = d['texel_records']
    b, m, n, e = (d[k]['image'] for k in ('base_color_tex', 'mr_tex', 'normal_tex', 'emissive_tex'))
    H, W = b.shape[:2]
    L = d['texel_tile_log2']
    T = 1 << L
    assert L == 2 and H % T == 0 and W % T == 0                      # tile-major: (H/T, W/T, T, T, 12)
    assert rec.shape == (H // T, W // T, T, T, 12) and rec.dtype == torch.float32 and rec.is_contiguous()
    flat = rec.permute(0, 2, 1, 3, 4).reshape(H, W, 12)              # back to row-major
    assert torch.equal(flat[..., 0:4], b) and torch.equal(flat[..., 4:6], m[..., 1:3])
    assert torch.equal(flat[..., 6:9], n[..., :3]) and torch.equal(flat[..., 9:12], e[..., :3])
    # the kernel's tap index (csrc/shade.cuh: tex_taps): texel (y, x) sits at ((y>>L)*(W>>L) + (x>>L)) << 2L | (y & (T-1)) << L | (x & (T-1))
    ys, xs = torch.meshgrid(torch.arange(H), torch.arange(W), indexing='ij')
    idx = (((ys >> L) * (W >> L) + (xs >> L)) << (2 * L)) | ((ys & (T - 1)) << L) | (xs & (T - 1))
    assert torch.equal(rec.reshape(-1, 12)[idx.reshape(-1)].reshape(H, W, 12), flat)
    # packing into the C struct keeps the pointer and checks the shape
    keep = []
    mat = _abi.pack_material(d, lambda t: t.data_ptr(), keep)
    assert mat.texel_records == rec.data_ptr() and any(k is rec for k in keep) and mat.texel_tile_log2 == 2
    # fall-backs: a texture of another size, another wrap mode, a missing texture
    other = dict(d, mr_tex=dict(d['mr_tex'], image=pad_rgba(torch.rand(8, 8, 3))))
    assert texel_records(other) is None
    assert texel_records(dict(d, normal_tex=dict(d['normal_tex'], wrap='clamp'))) is None
    assert texel_records(dict(d, emissive_tex=None)) is None
    assert texel_records(dict(kind='default', tint=None)) is None


@pytest.mark.gpu
def test_library_builds_from_source_on_this_box_and_traces(tmp_path):
    """The shipped .so is what the other tests load; this one proves the SOURCES build where they run: nvcc compiles csrc/ for sm_100a into a
    scratch directory, a fresh interpreter loads that library (DIFFRP_B200_LIB) and reproduces the oracle's hits bit for bit."""
    import subprocess
    import sys
    from diffrp_b200 import build as b
    out = str(tmp_path / "libdiffrp_b200_fresh.so")
    cmd = [b.find_nvcc()] + [f for f in b.NVCC_FLAGS if f not in ("-Xptxas", "-v")] + ["-o", out] + [os.path.join(b.CSRC, s) for s in b.SOURCES]
    res = subprocess.run(cmd, capture_output=True, text=True)
    assert res.returncode == 0 and os.path.getsize(out) > 1_000_000, res.stderr[-2000:]
    code = (
        "import numpy as np, torch, oracle\n"
        "from diffrp_b200 import synthetic as syn, _lib\n"
        "from diffrp_b200.raycaster import B200Raycaster\n"
        "v, f = syn.icosphere(3, 0.8); o, d = syn.random_rays(50000, seed=4)\n"
        "rc = B200Raycaster(torch.from_numpy(v).cuda(), torch.from_numpy(f).cuda(), {'epsilon': 1e-8})\n"
        "t, i = rc.query(torch.from_numpy(o).cuda(), torch.from_numpy(d).cuda(), 10.0)\n"
        "ot, oi = oracle.bruteforce(v, f, o, d, 10.0, 1e-8)\n"
        "assert np.array_equal(t.cpu().numpy().view(np.int32), ot.view(np.int32)) and np.array_equal(i.cpu().numpy(), oi)\n"
        "print('FRESH_OK', _lib.build_config())\n")
    env = dict(os.environ, DIFFRP_B200_LIB=out, DIFFRP_B200_NO_BUILD="1")
    res = subprocess.run([sys.executable, "-c", code], cwd=ROOT, env=env, capture_output=True, text=True)
    assert res.returncode == 0 and "FRESH_OK" in res.stdout, res.stderr[-2000:]


def test_every_intra_package_import_resolves():
    """`from .module import name` statements inside functions only fail when that code path runs -- on the GPU box.  Resolve all of them here."""
    import ast
    import importlib
    pkg = os.path.join(ROOT, "diffrp_b200")
    missing = []
    for fn in sorted(os.listdir(pkg)):
        if not fn.endswith(".py"):
            continue
        tree = ast.parse(open(os.path.join(pkg, fn)).read())
        for node in ast.walk(tree):
            if isinstance(node, ast.ImportFrom) and node.level == 1:
                mod = importlib.import_module("diffrp_b200" + ("." + node.module if node.module else ""))
                for alias in node.names:
                    if alias.name != "*" and not hasattr(mod, alias.name):
                        try:
                            importlib.import_module(mod.__name__ + "." + alias.name)
                        except ImportError:
                            missing.append((fn, node.module, alias.name))
    assert not missing, missing


def test_pinned_arena_scene_uploads_in_a_few_large_copies():
    """Scene.pin_memory() lays the scene out in ONE arena in upload order, so PackedUpload's copy list collapses: one run for the geometry,
    one for the textures (layout logic only: pageable arena, no GPU)."""
    import sys
    import torch
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import scenes
    from diffrp_b200.flatten import PackedUpload, merge_runs
    src = scenes.mixed_scene()
    sc = src.pin_memory(pin=False)
    arena = (sc._arena.data_ptr(), sc._arena.data_ptr() + sc._arena.numel())
    for a, b in zip(src.objects, sc.objects):
        for f in ('verts', 'normals', 'color', 'uv', 'tangents', 'tris'):
            assert torch.equal(getattr(a, f), getattr(b, f)) and arena[0] <= getattr(b, f).data_ptr() < arena[1]
        assert b.material is not a.material or not hasattr(a.material, 'base_color_texture')
    assert torch.equal(src.lights[0].image, sc.lights[0].image)
    items, off = [], 0
    for o in sc.objects:            # the registration order of flatten_scene_cuda
        for t in (o.verts, o.normals, o.color, o.uv, o.tangents, o.tris):
            nbytes = t.numel() * t.element_size()
            items.append((off, t.data_ptr(), nbytes))
            off += -(-nbytes // PackedUpload.ALIGN) * PackedUpload.ALIGN
    runs = merge_runs(items, arena)
    assert len(items) == 6 * len(sc.objects) and len(runs) == 1 and runs[0][2] == items[-1][0] + items[-1][2]
    assert len(merge_runs(items, None)) == len(items)                  # unrelated host tensors are never merged
    shuffled = [items[1], items[0]] + items[2:]
    assert len(merge_runs([(o, p, n) for (o, _, _), (_, p, n) in zip(items, shuffled)], arena)) > 1


def test_instance_table_of_a_scene_and_shared_tensors_survive_moves():
    """Host logic of the instanced build: objects sharing vertex + index tensors form the instance table (first_tri, mesh); Scene.to() and
    Scene.pin_memory() keep shared tensors shared, so the table is the same after a move."""
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from diffrp_b200 import synthetic as syn
    from diffrp_b200.path_tracing import scene_instances
    sc, _ = syn.instanced_scene('cpu', n_instances=16, mesh_res=(6, 4), env_res=(8, 16))
    first, mesh = scene_instances(sc.objects)
    assert mesh.tolist() == [0] * 16 and first.tolist() == [48 * k for k in range(17)]
    for moved in (sc.to('cpu'), sc.pin_memory(pin=False)):
        f2, m2 = scene_instances(moved.objects)
        assert m2.tolist() == mesh.tolist() and f2.tolist() == first.tolist()
        assert len({o.verts.data_ptr() for o in moved.objects}) == 1
    few, _ = syn.instanced_scene('cpu', n_instances=4, mesh_res=(6, 4), env_res=(8, 16))
    assert scene_instances(few.objects) is None                       # too few objects to pay
    import scenes
    assert scene_instances(scenes.mixed_scene().objects) is None      # nothing shared


def test_frames_beyond_one_call_are_tiled_within_the_per_call_pixel_limit():
    """drp_render covers at most 2^24 pixels per call; render_accumulators() tiles larger frames (8K: 33 M pixels).  The tiles must partition
    the frame, and each must fit one call."""
    import numpy as np
    from diffrp_b200.path_tracing import frame_tiles, MAX_PIXELS_PER_CALL
    for H, W in ((4320, 7680), (4096, 4097), (8192, 8192), (1, 1 << 25)):
        assert H * W > MAX_PIXELS_PER_CALL
        cover = np.zeros((H, W), np.uint8) if H * W <= (1 << 26) else None
        total = 0
        for (x0, y0, w, h) in frame_tiles(H, W, 4096):
            assert 0 < w * h <= MAX_PIXELS_PER_CALL and x0 + w <= W and y0 + h <= H
            total += w * h
            if cover is not None:
                cover[y0:y0 + h, x0:x0 + w] += 1
        assert total == H * W and (cover is None or (cover == 1).all())
