"""
The per-thread logic of the CUDA kernels (diffrp_b200/csrc/*.cuh), compiled for the host by tests/hostsim, against the
CPU oracle.  This is how kernel logic is debugged in the GPU-less build container; the GPU tests repeat the same
comparisons through the real library.
"""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import oracle
import scenes
import diffrp_b200 as drp
from diffrp_b200 import _abi, synthetic as syn
from test_oracle_golden import make_camera

HS_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "hostsim")


@pytest.fixture(scope="module")
def hs():
    subprocess.run(["make", "-C", HS_DIR], check=True, capture_output=True)
    L = C.CDLL(os.path.join(HS_DIR, "libhostsim.so"))
    vp = C.c_void_p
    L.hs_build.restype = vp
    L.hs_build.argtypes = [vp, vp, C.c_int64, C.c_int64]
    L.hs_free.argtypes = [vp]
    L.hs_trace.restype = C.c_int64
    L.hs_trace.argtypes = [vp, vp, vp, C.c_int64, C.c_float, C.c_float, vp, vp]
    L.hs_stats.argtypes = [vp, vp, vp]
    L.hs_render.restype = C.c_int64
    L.hs_render.argtypes = [vp, C.POINTER(_abi.Scene), C.POINTER(_abi.RenderParams), C.c_float, vp]
    return L


def hs_query(L, v, f, o, d, far=10.0, eps=1e-8):
    v, f = np.ascontiguousarray(v, np.float32), np.ascontiguousarray(f, np.int32)
    h = L.hs_build(v.ctypes.data, f.ctypes.data, len(v), len(f))
    st = np.zeros(5, np.int64)
    sah = C.c_float()
    L.hs_stats(h, st.ctypes.data, C.byref(sah))
    t, i = np.empty(len(o), np.float32), np.empty(len(o), np.int32)
    overflow = L.hs_trace(h, o.ctypes.data, d.ctypes.data, len(o), far, eps, t.ctypes.data, i.ctypes.data)
    L.hs_free(h)
    return t, i, st, overflow


@pytest.mark.parametrize("n_tris", [0, 1, 2, 3, 5, 1280])
def test_lbvh_and_traversal_equal_bruteforce(hs, n_tris):
    v, f = syn.icosphere(3, 0.8)
    f = f[:n_tris].copy()
    o, d = syn.random_rays(30_000, seed=n_tris)
    t, i, st, overflow = hs_query(hs, v, f, o, d)
    assert overflow == 0 and st[3] == n_tris and st[4] <= 4  # every triangle in exactly one leaf, leaves <= 4 tris
    if n_tris == 0:
        assert (t == np.float32(10.0)).all() and (i == 0).all()
        return
    ot, oi = oracle.bruteforce(v, f, o, d, 10.0, 1e-8)
    assert np.array_equal(t.view(np.int32), ot.view(np.int32)) and np.array_equal(i, oi)


def test_degenerate_and_duplicate_triangles(hs):
    v, f = syn.uv_sphere(64, 32)  # pole rows are zero-area triangles
    f = np.concatenate([f, f[:200]])  # exact duplicates: ties must resolve to the smaller id
    o, d = syn.random_rays(50_000, seed=4)
    t, i, st, overflow = hs_query(hs, v, f, o, d)
    ot, oi = oracle.bruteforce(v, f, o, d, 10.0, 1e-8)
    assert overflow == 0 and np.array_equal(t.view(np.int32), ot.view(np.int32)) and np.array_equal(i, oi)


def test_axis_aligned_rays_on_box_planes(hs):
    """Reference defect B9 (NaN slab test) must not be reproduced: axis-parallel rays starting on box planes still hit."""
    v = np.array([[0, 0, 1], [1, 0, 1], [1, 1, 1], [0, 1, 1]], np.float32)
    f = np.array([[0, 1, 2], [0, 2, 3]], np.int32)
    o = np.array([[0.0, 0.25, 2.0], [0.5, 0.0, 2.0], [0.25, 0.25, 2.0], [1.0, 0.5, 2.0]], np.float32)
    d = np.tile(np.array([[0, 0, -1]], np.float32), (4, 1))
    t, i, st, overflow = hs_query(hs, v, f, o, d)
    ot, oi = oracle.bruteforce(v, f, o, d, 10.0, 1e-8)
    assert np.array_equal(t, ot) and np.array_equal(i, oi)
    assert (t == 1.0).all()


@pytest.mark.parametrize("scene_name,last,records", [("ico", "void", False), ("mixed", "void", False), ("mixed", "skybox", True),
                                                     ("affine", "skybox", True)])
def test_shading_logic_matches_oracle(hs, scene_name, last, records):
    scene = {"ico": scenes.icosphere_scene, "mixed": scenes.mixed_scene, "affine": scenes.affine_instances_scene}[scene_name]()
    cam = make_camera(None, dict(h=40, w=56, radius=3.0, azim=25, elev=15, origin=[0.0, -0.1, 0.0], fov=32))
    vao, hscene, p, keep = scenes.oracle_inputs(scene, cam, 3, 3, last_bounce=last, seed=123)
    acc, n = oracle.render(oracle.BVH(vao.world_pos.numpy(), vao.tris.numpy()), hscene, p)
    wp, tr = hscene.arrays['world_pos'], hscene.arrays['tris']
    if records:  # the interleaved 64-byte vertex records the CUDA path prefers must give the same result as the five arrays
        a = hscene.arrays
        rec = np.ascontiguousarray(np.concatenate([a['world_pos'], a['world_nrm'], a['uv'], a['color'], a['world_tan']], 1), np.float32)
        assert rec.shape[1] == 16
        hscene.struct.vertex_records = rec.ctypes.data
    h = hs.hs_build(wp.ctypes.data, tr.ctypes.data, len(wp), len(tr))
    acc2 = np.zeros_like(acc)
    hs.hs_render(h, C.byref(hscene.struct), C.byref(p), 1e-8, acc2.ctypes.data)
    hs.hs_free(h)
    np.testing.assert_allclose(acc2, acc, rtol=2e-5, atol=2e-5)


def test_texel_records_match_separate_textures_on_the_host(hs):
    """drp_material_t.texel_records (interleaved 48-byte texels) through the same shade.cuh code the kernels compile:
    identical accumulators to the four separate textures."""
    import torch
    from diffrp_b200.flatten import material_descriptions, pad_rgba, texel_records
    scene = scenes.mixed_scene()
    cam = make_camera(None, dict(h=40, w=56, radius=3.0, azim=25, elev=15, origin=[0.0, -0.1, 0.0], fov=32))
    vao, hscene, p, keep = scenes.oracle_inputs(scene, cam, 2, 3, seed=7)
    wp, tr = hscene.arrays['world_pos'], hscene.arrays['tris']
    h = hs.hs_build(wp.ctypes.data, tr.ctypes.data, len(wp), len(tr))
    acc_a = np.zeros((40 * 56, 16), np.float32)
    hs.hs_render(h, C.byref(hscene.struct), C.byref(p), 1e-8, acc_a.ctypes.data)
    recs, n_rec = [], 0
    for k, d in enumerate(material_descriptions(scene.objects, 'cpu')):
        d = dict(d)
        for name in ('base_color_tex', 'mr_tex', 'normal_tex', 'emissive_tex'):
            if d.get(name) is not None:
                d[name] = dict(d[name], image=pad_rgba(d[name]['image']))
        rec = texel_records(d)
        if rec is not None:
            recs.append(rec)
            hscene.struct.materials[k].texel_records = rec.data_ptr()
            hscene.struct.materials[k].texel_tile_log2 = d['texel_tile_log2']   # tiled layout: the tap index is remapped in tex_taps
            assert d['texel_tile_log2'] == 2
            n_rec += 1
    assert n_rec >= 1  # the OPAQUE / repeat / linear sphere has all four textures
    acc_b = np.zeros_like(acc_a)
    hs.hs_render(h, C.byref(hscene.struct), C.byref(p), 1e-8, acc_b.ctypes.data)
    hs.hs_free(h)
    assert np.abs(acc_a).sum() > 0
    assert np.array_equal(acc_a, acc_b)


def test_tile_render_matches_oracle_tile(hs):
    scene = scenes.mixed_scene()
    cam = make_camera(None, dict(h=40, w=56, radius=3.0, azim=25, elev=15, origin=[0.0, -0.1, 0.0], fov=32))
    vao, hscene, p, keep = scenes.oracle_inputs(scene, cam, 2, 3, seed=5)
    bvh = oracle.BVH(vao.world_pos.numpy(), vao.tris.numpy())
    whole, _ = oracle.render(bvh, hscene, p)
    p.tile_x0, p.tile_y0, p.tile_w, p.tile_h = 16, 8, 24, 20
    tile, _ = oracle.render(bvh, hscene, p)
    img_w, img_t = whole.reshape(40, 56, 16), tile.reshape(40, 56, 16)
    np.testing.assert_allclose(img_t[8:28, 16:40], img_w[8:28, 16:40], rtol=1e-6, atol=1e-6)  # same pixels, same RNG keys
    mask = np.ones((40, 56), bool); mask[8:28, 16:40] = False
    assert (img_t[mask] == 0).all()
    wp, tr = hscene.arrays['world_pos'], hscene.arrays['tris']
    h = hs.hs_build(wp.ctypes.data, tr.ctypes.data, len(wp), len(tr))
    acc2 = np.zeros_like(tile)
    hs.hs_render(h, C.byref(hscene.struct), C.byref(p), 1e-8, acc2.ctypes.data)
    hs.hs_free(h)
    np.testing.assert_allclose(acc2, tile, rtol=2e-5, atol=2e-5)


def hs_query_wide(L, v, f, o, d, far, eps=1e-8):
    vp = C.c_void_p
    L.hs_build_wide.restype = vp
    L.hs_build_wide.argtypes = [vp, vp, C.c_int64, C.c_int64]
    L.hs_trace_wide.restype = C.c_int64
    L.hs_trace_wide.argtypes = [vp, vp, vp, C.c_int64, C.c_float, C.c_float, vp, vp]
    v, f = np.ascontiguousarray(v, np.float32), np.ascontiguousarray(f, np.int32)
    h = L.hs_build_wide(v.ctypes.data, f.ctypes.data, len(v), len(f))
    t, i = np.empty(len(o), np.float32), np.empty(len(o), np.int32)
    overflow = L.hs_trace_wide(h, o.ctypes.data, d.ctypes.data, len(o), far, eps, t.ctypes.data, i.ctypes.data)
    L.hs_free(h)
    return t, i, overflow


@pytest.mark.parametrize("scale,offset", [(1.0, 0.0), (1e-3, 0.0), (1e3, 0.0), (1.0, 100.0), (1.0, 1e4), (1e-2, 37.5)])
@pytest.mark.parametrize("seed", [0, 1])
def test_random_triangle_soups_all_layouts_equal_bruteforce(hs, scale, offset, seed):
    """Conservativeness of the padded / quantised slab tests under extreme coordinates: random soups (sliver, degenerate and
    duplicate triangles included) scaled by 1e-3..1e3 and translated up to 1e4 from the origin -- where fp32 spacing is ~1e-3
    and every slab distance cancels catastrophically -- must still reproduce the exhaustive search bit for bit."""
    rng = np.random.default_rng(seed)
    n = 600
    c = rng.uniform(-1, 1, (n, 1, 3))
    tri = c + rng.normal(0, 0.08, (n, 3, 3)) * rng.uniform(0.02, 1.0, (n, 1, 1))
    tri[:20, 2] = tri[:20, 1]                      # degenerate (zero area)
    tri[20:40] = tri[40:60]                        # exact duplicates
    tri[60:80, :, 1] = tri[60:80, :1, 1]           # axis-aligned (flat in y): zero-thickness boxes
    v = ((tri.reshape(-1, 3) * scale) + offset).astype(np.float32)
    f = np.arange(3 * n, dtype=np.int32).reshape(n, 3)
    m = 20000
    org = (rng.uniform(-3, 3, (m, 3)) * scale + offset).astype(np.float32)
    tgt = (rng.uniform(-1, 1, (m, 3)) * scale + offset).astype(np.float32)
    d = tgt - org
    d = (d / np.linalg.norm(d, axis=-1, keepdims=True)).astype(np.float32)
    d[:200, 0] = 0.0                               # axis-parallel components
    d[200:400, 1] = 0.0
    d[:400] /= np.maximum(np.linalg.norm(d[:400], axis=-1, keepdims=True), 1e-20)
    far = np.float32(20.0 * scale)
    ot, oi = oracle.bruteforce(v, f, org, d, float(far), 0.0)
    for name, (t, i, overflow) in (("binary", hs_query(hs, v, f, org, d, float(far), 0.0)[:2] + (0,)), ("wide", hs_query_wide(hs, v, f, org, d, float(far), 0.0))):
        assert overflow == 0
        bad = np.nonzero((t.view(np.int32) != ot.view(np.int32)) | (i != oi))[0]
        assert len(bad) == 0, "%s layout: %d rays differ from the exhaustive search (first: %s)" % (name, len(bad), bad[:5])


def test_tonemap_pixel_function_equals_oracle(hs):
    """csrc/tonemap.cuh (host build) vs oracle/orc_tonemap.c on the golden inputs: both use libm here, so floats are bit-identical."""
    g = dict(np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "tonemap.npz")))
    hs.hs_tonemap.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.POINTER(_abi.TonemapParams), C.c_void_p, C.c_void_p]
    rgba = np.ascontiguousarray(np.concatenate([g['rgb'], g['alpha']], -1))
    lut = np.ascontiguousarray(g['lut'])
    for tone, name in ((_abi.TONE_AGX, 'agx'), (_abi.TONE_SRGB, 'srgb'), (_abi.TONE_LINEAR, 'linear')):
        for alpha_offset in (3, -1):
            c = 4 if alpha_offset >= 0 else 3
            p = _abi.TonemapParams(tone, lut.shape[0], lut.ctypes.data, 4, alpha_offset, 0, 0.5)
            f, b = np.empty((len(rgba), c), np.float32), np.empty((len(rgba), c), np.uint8)
            hs.hs_tonemap(rgba.ctypes.data, 1, len(rgba), C.byref(p), b.ctypes.data, f.ctypes.data)
            fo, bo = oracle.tonemap(rgba, name, lut=lut, scale=0.5, alpha_offset=alpha_offset)
            assert np.array_equal(f.view(np.uint32), fo.view(np.uint32)), (name, alpha_offset, np.abs(f - fo).max())
            assert np.array_equal(b, bo)


def test_node_format2_tables_and_triangle_index(hs):
    """The two table expansions of the per-slot hit byte (cwbvh.cuh, node format 2) against their definitions: internal children move to
    bit (slot ^ octinv) -- so that the highest set bit is the child to visit first --, leaf slots spread to 3 triangle bits each; triangle bit
    j of a node is the popc(V below j)-th triangle after tri_base."""
    if not hasattr(hs, 'hs_cw_tables'):
        pytest.skip("host simulator built for node format 1")
    perm = np.zeros(8 * 256, np.uint8)
    spread = np.zeros(256, np.uint32)
    hs.hs_cw_tables.argtypes = [C.c_void_p, C.c_void_p]
    hs.hs_cw_tables(perm.ctypes.data, spread.ctypes.data)
    for octinv in range(8):
        for byte in range(256):
            want = 0
            for slot in range(8):
                if byte >> slot & 1:
                    want |= 1 << (slot ^ octinv)
            assert perm[octinv * 256 + byte] == want
    for byte in range(256):
        want = 0
        for slot in range(8):
            if byte >> slot & 1:
                want |= 7 << (3 * slot)
        assert spread[byte] == want
    hs.hs_cw_tri_index.restype = C.c_int
    hs.hs_cw_tri_index.argtypes = [C.c_uint32, C.c_uint32, C.c_int]
    rng = np.random.default_rng(0)
    for _ in range(200):
        counts = rng.integers(0, 4, 8)                      # triangles per slot (0 = empty or internal)
        vmask = 0
        for slot, c in enumerate(counts):
            vmask |= ((1 << int(c)) - 1) << (3 * slot)
        base = int(rng.integers(0, 1 << 20))
        k = 0
        for slot, c in enumerate(counts):                   # compact storage in slot order
            for t in range(int(c)):
                assert hs.hs_cw_tri_index(base, vmask, 3 * slot + t) == base + k
                k += 1


# ---- refit and instanced assembly, node-level code of api.cu run serially ------------------------------------------------
def _hs_wide_api(L):
    vp = C.c_void_p
    L.hs_build_wide.restype = vp
    L.hs_build_wide.argtypes = [vp, vp, C.c_int64, C.c_int64]
    L.hs_trace_wide.restype = C.c_int64
    L.hs_trace_wide.argtypes = [vp, vp, vp, C.c_int64, C.c_float, C.c_float, vp, vp]
    L.hs_refit_wide.restype = C.c_int
    L.hs_refit_wide.argtypes = [vp, vp, vp]
    L.hs_build_instanced.restype = vp
    L.hs_build_instanced.argtypes = [vp, vp, C.c_int64, vp, vp, C.c_int64]
    L.hs_free.argtypes = [vp]


def _hs_trace(L, h, o, d, far=10.0, eps=1e-8):
    t, i = np.empty(len(o), np.float32), np.empty(len(o), np.int32)
    overflow = L.hs_trace_wide(h, o.ctypes.data, d.ctypes.data, len(o), far, eps, t.ctypes.data, i.ctypes.data)
    assert overflow == 0
    return t, i


@pytest.mark.parametrize("amp", [0.02, 0.7])
def test_refit_keeps_closest_hits_exact(hs, amp):
    """drp_refit's per-node code (cw_refit_node) over the topology of the undeformed mesh, applied to a deformed one: hits equal the oracle's
    exhaustive search over the deformed mesh bit for bit, for a small and for a topology-scrambling deformation; then refit back."""
    import oracle
    from diffrp_b200 import synthetic as syn
    _hs_wide_api(hs)
    v, f = syn.uv_sphere(48, 24, radius=0.8, bump=0.05, noise=0.01, seed=0)
    v, f = np.ascontiguousarray(v, np.float32), np.ascontiguousarray(f, np.int32)
    o, d = syn.random_rays(6000, seed=3)
    ang = amp * v[:, 1:2] * 4.0
    v2 = np.concatenate([v[:, 0:1] * np.cos(ang) - v[:, 2:3] * np.sin(ang), v[:, 1:2] * (1 + amp), v[:, 0:1] * np.sin(ang) + v[:, 2:3] * np.cos(ang)], 1)
    v2 = np.ascontiguousarray(v2 + np.random.default_rng(1).normal(0, amp * 0.05, v2.shape), np.float32)
    h = hs.hs_build_wide(v.ctypes.data, f.ctypes.data, len(v), len(f))
    assert hs.hs_refit_wide(h, v2.ctypes.data, f.ctypes.data) == 0
    t, i = _hs_trace(hs, h, o, d)
    ot, oi = oracle.bruteforce(v2, f, o, d, 10.0, 1e-8)
    assert np.array_equal(t.view(np.int32), ot.view(np.int32)) and np.array_equal(i, oi)
    assert 0.2 < (t < 10.0).mean() < 0.95
    assert hs.hs_refit_wide(h, v.ctypes.data, f.ctypes.data) == 0
    t, i = _hs_trace(hs, h, o, d)
    ot, oi = oracle.bruteforce(v, f, o, d, 10.0, 1e-8)
    assert np.array_equal(t.view(np.int32), ot.view(np.int32)) and np.array_equal(i, oi)
    hs.hs_free(h)


@pytest.mark.parametrize("n_inst,two_meshes", [(1, False), (2, False), (9, False), (70, False), (12, True)])
def test_instanced_assembly_keeps_closest_hits_exact(hs, n_inst, two_meshes):
    """drp_build_instanced's logic (template per mesh, replicate + refit per instance, host-built instance level, copied roots) over
    rigidly transformed copies of one or two meshes: hits and GLOBAL primitive ids equal the exhaustive search over the flattened geometry."""
    import oracle
    from diffrp_b200 import synthetic as syn
    _hs_wide_api(hs)
    rng = np.random.default_rng(n_inst)
    va, fa = syn.uv_sphere(10, 6, radius=0.12, bump=0.01, noise=0.002, seed=1)
    vb, fb = syn.icosphere(1, 0.1)
    verts, tris, first, mesh = [], [], [0], []
    voff = 0
    for q in range(n_inst):
        use_b = two_meshes and q % 3 == 1
        v, f = (vb, fb) if use_b else (va, fa)
        M = syn.rigid_matrix(rng, 0.6 + rng.random(), rng.uniform(-0.6, 0.6, 3))
        verts.append(v @ M[:3, :3].T + M[:3, 3])
        tris.append(f + voff)
        voff += len(v)
        first.append(first[-1] + len(f))
        mesh.append(7 if use_b else 3)
    verts = np.ascontiguousarray(np.concatenate(verts), np.float32)
    tris = np.ascontiguousarray(np.concatenate(tris), np.int32)
    first, mesh = np.asarray(first, np.int64), np.asarray(mesh, np.int32)
    h = hs.hs_build_instanced(verts.ctypes.data, tris.ctypes.data, len(tris), first.ctypes.data, mesh.ctypes.data, n_inst)
    assert h
    o, d = syn.random_rays(5000, seed=8)
    o = np.ascontiguousarray(o * 0.8, np.float32)
    t, i = _hs_trace(hs, h, o, d)
    ot, oi = oracle.bruteforce(verts, tris, o, d, 10.0, 1e-8)
    assert np.array_equal(t.view(np.int32), ot.view(np.int32)) and np.array_equal(i, oi)
    assert (t < 10.0).mean() > (0.01 if n_inst < 3 else 0.1)
    if n_inst > 2:
        assert len(np.unique(np.searchsorted(first, i[t < 10.0], side='right'))) > 2      # hits land in several instances
    hs.hs_free(h)


# ---- deep hierarchies: the stack bound behind k_extend_fixup (wavefront.cu) ------------------------------------------------------------------
def _bind_strided(L):
    vp = C.c_void_p
    L.hs_build_wide.restype = vp
    L.hs_build_wide.argtypes = [vp, vp, C.c_int64, C.c_int64]
    L.hs_trace_wide_strided.restype = C.c_int64
    L.hs_trace_wide_strided.argtypes = [vp, vp, vp, C.c_int64, C.c_float, C.c_float, C.c_int, vp, vp, vp, C.POINTER(C.c_int)]
    L.hs_stats_wide.argtypes = [vp, vp]


def _trace_strided(L, h, o, d, far, eps, cap):
    t, i, of = np.empty(len(o), np.float32), np.empty(len(o), np.int32), np.zeros(len(o), np.uint8)
    deepest = C.c_int(0)
    n_of = L.hs_trace_wide_strided(h, o.ctypes.data, d.ctypes.data, len(o), far, eps, cap, t.ctypes.data, i.ctypes.data, of.ctypes.data, C.byref(deepest))
    return t, i, of.astype(bool), int(n_of), deepest.value


def _adversarial_scene(kind, rng, m=6000):
    """Triangle sets that drive the Karras tree to its depth limit, and rays aimed where the hierarchy is deepest."""
    if kind == "morton_chain":
        # centres whose 63-bit Morton keys are 2^k, k = 0..62: every split of the radix tree peels off ONE key -> a 63-level chain of nested
        # boxes around one corner; 24 exact duplicates per centre add index-bit levels underneath (equal keys are split by primitive index);
        # triangle size ~ distance from the corner, so that the peeled-off clusters overlap the rest of the chain
        ext, cs = 4.0, [np.zeros(3), np.full(3, 1.0)]
        for k in range(63):
            c = np.zeros(3)
            c[k % 3] = 2.0 ** (k // 3) / 2.0 ** 21
            cs.append(c)
        cs = np.asarray(cs) * ext
        size = np.maximum(np.linalg.norm(cs, axis=-1), 1e-7)[:, None, None]
        tri = cs[:, None, :] + size * np.asarray([[0, 0, 0], [1, 0, 0.3], [0, 1, 0.6]])[None]
        tri = np.repeat(tri, 24, axis=0) - ext / 2
        corner = np.full(3, -ext / 2)
        org = corner + rng.normal(0, 1, (m, 3)) * np.exp(rng.uniform(-14, 1, (m, 1)))    # every scale of the chain gets rays
        tgt = corner + rng.normal(0, 1, (m, 3)) * np.exp(rng.uniform(-14, 1, (m, 1)))
    elif kind == "duplicates":
        one = np.asarray([[-0.3, -0.2, 0.0], [0.4, -0.1, 0.1], [0.0, 0.5, -0.1]])
        tri = np.repeat(one[None], 6000, axis=0)
        org = rng.uniform(-3, 3, (m, 3))
        tgt = rng.uniform(-0.3, 0.3, (m, 3))
    else:  # "slivers": long needles through the whole scene, every box overlaps every other
        a = rng.uniform(-1, 1, (3000, 3))
        b = a + rng.normal(0, 1.0, (3000, 3))
        tri = np.stack([a, b, b + rng.normal(0, 1e-3, (3000, 3))], 1)
        org = rng.uniform(-3, 3, (m, 3))
        tgt = rng.uniform(-1, 1, (m, 3))
    v = tri.reshape(-1, 3).astype(np.float32)
    f = np.arange(len(v), dtype=np.int32).reshape(-1, 3)
    d = tgt - org
    d = d / np.linalg.norm(d, axis=-1, keepdims=True)
    return v, f, org.astype(np.float32), d.astype(np.float32)


@pytest.mark.parametrize("kind", ["morton_chain", "duplicates", "slivers"])
def test_deep_hierarchies_fit_the_deep_stack_and_overflow_is_always_flagged(hs, kind):
    """k_extend_cw walks with CW_STACK = 48 entries and hands rays that need more to k_extend_fixup, whose CW_DEEP_STACK = 256 entries 'cannot run
    out' (wavefront.cu).  Checked here on the host build of the same traversal: on scenes built to maximise the depth (a 63-level Morton chain
    with duplicate-key subtrees, thousands of identical triangles, scene-spanning slivers) (i) 256 entries are never exhausted and the hits equal
    the exhaustive search bit for bit, (ii) with a deliberately tiny stack every ray whose result could be wrong is FLAGGED -- a ray that is not
    flagged is exact -- which is what makes the hand-off to the fix-up kernel (and the sticky error flag behind it) sufficient."""
    L = hs
    _bind_strided(L)
    rng = np.random.default_rng(7)
    v, f, org, d = _adversarial_scene(kind, rng)
    far = 50.0
    ot, oi = oracle.bruteforce(v, f, org, d, far, 0.0)
    h = L.hs_build_wide(v.ctypes.data, f.ctypes.data, len(v), len(f))
    try:
        st = np.zeros(6, np.int64)
        L.hs_stats_wide(h, st.ctypes.data)
        t, i, flagged, n_of, deepest = _trace_strided(L, h, org, d, far, 0.0, 256)
        assert n_of == 0 and not flagged.any(), "the 256-entry deep stack overflowed on '%s' (%d rays)" % (kind, n_of)
        assert deepest <= st[1], "a ray stacked %d entries in a hierarchy of %d levels: more than one node group per level" % (deepest, st[1])
        assert deepest < 128, "deepest stack use %d leaves less than 2x headroom in CW_DEEP_STACK" % deepest
        assert np.array_equal(t.view(np.int32), ot.view(np.int32)) and np.array_equal(i, oi)
        assert deepest >= 2, "no ray ever stacked two entries: the scene does not exercise the stack"
        for cap in (0, 1, deepest - 1):
            t, i, flagged, n_of, _ = _trace_strided(L, h, org, d, far, 0.0, cap)
            assert n_of == int(flagged.sum()) and flagged.any(), "a %d-entry stack did not overflow although %d entries are needed" % (cap, deepest)
            ok = ~flagged
            assert np.array_equal(t[ok].view(np.int32), ot[ok].view(np.int32)) and np.array_equal(i[ok], oi[ok]), \
                "cap %d: an unflagged ray differs from the exhaustive search" % cap
        print("%s: %d wide nodes in %d levels, deepest stack use %d" % (kind, st[0], st[1], deepest))
    finally:
        L.hs_free(h)
