"""
Generate the golden fixtures under tests/golden/ by running the REFERENCE itself (imported from /root/reference
through refharness.py, CPU device, Patch A+B -- see refharness docstring).  Run in the build container:

    python tests/golden/make_golden.py

The fixtures pin the oracle (tests/test_oracle_golden.py) and, through replayed RNG, the CUDA path
(tests/test_render_gpu.py).  The reference ships no tests of its own for this path.
"""
import os
import sys
import zlib
import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import refharness  # noqa: E402

ref = refharness.load_reference('cpu')
import diffrp  # noqa: E402  (the reference)
from diffrp.utils.raycaster import BruteForceRaycaster, NaivePBBVH  # noqa: E402
from diffrp.utils.geometry import barycentric  # noqa: E402
from diffrp.utils.coordinates import unit_direction_to_latlong_uv, near_plane_ndc_grid  # noqa: E402
import scenes  # noqa: E402  (tests/scenes.py, builds diffrp_b200 scenes)
import diffrp_b200 as drp  # noqa: E402
from diffrp_b200 import synthetic as syn  # noqa: E402

FAR = 10.0


def save(name, **arrays):
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **{k: np.asarray(v) for k, v in arrays.items()})
    print("wrote %s (%.1f KiB)" % (path, os.path.getsize(path) / 1024))


def crc(a):
    return np.uint32(zlib.crc32(np.ascontiguousarray(a).tobytes()))


def to_reference_scene(scene):
    """diffrp_b200 Scene (already preprocessed objects) -> reference Scene with the reference's own classes."""
    out = diffrp.Scene()
    for o in scene.objects:
        m = o.material
        if isinstance(m, drp.GLTFMaterial):
            smp = lambda s: None if s is None else diffrp.GLTFSampler(s.image, s.wrap_mode, s.interpolation)
            rm = diffrp.GLTFMaterial(m.base_color_factor, smp(m.base_color_texture), m.metallic_factor, m.roughness_factor,
                                     smp(m.metallic_roughness_texture), smp(m.normal_texture), smp(m.occlusion_texture),
                                     m.emissive_factor, smp(m.emissive_texture), m.alpha_cutoff, m.alpha_mode)
        else:
            rm = diffrp.DefaultMaterial(m.tint)
        col = o.color if o.color.shape[-1] == 4 else torch.cat([o.color, torch.ones_like(o.color[:, :1])], -1)
        out.objects.append(diffrp.MeshObject(rm, o.verts, o.tris, o.normals, o.M, col, o.uv, o.tangents, dict(o.custom_attrs), {}))
    for l in scene.lights:
        out.add_light(diffrp.ImageEnvironmentLight(l.intensity, l.color, l.image, l.render_skybox))
    return out


def raycast_fixtures():
    v, f = syn.icosphere(3, 0.8)
    o, d = syn.random_rays(20000)
    tv, tf, to, td = map(torch.from_numpy, (v, f, o, d))
    bt, bi = BruteForceRaycaster(tv, tf, {'epsilon': 1e-8}).query(to, td, FAR)
    out = dict(verts=v, tris=f, rays_o=o, rays_d=d, far=np.float32(FAR), brute_t=bt.numpy(), brute_i=bi.numpy().astype(np.int32))
    for builder in ('splitaxis', 'morton'):
        t, i = NaivePBBVH(tv, tf, {'epsilon': 1e-8, 'builder': builder}).query(to, td, FAR)
        out['bvh_%s_t' % builder] = t.numpy()
        out['bvh_%s_i' % builder] = i.numpy().astype(np.int32)
    save("raycast_icosphere", **out)

    # config 1 primary rays: unrotated icosphere, 64x64, 16 spp -- the symmetric edge-plane stress case
    scene = to_reference_scene(scenes.icosphere_scene(rotate=False, colors=False))
    cam = diffrp.PerspectiveCamera(h=64, w=64)
    sess = diffrp.PathTracingSession(scene, cam, diffrp.PathTracingSessionOptions(ray_spp=16, raycaster_impl='brute-force'))
    H, W = 64, 64
    grid = near_plane_ndc_grid(H, W, torch.float32, torch.device('cpu')).reshape(-1, 4)
    qx, qy = diffrp.hammersley(16, True, torch.device('cpu'))
    qx, qy = qx[..., None, None], qy[..., None, None]
    g = torch.cat([grid[..., 0:1] + (qx - 0.5) * (2 / W), grid[..., 1:2] + (qy - 0.5) * (2 / H), grid[..., 2:].expand(16, -1, -1)], -1).reshape(-1, 4)
    ro, rd = sess._view_dir_impl(g, sess.camera_V(), sess.camera_P(), sess.camera_VP())
    ro, rd = ro[:4 * H * W].contiguous(), rd[:4 * H * W].contiguous()  # samples 0..3 (sample 0 holds the edge-plane rays)
    vao = sess.vertex_array_object()
    far = sess.camera_far()
    bt, bi = BruteForceRaycaster(vao.world_pos, vao.tris, {'epsilon': 1e-8}).query(ro, rd, far)
    nt, ni = NaivePBBVH(vao.world_pos, vao.tris, {'epsilon': 1e-8, 'builder': 'splitaxis'}).query(ro, rd, far)
    save("raycast_c1_primary", verts=vao.world_pos.numpy(), tris=vao.tris.numpy(), rays_o=ro.numpy(), rays_d=rd.numpy(), far=np.float32(far),
         brute_t=bt.numpy(), brute_i=bi.numpy().astype(np.int32), bvh_t=nt.numpy(), bvh_i=ni.numpy().astype(np.int32),
         V=sess.camera_V().numpy(), P=sess.camera_P().numpy())


def function_fixtures():
    g = torch.Generator().manual_seed(11)
    n = 2048
    # _sampler_brdf_impl with replayed uniforms
    albedo = torch.rand(n, 3, generator=g)
    normal = torch.nn.functional.normalize(torch.randn(n, 3, generator=g), dim=-1)
    normal[:8] = torch.tensor([0.0, 1.0, 0.0])  # the up-vector switch of combine_fixed_tangent_space
    metal, smooth = torch.rand(n, 1, generator=g), torch.rand(n, 1, generator=g)
    alpha = torch.where(torch.rand(n, 1, generator=g) < 0.3, torch.rand(n, 1, generator=g), torch.ones(n, 1))
    emission = torch.rand(n, 3, generator=g) * 0.2
    attrs = torch.cat([albedo, normal, metal, smooth, alpha, emission], -1)
    attrs[-64:] = 0  # misses: zero g-buffer rows
    t = torch.rand(n, generator=g) * 3 + 0.1
    ro = torch.randn(n, 3, generator=g)
    rd = torch.nn.functional.normalize(-normal + 0.5 * torch.randn(n, 3, generator=g), dim=-1)
    env = torch.rand(n, 3, generator=g)
    torch.manual_seed(1234)
    u = torch.stack([torch.rand(n, 1) for _ in range(6)]).reshape(6, n)
    torch.manual_seed(1234)
    outs = diffrp.PathTracingSession._sampler_brdf_impl(attrs, t, ro, rd, env)
    names = ('albedo', 'emission', 'world_normal', 'alpha', 'radiance', 'transfer', 'next_o', 'next_d')
    save("sampler_brdf", attrs=attrs.numpy(), t=t.numpy(), rays_o=ro.numpy(), rays_d=rd.numpy(), env=env.numpy(), u6=u.numpy(),
         **{'out_' + k: (v.expand(n, 3) if k == 'transfer' else v).numpy() for k, v in zip(names, outs)})

    # textures: sample2d for every GLTFSampler wrap / interpolation + the env lookup path
    img = torch.from_numpy(syn.smooth_texture(13, 17, 4, 3))
    uv = torch.rand(n, 2, generator=g) * 5 - 2
    uv[:16] = torch.tensor([[0.0, 0.0], [1.0, 1.0], [0.5, 0.5], [1.0, 0.0]]).repeat(4, 1)
    tex = {}
    for wrap in ('repeat', 'clamp', 'mirror'):
        for interp in ('linear', 'point'):
            tex['%s_%s' % (wrap, interp)] = diffrp.GLTFSampler(img, wrap, interp).sample(uv).numpy()
    d = torch.nn.functional.normalize(torch.randn(n, 3, generator=g), dim=-1)
    d[:6] = torch.tensor([[0, 1, 0], [0, -1, 0], [0, 0, 1], [0, 0, -1], [1, 0, 0], [-1, 0, 0]], dtype=torch.float32)
    envimg = torch.from_numpy(syn.smooth_texture(16, 32, 3, 4, 0, 2))
    light = diffrp.ImageEnvironmentLight(1.5, torch.tensor([1.0, 0.9, 0.8]), envimg)
    lu, lv = unit_direction_to_latlong_uv(d)
    envs = diffrp.sample2d(light.image_rh(), torch.cat([lu, lv], -1))
    save("textures", image=img.numpy(), uv=uv.numpy(), dirs=d.numpy(), env_image_rh=light.image_rh().numpy(), env_out=envs.numpy(),
         latlong_u=lu.numpy(), latlong_v=lv.numpy(), **tex)

    # barycentric from the hit point + hammersley + camera
    a, b, c = (torch.randn(n, 3, generator=g) for _ in range(3))
    w = torch.rand(n, 3, generator=g)
    w = w / w.sum(-1, keepdim=True)
    pnt = w[:, :1] * a + w[:, 1:2] * b + w[:, 2:] * c + 0.01 * torch.randn(n, 3, generator=g)
    c[:4] = a[:4]  # degenerate triangles: nan -> 0
    bary = barycentric(a, b, c, pnt)
    hx, hy = diffrp.hammersley(16, True, torch.device('cpu'))
    hx2, hy2 = diffrp.hammersley(1024, True, torch.device('cpu'))
    cam = diffrp.PerspectiveCamera.from_orbit(h=36, w=48, radius=2.5, azim=40, elev=25, origin=[0.1, 0.0, -0.2], fov=35, near=0.05, far=20.0)
    save("misc", a=a.numpy(), b=b.numpy(), c=c.numpy(), p=pnt.numpy(), bary=bary.numpy(), ham16_x=hx.numpy(), ham16_y=hy.numpy(),
         ham1024_x=hx2.numpy(), ham1024_y=hy2.numpy(), orbit_V=cam.V().numpy(), orbit_P=cam.P().numpy(),
         default_V=diffrp.PerspectiveCamera(h=64, w=64).V().numpy(), default_P=diffrp.PerspectiveCamera(h=64, w=64).P().numpy())


def pbr_fixture(name, scene, cam_kwargs, spp, depth, last_bounce, seed, orbit=None, impl='brute-force'):
    rscene = to_reference_scene(scene)
    cam = diffrp.PerspectiveCamera.from_orbit(**orbit) if orbit else diffrp.PerspectiveCamera(**cam_kwargs)
    opts = diffrp.PathTracingSessionOptions(ray_spp=spp, ray_depth=depth, raycaster_impl=impl, pbr_ray_last_bounce=last_bounce)
    sess = diffrp.PathTracingSession(rscene, cam, opts)
    H, W = cam.resolution()
    torch.manual_seed(seed)
    rad, alpha, extras = sess.pbr()
    torch.manual_seed(seed)
    u = torch.stack([torch.rand(spp * H * W, 1) for _ in range(depth * 6)])
    save(name, radiance=rad.numpy(), alpha=alpha.numpy(), **{k: v.numpy() for k, v in extras.items()}, spp=spp, depth=depth, seed=seed,
         last_bounce=last_bounce, H=H, W=W, u_crc=crc(u.numpy()), far=np.float32(sess.camera_far()))


def scene_api_fixture():
    """MeshObject.preprocess ('flat' face soup / 'smooth' normals, objects.py:71-98) and Scene.static_batching (scene.py:33-75)."""
    v, f = syn.icosphere(1, 0.5)
    V, F = torch.from_numpy(v), torch.from_numpy(f)
    rnd = lambda c, s: torch.rand(len(v), c, generator=torch.Generator().manual_seed(s))  # noqa: E731
    out = {}
    for mode in ('flat', 'smooth'):
        o = diffrp.MeshObject(diffrp.DefaultMaterial(), V.clone(), F.clone(), normals=mode, color=rnd(4, 2), uv=rnd(2, 3), tangents=rnd(4, 4),
                              custom_attrs={'w': rnd(2, 5)}).preprocess()
        for k in ('verts', 'tris', 'normals', 'color', 'uv', 'tangents', 'M'):
            out['%s_%s' % (mode, k)] = getattr(o, k).numpy()
        out['%s_custom_w' % mode] = o.custom_attrs['w'].numpy()
    m1, m2 = diffrp.DefaultMaterial(), diffrp.DefaultMaterial(torch.tensor([0.5, 0.6, 0.7]))
    sc = diffrp.Scene()
    for k, (mat, seed) in enumerate([(m1, 10), (m2, 11), (m1, 12)]):
        sc.add_mesh_object(diffrp.MeshObject(mat, V.clone() * (1 + 0.1 * k), F.clone(), normals='smooth', M=scenes.rigid(seed, 1.0 + 0.2 * k, (k * 0.3, 0, 0)),
                                             tangents=rnd(4, 20 + k)))
    sc.static_batching()
    out['batched_n'] = np.int32(len(sc.objects))
    for j, o in enumerate(sc.objects):
        for k in ('verts', 'tris', 'normals', 'color', 'uv', 'tangents', 'M'):
            out['batched%d_%s' % (j, k)] = getattr(o, k).numpy()
    save("scene_api", **out)


def tonemap_fixture():
    """Colour epilogue: linear_to_srgb (colors.py:33-42), linear_to_alexa_logc_ei1000 (colors.py:94-102),
    agx_base_contrast (tone_mapping.py:21-35) run unmodified with its LUT loader swapped for a synthetic LUT (the shipped LUT is a
    data resource of the reference and is not copied), to_pil's byte conversion (exchange.py:7-18)."""
    from diffrp.utils import tone_mapping
    from diffrp.utils.colors import linear_to_srgb, linear_to_alexa_logc_ei1000
    from diffrp.utils.shader_ops import saturate
    g = torch.Generator().manual_seed(77)
    n = 6000
    rgb = torch.exp(torch.randn(n, 3, generator=g) * 2.5 - 1.5)                     # HDR, log-normal
    rgb[:64] = torch.tensor([0.0, 0.0031308, 0.0031307, 0.010591, 0.010592, 1.0, 1e-8, 65504.0]).repeat(8)[:, None]
    rgb[64:128] = -torch.rand(64, 3, generator=g) * 0.01                           # slightly negative (fireflies after filtering)
    rgb[128:192] = torch.rand(64, 3, generator=g) * 0.02                           # around both cuts
    # smooth synthetic LUT, stored the way AgxLutLoader.load returns it (fliplr already applied), z y x 3
    N = 16
    zz, yy, xx = torch.meshgrid(*(torch.linspace(0, 1, N),) * 3, indexing='ij')
    lut_file = torch.stack([xx ** 1.5 * (1 - 0.2 * yy), 0.9 * yy + 0.1 * zz * xx, zz ** 0.7 * (0.5 + 0.5 * yy)], -1).contiguous()
    lut_loaded = torch.fliplr(lut_file).contiguous()

    class Loader:
        def load(self, variant):
            assert variant == "base-contrast"
            return lut_loaded
    tone_mapping.agx_lut_loader = Loader()
    agx = tone_mapping.agx_base_contrast(rgb)
    alpha = torch.rand(n, 1, generator=g) * 1.2 - 0.1
    srgb = linear_to_srgb(rgb)
    byte_agx = (saturate(torch.cat([agx, alpha], -1).float()) * 255).byte()
    byte_srgb = (saturate(torch.cat([srgb, alpha], -1).float()) * 255).byte()
    save("tonemap", rgb=rgb.numpy(), alpha=alpha.numpy(), lut=lut_loaded.numpy(), logc=linear_to_alexa_logc_ei1000(rgb).numpy(),
         srgb=srgb.numpy(), agx=agx.numpy(), byte_agx=byte_agx.numpy(), byte_srgb=byte_srgb.numpy())


def denoiser_fixture():
    """run_denoiser + UNet(9, 3) of the reference (rendering/denoiser.py:24-35, 72-173), unmodified, on the CPU in fp32, with the seeded
    stand-in parameters of diffrp_b200.denoiser.UNetWeights.random (the OIDN weight file is a data resource and is not copied).
    Odd image size so that the reflection padding (3 / 4 rows, 5 / 6 columns) is exercised."""
    from diffrp.rendering import denoiser as rd
    from diffrp_b200.denoiser import UNetWeights
    _, sd = UNetWeights.random(seed=11, device='cpu')
    net = rd.UNet(9, 3)
    net.load_state_dict(sd)
    g = torch.Generator().manual_seed(12)
    h, w = 41, 53
    hdr = torch.exp(torch.randn(h, w, 3, generator=g) * 1.5 - 1.0)
    hdr[:2] = 0.0
    hdr[2, :8] = torch.tensor([1e-7, 1e-6, 1e-3, 0.03, 0.04, 1.0, 100.0, 60000.0])[:, None]
    albedo = torch.rand(h, w, 3, generator=g)
    normal = torch.nn.functional.normalize(torch.randn(h, w, 3, generator=g), dim=-1)
    with torch.no_grad():
        out = rd.run_denoiser(net.float(), hdr, albedo, normal)
    save("denoiser", hdr=hdr.numpy(), albedo=albedo.numpy(), normal=normal.numpy(), out=out.numpy(), seed=np.int32(11))


AFFINE_ORBIT = dict(h=40, w=56, radius=3.2, azim=-35, elev=22, origin=[0.0, 0.0, 0.0], fov=34, near=0.1, far=10.0)


if __name__ == "__main__":
    torch.set_num_threads(8)
    if len(sys.argv) > 1 and sys.argv[1] == "affine":   # only the fixture added last (the others regenerate bit for bit)
        pbr_fixture("pbr_affine_instances", scenes.affine_instances_scene(), None, 3, 3, 'skybox', 7, orbit=AFFINE_ORBIT)
        sys.exit(0)
    denoiser_fixture()
    tonemap_fixture()
    scene_api_fixture()
    raycast_fixtures()
    function_fixtures()
    pbr_fixture("pbr_icosphere", scenes.icosphere_scene(), dict(h=32, w=32), 4, 2, 'void', 0)
    pbr_fixture("pbr_icosphere_skybox", scenes.icosphere_scene(), dict(h=24, w=40), 3, 3, 'skybox', 1)
    orbit = dict(h=48, w=64, radius=3.0, azim=25, elev=15, origin=[0.0, -0.1, 0.0], fov=32, near=0.1, far=10.0)
    pbr_fixture("pbr_mixed_void", scenes.mixed_scene(), None, 4, 3, 'void', 2, orbit=orbit)
    pbr_fixture("pbr_mixed_skybox", scenes.mixed_scene(), None, 2, 4, 'skybox', 3, orbit=orbit)
    # high-spp image for the statistical (PSNR) comparison of the native counter-based RNG against the reference's torch RNG
    pbr_fixture("pbr_mixed_256spp", scenes.mixed_scene(), None, 256, 3, 'skybox', 5, orbit=dict(orbit, h=36, w=48), impl='naive-pbbvh')
    # config 1 in full: icosphere, 64x64, 16 spp, 2 bounces (reference on CPU)
    pbr_fixture("pbr_config1", scenes.icosphere_scene(rotate=False, colors=False), dict(h=64, w=64), 16, 2, 'void', 0, impl='naive-pbbvh')
    # config-5-like: objects sharing one mesh under NON-rigid transforms (normals by M, not its inverse transpose) + a sheared GLTF sphere
    pbr_fixture("pbr_affine_instances", scenes.affine_instances_scene(), None, 3, 3, 'skybox', 7, orbit=AFFINE_ORBIT)
