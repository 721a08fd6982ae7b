"""Seeded test scenes shared by the CPU and GPU tests (torch CPU tensors; sessions move them to the GPU)."""
import numpy as np
import torch

import diffrp_b200 as drp
from diffrp_b200 import synthetic as syn
from diffrp_b200.flatten import flatten_scene, material_descriptions
from diffrp_b200.path_tracing import raygen_tables


def T(a):
    return torch.from_numpy(np.ascontiguousarray(a))


def rigid(seed, scale=1.0, translate=(0, 0, 0)):
    rng = np.random.default_rng(seed)
    q, _ = np.linalg.qr(rng.standard_normal((3, 3)))
    if np.linalg.det(q) < 0:
        q[:, 0] = -q[:, 0]
    m = np.eye(4, dtype=np.float32)
    m[:3, :3] = q * scale
    m[:3, 3] = translate
    return T(m)


def icosphere_scene(env=True, colors=True, rotate=True, subdiv=3):
    """Config 1: icosphere r=0.8 with DefaultMaterial (flat normals -> face soup), gradient env."""
    v, f = syn.icosphere(subdiv, 0.8)
    col = torch.rand(len(v), 4, generator=torch.Generator().manual_seed(5)) if colors else None
    scene = drp.Scene()
    scene.add_mesh_object(drp.MeshObject(drp.DefaultMaterial(), T(v), T(f), color=col, M=rigid(3) if rotate else None))
    if env:
        scene.add_light(drp.ImageEnvironmentLight(intensity=1.0, color=torch.ones(3), image=T(syn.gradient_env())))
    return scene


def gltf_material(seed, size=64, alpha_mode='OPAQUE', normal_map=True, emissive=True, wrap='repeat', interp='linear'):
    tex = lambda c, s, lo=0.0, hi=1.0: drp.GLTFSampler(T(syn.smooth_texture(size, size, c, seed * 10 + s, lo, hi)), wrap, interp)
    nrm = syn.smooth_texture(size, size, 3, seed * 10 + 7, 0.35, 0.65)
    nrm[..., 2] = 0.9
    return drp.GLTFMaterial(
        base_color_factor=torch.tensor([0.9, 0.8, 0.7, 0.9]), base_color_texture=tex(4, 1, 0.2, 1.0),
        metallic_factor=0.8, roughness_factor=0.9, metallic_roughness_texture=tex(3, 2, 0.05, 0.95),
        normal_texture=drp.GLTFSampler(T(nrm), wrap, interp) if normal_map else None, occlusion_texture=None,
        emissive_factor=torch.tensor([0.5, 0.4, 0.3]) if emissive else None, emissive_texture=tex(3, 3, 0.0, 0.3),
        alpha_cutoff=0.55, alpha_mode=alpha_mode)


def mixed_scene(n_theta=48, n_phi=24):
    """Multi-object scene: textured + normal-mapped + emissive GLTF spheres (all alpha modes, all wrap modes),
    a tinted DefaultMaterial ground grid and a smooth-normal DefaultMaterial sphere, env-lit."""
    scene = drp.Scene()
    gv, gf, gn, guv, gt = syn.ground_grid(8, 1.6, -0.75)
    scene.add_mesh_object(drp.MeshObject(drp.DefaultMaterial(torch.tensor([0.7, 0.8, 0.9])), T(gv), T(gf), normals=T(gn), uv=T(guv), tangents=T(gt)))
    specs = [((-0.55, -0.1, 0.0), 'OPAQUE', 'repeat', 'linear'), ((0.55, -0.1, 0.1), 'BLEND', 'mirror', 'linear'),
             ((0.0, 0.45, -0.3), 'MASK', 'clamp', 'point')]
    for k, (c, am, wrap, interp) in enumerate(specs):
        v, f, n, uv, tg = syn.uv_sphere(n_theta, n_phi, radius=0.33, bump=0.02, noise=0.004, seed=k, with_attrs=True)
        mat = gltf_material(k + 1, alpha_mode=am, wrap=wrap, interp=interp, normal_map=(k != 2), emissive=(k != 1))
        col = torch.rand(len(v), 4, generator=torch.Generator().manual_seed(20 + k)) * 0.5 + 0.5
        scene.add_mesh_object(drp.MeshObject(mat, T(v), T(f), normals=T(n), M=rigid(30 + k, 1.0 + 0.1 * k, c), color=col,
                                             uv=T(uv * 3.0 - 0.7), tangents=T(tg)))
    v, f = syn.icosphere(2, 0.25)
    scene.add_mesh_object(drp.MeshObject(drp.DefaultMaterial(), T(v), T(f), normals='smooth', M=rigid(40, 1.0, (0.0, -0.35, 0.55)),
                                         color=torch.rand(len(v), 3, generator=torch.Generator().manual_seed(9))))
    scene.add_light(drp.ImageEnvironmentLight(intensity=1.5, color=torch.tensor([1.0, 0.9, 0.8]), image=T(syn.smooth_texture(32, 64, 3, 77, 0.0, 2.0))))
    return scene


def to_device(scene, dev):
    """Copy of the scene with every tensor on ``dev`` (the reference requires GPU tensors in MeshObject)."""
    mv = lambda x: x.to(dev) if isinstance(x, torch.Tensor) else x
    out = drp.Scene()
    for o in scene.objects:
        m = o.material
        out.objects.append(drp.MeshObject(m, mv(o.verts), mv(o.tris), mv(o.normals), mv(o.M), mv(o.color), mv(o.uv), mv(o.tangents),
                                          {k: mv(v) for k, v in o.custom_attrs.items()}, o.metadata))
    for l in scene.lights:
        out.lights.append(drp.ImageEnvironmentLight(l.intensity, mv(l.color), mv(l.image), l.render_skybox))
    return out


def oracle_inputs(scene, camera, spp, depth, last_bounce='void', step_eps=1e-3, replay_u=None, seed=0, sample_ids=None):
    """Flatten on the CPU and build the oracle's HostScene + host-pointer render params."""
    import oracle
    vao = flatten_scene(scene.objects, 'cpu')
    descs = material_descriptions(scene.objects, 'cpu')
    mats = []
    for d in descs:
        d = dict(d)
        for k in ('base_color_tex', 'mr_tex', 'normal_tex', 'emissive_tex'):
            if d.get(k) is not None:
                d[k] = dict(d[k], image=d[k]['image'].numpy())
        mats.append(d)
    env = None
    for l in scene.lights:
        env = l.image_rh().numpy()
    hs = oracle.HostScene(vao.world_pos.numpy(), vao.world_nrm.numpy(), vao.color.numpy(), vao.uv.numpy(), vao.world_tan.numpy(),
                          vao.tris.numpy(), vao.tri_material.numpy(), mats, env=env)
    H, W = camera.resolution()
    tab = raygen_tables(camera.V().cpu(), camera.P().cpu(), H, W, spp, True, 'cpu')
    ids = np.arange(spp) if sample_ids is None else np.asarray(sample_ids)
    p, keep = oracle.make_params(H, W, depth, tab['t_far'], tab['t_near'], tab['cam_pos'], tab['inv_vp'], tab['ndc_x'].numpy(),
                                 tab['ndc_y'].numpy(), tab['jitter_x'].numpy()[ids], tab['jitter_y'].numpy()[ids], sample_ids=ids,
                                 step_epsilon=step_eps, last_bounce_skybox=(last_bounce == 'skybox'), seed=seed, replay_u=replay_u)
    return vao, hs, p, keep
