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


def affine(seed, translate=(0, 0, 0), scale=(1.0, 1.0, 1.0), shear=0.0):
    """Seeded NON-rigid transform: rotation x anisotropic scale x shear + translation (normals must follow the reference's rule --
    normalize(M3x3 n), not the inverse transpose, base_material.py:183-210 -- which only shows under such matrices)."""
    rng = np.random.default_rng(seed)
    q, _ = np.linalg.qr(rng.standard_normal((3, 3)))
    if np.linalg.det(q) < 0:
        q[:, 0] = -q[:, 0]
    sh = np.eye(3)
    sh[0, 1], sh[2, 0] = shear, -0.5 * shear
    m = np.eye(4, dtype=np.float32)
    m[:3, :3] = q @ np.diag(scale) @ sh
    m[:3, 3] = translate
    return T(m)


def affine_instances_scene(n_inst=6):
    """Config-5-like scene at test size: ``n_inst`` MeshObjects SHARING one mesh (vertex, index and normal tensors) under seeded non-rigid
    transforms with tinted DefaultMaterials, plus one textured, normal-mapped GLTF sphere under a sheared transform (tangent frame under a
    non-rigid M), env-lit."""
    scene = drp.Scene()
    v, f, n, uv, tg = syn.uv_sphere(24, 12, radius=0.22, bump=0.03, noise=0.004, seed=3, with_attrs=True)
    Vt, Ft, Nt = T(v), T(f), T(n)
    tints = ([1, .3, .3], [.3, 1, .3], [.3, .3, 1], [1, 1, .3], [1, .3, 1], [.3, 1, 1])
    rng = np.random.default_rng(11)
    for k in range(n_inst):
        pos = (0.9 * np.cos(2 * np.pi * k / n_inst), 0.25 * np.sin(3.0 * k), 0.9 * np.sin(2 * np.pi * k / n_inst))
        sc = tuple(rng.uniform(0.5, 1.8, 3))
        scene.add_mesh_object(drp.MeshObject(drp.DefaultMaterial(torch.tensor(tints[k % 6], dtype=torch.float32)), Vt, Ft, normals=Nt,
                                             M=affine(50 + k, pos, sc, shear=0.15 * (k % 3))))
    v2, f2, n2, uv2, tg2 = syn.uv_sphere(32, 16, radius=0.3, bump=0.02, noise=0.003, seed=5, with_attrs=True)
    scene.add_mesh_object(drp.MeshObject(gltf_material(4), T(v2), T(f2), normals=T(n2), M=affine(60, (0.0, -0.05, 0.0), (1.4, 0.7, 1.0), shear=0.3),
                                         uv=T(uv2 * 2.0), tangents=T(tg2)))
    scene.add_light(drp.ImageEnvironmentLight(intensity=1.2, color=torch.tensor([0.9, 1.0, 1.0]), image=T(syn.smooth_texture(32, 64, 3, 78, 0.0, 2.0))))
    return scene


def to_device(scene, dev):
    return scene.to(dev)


def oracle_inputs(scene, camera, spp, depth, **kw):
    import oracle
    return oracle.inputs_from_scene(scene, camera, spp, depth, **kw)
