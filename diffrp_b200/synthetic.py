"""
Seeded synthetic geometry / rays / textures for tests and benchmarks (numpy, host side).

The shapes follow SURVEY.md section 8(d): icosphere (config 1), displaced UV sphere (configs 2/4),
ground grid + displaced spheres (config 3).  Nothing here touches the GPU.
"""
import math
import numpy as np


def icosphere(subdiv: int = 3, radius: float = 0.8):
    """Icosahedron subdivided ``subdiv`` times: 10*4^s + 2 verts, 20*4^s faces (642 / 1280 at s=3)."""
    p = (1.0 + math.sqrt(5.0)) / 2.0
    v = [(-1, p, 0), (1, p, 0), (-1, -p, 0), (1, -p, 0), (0, -1, p), (0, 1, p), (0, -1, -p), (0, 1, -p),
         (p, 0, -1), (p, 0, 1), (-p, 0, -1), (-p, 0, 1)]
    f = [(0, 11, 5), (0, 5, 1), (0, 1, 7), (0, 7, 10), (0, 10, 11), (1, 5, 9), (5, 11, 4), (11, 10, 2), (10, 7, 6),
         (7, 1, 8), (3, 9, 4), (3, 4, 2), (3, 2, 6), (3, 6, 8), (3, 8, 9), (4, 9, 5), (2, 4, 11), (6, 2, 10),
         (8, 6, 7), (9, 8, 1)]
    verts = [np.array(x, dtype=np.float64) / math.sqrt(1 + p * p) for x in v]
    faces = list(f)
    for _ in range(subdiv):
        cache = {}

        def mid(a, b):
            key = (a, b) if a < b else (b, a)
            if key not in cache:
                m = verts[a] + verts[b]
                verts.append(m / np.linalg.norm(m))
                cache[key] = len(verts) - 1
            return cache[key]
        nf = []
        for a, b, c in faces:
            ab, bc, ca = mid(a, b), mid(b, c), mid(c, a)
            nf += [(a, ab, ca), (b, bc, ab), (c, ca, bc), (ab, bc, ca)]
        faces = nf
    return (np.array(verts) * radius).astype(np.float32), np.array(faces, dtype=np.int32)


def uv_sphere(n_theta: int, n_phi: int, radius: float = 0.8, bump: float = 0.05, noise: float = 0.01, seed: int = 0,
              center=(0.0, 0.0, 0.0), with_attrs: bool = False):
    """
    Displaced UV sphere with n_theta x n_phi quads = 2*n_theta*n_phi triangles (pole rows are zero-area
    triangles on purpose).  r = radius + bump*sin(7 theta)cos(5 phi) + noise*U(0,1).
    With ``with_attrs`` also returns smooth normals (radial), uv and analytic tangents (d/dtheta, w=1).
    """
    rng = np.random.default_rng(seed)
    th = np.linspace(0.0, 2.0 * np.pi, n_theta + 1)
    ph = np.linspace(-0.5 * np.pi, 0.5 * np.pi, n_phi + 1)
    T, P = np.meshgrid(th, ph, indexing='xy')  # (n_phi+1, n_theta+1)
    r = radius + bump * np.sin(7 * T) * np.cos(5 * P) + noise * rng.random(T.shape)
    r[:, -1] = r[:, 0]
    x, y, z = r * np.sin(T) * np.cos(P), r * np.sin(P), r * np.cos(T) * np.cos(P)
    verts = (np.stack([x, y, z], -1).reshape(-1, 3) + np.asarray(center)).astype(np.float32)
    j, i = np.meshgrid(np.arange(n_phi), np.arange(n_theta), indexing='ij')
    a = (j * (n_theta + 1) + i).reshape(-1)
    b, c, d = a + 1, a + (n_theta + 1), a + (n_theta + 2)
    tris = np.concatenate([np.stack([a, b, d], -1), np.stack([a, d, c], -1)], 0).astype(np.int32)
    if not with_attrs:
        return verts, tris
    nrm = np.stack([np.sin(T) * np.cos(P), np.sin(P), np.cos(T) * np.cos(P)], -1).reshape(-1, 3).astype(np.float32)
    uv = np.stack([T / (2 * np.pi), (P + 0.5 * np.pi) / np.pi], -1).reshape(-1, 2).astype(np.float32)
    tan = np.stack([np.cos(T), np.zeros_like(T), -np.sin(T), np.ones_like(T)], -1).reshape(-1, 4).astype(np.float32)
    return verts, tris, nrm, uv, tan


def ground_grid(n: int, half: float = 2.0, y: float = -1.0):
    """n x n quad grid in the plane y = const: 2 n^2 triangles, normals +y, uv tiling 4x."""
    g = np.linspace(-half, half, n + 1)
    X, Z = np.meshgrid(g, g, indexing='xy')
    verts = np.stack([X, np.full_like(X, y), Z], -1).reshape(-1, 3).astype(np.float32)
    j, i = np.meshgrid(np.arange(n), np.arange(n), indexing='ij')
    a = (j * (n + 1) + i).reshape(-1)
    b, c, d = a + 1, a + (n + 1), a + (n + 2)
    tris = np.concatenate([np.stack([a, d, b], -1), np.stack([a, c, d], -1)], 0).astype(np.int32)
    nrm = np.tile(np.array([[0, 1, 0]], np.float32), (len(verts), 1))
    uv = (np.stack([X, Z], -1).reshape(-1, 2) / (2 * half) * 4.0 + 0.5).astype(np.float32)
    tan = np.tile(np.array([[1, 0, 0, 1]], np.float32), (len(verts), 1))
    return verts, tris, nrm, uv, tan


def random_rays(n: int, origin_radius: float = 3.0, target_sigma: float = 0.5, seed: int = 1):
    """Origins uniform on a sphere, directions towards N(0, sigma^2) targets (config 2)."""
    rng = np.random.default_rng(seed)
    o = rng.standard_normal((n, 3))
    o = o / np.linalg.norm(o, axis=-1, keepdims=True) * origin_radius
    tgt = np.random.default_rng(seed + 1).standard_normal((n, 3)) * target_sigma
    d = tgt - o
    d = d / np.linalg.norm(d, axis=-1, keepdims=True)
    return o.astype(np.float32), d.astype(np.float32)


def smooth_texture(h: int, w: int, c: int, seed: int, lo: float = 0.0, hi: float = 1.0, octaves: int = 4):
    """Seeded band-limited fp32 texture in [lo, hi]."""
    rng = np.random.default_rng(seed)
    yy, xx = np.meshgrid(np.linspace(0, 1, h, endpoint=False), np.linspace(0, 1, w, endpoint=False), indexing='ij')
    img = np.zeros((h, w, c))
    for k in range(c):
        acc = np.zeros((h, w))
        for o in range(octaves):
            f = 2 ** o
            ph = rng.random(4) * 2 * np.pi
            acc += (np.sin(2 * np.pi * f * xx + ph[0]) * np.cos(2 * np.pi * f * yy + ph[1])
                    + np.sin(2 * np.pi * f * (xx + yy) + ph[2]) * 0.5) / f
        acc = (acc - acc.min()) / (acc.max() - acc.min() + 1e-12)
        img[..., k] = lo + (hi - lo) * acc
    return img.astype(np.float32)


def gradient_env(h: int = 64, w: int = 128, top: float = 2.0):
    """Deterministic lat-long environment: linspace(0, top) ramp (config 1)."""
    return np.linspace(0.0, top, h * w * 3, dtype=np.float32).reshape(h, w, 3)


# ---- benchmark scenes -------------------------------------------------------------------------------

def torch_noise_texture(h, w, c, seed, lo=0.0, hi=1.0, octaves=5, device='cpu'):
    """Seeded multi-octave value noise (bilinear-upsampled random grids) as an (h, w, c) fp32 torch tensor."""
    import torch
    import torch.nn.functional as F
    g = torch.Generator().manual_seed(seed)
    acc = torch.zeros(1, c, h, w)
    amp, tot = 1.0, 0.0
    for o in range(octaves):
        n = 4 * 2 ** o
        grid = torch.rand(1, c, n, n, generator=g)
        acc += amp * F.interpolate(grid, size=(h, w), mode='bilinear', align_corners=False)
        tot += amp
        amp *= 0.5
    img = (acc / tot)[0].permute(1, 2, 0)
    return (lo + (hi - lo) * img).contiguous().to(device)


def teaser_scene(device='cpu', n_spheres=48, sphere_res=(160, 128), ground=256, tex=1024, n_materials=8, seed=0, pin=False):
    """
    Config 3: ground quad grid + ``n_spheres`` displaced UV spheres, ``n_materials`` GLTFMaterials with seeded
    tex x tex fp32 textures (base RGBA, metallic-roughness, normal, emissive), analytic tangents, seeded HDR env.
    Defaults give exactly 48*2*160*128 + 2*256^2 = 2,097,152 triangles.  Returns (Scene, orbit-camera kwargs).
    """
    import torch
    from .scene import Scene, MeshObject, ImageEnvironmentLight
    from .materials import GLTFMaterial, GLTFSampler
    put = (lambda t: t.pin_memory()) if pin else (lambda t: t.to(device))
    Tn = lambda a: put(torch.from_numpy(np.ascontiguousarray(a)))
    mats = []
    for k in range(n_materials):
        nrm = torch_noise_texture(tex, tex, 3, seed * 100 + k * 10 + 3, 0.35, 0.65)
        nrm[..., 2] = 0.92
        mats.append(GLTFMaterial(
            base_color_factor=put(torch.tensor([1.0, 1.0, 1.0, 1.0])),
            base_color_texture=GLTFSampler(put(torch_noise_texture(tex, tex, 4, seed * 100 + k * 10 + 1, 0.15, 0.95))),
            metallic_factor=0.2 + 0.1 * k, roughness_factor=0.9,
            metallic_roughness_texture=GLTFSampler(put(torch_noise_texture(tex, tex, 3, seed * 100 + k * 10 + 2, 0.1, 0.9))),
            normal_texture=GLTFSampler(put(nrm)), occlusion_texture=None,
            emissive_factor=put(torch.tensor([0.3, 0.25, 0.2])) if k % 4 == 0 else None,
            emissive_texture=GLTFSampler(put(torch_noise_texture(tex, tex, 3, seed * 100 + k * 10 + 4, 0.0, 0.5))),
            alpha_cutoff=0.5, alpha_mode='OPAQUE'))
    scene = Scene()
    gv, gf, gn, guv, gt = ground_grid(ground, 2.2, -0.6)
    scene.add_mesh_object(MeshObject(mats[0], Tn(gv), Tn(gf), normals=Tn(gn), uv=Tn(guv), tangents=Tn(gt)))
    rng = np.random.default_rng(seed)
    cols = int(np.ceil(np.sqrt(n_spheres * 4 / 3)))
    rows = int(np.ceil(n_spheres / cols))
    for k in range(n_spheres):
        cx = (k % cols - (cols - 1) / 2) * (3.6 / cols)
        cz = (k // cols - (rows - 1) / 2) * (3.0 / rows)
        r = 0.16 + 0.06 * rng.random()
        v, f, n, uv, tg = uv_sphere(sphere_res[0], sphere_res[1], radius=r, bump=0.012, noise=0.002, seed=seed + k,
                                    center=(cx, -0.6 + r + 0.02 + 0.25 * rng.random(), cz), with_attrs=True)
        scene.add_mesh_object(MeshObject(mats[k % n_materials], Tn(v), Tn(f), normals=Tn(n), uv=Tn(uv * 2.0), tangents=Tn(tg)))
    env = torch_noise_texture(256, 512, 3, seed * 100 + 99, 0.0, 1.0)
    env = env ** 3 * 6.0 + 0.05  # a few bright regions
    scene.add_light(ImageEnvironmentLight(intensity=1.0, color=put(torch.ones(3)), image=put(env.contiguous())))
    cam = dict(radius=4.6, azim=32.0, elev=24.0, origin=[0.0, -0.35, 0.0], fov=40.0, near=0.1, far=20.0)
    return scene, cam


def rigid_matrix(rng, scale=1.0, translate=(0.0, 0.0, 0.0)):
    """Seeded rigid transform (random rotation x uniform scale + translation) as a (4, 4) fp32 numpy array."""
    q, _ = np.linalg.qr(rng.standard_normal((3, 3)))
    if np.linalg.det(q) < 0:
        q[:, 0] = -q[:, 0]
    m = np.eye(4, dtype=np.float32)
    m[:3, :3] = q * scale
    m[:3, 3] = translate
    return m


def datagen_scene(device='cpu', n_theta=500, n_phi=500, env_res=(256, 512)):
    """
    Config 4 (multi-view datagen): one displaced UV sphere of 2 * n_theta * n_phi triangles (500,000 by default) with a
    ``DefaultMaterial`` and seeded vertex colours, env-lit.  Returns (Scene, camera factory: (k, n_views, res) -> orbit kwargs).
    """
    import torch
    from .scene import Scene, MeshObject, ImageEnvironmentLight
    from .materials import DefaultMaterial
    Tn = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(device)
    v, f, n, uv, tg = uv_sphere(n_theta, n_phi, radius=0.8, bump=0.05, noise=0.01, seed=0, with_attrs=True)
    col = (torch.rand(len(v), 4, generator=torch.Generator().manual_seed(3)) * 0.6 + 0.4).to(device)
    scene = Scene().add_mesh_object(MeshObject(DefaultMaterial(), Tn(v), Tn(f), normals=Tn(n), color=col, uv=Tn(uv)))
    env = (torch_noise_texture(env_res[0], env_res[1], 3, 7, 0.0, 1.0) ** 3 * 5.0 + 0.1).to(device)
    scene.add_light(ImageEnvironmentLight(1.0, torch.ones(3, device=device), env))

    def orbit(k, n_views, res=512):
        return dict(h=res, w=res, radius=3.0, azim=360.0 * k / n_views, elev=20.0 * math.sin(2 * math.pi * k / n_views), origin=[0, 0, 0])
    return scene, orbit


def instanced_scene(device='cpu', n_instances=1000, mesh_res=(100, 50), env_res=(256, 512), seed=0, spread=(1.6, 0.9, 1.0)):
    """
    Config 5 (instanced scene): ``n_instances`` MeshObjects that SHARE one displaced-sphere mesh (2 * mesh_res[0] * mesh_res[1]
    triangles: 10,000 by default -> 10 M triangles flattened) with seeded rigid transforms and one ``DefaultMaterial`` each of 8
    tints, env-lit.  Returns (Scene, orbit-camera kwargs without the resolution).
    """
    import torch
    from .scene import Scene, MeshObject, ImageEnvironmentLight
    from .materials import DefaultMaterial
    Tn = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(device)
    v, f, n, uv, tg = uv_sphere(mesh_res[0], mesh_res[1], radius=0.045, bump=0.004, noise=0.001, seed=1, with_attrs=True)
    Vt, Ft, Nt = Tn(v), Tn(f), Tn(n)
    tints = ([1, .3, .3], [.3, 1, .3], [.3, .3, 1], [1, 1, .3], [1, .3, 1], [.3, 1, 1], [.9, .9, .9], [.5, .5, .5])
    mats = [DefaultMaterial(torch.tensor(c, dtype=torch.float32, device=device)) for c in tints]
    rng = np.random.default_rng(seed)
    scene = Scene()
    lo, hi = [-spread[0], -spread[1], -spread[2]], list(spread)
    for k in range(n_instances):
        pos = rng.uniform(lo, hi)
        scene.add_mesh_object(MeshObject(mats[k % 8], Vt, Ft, normals=Nt, M=Tn(rigid_matrix(rng, 0.6 + rng.random(), pos))))
    env = (torch_noise_texture(env_res[0], env_res[1], 3, 7, 0.0, 1.0) ** 3 * 5.0 + 0.1).to(device)
    scene.add_light(ImageEnvironmentLight(1.0, torch.ones(3, device=device), env))
    return scene, dict(radius=4.0, azim=20.0, elev=10.0, origin=[0, 0, 0], fov=35)
