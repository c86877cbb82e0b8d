"""Scene fixtures shared by the CPU and GPU tests (TEST INFRASTRUCTURE)."""
from __future__ import annotations

import os
import tempfile

import numpy as np

import support as S

b2pt = S.b2pt

from b2pt import scenes as _pkg  # noqa: E402  (the scene builders live in the package: bench.py and tools/ use them too)

_keep = _pkg._keep
cornell = _pkg.cornell
chess = _pkg.chess
cornell_sweep = _pkg.cornell_sweep


def two_triangle_scene():
    """Smallest scene: an emissive quad above a rough white floor quad."""
    sc = b2pt.HostScene.empty()
    m = b2pt.Material(b2pt.ROUGH_CONDUCTOR, (20.0, 18.0, 15.0), 1.74, 0.1, 1.0, (0, 0, 0), 0, 0)
    light = sc.add_material("light", m)
    floor = np.array([[-50, 0, -50, 50, 0, -50, 50, 0, 50], [-50, 0, -50, 50, 0, 50, -50, 0, 50]], np.float32)
    quad = np.array([[-10, 40, -10, 10, 40, 10, 10, 40, -10], [-10, 40, -10, -10, 40, 10, 10, 40, 10]], np.float32)
    sc.add_triangles(floor, sc.find_material("rough_white_conductor"))
    sc.add_triangles(quad, light)
    sc.set_camera(32, 32, 60.0, (0, 30, -80), (0, 10, 0))
    return sc.build_tree(), None


def ray_batch(ref: "S.Ref", scene, n_pixels=600, samples=2, seed=1):
    """Rays as the path tracer meets them: camera rays, continuation rays leaving the surfaces they hit,
    and shadow rays towards points on the light (with their reference distances)."""
    rng = np.random.RandomState(seed)
    cam = scene.camera
    px = rng.choice(cam.width * cam.height, size=min(n_pixels, cam.width * cam.height), replace=False).astype(np.int32)
    o, d = ref.camera_rays(px, 0, samples)
    prim, t, co, nn, uv = ref.intersect(o, d)
    hit = prim >= 0
    p = co[hit] + nn[hit] * np.float32(1e-4)
    # continuation rays: random directions
    v = rng.normal(size=p.shape).astype(np.float32)
    v /= np.linalg.norm(v, axis=1, keepdims=True).astype(np.float32)
    # shadow rays: towards reference light samples
    u4 = (np.floor(rng.rand(len(p), 4) * 16777216.0) / 16777216.0).astype(np.float32)
    lp, ln, le, lpdf = ref.sample_light(u4)
    w = (lp - p).astype(np.float32)
    dist = np.sqrt((w * w).sum(1)).astype(np.float32)
    ws = (w / dist[:, None]).astype(np.float32)
    rays_o = np.concatenate([o, p, p]).astype(np.float32)
    rays_d = np.concatenate([d, v, ws]).astype(np.float32)
    return rays_o, rays_d, (p, ws, dist, u4)


def random_scene(seed, width=48, height=36):
    """A random small scene: a lit box of random triangles and spheres with random materials from the nine named ones plus
    randomised parameter variations (roughness, Cauchy coefficients, reflectance), one or two emissive quads, optional sky."""
    rng = np.random.RandomState(seed)
    sc = b2pt.HostScene.empty()
    names = list(b2pt.NAMED_MATERIALS)
    mats = [sc.find_material(n) for n in names]
    for k in range(3):  # parameter variations
        t = int(rng.randint(0, 4))
        m = b2pt.Material(t, (0, 0, 0), float(rng.uniform(1.1, 2.2)), float(rng.uniform(0.0, 0.3)), float(rng.choice([0.01, 0.05, 0.2, 0.6, 1.0])),
                          tuple(rng.uniform(0.05, 1.0, 3).tolist()), 0, 0)
        mats.append(sc.add_material(f"var{k}", m))
    light = sc.add_material("light", b2pt.Material(b2pt.ROUGH_CONDUCTOR, tuple(rng.uniform(8, 40, 3).tolist()), 1.74, 0.1, 1.0, (0, 0, 0), 0, 0))
    # floor and back wall (two triangles each), random objects in front
    def quad(a, b, c, d):
        return np.array([list(a) + list(b) + list(c), list(a) + list(c) + list(d)], np.float32)
    sc.add_triangles(quad((-60, 0, -20), (60, 0, -20), (60, 0, 100), (-60, 0, 100)), int(rng.choice(mats)))
    sc.add_triangles(quad((-60, 0, 100), (60, 0, 100), (60, 80, 100), (-60, 80, 100)), int(rng.choice(mats)))
    for _ in range(int(rng.randint(2, 6))):
        c = rng.uniform([-40, 5, 20], [40, 40, 80])
        tris = (c + rng.normal(0, 9, (int(rng.randint(1, 12)), 3, 3))).reshape(-1, 9).astype(np.float32)
        sc.add_triangles(tris, int(rng.choice(mats)))
    for _ in range(int(rng.randint(0, 4))):
        sc.add_sphere(tuple(rng.uniform([-35, 6, 25], [35, 30, 75]).tolist()), float(rng.uniform(3, 12)), int(rng.choice(mats)))
    for _ in range(int(rng.randint(1, 3))):
        cx, cz, y, s = rng.uniform(-30, 30), rng.uniform(30, 70), rng.uniform(50, 78), rng.uniform(6, 18)
        sc.add_triangles(quad((cx - s, y, cz - s), (cx + s, y, cz + s), (cx + s, y, cz - s), (cx - s, y, cz - s))[:1], light)
        sc.add_triangles(quad((cx - s, y, cz - s), (cx - s, y, cz + s), (cx + s, y, cz + s), (cx + s, y, cz + s))[:1], light)
    env_png = None
    if rng.rand() < 0.5:
        tmp = tempfile.TemporaryDirectory(prefix="b2pt_rand_")
        _keep.append(tmp)
        env_png = _pkg.write_sky_png(os.path.join(tmp.name, "sky.png"), 64, 32, seed)
        sc.load_env_png(env_png)
    else:
        sc.set_background(tuple(rng.uniform(0, 0.3, 3).tolist()))
    dof = bool(rng.rand() < 0.5)
    sc.set_camera(width, height, float(rng.uniform(35, 75)), (float(rng.uniform(-10, 10)), float(rng.uniform(15, 40)), -15.0), (0.0, 20.0, 50.0), (0, 1, 0), dof, 65.0, 1.5)
    sc.set_render(0, float(rng.choice([0.3, 0.5, 0.7, 0.85])), 1, int(rng.choice([1, 2, 4, 5])))
    return sc.build_tree(), env_png
