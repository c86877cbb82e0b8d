"""Seeded input generators shared by the CPU and GPU parity tests (TEST INFRASTRUCTURE)."""
import numpy as np


def uniforms(rng, *shape):
    """Uniforms as the sample streams produce them: multiples of 2^-24 in [0, 1)."""
    return (np.floor(rng.rand(*shape) * 16777216.0) / 16777216.0).astype(np.float32)

def rel_close(a, b, rel, abs_=0.0):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    both_nan = np.isnan(a) & np.isnan(b)
    both_inf = np.isinf(a) & np.isinf(b) & (np.sign(a) == np.sign(b))
    with np.errstate(invalid="ignore"):
        ok = np.abs(a - b) <= abs_ + rel * np.maximum(np.abs(a), np.abs(b))
    return ok | both_nan | both_inf

def adversarial_triangle_cases(rng, n):
    v = (rng.rand(n, 9).astype(np.float32) - 0.5) * np.float32(200)
    o = (rng.rand(n, 3).astype(np.float32) - 0.5) * np.float32(400)
    # aim at a point of the triangle's plane: inside, on edges, at vertices, just outside
    w = rng.rand(n, 2).astype(np.float32)
    kind = rng.randint(0, 8, n)
    w[kind == 1, 1] = 0            # on edge
    w[kind == 2] = 0               # at vertex v0
    w[kind == 3] *= np.float32(1.5)  # may leave the triangle
    flip = w.sum(1) > 1
    w[flip & (kind != 3)] = 1 - w[flip & (kind != 3)]
    w[kind == 5, 1] = 1 - w[kind == 5, 0]   # on the far edge: u + v == 1 up to rounding
    w[kind == 6] = [1, 0]                   # at v1: u == 1
    w[kind == 7] = [0, 1]                   # at v2: v == 1
    v0, v1, v2 = v[:, 0:3], v[:, 3:6], v[:, 6:9]
    p = v0 + (v1 - v0) * w[:, :1] + (v2 - v0) * w[:, 1:]
    d = p - o
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    d = d.astype(np.float32)
    # grazing / degenerate: tiny triangles make |det| approach the absolute 1e-4 threshold
    small = kind == 4
    v[small] = v[small, :3].repeat(3, 0).reshape(-1, 9) + (rng.rand(small.sum(), 9).astype(np.float32) - 0.5) * np.float32(0.05)
    return v.astype(np.float32), o.astype(np.float32), d

def bsdf_inputs(rng, n):
    def unit(k):
        v = rng.normal(size=(k, 3)).astype(np.float32)
        return (v / np.linalg.norm(v, axis=1, keepdims=True)).astype(np.float32)
    nrm = unit(n)
    wi, wo = unit(n), unit(n)
    # half of the cases: wo near the mirror / refracted direction of wi (where the smooth lobes are non-zero)
    k = n // 2
    refl = 2 * (wi[:k] * nrm[:k]).sum(1, keepdims=True) * nrm[:k] - wi[:k]
    wo[:k] = refl + rng.normal(size=(k, 3)).astype(np.float32) * np.float32(0.01) * (rng.rand(k, 1) < 0.5)
    wo[:k] /= np.linalg.norm(wo[:k], axis=1, keepdims=True)
    wl = rng.randint(0, 3, n).astype(np.int32)
    uv = rng.rand(n, 2).astype(np.float32)
    rf = rng.randint(0, 2, n).astype(np.int32)
    return wi, wo.astype(np.float32), nrm, wl, uv, rf


def box_cases(rng, n):
    c = (rng.rand(n, 3).astype(np.float32) - 0.5) * 100
    e = rng.rand(n, 3).astype(np.float32) * 30
    e[rng.rand(n) < 0.1, 1] = 0  # flat boxes (axis-aligned quads)
    b6 = np.concatenate([c - e, c + e], 1).astype(np.float32)
    o = (rng.rand(n, 3).astype(np.float32) - 0.5) * 300
    tgt = c + (rng.rand(n, 3).astype(np.float32) - 0.5) * e * np.float32(3)
    d = (tgt - o).astype(np.float32)
    d[rng.rand(n) < 0.1, 0] = 0  # axis-parallel rays: infinite inverse direction
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    return b6, o.astype(np.float32), d.astype(np.float32)


def sphere_cases(rng, n):
    c4 = np.concatenate([(rng.rand(n, 3) - 0.5) * 200, rng.rand(n, 1) * 80 + 1], 1).astype(np.float32)
    o = ((rng.rand(n, 3) - 0.5) * 400).astype(np.float32)
    inside = rng.rand(n) < 0.2
    o[inside] = c4[inside, :3] + (rng.rand(inside.sum(), 3).astype(np.float32) - 0.5) * c4[inside, 3:4] * np.float32(0.5)
    tgt = c4[:, :3] + (rng.rand(n, 3).astype(np.float32) - 0.5) * c4[:, 3:4] * np.float32(2.2)
    d = tgt - o
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    return c4, o, d.astype(np.float32)


def degenerate_rays(rng, scene_min, scene_max, n):
    """Rays whose slab products can be NaN or infinite: zero direction components, origins on axis-aligned planes
    (y = 0 is the chess floor / Cornell floor), zero-length directions (what Material::refract returns on total
    internal reflection), denormal components."""
    lo, hi = np.asarray(scene_min, np.float32), np.asarray(scene_max, np.float32)
    o = (lo + (hi - lo) * rng.rand(n, 3)).astype(np.float32)
    d = rng.normal(size=(n, 3)).astype(np.float32)
    k = rng.randint(0, 6, n)
    d[k == 0, 0] = 0
    d[k == 1, 1] = 0
    d[k == 2, 2] = 0
    d[k == 3] = d[k == 3] * np.array([1, 0, 0], np.float32)   # axis-parallel
    d[k == 4] = 0                                              # zero vector
    d[k == 5, 1] = np.float32(1e-41)                           # denormal: 1/d overflows to inf
    nz = np.linalg.norm(d, axis=1) > 0
    d[nz] /= np.linalg.norm(d[nz], axis=1, keepdims=True)
    d[k == 5, 1] = np.float32(1e-41)
    on_plane = rng.rand(n) < 0.5
    o[on_plane, 1] = 0                                         # exactly on the floor plane
    return o, d.astype(np.float32)
