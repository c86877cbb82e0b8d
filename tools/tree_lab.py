"""Tree experiments on the CPU (TEST INFRASTRUCTURE; uses tests/hostcheck = csrc/pt_math.cuh compiled for the host).

Counts the sibling-pair nodes and primitive tests per ray of the library's traversal for three ray populations of the
chess scene — camera rays, continuation rays (mirror directions off the first hit) and occluder searches towards light
samples — for a set of builder options (bins, face weights of the area measure).  Any tree over the reference's leaves
returns identical hits (csrc/pt_build.hpp), so these numbers are the only thing a builder changes.

    python tools/tree_lab.py [n_pixels]
"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import scenes  # noqa: E402
import support as S  # noqa: E402


class Lab(S.HostCheck):
    def __init__(self, scene, bins=16, w=(1, 1, 1)):
        self.L = S.hc_lib()
        self.L.hc_scene_new_opts.restype = C.c_void_p
        self.L.hc_scene_new_opts.argtypes = [C.POINTER(S.b2pt.SceneDesc), C.c_int, C.c_float, C.c_float, C.c_float]
        self.h = C.c_void_p(self.L.hc_scene_new_opts(C.byref(scene.desc), bins, *[float(x) for x in w]))

    def occluders(self, o, d, dist, wide=False):
        o, d, s = S.f32(o).reshape(-1, 3), S.f32(d).reshape(-1, 3), S.f32(dist)
        vis = np.zeros(len(o), np.int32)
        cnt = (C.c_ulonglong * 2)()
        (self.L.hc_occluder_counts4 if wide else self.L.hc_occluder_counts)(self.h, S.fp(o), S.fp(d), S.fp(s), C.c_long(len(o)), S.ip(vis), cnt)
        return vis, (cnt[0], cnt[1])

    def intersect4(self, o, d):
        o, d = S.f32(o).reshape(-1, 3), S.f32(d).reshape(-1, 3)
        prim = np.zeros(len(o), np.int32)
        t = np.zeros(len(o), np.float64)
        cnt = (C.c_ulonglong * 2)()
        self.L.hc_intersect4(self.h, S.fp(o), S.fp(d), C.c_long(len(o)), S.ip(prim), t.ctypes.data_as(S.c_double_p), cnt)
        return prim, t, (cnt[0], cnt[1])


def populations(sc, n_pixels, seed=7):
    rng = np.random.RandomState(seed)
    cam = sc.camera
    px = rng.choice(cam.width * cam.height, n_pixels, replace=False).astype(np.int32)
    o, d = S.hc_camera_rays(cam, px, 0, 1)
    hc = S.HostCheck(sc)
    prim, t = hc.intersect(o, d)
    co, nn, _ = hc.surface(o, d)
    hit = prim >= 0
    p = (co[hit] + nn[hit] * np.float32(1e-4)).astype(np.float32)
    dn = (d[hit] * nn[hit]).sum(1, keepdims=True)
    refl = (d[hit] - 2 * dn * nn[hit]).astype(np.float32)
    u4 = (np.floor(rng.rand(len(p), 4) * 16777216.0) / 16777216.0).astype(np.float32)
    lprim = hc.sample_light_node(u4)
    # light sample positions: uniform points of the sampled light triangle are enough for a direction population
    v0 = np.ctypeslib.as_array(sc.desc.prim_v0, shape=(sc.desc.n_prims, 4))
    e1 = np.ctypeslib.as_array(sc.desc.prim_e1, shape=(sc.desc.n_prims, 4))
    e2 = np.ctypeslib.as_array(sc.desc.prim_e2, shape=(sc.desc.n_prims, 4))
    ln_prim = np.ctypeslib.as_array(sc.desc.light_node_prim, shape=(sc.desc.n_light_nodes,))
    tri = ln_prim[lprim]
    a, b = np.sqrt(u4[:, 2:3]), u4[:, 3:4]
    lp = v0[tri, :3] + (a * (1 - b)) * 0 + e1[tri, :3] * (a * b) * 0  # placeholder, replaced below
    r1, r2 = rng.rand(len(p), 1).astype(np.float32), rng.rand(len(p), 1).astype(np.float32)
    flip = (r1 + r2) > 1
    r1, r2 = np.where(flip, 1 - r1, r1), np.where(flip, 1 - r2, r2)
    lp = (v0[tri, :3] + e1[tri, :3] * r1 + e2[tri, :3] * r2).astype(np.float32)
    w = lp - p
    dist = np.sqrt((w * w).sum(1)).astype(np.float32)
    ws = (w / dist[:, None]).astype(np.float32)
    facing = (ws * nn[hit]).sum(1) > 0  # the others are rejected by the window test in practice; keep all for the count anyway
    hc.close()
    return (o, d), (p, refl), (p[facing], ws[facing], dist[facing])


def main():
    n_pixels = int(sys.argv[1]) if len(sys.argv) > 1 else 40000
    S.ensure_built()
    sc, _ = scenes.chess(1920, 1080, dof=True, sky=False)
    cam, cont, sh = populations(sc, n_pixels)
    print(f"{len(cam[0])} camera rays, {len(cont[0])} continuation rays, {len(sh[0])} occluder searches")
    base = None
    variants = [("sah16 (current)", 16, (1, 1, 1)), ("sah32", 32, (1, 1, 1)), ("sah64", 64, (1, 1, 1)),
                ("w(1,2,1)", 16, (1, 2, 1)), ("w(1,4,1)", 16, (1, 4, 1)), ("w(1,8,1)", 16, (1, 8, 1)), ("w(1,16,1)", 16, (1, 16, 1)),
                ("w(2,1,2)", 16, (2, 1, 2)), ("w(1,1,2)", 16, (1, 1, 2)), ("w(1,1,4)", 16, (1, 1, 4))]
    for name, bins, w in variants:
        lab = Lab(sc, bins, w)
        _, _, c1 = lab.intersect(*cam, counts=True)
        _, _, c2 = lab.intersect(*cont, counts=True)
        vis, c3 = lab.occluders(*sh)
        row = [c1[0] / len(cam[0]), c1[1] / len(cam[0]), c2[0] / len(cont[0]), c2[1] / len(cont[0]), c3[0] / len(sh[0]), c3[1] / len(sh[0])]
        print(f"{name:18s} camera {row[0]:6.2f} nodes {row[1]:5.2f} prims | continuation {row[2]:6.2f} / {row[3]:5.2f} | occluder {row[4]:6.2f} / {row[5]:5.2f}"
              f" | visible {vis.mean():.3f}")
        if name.startswith("sah16"):
            q1, q2 = lab.intersect4(*cam), lab.intersect4(*cont)
            b1, b2 = lab.intersect(*cam), lab.intersect(*cont)
            same = (q1[0] == b1[0]).all() and (q1[1].view(np.uint64) == b1[1].view(np.uint64)).all() and (q2[0] == b2[0]).all() and \
                (q2[1].view(np.uint64) == b2[1].view(np.uint64)).all()
            vis4, c4 = lab.occluders(*sh, wide=True)
            same = same and (vis4 == vis).all()
            print(f"{'  four-wide':18s} camera {q1[2][0] / 4 / len(cam[0]):6.2f} steps {q1[2][1] / len(cam[0]):5.2f} prims | continuation "
                  f"{q2[2][0] / 4 / len(cont[0]):6.2f} / {q2[2][1] / len(cont[0]):5.2f} | occluder {c4[0] / 4 / len(sh[0]):6.2f} / {c4[1] / len(sh[0]):5.2f}"
                  f" | identical results: {same}; binary steps = nodes / 2; quad stack need {lab.L.hc_quad_stack_need(lab.h)}")
        lab.close()
    sc.close()


if __name__ == "__main__":
    main()
