"""Test support: bindings of the two CPU checkers and scene mirroring.  TEST INFRASTRUCTURE.

* ``Ref``       — oracle/_ref/libref_oracle.so: the UNMODIFIED reference sources compiled against
                  oracle/eigen_shim (see oracle/ref_harness.cpp).  Built by oracle/Makefile here (where
                  /root/reference exists); the prebuilt library travels to the GPU box.
* ``HostCheck`` — tests/hostcheck: csrc/pt_math.cuh compiled for the host, so the device arithmetic
                  can be compared with the oracle on a machine without a GPU.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import b2pt_loader  # noqa: E402

b2pt = b2pt_loader.load()

REF_LIB = os.path.join(ROOT, "oracle", "_ref", "libref_oracle.so")
HC_LIB = os.path.join(ROOT, "tests", "hostcheck", "_build", "libhostcheck.so")
REFERENCE_DIR = "/root/reference"
GOLDEN = os.path.join(ROOT, "tests", "golden")
SEED = 0x5EED0001

fp, ip, f32, i32 = b2pt.fp, b2pt.ip, b2pt.f32, b2pt.i32
c_double_p = C.POINTER(C.c_double)


def ensure_built():
    """Builds the host library, the CUDA library and both checkers when missing (no GPU needed)."""
    if not (os.path.exists(b2pt.HOST_LIB) and os.path.exists(b2pt.GPU_LIB)):
        b2pt.build()
    if not os.path.exists(HC_LIB):
        subprocess.run(["make", "-C", os.path.join(ROOT, "tests", "hostcheck")], check=True, capture_output=True)
    if not os.path.exists(REF_LIB) and os.path.isdir(os.path.join(REFERENCE_DIR, "src")):
        subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "ref"], check=True, capture_output=True)


# ---- the two CPU checkers: bindings live next to them in oracle/refbind.py ------------------------------------------
from oracle.refbind import (PTO_LIB, Ref, Restated, have_ref, pto_lib, pto_philox_block, pto_stream_uniforms, pto_tri, ref_box, ref_lib, ref_sphere, ref_tri,  # noqa: E402,F401
                            write_obj_soup)


_hc = None


def hc_lib():
    global _hc
    if _hc is None:
        L = C.CDLL(HC_LIB)
        L.hc_scene_new.restype = C.c_void_p
        L.hc_scene_new.argtypes = [C.POINTER(b2pt.SceneDesc)]
        L.hc_scene_free.argtypes = [C.c_void_p]
        _hc = L
    return _hc


class HostCheck:
    """csrc/pt_math.cuh compiled for the host, over a flattened scene."""

    def __init__(self, scene: "b2pt.HostScene"):
        self.L = hc_lib()
        self.h = C.c_void_p(self.L.hc_scene_new(C.byref(scene.desc)))
        if not self.h:
            raise RuntimeError("hostcheck rejected the scene")

    def close(self):
        if self.h:
            self.L.hc_scene_free(self.h)
            self.h = None

    def intersect(self, o, d, counts=False, binary=False):
        """Closest hits as the extend kernel finds them: the four-wide walk (pt::closest_hit4), or with binary=True the
        sibling-pair walk the shadow kernel and the degenerate rays use.  counts: (boxes tested, primitives tested)."""
        o, d = f32(o).reshape(-1, 3), f32(d).reshape(-1, 3)
        prim = np.zeros(len(o), np.int32)
        t = np.zeros(len(o), np.float64)
        cnt = (C.c_ulonglong * 2)()
        (self.L.hc_intersect if binary else self.L.hc_intersect4)(self.h, fp(o), fp(d), C.c_long(len(o)), ip(prim), t.ctypes.data_as(c_double_p), cnt)
        return (prim, t, (cnt[0], cnt[1])) if counts else (prim, t)

    def quad_stats(self):
        """(number of quads of the four-wide tree, stack entries its walk can need)."""
        self.L.hc_quad_count.restype = C.c_long
        return int(self.L.hc_quad_count(self.h)), int(self.L.hc_quad_stack_need(self.h))

    def shadow(self, o, d, dist):
        o, d, s = f32(o).reshape(-1, 3), f32(d).reshape(-1, 3), f32(dist)
        vis = np.zeros(len(o), np.int32)
        self.L.hc_shadow(self.h, fp(o), fp(d), fp(s), C.c_long(len(o)), ip(vis))
        return vis

    def shadow4(self, o, d, dist, counts=False):
        """The visibility decision by the four-wide walk of the shadow kernel (pt::light_visible4)."""
        o, d, s = f32(o).reshape(-1, 3), f32(d).reshape(-1, 3), f32(dist)
        vis = np.zeros(len(o), np.int32)
        cnt = (C.c_ulonglong * 2)()
        self.L.hc_shadow4(self.h, fp(o), fp(d), fp(s), C.c_long(len(o)), ip(vis), cnt)
        return (vis, (cnt[0], cnt[1])) if counts else vis

    def flat_intersect(self, o, d):
        """Closest hits by the treeless walk of small scenes (pt::flat_closest); None when the scene is not small."""
        o, d = f32(o).reshape(-1, 3), f32(d).reshape(-1, 3)
        prim = np.zeros(len(o), np.int32)
        t = np.zeros(len(o), np.float64)
        if not self.L.hc_flat_intersect(self.h, fp(o), fp(d), C.c_long(len(o)), ip(prim), t.ctypes.data_as(c_double_p)):
            return None
        return prim, t

    def flat_shadow(self, o, d, dist):
        o, d, s = f32(o).reshape(-1, 3), f32(d).reshape(-1, 3), f32(dist)
        vis = np.zeros(len(o), np.int32)
        if not self.L.hc_flat_shadow(self.h, fp(o), fp(d), fp(s), C.c_long(len(o)), ip(vis)):
            return None
        return vis

    def shadow_with_light_node(self, o, d, dist, light_prim):
        o, d, s, lp = f32(o).reshape(-1, 3), f32(d).reshape(-1, 3), f32(dist), i32(light_prim)
        vis = np.zeros(len(o), np.int32)
        self.L.hc_shadow_lnode(self.h, fp(o), fp(d), fp(s), ip(lp), C.c_long(len(o)), ip(vis))
        return vis

    def sample_light_node(self, u4):
        u = f32(u4).reshape(-1, 4)
        prim = np.zeros(len(u), np.int32)
        self.L.hc_sample_light_node(self.h, fp(u), C.c_long(len(u)), ip(prim))
        return prim

    def surface(self, o, d):
        o, d = f32(o).reshape(-1, 3), f32(d).reshape(-1, 3)
        co, nn, uv = np.zeros((len(o), 3), np.float32), np.zeros((len(o), 3), np.float32), np.zeros((len(o), 2), np.float32)
        self.L.hc_surface(self.h, fp(o), fp(d), C.c_long(len(o)), fp(co), fp(nn), fp(uv))
        return co, nn, uv

    def bsdf_eval(self, mat, wi, wo, n, wl, uv, rf):
        wi, wo, n, uv, wl, rf = f32(wi), f32(wo), f32(n), f32(uv), i32(wl), i32(rf)
        out = np.zeros(len(wl), np.float32)
        self.L.hc_eval(self.h, mat, fp(wi), fp(wo), fp(n), ip(wl), fp(uv), ip(rf), C.c_long(len(wl)), fp(out))
        return out

    def nee_dead(self, o, d, samples, rng):
        """For the first hit of every ray: (vertex-level verdict pt::nee_vertex_is_dead, number of `samples` random light samples
        whose summand is not known to be zero; -1 where the ray has no shaded vertex)."""
        o, d = f32(o).reshape(-1, 3), f32(d).reshape(-1, 3)
        n = len(o)
        u4 = (np.floor(rng.rand(n, samples, 4) * 16777216.0) / 16777216.0).astype(np.float32)
        v, a = np.zeros(n, np.int32), np.zeros(n, np.int32)
        self.L.hc_nee_dead(self.h, fp(o), fp(d), fp(u4), C.c_long(n), samples, ip(v), ip(a))
        return v, a

    def bsdf_eval_returns_zero(self, mat, wi, wo, n, wl, rf):
        """pt::mat_eval_returns_zero: the predicate the nee kernel uses to drop light samples whose summand is zero."""
        wi, wo, n, wl, rf = f32(wi), f32(wo), f32(n), i32(wl), i32(rf)
        out = np.zeros(len(wl), np.int32)
        self.L.hc_eval_returns_zero(self.h, mat, fp(wi), fp(wo), fp(n), ip(wl), ip(rf), C.c_long(len(wl)), ip(out))
        return out

    def bsdf_pdf(self, mat, wi, wo, n, wl, rf):
        wi, wo, n, wl, rf = f32(wi), f32(wo), f32(n), i32(wl), i32(rf)
        out = np.zeros(len(wl), np.float32)
        self.L.hc_pdf(self.h, mat, fp(wi), fp(wo), fp(n), ip(wl), ip(rf), C.c_long(len(wl)), fp(out))
        return out

    def fresnel(self, mat, I, n, wl):
        I, n, wl = f32(I), f32(n), i32(wl)
        out = np.zeros(len(wl), np.float32)
        self.L.hc_fresnel(self.h, mat, fp(I), fp(n), ip(wl), C.c_long(len(wl)), fp(out))
        return out

    def refract(self, mat, I, n, wl):
        I, n, wl = f32(I), f32(n), i32(wl)
        out = np.zeros((len(wl), 3), np.float32)
        self.L.hc_refract(self.h, mat, fp(I), fp(n), ip(wl), C.c_long(len(wl)), fp(out))
        return out

    def reflect(self, I, n):
        I, n = f32(I).reshape(-1, 3), f32(n).reshape(-1, 3)
        out = np.zeros((len(I), 3), np.float32)
        self.L.hc_reflect(fp(I), fp(n), C.c_long(len(I)), fp(out))
        return out

    def material_sample(self, mat, n, u2):
        n, u2 = f32(n).reshape(-1, 3), f32(u2).reshape(-1, 2)
        out = np.zeros((len(n), 3), np.float32)
        self.L.hc_material_sample(self.h, mat, fp(n), fp(u2), C.c_long(len(n)), fp(out))
        return out

    def env(self, d):
        d = f32(d).reshape(-1, 3)
        out = np.zeros((len(d), 3), np.float32)
        self.L.hc_env(self.h, fp(d), C.c_long(len(d)), fp(out))
        return out

    def sample_light(self, u4):
        u = f32(u4).reshape(-1, 4)
        co, nn, em = (np.zeros((len(u), 3), np.float32) for _ in range(3))
        pdf = np.zeros(len(u), np.float32)
        self.L.hc_sample_light(self.h, fp(u), C.c_long(len(u)), fp(co), fp(nn), fp(em), fp(pdf))
        return co, nn, em, pdf


def hc_tri(v9, o, d):
    v, o, d = f32(v9).reshape(-1, 9), f32(o).reshape(-1, 3), f32(d).reshape(-1, 3)
    hit = np.zeros(len(o), np.int32)
    t = np.zeros(len(o), np.float64)
    hc_lib().hc_tri(fp(v), fp(o), fp(d), C.c_long(len(o)), ip(hit), t.ctypes.data_as(c_double_p))
    return hit, t


def hc_box(b6, o, d):
    b, o, d = f32(b6).reshape(-1, 6), f32(o).reshape(-1, 3), f32(d).reshape(-1, 3)
    hit = np.zeros(len(o), np.int32)
    hc_lib().hc_box(fp(b), fp(o), fp(d), C.c_long(len(o)), ip(hit))
    return hit


def hc_sphere(c4, o, d):
    c, o, d = f32(c4).reshape(-1, 4), f32(o).reshape(-1, 3), f32(d).reshape(-1, 3)
    hit = np.zeros(len(o), np.int32)
    t = np.zeros(len(o), np.float64)
    hc_lib().hc_sphere(fp(c), fp(o), fp(d), C.c_long(len(o)), ip(hit), t.ctypes.data_as(c_double_p))
    return hit, t


def hc_camera_rays(cam, pixels, sample_begin, sample_count, seed=SEED):
    px = i32(pixels)
    n = len(px) * sample_count
    o, d = np.zeros((n, 3), np.float32), np.zeros((n, 3), np.float32)
    hc_lib().hc_camera_rays(C.byref(cam), ip(px), len(px), sample_begin, sample_count, C.c_ulonglong(seed), fp(o), fp(d))
    return o, d


def hc_stream_uniforms(seed, pixel, sample, tag, dim_begin, count):
    out = np.zeros(count, np.float32)
    hc_lib().hc_stream_uniforms(C.c_ulonglong(seed), pixel, sample, tag, dim_begin, count, fp(out))
    return out


# ---- scenes (the builders live in the package: b2pt.scenes) ----------------------------------------------------------
from b2pt import scenes as _scenes  # noqa: E402

synthetic_sky = _scenes.synthetic_sky
write_sky_png = _scenes.write_sky_png
CHESS_CONF = _scenes.CHESS_CONF
chess_conf_text = _scenes.chess_conf_text


# ---- oracle/pt_oracle.c: the plain-C restatement ----------------------------------------------------------------