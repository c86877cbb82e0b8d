"""oracle/refbind.py — TEST INFRASTRUCTURE: ctypes bindings of the two CPU checkers.

* ``Ref``      — oracle/_ref/libref_oracle.so: the UNMODIFIED reference sources compiled against oracle/eigen_shim
                 (oracle/ref_harness.cpp).  Built by oracle/Makefile where /root/reference exists; the prebuilt library
                 travels to the GPU box.
* ``Restated`` — oracle/libpt_oracle.so: the plain-C restatement (oracle/pt_oracle.c).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / RMSE / --impl reference legs import this module, and only
as the checker; nothing in the package does.
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
REFERENCE_DIR = "/root/reference"
SEED = 0x5EED0001

fp, ip, f32, i32 = b2pt.fp, b2pt.ip, b2pt.f32, b2pt.i32
c_double_p = C.POINTER(C.c_double)


def have_ref():
    return os.path.exists(REF_LIB)


_ref = None


def ref_lib():
    global _ref
    if _ref is None:
        L = C.CDLL(REF_LIB)
        L.ref_describe.restype = C.c_char_p
        L.ref_scene_new.restype = C.c_void_p
        L.ref_scene_free.argtypes = [C.c_void_p]
        L.ref_add_material.argtypes = [C.c_void_p, C.c_int, b2pt.c_float_p, C.c_float, C.c_float, C.c_float, b2pt.c_float_p, C.c_int]
        L.ref_material_defaults.argtypes = [C.c_int, b2pt.c_float_p, b2pt.c_float_p, b2pt.c_float_p, b2pt.c_int_p]
        L.ref_add_mesh.argtypes = [C.c_void_p, C.c_char_p, C.c_int, b2pt.c_float_p, C.c_float]
        L.ref_add_sphere.argtypes = [C.c_void_p, b2pt.c_float_p, C.c_float, C.c_int]
        L.ref_set_rr.argtypes = [C.c_void_p, C.c_float]
        L.ref_set_shadow.argtypes = [C.c_void_p, C.c_int]
        L.ref_set_n_dir.argtypes = [C.c_void_p, C.c_int]
        L.ref_set_background.argtypes = [C.c_void_p, b2pt.c_float_p]
        L.ref_load_env.argtypes = [C.c_void_p, C.c_char_p]
        L.ref_set_camera.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_float, b2pt.c_float_p, b2pt.c_float_p, b2pt.c_float_p, C.c_int,
                                     C.c_float, C.c_float]
        L.ref_build.argtypes = [C.c_void_p]
        L.ref_mesh_triangle_count.argtypes = [C.c_void_p, C.c_int]
        L.ref_mesh_triangles.argtypes = [C.c_void_p, C.c_int, b2pt.c_float_p, b2pt.c_float_p, b2pt.c_float_p]
        L.ref_object_area.argtypes = [C.c_void_p, C.c_int]
        L.ref_object_area.restype = C.c_float
        L.ref_object_bounds.argtypes = [C.c_void_p, C.c_int, b2pt.c_float_p]
        L.ref_camera_orientation.argtypes = [C.c_void_p, b2pt.c_float_p]
        L.ref_camera_scale.argtypes = [C.c_void_p]
        L.ref_camera_scale.restype = C.c_float
        L.ref_intersect.argtypes = [C.c_void_p, b2pt.c_float_p, b2pt.c_float_p, C.c_int, b2pt.c_int_p, b2pt.c_int_p, c_double_p,
                                    b2pt.c_float_p, b2pt.c_float_p, b2pt.c_float_p]
        L.ref_tri_intersect.argtypes = [b2pt.c_float_p, b2pt.c_float_p, b2pt.c_float_p, C.c_int, b2pt.c_int_p, c_double_p, b2pt.c_float_p]
        L.ref_box_intersect.argtypes = [b2pt.c_float_p, b2pt.c_float_p, b2pt.c_float_p, C.c_int, b2pt.c_int_p]
        L.ref_sphere_intersect.argtypes = [b2pt.c_float_p, b2pt.c_float_p, b2pt.c_float_p, C.c_int, b2pt.c_int_p, c_double_p, b2pt.c_float_p,
                                           b2pt.c_float_p]
        L.ref_bsdf_eval.argtypes = [C.c_void_p, C.c_int, b2pt.c_float_p, b2pt.c_float_p, b2pt.c_float_p, b2pt.c_int_p, b2pt.c_float_p,
                                    b2pt.c_int_p, C.c_int, b2pt.c_float_p]
        L.ref_bsdf_pdf.argtypes = [C.c_void_p, C.c_int, b2pt.c_float_p, b2pt.c_float_p, b2pt.c_float_p, b2pt.c_int_p, b2pt.c_int_p, C.c_int,
                                   b2pt.c_float_p]
        L.ref_fresnel.argtypes = [C.c_void_p, C.c_int, b2pt.c_float_p, b2pt.c_float_p, b2pt.c_int_p, C.c_int, b2pt.c_float_p]
        L.ref_refract.argtypes = [C.c_void_p, C.c_int, b2pt.c_float_p, b2pt.c_float_p, b2pt.c_int_p, C.c_int, b2pt.c_float_p]
        L.ref_reflect.argtypes = [C.c_void_p, C.c_int, b2pt.c_float_p, b2pt.c_float_p, C.c_int, b2pt.c_float_p]
        L.ref_material_sample.argtypes = [C.c_void_p, C.c_int, b2pt.c_float_p, b2pt.c_float_p, b2pt.c_float_p, C.c_int, b2pt.c_float_p]
        L.ref_material_ior.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.ref_material_ior.restype = C.c_float
        L.ref_material_has_emission.argtypes = [C.c_void_p, C.c_int]
        L.ref_sample_env.argtypes = [C.c_void_p, b2pt.c_float_p, C.c_int, b2pt.c_float_p]
        L.ref_sample_light.argtypes = [C.c_void_p, b2pt.c_float_p, C.c_int, b2pt.c_float_p, b2pt.c_float_p, b2pt.c_float_p, b2pt.c_float_p]
        L.ref_uniform_from_script.argtypes = [C.c_float]
        L.ref_uniform_from_script.restype = C.c_float
        L.ref_cast_ray_scripted.argtypes = [C.c_void_p, b2pt.c_float_p, b2pt.c_float_p, b2pt.c_int_p, b2pt.c_float_p, C.c_int, C.c_int,
                                            b2pt.c_float_p, b2pt.c_int_p]
        L.ref_camera_rays_philox.argtypes = [C.c_void_p, b2pt.c_int_p, C.c_int, C.c_int, C.c_int, C.c_uint32, C.c_uint32, b2pt.c_float_p,
                                             b2pt.c_float_p]
        L.ref_render_samples_philox.argtypes = [C.c_void_p, b2pt.c_int_p, C.c_int, C.c_int, C.c_int, C.c_uint32, C.c_uint32, b2pt.c_float_p,
                                                C.POINTER(C.c_ulonglong)]
        L.ref_render_frame_philox.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_uint32, C.c_uint32, C.c_int, b2pt.c_float_p]
        L.ref_render_real.argtypes = [C.c_void_p, C.c_int, C.c_char_p]
        L.ref_render_frame_free.argtypes = [C.c_void_p, C.c_int, C.c_uint32, C.c_int, b2pt.c_float_p, b2pt.c_float_p]
        L.ref_render_samples_free.argtypes = [C.c_void_p, b2pt.c_int_p, C.c_int, C.c_int, C.c_uint32, C.c_int, b2pt.c_float_p]
        L.ref_render_samples_philox_split.argtypes = [C.c_void_p, b2pt.c_int_p, C.c_int, C.c_int, C.c_int, C.c_uint32, C.c_uint32, b2pt.c_float_p]
        L.ref_set_free_engine.argtypes = [C.c_int, C.c_uint32]
        L.ref_uniform_from_word.argtypes = [C.c_uint32]
        L.ref_uniform_from_word.restype = C.c_float
        _ref = L
    return _ref


def write_obj_soup(path, v9, uv6):
    """Triangle-soup OBJ whose %.9g literals read back to the same float32 values."""
    v = np.asarray(v9, np.float32).reshape(-1, 3)
    t = np.asarray(uv6, np.float32).reshape(-1, 2)
    with open(path, "w") as f:
        f.write("".join("v %.9g %.9g %.9g\n" % tuple(r) for r in v.tolist()))
        f.write("".join("vt %.9g %.9g\n" % tuple(r) for r in t.tolist()))
        n = len(v)
        f.write("".join("f %d/%d %d/%d %d/%d\n" % (i + 1, i + 1, i + 2, i + 2, i + 3, i + 3) for i in range(0, n - 2, 3)))


class Ref:
    """A scene held by the real reference code, mirrored from a HostScene."""

    def __init__(self, scene: "b2pt.HostScene", env_png: str | None = None):
        self.L = ref_lib()
        self.h = C.c_void_p(self.L.ref_scene_new())
        self.scene = scene
        self._tmp = tempfile.TemporaryDirectory(prefix="b2pt_ref_")
        for _, m in scene.materials():
            em, rf = f32(list(m.emission)), f32(list(m.base_reflectance))
            self.L.ref_add_material(self.h, m.type, fp(em), m.ior_a, m.ior_b, m.roughness, fp(rf), m.textured)
        zero = f32([0, 0, 0])
        for k in range(scene.n_objects):
            o = scene.object_info(k)
            if o["kind"] == "sphere":
                c = f32(o["center"])
                self.L.ref_add_sphere(self.h, fp(c), o["radius"], o["material"])
            else:
                path = os.path.join(self._tmp.name, f"obj{k}.obj")
                write_obj_soup(path, o["v9"], o["uv6"])
                self.L.ref_add_mesh(self.h, path.encode(), o["material"], fp(zero), 1.0)
        d = scene.desc
        self.L.ref_set_rr(self.h, d.rr_rate)
        self.L.ref_set_shadow(self.h, d.enable_shadow)
        self.L.ref_set_n_dir(self.h, d.n_dir_sample)
        bg = f32(list(d.background))
        self.L.ref_set_background(self.h, fp(bg))
        if d.use_env_map:
            if not env_png:
                raise ValueError("scene uses an env map: pass the PNG it was loaded from")
            if self.L.ref_load_env(self.h, env_png.encode()) != 1:
                raise RuntimeError("reference failed to load env map")
        fov, pos, tgt, up = scene.camera_params()
        cam = scene.camera
        self.L.ref_set_camera(self.h, cam.width, cam.height, fov, fp(pos), fp(tgt), fp(up), cam.use_dof, cam.focal_distance, cam.aperture_radius)
        self.L.ref_build(self.h)
        self._prim_tables = None

    def close(self):
        if self.h:
            self.L.ref_scene_free(self.h)
            self.h = None
            self._tmp.cleanup()

    def set_params(self, rr_rate=None, enable_shadow=None, n_dir=None):
        if rr_rate is not None:
            self.L.ref_set_rr(self.h, rr_rate)
        if enable_shadow is not None:
            self.L.ref_set_shadow(self.h, int(enable_shadow))
        if n_dir is not None:
            self.L.ref_set_n_dir(self.h, n_dir)

    def intersect(self, o, d):
        """Scene::intersect: (prim id in the flattened numbering, -1 miss; distance; coords; normal; uv)."""
        o, d = f32(o).reshape(-1, 3), f32(d).reshape(-1, 3)
        n = len(o)
        obj, tri = np.zeros(n, np.int32), np.zeros(n, np.int32)
        t = np.zeros(n, np.float64)
        co, nn, uv = np.zeros((n, 3), np.float32), np.zeros((n, 3), np.float32), np.zeros((n, 2), np.float32)
        self.L.ref_intersect(self.h, fp(o), fp(d), n, ip(obj), ip(tri), t.ctypes.data_as(c_double_p), fp(co), fp(nn), fp(uv))
        return self.to_prim(obj, tri), t, co, nn, uv

    def to_prim(self, obj, tri):
        """(object, face) -> primitive id of the flattened scene (depth-first leaf order)."""
        if self._prim_tables is None:
            po, pf = self.scene.prim_origins()
            key = {}
            for pid, (a, b) in enumerate(zip(po.tolist(), pf.tolist())):
                key[(a, b)] = pid
            self._prim_tables = key
        out = np.full(len(obj), -1, np.int32)
        for i, (a, b) in enumerate(zip(obj.tolist(), tri.tolist())):
            if a >= 0:
                out[i] = self._prim_tables[(a, b)]
        return out

    def bsdf_eval(self, mat, wi, wo, n, wl, uv, rf):
        wi, wo, n, uv, wl, rf = f32(wi), f32(wo), f32(n), f32(uv), i32(wl), i32(rf)
        out = np.zeros(len(wl), np.float32)
        self.L.ref_bsdf_eval(self.h, mat, fp(wi), fp(wo), fp(n), ip(wl), fp(uv), ip(rf), len(wl), fp(out))
        return out

    def bsdf_pdf(self, mat, wi, wo, n, wl, rf):
        wi, wo, n, wl, rf = f32(wi), f32(wo), f32(n), i32(wl), i32(rf)
        out = np.zeros(len(wl), np.float32)
        self.L.ref_bsdf_pdf(self.h, mat, fp(wi), fp(wo), fp(n), ip(wl), ip(rf), len(wl), fp(out))
        return out

    def fresnel(self, mat, I, n, wl):
        I, n, wl = f32(I), f32(n), i32(wl)
        out = np.zeros(len(wl), np.float32)
        self.L.ref_fresnel(self.h, mat, fp(I), fp(n), ip(wl), len(wl), fp(out))
        return out

    def refract(self, mat, I, n, wl):
        I, n, wl = f32(I), f32(n), i32(wl)
        out = np.zeros((len(wl), 3), np.float32)
        self.L.ref_refract(self.h, mat, fp(I), fp(n), ip(wl), len(wl), fp(out))
        return out

    def reflect(self, mat, I, n):
        I, n = f32(I).reshape(-1, 3), f32(n).reshape(-1, 3)
        out = np.zeros((len(I), 3), np.float32)
        self.L.ref_reflect(self.h, mat, fp(I), fp(n), len(I), fp(out))
        return out

    def material_sample(self, mat, wo, n, u2):
        wo, n, u2 = f32(wo).reshape(-1, 3), f32(n).reshape(-1, 3), f32(u2).reshape(-1, 2)
        out = np.zeros((len(n), 3), np.float32)
        self.L.ref_material_sample(self.h, mat, fp(wo), fp(n), fp(u2), len(n), fp(out))
        return out

    def sample_env(self, d):
        d = f32(d).reshape(-1, 3)
        out = np.zeros((len(d), 3), np.float32)
        self.L.ref_sample_env(self.h, fp(d), len(d), fp(out))
        return out

    def sample_light(self, u4):
        u = f32(u4).reshape(-1, 4)
        co, nn, em = (np.zeros((len(u), 3), np.float32) for _ in range(3))
        pdf = np.zeros(len(u), np.float32)
        self.L.ref_sample_light(self.h, fp(u), len(u), fp(co), fp(nn), fp(em), fp(pdf))
        return co, nn, em, pdf

    def camera_rays(self, pixels, sample_begin, sample_count, seed=SEED):
        px = i32(pixels)
        n = len(px) * sample_count
        o, d = np.zeros((n, 3), np.float32), np.zeros((n, 3), np.float32)
        self.L.ref_camera_rays_philox(self.h, ip(px), len(px), sample_begin, sample_count, seed & 0xFFFFFFFF, seed >> 32, fp(o), fp(d))
        return o, d

    def render_samples(self, pixels, sample_begin, sample_count, seed=SEED):
        """The three castRay values of Renderer.cpp:77-79 per (pixel, sample) on the Philox sample streams."""
        px = i32(pixels)
        out = np.zeros((len(px), sample_count, 3), np.float32)
        draws = C.c_ulonglong()
        self.L.ref_render_samples_philox(self.h, ip(px), len(px), sample_begin, sample_count, seed & 0xFFFFFFFF, seed >> 32, fp(out),
                                         C.byref(draws))
        return out

    def render_free(self, spp, seed=1, threads=0):
        """The reference's pixel loop on its own sampling scheme (free-running mt19937 per thread, independent draws for the
        three castRay calls): (mean frame, per-pixel variance of one sample), both [H, W, 3]."""
        cam = self.scene.camera
        mean = np.zeros((cam.height, cam.width, 3), np.float32)
        m2 = np.zeros((cam.height, cam.width, 3), np.float32)
        self.L.ref_render_frame_free(self.h, spp, seed & 0xFFFFFFFF, threads, fp(mean), fp(m2))
        var = np.maximum(m2.astype(np.float64) - mean.astype(np.float64) ** 2, 0.0)
        return mean, var

    def render_samples_split(self, pixels, sample_begin, sample_count, seed=SEED):
        """castRay per (pixel, sample, wavelength) on keyed Philox streams with R, G and B reading THREE DIFFERENT streams
        (tags 0, 2, 3): the oracle of B2PT_FLAG_INDEPENDENT_WAVELENGTHS."""
        px = i32(pixels)
        out = np.zeros((len(px), sample_count, 3), np.float32)
        self.L.ref_render_samples_philox_split(self.h, ip(px), len(px), sample_begin, sample_count, seed & 0xFFFFFFFF, seed >> 32, fp(out))
        return out

    def render_samples_free(self, pixels, sample_count, seed=1, threads=0):
        """Per-sample values [pixels, sample_count, 3] on the reference's own sampling scheme (free-running mt19937 per thread)."""
        px = i32(pixels)
        out = np.zeros((len(px), sample_count, 3), np.float32)
        self.L.ref_render_samples_free(self.h, ip(px), len(px), sample_count, seed & 0xFFFFFFFF, threads, fp(out))
        return out

    def render_frame(self, sample_begin, sample_count, spp_total, seed=SEED, threads=0, fb=None):
        cam = self.scene.camera
        if fb is None:
            fb = np.zeros((cam.height, cam.width, 3), np.float32)
        self.L.ref_render_frame_philox(self.h, sample_begin, sample_count, spp_total, seed & 0xFFFFFFFF, seed >> 32, threads, fp(fb))
        return fb



def ref_tri(v9, o, d):
    v, o, d = f32(v9).reshape(-1, 9), f32(o).reshape(-1, 3), f32(d).reshape(-1, 3)
    hit = np.zeros(len(o), np.int32)
    t = np.zeros(len(o), np.float64)
    ref_lib().ref_tri_intersect(fp(v), fp(o), fp(d), len(o), ip(hit), t.ctypes.data_as(c_double_p), None)
    return hit, t


def ref_box(b6, o, d):
    b, o, d = f32(b6).reshape(-1, 6), f32(o).reshape(-1, 3), f32(d).reshape(-1, 3)
    hit = np.zeros(len(o), np.int32)
    ref_lib().ref_box_intersect(fp(b), fp(o), fp(d), len(o), ip(hit))
    return hit


def ref_sphere(c4, o, d):
    c, o, d = f32(c4).reshape(-1, 4), f32(o).reshape(-1, 3), f32(d).reshape(-1, 3)
    hit = np.zeros(len(o), np.int32)
    t = np.zeros(len(o), np.float64)
    co, nn = np.zeros((len(o), 3), np.float32), np.zeros((len(o), 3), np.float32)
    ref_lib().ref_sphere_intersect(fp(c), fp(o), fp(d), len(o), ip(hit), t.ctypes.data_as(c_double_p), fp(co), fp(nn))
    return hit, t, co, nn



PTO_LIB = os.path.join(ROOT, "oracle", "libpt_oracle.so")
_pto = None


def pto_lib():
    global _pto
    if _pto is None:
        if not os.path.exists(PTO_LIB):
            subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "libpt_oracle.so"], check=True, capture_output=True)
        L = C.CDLL(PTO_LIB)
        L.pto_scene_new.restype = C.c_void_p
        L.pto_scene_new.argtypes = [C.POINTER(b2pt.SceneDesc)]
        L.pto_scene_free.argtypes = [C.c_void_p]
        L.pto_describe.restype = C.c_char_p
        _pto = L
    return _pto


class Restated:
    """oracle/pt_oracle.c over a flattened scene (recursive castRay, exhaustive BVH walk)."""

    def __init__(self, scene: "b2pt.HostScene"):
        self.L = pto_lib()
        self.scene = scene
        self.h = C.c_void_p(self.L.pto_scene_new(C.byref(scene.desc)))

    def close(self):
        if self.h:
            self.L.pto_scene_free(self.h)
            self.h = None

    def intersect(self, o, d):
        o, d = f32(o).reshape(-1, 3), f32(d).reshape(-1, 3)
        prim = np.zeros(len(o), np.int32)
        t = np.zeros(len(o), np.float64)
        self.L.pto_intersect(self.h, fp(o), fp(d), C.c_long(len(o)), ip(prim), t.ctypes.data_as(c_double_p))
        return prim, t

    def bsdf_eval(self, mat, wi, wo, n, wl, uv, rf):
        wi, wo, n, uv, wl, rf = f32(wi), f32(wo), f32(n), f32(uv), i32(wl), i32(rf)
        out = np.zeros(len(wl), np.float32)
        self.L.pto_bsdf_eval(self.h, mat, fp(wi), fp(wo), fp(n), ip(wl), fp(uv), ip(rf), C.c_long(len(wl)), fp(out))
        return out

    def bsdf_pdf(self, mat, wi, wo, n, wl, rf):
        wi, wo, n, wl, rf = f32(wi), f32(wo), f32(n), i32(wl), i32(rf)
        out = np.zeros(len(wl), np.float32)
        self.L.pto_bsdf_pdf(self.h, mat, fp(wi), fp(wo), fp(n), ip(wl), ip(rf), C.c_long(len(wl)), fp(out))
        return out

    def sample_env(self, d):
        d = f32(d).reshape(-1, 3)
        out = np.zeros((len(d), 3), np.float32)
        self.L.pto_sample_env(self.h, fp(d), C.c_long(len(d)), fp(out))
        return out

    def sample_light(self, u4):
        u = f32(u4).reshape(-1, 4)
        co, nn, em = (np.zeros((len(u), 3), np.float32) for _ in range(3))
        pdf = np.zeros(len(u), np.float32)
        self.L.pto_sample_light(self.h, fp(u), C.c_long(len(u)), fp(co), fp(nn), fp(em), fp(pdf))
        return co, nn, em, pdf

    def camera_rays(self, pixels, sample_begin, sample_count, seed=SEED):
        px = i32(pixels)
        n = len(px) * sample_count
        o, d = np.zeros((n, 3), np.float32), np.zeros((n, 3), np.float32)
        cam = self.scene.camera
        self.L.pto_camera_rays(C.byref(cam), ip(px), len(px), sample_begin, sample_count, C.c_uint64(seed), fp(o), fp(d))
        return o, d

    def render_samples(self, pixels, sample_begin, sample_count, seed=SEED):
        px = i32(pixels)
        out = np.zeros((len(px), sample_count, 3), np.float32)
        cam = self.scene.camera
        self.L.pto_render_samples(self.h, C.byref(cam), ip(px), len(px), sample_begin, sample_count, C.c_uint64(seed), fp(out))
        return out

    def render_samples_counted(self, pixels, sample_begin, sample_count, seed=SEED):
        """(per-sample radiance, rays the reference algorithm needs for these paths — SURVEY 8d — and vertices it shades)."""
        px = i32(pixels)
        out = np.zeros((len(px), sample_count, 3), np.float32)
        cam = self.scene.camera
        rays, verts = C.c_ulonglong(), C.c_ulonglong()
        self.L.pto_render_samples_counted(self.h, C.byref(cam), ip(px), len(px), sample_begin, sample_count, C.c_uint64(seed), fp(out),
                                          C.byref(rays), C.byref(verts))
        return out, rays.value, verts.value

    def render_frame(self, sample_begin, sample_count, spp_total, seed=SEED, fb=None):
        cam = self.scene.camera
        if fb is None:
            fb = np.zeros((cam.height, cam.width, 3), np.float32)
        self.L.pto_render_frame(self.h, C.byref(cam), sample_begin, sample_count, spp_total, C.c_uint64(seed), fp(fb))
        return fb


def pto_stream_uniforms(seed, pixel, sample, tag, dim_begin, count):
    """Uniforms of a sample stream from the oracle's own Philox (oracle/b2pt_portable.h)."""
    out = np.zeros(count, np.float32)
    pto_lib().pto_stream_uniforms(C.c_uint64(seed), C.c_uint32(pixel), C.c_uint32(sample), C.c_uint32(tag), C.c_uint32(dim_begin), count, fp(out))
    return out


def pto_philox_block(ctr, key):
    c, k, o = (C.c_uint32 * 4)(*ctr), (C.c_uint32 * 2)(*key), (C.c_uint32 * 4)()
    pto_lib().pto_philox_block(c, k, o)
    return [int(x) for x in o]


def pto_tri(v9, o, d):
    v, o, d = f32(v9).reshape(-1, 9), f32(o).reshape(-1, 3), f32(d).reshape(-1, 3)
    hit = np.zeros(len(o), np.int32)
    t = np.zeros(len(o), np.float64)
    pto_lib().pto_tri_intersect(fp(v), fp(o), fp(d), C.c_long(len(o)), ip(hit), t.ctypes.data_as(c_double_p))
    return hit, t

