"""B200 wavefront path tracer — Python bindings over the two C ABIs.

* ``libb2pt_host.so`` (include/b2pt_host.h): CPU scene assembler mirroring the reference's
  ``main()`` up to ``scene.buildBVH()`` (src/main.cpp:19-330).
* ``libb2pt.so`` (include/b2pt.h): the CUDA hot path replacing ``Renderer::Render``'s pixel loop
  (src/Renderer.cpp:36-92) and everything below it.  There is no CPU fallback: creating a
  :class:`Context` without the library or without a GPU raises.

The package directory name is not an importable identifier; load it with ``b2pt_loader.load()``
at the repository root (tests, bench.py and __graft_entry__.py do).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG_DIR)
ASSET_DIR = os.path.join(ROOT, "assets", "models")
HOST_LIB = os.path.join(PKG_DIR, "libb2pt_host.so")
GPU_LIB = os.environ.get("B2PT_GPU_LIB", os.path.join(PKG_DIR, "libb2pt.so"))  # override: experiments with kernel variants

c_float_p = C.POINTER(C.c_float)
c_int_p = C.POINTER(C.c_int32)
c_uint_p = C.POINTER(C.c_uint32)
c_double_p = C.POINTER(C.c_double)

SMOOTH_CONDUCTOR, ROUGH_CONDUCTOR, SMOOTH_DIELECTRIC, ROUGH_DIELECTRIC = 0, 1, 2, 3
FLAG_COUNT_TRAVERSAL = 1
FLAG_SPLIT_WAVELENGTHS = 2
FLAG_FRESH_FRAME = 4
FLAG_INDEPENDENT_WAVELENGTHS = 8
FIX_DIRECT_LIGHT_SAMPLE, FIX_MODEL_QUALITY, FIX_ADD_DIAMOND, FIX_OUTPUT_PATH = 1, 2, 4, 8
NAMED_MATERIALS = ["rough_red_conductor", "rough_white_conductor", "green_mirror", "gold_conductor", "silver_mirror",
                   "smooth_glass", "smooth_glass_gem", "clear_rough_plastic", "rough_plastic"]


class Material(C.Structure):
    _fields_ = [("type", C.c_int32), ("emission", C.c_float * 3), ("ior_a", C.c_float), ("ior_b", C.c_float),
                ("roughness", C.c_float), ("base_reflectance", C.c_float * 3), ("textured", C.c_int32), ("_pad", C.c_int32)]


class Node(C.Structure):
    _fields_ = [("bmin", C.c_float * 3), ("a", C.c_uint32), ("bmax", C.c_float * 3), ("kind", C.c_uint32)]


class SceneDesc(C.Structure):
    _fields_ = [("n_nodes", C.c_uint32), ("nodes", C.POINTER(Node)), ("n_prims", C.c_uint32),
                ("prim_v0", c_float_p), ("prim_e1", c_float_p), ("prim_e2", c_float_p), ("prim_v1v2", c_float_p),
                ("prim_normal", c_float_p), ("prim_uv", c_float_p), ("prim_material", c_uint_p), ("prim_kind", c_uint_p),
                ("n_materials", C.c_uint32), ("materials", C.POINTER(Material)),
                ("n_lights", C.c_uint32), ("light_area", c_float_p), ("light_root", c_uint_p), ("light_material", c_uint_p),
                ("n_light_nodes", C.c_uint32), ("light_node_area", c_float_p), ("light_node_left", c_int_p),
                ("light_node_right", c_int_p), ("light_node_prim", c_int_p),
                ("use_env_map", C.c_int32), ("env_width", C.c_uint32), ("env_height", C.c_uint32), ("env_rgb", c_float_p),
                ("background", C.c_float * 3), ("rr_rate", C.c_float), ("inv_rr", C.c_float),
                ("enable_shadow", C.c_int32), ("n_dir_sample", C.c_int32), ("max_depth", C.c_uint32)]


class Camera(C.Structure):
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("position", C.c_float * 3), ("orientation", C.c_float * 9),
                ("scale", C.c_float), ("aspect", C.c_float), ("use_dof", C.c_int32), ("focal_distance", C.c_float),
                ("aperture_radius", C.c_float)]


class RenderParams(C.Structure):
    _fields_ = [("spp_total", C.c_int32), ("sample_begin", C.c_int32), ("sample_count", C.c_int32), ("seed", C.c_uint64),
                ("max_wave_bundles", C.c_int32), ("flags", C.c_int32)]


class Stats(C.Structure):
    _fields_ = [("gpu_ms", C.c_double), ("extend_ms", C.c_double), ("shadow_ms", C.c_double),
                ("kernel_launches", C.c_uint64), ("extend_launches", C.c_uint64), ("shadow_launches", C.c_uint64),
                ("bundles", C.c_uint64), ("paths", C.c_uint64), ("rays_traced_closest", C.c_uint64),
                ("rays_traced_shadow", C.c_uint64), ("rays_reference", C.c_uint64), ("nodes_fetched", C.c_uint64),
                ("prims_tested", C.c_uint64), ("extend_nodes", C.c_uint64), ("extend_prims", C.c_uint64),
                ("shadow_nodes", C.c_uint64), ("shadow_prims", C.c_uint64), ("vertices_shaded", C.c_uint64), ("max_depth", C.c_uint32), ("waves", C.c_uint32)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


def build(verbose: bool = False) -> None:
    """Compile both libraries and the RayTracing program in-tree (nvcc for sm_100a; no GPU needed)."""
    r = subprocess.run(["make", "-C", PKG_DIR, "all"], capture_output=True, text=True)
    if verbose or r.returncode != 0:
        print(r.stdout + r.stderr)
    if r.returncode != 0:
        raise RuntimeError("building the b2pt libraries failed")


_host = None
_gpu = None


def f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def fp(a):
    return a.ctypes.data_as(c_float_p)


def ip(a):
    return a.ctypes.data_as(c_int_p)


def host_lib():
    global _host
    if _host is None:
        if not os.path.exists(HOST_LIB):
            raise RuntimeError(f"{HOST_LIB} is missing: run build() / make first")
        L = C.CDLL(HOST_LIB)
        L.b2pt_host_last_error.restype = C.c_char_p
        for n in ("b2pt_host_scene_new", "b2pt_host_scene_demo", "b2pt_host_scene_from_conf"):
            getattr(L, n).restype = C.c_void_p
        L.b2pt_host_scene_demo.argtypes = [C.c_char_p, C.c_int, C.c_int]
        L.b2pt_host_scene_from_conf.argtypes = [C.c_char_p, C.c_char_p, C.c_int]
        L.b2pt_host_scene_free.argtypes = [C.c_void_p]
        L.b2pt_host_set_asset_dir.argtypes = [C.c_char_p]
        L.b2pt_host_find_material.argtypes = [C.c_void_p, C.c_char_p]
        L.b2pt_host_add_material.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(Material)]
        L.b2pt_host_set_material.argtypes = [C.c_void_p, C.c_int, C.POINTER(Material)]
        L.b2pt_host_get_material.argtypes = [C.c_void_p, C.c_int, C.POINTER(Material)]
        L.b2pt_host_add_mesh.argtypes = [C.c_void_p, C.c_char_p, C.c_int, c_float_p, C.c_float]
        L.b2pt_host_add_mesh_triangles.argtypes = [C.c_void_p, c_float_p, c_float_p, C.c_int, C.c_int]
        L.b2pt_host_add_sphere.argtypes = [C.c_void_p, c_float_p, C.c_float, C.c_int]
        L.b2pt_host_set_camera.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_float, c_float_p, c_float_p, c_float_p, C.c_int,
                                           C.c_float, C.c_float]
        L.b2pt_host_set_resolution.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.b2pt_host_set_dof.argtypes = [C.c_void_p, C.c_int, C.c_float, C.c_float]
        L.b2pt_host_set_render.argtypes = [C.c_void_p, C.c_int, C.c_float, C.c_int, C.c_int]
        L.b2pt_host_set_background.argtypes = [C.c_void_p, c_float_p]
        L.b2pt_host_load_env_png.argtypes = [C.c_void_p, C.c_char_p]
        L.b2pt_host_set_env_pixels.argtypes = [C.c_void_p, c_float_p, C.c_int, C.c_int]
        L.b2pt_host_scene_build.argtypes = [C.c_void_p]
        L.b2pt_host_scene_desc.argtypes = [C.c_void_p]
        L.b2pt_host_scene_desc.restype = C.POINTER(SceneDesc)
        L.b2pt_host_scene_camera.argtypes = [C.c_void_p]
        L.b2pt_host_scene_camera.restype = C.POINTER(Camera)
        L.b2pt_host_scene_spp.argtypes = [C.c_void_p]
        L.b2pt_host_scene_output_path.argtypes = [C.c_void_p]
        L.b2pt_host_scene_output_path.restype = C.c_char_p
        L.b2pt_host_scene_max_depth.argtypes = [C.c_void_p]
        L.b2pt_host_n_objects.argtypes = [C.c_void_p]
        L.b2pt_host_object_kind.argtypes = [C.c_void_p, C.c_int]
        L.b2pt_host_object_path.argtypes = [C.c_void_p, C.c_int]
        L.b2pt_host_object_path.restype = C.c_char_p
        L.b2pt_host_object_material.argtypes = [C.c_void_p, C.c_int]
        L.b2pt_host_object_transform.argtypes = [C.c_void_p, C.c_int, c_float_p, c_float_p]
        L.b2pt_host_object_sphere.argtypes = [C.c_void_p, C.c_int, c_float_p, c_float_p]
        L.b2pt_host_object_n_tris.argtypes = [C.c_void_p, C.c_int]
        L.b2pt_host_object_triangles.argtypes = [C.c_void_p, C.c_int, c_float_p, c_float_p]
        L.b2pt_host_n_materials.argtypes = [C.c_void_p]
        L.b2pt_host_material_name.argtypes = [C.c_void_p, C.c_int]
        L.b2pt_host_material_name.restype = C.c_char_p
        L.b2pt_host_prim_origin.argtypes = [C.c_void_p, C.c_int, c_int_p, c_int_p]
        L.b2pt_host_prim_of.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.b2pt_host_camera_params.argtypes = [C.c_void_p, c_float_p, c_float_p, c_float_p, c_float_p]
        L.b2pt_host_pack_obj.argtypes = [C.c_char_p, C.c_char_p]
        L.b2pt_host_unpack_to_obj.argtypes = [C.c_char_p, C.c_char_p]
        L.b2pt_host_tonemap_rgba8.argtypes = [c_float_p, C.c_int, C.POINTER(C.c_ubyte)]
        L.b2pt_host_write_png_rgba8.argtypes = [C.c_char_p, C.POINTER(C.c_ubyte), C.c_int, C.c_int]
        L.b2pt_host_read_png_rgba8.argtypes = [C.c_char_p, C.POINTER(C.POINTER(C.c_ubyte)), C.POINTER(C.c_uint), C.POINTER(C.c_uint)]
        L.b2pt_host_free.argtypes = [C.c_void_p]
        L.b2pt_host_set_asset_dir(ASSET_DIR.encode())
        _host = L
    return _host


def gpu_lib():
    """The CUDA library.  Raises when it has not been built: there is no CPU path to fall back to."""
    global _gpu
    if _gpu is None:
        if not os.path.exists(GPU_LIB):
            raise RuntimeError(f"{GPU_LIB} is missing: the CUDA extension must be built (there is no CPU fallback)")
        L = C.CDLL(GPU_LIB)
        L.b2pt_last_error.restype = C.c_char_p
        L.b2pt_last_error.argtypes = [C.c_void_p]
        L.b2pt_create.argtypes = [C.POINTER(C.c_void_p), C.c_int]
        L.b2pt_destroy.argtypes = [C.c_void_p]
        L.b2pt_set_stream.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        L.b2pt_upload_scene.argtypes = [C.c_void_p, C.POINTER(SceneDesc)]
        L.b2pt_update_scene_params.argtypes = [C.c_void_p, C.c_float, C.c_int, C.c_int]
        rp, st = C.POINTER(RenderParams), C.POINTER(Stats)
        L.b2pt_render.argtypes = [C.c_void_p, C.POINTER(Camera), rp, c_float_p, st]
        L.b2pt_group_render.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.POINTER(Camera), rp, c_float_p, st]
        L.b2pt_render_device.argtypes = [C.c_void_p, C.POINTER(Camera), rp, C.c_void_p, st]
        L.b2pt_render_samples.argtypes = [C.c_void_p, C.POINTER(Camera), rp, c_int_p, C.c_int32, c_float_p, st]
        L.b2pt_intersect_batch.argtypes = [C.c_void_p, c_float_p, c_float_p, C.c_int64, c_int_p, c_double_p, st]
        L.b2pt_shadow_batch.argtypes = [C.c_void_p, c_float_p, c_float_p, c_float_p, C.c_int64, c_int_p, st]
        L.b2pt_tri_intersect_batch.argtypes = [C.c_void_p, c_float_p, c_float_p, c_float_p, C.c_int64, c_int_p, c_double_p]
        L.b2pt_box_intersect_batch.argtypes = [C.c_void_p, c_float_p, c_float_p, c_float_p, C.c_int64, c_int_p]
        L.b2pt_sphere_intersect_batch.argtypes = [C.c_void_p, c_float_p, c_float_p, c_float_p, C.c_int64, c_int_p, c_double_p,
                                                  c_float_p, c_float_p]
        L.b2pt_bsdf_eval_batch.argtypes = [C.c_void_p, C.c_int, c_float_p, c_float_p, c_float_p, c_int_p, c_float_p, c_int_p,
                                           C.c_int64, c_float_p]
        L.b2pt_bsdf_pdf_batch.argtypes = [C.c_void_p, C.c_int, c_float_p, c_float_p, c_float_p, c_int_p, c_int_p, C.c_int64,
                                          c_float_p]
        L.b2pt_fresnel_batch.argtypes = [C.c_void_p, C.c_int, c_float_p, c_float_p, c_int_p, C.c_int64, c_float_p]
        L.b2pt_refract_batch.argtypes = [C.c_void_p, C.c_int, c_float_p, c_float_p, c_int_p, C.c_int64, c_float_p]
        L.b2pt_reflect_batch.argtypes = [C.c_void_p, c_float_p, c_float_p, C.c_int64, c_float_p]
        L.b2pt_material_sample_batch.argtypes = [C.c_void_p, C.c_int, c_float_p, c_float_p, c_float_p, C.c_int64, c_float_p]
        L.b2pt_env_lookup_batch.argtypes = [C.c_void_p, c_float_p, C.c_int64, c_float_p]
        L.b2pt_sample_light_batch.argtypes = [C.c_void_p, c_float_p, C.c_int64, c_float_p, c_float_p, c_float_p, c_float_p]
        L.b2pt_camera_rays_batch.argtypes = [C.c_void_p, C.POINTER(Camera), c_int_p, C.c_int32, C.c_int32, C.c_int32, C.c_uint64,
                                             c_float_p, c_float_p]
        L.b2pt_stream_uniforms.argtypes = [C.c_void_p, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int32,
                                           c_float_p]
        L.b2pt_measure_copy_gbs.argtypes = [C.c_void_p, C.c_size_t, C.c_int, c_double_p]
        L.b2pt_tonemap_rgba8.argtypes = [C.c_void_p, c_float_p, C.c_int, C.POINTER(C.c_ubyte)]
        L.b2pt_measure_l2_read_gbs.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_int, c_double_p]
        _gpu = L
    return _gpu


class HostScene:
    """A scene under assembly (mirrors what the reference's main() holds before Render)."""

    def __init__(self, handle):
        if not handle:
            raise RuntimeError("scene assembly failed: " + host_lib().b2pt_host_last_error().decode())
        self.h = C.c_void_p(handle)
        self.L = host_lib()

    @classmethod
    def empty(cls):
        return cls(host_lib().b2pt_host_scene_new())

    @classmethod
    def demo(cls, width=0, height=0, models_dir=None):
        """DEMO Cornell scene, src/main.cpp:99-129."""
        md = models_dir or os.path.join(ROOT, "_no_models_dir_")  # falls back to assets/*.b2m
        return cls(host_lib().b2pt_host_scene_demo(md.encode(), width, height))

    @classmethod
    def from_conf(cls, conf_path, run_dir=None, fix_flags=0):
        """conf.json chess scene, src/main.cpp:137-316."""
        rd = run_dir or os.path.dirname(os.path.abspath(conf_path))
        return cls(host_lib().b2pt_host_scene_from_conf(conf_path.encode(), rd.encode(), fix_flags))

    def close(self):
        if self.h:
            self.L.b2pt_host_scene_free(self.h)
            self.h = None

    def _ck(self, r):
        if r < 0:
            raise RuntimeError(self.L.b2pt_host_last_error().decode())
        return r

    def find_material(self, name):
        return self.L.b2pt_host_find_material(self.h, name.encode())

    def add_material(self, name, m: Material):
        return self._ck(self.L.b2pt_host_add_material(self.h, name.encode(), C.byref(m)))

    def get_material(self, idx) -> Material:
        m = Material()
        self._ck(self.L.b2pt_host_get_material(self.h, idx, C.byref(m)))
        return m

    def set_material(self, idx, m: Material):
        self._ck(self.L.b2pt_host_set_material(self.h, idx, C.byref(m)))

    def add_mesh(self, path, material, translation=(0, 0, 0), zoom=1.0):
        t = f32(translation)
        return self._ck(self.L.b2pt_host_add_mesh(self.h, path.encode(), material, fp(t), zoom))

    def add_triangles(self, v9, material, uv6=None):
        v = f32(v9).reshape(-1, 9)
        u = f32(uv6).reshape(-1, 6) if uv6 is not None else None
        return self._ck(self.L.b2pt_host_add_mesh_triangles(self.h, fp(v), fp(u) if u is not None else None, len(v), material))

    def add_sphere(self, center, radius, material):
        c = f32(center)
        return self._ck(self.L.b2pt_host_add_sphere(self.h, fp(c), radius, material))

    def set_camera(self, width, height, fov, pos, target, up=(0, 1, 0), use_dof=False, focal_distance=100.0, aperture_radius=5.0):
        p, t, u = f32(pos), f32(target), f32(up)
        self.L.b2pt_host_set_camera(self.h, width, height, fov, fp(p), fp(t), fp(u), int(use_dof), focal_distance, aperture_radius)

    def set_resolution(self, width, height):
        self.L.b2pt_host_set_resolution(self.h, width, height)

    def set_dof(self, use_dof, focal_distance=-1.0, aperture_radius=-1.0):
        self.L.b2pt_host_set_dof(self.h, int(use_dof), focal_distance, aperture_radius)

    def set_render(self, spp=0, rr_rate=-1.0, enable_shadow=-1, n_dir_sample=0):
        self.L.b2pt_host_set_render(self.h, spp, rr_rate, enable_shadow, n_dir_sample)

    def set_background(self, rgb):
        b = f32(rgb)
        self.L.b2pt_host_set_background(self.h, fp(b))

    def set_env_pixels(self, rgb):
        a = f32(rgb)
        self._ck(self.L.b2pt_host_set_env_pixels(self.h, fp(a), a.shape[1], a.shape[0]))

    def load_env_png(self, path):
        return self.L.b2pt_host_load_env_png(self.h, path.encode())

    def build_tree(self):
        self._ck(self.L.b2pt_host_scene_build(self.h))
        return self

    @property
    def desc(self) -> SceneDesc:
        p = self.L.b2pt_host_scene_desc(self.h)
        if not p:
            raise RuntimeError("scene not built")
        return p.contents

    @property
    def camera(self) -> Camera:
        return self.L.b2pt_host_scene_camera(self.h).contents

    @property
    def spp(self):
        return self.L.b2pt_host_scene_spp(self.h)

    @property
    def max_depth(self):
        return self.L.b2pt_host_scene_max_depth(self.h)

    @property
    def n_objects(self):
        return self.L.b2pt_host_n_objects(self.h)

    def object_info(self, k):
        kind = self.L.b2pt_host_object_kind(self.h, k)
        mat = self.L.b2pt_host_object_material(self.h, k)
        if kind == 1:
            c = np.zeros(3, np.float32)
            r = C.c_float()
            self.L.b2pt_host_object_sphere(self.h, k, fp(c), C.byref(r))
            return {"kind": "sphere", "material": mat, "center": c, "radius": r.value}
        n = self.L.b2pt_host_object_n_tris(self.h, k)
        v = np.zeros((n, 9), np.float32)
        uv = np.zeros((n, 6), np.float32)
        self.L.b2pt_host_object_triangles(self.h, k, fp(v), fp(uv))
        return {"kind": "mesh", "material": mat, "v9": v, "uv6": uv, "path": self.L.b2pt_host_object_path(self.h, k).decode()}

    def materials(self):
        n = self.L.b2pt_host_n_materials(self.h)
        return [(self.L.b2pt_host_material_name(self.h, i).decode(), self.get_material(i)) for i in range(n)]

    def prim_origin(self, prim):
        o, f = C.c_int32(), C.c_int32()
        self.L.b2pt_host_prim_origin(self.h, int(prim), C.byref(o), C.byref(f))
        return o.value, f.value

    def prim_origins(self):
        """(object, face) of every primitive id as two int arrays."""
        n = self.desc.n_prims
        o = np.zeros(n, np.int32)
        f = np.zeros(n, np.int32)
        for i in range(n):
            o[i], f[i] = self.prim_origin(i)
        return o, f

    def camera_params(self):
        fov = C.c_float()
        p, t, u = (np.zeros(3, np.float32) for _ in range(3))
        self.L.b2pt_host_camera_params(self.h, C.byref(fov), fp(p), fp(t), fp(u))
        return fov.value, p, t, u


def tonemap_rgba8(rgb):
    """Renderer.cpp:93-102: gamma 0.45, clamp (NaN -> 255), truncation to 8 bit, alpha 255."""
    a = f32(rgb).reshape(-1, 3)
    out = np.zeros((len(a), 4), np.uint8)
    host_lib().b2pt_host_tonemap_rgba8(fp(a), len(a), out.ctypes.data_as(C.POINTER(C.c_ubyte)))
    return out


def write_png(path, rgba, width, height):
    a = np.ascontiguousarray(rgba, np.uint8)
    if host_lib().b2pt_host_write_png_rgba8(path.encode(), a.ctypes.data_as(C.POINTER(C.c_ubyte)), width, height) != 0:
        raise RuntimeError(host_lib().b2pt_host_last_error().decode())


def read_png(path):
    p = C.POINTER(C.c_ubyte)()
    w, h = C.c_uint(), C.c_uint()
    if host_lib().b2pt_host_read_png_rgba8(path.encode(), C.byref(p), C.byref(w), C.byref(h)) != 0:
        raise RuntimeError(host_lib().b2pt_host_last_error().decode())
    a = np.ctypeslib.as_array(p, shape=(h.value, w.value, 4)).copy()
    host_lib().b2pt_host_free(p)
    return a


class Context:
    """One GPU context of the CUDA library (b2pt_create .. b2pt_destroy)."""

    def __init__(self, device=0):
        self.L = gpu_lib()
        h = C.c_void_p()
        r = self.L.b2pt_create(C.byref(h), device)
        if r != 0:
            raise RuntimeError(f"b2pt_create failed ({r}): " + self.L.b2pt_last_error(None).decode())
        self.h = h

    def close(self):
        if self.h:
            self.L.b2pt_destroy(self.h)
            self.h = None

    def _ck(self, r):
        if r != 0:
            raise RuntimeError(f"b2pt error {r}: " + self.L.b2pt_last_error(self.h).decode())

    def set_stream(self, cuda_stream=None):
        """Issue work on `cuda_stream` (an int cudaStream_t handle, 0 = default stream); None = the context's own stream."""
        self._ck(self.L.b2pt_set_stream(self.h, C.c_void_p(cuda_stream or 0), 0 if cuda_stream is None else 1))

    def upload(self, scene: HostScene):
        self._ck(self.L.b2pt_upload_scene(self.h, C.byref(scene.desc)))
        return self

    def set_params(self, rr_rate=-1.0, enable_shadow=-1, n_dir_sample=0):
        self._ck(self.L.b2pt_update_scene_params(self.h, rr_rate, enable_shadow, n_dir_sample))

    @staticmethod
    def _params(spp_total, sample_begin, sample_count, seed, max_wave_bundles, flags):
        return RenderParams(spp_total, sample_begin, sample_count if sample_count else spp_total - sample_begin, seed,
                            max_wave_bundles, flags)

    def render(self, cam: Camera, spp, seed=0x5EED0001, sample_begin=0, sample_count=0, out=None, max_wave_bundles=0, flags=0):
        """Host-buffer frame: adds sum_k rgb_k / spp into `out` ([H, W, 3] fp32) and returns (out, stats)."""
        if out is None:
            out = np.empty((cam.height, cam.width, 3), np.float32)
            flags |= FLAG_FRESH_FRAME  # nothing to accumulate into
        p = self._params(spp, sample_begin, sample_count, seed, max_wave_bundles, flags)
        st = Stats()
        self._ck(self.L.b2pt_render(self.h, C.byref(cam), C.byref(p), fp(out), C.byref(st)))
        return out, st

    def render_device(self, cam: Camera, spp, device_ptr, seed=0x5EED0001, sample_begin=0, sample_count=0, max_wave_bundles=0,
                      flags=0):
        p = self._params(spp, sample_begin, sample_count, seed, max_wave_bundles, flags)
        st = Stats()
        self._ck(self.L.b2pt_render_device(self.h, C.byref(cam), C.byref(p), C.c_void_p(device_ptr), C.byref(st)))
        return st

    def render_samples(self, cam: Camera, pixels, sample_begin, sample_count, seed=0x5EED0001, flags=0):
        px = i32(pixels)
        out = np.zeros((len(px), sample_count, 3), np.float32)
        p = RenderParams(sample_count, sample_begin, sample_count, seed, 0, flags)
        st = Stats()
        self._ck(self.L.b2pt_render_samples(self.h, C.byref(cam), C.byref(p), ip(px), len(px), fp(out), C.byref(st)))
        return out, st

    def intersect(self, origins, dirs, count=False):
        o, d = f32(origins).reshape(-1, 3), f32(dirs).reshape(-1, 3)
        prim = np.zeros(len(o), np.int32)
        t = np.zeros(len(o), np.float64)
        st = Stats()
        self._ck(self.L.b2pt_intersect_batch(self.h, fp(o), fp(d), len(o), ip(prim), t.ctypes.data_as(c_double_p),
                                             C.byref(st) if count else None))
        return (prim, t, st) if count else (prim, t)

    def shadow(self, origins, dirs, dist):
        o, d, s = f32(origins).reshape(-1, 3), f32(dirs).reshape(-1, 3), f32(dist)
        vis = np.zeros(len(o), np.int32)
        self._ck(self.L.b2pt_shadow_batch(self.h, fp(o), fp(d), fp(s), len(o), ip(vis), None))
        return vis

    def tri_intersect(self, v9, origins, dirs):
        v, o, d = f32(v9).reshape(-1, 9), f32(origins).reshape(-1, 3), f32(dirs).reshape(-1, 3)
        hit = np.zeros(len(o), np.int32)
        t = np.zeros(len(o), np.float64)
        self._ck(self.L.b2pt_tri_intersect_batch(self.h, fp(v), fp(o), fp(d), len(o), ip(hit), t.ctypes.data_as(c_double_p)))
        return hit, t

    def box_intersect(self, b6, origins, dirs):
        b, o, d = f32(b6).reshape(-1, 6), f32(origins).reshape(-1, 3), f32(dirs).reshape(-1, 3)
        hit = np.zeros(len(o), np.int32)
        self._ck(self.L.b2pt_box_intersect_batch(self.h, fp(b), fp(o), fp(d), len(o), ip(hit)))
        return hit

    def sphere_intersect(self, c4, origins, dirs):
        c, o, d = f32(c4).reshape(-1, 4), f32(origins).reshape(-1, 3), f32(dirs).reshape(-1, 3)
        hit = np.zeros(len(o), np.int32)
        t = np.zeros(len(o), np.float64)
        co = np.zeros((len(o), 3), np.float32)
        nn = np.zeros((len(o), 3), np.float32)
        self._ck(self.L.b2pt_sphere_intersect_batch(self.h, fp(c), fp(o), fp(d), len(o), ip(hit), t.ctypes.data_as(c_double_p),
                                                    fp(co), fp(nn)))
        return hit, t, co, nn

    def bsdf_eval(self, material, wi, wo, n, wl, uv, is_reflect):
        wi, wo, n, uv = f32(wi), f32(wo), f32(n), f32(uv)
        wl, rf = i32(wl), i32(is_reflect)
        out = np.zeros(len(wl), np.float32)
        self._ck(self.L.b2pt_bsdf_eval_batch(self.h, material, fp(wi), fp(wo), fp(n), ip(wl), fp(uv), ip(rf), len(wl), fp(out)))
        return out

    def bsdf_pdf(self, material, wi, wo, n, wl, is_reflect):
        wi, wo, n = f32(wi), f32(wo), f32(n)
        wl, rf = i32(wl), i32(is_reflect)
        out = np.zeros(len(wl), np.float32)
        self._ck(self.L.b2pt_bsdf_pdf_batch(self.h, material, fp(wi), fp(wo), fp(n), ip(wl), ip(rf), len(wl), fp(out)))
        return out

    def fresnel(self, material, I, n, wl):
        I, n, wl = f32(I), f32(n), i32(wl)
        out = np.zeros(len(wl), np.float32)
        self._ck(self.L.b2pt_fresnel_batch(self.h, material, fp(I), fp(n), ip(wl), len(wl), fp(out)))
        return out

    def refract(self, material, I, n, wl):
        I, n, wl = f32(I), f32(n), i32(wl)
        out = np.zeros((len(wl), 3), np.float32)
        self._ck(self.L.b2pt_refract_batch(self.h, material, fp(I), fp(n), ip(wl), len(wl), fp(out)))
        return out

    def reflect(self, I, n):
        I, n = f32(I).reshape(-1, 3), f32(n).reshape(-1, 3)
        out = np.zeros((len(I), 3), np.float32)
        self._ck(self.L.b2pt_reflect_batch(self.h, fp(I), fp(n), len(I), fp(out)))
        return out

    def material_sample(self, material, wo, n, u2):
        wo, n, u2 = f32(wo).reshape(-1, 3), f32(n).reshape(-1, 3), f32(u2).reshape(-1, 2)
        out = np.zeros((len(n), 3), np.float32)
        self._ck(self.L.b2pt_material_sample_batch(self.h, material, fp(wo), fp(n), fp(u2), len(n), fp(out)))
        return out

    def env_lookup(self, dirs):
        d = f32(dirs).reshape(-1, 3)
        out = np.zeros((len(d), 3), np.float32)
        self._ck(self.L.b2pt_env_lookup_batch(self.h, fp(d), len(d), fp(out)))
        return out

    def sample_light(self, u4):
        u = f32(u4).reshape(-1, 4)
        co, nn, em = (np.zeros((len(u), 3), np.float32) for _ in range(3))
        pdf = np.zeros(len(u), np.float32)
        self._ck(self.L.b2pt_sample_light_batch(self.h, fp(u), len(u), fp(co), fp(nn), fp(em), fp(pdf)))
        return co, nn, em, pdf

    def camera_rays(self, cam: Camera, pixels, sample_begin, sample_count, seed=0x5EED0001):
        px = i32(pixels)
        o = np.zeros((len(px) * sample_count, 3), np.float32)
        d = np.zeros((len(px) * sample_count, 3), np.float32)
        self._ck(self.L.b2pt_camera_rays_batch(self.h, C.byref(cam), ip(px), len(px), sample_begin, sample_count, seed, fp(o), fp(d)))
        return o, d

    def stream_uniforms(self, seed, pixel, sample, tag, dim_begin, count):
        out = np.zeros(count, np.float32)
        self._ck(self.L.b2pt_stream_uniforms(self.h, seed, pixel, sample, tag, dim_begin, count, fp(out)))
        return out

    def measure_copy_gbs(self, nbytes=1 << 30, iters=10):
        g = C.c_double()
        self._ck(self.L.b2pt_measure_copy_gbs(self.h, nbytes, iters, C.byref(g)))
        return g.value

    def tonemap_rgba8(self, rgb=None, n_pixels=0):
        """Renderer.cpp:93-102 on the device: RGBA8 bytes of `rgb` ([n][3] fp32), or with rgb=None of the frame the last
        render() left on the device.  Byte-identical to the host loop (b2pt_host_tonemap_rgba8)."""
        if rgb is not None:
            rgb = np.ascontiguousarray(rgb, np.float32).reshape(-1, 3)
            n_pixels = len(rgb)
        out = np.zeros((n_pixels, 4), np.uint8)
        self._ck(self.L.b2pt_tonemap_rgba8(self.h, fp(rgb) if rgb is not None else None, n_pixels, out.ctypes.data_as(C.POINTER(C.c_ubyte))))
        return out

    def measure_l2_read_gbs(self, nbytes=32 << 20, repeats=64, iters=5):
        """Read bandwidth out of the L2 (GB/s): the denominator for traversal kernels whose tree is cache-resident."""
        g = C.c_double()
        self._ck(self.L.b2pt_measure_l2_read_gbs(self.h, nbytes, repeats, iters, C.byref(g)))
        return g.value


def group_render(contexts, cam: Camera, spp, seed=0x5EED0001, sample_begin=0, sample_count=0, out=None, flags=0):
    """One frame over several GPUs from one host thread: spp split across `contexts` (one per device, same scene uploaded),
    one ncclReduce of the fp32 frame onto the first device (b2pt_group_render)."""
    L = gpu_lib()
    if out is None:
        out = np.empty((cam.height, cam.width, 3), np.float32)
        flags |= FLAG_FRESH_FRAME
    p = Context._params(spp, sample_begin, sample_count, seed, 0, flags)
    st = Stats()
    arr = (C.c_void_p * len(contexts))(*[c.h for c in contexts])
    r = L.b2pt_group_render(arr, len(contexts), C.byref(cam), C.byref(p), fp(out), C.byref(st))
    if r != 0:
        raise RuntimeError(f"b2pt error {r}: " + L.b2pt_last_error(contexts[0].h).decode())
    return out, st


def device_count():
    """Number of CUDA devices visible to the library's process (via the runtime the library links)."""
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0
