"""The reference-side binding of INTEGRATION.md (oracle/b2pt_bind.hpp, compiled against the reference's own headers) applied
to the reference's OWN pointer trees produces exactly the arrays the host assembler (host/scene.cpp) builds from scratch:
same nodes in the same order (so the same BVHAccel::recursiveBuild topology, std::sort permutation included), same
primitive numbering, same light trees, same camera.  CPU-only."""
import ctypes as C

import numpy as np
import pytest

import scenes
import support as S

b2pt = S.b2pt
pytestmark = pytest.mark.skipif(not S.have_ref(), reason="oracle/_ref/libref_oracle.so not built")


def arr(ptr, n, dtype, width=1):
    a = np.ctypeslib.as_array(ptr, shape=(n * width,)).view(dtype)
    return a.reshape(n, width) if width > 1 else a


@pytest.mark.parametrize("which", ["cornell", "chess_sky_dof", "two_triangles"])
def test_binding_equals_host_flattening(which):
    if which == "cornell":
        sc, env = scenes.cornell(64, 64)
    elif which == "chess_sky_dof":
        sc, env = scenes.chess(96, 54, dof=True, sky=True)
    else:
        sc, env = scenes.two_triangle_scene()
    ref = S.Ref(sc, env)
    L = ref.L
    L.ref_flatten.restype = C.c_void_p
    L.ref_flatten.argtypes = [C.c_void_p]
    L.ref_flat_desc.restype = C.POINTER(b2pt.SceneDesc)
    L.ref_flat_desc.argtypes = [C.c_void_p]
    L.ref_flat_camera.restype = C.POINTER(b2pt.Camera)
    L.ref_flat_camera.argtypes = [C.c_void_p]
    L.ref_flat_material_index.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
    L.ref_flat_free.argtypes = [C.c_void_p]
    f = C.c_void_p(L.ref_flatten(ref.h))
    a, b = L.ref_flat_desc(f).contents, sc.desc
    assert a.n_nodes == b.n_nodes and a.n_prims == b.n_prims and a.max_depth == b.max_depth
    na = np.ctypeslib.as_array(C.cast(a.nodes, C.POINTER(C.c_uint32)), shape=(a.n_nodes, 8))
    nb = np.ctypeslib.as_array(C.cast(b.nodes, C.POINTER(C.c_uint32)), shape=(b.n_nodes, 8))
    assert np.array_equal(na, nb)  # boxes (as bits), child / primitive indices, kinds: identical topology
    n = a.n_prims
    for field, width in (("prim_v0", 4), ("prim_e1", 4), ("prim_e2", 4), ("prim_normal", 4), ("prim_v1v2", 6), ("prim_uv", 6)):
        xa = np.ctypeslib.as_array(getattr(a, field), shape=(n, width)).view(np.uint32)
        xb = np.ctypeslib.as_array(getattr(b, field), shape=(n, width)).view(np.uint32)
        assert np.array_equal(xa, xb), field
    assert np.array_equal(np.ctypeslib.as_array(a.prim_kind, shape=(n,)), np.ctypeslib.as_array(b.prim_kind, shape=(n,)))
    # materials are numbered in first-use order by the binding and in creation order by the host: compare through the map
    ma, mb = np.ctypeslib.as_array(a.prim_material, shape=(n,)), np.ctypeslib.as_array(b.prim_material, shape=(n,))
    for host_index in np.unique(mb):
        bind_index = L.ref_flat_material_index(ref.h, f, int(host_index))
        assert bind_index >= 0 and np.array_equal(ma == bind_index, mb == host_index)
        x, y = a.materials[bind_index], b.materials[int(host_index)]
        assert (x.type, x.ior_a, x.ior_b, x.roughness, x.textured, list(x.emission), list(x.base_reflectance)) == \
               (y.type, y.ior_a, y.ior_b, y.roughness, y.textured, list(y.emission), list(y.base_reflectance))
    # light trees
    assert a.n_lights == b.n_lights and a.n_light_nodes == b.n_light_nodes
    for field in ("light_area", "light_root"):
        assert np.array_equal(np.ctypeslib.as_array(getattr(a, field), shape=(a.n_lights,)), np.ctypeslib.as_array(getattr(b, field), shape=(b.n_lights,)))
    for field in ("light_node_area", "light_node_left", "light_node_right", "light_node_prim"):
        xa = np.ctypeslib.as_array(getattr(a, field), shape=(a.n_light_nodes,))
        xb = np.ctypeslib.as_array(getattr(b, field), shape=(b.n_light_nodes,))
        assert np.array_equal(xa.view(np.uint32), xb.view(np.uint32)), field
    assert (a.rr_rate, a.inv_rr, a.enable_shadow, a.n_dir_sample, a.use_env_map, a.env_width, a.env_height) == \
           (b.rr_rate, b.inv_rr, b.enable_shadow, b.n_dir_sample, b.use_env_map, b.env_width, b.env_height)
    if a.use_env_map:
        k = a.env_width * a.env_height * 3
        assert np.array_equal(np.ctypeslib.as_array(a.env_rgb, shape=(k,)), np.ctypeslib.as_array(b.env_rgb, shape=(k,)))
    ca, cb = L.ref_flat_camera(f).contents, sc.camera
    assert (ca.width, ca.height, ca.scale, ca.aspect, ca.use_dof, ca.focal_distance, ca.aperture_radius, list(ca.position), list(ca.orientation)) == \
           (cb.width, cb.height, cb.scale, cb.aspect, cb.use_dof, cb.focal_distance, cb.aperture_radius, list(cb.position), list(cb.orientation))
    L.ref_flat_free(f)
    ref.close(); sc.close()
