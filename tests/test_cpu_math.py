"""CPU-only parity tests (`-m "not gpu"`): the device arithmetic of csrc/pt_math.cuh, compiled for the host
by tests/hostcheck, against the oracle (the reference's own sources).  The `-m gpu` tests repeat these
through the C ABI on the device; these keep the arithmetic honest on a machine without a GPU."""
import numpy as np
import pytest

import scenes
import support as S
from gen import adversarial_triangle_cases, bsdf_inputs, box_cases, degenerate_rays, rel_close, sphere_cases, uniforms

b2pt = S.b2pt
pytestmark = pytest.mark.skipif(not S.have_ref(), reason="oracle/_ref/libref_oracle.so not built")


@pytest.fixture(scope="module", params=["cornell", "chess_sky_dof", "chess_dark"])
def world(request):
    if request.param == "cornell":
        sc, env = scenes.cornell(96, 96)
    elif request.param == "chess_sky_dof":
        sc, env = scenes.chess(160, 90, dof=True, sky=True)
    else:
        sc, env = scenes.chess(160, 90, dof=False, sky=False)
    ref = S.Ref(sc, env)
    hc = S.HostCheck(sc)
    yield request.param, sc, ref, hc
    hc.close()
    ref.close()
    sc.close()


def test_triangle_bit_exact():
    v, o, d = adversarial_triangle_cases(np.random.RandomState(7), 100000)
    hit_h, t_h = S.hc_tri(v, o, d)
    hit_r, t_r = S.ref_tri(v, o, d)
    assert np.array_equal(hit_h, hit_r)
    assert 0.2 < hit_r.mean() < 0.95
    assert np.array_equal(t_h[hit_r == 1].view(np.uint64), t_r[hit_r == 1].view(np.uint64))


def test_box_bit_exact():
    b6, o, d = box_cases(np.random.RandomState(8), 100000)
    hit_r = S.ref_box(b6, o, d)
    assert np.array_equal(S.hc_box(b6, o, d), hit_r)
    assert 0.02 < hit_r.mean() < 0.9


def test_sphere_bit_exact():
    c4, o, d = sphere_cases(np.random.RandomState(9), 50000)
    hit_h, t_h = S.hc_sphere(c4, o, d)
    hit_r, t_r, _, _ = S.ref_sphere(c4, o, d)
    assert np.array_equal(hit_h, hit_r)
    assert 0.2 < hit_r.mean() < 0.99
    assert np.array_equal(t_h[hit_r == 1].view(np.uint64), t_r[hit_r == 1].view(np.uint64))


def test_scene_intersect_bit_exact(world):
    name, sc, ref, hc = world
    o, d, _ = scenes.ray_batch(ref, sc, n_pixels=1500, samples=2, seed=3)
    prim_r, t_r, co_r, nn_r, uv_r = ref.intersect(o, d)
    prim_h, t_h, (nodes, prims) = hc.intersect(o, d, counts=True)
    assert np.array_equal(prim_h, prim_r), f"{name}: {(prim_h != prim_r).sum()} hit ids differ"
    assert np.array_equal(t_h.view(np.uint64), t_r.view(np.uint64))
    hit = prim_r >= 0
    assert hit.mean() > 0.3
    co_h, nn_h, uv_h = hc.surface(o, d)
    assert np.array_equal(co_h[hit].view(np.uint32), co_r[hit].view(np.uint32))
    assert np.array_equal(nn_h[hit].view(np.uint32), nn_r[hit].view(np.uint32))
    assert nodes > 0 and prims > 0


def test_binary_and_four_wide_walks_agree(world):
    """The sibling-pair walk (shadow kernel, degenerate rays) and the four-wide walk (extend kernel) over the same leaves
    return the same hits bit for bit; the four-wide one needs about half the steps (one step = four boxes)."""
    name, sc, ref, hc = world
    o, d, _ = scenes.ray_batch(ref, sc, n_pixels=1500, samples=2, seed=5)
    prim_b, t_b, (boxes_b, _) = hc.intersect(o, d, counts=True, binary=True)
    prim_q, t_q, (boxes_q, _) = hc.intersect(o, d, counts=True)
    assert np.array_equal(prim_b, prim_q) and np.array_equal(t_b.view(np.uint64), t_q.view(np.uint64))
    quads, need = hc.quad_stats()
    assert quads > 0 and need + 2 < 64
    if name != "cornell":  # 35 primitives: nothing to collapse
        assert boxes_q / 4 < 0.7 * boxes_b / 2, (boxes_q, boxes_b)


def test_degenerate_rays_take_the_reference_tree(world):
    """Zero / denormal direction components and origins on box planes: slab products are NaN or infinite."""
    name, sc, ref, hc = world
    root = sc.desc.nodes[0]
    o, d = degenerate_rays(np.random.RandomState(31), list(root.bmin), list(root.bmax), 4000)
    prim_r, t_r, *_ = ref.intersect(o, d)
    prim_h, t_h = hc.intersect(o, d)
    assert np.array_equal(prim_h, prim_r), f"{name}: {(prim_h != prim_r).sum()} hit ids differ"
    assert np.array_equal(t_h.view(np.uint64), t_r.view(np.uint64))
    assert (prim_r >= 0).mean() > 0.05


def test_textured_uv(world):
    name, sc, ref, hc = world
    if name == "cornell":
        pytest.skip("no textured mesh in the Cornell scene")
    # rays straight down onto the chessboard floor
    rng = np.random.RandomState(5)
    n = 4000
    o = np.stack([rng.uniform(-600, 1100, n), np.full(n, 500.0), rng.uniform(-2400, 200, n)], 1).astype(np.float32)
    d = np.tile(np.array([[0.01, -1.0, 0.02]], np.float32), (n, 1))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    prim_r, t_r, co_r, nn_r, uv_r = ref.intersect(o, d)
    po, pf = sc.prim_origins()
    floor_obj = [k for k in range(sc.n_objects) if sc.object_info(k)["kind"] == "mesh" and len(sc.object_info(k)["v9"]) == 2
                 and sc.get_material(sc.object_info(k)["material"]).textured][0]
    on_floor = (prim_r >= 0) & (po[np.maximum(prim_r, 0)] == floor_obj)
    assert on_floor.mean() > 0.2
    co_h, nn_h, uv_h = hc.surface(o, d)
    assert np.array_equal(uv_h[on_floor].view(np.uint32), uv_r[on_floor].view(np.uint32))


def test_treeless_walk_of_small_scenes(world):
    """Scenes of at most pt::kFlatMax primitives (the Cornell box) are not walked through a tree at all: every primitive's own leaf
    box is tested, then the primitive (pt::flat_closest / flat_unoccluded).  Same hits, same t bits, same visibility decisions."""
    name, sc, ref, hc = world
    o, d, (p, ws, dist, u4) = scenes.ray_batch(ref, sc, n_pixels=2000, samples=2, seed=6)
    got = hc.flat_intersect(o, d)
    if name != "cornell":
        assert got is None  # the chess scene has 38 458 primitives
        return
    prim_r, t_r, *_ = ref.intersect(o, d)
    assert np.array_equal(got[0], prim_r)
    hit = prim_r >= 0
    assert np.array_equal(got[1][hit].view(np.uint64), t_r[hit].view(np.uint64))
    prim_s, t_s, *_ = ref.intersect(p, ws)
    want = ((prim_s >= 0) & (np.abs(t_s - dist.astype(np.float64)) < np.float64(np.float32(1e-4)))).astype(np.int32)
    assert np.array_equal(hc.flat_shadow(p, ws, dist), want)
    # degenerate rays (zero direction components, origins on box planes) go through the reference's topology: still equal
    root = sc.desc.nodes[0]
    od, dd = degenerate_rays(np.random.RandomState(3), list(root.bmin), list(root.bmax), 4000)
    pr, tr, *_ = ref.intersect(od, dd)
    g = hc.flat_intersect(od, dd)
    assert np.array_equal(g[0], pr) and np.array_equal(g[1][pr >= 0].view(np.uint64), tr[pr >= 0].view(np.uint64))


def test_shadow_decision(world):
    name, sc, ref, hc = world
    _, _, (p, ws, dist, u4) = scenes.ray_batch(ref, sc, n_pixels=1500, samples=2, seed=4)
    prim_r, t_r, *_ = ref.intersect(p, ws)
    want = ((prim_r >= 0) & (np.abs(t_r - dist.astype(np.float64)) < np.float64(np.float32(1e-4)))).astype(np.int32)
    got = hc.shadow(p, ws, dist)
    assert np.array_equal(got, want), f"{name}: {(got != want).sum()} of {len(want)} visibility decisions differ"
    assert 0.01 < want.mean() < 0.99
    # the four-wide walk over the padded quads (what the shadow kernel runs): the same decisions
    got4, (boxes, prims) = hc.shadow4(p, ws, dist, counts=True)
    assert np.array_equal(got4, want), f"{name}: {(got4 != want).sum()} of {len(want)} four-wide visibility decisions differ"
    assert boxes > 0
    # the render path also passes the sampled light triangle (tested first): same decisions
    got2 = hc.shadow_with_light_node(p, ws, dist, hc.sample_light_node(u4))
    assert np.array_equal(got2, want)


def test_bsdf_parity():
    sc, _ = scenes.two_triangle_scene()
    ref, hc = S.Ref(sc), S.HostCheck(sc)
    rng = np.random.RandomState(11)
    n = 20000
    wi, wo, nrm, wl, uv, rf = bsdf_inputs(rng, n)
    for mat in range(len(b2pt.NAMED_MATERIALS)):
        name = b2pt.NAMED_MATERIALS[mat]
        assert rel_close(hc.bsdf_eval(mat, wi, wo, nrm, wl, uv, rf), ref.bsdf_eval(mat, wi, wo, nrm, wl, uv, rf), 1e-5, 1e-7).all(), (name, "eval")
        assert rel_close(hc.bsdf_pdf(mat, wi, wo, nrm, wl, rf), ref.bsdf_pdf(mat, wi, wo, nrm, wl, rf), 1e-5, 1e-7).all(), (name, "pdf")
        assert rel_close(hc.fresnel(mat, wi, nrm, wl), ref.fresnel(mat, wi, nrm, wl), 1e-5, 1e-7).all(), (name, "fresnel")
        assert np.array_equal(hc.refract(mat, wi, nrm, wl).view(np.uint32), ref.refract(mat, wi, nrm, wl).view(np.uint32)), (name, "refract")
        u2 = uniforms(rng, n, 2)
        assert np.array_equal(hc.material_sample(mat, nrm, u2).view(np.uint32), ref.material_sample(mat, wo, nrm, u2).view(np.uint32)), (name, "sample")
    assert np.array_equal(hc.reflect(wi, nrm).view(np.uint32), ref.reflect(0, wi, nrm).view(np.uint32))
    # the smooth lobes and the rough lobes are both exercised (not all zeros)
    assert (ref.bsdf_eval(b2pt.NAMED_MATERIALS.index("silver_mirror"), wi, wo, nrm, wl, uv, np.ones(n, np.int32)) > 0).mean() > 0.05
    assert (ref.bsdf_eval(b2pt.NAMED_MATERIALS.index("rough_plastic"), wi, wo, nrm, wl, uv, rf) > 0).mean() > 0.2
    hc.close(); ref.close(); sc.close()


def test_zero_summand_predicate_implies_eval_zero():
    """mat_eval_returns_zero (used to drop light samples before any visibility work) is true ONLY where the reference's own
    Material::eval returns its literal 0 — for every material, lobe and wavelength, including directions on the smooth cones'
    edges — and it catches the bulk of them for the smooth materials (otherwise it would not be worth having)."""
    sc, _ = scenes.two_triangle_scene()
    ref, hc = S.Ref(sc), S.HostCheck(sc)
    rng = np.random.RandomState(12)
    wi, wo, nrm, wl, uv, rf = bsdf_inputs(rng, 40000)
    for mat, name in enumerate(b2pt.NAMED_MATERIALS):
        gate = hc.bsdf_eval_returns_zero(mat, wi, wo, nrm, wl, rf) != 0
        f_ref = ref.bsdf_eval(mat, wi, wo, nrm, wl, uv, rf)
        f_hc = hc.bsdf_eval(mat, wi, wo, nrm, wl, uv, rf)
        assert (f_ref[gate] == 0).all() and (f_hc[gate] == 0).all(), name
        assert gate.mean() > 0.3, (name, gate.mean())
        if "smooth" in name or name in ("gold_conductor", "silver_mirror"):
            assert (gate | (f_ref != 0)).mean() > 0.999, name  # nearly every zero of a smooth material is one the predicate names
    ref.close(); hc.close(); sc.close()


def test_eval_shared_by_the_wavelengths_is_bit_identical():
    """pt::mat_eval3 / mat_eval_reflect3 (half vector, D, G computed once for the wavelength paths that share wi: nee_eval and the
    shade kernels) against one pt::mat_eval call per path, for all nine materials plus a textured one, every subset of paths."""
    import ctypes as C
    sc, _ = scenes.cornell(32, 32)
    mi = sc.find_material("silver_mirror")
    m = sc.get_material(mi)
    m.textured = 1
    tex = sc.add_material("textured_silver", m)
    m2 = sc.get_material(sc.find_material("rough_white_conductor"))
    m2.textured = 1
    tex2 = sc.add_material("textured_rough", m2)
    sc.build_tree()
    hc = S.HostCheck(sc)
    hc.L.hc_check_eval3.restype = C.c_long
    rng = np.random.RandomState(31)
    wi, wo, nrm, wl, uv, rf = bsdf_inputs(rng, 30000)
    for mat in list(range(len(b2pt.NAMED_MATERIALS))) + [tex, tex2]:
        bad = hc.L.hc_check_eval3(hc.h, mat, S.fp(wi), S.fp(wo), S.fp(nrm), S.fp(uv), C.c_long(len(wi)))
        assert bad == 0, (mat, bad)
    hc.close(); sc.close()


def test_vertex_level_verdict_never_drops_a_live_light_sample(world):
    """pt::nee_vertex_is_dead (light_kernel gives such vertices no light samples at all) is conservative: wherever it holds,
    every one of 64 random light samples has a summand that the per-sample predicate also knows to be zero."""
    name, sc, ref, hc = world
    rng = np.random.RandomState(21)
    o, d, (p, ws, dist, u4) = scenes.ray_batch(ref, sc, n_pixels=3000, samples=1, seed=9)
    # add rays that start just below the surfaces (vertices seen from the inside)
    prim, t, co, nn, uv = ref.intersect(o[:3000], d[:3000])
    hit = prim >= 0
    inside_o = (co[hit] - nn[hit] * np.float32(1e-3)).astype(np.float32)
    O, D = np.concatenate([o, inside_o]), np.concatenate([d, d[:3000][hit]])
    v, a = hc.nee_dead(O, D, 64, rng)
    assert (a != -1000).all(), "nee_sample_is_dead (all wavelengths at once) disagrees with the per-wavelength predicate"
    has = a >= 0
    assert has.sum() > 1000
    assert not ((v == 1) & (a > 0)).any(), f"{name}: {((v == 1) & (a > 0)).sum()} vertices called dead have live samples"
    assert (v[has] == 1).mean() > 0.1, (name, (v[has] == 1).mean())


def test_textured_reflectance():
    """Checkerboard reflectance (Material.hpp:134-151) through eval on a textured smooth conductor."""
    sc, _ = scenes.two_triangle_scene()
    mi = sc.find_material("silver_mirror")
    m = sc.get_material(mi)
    m.textured = 1
    sc.set_material(mi, m)
    sc.build_tree()
    ref, hc = S.Ref(sc), S.HostCheck(sc)
    rng = np.random.RandomState(2)
    n = 20000
    nrm = np.tile(np.array([[0, 1, 0]], np.float32), (n, 1))
    wi = rng.normal(size=(n, 3)).astype(np.float32)
    wi[:, 1] = np.abs(wi[:, 1]) + 0.1
    wi /= np.linalg.norm(wi, axis=1, keepdims=True)
    wo = wi * np.array([-1, 1, -1], np.float32)
    wl = rng.randint(0, 3, n).astype(np.int32)
    uv = (rng.rand(n, 2) * 1.4 - 0.2).astype(np.float32)
    rf = np.ones(n, np.int32)
    a, b = hc.bsdf_eval(mi, wi, wo, nrm, wl, uv, rf), ref.bsdf_eval(mi, wi, wo, nrm, wl, uv, rf)
    assert rel_close(a, b, 1e-6, 1e-8).all()
    assert len(np.unique(np.round(b, 2))) > 3
    hc.close(); ref.close(); sc.close()


def test_sample_light_bit_exact(world):
    name, sc, ref, hc = world
    u4 = uniforms(np.random.RandomState(12), 20000, 4)
    got, want = hc.sample_light(u4), ref.sample_light(u4)
    for a, b in zip(got, want):
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32))
    # the sqrt-warped triangle selector (BVH.cpp:132): the two triangles of the quad are NOT picked 50/50
    assert len(np.unique(want[3])) >= 1


def test_env_lookup(world):
    name, sc, ref, hc = world
    d = np.random.RandomState(13).normal(size=(20000, 3)).astype(np.float32)
    d[:6] = [[0, 1, 0], [0, -1, 0], [1, 0, 0], [-1, 0, 0], [0, 0, 1], [0, 0, -1]]
    g, r = hc.env(d), ref.sample_env(d)
    assert np.array_equal(g.view(np.uint32), r.view(np.uint32))  # same libm on the host: bit-exact
    if name == "chess_sky_dof":
        assert r.std() > 0.01
    else:
        assert (r == 0).all()


def test_atan2f_acosf_match_the_c_library():
    """Scene::sampleEnv calls glibc's atan2f / acosf (src/Scene.hpp:66-67).  pt::atanf_ref / atan2f_ref / acosf_ref restate those
    routines; here the host compile of them is compared with the C library of this image bit for bit: every float for atanf,
    every float of [-1, 1] (and a stride of the rest) for acosf, 5e7 + 5e7 pairs for atan2f."""
    import ctypes as C
    L = S.hc_lib()
    for f in (L.hc_check_atanf, L.hc_check_acosf):
        f.restype = C.c_ulonglong
        f.argtypes = [C.c_uint, C.c_uint, C.c_uint]
    L.hc_check_atan2f.restype = C.c_ulonglong
    L.hc_check_atan2f.argtypes = [C.c_longlong, C.c_uint, C.c_int]
    assert L.hc_check_atanf(0, 0xFFFFFFFF, 1) == 0
    assert L.hc_check_acosf(0, 0x3F800001, 1) == 0
    assert L.hc_check_acosf(0x80000000, 0xBF800001, 1) == 0
    assert L.hc_check_acosf(0, 0xFFFFFFFF, 97) == 0
    assert L.hc_check_atan2f(50_000_000, 1, 0) == 0  # components of unit vectors, poles and seam over-represented
    assert L.hc_check_atan2f(50_000_000, 2, 1) == 0  # arbitrary bit patterns (zeros, infinities, NaN, denormals)


def test_camera_rays_bit_exact(world):
    name, sc, ref, hc = world
    cam = sc.camera
    px = np.random.RandomState(14).choice(cam.width * cam.height, 1500, replace=False).astype(np.int32)
    o_h, d_h = S.hc_camera_rays(cam, px, 3, 4)
    o_r, d_r = ref.camera_rays(px, 3, 4)
    assert np.array_equal(o_h.view(np.uint32), o_r.view(np.uint32))
    assert np.array_equal(d_h.view(np.uint32), d_r.view(np.uint32))
    if name == "chess_sky_dof":
        assert len(np.unique(o_r[:, 0])) > 100  # thin lens: origins spread over the aperture
    else:
        assert len(np.unique(o_r[:, 0])) == 1


def test_uniform_mapping():
    """pt::unit_from_word and the oracle's b2pt_u01 are what libstdc++'s uniform_real_distribution<float> (through the reference's own
    get_random_float, global.hpp:42-53) yields for an engine word — the FULL 32-bit word, round to nearest, never 1."""
    import ctypes as C
    u = S.hc_stream_uniforms(S.SEED, 99, 7, 0, 0, 64)
    assert (u >= 0).all() and (u < 1).all()
    lib, hc, pto = S.ref_lib(), S.hc_lib(), S.pto_lib()
    lib.ref_uniform_from_word.restype = C.c_float
    lib.ref_uniform_from_word.argtypes = [C.c_uint32]
    hc.hc_u01.restype = C.c_float
    hc.hc_u01.argtypes = [C.c_uint32]
    pto.pto_u01.restype = C.c_float
    pto.pto_u01.argtypes = [C.c_uint32]
    words = [0, 1, 255, 256, 0x7FFFFFFF, 0x80000000, 0xFFFFFF00, 0xFFFFFF7F, 0xFFFFFF80, 0xFFFFFFFF, 0x00FFFFFF, 0x01000001, 0x3FFFFFE0, 0x3FFFFFF0]
    words += [int(x) for x in np.random.RandomState(5).randint(0, 2 ** 32, 4000, dtype=np.uint64)]
    for w in words:
        want = lib.ref_uniform_from_word(w)
        assert hc.hc_u01(w) == want and pto.pto_u01(w) == want, hex(w)
        assert 0.0 <= want < 1.0
    assert lib.ref_uniform_from_word(0xFFFFFFFF) == np.float32(0.99999994)
    # uniforms below 1/2 keep mantissa bits a 24-bit truncation would zero
    small = S.hc_stream_uniforms(S.SEED, 3, 1, 0, 0, 4096)
    small = small[(small < 0.25) & (small > 0)]
    assert (np.frombuffer(small.tobytes(), np.uint32) & 1).mean() > 0.3
    # streams are functions of (seed, pixel, sample, tag, dim) only
    assert np.array_equal(S.hc_stream_uniforms(S.SEED, 99, 7, 0, 5, 10), u[5:15])
    assert not np.array_equal(S.hc_stream_uniforms(S.SEED, 99, 8, 0, 0, 64), u)
    assert not np.array_equal(S.hc_stream_uniforms(S.SEED, 99, 7, 1, 0, 64), u)
    assert abs(float(np.mean(S.hc_stream_uniforms(S.SEED, 5, 5, 0, 0, 4096))) - 0.5) < 0.02
