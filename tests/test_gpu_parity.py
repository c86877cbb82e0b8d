"""GPU parity tests: the CUDA path, called through the C ABI (include/b2pt.h), against the oracle
(the reference's own sources, oracle/_ref/libref_oracle.so) on identical seeded inputs.

Bars (BASELINE.json north_star): (a) intersection hit index and t bit-exact; (b) BSDF eval/pdf within
1e-5 relative; (c) per-sample radiance on shared sample streams within 2e-4 relative (+1e-5 absolute).
"""
import numpy as np
import pytest

import scenes
import support as S
from gen import adversarial_triangle_cases, bsdf_inputs, box_cases, degenerate_rays, rel_close, sphere_cases, uniforms

b2pt = S.b2pt
pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not S.have_ref(), reason="oracle/_ref/libref_oracle.so not built")]


@pytest.fixture(scope="module")
def ctx():
    c = b2pt.Context(0)
    yield c
    c.close()


@pytest.fixture(scope="module", params=["cornell", "chess_sky_dof", "chess_dark"])
def world(request, ctx):
    if request.param == "cornell":
        sc, env = scenes.cornell(96, 96)
    elif request.param == "chess_sky_dof":
        sc, env = scenes.chess(160, 90, dof=True, sky=True)
    else:
        sc, env = scenes.chess(160, 90, dof=False, sky=False)
    ref = S.Ref(sc, env)
    ctx.upload(sc)
    yield request.param, sc, ref
    ref.close()
    sc.close()




# ---- (a) primitives: bit-exact -------------------------------------------------------------------------


def test_triangle_bit_exact(ctx):
    rng = np.random.RandomState(7)
    v, o, d = adversarial_triangle_cases(rng, 200000)
    hit_g, t_g = ctx.tri_intersect(v, o, d)
    hit_r, t_r = S.ref_tri(v, o, d)
    assert np.array_equal(hit_g, hit_r)
    assert 0.2 < hit_r.mean() < 0.95
    assert np.array_equal(t_g[hit_r == 1].view(np.uint64), t_r[hit_r == 1].view(np.uint64))


def test_box_bit_exact(ctx):
    rng = np.random.RandomState(8)
    b6, o, d = box_cases(rng, 200000)
    hit_g = ctx.box_intersect(b6, o, d)
    hit_r = S.ref_box(b6, o, d)
    assert np.array_equal(hit_g, hit_r)
    assert 0.02 < hit_r.mean() < 0.9


def test_sphere_bit_exact(ctx):
    rng = np.random.RandomState(9)
    c4, o, d = sphere_cases(rng, 100000)
    hit_g, t_g, co_g, nn_g = ctx.sphere_intersect(c4, o, d)
    hit_r, t_r, co_r, nn_r = S.ref_sphere(c4, o, d)
    assert np.array_equal(hit_g, hit_r)
    h = hit_r == 1
    assert np.array_equal(t_g[h].view(np.uint64), t_r[h].view(np.uint64))
    assert np.array_equal(co_g[h].view(np.uint32), co_r[h].view(np.uint32))
    assert np.array_equal(nn_g[h].view(np.uint32), nn_r[h].view(np.uint32))


# ---- (a) scene intersection: hit primitive and t bit-exact ------------------------------------------------
def test_scene_intersect_bit_exact(ctx, world):
    name, sc, ref = world
    o, d, _ = scenes.ray_batch(ref, sc, n_pixels=3000, samples=2, seed=3)
    prim_r, t_r, *_ = ref.intersect(o, d)
    prim_g, t_g, st = ctx.intersect(o, d, count=True)
    assert np.array_equal(prim_g, prim_r), f"{name}: {(prim_g != prim_r).sum()} of {len(o)} hit ids differ"
    assert np.array_equal(t_g.view(np.uint64), t_r.view(np.uint64))
    assert (prim_r >= 0).mean() > 0.3
    assert st.nodes_fetched > 0 and st.prims_tested > 0


def test_degenerate_rays_take_the_reference_tree(ctx, world):
    """Zero / denormal direction components and origins on box planes: slab products are NaN or infinite."""
    name, sc, ref = world
    root = sc.desc.nodes[0]
    o, d = degenerate_rays(np.random.RandomState(31), list(root.bmin), list(root.bmax), 20000)
    prim_r, t_r, *_ = ref.intersect(o, d)
    prim_g, t_g = ctx.intersect(o, d)
    assert np.array_equal(prim_g, prim_r), f"{name}: {(prim_g != prim_r).sum()} hit ids differ"
    assert np.array_equal(t_g.view(np.uint64), t_r.view(np.uint64))


def test_shadow_decision(ctx, world):
    name, sc, ref = world
    _, _, (p, ws, dist, u4) = scenes.ray_batch(ref, sc, n_pixels=3000, samples=2, seed=4)
    prim_r, t_r, *_ = ref.intersect(p, ws)
    want = ((prim_r >= 0) & (np.abs(t_r - dist.astype(np.float64)) < np.float64(np.float32(1e-4)))).astype(np.int32)
    got = ctx.shadow(p, ws, dist)
    assert np.array_equal(got, want), f"{name}: {(got != want).sum()} of {len(want)} visibility decisions differ"
    assert 0.01 < want.mean() < 0.99


# ---- (b) BSDF: 1e-5 relative ---------------------------------------------------------------------------------


def test_bsdf_parity(ctx, world):
    name, sc, ref = world
    if name != "cornell":
        pytest.skip("material table is the same in every scene")
    rng = np.random.RandomState(11)
    n = 40000
    wi, wo, nrm, wl, uv, rf = bsdf_inputs(rng, n)
    for mat in range(len(b2pt.NAMED_MATERIALS)):
        e_r, e_g = ref.bsdf_eval(mat, wi, wo, nrm, wl, uv, rf), ctx.bsdf_eval(mat, wi, wo, nrm, wl, uv, rf)
        assert rel_close(e_g, e_r, 1e-5, 1e-7).all(), (b2pt.NAMED_MATERIALS[mat], "eval")
        p_r, p_g = ref.bsdf_pdf(mat, wi, wo, nrm, wl, rf), ctx.bsdf_pdf(mat, wi, wo, nrm, wl, rf)
        assert rel_close(p_g, p_r, 1e-5, 1e-7).all(), (b2pt.NAMED_MATERIALS[mat], "pdf")
        f_r, f_g = ref.fresnel(mat, wi, nrm, wl), ctx.fresnel(mat, wi, nrm, wl)
        assert rel_close(f_g, f_r, 1e-5, 1e-7).all(), (b2pt.NAMED_MATERIALS[mat], "fresnel")
        r_r, r_g = ref.refract(mat, wi, nrm, wl), ctx.refract(mat, wi, nrm, wl)
        assert np.array_equal(r_g.view(np.uint32), r_r.view(np.uint32)), (b2pt.NAMED_MATERIALS[mat], "refract")
        u2 = uniforms(rng, n, 2)
        s_r, s_g = ref.material_sample(mat, wo, nrm, u2), ctx.material_sample(mat, wo, nrm, u2)
        assert np.array_equal(s_g.view(np.uint32), s_r.view(np.uint32)), (b2pt.NAMED_MATERIALS[mat], "sample")
    assert np.array_equal(ctx.reflect(wi, nrm).view(np.uint32), ref.reflect(0, wi, nrm).view(np.uint32))


# ---- light sampling, env lookup, camera rays, sample streams ----------------------------------------------------
def test_sample_light_bit_exact(ctx, world):
    name, sc, ref = world
    u4 = uniforms(np.random.RandomState(12), 50000, 4)
    for a, b in zip(ctx.sample_light(u4), ref.sample_light(u4)):
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32))


def test_env_lookup(ctx, world):
    name, sc, ref = world
    d = np.random.RandomState(13).normal(size=(50000, 3)).astype(np.float32)
    d[:10] = [[0, 1, 0], [0, -1, 0], [1, 0, 0], [-1, 0, 0], [0, 0, 1], [0, 0, -1], [-1, 0, 1e-8], [-1, 0, -1e-8], [1e-8, 1, 0], [0, -1, 1e-8]]
    g, r = ctx.env_lookup(d), ref.sample_env(d)
    # atan2f / acosf are the C library's algorithms restated (pt::atan2f_ref / acosf_ref, bit-exact against glibc on the host:
    # tests/test_cpu_math.py::test_atan2f_acosf_match_the_c_library), so the texel position is the reference's: SURVEY 8c(4) asks 1e-6
    assert np.abs(g - r).max() <= 1e-6


def test_camera_rays_bit_exact(ctx, world):
    name, sc, ref = world
    cam = sc.camera
    px = np.random.RandomState(14).choice(cam.width * cam.height, 2000, replace=False).astype(np.int32)
    o_g, d_g = ctx.camera_rays(cam, px, 3, 4)
    o_r, d_r = ref.camera_rays(px, 3, 4)
    assert np.array_equal(o_g.view(np.uint32), o_r.view(np.uint32))
    assert np.array_equal(d_g.view(np.uint32), d_r.view(np.uint32))


def test_stream_uniforms(ctx):
    """Device Philox streams against the ORACLE's own implementation (oracle/b2pt_portable.h, pinned to the published Philox4x32-10
    known answers by tests/test_oracle.py) — not against a host compile of the device header."""
    for seed, pixel, sample, tag, dim0, n in ((S.SEED, 1234, 56, 0, 3, 50), (S.SEED, 0, 0, 1, 0, 9), (0xDEADBEEFCAFEF00D, 2073599, 2047, 0, 1021, 40),
                                              (1, 77, 4095, 1, 2, 7)):
        g = ctx.stream_uniforms(seed, pixel, sample, tag, dim0, n)
        o = S.pto_stream_uniforms(seed, pixel, sample, tag, dim0, n)
        assert np.array_equal(g.view(np.uint32), o.view(np.uint32))
        assert np.array_equal(g, S.hc_stream_uniforms(seed, pixel, sample, tag, dim0, n))
        assert (g >= 0).all() and (g < 1).all()


def test_textured_reflectance(ctx, world):
    """Material::getReflectance's checkerboard (src/Material.hpp:134-151) through eval on the device: the chess floor is the
    textured silver mirror (floor_isTextured), Schlick's F0 comes from (u, v).  uv values sit on and around the cell borders."""
    name, sc, ref = world
    textured = [i for i, (_, m) in enumerate(sc.materials()) if m.textured]
    if name == "cornell":
        assert not textured
        pytest.skip("no textured material in the DEMO scene")
    assert textured
    rng = np.random.RandomState(17)
    n = 60000
    wi, wo, nrm, wl, uv, rf = bsdf_inputs(rng, n)
    wo[: n // 2] = (2 * (wi[: n // 2] * nrm[: n // 2]).sum(1, keepdims=True) * nrm[: n // 2] - wi[: n // 2]).astype(np.float32)  # exact mirror pairs: F is returned
    rf[:] = 1
    # uv on the checker's cell borders (col = int((u - .05) * 10), row = int(v * 12)) and outside [0, 1]
    k = n // 4
    uv[:k, 0] = (0.05 + rng.randint(-1, 12, k) / 10.0 + rng.choice([-1e-7, 0.0, 1e-7], k)).astype(np.float32)
    uv[:k, 1] = (rng.randint(-1, 14, k) / 12.0 + rng.choice([-1e-7, 0.0, 1e-7], k)).astype(np.float32)
    for mat in textured:
        e_r, e_g = ref.bsdf_eval(mat, wi, wo, nrm, wl, uv, rf), ctx.bsdf_eval(mat, wi, wo, nrm, wl, uv, rf)
        assert rel_close(e_g, e_r, 1e-5, 1e-7).all()
        nz = e_r > 0
        assert nz.mean() > 0.1
        # both checker values occur (0.9 / 0.1 cells), so the test sees the texture and not a constant F0
        assert len(np.unique(np.round(e_r[nz], 2))) > 5


# ---- (c) radiance per sample on shared sample streams -------------------------------------------------------------
def compare_samples(name, g, r, frac_ok=0.999):
    ok = rel_close(g, r, 2e-4, 1e-5)
    bad = (~ok).sum()
    assert ok.mean() >= frac_ok, f"{name}: {bad} of {ok.size} per-sample radiances differ; worst |d| = {np.nanmax(np.abs(g - r)):.3g}"
    # the few that may differ (a libm ulp flipping a branch) must not move the mean
    assert abs(np.nanmean(g) - np.nanmean(r)) <= 2e-3 * max(abs(np.nanmean(r)), 1e-3)


def test_radiance_per_sample(ctx, world):
    name, sc, ref = world
    cam = sc.camera
    rng = np.random.RandomState(21)
    px = rng.choice(cam.width * cam.height, 700, replace=False).astype(np.int32)
    spp = 8
    g, st = ctx.render_samples(cam, px, 0, spp)
    r = ref.render_samples(px, 0, spp)
    assert st.paths == 3 * len(px) * spp
    assert np.isfinite(r).all()
    compare_samples(name, g, r)
    # tracing the three wavelengths as separate rays gives the same numbers
    g3, st3 = ctx.render_samples(cam, px, 0, spp, flags=b2pt.FLAG_SPLIT_WAVELENGTHS)
    assert np.array_equal(g3.view(np.uint32), g.view(np.uint32))
    assert st3.rays_reference == st.rays_reference
    assert st3.rays_traced_closest >= st.rays_traced_closest


def test_frame_matches_oracle_frame(ctx, world):
    name, sc, ref = world
    cam = sc.camera
    spp = 4
    fb_g, st = ctx.render(cam, spp)
    fb_r = ref.render_frame(0, spp, spp)
    assert fb_g.shape == (cam.height, cam.width, 3)
    ok = rel_close(fb_g, fb_r, 5e-4, 2e-5)
    assert ok.mean() > 0.998, f"{name}: {(~ok).sum()} of {ok.size} framebuffer values differ"
    assert abs(fb_g.mean() - fb_r.mean()) <= 1e-3 * fb_r.mean()
    # sample ranges compose: two half-frames accumulate to the same image (multi-GPU split)
    half, _ = ctx.render(cam, spp, sample_begin=0, sample_count=2)
    half, _ = ctx.render(cam, spp, sample_begin=2, sample_count=2, out=half)
    assert np.allclose(half, fb_g, rtol=1e-5, atol=1e-6)
    # FRESH_FRAME overwrites whatever the caller's buffer held
    junk = np.full_like(fb_g, 123.0)
    fresh, _ = ctx.render(cam, spp, out=junk, flags=b2pt.FLAG_FRESH_FRAME)
    assert fresh is junk and np.allclose(fresh, fb_g, rtol=1e-5, atol=1e-6)
    # small waves give the same frame as one big wave
    small, st2 = ctx.render(cam, spp, max_wave_bundles=4096)
    assert st2.waves > st.waves
    assert np.allclose(small, fb_g, rtol=1e-5, atol=1e-6)
    assert st2.rays_reference == st.rays_reference


def test_no_shadow_and_ndir(ctx):
    sc, _ = scenes.cornell(48, 48, n_dir=7, rr=0.55)
    ref = S.Ref(sc)
    ctx.upload(sc)
    px = np.arange(0, 48 * 48, 5, dtype=np.int32)
    compare_samples("ndir7", ctx.render_samples(sc.camera, px, 0, 4)[0], ref.render_samples(px, 0, 4))
    ctx.set_params(enable_shadow=0)
    ref.set_params(enable_shadow=False)
    compare_samples("noshadow", ctx.render_samples(sc.camera, px, 0, 4)[0], ref.render_samples(px, 0, 4))
    ref.close()
    sc.close()


def test_errors(ctx):
    c2 = b2pt.Context(0)
    with pytest.raises(RuntimeError):
        c2.render(scenes.cornell(8, 8)[0].camera, 1)  # no scene uploaded
    with pytest.raises(RuntimeError):
        b2pt.Context(99)
    c2.close()
