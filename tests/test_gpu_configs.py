"""GPU parity on the BASELINE configurations the small-scene tests do not reach ("untested configurations"):

* C4, the 296 k-triangle glass-gem scene: Scene::intersect hit id + t bits on camera / continuation / shadow rays, and
  per-sample radiance — the SAH tree and its four-wide collapse are what changes with scene size;
* C2 at the north-star light-sample count (32): per-sample replay on the chess scene, and the ray accounting of SURVEY 8d
  (`rays_reference`) against the oracle's own count for the same paths;
* C5's frame size (1024 x 1024 Cornell box) on a pixel subset.
"""
import numpy as np
import pytest

import scenes
import support as S

b2pt = S.b2pt
pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not S.have_ref(), reason="oracle/_ref/libref_oracle.so not built")]


@pytest.fixture(scope="module")
def ctx():
    c = b2pt.Context(0)
    yield c
    c.close()


def close_per_sample(got, want):
    return np.abs(got - want) <= 1e-5 + 2e-4 * np.maximum(np.abs(got), np.abs(want))


def test_high_poly_gem_scene_hits_and_radiance(ctx):
    """configs[3]: high_king / high_soldier (296 242 triangles), every piece smooth_glass_gem."""
    sc, env = scenes.chess(320, 180, dof=True, sky=True, quality="high", fix=b2pt.FIX_MODEL_QUALITY, king="smooth_glass_gem",
                           left="smooth_glass_gem", right="smooth_glass_gem")
    assert sc.desc.n_prims > 290000
    ref = S.Ref(sc, env)
    ctx.upload(sc)
    o, d, (p, ws, dist, u4) = scenes.ray_batch(ref, sc, n_pixels=40000, samples=2, seed=3)
    assert len(o) >= 100000
    prim_r, t_r, _, _, _ = ref.intersect(o, d)
    prim_g, t_g, st = ctx.intersect(o, d, count=True)
    assert np.array_equal(prim_g, prim_r)
    hit = prim_r >= 0
    assert 0.3 < hit.mean() < 0.9
    assert np.array_equal(t_g[hit].view(np.uint64), t_r[hit].view(np.uint64))
    # the walk stays logarithmic on the big tree (exhaustive reference walk: > 100 boxes per ray here)
    assert st.extend_nodes / len(o) < 80, st.extend_nodes / len(o)
    # rays that leave the glass pieces from inside (refraction chains start there): origins on hit points, pushed inwards
    pin = (o[: len(prim_r)][hit] + d[: len(prim_r)][hit] * t_r[hit, None].astype(np.float32)).astype(np.float32)
    rng = np.random.RandomState(9)
    v = rng.normal(size=pin.shape).astype(np.float32)
    v /= np.linalg.norm(v, axis=1, keepdims=True).astype(np.float32)
    prim_r2, t_r2, _, _, _ = ref.intersect(pin, v)
    prim_g2, t_g2 = ctx.intersect(pin, v)
    assert np.array_equal(prim_g2, prim_r2)
    h2 = prim_r2 >= 0
    assert np.array_equal(t_g2[h2].view(np.uint64), t_r2[h2].view(np.uint64))
    # visibility decisions towards the light
    vis_g = ctx.shadow(p, ws, dist)
    prim_s, t_s, _, _, _ = ref.intersect(p, ws)
    vis_r = ((prim_s >= 0) & (np.abs(t_s - dist.astype(np.float64)) < 1e-4)).astype(np.int32)
    assert np.array_equal(vis_g, vis_r)
    # per-sample radiance: deep dielectric chains (dispersion splits the three wavelengths)
    px = np.random.RandomState(5).choice(320 * 180, 200, replace=False).astype(np.int32)
    got, st = ctx.render_samples(sc.camera, px, 0, 4)
    want = ref.render_samples(px, 0, 4)
    ok = close_per_sample(got, want)
    assert ok.mean() >= 0.999, ok.mean()
    assert st.max_depth >= 4
    ref.close()
    sc.close()


def test_chess_with_32_light_samples_replay_and_ray_accounting(ctx):
    """configs[1] as north_star writes it: 32 next-event samples per vertex (conf.json:23 through setDirectLightSample)."""
    sc, env = scenes.chess(160, 90, dof=True, sky=True, n_dir=32)
    assert sc.desc.n_dir_sample == 32
    ref = S.Ref(sc, env)
    ctx.upload(sc)
    px = np.arange(0, 160 * 90, 3, dtype=np.int32)
    spp = 4
    got, st = ctx.render_samples(sc.camera, px, 0, spp)
    want = ref.render_samples(px, 0, spp)
    ok = close_per_sample(got, want)
    assert ok.mean() >= 0.999, ok.mean()
    # ray accounting: the oracle restatement counts the rays the reference algorithm needs for exactly these paths
    rs = S.Restated(sc)
    want2, rays, verts = rs.render_samples_counted(px, 0, spp)
    assert close_per_sample(want2, want).mean() >= 0.999
    paths = len(px) * spp * 3
    assert st.paths == paths
    rpp_oracle, rpp_gpu = rays / paths, st.rays_reference / paths
    assert 13.0 < rpp_oracle < 16.0, rpp_oracle  # 14.4 - 14.6 at 1080p (profiles/)
    # the GPU never counts a ray the reference does not need; it misses only the continuation rays of paths it ends early
    # because their pixel value is already settled (DESIGN.md "settled paths"), a per-cent effect
    assert st.rays_reference <= rays
    assert st.rays_reference >= 0.97 * rays, (rpp_gpu, rpp_oracle)
    assert st.vertices_shaded <= verts and st.vertices_shaded >= 0.97 * verts
    # what is really traced is a small part of that: mirrors and glass get no light samples at all
    assert st.rays_traced_closest + st.rays_traced_shadow < 0.5 * st.rays_reference
    rs.close()
    ref.close()
    sc.close()


def test_cornell_1024_subset_replay(ctx):
    """configs[4] frame size: the 1024 x 1024 Cornell box of the material sweep, per-sample replay on a pixel subset, for the
    slowest (clear rough plastic: rough dielectric everywhere) and a dispersive (glass) member of the sweep."""
    for name in ("clear_rough_plastic", "smooth_glass"):
        sc, _ = scenes.cornell_sweep(name, 1024, 1024)
        ref = S.Ref(sc)
        ctx.upload(sc)
        px = np.random.RandomState(21).choice(1024 * 1024, 1500, replace=False).astype(np.int32)
        got, _ = ctx.render_samples(sc.camera, px, 0, 4)
        want = ref.render_samples(px, 0, 4)
        assert close_per_sample(got, want).mean() >= 0.999
        assert want.mean() > 0.05
        ref.close()
        sc.close()
