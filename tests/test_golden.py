"""Golden vectors (tests/golden/*.npz, generated from the reference itself by tests/golden/make_golden.py) replayed
against (CPU) the plain-C restatement oracle/pt_oracle.c and the host build of the device arithmetic, and (GPU) the CUDA
path through the C ABI.  Nothing here needs /root/reference or oracle/_ref at run time."""
import os

import numpy as np
import pytest

import support as S
from gen import rel_close
from golden.make_golden import golden_scenes

b2pt = S.b2pt
G = S.GOLDEN


def load(name):
    return np.load(os.path.join(G, name + ".npz"))


def bits(a):
    return a.view(np.uint64) if a.dtype == np.float64 else a.view(np.uint32)


# ---------------------------------------------------------------- CPU: restatement and hostcheck --------------------
def test_primitives_cpu():
    g = load("primitives")
    hit, t = S.pto_tri(g["tri_v"], g["tri_o"], g["tri_d"])
    assert np.array_equal(hit, g["tri_hit"]) and np.array_equal(bits(t[hit == 1]), bits(g["tri_t"][hit == 1]))
    hit, t = S.hc_tri(g["tri_v"], g["tri_o"], g["tri_d"])
    assert np.array_equal(hit, g["tri_hit"]) and np.array_equal(bits(t[hit == 1]), bits(g["tri_t"][hit == 1]))
    assert np.array_equal(S.hc_box(g["box_b"], g["box_o"], g["box_d"]), g["box_hit"])
    hit, t = S.hc_sphere(g["sph_c"], g["sph_o"], g["sph_d"])
    assert np.array_equal(hit, g["sph_hit"]) and np.array_equal(bits(t[hit == 1]), bits(g["sph_t"][hit == 1]))


@pytest.mark.parametrize("name", ["cornell", "chess_sky_dof", "chess_dark"])
def test_scene_cpu(name):
    g = load(name)
    sc, env = golden_scenes()[name]()
    pto, hc = S.Restated(sc), S.HostCheck(sc)
    for impl in (pto, hc):
        prim, t = impl.intersect(g["ray_o"], g["ray_d"])
        assert np.array_equal(prim, g["prim"]) and np.array_equal(bits(t), bits(g["t"]))
        for a, b in zip(impl.sample_light(g["u4"]), (g["light_p"], g["light_n"], g["light_e"], g["light_pdf"])):
            assert np.array_equal(bits(a), bits(b))
    assert np.array_equal(hc.shadow(g["sh_o"], g["sh_d"], g["sh_dist"]), g["sh_visible"])
    assert np.array_equal(hc.shadow4(g["sh_o"], g["sh_d"], g["sh_dist"]), g["sh_visible"])
    assert np.array_equal(bits(pto.sample_env(g["env_d"])), bits(g["env_rgb"]))
    assert np.array_equal(bits(hc.env(g["env_d"])), bits(g["env_rgb"]))
    o, d = pto.camera_rays(g["pixels"], 2, 3)
    assert np.array_equal(bits(o), bits(g["cam_o"])) and np.array_equal(bits(d), bits(g["cam_d"]))
    rad = pto.render_samples(g["pixels"], 0, 4, seed=int(g["seed"]))
    assert rel_close(rad, g["radiance"], 1e-5, 1e-7).all()
    pto.close(); hc.close(); sc.close()


def test_bsdf_cpu():
    import scenes
    g = load("bsdf")
    sc, _ = scenes.two_triangle_scene()
    pto, hc = S.Restated(sc), S.HostCheck(sc)
    wi, wo, n, wl, uv, rf = (g[k] for k in ("wi", "wo", "n", "wl", "uv", "rf"))
    for mat in range(len(b2pt.NAMED_MATERIALS)):
        for impl in (pto, hc):
            assert rel_close(impl.bsdf_eval(mat, wi, wo, n, wl, uv, rf), g[f"eval_{mat}"], 1e-5, 1e-7).all()
            assert rel_close(impl.bsdf_pdf(mat, wi, wo, n, wl, rf), g[f"pdf_{mat}"], 1e-5, 1e-7).all()
        assert rel_close(hc.fresnel(mat, wi, n, wl), g[f"fresnel_{mat}"], 1e-5, 1e-7).all()
        assert np.array_equal(bits(hc.refract(mat, wi, n, wl)), bits(g[f"refract_{mat}"]))
        assert np.array_equal(bits(hc.material_sample(mat, n, g["u2"])), bits(g[f"sample_{mat}"]))
    assert np.array_equal(bits(hc.reflect(wi, n)), bits(g["reflect"]))
    pto.close(); hc.close(); sc.close()


# ---------------------------------------------------------------- GPU: the CUDA path through the C ABI ---------------
@pytest.fixture(scope="module")
def ctx():
    c = b2pt.Context(0)
    yield c
    c.close()


@pytest.mark.gpu
def test_primitives_gpu(ctx):
    g = load("primitives")
    hit, t = ctx.tri_intersect(g["tri_v"], g["tri_o"], g["tri_d"])
    assert np.array_equal(hit, g["tri_hit"]) and np.array_equal(bits(t[hit == 1]), bits(g["tri_t"][hit == 1]))
    assert np.array_equal(ctx.box_intersect(g["box_b"], g["box_o"], g["box_d"]), g["box_hit"])
    hit, t, p, n = ctx.sphere_intersect(g["sph_c"], g["sph_o"], g["sph_d"])
    h = hit == 1
    assert np.array_equal(hit, g["sph_hit"]) and np.array_equal(bits(t[h]), bits(g["sph_t"][h]))
    assert np.array_equal(bits(p[h]), bits(g["sph_p"][h])) and np.array_equal(bits(n[h]), bits(g["sph_n"][h]))


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["cornell", "chess_sky_dof", "chess_dark"])
def test_scene_gpu(ctx, name):
    g = load(name)
    sc, env = golden_scenes()[name]()
    ctx.upload(sc)
    prim, t = ctx.intersect(g["ray_o"], g["ray_d"])
    assert np.array_equal(prim, g["prim"]) and np.array_equal(bits(t), bits(g["t"]))  # (a) hit index and t bit-exact
    assert np.array_equal(ctx.shadow(g["sh_o"], g["sh_d"], g["sh_dist"]), g["sh_visible"])
    for a, b in zip(ctx.sample_light(g["u4"]), (g["light_p"], g["light_n"], g["light_e"], g["light_pdf"])):
        assert np.array_equal(bits(a), bits(b))
    assert np.abs(ctx.env_lookup(g["env_d"]) - g["env_rgb"]).max() <= 2e-3
    o, d = ctx.camera_rays(sc.camera, g["pixels"], 2, 3, seed=int(g["seed"]))
    assert np.array_equal(bits(o), bits(g["cam_o"])) and np.array_equal(bits(d), bits(g["cam_d"]))
    rad, _ = ctx.render_samples(sc.camera, g["pixels"], 0, 4, seed=int(g["seed"]))
    ok = rel_close(rad, g["radiance"], 2e-4, 1e-5)  # (c) per-sample radiance on shared sample streams
    assert ok.mean() >= 0.999, f"{(~ok).sum()} of {ok.size} differ"
    sc.close()


@pytest.mark.gpu
def test_bsdf_gpu(ctx):
    import scenes
    g = load("bsdf")
    sc, _ = scenes.two_triangle_scene()
    ctx.upload(sc)
    wi, wo, n, wl, uv, rf = (g[k] for k in ("wi", "wo", "n", "wl", "uv", "rf"))
    for mat in range(len(b2pt.NAMED_MATERIALS)):
        assert rel_close(ctx.bsdf_eval(mat, wi, wo, n, wl, uv, rf), g[f"eval_{mat}"], 1e-5, 1e-7).all()  # (b) 1e-5 relative
        assert rel_close(ctx.bsdf_pdf(mat, wi, wo, n, wl, rf), g[f"pdf_{mat}"], 1e-5, 1e-7).all()
        assert rel_close(ctx.fresnel(mat, wi, n, wl), g[f"fresnel_{mat}"], 1e-5, 1e-7).all()
        assert np.array_equal(bits(ctx.refract(mat, wi, n, wl)), bits(g[f"refract_{mat}"]))
        assert np.array_equal(bits(ctx.material_sample(mat, wo, n, g["u2"])), bits(g[f"sample_{mat}"]))
    sc.close()
