"""Randomised scenes (random triangle soups, spheres, material parameters, one or two lights, sky or constant background,
DoF on/off, rr and NEE counts varied): the plain-C restatement (CPU) and the CUDA path (GPU) replay the reference's
castRay per sample on shared sample streams."""
import numpy as np
import pytest

import scenes
import support as S
from gen import rel_close

b2pt = S.b2pt
need_ref = pytest.mark.skipif(not S.have_ref(), reason="oracle/_ref/libref_oracle.so not built")


@need_ref
@pytest.mark.parametrize("seed", range(6))
def test_restatement_on_random_scenes(seed):
    sc, env = scenes.random_scene(100 + seed)
    ref, pto = S.Ref(sc, env), S.Restated(sc)
    cam = sc.camera
    px = np.random.RandomState(seed).choice(cam.width * cam.height, 250, replace=False).astype(np.int32)
    a, b = pto.render_samples(px, 0, 4), ref.render_samples(px, 0, 4)
    ok = rel_close(a, b, 1e-5, 1e-7)
    assert ok.all(), f"seed {seed}: {(~ok).sum()} of {ok.size} differ"
    o, d, _ = scenes.ray_batch(ref, sc, n_pixels=300, samples=1, seed=seed)
    pr, tr, *_ = ref.intersect(o, d)
    pp, tp = pto.intersect(o, d)
    assert np.array_equal(pp, pr) and np.array_equal(tp.view(np.uint64), tr.view(np.uint64))
    pto.close(); ref.close(); sc.close()


@need_ref
@pytest.mark.gpu
def test_cuda_on_random_scenes():
    ctx = b2pt.Context(0)
    total = bad = 0
    for seed in range(24):
        sc, env = scenes.random_scene(200 + seed)
        ref = S.Ref(sc, env)
        ctx.upload(sc)
        cam = sc.camera
        px = np.random.RandomState(seed).choice(cam.width * cam.height, 400, replace=False).astype(np.int32)
        g, _ = ctx.render_samples(cam, px, 0, 6)
        r = ref.render_samples(px, 0, 6)
        ok = rel_close(g, r, 2e-4, 1e-5)
        total += ok.size
        bad += int((~ok).sum())
        assert ok.mean() >= 0.998, f"seed {seed}: {(~ok).sum()} of {ok.size} per-sample radiances differ"
        o, d, (p, ws, dist, u4) = scenes.ray_batch(ref, sc, n_pixels=800, samples=1, seed=seed)
        pr, tr, *_ = ref.intersect(o, d)
        pg, tg = ctx.intersect(o, d)
        assert np.array_equal(pg, pr) and np.array_equal(tg.view(np.uint64), tr.view(np.uint64)), seed
        ref.close(); sc.close()
    assert bad <= total * 2e-4, (bad, total)
    ctx.close()
