"""Pins the plain-C restatement (oracle/pt_oracle.c) against the reference itself (oracle/_ref: the reference's
unmodified sources): hit ids and t bit-exact, BSDF values, light samples, camera rays, and per-sample radiance of
Scene::castRay on shared sample streams.  CPU-only."""
import numpy as np
import pytest

import scenes
import support as S
from gen import adversarial_triangle_cases, bsdf_inputs, rel_close, uniforms

b2pt = S.b2pt
pytestmark = pytest.mark.skipif(not S.have_ref(), reason="oracle/_ref/libref_oracle.so not built")


@pytest.fixture(scope="module", params=["cornell", "chess_sky_dof"])
def world(request):
    sc, env = scenes.cornell(64, 64) if request.param == "cornell" else scenes.chess(96, 54, dof=True, sky=True)
    ref, pto = S.Ref(sc, env), S.Restated(sc)
    yield request.param, sc, ref, pto
    pto.close(); ref.close(); sc.close()


def test_triangle():
    v, o, d = adversarial_triangle_cases(np.random.RandomState(3), 50000)
    hit_p, t_p = S.pto_tri(v, o, d)
    hit_r, t_r = S.ref_tri(v, o, d)
    assert np.array_equal(hit_p, hit_r)
    assert np.array_equal(t_p[hit_r == 1].view(np.uint64), t_r[hit_r == 1].view(np.uint64))


def test_intersect(world):
    name, sc, ref, pto = world
    o, d, _ = scenes.ray_batch(ref, sc, n_pixels=800, samples=1, seed=9)
    prim_r, t_r, *_ = ref.intersect(o, d)
    prim_p, t_p = pto.intersect(o, d)
    assert np.array_equal(prim_p, prim_r)
    assert np.array_equal(t_p.view(np.uint64), t_r.view(np.uint64))


def test_bsdf(world):
    name, sc, ref, pto = world
    if name != "cornell":
        pytest.skip("same material table")
    wi, wo, nrm, wl, uv, rf = bsdf_inputs(np.random.RandomState(4), 10000)
    for mat in range(len(b2pt.NAMED_MATERIALS)):
        assert rel_close(pto.bsdf_eval(mat, wi, wo, nrm, wl, uv, rf), ref.bsdf_eval(mat, wi, wo, nrm, wl, uv, rf), 1e-6, 1e-9).all()
        assert rel_close(pto.bsdf_pdf(mat, wi, wo, nrm, wl, rf), ref.bsdf_pdf(mat, wi, wo, nrm, wl, rf), 1e-6, 1e-9).all()


def test_light_env_camera(world):
    name, sc, ref, pto = world
    u4 = uniforms(np.random.RandomState(5), 5000, 4)
    for a, b in zip(pto.sample_light(u4), ref.sample_light(u4)):
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32))
    d = np.random.RandomState(6).normal(size=(5000, 3)).astype(np.float32)
    assert np.array_equal(pto.sample_env(d).view(np.uint32), ref.sample_env(d).view(np.uint32))
    px = np.arange(0, sc.camera.width * sc.camera.height, 11, dtype=np.int32)
    for a, b in zip(pto.camera_rays(px, 1, 3), ref.camera_rays(px, 1, 3)):
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32))


def test_radiance_per_sample(world):
    """castRay of the restatement == castRay of the reference on the same sample streams."""
    name, sc, ref, pto = world
    cam = sc.camera
    px = np.random.RandomState(7).choice(cam.width * cam.height, 300, replace=False).astype(np.int32)
    a, b = pto.render_samples(px, 0, 6), ref.render_samples(px, 0, 6)
    ok = rel_close(a, b, 1e-5, 1e-7)
    assert ok.all(), f"{name}: {(~ok).sum()} of {ok.size} differ, worst {np.abs(a - b).max()}"
    assert b.mean() > 1e-3


def test_philox_known_answers():
    """The oracle's Philox4x32-10 (oracle/b2pt_portable.h) against the known-answer vectors published with Random123
    (kat_vectors: philox4x32 10 rounds), and the device header's host compile against the oracle's streams."""
    kat = [([0, 0, 0, 0], [0, 0], [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]),
           ([0xFFFFFFFF] * 4, [0xFFFFFFFF] * 2, [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]),
           ([0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344], [0xA4093822, 0x299F31D0], [0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1])]
    for ctr, key, want in kat:
        assert S.pto_philox_block(ctr, key) == want
    for seed, pixel, sample, tag, dim0, n in ((S.SEED, 1234, 56, 0, 3, 50), (0xDEADBEEFCAFEF00D, 2073599, 2047, 1, 1021, 40)):
        a = S.pto_stream_uniforms(seed, pixel, sample, tag, dim0, n)
        b = S.hc_stream_uniforms(seed, pixel, sample, tag, dim0, n)
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32))
