"""Several GPUs of one box from one host thread (b2pt_group_render, ./RayTracing --gpus N): the frame equals the single-GPU
frame — sample streams are keyed by the global sample index.  Skipped on single-GPU boxes."""
import os
import subprocess

import numpy as np
import pytest

import scenes
import support as S

b2pt = S.b2pt
pytestmark = [pytest.mark.gpu, pytest.mark.skipif(b2pt.device_count() < 2, reason="needs at least 2 GPUs")]


def test_group_render_equals_single_gpu():
    n = min(b2pt.device_count(), 4)
    sc, env = scenes.chess(320, 180, dof=True, sky=True)
    ctxs = [b2pt.Context(i).upload(sc) for i in range(n)]
    one, st1 = ctxs[0].render(sc.camera, 10)
    many, stn = b2pt.group_render(ctxs, sc.camera, 10)
    assert np.allclose(many, one, rtol=2e-5, atol=2e-6)
    assert stn.rays_reference == st1.rays_reference and stn.bundles == st1.bundles
    # accumulating variant: a second half added onto the first
    acc, _ = b2pt.group_render(ctxs, sc.camera, 10, sample_begin=0, sample_count=3)
    acc, _ = b2pt.group_render(ctxs, sc.camera, 10, sample_begin=3, sample_count=7, out=acc)
    assert np.allclose(acc, one, rtol=2e-5, atol=2e-6)
    for c in ctxs:
        c.close()
    sc.close()


def test_program_with_two_gpus(tmp_path):
    exe = os.path.join(b2pt.PKG_DIR, "RayTracing")
    env = dict(os.environ, B2PT_ASSET_DIR=b2pt.ASSET_DIR)
    outs = []
    for g in (1, 2):
        build = tmp_path / f"build{g}"
        build.mkdir()
        r = subprocess.run([exe, "--demo", "--spp", "8", "--width", "96", "--height", "96", "--gpus", str(g)], cwd=str(build), env=env,
                           capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stderr
        outs.append(b2pt.read_png(str(build / "output.png")))
    assert (np.abs(outs[0].astype(int) - outs[1].astype(int)) <= 1).mean() > 0.999
