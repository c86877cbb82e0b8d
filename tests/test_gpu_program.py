"""The drop-in program: ./RayTracing run the way the reference is run (cwd with conf.json beside ../models) produces the same
PNG as the library API followed by the reference's tone map, prints the reference's stdout lines, and fails loudly on errors."""
import os
import shutil
import subprocess

import numpy as np
import pytest

import scenes
import support as S

b2pt = S.b2pt
pytestmark = pytest.mark.gpu
EXE = os.path.join(b2pt.PKG_DIR, "RayTracing")


def run(cmd, cwd):
    env = dict(os.environ, B2PT_ASSET_DIR=b2pt.ASSET_DIR)
    return subprocess.run(cmd, cwd=cwd, env=env, capture_output=True, text=True, timeout=600)


def test_conf_json_to_output_png(tmp_path):
    build = tmp_path / "build"
    build.mkdir()
    (tmp_path / "models" / "envoMaps").mkdir(parents=True)
    S.write_sky_png(str(tmp_path / "models" / "envoMaps" / "sky.png"))
    (build / "conf.json").write_text(S.chess_conf_text(160, 90, 6, True))
    r = run([EXE], str(build))
    assert r.returncode == 0, r.stderr
    out = r.stdout
    for line in (" - Generating BVH...", "SPP: 6", "Writing image to ./output.png", "Rendering finished in "):
        assert line in out, out
    assert "] 100 %" in out
    img = b2pt.read_png(str(build / "output.png"))
    assert img.shape == (90, 160, 4) and (img[..., 3] == 255).all()
    # the same frame through the library API + the reference's tone map (Renderer.cpp:93-102)
    sc = b2pt.HostScene.from_conf(str(build / "conf.json"), str(build)).build_tree()
    ctx = b2pt.Context(0).upload(sc)
    fb, _ = ctx.render(sc.camera, 6)
    want = b2pt.tonemap_rgba8(fb).reshape(90, 160, 4)
    assert (np.abs(img.astype(int) - want.astype(int)) <= 1).mean() > 0.999  # atomics order: last-bit differences at most
    assert img[..., :3].mean() > 20
    ctx.close(); sc.close()


def test_demo_and_overrides(tmp_path):
    build = tmp_path / "build"
    build.mkdir()
    r = run([EXE, "--demo", "--spp", "4", "--width", "64", "--height", "48", "--chunk", "2"], str(build))
    assert r.returncode == 0, r.stderr
    assert "SPP: 4" in r.stdout
    img = b2pt.read_png(str(build / "output.png"))
    assert img.shape == (48, 64, 4) and img[..., :3].mean() > 10
    if S.have_ref():
        sc, _ = scenes.cornell(64, 48)
        ref = S.Ref(sc)
        fb = ref.render_frame(0, 4, 4)
        want = b2pt.tonemap_rgba8(fb).reshape(48, 64, 4)
        assert (np.abs(img.astype(int) - want.astype(int)) <= 1).mean() > 0.995  # same sample streams as the oracle
        ref.close(); sc.close()


def test_checkpoint_resume_and_preview(tmp_path):
    """A run stopped after half of its samples and resumed from its checkpoint ends with the image of the uninterrupted run
    (same sample streams: the frame does not depend on how the samples are split over calls or processes)."""
    build = tmp_path / "build"
    build.mkdir()
    common = [EXE, "--demo", "--spp", "8", "--width", "64", "--height", "48", "--chunk", "2"]
    r = run(common, str(build))
    assert r.returncode == 0, r.stderr
    whole = b2pt.read_png(str(build / "output.png")).astype(int)
    ck = str(build / "half.ckpt")
    r = run(common + ["--checkpoint", ck, "--stop-after", "4", "--preview-every", "1"], str(build))
    assert r.returncode == 0, r.stderr
    half = b2pt.read_png(str(build / "output.png")).astype(int)
    assert os.path.getsize(ck) == 40 + 64 * 48 * 3 * 4  # header (magic, size, spp, done, seed, scene hash) + fp32 frame
    # the 4-of-8-spp preview is scaled by spp / samples done: about as bright as the whole frame, only noisier
    assert abs(half[..., :3].mean() - whole[..., :3].mean()) < 0.08 * whole[..., :3].mean()
    r = run(common + ["--resume", ck], str(build))
    assert r.returncode == 0, r.stderr
    resumed = b2pt.read_png(str(build / "output.png")).astype(int)
    assert (np.abs(resumed - whole) <= 1).mean() > 0.999  # atomics order: last-bit differences at most
    # a checkpoint written for another spp is refused
    r = run([EXE, "--demo", "--spp", "16", "--width", "64", "--height", "48", "--resume", ck], str(build))
    assert r.returncode != 0 and "cannot resume" in r.stderr
    # ... and so is one written for another scene configuration (here: another light-sample count), instead of mixing the two
    r = run(common + ["--ndir", "7", "--resume", ck], str(build))
    assert r.returncode != 0 and "cannot resume" in r.stderr


def test_bad_arguments_fail_loudly(tmp_path):
    build = tmp_path / "build"
    build.mkdir()
    assert run([EXE, "--nonsense"], str(build)).returncode == 2
    r = run([EXE, "--demo", "--device", "99", "--spp", "1"], str(build))
    assert r.returncode != 0 and "b2pt_create" in r.stderr
