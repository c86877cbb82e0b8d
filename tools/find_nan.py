"""Finds pixel-samples of the chess frame whose radiance is NaN / inf on the GPU and replays them through the reference's own castRay
(oracle/_ref) on the same sample streams: the reference poisons those pixels too (framebuffer += NaN, Renderer.cpp:80), so they
must be the SAME samples with the same non-finite class.   python tools/find_nan.py [spp] [ndir]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import b2pt_loader
b2pt = b2pt_loader.load()
from b2pt import scenes
from oracle import refbind as R
spp = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
ndir = int(sys.argv[2]) if len(sys.argv) > 2 else 32
sc, env = scenes.chess(1920, 1080, dof=True, sky=True, n_dir=ndir, sky_size=(2048, 1024))
ctx = b2pt.Context(0).upload(sc)
fb, st = ctx.render(sc.camera, spp)
bad = np.argwhere(~np.isfinite(fb).all(axis=2))
print(f"{spp} spp, nee{ndir}: {len(bad)} pixels with a non-finite channel of {fb.shape[0] * fb.shape[1]}")
if len(bad):
    px = (bad[:, 0] * sc.camera.width + bad[:, 1]).astype(np.int32)[:64]
    got, _ = ctx.render_samples(sc.camera, px, 0, spp)
    ref = R.Ref(sc, env)
    want = ref.render_samples(px, 0, spp)
    nf_g, nf_r = ~np.isfinite(got), ~np.isfinite(want)
    print(f"  non-finite per-sample values: GPU {int(nf_g.sum())}, reference replay {int(nf_r.sum())}, same positions: {bool((nf_g == nf_r).all())}")
    same_class = bool((np.isnan(got) == np.isnan(want)).all() and (np.isposinf(got) == np.isposinf(want)).all() and (np.isneginf(got) == np.isneginf(want)).all())
    print(f"  same class (NaN / +inf / -inf): {same_class}")
    fin = ~(nf_g | nf_r)
    print(f"  the finite samples of those pixels: max |GPU - reference| = {np.abs(got[fin] - want[fin]).max():.3g}")
    i = np.argwhere(nf_g)[:5]
    for q, k, c in i:
        print(f"    pixel {int(px[q])} sample {int(k)} channel {int(c)}: GPU {got[q, k, c]} reference {want[q, k, c]}")
