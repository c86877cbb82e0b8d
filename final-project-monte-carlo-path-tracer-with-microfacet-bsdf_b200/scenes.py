"""The reference's scenes as the BASELINE configurations name them, assembled through the host library.

* :func:`cornell` — the DEMO scene of src/main.cpp:99-129 (configs[0]; `make DEMO=1`), with the size override BASELINE asks.
* :func:`chess` — the conf.json scene of src/main.cpp:137-316 with the shipped values (configs[1]-[3]): a conf.json is
  written into a scratch run directory and read back by the same code path `./RayTracing` uses.
* :func:`cornell_sweep` — configs[4]: material *i* of the nine on the three spheres and the two boxes of the Cornell box.
* :func:`synthetic_sky` — the deterministic equirect sky that stands in for `models/envoMaps/sky.png` (a missing blob upstream).

Used by bench.py, tools/ and the tests; nothing here touches the GPU.
"""
from __future__ import annotations

import os
import tempfile

import numpy as np

import b2pt

_keep = []  # scratch directories that must outlive the scenes built from them

CHESS_CONF = """{
  "camera": {"width": %(w)d, "height": %(h)d, "fov": 70,
             "position": [278, 150, -2550], "target": [278, 0, 0], "up": [0, 1, 0],
             "useDOF": %(dof)s, "focusDistance": 3036.98, "apertureRadius": 10},
  "renderer": {"spp": %(spp)d, "path": "./output.png", "parrallelism": 8},
  "scene": {"addDiamond": true, "model_quality": "%(quality)s", "includeShadow": true,
            "RussianRouletteRate": 0.4, "directLightSample": 32, "envMap": %(env)s,
            "kingPosition": [0, 0, 0], "kingMaterial": "%(king)s",
            "soldierLeftRowPosition": [-559, 0, -200], "soldierRightRowPosition": [160, 0, -200],
            "soldierXSpacing": 0, "soldierYSpacing": 0, "soldierZSpacing": -356, "soldierCountPerRow": 7,
            "soldierMaterials": [%(soldiers)s],
            "lightPosition": [278, 1300, 0], "lightBrightness": 100.0,
            "floorMaterial": "silver_mirror", "floor_isTextured": true, "wallMaterial": "rough_white_conductor"}
}
"""


def chess_conf_text(w=1920, h=1080, spp=2048, dof=True, env='"../models/envoMaps/sky.png"', quality="low", king="gold_conductor",
                    left="smooth_glass", right="rough_white_conductor"):
    """The shipped conf.json (/root/reference/conf.json) with the fields the BASELINE configurations vary."""
    soldiers = ", ".join(['"%s"' % left] * 7 + ['"%s"' % right] * 7)
    return CHESS_CONF % dict(w=w, h=h, spp=spp, dof="true" if dof else "false", env=env, quality=quality, king=king, soldiers=soldiers)


def synthetic_sky(width=256, height=128, seed=0):
    """Deterministic RGBA8 equirect sky (blue-to-white gradient + value-noise clouds); sky.png is absent upstream."""
    rng = np.random.RandomState(seed)
    coarse = rng.rand(height // 8 + 2, width // 8 + 2)
    ys, xs = np.mgrid[0:height, 0:width]
    gy, gx = ys / 8.0, xs / 8.0
    y0, x0 = gy.astype(int), gx.astype(int)
    fy, fx = gy - y0, gx - x0
    n = (coarse[y0, x0] * (1 - fx) + coarse[y0, x0 + 1] * fx) * (1 - fy) + (coarse[y0 + 1, x0] * (1 - fx) + coarse[y0 + 1, x0 + 1] * fx) * fy
    v = ys / max(height - 1, 1)
    cloud = np.clip((n - 0.55) * 3.0, 0, 1) * (v < 0.5)
    r = 0.35 + 0.55 * v + 0.4 * cloud
    g = 0.55 + 0.40 * v + 0.3 * cloud
    b = 0.95 * np.ones_like(v)
    img = np.stack([r, g, b, np.ones_like(v)], -1)
    return (np.clip(img, 0, 1) * 255).astype(np.uint8)


def write_sky_png(path, width=256, height=128, seed=0):
    b2pt.write_png(path, synthetic_sky(width, height, seed), width, height)
    return path


def cornell(width=96, height=96, n_dir=0, rr=-1.0):
    """DEMO scene of src/main.cpp:99-129 (BASELINE configs[0] geometry).  Returns (scene, None)."""
    sc = b2pt.HostScene.demo(width, height)
    sc.set_render(0, rr, -1, n_dir)
    return sc.build_tree(), None


def chess(width=160, height=90, dof=True, sky=True, quality="low", n_dir=0, fix=0, king="gold_conductor", left="smooth_glass",
          right="rough_white_conductor", spp=32, sky_size=(256, 128)):
    """conf.json scene of src/main.cpp:137-316 with the shipped values (BASELINE configs[1]-[3]).  Returns (scene, path of the
    env-map PNG or None)."""
    tmp = tempfile.TemporaryDirectory(prefix="b2pt_chess_")
    _keep.append(tmp)
    run = os.path.join(tmp.name, "build")
    os.makedirs(run)
    os.makedirs(os.path.join(tmp.name, "models", "envoMaps"))
    env_png = None
    if sky:
        env_png = write_sky_png(os.path.join(tmp.name, "models", "envoMaps", "sky.png"), sky_size[0], sky_size[1])
        env = '"../models/envoMaps/sky.png"'
    else:
        env = "[0, 0, 0]"
    conf = os.path.join(run, "conf.json")
    with open(conf, "w") as f:
        f.write(chess_conf_text(width, height, spp, dof, env, quality, king, left, right))
    sc = b2pt.HostScene.from_conf(conf, run, fix)
    if n_dir:
        sc.set_render(0, -1.0, -1, n_dir)
    return sc.build_tree(), env_png


def cornell_sweep(material, width=1024, height=1024, n_dir=0):
    """BASELINE configs[4]: the DEMO Cornell box with `material` (a name of b2pt.NAMED_MATERIALS or an index) on the three spheres
    and the two boxes; walls, floor and light unchanged (same objects in the same Add order, hence the same primitive numbering)."""
    demo = b2pt.HostScene.demo(width, height)
    objs = [demo.object_info(k) for k in range(demo.n_objects)]
    mats = demo.materials()
    fov, pos, tgt, up = demo.camera_params()
    demo.close()
    sc = b2pt.HostScene.empty()
    n_named = len(sc.materials())
    remap = {}
    for idx, (name, m) in enumerate(mats):  # materials main() creates beyond the nine named ones (the light)
        remap[idx] = idx if idx < n_named else sc.add_material(name, m)
    mi = sc.find_material(material) if isinstance(material, str) else int(material)
    if mi < 0:
        raise ValueError(f"unknown material {material!r}")
    for k, o in enumerate(objs):
        if o["kind"] == "sphere":
            sc.add_sphere(o["center"], o["radius"], mi)
        elif k in (1, 2):  # shortbox, tallbox (src/main.cpp:108-109,118-119)
            sc.add_triangles(o["v9"], mi)
        else:
            sc.add_triangles(o["v9"], remap[o["material"]])
    sc.set_camera(width, height, fov, tuple(pos), tuple(tgt), tuple(up))
    if n_dir:
        sc.set_render(0, -1.0, -1, n_dir)
    return sc.build_tree(), None
