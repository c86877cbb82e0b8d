"""Host-side scene assembly (libb2pt_host.so: host/scene.cpp, obj_load.cpp, png_io.cpp) against the reference's own
MeshTriangle / Camera / Material code (oracle/_ref), plus the conf.json behaviours of src/main.cpp:137-316.  CPU-only."""
import ctypes as C
import os

import numpy as np
import pytest

import scenes
import support as S

b2pt = S.b2pt
fp = S.fp
REF_MODELS = os.path.join(S.REFERENCE_DIR, "models")
need_ref = pytest.mark.skipif(not S.have_ref(), reason="oracle/_ref/libref_oracle.so not built")
need_models = pytest.mark.skipif(not os.path.isdir(REF_MODELS), reason="/root/reference/models not present (GPU box)")


@need_ref
@need_models
@pytest.mark.parametrize("obj,tr,zoom", [("low_king.obj", (278.0, 0.0, 150.5), 1.0), ("low_soldier.obj", (-559.0, 0.0, -912.0), 1.0),
                                         ("diamond.obj", (0, 0, 0), 1.0), ("bottom.obj", (0, 0, 0), 1.0), ("light.obj", (278, 1300, 0), 1.0),
                                         ("cornellbox/tallbox.obj", (0.5, 0.25, -3), 1.7)])
def test_mesh_matches_reference(obj, tr, zoom):
    """MeshTriangle ctor (src/Triangle.hpp:83-135): vertices, normals, areas, bounds bit-identical; OBJ parse == .b2m pack."""
    path = os.path.join(REF_MODELS, obj)
    sc = b2pt.HostScene.empty()
    mat = sc.find_material("rough_plastic")
    k = sc.add_mesh(path, mat, tr, zoom)
    info = sc.object_info(k)
    L = S.ref_lib()
    h = C.c_void_p(L.ref_scene_new())
    zero, one = S.f32([0, 0, 0]), S.f32([1, 1, 1])
    L.ref_add_material(h, 3, fp(zero), 1.5, 0.01, 0.4, fp(zero), 0)
    t = S.f32(tr)
    L.ref_add_mesh(h, path.encode(), 0, fp(t), zoom)
    n = L.ref_mesh_triangle_count(h, 0)
    assert n == len(info["v9"])
    v9, n3, area = np.zeros((n, 9), np.float32), np.zeros((n, 3), np.float32), np.zeros(n, np.float32)
    L.ref_mesh_triangles(h, 0, fp(v9), fp(n3), fp(area))
    assert np.array_equal(v9.view(np.uint32), info["v9"].view(np.uint32))
    # normals / areas / boxes come out of the flattened scene
    sc.build_tree()
    d = sc.desc
    nrm = np.ctypeslib.as_array(d.prim_normal, shape=(d.n_prims, 4))
    po, pf = sc.prim_origins()
    assert np.array_equal(nrm[:, :3][np.argsort(pf)].view(np.uint32), n3.view(np.uint32))
    assert np.array_equal(nrm[:, 3][np.argsort(pf)].view(np.uint32), area.view(np.uint32))
    b6 = np.zeros(6, np.float32)
    L.ref_object_bounds(h, 0, fp(b6))
    root = d.nodes[0]
    assert np.array_equal(np.array(list(root.bmin) + list(root.bmax), np.float32).view(np.uint32), b6.view(np.uint32))
    assert np.float32(L.ref_object_area(h, 0)) == _mesh_area(area)
    # the packed asset holds the same stream
    pack = os.path.join(b2pt.ASSET_DIR, obj.replace("/", "_")[:-4] + ".b2m")
    sc2 = b2pt.HostScene.empty()
    k2 = sc2.add_mesh(pack, mat, tr, zoom)
    assert np.array_equal(sc2.object_info(k2)["v9"].view(np.uint32), info["v9"].view(np.uint32))
    assert np.array_equal(sc2.object_info(k2)["uv6"].view(np.uint32), info["uv6"].view(np.uint32))
    L.ref_scene_free(h)
    sc.close(); sc2.close()


def _mesh_area(area):
    s = np.float32(0)
    for a in area:  # `area += triangle area` in file order, float accumulation (src/Triangle.hpp:126)
        s = np.float32(s + a)
    return s


@need_ref
def test_camera_matches_reference():
    for (w, h, fov, pos, tgt) in [(1920, 1080, 70.0, (278, 150, -2550), (278, 0, 0)), (384, 384, 40.0, (278, 273, -800), (278, 273, 0)),
                                  (123, 77, 33.3, (1, 2, 3), (-4, 5, 60))]:
        sc, _ = scenes.two_triangle_scene()
        sc.set_camera(w, h, fov, pos, tgt, (0, 1, 0), True, 900.0, 4.5)
        ref = S.Ref(sc)
        m9 = np.zeros(9, np.float32)
        ref.L.ref_camera_orientation(ref.h, fp(m9))
        cam = sc.camera
        assert np.array_equal(np.array(list(cam.orientation), np.float32).view(np.uint32), m9.view(np.uint32))
        assert np.float32(ref.L.ref_camera_scale(ref.h)) == np.float32(cam.scale)
        assert np.float32(cam.aspect) == np.float32(w / np.float32(h))
        ref.close(); sc.close()


@need_ref
def test_material_defaults_match_reference_ctor():
    L = S.ref_lib()
    sc = b2pt.HostScene.empty()
    for t in range(4):
        a, b, r = C.c_float(), C.c_float(), C.c_float()
        d = C.c_int()
        L.ref_material_defaults(t, C.byref(a), C.byref(b), C.byref(r), C.byref(d))
        # the named materials override what they set; defaults show through elsewhere
        assert d.value == (1 if t in (0, 2) else 0)
    names = [n for n, _ in sc.materials()]
    assert names == b2pt.NAMED_MATERIALS
    mats = dict(sc.materials())
    assert mats["gold_conductor"].type == b2pt.SMOOTH_CONDUCTOR and np.allclose(list(mats["gold_conductor"].base_reflectance), [1.0, 0.85, 0.57])
    assert mats["smooth_glass_gem"].ior_a == np.float32(1.3) and mats["smooth_glass_gem"].ior_b == np.float32(0.2)
    assert mats["green_mirror"].type == b2pt.ROUGH_CONDUCTOR and mats["green_mirror"].roughness == np.float32(0.01)
    assert mats["clear_rough_plastic"].roughness == np.float32(0.02) and mats["rough_plastic"].roughness == np.float32(0.4)
    sc.close()


def test_conf_scene_and_quirks():
    """src/main.cpp:137-316 with the shipped conf.json values, quirks kept by default."""
    sc, env = scenes.chess(320, 180, dof=True, sky=True)
    assert sc.n_objects == 14 + 4                      # soldiers (Add order: as constructed), light, floor, king, diamond
    kinds = [sc.object_info(k) for k in range(sc.n_objects)]
    assert [len(o["v9"]) for o in kinds[14:]] == [2, 2, len(kinds[16]["v9"]), len(kinds[17]["v9"])]
    names = [n for n, _ in sc.materials()]
    assert names[kinds[0]["material"]] == "smooth_glass" and names[kinds[1]["material"]] == "rough_white_conductor"
    assert names[kinds[14]["material"]] == "light" and names[kinds[15]["material"]] == "silver_mirror" and names[kinds[16]["material"]] == "gold_conductor"
    d = sc.desc
    assert d.n_dir_sample == 4                         # directLightSample: 32 is never read (src/Scene.hpp:114-116)
    assert d.rr_rate == np.float32(0.4) and d.inv_rr == np.float32(1) / np.float32(0.4)
    assert d.use_env_map == 1 and d.n_lights == 1
    assert sc.camera.use_dof == 1 and sc.camera.focal_distance == np.float32(3036.98) and sc.camera.aperture_radius == 10
    assert sc.spp == 32
    assert dict(sc.materials())["silver_mirror"].textured == 1   # floor_isTextured mutates the shared material
    light = dict(sc.materials())["light"]
    assert abs(light.emission[0] - 100 * 47.8348) < 0.1  # lightBrightness 100.0 (a JSON float)
    assert d.n_prims == sum(len(o["v9"]) for o in kinds) and len(kinds[0]["v9"]) == 2560
    n_low = d.n_prims
    # model_quality has no effect unless the fix is requested; directLightSample likewise
    hi, _ = scenes.chess(64, 36, quality="high")
    assert hi.desc.n_prims == n_low
    hi.close()
    hi, _ = scenes.chess(64, 36, quality="high", fix=b2pt.FIX_MODEL_QUALITY | b2pt.FIX_DIRECT_LIGHT_SAMPLE)
    assert hi.desc.n_prims > 250000 and hi.desc.n_dir_sample == 32
    hi.close()
    dark, _ = scenes.chess(64, 36, dof=False, sky=False)
    assert dark.desc.use_env_map == 0 and list(dark.desc.background) == [0, 0, 0] and dark.camera.use_dof == 0
    dark.close(); sc.close()


def test_conf_errors_are_soft(tmp_path):
    """A malformed conf.json prints and carries on with defaults (src/main.cpp:291-294); a missing model fails loudly."""
    run = tmp_path / "build"
    run.mkdir()
    (run / "conf.json").write_text("{ this is not json")
    sc = b2pt.HostScene.from_conf(str(run / "conf.json"), str(run))
    assert sc.n_objects == 3 and sc.camera.width == 384     # light, floor, king with defaults; no soldiers, no diamond
    sc.close()
    L = b2pt.host_lib()
    L.b2pt_host_set_asset_dir(b"/nonexistent")
    try:
        with pytest.raises(RuntimeError):
            b2pt.HostScene.from_conf(str(run / "conf.json"), str(run))
    finally:
        L.b2pt_host_set_asset_dir(b2pt.ASSET_DIR.encode())


def test_demo_scene():
    sc, _ = scenes.cornell(0, 0)
    assert (sc.camera.width, sc.camera.height) == (384, 384) and sc.n_objects == 9 and sc.desc.n_prims == 35
    assert sc.desc.rr_rate == np.float32(0.7) and sc.desc.n_dir_sample == 4 and sc.spp == 2048
    assert sc.desc.n_lights == 1 and abs(sc.desc.light_area[0] - 13650.0) < 1e-3
    sc.close()


def test_tonemap_and_png(tmp_path):
    rgb = np.array([[0.0, 0.5, 1.0], [2.0, np.nan, -1.0], [1e-6, 0.25, np.inf]], np.float32)
    out = b2pt.tonemap_rgba8(rgb)
    want = np.zeros((3, 4), np.uint8)
    for i in range(3):
        for c in range(3):
            v = np.float32(255 * np.power(np.float64(rgb[i, c]), np.float64(np.float32(0.45)))) if not np.isnan(rgb[i, c]) and rgb[i, c] >= 0 else np.float32(np.nan)
            m = v if v < 255 else np.float32(255)   # std::min(hi, v): NaN -> hi
            r = m if 0 < m else np.float32(0)
            want[i, c] = np.uint8(r)
        want[i, 3] = 255
    assert np.array_equal(out, want)
    assert out[1, 1] == 255 and out[1, 2] == 255   # NaN -> 255 (pow of a negative is NaN too), src/global.hpp:16-18
    img = (np.random.RandomState(0).rand(7, 5, 4) * 255).astype(np.uint8)
    p = str(tmp_path / "x.png")
    b2pt.write_png(p, img, 5, 7)
    assert np.array_equal(b2pt.read_png(p), img)


@need_models
def test_demo_png_channel_means():
    """The one end-to-end fixture the reference ships: cornellbox_demo.png.  The oracle render (the reference's own castRay)
    reproduces its linear channel means; a cheap low-spp check that the shim-built reference is the published program."""
    png = os.path.join(S.REFERENCE_DIR, "cornellbox_demo.png")
    if not (os.path.exists(png) and S.have_ref()):
        pytest.skip("fixture or oracle not present")
    img = b2pt.read_png(png).astype(np.float64)[..., :3] / 255.0
    lin = np.power(img, 1 / 0.45).mean(axis=(0, 1))
    sc, _ = scenes.cornell(96, 96)
    ref = S.Ref(sc)
    fb = np.clip(ref.render_frame(0, 6, 6), 0, 1).mean(axis=(0, 1))
    assert np.allclose(fb, lin, rtol=0.06), (fb, lin)
    ref.close(); sc.close()


def test_upload_validation_rejects_what_the_kernels_cannot_index():
    """b2pt_upload_scene's validator (csrc/pt_pack.hpp, shared with tests/hostcheck): material types outside the four of
    src/Material.hpp:13-18, leaves whose kind differs from prim_kind, light-sample counts beyond the slot arithmetic and trees
    deeper than the walk stacks are refused; the depth is computed from the links, not taken from the caller."""
    sc, _ = scenes.cornell(32, 32)
    d = sc.desc
    L = S.hc_lib()
    L.hc_scene_new.restype = C.c_void_p

    def accepted():
        h = L.hc_scene_new(C.byref(d))
        if h:
            L.hc_scene_free(C.c_void_p(h))
        return bool(h)

    assert accepted()
    for bad in (-1, 4, 7, 1 << 20):
        keep = d.materials[2].type
        d.materials[2].type = bad
        assert not accepted()
        d.materials[2].type = keep
    keep = d.prim_kind[0]
    d.prim_kind[0] = 2 if keep == 1 else 1
    assert not accepted()
    d.prim_kind[0] = 9
    assert not accepted()
    d.prim_kind[0] = keep
    for bad in (0, -3, 1025):
        keep = d.n_dir_sample
        d.n_dir_sample = bad
        assert not accepted()
        d.n_dir_sample = keep
    # a caller-supplied depth is not trusted: lying about it changes nothing ...
    keep = d.max_depth
    d.max_depth = 1000
    assert accepted()
    d.max_depth = keep
    # ... and links that loop are caught by the walk
    root_a = d.nodes[0].a
    idx = next(i for i in range(2, d.n_nodes) if d.nodes[i].kind == 0)
    keep = d.nodes[idx].a
    d.nodes[idx].a = root_a
    assert not accepted()
    d.nodes[idx].a = keep
    assert accepted()
    # the host assembler refuses such materials too
    with pytest.raises(RuntimeError):
        sc.add_material("bad", b2pt.Material(5, (0, 0, 0), 1.5, 0.0, 0.1, (0, 0, 0), 0, 0))
    sc.close()


def _png(width, height, depth, ctype, rows_of, interlace=0, plte=None):
    """A PNG file image as bytes: rows_of(pass_width, xs, ys, dx, dy, y) -> packed row bytes (filter 0)."""
    import struct
    import zlib

    def chunk(t, d):
        return struct.pack(">I", len(d)) + t + d + struct.pack(">I", zlib.crc32(t + d) & 0xFFFFFFFF)

    passes = [(0, 0, 1, 1)] if not interlace else [(0, 0, 8, 8), (4, 0, 8, 8), (0, 4, 4, 8), (2, 0, 4, 4), (0, 2, 2, 4), (1, 0, 2, 2), (0, 1, 1, 2)]
    raw = b""
    for xs, ys, dx, dy in passes:
        pw = (width - xs + dx - 1) // dx if width > xs else 0
        ph = (height - ys + dy - 1) // dy if height > ys else 0
        if pw == 0 or ph == 0:
            continue
        for y in range(ph):
            raw += b"\x00" + rows_of(pw, xs, ys, dx, dy, y)
    out = b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", width, height, depth, ctype, 0, 0, interlace))
    if plte is not None:
        out += chunk(b"PLTE", plte)
    return out + chunk(b"IDAT", zlib.compress(raw)) + chunk(b"IEND", b"")


def test_png_reader_formats_and_hostile_headers(tmp_path):
    """The env-map decoder (host/png_io.cpp) stands in for lodepng::decode (src/Scene.hpp:39-57): RGBA8 out of any colour type and
    bit depth, Adam7-interlaced files included; malformed headers are errors, never crashes or exceptions across the C ABI."""
    rng = np.random.RandomState(3)
    W, H = 13, 11  # not multiples of 8: the last Adam7 passes are partial
    img = rng.randint(0, 256, (H, W, 4)).astype(np.uint8)

    def rgba8(pw, xs, ys, dx, dy, y):
        return img[ys + y * dy, xs:xs + pw * dx:dx].tobytes()

    def rgb16(pw, xs, ys, dx, dy, y):
        px = img[ys + y * dy, xs:xs + pw * dx:dx, :3].astype(np.uint16)
        return ((px << 8) | 0x5A).astype(">u2").tobytes()  # lodepng keeps the high byte

    pal = rng.randint(0, 256, (16, 3)).astype(np.uint8)
    idx = rng.randint(0, 16, (H, W)).astype(np.uint8)

    def pal4(pw, xs, ys, dx, dy, y):
        v = idx[ys + y * dy, xs:xs + pw * dx:dx]
        v = np.concatenate([v, np.zeros(len(v) % 2, np.uint8)])
        return ((v[0::2] << 4) | v[1::2]).astype(np.uint8).tobytes()

    cases = {"rgba8": (8, 6, rgba8, None, img), "rgb16": (16, 2, rgb16, None, np.concatenate([img[..., :3], np.full((H, W, 1), 255, np.uint8)], -1)),
             "pal4": (4, 3, pal4, pal.tobytes(), np.concatenate([pal[idx], np.full((H, W, 1), 255, np.uint8)], -1))}
    for name, (depth, ctype, rows, plte, want) in cases.items():
        for interlace in (0, 1):
            p = tmp_path / f"{name}_{interlace}.png"
            p.write_bytes(_png(W, H, depth, ctype, rows, interlace, plte))
            got = b2pt.read_png(str(p))
            assert got.shape == (H, W, 4)
            assert np.array_equal(got, want), (name, interlace)
    # hostile headers
    for depth, ctype, w, h in ((0, 6, 4, 4), (3, 2, 4, 4), (8, 5, 4, 4), (8, 6, 0, 4), (8, 6, 0x7FFFFFFF, 0x7FFFFFFF), (16, 3, 4, 4)):
        p = tmp_path / "bad.png"
        p.write_bytes(_png(min(w, 4), min(h, 4), 8, 6, lambda pw, *a: bytes(4 * pw)).replace(
            __import__("struct").pack(">IIBB", min(w, 4), min(h, 4), 8, 6), __import__("struct").pack(">IIBB", w, h, depth, ctype)))
        with pytest.raises(Exception):
            b2pt.read_png(str(p))
    # an env map the reader refuses leaves the scene on its background colour, as Scene::loadEnvMap does
    sc = b2pt.HostScene.empty()
    assert sc.load_env_png(str(tmp_path / "bad.png")) != 0
    sc.close()
