"""BASELINE configs[4]: each of the nine conf.json materials on the Cornell spheres and boxes, 1024x1024, spp 4096, one GPU.
    python tools/exp_material_sweep.py [spp]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import support as S
b2pt = S.b2pt
spp = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
W = H = 1024
ctx = b2pt.Context(0)
demo = b2pt.HostScene.demo(W, H)
objs = [demo.object_info(k) for k in range(demo.n_objects)]
demo.close()
print(f"Cornell {W}x{H}, {spp} spp, material i on the three spheres and the two boxes (walls unchanged)")
for mi, name in enumerate(b2pt.NAMED_MATERIALS):
    sc = b2pt.HostScene.empty()
    light = sc.add_material("light", b2pt.Material(b2pt.ROUGH_CONDUCTOR, tuple(3.9 * x for x in (47.8348, 38.5664, 31.0808)), 1.74, 0.1, 1.0, (0, 0, 0), 0, 0))
    for k, o in enumerate(objs):
        if o["kind"] == "sphere":
            sc.add_sphere(o["center"], o["radius"], mi)
        elif k in (1, 2):
            sc.add_triangles(o["v9"], mi)
        elif k == 5:
            sc.add_triangles(o["v9"], light)
        else:
            sc.add_triangles(o["v9"], o["material"])
    sc.set_camera(W, H, 40.0, (278, 273, -800), (278, 273, 0))
    sc.build_tree()
    ctx.upload(sc)
    ctx.render(sc.camera, 64, flags=b2pt.FLAG_FRESH_FRAME)  # warm-up
    ms, rays, traced = 0.0, 0, 0
    fb = None
    for s0 in range(0, spp, 512):
        n = min(512, spp - s0)
        fb, st = ctx.render(sc.camera, spp, sample_begin=s0, sample_count=n, out=fb, flags=b2pt.FLAG_FRESH_FRAME if s0 == 0 else 0)
        ms += st.gpu_ms; rays += st.rays_reference; traced += st.rays_traced_closest + st.rays_traced_shadow
    print(f"{name:24s} {ms / 1e3:6.2f} s  {rays / ms / 1e3:8.0f} Mrays/s (reference-definition)  {traced / ms / 1e3:6.0f} M traced rays/s  "
          f"{W * H * spp / ms / 1e3:6.0f} M pixel-samples/s  frame mean {fb.mean():.4f}", flush=True)
    sc.close()
ctx.close()
