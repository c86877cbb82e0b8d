"""Where the wall time of one ./RayTracing-like job goes, phase by phase (f1 of SURVEY 8f): host scene assembly, context,
scene upload (SAH tree, quads, tables), first render (includes the wave allocation), steady render, teardown.
    python tools/time_phases.py [spp] [ndir] [quality]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
t00 = time.perf_counter()
import b2pt_loader
b2pt = b2pt_loader.load()
from b2pt import scenes
import numpy as np
spp = int(sys.argv[1]) if len(sys.argv) > 1 else 32
ndir = int(sys.argv[2]) if len(sys.argv) > 2 else 4
quality = sys.argv[3] if len(sys.argv) > 3 else "low"
def lap(name, t0):
    t = time.perf_counter(); print(f"{name:34s} {1e3 * (t - t0):9.1f} ms", flush=True); return t
t = lap("python imports", t00)
sc, env = scenes.chess(1920, 1080, dof=True, sky=True, quality=quality, n_dir=ndir, fix=b2pt.FIX_MODEL_QUALITY if quality == "high" else 0, sky_size=(2048, 1024))
t = lap("scene assembly (host, incl. sky png)", t)
ctx = b2pt.Context(0)
t = lap("b2pt_create (CUDA context)", t)
ctx.upload(sc)
t = lap("b2pt_upload_scene (trees + copies)", t)
fb, st = ctx.render(sc.camera, spp)
t = lap(f"first b2pt_render ({spp} spp)", t)
print(f"    of which device time {st.gpu_ms:.1f} ms, launches {st.kernel_launches}")
fb, st = ctx.render(sc.camera, spp)
t = lap(f"second b2pt_render ({spp} spp)", t)
print(f"    of which device time {st.gpu_ms:.1f} ms")
rgba = ctx.tonemap_rgba8(None, 1920 * 1080)
t = lap("device tone map", t)
b2pt.write_png("/tmp/out.png", rgba, 1920, 1080)
t = lap("PNG encode + write", t)
ctx.close()
t = lap("b2pt_destroy", t)
