"""Times the Cornell box (BASELINE configs[0] / [4] geometry) at 1024x1024: python tools/exp_cornell.py [spp]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import support as S, scenes
b2pt = S.b2pt
spp = int(sys.argv[1]) if len(sys.argv) > 1 else 256
sc, _ = scenes.cornell(1024, 1024)
ctx = b2pt.Context(0); ctx.upload(sc)
for rep in range(3):
    fb, st = ctx.render(sc.camera, spp, flags=b2pt.FLAG_FRESH_FRAME)
    print(f"{os.environ.get('B2PT_GPU_LIB', 'default')}: {st.gpu_ms:.1f} ms, {st.rays_reference / st.gpu_ms / 1e3:.0f} Mrays/s (reference-definition), "
          f"{(st.rays_traced_closest + st.rays_traced_shadow) / st.gpu_ms / 1e3:.0f} M traced rays/s, mean {fb.mean():.5f}", flush=True)
ctx.close()
