import sys, time
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import numpy as np
import support as S, scenes
b2pt = S.b2pt
sc, env = scenes.chess(1920, 1080, dof=True, sky=True)
for q in (16 << 20, 48 << 20, 0):
    t0 = time.perf_counter(); ctx = b2pt.Context(0); ctx.upload(sc); t1 = time.perf_counter()
    out = []
    for rep in range(3):
        t2 = time.perf_counter()
        fb, st = ctx.render(sc.camera, 32, max_wave_bundles=q, flags=b2pt.FLAG_FRESH_FRAME)
        out.append((round((time.perf_counter() - t2) * 1e3, 1), round(st.gpu_ms, 1), st.kernel_launches))
    t3 = time.perf_counter(); ctx.close(); t4 = time.perf_counter()
    print(f"queue {q >> 20} Mi: create+upload {1e3 * (t1 - t0):.0f} ms, renders (wall ms, gpu ms, launches) {out}, close {1e3 * (t4 - t3):.0f} ms", flush=True)
