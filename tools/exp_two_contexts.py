"""Experiment: does running TWO render contexts at once on one GPU (each half of the samples, own stream, own host thread)
hide the ramp-down tails of the kernels?   python tools/exp_two_contexts.py [spp]"""
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402
import bench  # noqa: E402
import support as S  # noqa: E402

b2pt = S.b2pt


def main():
    spp = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    args = bench.parse.__wrapped__() if hasattr(bench.parse, "__wrapped__") else None
    sys.argv = sys.argv[:1]
    args = bench.parse()
    sc, _ = bench.make_scene(args)
    cam = sc.camera
    dev = torch.device("cuda", 0)
    for n_ctx, queue in ((1, 0), (2, 24 << 20), (2, 16 << 20), (3, 16 << 20), (1, 0)):
        ctxs = [b2pt.Context(0) for _ in range(n_ctx)]
        for c in ctxs:
            c.upload(sc)
        fbs = [torch.zeros((cam.height, cam.width, 3), dtype=torch.float32, device=dev) for _ in range(n_ctx)]
        per = spp // n_ctx

        def work(i, seed):
            ctxs[i].render_device(cam, spp, fbs[i].data_ptr(), sample_begin=i * per, sample_count=per, seed=seed, max_wave_bundles=queue)

        def step(seed):
            ts = [threading.Thread(target=work, args=(i, seed)) for i in range(n_ctx)]
            for t in ts:
                t.start()
            for t in ts:
                t.join()
            torch.cuda.synchronize()

        for w in range(2):
            step(w)
        t0 = time.perf_counter()
        reps = 4
        for k in range(reps):
            step(10 + k)
        dt = (time.perf_counter() - t0) / reps
        print(f"{n_ctx} context(s), queue {queue >> 20} Mi: {dt * 1e3:.2f} ms per {spp}-spp frame", flush=True)
        for c in ctxs:
            c.close()
        del fbs
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
