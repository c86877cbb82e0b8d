#!/usr/bin/env python
"""bench.py — the reference's headline benchmark on B200: the 1080p chess frame of BASELINE.json configs[1].

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--frame-spp S] [--ndir D] [--scene ...]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

Workload (configs[1] AS WRITTEN, the north-star experiment): the shipped conf.json chess scene, 1920x1080, thin-lens DoF, sky
env map (synthetic 2048x1024: the upstream sky.png is absent), Russian roulette 0.4, **32 next-event samples per vertex**
(conf.json:23 through Scene::setDirectLightSample; the reference's own main() never reads the key and runs 4 — that
configuration is timed too and reported under `variants.nee4`).

One step = ONE FRAME of --frame-spp (2048) samples per pixel: one pass of the hot path (Renderer::Render's pixel loop) over the
whole job.  With N GPUs the samples of the frame are split across the ranks (STRONG scaling: the frame is fixed, rank g renders
samples [g S/N, (g+1) S/N) of every pixel) and the fp32 radiance buffers are summed on rank 0 with one NCCL reduce per
frame.  `ms_per_step` therefore IS the time of the 2048-spp frame the north star bounds by 10 s on 8 GPUs.

`value` = reference-definition rays per second (every closest-hit or visibility query the reference algorithm
needs, per wavelength path) with the frame buffer resident in HBM; `spp_per_s` (pixel-samples per second, needs no ray
definition) and `traced_rays_per_s_M` (what the GPU really traverses) sit beside it.  `e2e` = the same through the host-buffer
entry point b2pt_render (camera and parameters down, the frame back up inside the timed region).  Prints ONE JSON line on rank 0.

Other BASELINE configurations run under the same harness: --scene cornell --width 512 --height 512 --frame-spp 32 (configs[0]),
--no-dof (configs[2]), --quality high --gem (configs[3]), --scene sweep:<material> --width 1024 --height 1024 --frame-spp 4096
(configs[4]).  --single-process times the library's own multi-GPU entry point (b2pt_group_render: one process, N contexts,
ncclReduce through dlopen — what ./RayTracing --gpus N runs) instead of one torchrun rank per GPU.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WIDTH, HEIGHT = 1920, 1080
RAY_RECORD_BYTES = 64  # R of SURVEY.md 8(d): ray in + hit out
SEED = 0x5EED0001


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--frame-spp", type=int, default=2048, help="samples per pixel of the frame = one step (split over the GPUs)")
    ap.add_argument("--ndir", type=int, default=32, help="next-event samples per vertex (configs[1]: 32; the reference's effective value: 4)")
    ap.add_argument("--scene", default="chess", help="chess | cornell | sweep:<material name>")
    ap.add_argument("--width", type=int, default=0)
    ap.add_argument("--height", type=int, default=0)
    ap.add_argument("--no-dof", action="store_true", help="configs[2]: DoF off, black environment")
    ap.add_argument("--quality", default="low", choices=["low", "high"])
    ap.add_argument("--gem", action="store_true", help="configs[3]: every chess piece smooth_glass_gem")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-variants", action="store_true")
    ap.add_argument("--queue", type=int, default=0, help="ray queue target (0 = library default)")
    ap.add_argument("--cpu-sample", default="", help="WxHxSPP of the bounded CPU sample (default: quarter-size frame, 1 spp at ndir > 8, else 4)")
    ap.add_argument("--single-process", action="store_true", help="time b2pt_group_render over --gpus devices from one process")
    a = ap.parse_args()
    if a.width <= 0 or a.height <= 0:
        a.width, a.height = (WIDTH, HEIGHT) if a.scene == "chess" else ((512, 512) if a.scene == "cornell" else (1024, 1024))
    return a


def load_pkg():
    import b2pt_loader
    return b2pt_loader.load()


def make_scene(args, width=None, height=None):
    b2pt = load_pkg()
    from b2pt import scenes
    w, h = width or args.width, height or args.height
    if args.scene == "chess":
        fix = b2pt.FIX_MODEL_QUALITY if args.quality == "high" else 0
        gem = dict(king="smooth_glass_gem", left="smooth_glass_gem", right="smooth_glass_gem") if args.gem else {}
        return scenes.chess(w, h, dof=not args.no_dof, sky=not args.no_dof, quality=args.quality, n_dir=args.ndir, fix=fix, sky_size=(2048, 1024), **gem)
    if args.scene == "cornell":
        return scenes.cornell(w, h, n_dir=args.ndir)
    if args.scene.startswith("sweep:"):
        return scenes.cornell_sweep(args.scene.split(":", 1)[1], w, h, n_dir=args.ndir)
    raise SystemExit(f"unknown --scene {args.scene}")


def workload_name(args):
    if args.scene == "chess":
        cfg = 3 if (args.gem or args.quality == "high") else (2 if args.no_dof else 1)
        return (f"chess {args.width}x{args.height} {'DoF+sky' if not args.no_dof else 'noDoF+dark'} {args.quality}-poly{' gem' if args.gem else ''} "
                f"rr0.4 nee{args.ndir} spp{args.frame_spp} (BASELINE configs[{cfg}])")
    if args.scene == "cornell":
        return f"Cornell box DEMO {args.width}x{args.height} rr0.7 nee{args.ndir} spp{args.frame_spp} (BASELINE configs[0])"
    return f"Cornell box, {args.scene.split(':', 1)[1]} on spheres and boxes, {args.width}x{args.height} nee{args.ndir} spp{args.frame_spp} (BASELINE configs[4])"


def metric_name(args):
    return "Mrays/s (1080p chess scene)" if args.scene == "chess" else "Mrays/s (Cornell box)"


class ClockSampler(threading.Thread):
    """Samples SM clocks and throttle reasons during the timed region (NVML; nvidia-smi as a fallback)."""

    REASONS = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.stop_flag = threading.Event()
        self.sm, self.max_sm, self.mask = [], None, 0
        self.nv = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            # NVML indexes physical devices; honour CUDA_VISIBLE_DEVICES when it lists integers
            vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
            ids = [int(x) for x in vis.split(",") if x.strip().isdigit()]
            phys = ids[index] if index < len(ids) else index
            self.h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_sm = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def sample(self):
        if self.nv is not None:
            try:
                self.sm.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                self.mask |= int(self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                return
            except Exception:
                try:
                    self.mask |= int(self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                    return
                except Exception:
                    pass
        try:
            q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
                "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
            out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                 capture_output=True, text=True, timeout=5).stdout.strip().split(",")
            out = [x.strip() for x in out]
            if out and out[0].isdigit():
                self.sm.append(int(out[0]))
                self.max_sm = int(out[1])
                for k, n in enumerate(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]):
                    if len(out) > 2 + k and out[2 + k].lower().startswith("active"):
                        self.mask |= self.REASONS[n]
        except Exception:
            pass

    def run(self):
        while not self.stop_flag.is_set():
            self.sample()
            self.stop_flag.wait(0.005 if self.nv is not None else 0.2)

    def summary(self):
        self.stop_flag.set()
        self.join(timeout=6)
        sm = sorted(self.sm)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": self.max_sm, "reasons": [n for n, b in self.REASONS.items() if self.mask & b],
                "samples": len(sm), "source": "nvml" if self.nv is not None else "nvidia-smi"}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ---- the reference's own CPU implementation (oracle/_ref: its unmodified sources) --------------------------------------
def cpu_sample_shape(args):
    """(width, height, spp) of the bounded CPU sample: the same scene and camera on a quarter-size frame (a pixel-sample costs the
    same whatever the frame size), sized for ~10 s of CPU work per step."""
    if args.cpu_sample:
        w, h, s = (int(x) for x in args.cpu_sample.lower().split("x"))
        return w, h, s
    heavy = args.ndir > 8
    if args.scene == "chess":
        return args.width // 2, args.height // 2, 1 if heavy else 4
    return min(args.width, 256), min(args.height, 256), 2 if heavy else 8


def cpu_reference_run(args, rays_per_path, repeats=1, warmup=0):
    """Times Renderer::Render of the reference (8 OpenMP threads hard-coded, Renderer.cpp:16,36).  The unmodified reference
    cannot count rays: rays = paths x the rays-per-path ratio of the same algorithm measured by the GPU counters (and
    cross-checked against the oracle's own count by tests/test_gpu_configs.py)."""
    from oracle import refbind as R

    w, h, spp = cpu_sample_shape(args)
    sc, env_png = make_scene(args, w, h)
    ref = R.Ref(sc, env_png)
    times = []
    with tempfile.TemporaryDirectory() as td:
        for it in range(warmup + repeats):
            t0 = time.perf_counter()
            ref.L.ref_render_real(ref.h, spp, os.path.join(td, "out.png").encode())
            if it >= warmup:
                times.append(time.perf_counter() - t0)
    ref.close()
    sc.close()
    paths = 3.0 * w * h * spp
    return times, (w, h, spp), paths * rays_per_path


def rays_per_path_for(args, measured=None):
    ratio_file = os.path.join(ROOT, "profiles", "rays_per_path.json")
    key = workload_name(args)
    table = json.load(open(ratio_file)) if os.path.exists(ratio_file) else {}
    if measured is not None:
        try:
            table[key] = measured
            os.makedirs(os.path.dirname(ratio_file), exist_ok=True)
            json.dump(table, open(ratio_file, "w"), indent=1, sort_keys=True)
        except Exception:
            pass
        return measured
    if key in table:
        return table[key]
    return 14.5 if args.ndir == 32 else 2.82  # chess scene, device counters = oracle count (tests/test_gpu_configs.py)


def cpu_baseline_record(args, times, shape, rays):
    w, h, spp = shape
    t = sum(times) / len(times)
    return {"value": rays / t / 1e6, "unit": "Mrays/s", "cores": 8, "host_cores": os.cpu_count(), "kind": "reference", "seconds": t,
            "spp_per_s": w * h * spp / t,
            "sample": f"Renderer::Render of the unmodified reference (oracle/_ref), same scene and camera on a {w}x{h} frame at spp={spp} "
                      f"({w * h * spp} pixel-samples per step), 8 OpenMP threads (hard-coded in Renderer.cpp:16), rays = paths x rays/path of the GPU counters"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import refbind as R

    if not R.have_ref():
        emit({"impl": "reference", "unavailable": "oracle/_ref/libref_oracle.so is not built (needs /root/reference at build time)"})
        return
    rpp = rays_per_path_for(args)
    times, shape, rays = cpu_reference_run(args, rpp, repeats=max(args.steps, 1), warmup=max(args.warmup, 0))
    rec = cpu_baseline_record(args, times, shape, rays)
    w, h, spp = shape
    line = {
        "impl": "reference", "metric": metric_name(args), "value": rec["value"], "unit": "Mrays/s", "n_gpus": args.gpus,
        "steps": len(times), "warmup": max(args.warmup, 0), "ms_per_step": rec["seconds"] * 1e3, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(args), "frame_spp": args.frame_spp, "spp_per_step": spp, "cpu_sample": f"{w}x{h}x{spp}", "warmup": max(args.warmup, 0),
                   "threads": 8, "note": "each step is a bounded sample of the workload (same scene, camera, light samples; smaller frame, fewer spp): "
                                         "the full 2048-spp frame takes the CPU path many hours"},
        "spp_per_s": rec["spp_per_s"],
        "projected_s_frame": args.width * args.height * args.frame_spp / rec["spp_per_s"],
        "cpu_baseline": rec,
        "e2e": {"value": rec["value"], "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)


def ctypes_sizeof_inputs(b2pt):
    """Bytes a render step sends to the device: the camera and the render parameters (the scene is resident)."""
    import ctypes
    return ctypes.sizeof(b2pt.Camera) + ctypes.sizeof(b2pt.RenderParams)


def roofline_record(args, ctx, st_count, rays_closest, rays_shadow, ext_ms, sh_ms, ext_launches, gpu_ms):
    """Roofline of the dominant kernel (extend): algorithmic bytes = 32 B x child boxes tested + 48 B x primitives tested + 64 B ray
    record per ray, counts from the stats pass on the same workload; achieved = bytes of an average launch /
    its duration (CUDA events around every launch, inside b2pt_render*).  The scene is cache-resident, so the bandwidth that can
    bound the walk is the L2's (measured live here); the HBM figure is kept beside it."""
    ext_rays = max(st_count.rays_traced_closest, 1)
    bytes_per_ray = (32.0 * st_count.extend_nodes + 48.0 * st_count.extend_prims) / ext_rays + RAY_RECORD_BYTES
    sh_rays = max(st_count.rays_traced_shadow, 1)
    sh_bytes_per_ray = (32.0 * st_count.shadow_nodes + 48.0 * st_count.shadow_prims) / sh_rays + RAY_RECORD_BYTES
    hbm, hbm_src = measured_peak()
    ext_launch_ms = ext_ms / max(ext_launches, 1)
    achieved = (rays_closest / max(ext_launches, 1)) * bytes_per_ray / (ext_launch_ms * 1e-3) / 1e9 if ext_ms > 0 else 0.0
    sh_achieved = rays_shadow * sh_bytes_per_ray / (sh_ms * 1e-3) / 1e9 if sh_ms > 0 else 0.0
    try:
        l2_gbs = ctx.measure_l2_read_gbs()
    except Exception:
        l2_gbs = None
    traffic, traffic_src = None, None
    try:  # DRAM bytes per launch from the committed ncu capture of this build (per ray x rays of an average launch)
        tj = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
        traffic = tj["extend_kernel"]["dram_bytes_per_ray"] * (rays_closest / max(ext_launches, 1))
        traffic_src = tj.get("source")
    except Exception:
        pass
    peak = l2_gbs if l2_gbs else hbm
    return {"bound": "l1/l2" if l2_gbs else "hbm", "kernel": "extend_kernel", "achieved": achieved, "peak": peak, "unit": "GB/s",
            "frac": achieved / peak if peak else None, "traffic": traffic, "traffic_source": traffic_src,
            "peak_source": "measured live: b2pt_measure_l2_read_gbs (64 passes of LDG.128 over a 32 MiB buffer, slices rotated between SMs)" if l2_gbs else hbm_src,
            "hbm": {"peak": hbm, "frac": achieved / hbm, "peak_source": hbm_src,
                    "note": "the tree and the primitives are cache-resident (DRAM traffic = the queue records), so the algorithmic bytes are not HBM bytes and this fraction can exceed 1"},
            "bytes_per_ray": bytes_per_ray, "nodes_per_ray": st_count.extend_nodes / ext_rays, "tris_per_ray": st_count.extend_prims / ext_rays,
            "avg_launch_ms": ext_launch_ms, "launches": ext_launches, "share_of_step": ext_ms / gpu_ms if gpu_ms else None,
            "shadow_kernel": {"achieved": sh_achieved, "frac": sh_achieved / peak if peak else None, "bytes_per_ray": sh_bytes_per_ray,
                              "nodes_per_ray": st_count.shadow_nodes / sh_rays, "share_of_step": sh_ms / gpu_ms if gpu_ms else None}}


def run_single_process(args):
    """The library's own multi-GPU path: one process, one context per device, b2pt_group_render (spp split, one ncclReduce per
    frame through the dlopen'ed NCCL).  Host-buffer entry point, so the timing is end to end by construction."""
    import numpy as np
    b2pt = load_pkg()
    n = args.gpus
    sc, env_png = make_scene(args)
    cam = sc.camera
    ctxs = [b2pt.Context(i).upload(sc) for i in range(n)]
    out = np.zeros((cam.height, cam.width, 3), np.float32)
    spp = args.frame_spp
    for w in range(args.warmup):
        b2pt.group_render(ctxs, cam, spp, seed=SEED + w, out=out, flags=b2pt.FLAG_FRESH_FRAME)
    sampler = ClockSampler(0)
    sampler.start()
    ms, rays, traced, launches = 0.0, 0, 0, 0
    for k in range(args.steps):
        t0 = time.perf_counter()
        _, st = b2pt.group_render(ctxs, cam, spp, seed=SEED + args.warmup + k, out=out, flags=b2pt.FLAG_FRESH_FRAME)
        ms += (time.perf_counter() - t0) * 1e3
        rays += st.rays_reference; traced += st.rays_traced_closest + st.rays_traced_shadow; launches += st.kernel_launches
    clocks = sampler.summary()
    ref, _ = ctxs[0].render(cam, spp, seed=SEED + args.warmup + args.steps - 1)
    fin = ~(np.isnan(ref) | np.isnan(out))  # NaN samples poison their pixel in the reference too: same pixels on both sides
    diff = float(np.abs(ref - out)[fin].max())
    nan_mismatches = int((np.isnan(ref) ^ np.isnan(out)).sum())
    pix = cam.width * cam.height
    K = args.steps
    val = rays / (ms * 1e-3) / 1e6
    line = {"metric": metric_name(args), "value": val, "unit": "Mrays/s", "n_gpus": n, "steps": K, "warmup": args.warmup, "ms_per_step": ms / K,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(args), "frame_spp": spp, "parallelism": f"ONE process, {n} contexts, b2pt_group_render: spp split, one ncclReduce (dlopen) per frame",
                       "timing": "host clock around the blocking call (camera + parameters down, frame back up): value == e2e"},
            "spp_per_s": pix * spp * K / (ms * 1e-3), "traced_rays_per_s_M": traced / (ms * 1e-3) / 1e6, "seconds_per_frame": ms / K / 1e3,
            "multi_gpu_max_abs_diff": diff, "multi_gpu_check": {"nan_values": int(np.isnan(out).sum()), "nan_mismatches": nan_mismatches}, "frame_mean": float(out[fin].mean()),
            "e2e": {"value": val, "unit": "Mrays/s", "h2d_bytes_per_step": ctypes_sizeof_inputs(b2pt), "d2h_bytes_per_step": pix * 12, "ms_per_step": ms / K},
            "gpu_launches": int(launches), "clocks": clocks}
    emit(line)
    for c in ctxs:
        c.close()


def run_ours(args):
    import numpy as np
    import torch
    b2pt = load_pkg()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)

    sc, env_png = make_scene(args)
    cam = sc.camera
    ctx = b2pt.Context(local)  # raises when libb2pt.so or the GPU is missing: no CPU fallback
    ctx.upload(sc)
    stream = torch.cuda.current_stream(dev)
    ctx.set_stream(stream.cuda_stream)

    from b2pt.multigpu import reduce_frame, sample_range

    spp_frame = args.frame_spp
    my_begin, my_count = sample_range(rank, world, spp_frame)  # strong scaling: the frame's samples are split over the ranks
    fb = torch.zeros((cam.height, cam.width, 3), dtype=torch.float32, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2
    host_fb = torch.zeros((cam.height, cam.width, 3), dtype=torch.float32).pin_memory()
    host_np = host_fb.numpy()

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def step_device(step, flags=0, begin=my_begin, count=my_count, total=spp_frame, reduce=True):
        fb.zero_()
        st = None
        if count > 0:
            st = ctx.render_device(cam, total, fb.data_ptr(), sample_begin=begin, sample_count=count, seed=SEED + step, flags=flags, max_wave_bundles=args.queue)
        if reduce:
            reduce_frame(fb, dist, 0)  # ONE reduce per frame
        return st

    def step_e2e(step):
        if dist is None:
            # a fresh frame into the caller's pinned buffer: camera + parameters go down, the frame comes back
            _, st = ctx.render(cam, spp_frame, seed=SEED + step, sample_begin=0, sample_count=spp_frame, out=host_np, max_wave_bundles=args.queue,
                               flags=b2pt.FLAG_FRESH_FRAME)
        else:
            st = step_device(step)
            if rank == 0:
                host_fb.copy_(fb, non_blocking=True)
            torch.cuda.synchronize(dev)
        return st

    # traversal counts of this workload (stats build of the kernels, outside the timed region; a 64-spp slice of the frame)
    st_count = step_device(1000, flags=b2pt.FLAG_COUNT_TRAVERSAL, begin=0, count=min(64, spp_frame), reduce=False)
    barrier()

    for w in range(args.warmup):
        step_device(w)
    barrier()

    sampler = ClockSampler(local)
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    gpu_ms, ext_ms, sh_ms = 0.0, 0.0, 0.0
    launches = ext_launches = sh_launches = 0
    rays = rays_closest = rays_shadow = 0
    for k in range(args.steps):
        flush.fill_(k & 0xFF)
        barrier()
        ev0.record(stream)
        st = step_device(args.warmup + k)
        ev1.record(stream)
        barrier()
        gpu_ms += ev0.elapsed_time(ev1)
        if st is not None:
            ext_ms += st.extend_ms; sh_ms += st.shadow_ms
            launches += st.kernel_launches; ext_launches += st.extend_launches; sh_launches += st.shadow_launches
            rays += st.rays_reference; rays_closest += st.rays_traced_closest; rays_shadow += st.rays_traced_shadow
    clocks = sampler.summary()
    last_seed_step = args.warmup + args.steps - 1

    # multi-GPU correctness, outside the timed region: the frame of the last timed step (still in fb on rank 0) against the same
    # samples rendered by rank 0 alone
    multi = None
    if dist is not None:
        multi_fb = fb.clone() if rank == 0 else None
        barrier()
        if rank == 0:
            fb.zero_()
            ctx.render_device(cam, spp_frame, fb.data_ptr(), sample_begin=0, sample_count=spp_frame, seed=SEED + last_seed_step, max_wave_bundles=args.queue)
            torch.cuda.synchronize(dev)
            # a sample that evaluates to NaN poisons its pixel in the reference too (framebuffer += NaN, Renderer.cpp:80; the tone map
            # sends it to 255): such pixels must be the SAME pixels on both sides, the others are compared numerically
            nan_a, nan_b = torch.isnan(fb), torch.isnan(multi_fb)
            fin = ~(nan_a | nan_b)
            d = (fb - multi_fb).abs()[fin]
            multi = {"max_abs_diff": float(d.max()), "mean_abs_diff": float(d.mean()), "frame_mean": float(fb[fin].mean()),
                     "nan_values": int(nan_a.sum()), "nan_mismatches": int((nan_a ^ nan_b).sum()),
                     "how": f"rank 0 alone renders the {spp_frame} samples of the last timed frame (same seed) and compares with the {world}-rank reduce; "
                            "differences are float summation order (atomic adds, reduce tree)"}
        barrier()

    # end to end through the host-buffer API
    step_e2e(0)
    barrier()
    e2e_ms = 0.0
    e2e_rays = 0
    for k in range(args.steps):
        flush.fill_(k & 0xFF)
        barrier()
        t0 = time.perf_counter()
        st = step_e2e(args.warmup + k)
        barrier()
        e2e_ms += (time.perf_counter() - t0) * 1e3
        e2e_rays += st.rays_reference if st is not None else 0

    # max over ranks of the times, sum over ranks of the work
    t = torch.tensor([gpu_ms, e2e_ms], dtype=torch.float64, device=dev)
    w = torch.tensor([float(rays), float(e2e_rays), float(launches), float(rays_closest + rays_shadow)], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(w, op=dist.ReduceOp.SUM)
    gpu_ms_max, e2e_ms_max = t.tolist()
    rays_all, e2e_rays_all, launches_all, traced_all = w.tolist()

    if rank == 0:
        K = args.steps
        value = rays_all / (gpu_ms_max * 1e-3) / 1e6
        e2e_value = e2e_rays_all / (e2e_ms_max * 1e-3) / 1e6
        pix = cam.width * cam.height
        spp_per_s = pix * spp_frame * K / (gpu_ms_max * 1e-3)
        paths_frame = 3.0 * pix * spp_frame
        line = {
            "metric": metric_name(args), "value": value, "unit": "Mrays/s", "n_gpus": world, "steps": K, "warmup": args.warmup,
            "ms_per_step": gpu_ms_max / K, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "spp_per_s": spp_per_s,
            "traced_rays_per_s_M": traced_all / (gpu_ms_max * 1e-3) / 1e6,
            "seconds_per_frame": gpu_ms_max / K / 1e3,
            "config": {"workload": workload_name(args), "frame_spp": spp_frame, "spp_per_gpu_per_step": my_count, "spp_per_step": spp_frame,
                       "step": f"one frame of {spp_frame} samples per pixel = the whole job; ms_per_step is the frame time",
                       "parallelism": f"spp-split x{world} of the fixed frame (strong scaling), ONE NCCL reduce of the fp32 frame per frame" if world > 1 else "single GPU",
                       "l2": "256 MiB buffer written between timed steps (L2 flush); queue buffers exceed L2",
                       "env_map": "synthetic 2048x1024 sky (upstream sky.png is absent)" if (args.scene == "chess" and not args.no_dof) else "constant background",
                       "ray_definition": "value counts the rays the reference algorithm needs, per wavelength path (SURVEY 8d): shadow rays of light samples "
                                         "proven to contribute exactly zero are counted (the reference traces them) but not traced here; "
                                         "traced_rays_per_s_M is what the GPU really traverses, spp_per_s needs no ray definition"},
            "mpaths_per_s": paths_frame * K / (gpu_ms_max * 1e-3) / 1e6,
            "rays_per_path": rays_all / (paths_frame * K),
            "e2e": {"value": e2e_value, "unit": "Mrays/s", "h2d_bytes_per_step": ctypes_sizeof_inputs(b2pt), "d2h_bytes_per_step": pix * 12,
                    "ms_per_step": e2e_ms_max / K, "spp_per_s": pix * spp_frame * K / (e2e_ms_max * 1e-3)},
            "gpu_launches": int(launches_all),
            "clocks": clocks,
            "roofline": roofline_record(args, ctx, st_count, rays_closest, rays_shadow, ext_ms, sh_ms, ext_launches, gpu_ms),
        }
        if multi is not None:
            line["multi_gpu_max_abs_diff"] = multi["max_abs_diff"]
            line["multi_gpu_check"] = multi
        rays_per_path_for(args, measured=line["rays_per_path"])  # kept for the reference arm (it cannot count rays itself)
        if world == 1 and not args.no_cpu_baseline:
            from oracle import refbind as R
            if R.have_ref():
                # RMSE vs the CPU reference (the metric's third part): the reference's own castRay replayed on the sample streams the
                # GPU used, for a pixel subset of the frame (same scene, same camera, 8 spp) — outside every timed region
                try:
                    ref = R.Ref(sc, env_png)
                    px = np.random.RandomState(0).choice(pix, 512, replace=False).astype(np.int32)
                    g, _ = ctx.render_samples(cam, px, 0, 8)
                    r = ref.render_samples(px, 0, 8)
                    ref.close()
                    gm, rm = g.mean(1), r.mean(1)
                    line["rmse_vs_cpu_ref"] = {"rmse": float(np.sqrt(np.mean((gm - rm) ** 2))), "mean_radiance": float(rm.mean()),
                                               "max_abs_diff_per_sample": float(np.abs(g - r).max()), "pixels": 512, "spp": 8,
                                               "how": "same sample streams on both sides (Philox keyed by pixel, sample)"}
                except Exception as e:  # the checker is optional: never fail the bench line over it
                    line["rmse_vs_cpu_ref"] = {"error": str(e)[:200]}
                times, shape, cpu_rays = cpu_reference_run(args, line["rays_per_path"])
                line["cpu_baseline"] = cpu_baseline_record(args, times, shape, cpu_rays)
        if world == 1 and not args.no_variants and args.scene == "chess" and args.ndir != 4:
            # the reference's EFFECTIVE configuration (its main() never calls setDirectLightSample): 4 light samples per vertex.
            # A fresh context, so that the ray queue is sized for this configuration like a program run with it would.
            try:
                ctx.close()
                ctx = b2pt.Context(local)
                ctx.upload(sc)
                ctx.set_stream(stream.cuda_stream)
                ctx.set_params(n_dir_sample=4)
                vs = min(512, spp_frame)
                step_device(2000, count=vs, begin=0)
                torch.cuda.synchronize(dev)
                vms, vrays, vtr = 0.0, 0, 0
                for k in range(2):
                    flush.fill_(k)
                    torch.cuda.synchronize(dev)
                    ev0.record(stream)
                    st = step_device(2001 + k, count=vs, begin=0)
                    ev1.record(stream)
                    torch.cuda.synchronize(dev)
                    vms += ev0.elapsed_time(ev1); vrays += st.rays_reference; vtr += st.rays_traced_closest + st.rays_traced_shadow
                line["variants"] = {"nee4": {"what": "the same frame with the reference's effective 4 light samples per vertex (conf.json's 32 is never read by its main())",
                                             "spp_per_step": vs, "steps": 2, "spp_per_s": pix * vs * 2 / (vms * 1e-3), "value_Mrays_per_s": vrays / (vms * 1e-3) / 1e6,
                                             "traced_rays_per_s_M": vtr / (vms * 1e-3) / 1e6, "seconds_per_2048spp_frame": 2048.0 * pix / (pix * vs * 2 / (vms * 1e-3))}}
            except Exception as e:
                line["variants"] = {"error": str(e)[:200]}
        emit(line)
    ctx.close()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def emit(line):
    """The ONE line of this run, on the real stdout."""
    os.write(REAL_STDOUT, (json.dumps(line) + "\n").encode())


REAL_STDOUT = 1


def main():
    global REAL_STDOUT
    # Everything else that writes to fd 1 (the reference's own printf/cout chatter when its scene is built and rendered)
    # goes to stderr, so stdout carries exactly one JSON line.
    sys.stdout.flush()
    REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    elif args.single_process:
        run_single_process(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
