#!/usr/bin/env python
"""bench.py — the reference's headline benchmark on B200: Mrays/s and spp/s on the 1080p chess scene.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--spp-per-step S] [--ndir D]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

Workload (BASELINE.json configs[1]): the shipped conf.json chess scene, 1920x1080, thin-lens DoF, sky env map
(synthetic: the upstream sky.png is absent), Russian roulette 0.4, and the reference's EFFECTIVE next-event
sample count 4 (conf.json's directLightSample:32 is never read, SURVEY.md note 2; --ndir 32 runs the north-star
variant).  One step = one pass of the hot path (Renderer::Render's pixel loop) over S samples per pixel of the
full frame on every GPU; with N GPUs the samples are split across ranks (weak scaling: S per GPU per step) and
the fp32 radiance buffers are summed on rank 0 with one NCCL reduce per step.

`value` = reference-definition rays per second (every closest-hit or visibility query the
reference algorithm needs, counted per wavelength path) with the frame buffer resident in HBM; `e2e` = the same
through the host-buffer entry point b2pt_render (H2D of the frame buffer, D2H of the result inside the timed
region).  Prints ONE JSON line on rank 0.
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
sys.path.insert(0, os.path.join(ROOT, "tests"))

WIDTH, HEIGHT = 1920, 1080
RAY_RECORD_BYTES = 64  # R of SURVEY.md 8(d): ray in + hit out


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--spp-per-step", type=int, default=64, help="samples per pixel per GPU per step")
    ap.add_argument("--ndir", type=int, default=4, help="next-event samples per vertex (reference effective value: 4)")
    ap.add_argument("--width", type=int, default=WIDTH)
    ap.add_argument("--height", type=int, default=HEIGHT)
    ap.add_argument("--no-dof", action="store_true", help="configs[2]: DoF off, black environment")
    ap.add_argument("--quality", default="low", choices=["low", "high"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--queue", type=int, default=0, help="ray queue target (0 = library default)")
    ap.add_argument("--cpu-sample-spp", type=int, default=3, help="spp of the bounded CPU sample (full frame)")
    return ap.parse_args()


def make_scene(args):
    import scenes
    import support as S

    fix = S.b2pt.FIX_MODEL_QUALITY if args.quality == "high" else 0
    sc, env_png = scenes.chess(args.width, args.height, dof=not args.no_dof, sky=not args.no_dof, quality=args.quality, n_dir=args.ndir, fix=fix)
    return sc, env_png


def workload_name(args):
    return (f"chess {args.width}x{args.height} {'DoF+sky' if not args.no_dof else 'noDoF+dark'} {args.quality}-poly "
            f"rr0.4 nee{args.ndir} (BASELINE configs[{2 if args.no_dof else 1}])")


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
def cpu_reference_run(sc, env_png, spp, rays_per_path, repeats=1):
    """Times Renderer::Render of the reference (8 OpenMP threads hard-coded, Renderer.cpp:16,36) on the full frame at
    `spp`.  Rays are not counted by the unmodified reference; rays = paths x the rays-per-path ratio of the same
    algorithm measured on the GPU path."""
    import support as S

    ref = S.Ref(sc, env_png)
    cam = sc.camera
    times = []
    with tempfile.TemporaryDirectory() as td:
        for _ in range(repeats):
            t0 = time.perf_counter()
            devnull = os.open(os.devnull, os.O_WRONLY)
            saved = os.dup(1)
            os.dup2(devnull, 1)  # the reference prints a progress bar
            try:
                ref.L.ref_render_real(ref.h, spp, os.path.join(td, "out.png").encode())
            finally:
                os.dup2(saved, 1)
                os.close(devnull)
                os.close(saved)
            times.append(time.perf_counter() - t0)
    ref.close()
    paths = 3.0 * cam.width * cam.height * spp
    return times, paths, paths * rays_per_path


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import support as S

    if not S.have_ref():
        emit({"impl": "reference", "unavailable": "oracle/_ref/libref_oracle.so is not built (needs /root/reference at build time)"})
        return
    sc, env_png = make_scene(args)
    ratio_file = os.path.join(ROOT, "profiles", "rays_per_path.json")
    key = workload_name(args)
    rays_per_path = None
    if os.path.exists(ratio_file):
        rays_per_path = json.load(open(ratio_file)).get(key)
    if rays_per_path is None:
        rays_per_path = 2.77  # CPU-counted unique rays per path, chess, NEE=4
    spp = args.cpu_sample_spp
    for _ in range(min(args.warmup, 1)):
        cpu_reference_run(sc, env_png, spp, rays_per_path)
    times, paths, rays = cpu_reference_run(sc, env_png, spp, rays_per_path, repeats=max(args.steps, 1))
    t = sum(times) / len(times)
    val = rays / t / 1e6
    cores = os.cpu_count()
    line = {
        "impl": "reference", "metric": "Mrays/s (1080p chess scene)", "value": val, "unit": "Mrays/s", "n_gpus": args.gpus,
        "steps": len(times), "warmup": min(args.warmup, 1), "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": key, "spp_per_step": spp, "threads": 8},
        "spp_per_s": cam_pixels(sc) * spp / t,
        "cpu_baseline": {"value": val, "unit": "Mrays/s", "cores": 8, "host_cores": cores, "kind": "reference",
                         "sample": f"Renderer::Render of the unmodified reference (oracle/_ref), full {args.width}x{args.height} frame at spp={spp}, "
                                   f"8 OpenMP threads (hard-coded), rays = paths x {rays_per_path:.3f} rays/path"},
        "e2e": {"value": val, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)


def ctypes_sizeof_inputs(b2pt):
    """Bytes a render step sends to the device: the camera and the render parameters (the scene is resident)."""
    import ctypes
    return ctypes.sizeof(b2pt.Camera) + ctypes.sizeof(b2pt.RenderParams)


def cam_pixels(sc):
    return sc.camera.width * sc.camera.height


def run_ours(args):
    import numpy as np
    import torch
    import support as S

    b2pt = S.b2pt
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

    S_step = args.spp_per_step
    spp_total = S_step * world  # samples per pixel of one step's frame
    my_begin, my_count = sample_range(rank, world, spp_total)
    fb = torch.zeros((cam.height, cam.width, 3), dtype=torch.float32, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2
    host_fb = torch.zeros((cam.height, cam.width, 3), dtype=torch.float32).pin_memory()
    host_np = host_fb.numpy()

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def step_device(step, flags=0):
        fb.zero_()
        st = ctx.render_device(cam, spp_total, fb.data_ptr(), sample_begin=my_begin, sample_count=my_count, seed=S.SEED + step, flags=flags, max_wave_bundles=args.queue)
        reduce_frame(fb, dist, 0)
        return st

    def step_e2e(step):
        if dist is None:
            # a fresh frame into the caller's pinned buffer: camera + parameters go down, the frame comes back
            _, st = ctx.render(cam, spp_total, seed=S.SEED + step, sample_begin=0, sample_count=S_step, out=host_np, max_wave_bundles=args.queue,
                               flags=b2pt.FLAG_FRESH_FRAME)
        else:
            fb.zero_()
            st = ctx.render_device(cam, spp_total, fb.data_ptr(), sample_begin=my_begin, sample_count=my_count, seed=S.SEED + step, max_wave_bundles=args.queue)
            reduce_frame(fb, dist, 0)
            if rank == 0:
                host_fb.copy_(fb, non_blocking=True)
            torch.cuda.synchronize(dev)
        return st

    # traversal counts of this workload (stats build of the kernels, outside the timed region)
    st_count = step_device(1000, flags=b2pt.FLAG_COUNT_TRAVERSAL)
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
        ext_ms += st.extend_ms; sh_ms += st.shadow_ms
        launches += st.kernel_launches; ext_launches += st.extend_launches; sh_launches += st.shadow_launches
        rays += st.rays_reference; rays_closest += st.rays_traced_closest; rays_shadow += st.rays_traced_shadow
    clocks = sampler.summary()

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
        e2e_rays += st.rays_reference

    # max over ranks of the times, sum over ranks of the work
    t = torch.tensor([gpu_ms, e2e_ms], dtype=torch.float64, device=dev)
    w = torch.tensor([float(rays), float(e2e_rays), float(launches)], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(w, op=dist.ReduceOp.SUM)
    gpu_ms_max, e2e_ms_max = t.tolist()
    rays_all, e2e_rays_all, launches_all = w.tolist()

    if rank == 0:
        K = args.steps
        value = rays_all / (gpu_ms_max * 1e-3) / 1e6
        e2e_value = e2e_rays_all / (e2e_ms_max * 1e-3) / 1e6
        pix = cam.width * cam.height
        spp_per_s = pix * spp_total * K / (gpu_ms_max * 1e-3)
        paths_step = 3.0 * pix * S_step
        # roofline of the dominant kernel (extend): algorithmic bytes = 32 B x nodes fetched + 48 B x triangles tested
        # + 64 B ray record per ray, counts from the stats pass on the same workload
        ext_rays = max(st_count.rays_traced_closest, 1)
        bytes_per_ray = (32.0 * st_count.extend_nodes + 48.0 * st_count.extend_prims) / ext_rays + RAY_RECORD_BYTES
        sh_rays = max(st_count.rays_traced_shadow, 1)
        sh_bytes_per_ray = (32.0 * st_count.shadow_nodes + 48.0 * st_count.shadow_prims) / sh_rays + RAY_RECORD_BYTES
        peak, peak_src = measured_peak()
        ext_launch_ms = ext_ms / max(ext_launches, 1)
        achieved = (rays_closest / max(ext_launches, 1)) * bytes_per_ray / (ext_launch_ms * 1e-3) / 1e9 if ext_ms > 0 else 0.0
        sh_achieved = rays_shadow * sh_bytes_per_ray / (sh_ms * 1e-3) / 1e9 if sh_ms > 0 else 0.0
        try:  # the scene is cache-resident: the bandwidth that can bound the walk is the L2's, measured here (outside the timed region)
            l2_gbs = ctx.measure_l2_read_gbs()
        except Exception:
            l2_gbs = None
        traffic = None
        try:  # DRAM bytes per launch from the committed ncu capture (per ray x rays of an average launch)
            tj = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))["extend_kernel"]
            traffic = tj["dram_bytes_per_ray"] * (rays_closest / max(ext_launches, 1))
        except Exception:
            pass
        line = {
            "metric": "Mrays/s (1080p chess scene)", "value": value, "unit": "Mrays/s", "n_gpus": world, "steps": K, "warmup": args.warmup,
            "ms_per_step": gpu_ms_max / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": workload_name(args), "spp_per_gpu_per_step": S_step, "spp_per_step": spp_total,
                       "parallelism": f"spp-split x{world}, one NCCL reduce of the fp32 frame per step" if world > 1 else "single GPU",
                       "l2": "256 MiB buffer written between timed steps (L2 flush); queue buffers exceed L2",
                       "ray_definition": "rays the reference algorithm needs, per wavelength path (SURVEY 8d): the shadow rays of light samples that are "
                                         "proven to contribute exactly zero are counted (the reference traces them) but not traced here; "
                                         "traced_rays_per_s_M is what the GPU really traces, spp_per_s needs no ray definition"},
            "spp_per_s": spp_per_s,
            "mpaths_per_s": paths_step * world * K / (gpu_ms_max * 1e-3) / 1e6,
            "projected_s_2048spp": 2048.0 * pix / spp_per_s,
            "rays_per_path": rays_all / (paths_step * world * K),
            "traced_rays_per_s_M": (rays_closest + rays_shadow) / (gpu_ms * 1e-3) / 1e6,
            "e2e": {"value": e2e_value, "unit": "Mrays/s", "h2d_bytes_per_step": ctypes_sizeof_inputs(b2pt), "d2h_bytes_per_step": pix * 12,
                    "ms_per_step": e2e_ms_max / K},
            "gpu_launches": int(launches_all),
            "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": "extend_kernel", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                         "note": "algorithmic bytes are served by L1/L2 (the scene is cache-resident), so frac can exceed 1; DRAM traffic is the "
                                 "45 B/ray of queue records; ncu (profiles/): the walk is bound by the L1 data pipe and issue slots",
                         "l2": {"peak": l2_gbs, "frac": (achieved / l2_gbs) if l2_gbs else None, "unit": "GB/s",
                                "how": "b2pt_measure_l2_read_gbs: 64 passes of LDG.128 over a 32 MiB buffer, slices rotated between SMs"},
                         "bytes_per_ray": bytes_per_ray, "nodes_per_ray": st_count.extend_nodes / ext_rays,
                         "tris_per_ray": st_count.extend_prims / ext_rays, "avg_launch_ms": ext_launch_ms,
                         "share_of_step": ext_ms / gpu_ms if gpu_ms else None,
                         "shadow_kernel": {"achieved": sh_achieved, "frac": sh_achieved / peak, "bytes_per_ray": sh_bytes_per_ray,
                                           "nodes_per_ray": st_count.shadow_nodes / sh_rays, "share_of_step": sh_ms / gpu_ms if gpu_ms else None}},
        }
        # keep the rays-per-path ratio for the reference arm (it cannot count rays itself)
        try:
            os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
            rp = os.path.join(ROOT, "profiles", "rays_per_path.json")
            table = json.load(open(rp)) if os.path.exists(rp) else {}
            table[workload_name(args)] = line["rays_per_path"]
            json.dump(table, open(rp, "w"), indent=1, sort_keys=True)
        except Exception:
            pass
        if world == 1 and not args.no_cpu_baseline and S.have_ref():
            # RMSE vs the CPU reference (the metric's third part): the reference's own castRay replayed on the sample streams the
            # GPU used, for a pixel subset of the frame (same scene, same camera, 8 spp) — outside every timed region
            try:
                ref = S.Ref(sc, env_png)
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
            times, paths, cpu_rays = cpu_reference_run(sc, env_png, args.cpu_sample_spp, line["rays_per_path"])
            tt = sum(times) / len(times)
            line["cpu_baseline"] = {
                "value": cpu_rays / tt / 1e6, "unit": "Mrays/s", "cores": 8, "host_cores": os.cpu_count(), "kind": "reference",
                "seconds": tt, "spp_per_s": pix * args.cpu_sample_spp / tt,
                "sample": f"Renderer::Render of the unmodified reference (oracle/_ref), full {cam.width}x{cam.height} frame at "
                          f"spp={args.cpu_sample_spp}, 8 OpenMP threads (hard-coded in Renderer.cpp:16), rays = paths x measured rays/path"}
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
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
