#!/usr/bin/env python
"""Benchmark of the rendering hot path: path-traced rays/s on BASELINE config 3.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c3|c4|c5]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A *step* is one full render of the workload (ImageTracer.fire_all_rays over every pixel sample).
`value` is whole-job rays/s (closest-hit + shadow queries, the events BASELINE.md §3.3 counts) with
the scene resident in HBM and the image left in HBM, timed with CUDA events on the launch stream;
`e2e` is the same metric through the public API (CudaImageTracer.fire_all_rays: flatten, scene
upload, launch, device->host copy of the image) timed on the host clock.  With N > 1 every rank
renders its strata of every pixel and one NCCL all-reduce(sum) of the fp32 image is inside the
timed region (`scaling: strong` — the image is fixed, the work is split).

`--impl reference` times the CPU restatement of the reference (oracle/pt_oracle.c, bit-exact with
the Python reference — see tests/test_oracle_golden.py) on all host threads; the Python reference
itself cannot travel to the GPU box (BASELINE.md holds its measured 14-19 k rays/s per core).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

# Rank 0 prints exactly ONE line on stdout: the JSON.  Libraries write there too (NCCL prints its
# "NCCL version ..." banner on fd 1 whenever NCCL_DEBUG is set), so fd 1 is pointed at stderr for the
# whole run and the JSON line goes to the saved descriptor.
sys.stdout.flush()
_REAL_STDOUT = os.dup(1)
os.dup2(2, 1)


def emit(line: dict) -> None:
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

SCALE = 1
METRIC = "path-traced rays/sec (demo.txt 1080p)"
UNIT = "rays/s"


def workload(name):
    """(world, camera, render kwargs, description, flops per ray)."""
    from pytracer_b200 import scenes

    if name == "c3":
        world, camera = scenes.demo_scene()
        kw = dict(width=1920, height=1080, samples_per_side=8, algorithm="pathtracing", num_of_rays=10, max_depth=3, rr_limit=3)
        desc = "BASELINE config 3: examples/demo.txt pathtracing 1920x1080, 64 spp, num-of-rays 10, max-depth 3, Russian roulette limit 3"
        n_sph, n_pl = 1, 2
    elif name == "c4":
        rs = scenes.random_spheres_scene(1024, 2024, 4, 20.0)
        world, camera = rs.world, rs.camera
        kw = dict(width=3840, height=2160, samples_per_side=4, algorithm="pathtracing", num_of_rays=10, max_depth=3, rr_limit=3)
        desc = "BASELINE config 4: 1024 random ellipsoids + 2 planes, checkered/image pigments, pathtracing 3840x2160, 16 spp"
        n_sph, n_pl = 1024, 2
    elif name == "c5":
        rs = scenes.random_spheres_scene(4096, 2025, 5, 40.0, with_light=True)
        world, camera = rs.world, rs.camera
        kw = dict(width=3840, height=2160, samples_per_side=2, algorithm="pointlight")
        desc = "BASELINE config 5: 4096 random ellipsoids + 2 planes, pointlight 3840x2160, 4 spp"
        n_sph, n_pl = 4096, 2
    else:
        raise SystemExit(f"unknown workload {name}")
    # 54 per sphere test, 12 per plane test, 60 for the winner's record, 46 scatter
    flops_per_ray = 54 * n_sph + 12 * n_pl + 60 + 46
    if SCALE > 1:
        kw["width"], kw["height"] = kw["width"] // SCALE, kw["height"] // SCALE
        desc += f" [SCALED DOWN {SCALE}x per side: experiment, not the benchmark]"
    return world, camera, kw, desc, flops_per_ray


def build_params(kw, camera, **extra):
    from pytracer_b200.params import make_params
    from pytracer_b200.pcg import PCG

    args = dict(kw)
    width, height = args.pop("width"), args.pop("height")
    args.update(aa_pcg=PCG(42, 54), pt_pcg=PCG(45, 54))  # the CLI's generators (main.py:168-185)
    args.update(extra)
    return make_params(width, height, camera, **args)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.proc, self.path = None, None
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={index}", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        try:
            for line in open(self.path):
                parts = [x.strip() for x in line.split(",")]
                if len(parts) < 7:
                    continue
                try:
                    sm.append(float(parts[0]))
                    mx.append(float(parts[1]))
                except ValueError:
                    continue
                for name, val in zip(names, parts[3:7]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            out = {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}
        return out


def cpu_baseline_sample(world, camera, kw, threads):
    """The oracle on a bounded sample of the same workload: full frame, 1 sample per pixel (the
    same scene, resolution, num_of_rays, max_depth), `threads` host threads."""
    from oracle import oracle
    from pytracer_b200.flatten import flatten_world

    flat = flatten_world(world)
    args = dict(kw)
    args["samples_per_side"] = 1
    if flat.n_shapes > 100:  # ms per ray with thousands of shapes: crop like BASELINE.md §3.2
        args["width"], args["height"] = 256, 144
    p = build_params(args, camera)
    t0 = time.perf_counter()
    r = oracle.render_threaded(flat, p, threads) if threads > 1 else oracle.render(flat, p, want_hit=False)
    dt = time.perf_counter() - t0
    rays = r["rays_closest"] + r["rays_shadow"]
    sample = f"{args['width']}x{args['height']} at 1 spp of the same scene/settings ({rays} rays, {dt:.1f} s)"
    return rays / dt, sample


def run_reference(args, rank, world_size):
    if rank != 0:
        return
    world, camera, kw, desc, _ = workload(args.workload)
    threads = os.cpu_count() or 1
    from oracle import oracle

    oracle.build()
    for _ in range(args.warmup):
        cpu_baseline_sample(world, camera, dict(kw, width=480, height=270), threads)
    rates, sample = [], ""
    t_all = time.perf_counter()
    for _ in range(args.steps):
        rate, sample = cpu_baseline_sample(world, camera, kw, threads)
        rates.append(rate)
    total = time.perf_counter() - t_all
    value = statistics.mean(rates)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total / max(1, args.steps), "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": desc, "note": "CPU restatement of the reference (oracle/pt_oracle.c, bit-exact with the Python "
                   "reference on the golden fixtures), one bounded sample per step"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


def run_tonemap(args, rank, local_rank, world_size):
    """Extra line (not the headline): the tone-mapping step that follows a render, HBM-bound.
    A step = average luminosity + normalise/clamp/quantise of one 7680x4320 fp32 frame resident in HBM."""
    import torch

    from oracle import tonemap_oracle
    from pytracer_b200 import tonemap

    if rank != 0:
        return
    torch.cuda.set_device(local_rank)
    h, w = 4320 // SCALE, 7680 // SCALE
    n = h * w
    gen = torch.Generator(device="cuda").manual_seed(11)
    img = torch.exp(torch.randn((h, w, 3), device="cuda", generator=gen)).contiguous()
    ldr = torch.empty((h, w, 3), dtype=torch.uint8, device="cuda")
    flush = torch.zeros(64 * 1024 * 1024, dtype=torch.float32, device="cuda")  # 256 MB, read to evict L2 cleanly
    lum_ms, map_ms, total_ms = [], [], []
    sampler = ClockSampler(local_rank)
    for i in range(args.warmup + args.steps):
        flush.sum()
        st = tonemap.tone_map_device(img.data_ptr(), n, 1.0, None, 1.0, 0, ldr.data_ptr())
        if i >= args.warmup:
            lum_ms.append(st["lum_ms"]); map_ms.append(st["map_ms"]); total_ms.append(st["lum_ms"] + st["map_ms"])
    clocks = sampler.stop()
    # end to end: host image in (pinned), LDR bytes out
    himg = torch.empty((h, w, 3), dtype=torch.float32).pin_memory()
    himg.copy_(img)
    e2e = []
    for i in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        _, out, _ = tonemap.tone_map(himg.numpy(), 1.0, None, 1.0, want_hdr=False)
        dt = time.perf_counter() - t0
        if i >= args.warmup:
            e2e.append(dt)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6534.5))
    traffic = None
    try:
        if SCALE == 1:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get("tonemap")
    except Exception:
        pass
    k_lum, k_map = statistics.mean(lum_ms), statistics.mean(map_ms)
    ms = statistics.mean(total_ms)
    line = {
        "metric": "tone-mapped pixels/sec (average luminosity + normalise/clamp/8-bit, fp32 frame resident in HBM)",
        "value": n / (ms * 1e-3), "unit": "pixels/s", "n_gpus": 1, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64 semantics (fp32 fast path with fp64 guard)",
        "data": "synthetic", "config": {"workload": f"tone mapping of one {w}x{h} frame (hdrimages.py:120-171)", "l2": "256 MB read between steps"},
        "e2e": {"value": n / statistics.mean(e2e), "unit": "pixels/s", "h2d_bytes_per_step": 12 * n, "d2h_bytes_per_step": 3 * n},
        "gpu_launches": 2 * args.steps, "clocks": clocks,
        "roofline": {"bound": "hbm", "achieved": 15 * n / (k_map * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                     "frac": 15 * n / (k_map * 1e-3) / 1e9 / peak, "traffic": traffic, "kernel": "k_tone_map_ldr (12 B read + 3 B written per pixel)",
                     "kernel_ms": k_map, "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6534.5 GB/s",
                     "other_kernel": {"kernel": "k_lum_sum (12 B read per pixel)", "kernel_ms": k_lum,
                                      "achieved": 12 * n / (k_lum * 1e-3) / 1e9, "frac": 12 * n / (k_lum * 1e-3) / 1e9 / peak}},
    }
    if not args.no_cpu_baseline:
        host = img[: max(1, h // 8)].cpu().numpy()
        t0 = time.perf_counter()
        tonemap_oracle.tone_map(host, 1.0, None, 1.0)
        dt = time.perf_counter() - t0
        line["cpu_baseline"] = {"value": host.shape[0] * w / dt, "unit": "pixels/s", "cores": 1, "kind": "port",
                                "sample": f"{w}x{host.shape[0]} rows of the same frame through oracle/tonemap_oracle.py ({dt:.1f} s)"}
    emit(line)


def run_ours(args, rank, local_rank, world_size):
    import torch

    from pytracer_b200 import _abi, device
    from pytracer_b200.device import DeviceScene
    from pytracer_b200.dist import TorchComm, partition_params
    from pytracer_b200.hdrimage import HdrImage
    from pytracer_b200.imagetracer import CudaImageTracer
    from pytracer_b200.pcg import PCG
    from pytracer_b200.render import CudaRenderer

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU path); use --impl reference for the CPU arm")
    torch.cuda.set_device(local_rank)
    comm = TorchComm.from_env("nccl") if world_size > 1 else None
    import torch.distributed as dist

    world, camera, kw, desc, flops_per_ray = workload(args.workload)
    scene = DeviceScene(world)
    params = build_params(kw, camera, variant=args.variant, precision=args.precision, accel=args.accel)
    if world_size > 1:
        params = partition_params(params, rank, world_size)
        if args.partition != "auto":
            params.part_mode = _abi.PARTITIONS[args.partition]
    H, W = params.height, params.width
    image = torch.empty((H, W, 3), dtype=torch.float32, device="cuda")
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")  # > 126 MB of L2
    stream = torch.cuda.current_stream().cuda_stream

    def step():
        scene.render_device(params, image.data_ptr(), 0, stream)
        if comm is not None:
            comm.all_reduce_sum(image)

    for _ in range(max(args.warmup, 0)):
        step()
        scene.finish(stream)
    torch.cuda.synchronize()
    if comm is not None:
        comm.barrier()
    torch.cuda.synchronize()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    events, rays_rank, kernel_ms, launches = [], 0, [], 0
    wall0 = time.perf_counter()
    for _ in range(args.steps):
        flush.zero_()  # evict L2 between timed iterations (not inside the timed interval)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        step()
        e1.record()
        st = scene.finish(stream)
        events.append((e0, e1))
        rays_rank += st["rays_closest"] + st["rays_shadow"]
        kernel_ms.append(st["kernel_ms"])
        launches += st["n_launches"]
    torch.cuda.synchronize()
    if comm is not None:
        comm.barrier()
    torch.cuda.synchronize()
    wall = time.perf_counter() - wall0
    clocks = sampler.stop() if sampler else None
    total_ms = sum(a.elapsed_time(b) for a, b in events)
    t = torch.tensor([total_ms], dtype=torch.float64, device="cuda")
    r = torch.tensor([rays_rank], dtype=torch.int64, device="cuda")
    if comm is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(r, op=dist.ReduceOp.SUM)
    total_ms, rays = float(t.item()), int(r.item())
    value = rays / (total_ms * 1e-3)

    # ---- end to end through the public API: new renderer (flatten + upload), launch, image to host
    scene_bytes = sum(a.nbytes for a in (scene.flat.shape_kind, scene.flat.shape_material, scene.flat.shape_m,
                                         scene.flat.shape_invm, scene.flat.texels)) + \
        sum(map(lambda s: len(bytes(s)), (scene.flat.materials, scene.flat.pigments, scene.flat.lights)))
    himg = HdrImage(W, H)
    e2e_rays, e2e_t = 0, 0.0
    for i in range(args.warmup + args.steps):
        renderer = CudaRenderer(world, algorithm=kw["algorithm"], pcg=PCG(45, 54), num_of_rays=kw.get("num_of_rays", 10),
                                max_depth=kw.get("max_depth", 10), russian_roulette_limit=kw.get("rr_limit", 3),
                                variant=args.variant, precision=args.precision, accel=args.accel)
        tracer = CudaImageTracer(himg, camera, samples_per_side=kw["samples_per_side"], pcg=PCG(42, 54))
        if comm is not None:
            comm.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        tracer.fire_all_rays(renderer, comm=comm)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if i >= args.warmup:
            e2e_t += dt
            e2e_rays += tracer.last_stats["rays_closest"] + tracer.last_stats["rays_shadow"]
    te = torch.tensor([e2e_t], dtype=torch.float64, device="cuda")
    if comm is not None:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = e2e_rays / float(te.item())

    if rank != 0:
        return
    # ---- roofline of the dominant kernel: FP32 FMA pipe, not HBM, not tensor cores
    peak_tf, _ = device.ffma_peak_tflops()
    k_ms = statistics.mean(kernel_ms)
    rays_per_launch_rank = rays_rank / max(1, args.steps)
    achieved_tf = rays_per_launch_rank * flops_per_ray / (k_ms * 1e-3) / 1e12
    prop = torch.cuda.get_device_properties(local_rank)
    nominal_tf = prop.multi_processor_count * 128 * 2 * 1.965e9 / 1e12
    traffic, issue_busy = None, None  # dram__bytes_read.sum + dram__bytes_write.sum per launch, from the committed ncu capture
    try:
        prof = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        if SCALE == 1 and world_size == 1 and args.accel == "none":
            traffic = prof.get(args.workload)
            issue_busy = prof.get(args.workload + "_issue_slots_busy")
    except Exception:
        pass
    roofline = {"bound": "fp32", "achieved": achieved_tf, "peak": peak_tf, "unit": "TFLOP/s", "frac": achieved_tf / peak_tf,
                "traffic": traffic, "peak_source": "FFMA micro-benchmark run in this job (rt_bench_ffma); MEASURED_PEAKS.json holds "
                "only HBM and bf16-tensor peaks, neither bounds this path", "peak_nominal": nominal_tf,
                "flops_per_ray": flops_per_ray, "kernel_ms": k_ms, "rays_per_launch": rays_per_launch_rank}
    if issue_busy is not None:  # demo.txt is 184 flop/ray by construction: the resource that binds it is the issue port
        roofline["issue_slots_busy"] = issue_busy
        roofline["issue_slots_source"] = "smsp__issue_active.avg.pct_of_peak_sustained_active, profiles/traffic.json (committed ncu capture)"
    if args.accel != "none":
        roofline["note"] = ("flops_per_ray is the linear scan's (the reference's algorithm); the hierarchy skips most of that work, "
                            "so `frac` is a speed-up over the roofline of the linear scan, not a utilisation")
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world_size, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": total_ms / max(1, args.steps), "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32" if kw["algorithm"] == "pathtracing" or args.precision == "f32" else "f64", "data": "synthetic",
        "config": {"workload": desc, "variant": args.variant, "accel": args.accel, "rays_per_step": rays // max(1, args.steps),
                   "partition": {0: "none", 1: "spp", 2: "rows"}[params.part_mode], "l2": "256 MB buffer zeroed between timed steps",
                   "wall_s_timed_region": wall},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(scene_bytes + len(bytes(params))),
                "d2h_bytes_per_step": int(H * W * 3 * 4 + 64)},
        "gpu_launches": launches, "clocks": clocks, "roofline": roofline,
    }
    if world_size == 1 and not args.no_cpu_baseline:
        v, sample = cpu_baseline_sample(world, camera, kw, 1)
        line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": 1, "kind": "port", "sample": sample}
    emit(line)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c3", choices=["c3", "c4", "c5", "tonemap"])
    ap.add_argument("--variant", default="auto", choices=["auto", "mega", "warp"])
    ap.add_argument("--precision", default="auto", choices=["auto", "f32", "f64"])
    ap.add_argument("--partition", default="auto", choices=["auto", "spp", "rows"],
                    help="multi-GPU split: strata of every pixel (path tracing default) or interleaved rows")
    ap.add_argument("--accel", default="none", choices=["none", "bvh"],
                    help="bvh: sphere hierarchy instead of the reference's loop over all shapes (same image; separately reported mode)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--scale", type=int, default=1, help="divide width and height by this (quick experiments only; "
                    "a scaled run is NOT the benchmark and says so in config)")
    args = ap.parse_args()
    global SCALE
    SCALE = max(1, args.scale)
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world_size = int(os.environ.get("WORLD_SIZE", "1"))
    if args.workload == "tonemap":
        run_tonemap(args, rank, local_rank, world_size)
    elif args.impl == "reference":
        run_reference(args, rank, world_size)
    else:
        run_ours(args, rank, local_rank, world_size)
        if world_size > 1:
            import torch.distributed as dist

            if dist.is_initialized():
                dist.destroy_process_group()


if __name__ == "__main__":
    main()
