#!/usr/bin/env python3
"""bench.py -- the headline benchmark: Mrays/s and ms/frame on data/cornell_box.fray (400x400, Monte-Carlo path tracing at
256 paths per pixel, BASELINE.json configs[2]) on N B200s of one node, next to fray's own CPU renderer.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config c1|c2|c3|c4|c5]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one frame. `value` is whole-job throughput with the scene resident in HBM: rays of the frame (closest-hit +
any-hit queries, counted on the device) over the device time of the render kernels + the multi-GPU exchange + resolve, max over
ranks. `e2e` is the same frame through the public call with host buffers (fray_gpu_render: kernel parameters up, framebuffer
down into pinned host memory). One JSON line on stdout (rank 0).

--config selects another BASELINE.json configuration (c1 boxed, c2 zaphod, c3 cornell_box = the default and the headline,
c4 smallpt at 1024 paths/pixel, c5 forest at 3840x2160); the default run also times the other four briefly on rank 0 at N=1
and reports them under `other_configs`, so that one driver run records all five.
"""
from __future__ import annotations

import argparse
import json
import os
import re
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

UNIT = "Mrays/s"
# flop_per_ray: ALGORITHMIC work per ray of the workload under the cost model of SURVEY.md section 8(d) -- what the reference's
# algorithm does per ray (per-node transforms and box tests included), counted with the instrumented oracle
# (tests/oracle_util.py: algorithmic_work_per_ray; DESIGN.md "Measurement"). It is NOT the number of instructions these kernels
# execute: the flat tables remove most of that work, so `roofline.frac` reads "reference work per second against the FP32
# peak", and the executed-FP32 share of the pipe is reported next to it from the ncu capture of the same build (`roofline.ncu`).
CONFIGS = {
    "c1": dict(scene="boxed", settings=None, flop_per_ray=1000.0, integrator="Whitted, 1 sample/pixel, two 4x4 rectangular lights",
               baseline="configs[0]", kernel="waveTraceKernel / waveShadeKernel / waveShadowKernel (wavefront Whitted)"),
    "c2": dict(scene="zaphod", settings=None, flop_per_ray=270.0, integrator="Whitted, depth of field 100 samples/pixel",
               baseline="configs[1]", kernel="renderKernel<float, Whitted, flat table + textures + lens>"),
    "c3": dict(scene="cornell_box", settings=dict(pathsPerPixel=256), flop_per_ray=937.0, bytes_per_ray=1430.0, integrator="GI 256 paths/pixel",
               baseline="configs[2]", kernel="renderKernel<float, GI, FRAY_F_FLAT|FRAY_F_HEX>"),
    "c4": dict(scene="smallpt", settings=dict(pathsPerPixel=1024), flop_per_ray=650.0, integrator="GI 1024 paths/pixel",
               baseline="configs[3]", kernel="renderKernel<float, GI, flat table + spheres + two-sided list>",
               cpu_settings=dict(pathsPerPixel=64)),  # the in-run CPU leg renders 64 of the 1024 paths per pixel (~10 s instead of ~3 min)
    "c5": dict(scene="forest", settings=dict(interactive="off", frameWidth=3840, frameHeight=2160), flop_per_ray=1100.0, integrator="Whitted, 1 sample/pixel",
               baseline="configs[4]", kernel="waveTraceKernel / waveShadeKernel / waveShadowKernel (wavefront Whitted)",
               reference_settings=dict(interactive="off", frameWidth=1920, frameHeight=1080),
               reference_note="the reference's framebuffer stops at 3000 pixels a side (VFB_MAX_SIZE, src/constants.h:27): its arm renders 1920x1080"),
}
HEADLINE = "c3"
NOMINAL_FP32_TFLOPS = 148 * 128 * 2 * 1.965e9 / 1e12  # SMs x FP32 lanes x 2 (FMA) x max SM clock


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# stdout carries exactly one JSON line: anything libraries write to file descriptor 1 meanwhile (NCCL prints its version
# banner there) is sent to stderr, and emit() writes the line to the real stdout
_REAL_STDOUT = None


def capture_stdout():
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line: dict):
    data = (json.dumps(line) + "\n").encode()
    sys.stdout.flush()
    if _REAL_STDOUT is None:
        os.write(1, data)
    else:
        os.write(_REAL_STDOUT, data)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons while the timed region runs (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.rows = []
        self.proc = None
        self.gpu = gpu_index
        self.t = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "25", "-i", str(self.gpu)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
            return
        self.t = threading.Thread(target=self._read, daemon=True)
        self.t.start()

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def stop(self, t0: float, t1: float) -> dict:
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        rows = [r for (t, r) in self.rows if t0 - 0.05 <= t <= t1 + 0.15 and len(r) >= 9] or [r for (_, r) in self.rows if len(r) >= 9]
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm = sorted(float(r[1]) for r in rows)
        reasons = set()
        for r in rows:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(rows[0][2]), "power_w_max": max(float(r[3]) for r in rows),
                "samples": len(rows), "reasons": sorted(reasons)}


def scene_file(cfg: dict, extra: dict | None = None, settings_key: str = "settings") -> str:
    from fray_b200.scenes import override_scene
    st = dict(cfg.get(settings_key) or cfg.get("settings") or {})
    st.update(extra or {})
    tag = "bench_" + ("_".join(f"{k}{v}" for k, v in sorted(st.items())) or "default")
    return override_scene(cfg["scene"], tag, st or None)


def workload_name(cfg: dict, W: int, H: int) -> str:
    return f"{cfg['scene']}.fray {W}x{H}, {cfg['integrator']} (BASELINE.json {cfg['baseline']})"


def metric_name(cfg: dict, W: int, H: int) -> str:
    return f"Mrays/s ({cfg['scene']}.fray {W}x{H}, {cfg['integrator']})"


def reference_frame_seconds(ou, scene_path: str, kind: str, sc, cores: int) -> float:
    if kind == "reference":
        out = subprocess.run([ou.REF_BIN, scene_path], capture_output=True, text=True, check=True).stdout
        return float(re.search(r"Render took ([0-9.]+)s", out).group(1))
    t = time.time()
    ou.oracle_render(sc, threads=cores)
    return time.time() - t


def run_reference(args, rank: int, world: int):
    """--impl reference: the reference's own multithreaded CPU renderer (oracle/_ref/fray_ref, unmodified sources built
    headless) on this host's cores, on the SAME configuration as our arm: each step renders the whole frame. If the whole run
    would not end within ~5 minutes the remaining steps fall back to a bounded sample of the workload (fewer paths per pixel;
    the metric is a rate) and the line says so. Rays are counted by the oracle on the identical configuration."""
    if rank != 0:
        return
    import fray_b200 as fb
    import oracle_util as ou
    cfg = CONFIGS[args.config]
    cores = min(64, os.cpu_count() or 1)  # the reference's pool holds at most 64 threads (src/cxxptl-sdl.h:48)
    kind = "reference" if os.path.exists(ou.REF_BIN) else "port"
    extra = dict(numThreads=cores, wantPrepass="off")
    f = scene_file(cfg, extra, "reference_settings")
    sc = fb.Scene(f)
    _, ostats = ou.oracle_render(sc, threads=cores)  # the ray count is a property of (scene, seed): GPU fp64 == oracle one for one (tests)
    rays = ostats.rays
    sample = f"the whole frame: {cfg['scene']}.fray {sc.width}x{sc.height} at {sc.spp} samples/pixel ({rays} rays)"
    same = "reference_settings" not in cfg
    first = reference_frame_seconds(ou, f, kind, sc, cores)  # warm-up frame (also sizes the run)
    if first * (args.steps + 1) > 300.0 and cfg["settings"] and "pathsPerPixel" in cfg["settings"]:
        spp = max(8, int(cfg["settings"]["pathsPerPixel"] * 240.0 / (first * (args.steps + 1))))
        f = scene_file(cfg, dict(extra, pathsPerPixel=spp))
        sc = fb.Scene(f)
        _, ostats = ou.oracle_render(sc, threads=cores)
        rays = ostats.rays
        sample = f"bounded sample: {cfg['scene']}.fray {sc.width}x{sc.height} at {spp} of {cfg['settings']['pathsPerPixel']} paths/pixel ({rays} rays) per step"
        same = False
    times = [reference_frame_seconds(ou, f, kind, sc, cores) for _ in range(args.steps)]
    sec = sum(times) / len(times)
    value = rays / sec / 1e6
    W, H = (sc.width, sc.height)
    full = fb.Scene(scene_file(cfg))
    line = {
        "impl": "reference", "metric": metric_name(cfg, full.width, full.height), "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(cfg, full.width, full.height), "host_threads": cores, "step": sample, "same_config": same,
                   "note": cfg.get("reference_note", "one warm-up frame, then every step renders the frame again")},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample + f", {kind} renderer, mean of {args.steps} frames"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


def cpu_baseline(cfg: dict, gpu_frame, log_fn) -> tuple[dict, dict]:
    """The reference's own CPU renderer on the same frame (rank 0, N=1 only): one pass of the oracle port (ray count, work
    profile and the image our frame is compared with), then the reference binary, best of two frames. ~25 s for the headline."""
    import numpy as np

    import fray_b200 as fb
    import oracle_util as ou
    cores = min(64, os.cpu_count() or 1)
    f = scene_file(cfg, dict(numThreads=cores, wantPrepass="off"), "cpu_settings" if "cpu_settings" in cfg else "reference_settings")
    sc = fb.Scene(f)
    t = time.time()
    want, ostats = ou.oracle_render(sc, threads=cores, seed=42)
    port_sec = time.time() - t
    prof = ou.oracle_work_profile()
    flop, byts = ou.algorithmic_work_per_ray(prof, ostats.rays)
    if os.path.exists(ou.REF_BIN):
        best = None
        for _ in range(2):
            sec = reference_frame_seconds(ou, f, "reference", sc, cores)
            best = sec if best is None else min(best, sec)
        kind, sec = "reference", best
    else:
        kind, sec = "port", port_sec
    log_fn(f"cpu baseline ({kind}, {cores} threads): {sec:.2f} s for {ostats.rays} rays; oracle port {port_sec:.2f} s; "
           f"algorithmic work {flop:.0f} flop/ray {byts:.0f} B/ray")
    base = {"value": ostats.rays / sec / 1e6, "unit": UNIT, "cores": cores, "kind": kind,
            "sample": f"{cfg['scene']}.fray {sc.width}x{sc.height} at {sc.spp} samples/pixel, the full frame ({ostats.rays} rays by the reference's count), best of 2 frames",
            "ms_per_frame_sample": sec * 1e3, "flop_per_ray_measured": flop, "bytes_per_ray_measured": byts}
    parity = None
    if gpu_frame is not None and gpu_frame.shape == want.shape and "cpu_settings" not in cfg:
        frac, rmse, mx = ou.compare(want, gpu_frame, 1e-3)
        parity = {"against": "oracle/ FP64 restatement of the reference (bit-exact with fray_ref_ctr on the goldens), same seed 42, same frame",
                  "rmse": rmse, "frac_within_1e-3": frac, "max_abs_diff": mx,
                  "mean_gpu": float(np.mean(gpu_frame)), "mean_oracle": float(np.mean(want))}
    return base, parity


def quick_config(fb, name: str, device: int, peak: float) -> dict:
    """One BASELINE configuration timed briefly on one GPU (device-resident, best of 5 frames after 2 warm-up frames)."""
    cfg = CONFIGS[name]
    sc = fb.Scene(scene_file(cfg))
    ctx = fb.GpuContext(sc, device, fb.FP32)
    best = None
    for it in range(7):
        _, st = ctx.render(seed=42)
        if it >= 2 and (best is None or st.device_ms < best.device_ms):
            best = st
    ctx.close()
    tf = best.rays * cfg["flop_per_ray"] / (best.device_ms * 1e-3) / 1e12
    return {"config": name, "workload": workload_name(cfg, sc.width, sc.height), "ms_per_frame": best.device_ms, "rays_per_frame": best.rays,
            "mrays_s": best.rays / best.device_ms / 1e3, "kernel_launches": best.kernel_launches, "algorithmic_flop_per_ray": cfg["flop_per_ray"],
            "roofline_frac_algorithmic": tf / peak, "kernel": cfg["kernel"], "timing": "kernel time (CUDA events), best of 5 frames, one GPU"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default=HEADLINE, choices=sorted(CONFIGS), help="BASELINE.json configuration (c3 = cornell_box 256 paths/pixel, the headline)")
    ap.add_argument("--spp", type=int, default=0, help="override the paths per pixel of a path-traced configuration")
    ap.add_argument("--precision", default="fp32", choices=["fp32", "fp64"])
    ap.add_argument("--split", default="auto", choices=["tiles", "p2p", "samples", "host", "auto"],
                    help="multi-GPU decomposition: tiles (BASELINE.json for cornell; NCCL reduce), p2p (tiles written straight into rank 0's frame), "
                         "samples, host (end to end: every GPU stores its tiles straight into one shared page-locked host frame; device-timed: "
                         "samples or tiles by sample count). auto = samples or tiles by sample count, NCCL reduce (measured end to end on 2 / 4 / 8 "
                         "B200: 4.29 / 2.20 / 1.20 ms against 4.19 / 2.34 / 1.17-1.19 ms with host)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-other-configs", action="store_true")
    args = ap.parse_args()
    capture_stdout()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import numpy as np
    import torch
    import torch.distributed as dist

    import fray_b200 as fb
    import fray_b200.dist as fdist

    if not torch.cuda.is_available():
        sys.exit("bench.py: no CUDA device; the fray_b200 render path is CUDA only")
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local_rank}"))
    else:
        torch.cuda.set_device(local_rank)
    dev = torch.device(f"cuda:{local_rank}")

    cfg = dict(CONFIGS[args.config])
    if args.spp:
        cfg["settings"] = dict(cfg.get("settings") or {}, pathsPerPixel=args.spp)
        cfg["integrator"] = re.sub(r"\d+ paths/pixel", f"{args.spp} paths/pixel", cfg["integrator"])
    precision = fb.FP32 if args.precision == "fp32" else fb.FP64
    scene = fb.Scene(scene_file(cfg))
    W, H, spp = scene.width, scene.height, scene.spp
    r = fdist.DistributedRenderer(scene, mode=args.split, precision=precision, device=local_rank)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # ---- device-resident throughput ----
    for _ in range(max(args.warmup, 3)):
        r.render_device()
        r.stats()
    torch.cuda.synchronize()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    starts = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    stops = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    kernel_ms = []
    rays_step = 0
    if sampler:
        sampler.start()
        time.sleep(0.25)
    # the GPU idled while the sampler started (rank 0) or while waiting for rank 0 (others): one more untimed frame brings
    # clocks and the NCCL channels back up, then all ranks enter the timed region together
    barrier()
    r.render_device()
    barrier()
    t0 = time.time()
    for i in range(args.steps):
        flush.fill_(i & 0xFF)  # evict L2 between timed frames (not timed)
        starts[i].record()
        r.render_device()
        stops[i].record()
        st = r.stats()  # syncs; per-step kernel time + ray counters of this rank
        kernel_ms.append(st.device_ms)
        rays_step = st.rays
        launches_step = st.kernel_launches + (1 if (rank == 0 and r.device_mode != "p2p") else 0)  # this rank's render kernels, and resolve on rank 0
    barrier()
    t1 = time.time()
    clocks = sampler.stop(t0, t1) if sampler else None
    step_ms = [a.elapsed_time(b) for a, b in zip(starts, stops)]
    if os.environ.get("FRAY_BENCH_DEBUG"):
        log(f"rank {rank}: step_ms {[round(x, 3) for x in step_ms]} kernel_ms {[round(x, 3) for x in kernel_ms]} wall {t1 - t0:.3f}s")
    tot = torch.tensor([sum(step_ms), float(rays_step), sum(kernel_ms)], dtype=torch.float64, device=dev)
    if world > 1:
        mx = tot.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = tot.clone()
        dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        total_ms, rays_frame, kernel_total_ms = float(mx[0]), float(sm[1]), float(mx[2])
    else:
        total_ms, rays_frame, kernel_total_ms = float(tot[0]), float(tot[1]), float(tot[2])
    ms_per_step = total_ms / args.steps
    value = rays_frame / (ms_per_step * 1e-3) / 1e6

    # ---- end to end through the public API with host buffers ----
    host = torch.empty((H, W, 3), dtype=torch.float32).pin_memory()
    host_np = host.numpy()
    cam = scene.head.camera
    frame_np = None

    def e2e_step():
        nonlocal frame_np
        if world == 1:
            r.ctx.update_camera(cam)          # this frame's input (kernel parameters travel host -> device with the launch)
            r.ctx.render(out=host_np, spp=spp)  # kernels + device -> pinned host copy of the frame
            frame_np = host_np
        else:
            r.ctx.update_camera(cam)
            frame_np = r.render()
    for _ in range(2):
        e2e_step()
    barrier()
    te0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
    barrier()
    e2e_ms = (time.perf_counter() - te0) * 1e3 / args.steps
    if os.environ.get("FRAY_DIST_DEBUG") and getattr(r, "dbg", None):
        import numpy as _np
        log("host-mode phases (us): wait-entry, launch, sync, wait-peers =", _np.round(_np.mean(_np.array(r.dbg[-args.steps:]), axis=0), 1), "kernel ms", r.stats().device_ms)
    e2e_t = torch.tensor([e2e_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
    e2e_ms = float(e2e_t[0])
    e2e_value = rays_frame / (e2e_ms * 1e-3) / 1e6

    # ---- multi-GPU: rank 0's assembled frame against the same frame rendered by ONE GPU ----
    multi_check = None
    if world > 1 and rank == 0:
        single, _ = r.ctx.render(seed=42, spp=spp)
        d = np.abs(single.astype(np.float64) - frame_np.astype(np.float64))
        multi_check = {"against": "the same frame rendered by rank 0's GPU alone (same seed)", "max_abs_diff": float(d.max()),
                       "bit_identical": bool(np.array_equal(single, frame_np)), "frac_within_1e-5": float((d.max(axis=-1) <= 1e-5).mean()),
                       "note": "the frame of the end-to-end path; a share cuts a pixel's samples into other chunks than one GPU does (FP32 regrouping, ~1e-5); "
                               "fray_gpu_multi_render's tile split keeps the single GPU's chunks and is bit-identical (single_process below)"}

    # ---- multi-GPU, ONE process: fray_gpu_multi_render on the same N GPUs (the C++ drop-in's own path), rank 0 only ----
    single_process = None
    if world > 1:
        store = dist.distributed_c10d._get_default_store()
        if rank == 0:
            try:
                multi = fb.MultiGpuContext(scene, list(range(world)), precision)
                alone, _ = r.ctx.render(seed=42, spp=spp)
                res = {}
                for split_name, split_id in (("tiles", fb.SPLIT_TILES), ("samples", fb.SPLIT_SAMPLES)):
                    for _ in range(3):
                        img, mst = multi.render(out=host_np, seed=42, spp=spp, split=split_id)
                    t_sp = time.perf_counter()
                    for _ in range(args.steps):
                        img, mst = multi.render(out=host_np, seed=42, spp=spp, split=split_id)
                    ms_sp = (time.perf_counter() - t_sp) * 1e3 / args.steps
                    d = np.abs(alone.astype(np.float64) - img.astype(np.float64))
                    res[split_name] = {"e2e_ms_per_frame": ms_sp, "e2e_mrays_s": mst.rays / (ms_sp * 1e-3) / 1e6, "slowest_share_kernel_ms": mst.device_ms,
                                       "bit_identical_to_one_gpu": bool(np.array_equal(alone, img)), "max_abs_diff": float(d.max())}
                multi.close()
                single_process = dict(res, note="fray_gpu_multi_render: one process, a context and a host thread per GPU, shares stored into GPU 0's "
                                                 "frame over NVLink peer access, host buffer in and out (the path of `fray --gpu --devices N`)")
            except fb.FrayError as e:
                single_process = {"error": str(e)}
            store.set("fray_single_process_done", "1")
        else:
            store.wait(["fray_single_process_done"])  # on the CPU: an NCCL barrier would spin on this rank's GPU while rank 0 measures

    if rank == 0:
        fp32_peak, l2_peak = fb.measure_peaks(local_rank, 20.0)
        kernel_ms_per_launch = kernel_total_ms / args.steps
        flop_per_ray = cfg["flop_per_ray"]
        achieved_tflops = rays_step * flop_per_ray / (kernel_ms_per_launch * 1e-3) / 1e12
        roofline = {
            "bound": "fp32", "achieved": achieved_tflops, "peak": fp32_peak, "unit": "TFLOP/s", "frac": achieved_tflops / fp32_peak,
            "frac_is": "ALGORITHMIC: rays/s x the reference algorithm's flop per ray (SURVEY.md 8d cost model), i.e. reference work per second "
                       "against the FP32 peak -- not the share of the pipe these kernels keep busy (see `ncu` for that)",
            "traffic": None,  # filled below from the ncu capture in profiles/ when there is one for this configuration
            "peak_source": "FFMA micro-benchmark run in this process (fray_gpu_measure_peaks); MEASURED_PEAKS.json holds no FP32 figure",
            "peak_nominal": NOMINAL_FP32_TFLOPS, "frac_of_nominal": achieved_tflops / NOMINAL_FP32_TFLOPS,
            "kernel": cfg["kernel"] if precision == fb.FP32 else "renderKernel<double, ...> (parity precision)",
            "kernel_ms_per_launch": kernel_ms_per_launch, "algorithmic_flop_per_ray": flop_per_ray,
        }
        if "bytes_per_ray" in cfg:
            roofline["algorithmic_bytes_per_ray"] = cfg["bytes_per_ray"]
            roofline["table_reads"] = {"algorithmic_gbs": rays_step * cfg["bytes_per_ray"] / (kernel_ms_per_launch * 1e-3) / 1e9, "l2_peak_gbs": l2_peak,
                                       "note": "the reference's per-ray table reads (node, box, triangle records); this kernel stages them once per CTA in "
                                               "shared memory, so they are served by LDS.128 broadcasts, not by L2"}
        ncu_file = os.path.join(ROOT, "profiles", f"ncu_{args.config}.json")  # written by tools/ncu_summary.py --json from the capture of this build
        if os.path.exists(ncu_file) and precision == fb.FP32:
            with open(ncu_file) as fp:
                roofline["ncu"] = json.load(fp)
            if world == 1:  # dram__bytes_read.sum + dram__bytes_write.sum of one launch in that capture (profiles/, same build)
                roofline["traffic"] = int((roofline["ncu"]["dram_read_mb"] + roofline["ncu"]["dram_write_mb"]) * 1e6)
                roofline["traffic_source"] = f"profiles/{os.path.basename(ncu_file)}: ncu --set full capture of this kernel on this workload"
        line = {
            "metric": metric_name(cfg, W, H), "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32" if precision == fb.FP32 else "f64", "data": "synthetic",
            "config": {"workload": workload_name(cfg, W, H), "rays_per_frame": int(rays_frame),
                       "split": (r.device_mode if world > 1 else "none"),
                       "e2e_split": ("tiles, every GPU storing its own straight into one shared page-locked host frame (fray_gpu_render_to_host)"
                                     if r.mode == "host" else (r.mode if world > 1 else "none")), "l2": "flushed between timed frames (256 MB write)", "seed": 42},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": e2e_ms, "h2d_bytes_per_step": 1024, "d2h_bytes_per_step": W * H * 12},
            "gpu_launches": args.steps * launches_step,
            "roofline": roofline,
        }
        if multi_check:
            line["multi_gpu_check"] = multi_check
        if single_process:
            line["single_process"] = single_process
        if world == 1 and not args.no_other_configs and args.config == HEADLINE and precision == fb.FP32:
            line["other_configs"] = [quick_config(fb, name, local_rank, fp32_peak) for name in sorted(CONFIGS) if name != HEADLINE]
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"], parity = cpu_baseline(cfg, np.array(frame_np) if (args.config != "c5") else None, log)
            if parity:
                line["parity"] = parity
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
