#!/usr/bin/env python3
"""bench.py -- the headline benchmark: Mrays/s and ms/frame on data/cornell_box.fray (400x400, Monte-Carlo path tracing at
256 paths per pixel, BASELINE.json configs[2]) on N B200s of one node, next to fray's own CPU renderer.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one frame. `value` is whole-job throughput with the scene resident in HBM: rays of the frame (closest-hit +
any-hit queries, counted on the device) over the device time of render kernel + NCCL reduce + resolve, max over ranks.
`e2e` is the same frame through the public call with host buffers (fray_gpu_render: kernel parameters up, framebuffer
down into pinned host memory). One JSON line on stdout (rank 0).
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

SCENE = "cornell_box"
SPP = 256
METRIC = "Mrays/s (cornell_box.fray 400x400, GI 256 paths/pixel)"
UNIT = "Mrays/s"
# algorithmic work per ray of this workload under the cost model of SURVEY.md section 8(d), derived with the instrumented
# oracle (tools/work_profile.py, DESIGN.md "Measurement"): 6.97 node tests, 6.97 root-box tests, 12.8 cull tests,
# 6.45 triangle tests, 0.54 light tests, 0.46 light samples, 0.86 hemisphere samples, 0.36 BRDF evals per ray
FLOP_PER_RAY = 937.0
BYTES_PER_RAY = 1430.0
NCU_DRAM_BYTES_PER_LAUNCH = 191744 + 12758272
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


def bench_scene_file(spp: int, extra: dict | None = None) -> str:
    from fray_b200.scenes import override_scene
    st = dict(pathsPerPixel=spp)
    st.update(extra or {})
    tag = "bench_" + "_".join(f"{k}{v}" for k, v in sorted(st.items()))
    return override_scene(SCENE, tag, st)


def run_reference(args, rank: int, world: int):
    """--impl reference: the reference's own multithreaded CPU renderer (oracle/_ref/fray_ref, unmodified sources built
    headless) on this host's cores. Each step renders a bounded sample of the workload: the same scene at 40 paths/pixel
    (the scene file's default) instead of 256; rays are counted by the oracle on the identical configuration."""
    if rank != 0:
        return
    import fray_b200 as fb
    import oracle_util as ou
    cores = min(64, os.cpu_count() or 1)  # the reference's pool holds at most 64 threads (src/cxxptl-sdl.h:48)
    sample_spp = 40
    f = bench_scene_file(sample_spp, dict(numThreads=cores, wantPrepass="off"))
    kind = "reference" if os.path.exists(ou.REF_BIN) else "port"
    sc = fb.Scene(f)
    # rays of the sample frame (the ray count is a property of (scene, seed), see tests: GPU fp64 == oracle one for one)
    _, ostats = ou.oracle_render(sc)
    rays = ostats.rays

    def one_frame() -> float:
        if kind == "reference":
            out = subprocess.run([ou.REF_BIN, f], capture_output=True, text=True, check=True).stdout
            return float(re.search(r"Render took ([0-9.]+)s", out).group(1))
        t = time.time()
        ou.oracle_render(sc, threads=cores)
        return time.time() - t

    for _ in range(args.warmup if args.warmup < 2 else 1):
        one_frame()
    times = [one_frame() for _ in range(args.steps)]
    sec = sum(times) / len(times)
    value = rays / sec / 1e6
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "cornell_box.fray 400x400 GI, bounded sample: 40 paths/pixel per step (full workload 256)", "host_threads": cores},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind,
                         "sample": f"cornell_box.fray 400x400 at 40 paths/pixel ({rays} rays), {kind} renderer, mean of {args.steps} frames"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


def cpu_baseline(log_fn) -> dict:
    """The reference's own CPU renderer on the same frame (rank 0, N=1 only): ~25 s of CPU work (one pass of the oracle port for
    the ray count and the work profile, two of the reference binary)."""
    import fray_b200 as fb
    import oracle_util as ou
    cores = min(64, os.cpu_count() or 1)
    sample_spp = SPP  # the whole 256-path frame: ~8 s per frame on 16 threads, ~25 s for this leg
    f = bench_scene_file(sample_spp, dict(numThreads=cores, wantPrepass="off"))
    sc = fb.Scene(f)
    t = time.time()
    _, ostats = ou.oracle_render(sc, threads=cores)
    port_sec = time.time() - t
    prof = ou.oracle_work_profile()
    flop, byts = ou.algorithmic_work_per_ray(prof, ostats.rays)
    if os.path.exists(ou.REF_BIN):
        best = None
        for _ in range(2):
            out = subprocess.run([ou.REF_BIN, f], capture_output=True, text=True, check=True).stdout
            sec = float(re.search(r"Render took ([0-9.]+)s", out).group(1))
            best = sec if best is None else min(best, sec)
        kind, sec = "reference", best
    else:
        kind, sec = "port", port_sec
    log_fn(f"cpu baseline ({kind}, {cores} threads): {sec:.2f} s for {ostats.rays} rays; oracle port {port_sec:.2f} s; "
           f"algorithmic work {flop:.0f} flop/ray {byts:.0f} B/ray")
    return {"value": ostats.rays / sec / 1e6, "unit": UNIT, "cores": cores, "kind": kind,
            "sample": f"cornell_box.fray 400x400 at {sample_spp} paths/pixel, the full frame ({ostats.rays} rays by the reference's count), best of 2 frames",
            "ms_per_frame_sample": sec * 1e3, "flop_per_ray_measured": flop, "bytes_per_ray_measured": byts}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--spp", type=int, default=SPP, help="paths per pixel (the headline config is 256)")
    ap.add_argument("--precision", default="fp32", choices=["fp32", "fp64"])
    ap.add_argument("--split", default="auto", choices=["tiles", "p2p", "samples", "auto"],
                    help="multi-GPU decomposition: tiles (BASELINE.json for cornell; NCCL reduce), p2p (tiles written straight into rank 0's frame), samples")
    ap.add_argument("--no-cpu-baseline", action="store_true")
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

    precision = fb.FP32 if args.precision == "fp32" else fb.FP64
    scene = fb.Scene(bench_scene_file(args.spp))
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
        launches_step = st.kernel_launches + (1 if (rank == 0 and r.mode != "p2p") else 0)  # render (+ combine), and resolve on rank 0
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

    def e2e_step():
        if world == 1:
            r.ctx.update_camera(cam)          # this frame's input (kernel parameters travel host -> device with the launch)
            r.ctx.render(out=host_np, spp=spp)  # kernels + device -> pinned host copy of the frame
        else:
            r.ctx.update_camera(cam)
            r.render()
    for _ in range(2):
        e2e_step()
    barrier()
    te0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
    barrier()
    e2e_ms = (time.perf_counter() - te0) * 1e3 / args.steps
    e2e_t = torch.tensor([e2e_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
    e2e_ms = float(e2e_t[0])
    e2e_value = rays_frame / (e2e_ms * 1e-3) / 1e6

    if rank == 0:
        fp32_peak, l2_peak = fb.measure_peaks(local_rank, 20.0)
        kernel_ms_per_launch = kernel_total_ms / args.steps
        achieved_tflops = rays_step * FLOP_PER_RAY / (kernel_ms_per_launch * 1e-3) / 1e12
        roofline = {
            "bound": "fp32", "achieved": achieved_tflops, "peak": fp32_peak, "unit": "TFLOP/s", "frac": achieved_tflops / fp32_peak,
            # dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of this kernel on this workload, from the ncu --set full
            # capture profiles/r01_v34_cornell256_ncu_summary.txt (0.19 MB read + 12.8 MB written: the part of the 61 MB chunk-sum scratch
            # that left the L2 during the cold, serialised ncu pass; the scene tables never leave the SMs)
            "traffic": NCU_DRAM_BYTES_PER_LAUNCH if (args.spp == SPP and world == 1 and precision == fb.FP32) else None,
            "peak_source": "FFMA micro-benchmark run in this process (fray_gpu_measure_peaks); MEASURED_PEAKS.json holds no FP32 figure",
            "peak_nominal": NOMINAL_FP32_TFLOPS, "frac_of_nominal": achieved_tflops / NOMINAL_FP32_TFLOPS,
            "kernel": "renderKernel<float, GI, FRAY_F_FLAT|FRAY_F_HEX>" if precision == fb.FP32 else "renderKernel<double, GI, FRAY_F_GENERIC>",
            "kernel_ms_per_launch": kernel_ms_per_launch, "algorithmic_flop_per_ray": FLOP_PER_RAY, "algorithmic_bytes_per_ray": BYTES_PER_RAY,
            "table_reads": {"algorithmic_gbs": rays_step * BYTES_PER_RAY / (kernel_ms_per_launch * 1e-3) / 1e9, "l2_peak_gbs": l2_peak,
                            "note": "the reference's per-ray table reads (node, box, triangle records); this kernel stages them once per CTA in shared "
                                    "memory, so they are served by LDS.128 broadcasts, not by L2 (lts throughput < 2 % in the ncu capture)"},
            "hbm_note": "scene (<1 MB) and framebuffer (1.9 MB) are L1/L2 resident; HBM traffic is negligible",
        }
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32" if precision == fb.FP32 else "f64", "data": "synthetic",
            "config": {"workload": f"cornell_box.fray {W}x{H}, GI {spp} paths/pixel (BASELINE.json configs[2])", "rays_per_frame": int(rays_frame),
                       "split": r.mode if world > 1 else "none", "l2": "flushed between timed frames (256 MB write)", "seed": 42},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": e2e_ms, "h2d_bytes_per_step": 1024, "d2h_bytes_per_step": W * H * 12},
            "gpu_launches": args.steps * launches_step,  # renderKernel + combineKernel + resolveKernel per frame (rank 0)
            "roofline": roofline,
        }
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(log)
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
