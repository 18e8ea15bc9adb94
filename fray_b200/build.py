"""Build recipes for the native parts of fray_b200 (explicit compiler invocations, everything in-tree).

  libfray_host.so   g++   fray_b200/host/*.cpp                       the CPU-side scene layer
  libfray_gpu.so    nvcc  fray_b200/csrc/{fray_gpu,render_fp32,render_wave,render_fp64}.cu   sm_100a only
  fray              g++   fray_b200/host/main.cpp                    the `fray [--gpu] scene.fray` command line tool
nvcc cross-compiles without a GPU, so this runs in the build container and the .so files travel to the GPU box.
"""
from __future__ import annotations

import concurrent.futures as cf
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
INCLUDE = os.path.join(ROOT, "include")
CSRC = os.path.join(HERE, "csrc")
HOST = os.path.join(HERE, "host")
BUILD = os.path.join(HERE, "_build")

NVCC = os.environ.get("NVCC") or shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
CXX = os.environ.get("CXX") or "g++"
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
NVCC_COMMON = ARCH + ["-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC", "-I", INCLUDE, "--expt-relaxed-constexpr"]

HOST_SRCS = ["parser.cpp", "elements.cpp", "mesh.cpp", "image.cpp", "exr.cpp", "flatten.cpp", "capi.cpp"]
HOST_FLAGS = ["-O2", "-ffp-contract=off", "-std=c++17", "-fPIC", "-Wall", "-Wno-format-security", "-I", INCLUDE]


def _newer(target: str, deps: list[str]) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def _run(cmd: list[str], log: bool = True):
    if log:
        print("+", " ".join(cmd), flush=True)
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError(f"build step failed: {' '.join(cmd)}")
    return r.stdout + r.stderr


def build_host(force: bool = False) -> str:
    out = os.path.join(HERE, "libfray_host.so")
    srcs = [os.path.join(HOST, s) for s in HOST_SRCS]
    deps = srcs + [os.path.join(HOST, h) for h in ("scene.h", "vec.h")] + [os.path.join(INCLUDE, h) for h in ("fray_gpu.h", "fray_host.h")] + [os.path.join(CSRC, "rng.cuh")]
    if force or _newer(out, deps):
        _run([CXX] + HOST_FLAGS + ["-shared"] + srcs + ["-o", out, "-lz"])
    return out


def build_cli(force: bool = False) -> str:
    out = os.path.join(HERE, "fray")
    src = os.path.join(HOST, "main.cpp")
    if not os.path.exists(src):
        return ""
    if force or _newer(out, [src, os.path.join(HERE, "libfray_host.so")]):
        _run([CXX] + HOST_FLAGS + [src, "-o", out, "-L", HERE, "-lfray_host", "-ldl", "-Wl,-rpath,$ORIGIN"])
    return out


def build_gpu(force: bool = False, verbose_ptxas: bool = False) -> str:
    out = os.path.join(HERE, "libfray_gpu.so")
    os.makedirs(BUILD, exist_ok=True)
    headers = [os.path.join(CSRC, h) for h in ("core.cuh", "flat.cuh", "rng.cuh", "render_kernels.cuh", "scene_image.h", "wave.cuh", "wave_kernels.cuh")] + [os.path.join(INCLUDE, "fray_gpu.h")]
    fast = ["-prec-div=false", "-prec-sqrt=false", "-ftz=true"]
    units = [("fray_gpu.cu", []), ("render_fp32.cu", fast), ("render_wave.cu", fast), ("render_fp64.cu", ["-fmad=false"])]
    extra = ["-Xptxas", "-v"] if verbose_ptxas else []
    jobs = []
    for src, flags in units:
        obj = os.path.join(BUILD, src.replace(".cu", ".o"))
        if force or verbose_ptxas or _newer(obj, [os.path.join(CSRC, src)] + headers):
            jobs.append([NVCC] + NVCC_COMMON + flags + extra + ["-c", os.path.join(CSRC, src), "-o", obj])
    logs = []
    if jobs:
        with cf.ThreadPoolExecutor(max_workers=len(jobs)) as ex:
            logs = list(ex.map(_run, jobs))
    objs = [os.path.join(BUILD, s.replace(".cu", ".o")) for s, _ in units]
    if jobs or force or _newer(out, objs):
        _run([NVCC] + ARCH + ["-shared", "-o", out] + objs + ["-cudart", "static"])
    if verbose_ptxas:
        print("\n".join(logs))
    return out


def build_all(force: bool = False):
    build_host(force)
    build_cli(force)
    build_gpu(force)


if __name__ == "__main__":
    import argparse
    ap = argparse.ArgumentParser()
    ap.add_argument("--force", action="store_true")
    ap.add_argument("--ptxas", action="store_true", help="print registers / spills of every kernel")
    a = ap.parse_args()
    build_host(a.force)
    build_cli(a.force)
    build_gpu(a.force, a.ptxas)
