#!/usr/bin/env python3
"""A second libfray_gpu.so with extra preprocessor flags for the wavefront / FP32 units, for A/B timing on the GPU box:

    python tools/build_variant.py kd8 -DFRAY_KD_SHORT=8         ->  fray_b200/_build/variants/libfray_gpu_kd8.so
    FRAY_GPU_LIB=fray_b200/_build/variants/libfray_gpu_kd8.so python tools/time_configs.py --fp32
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fray_b200 import build as b


def main():
    name, flags = sys.argv[1], sys.argv[2:]
    b.build_gpu()
    out_dir = os.path.join(b.BUILD, "variants")
    os.makedirs(out_dir, exist_ok=True)
    fast = ["-prec-div=false", "-prec-sqrt=false", "-ftz=true"]
    objs = []
    procs = []
    for unit in ("fray_gpu", "render_fp32", "render_wave"):
        obj = os.path.join(out_dir, f"{unit}_{name}.o")
        extra = fast if unit != "fray_gpu" else []
        procs.append(subprocess.Popen([b.NVCC] + b.NVCC_COMMON + extra + flags + ["-c", os.path.join(b.CSRC, unit + ".cu"), "-o", obj]))
        objs.append(obj)
    for p in procs:
        if p.wait() != 0:
            raise SystemExit("variant build failed")
    objs.append(os.path.join(b.BUILD, "render_fp64.o"))
    out = os.path.join(out_dir, f"libfray_gpu_{name}.so")
    subprocess.check_call([b.NVCC] + b.ARCH + ["-shared", "-o", out] + objs + ["-cudart", "static"])
    print(out)


if __name__ == "__main__":
    main()
