#!/usr/bin/env python3
"""Kernel timings of the benchmark configurations (and the two heavier bundled scenes) on cuda:0, both precisions, best of
three frames. Prints one line per case; `profiles/r01_*_all_configs_timing.log` are its FP32 lines. (Parity lives in tests/.)"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import fray_b200 as fb
from fray_b200 import scenes

def main():
    timing = [
        ("cornell_box", dict(pathsPerPixel=256), None),
        ("cornell_box", dict(pathsPerPixel=40), None),
        ("smallpt", dict(pathsPerPixel=256), None),
        ("boxed", None, None),
        ("zaphod", None, None),
        ("forest", dict(interactive="off", frameWidth=1920, frameHeight=1080), None),
        ("forest", dict(interactive="off", frameWidth=3840, frameHeight=2160), None),
        ("forest", dict(interactive="off", frameWidth=3840, frameHeight=2160, wantAA="on"), None),
        ("hw9/dragon", None, None),
    ]
    for name, st, cam in timing:
        f = scenes.override_scene(name, "time", st, cam)
        sc = fb.Scene(f)
        for prec, pn in ((fb.FP32, "f32"), (fb.FP64, "f64")):
            if "--fp32" in sys.argv and prec != fb.FP32:
                continue
            ctx = fb.GpuContext(sc, 0, prec)
            best = None
            for it in range(5):
                img, s = ctx.render()
                best = s if best is None or s.device_ms < best.device_ms else best
            print(f"TIME {name:14s} {pn} {sc.width}x{sc.height} spp {sc.spp}: {best.device_ms:9.2f} ms  rays {best.rays} "
                  f"-> {best.rays / best.device_ms / 1e3:9.1f} Mrays/s  mean {img.mean():.5f}", flush=True)
            ctx.close()


if __name__ == "__main__":
    main()
