#!/usr/bin/env python3
"""Render one bundled scene a few times on cuda:0 and print the device time (a short command to put under ncu).

    python tools/render_once.py hw9/dragon [--frames 3] [--fp64] [key=value ...]      e.g. pathsPerPixel=256 frameWidth=1920
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import fray_b200 as fb
from fray_b200 import scenes


def main():
    argv = sys.argv[1:]
    frames = 3
    if "--frames" in argv:
        i = argv.index("--frames")
        frames = int(argv[i + 1])
        del argv[i:i + 2]
    args = [a for a in argv if not a.startswith("--")]
    name = args[0]
    settings = dict(a.split("=", 1) for a in args[1:])
    precision = fb.FP64 if "--fp64" in sys.argv else fb.FP32
    if name == "forest":
        settings.setdefault("interactive", "off")
    sc = fb.Scene(scenes.override_scene(name, "once", settings or None))
    ctx = fb.GpuContext(sc, 0, precision)
    for _ in range(frames):
        img, s = ctx.render()
        print(f"{name} {sc.width}x{sc.height} spp {sc.spp}: {s.device_ms:.3f} ms, {s.rays} rays, {s.rays / s.device_ms / 1e3:.1f} Mrays/s, mean {img.mean():.5f}", flush=True)
    ctx.close()


if __name__ == "__main__":
    main()
