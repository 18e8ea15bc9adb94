#!/usr/bin/env python3
"""Render a scene on cuda:0 in both precisions and save the images and primary-hit AOVs as .npz (to look at GPU results in
the build container, which has no GPU):

    python tools/dump_frame.py local:csg_layered gpurun_out/csg.npz        (tests/scenes/csg_layered.fray)
    python tools/dump_frame.py hw9/dragon gpurun_out/dragon.npz frameWidth=96 frameHeight=64
"""
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np

import fray_b200 as fb
from fray_b200 import scenes


def main():
    name, out = sys.argv[1], sys.argv[2]
    settings = dict(a.split("=", 1) for a in sys.argv[3:])
    if name.startswith("local:"):
        path = os.path.join(scenes.DATA_DIR, name[6:] + "__test.fray")
        shutil.copyfile(os.path.join(ROOT, "tests", "scenes", name[6:] + ".fray"), path)
    else:
        path = scenes.override_scene(name, "dump", settings or None)
    sc = fb.Scene(path)
    res = {}
    for prec, tag in ((fb.FP32, "f32"), (fb.FP64, "f64")):
        ctx = fb.GpuContext(sc, 0, prec)
        res["rgb_" + tag], st = ctx.render(seed=42)
        res["aov_" + tag], _ = ctx.render(mode=fb.RENDER_AOV)
        print(f"{name} {tag}: {st.device_ms:.3f} ms, {st.rays} rays")
        ctx.close()
    np.savez_compressed(out, **res)


if __name__ == "__main__":
    main()
