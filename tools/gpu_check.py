#!/usr/bin/env python3
"""Developer check on a GPU box: parity of both device precisions against the oracle on every bundled scene (small
frames), then kernel timings of the benchmark configurations. Prints one line per case."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np

import fray_b200 as fb
import oracle_util as ou

CASES = [
    ("cornell_box", dict(frameWidth=100, frameHeight=100, pathsPerPixel=8), None),
    ("smallpt", dict(frameWidth=96, frameHeight=72, pathsPerPixel=8), None),
    ("boxed", dict(frameWidth=128, frameHeight=96), None),
    ("zaphod", dict(frameWidth=129, frameHeight=86), dict(numSamples=6)),
    ("forest", dict(frameWidth=160, frameHeight=120, interactive="off"), None),
    ("hw9/axe_test", dict(frameWidth=160, frameHeight=120), None),
    ("hw9/nonconvex", dict(frameWidth=160, frameHeight=120), None),
    ("hw10/bokeh", dict(frameWidth=96, frameHeight=72), dict(numSamples=4)),
    ("hw12/sphtri", dict(frameWidth=96, frameHeight=72, pathsPerPixel=8), None),
    ("hw9/dragon", dict(frameWidth=96, frameHeight=64), None),
]


def main():
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    if which in ("all", "parity"):
        for name, st, cam in CASES:
            f = ou.override_scene(name, "chk", st, cam)
            sc = fb.Scene(f)
            ref, rs = ou.oracle_render(sc)
            raov, _ = ou.oracle_render(sc, mode=fb.RENDER_AOV)
            for prec, pn in ((fb.FP64, "f64"), (fb.FP32, "f32")):
                ctx = fb.GpuContext(sc, 0, prec)
                img, s = ctx.render()
                aov, _ = ctx.render(mode=fb.RENDER_AOV)
                f3, rmse, mx = ou.compare(ref, img, 1e-3)
                f5, _, _ = ou.compare(ref, img, 1e-5)
                node_eq = float((aov[..., 0] == raov[..., 0]).mean())
                tri_eq = float((aov[..., 1] == raov[..., 1]).mean())
                print(f"PARITY {name:14s} {pn} frac<=1e-3 {f3:.5f} frac<=1e-5 {f5:.5f} rmse {rmse:.3g} max {mx:.3g} "
                      f"nodeEq {node_eq:.5f} triEq {tri_eq:.5f} rays {s.rays} (oracle {rs.rays}) {s.device_ms:.2f} ms", flush=True)
                ctx.close()
    if which in ("all", "time"):
        timing = [
            ("cornell_box", dict(pathsPerPixel=256), None),
            ("cornell_box", dict(pathsPerPixel=40), None),
            ("smallpt", dict(pathsPerPixel=256), None),
            ("boxed", None, None),
            ("zaphod", None, None),
            ("forest", dict(interactive="off", frameWidth=1920, frameHeight=1080), None),
            ("hw9/dragon", None, None),
        ]
        for name, st, cam in timing:
            f = ou.override_scene(name, "time", st, cam)
            sc = fb.Scene(f)
            for prec, pn in ((fb.FP32, "f32"), (fb.FP64, "f64")):
                ctx = fb.GpuContext(sc, 0, prec)
                best = None
                for it in range(3):
                    img, s = ctx.render()
                    best = s if best is None or s.device_ms < best.device_ms else best
                print(f"TIME {name:14s} {pn} {sc.width}x{sc.height} spp {sc.spp}: {best.device_ms:9.2f} ms  rays {best.rays} "
                      f"-> {best.rays / best.device_ms / 1e3:9.1f} Mrays/s  mean {img.mean():.5f}", flush=True)
                ctx.close()


if __name__ == "__main__":
    main()
