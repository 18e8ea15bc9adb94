#!/usr/bin/env python3
"""Device time of ONE rank's share of the headline frame on cuda:0 (what a rank of an N-GPU run computes), for several
sample-chunk sizes (FRAY_GPU_CHUNK): the fixed costs that limit strong scaling show up here without an N-GPU box.

    python tools/share_time.py [--world 8] [--spp 256] [--chunks 0,1,2,4]       (0 = the library's own rule)
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import fray_b200 as fb
from fray_b200 import scenes


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--world", type=int, default=8)
    ap.add_argument("--spp", type=int, default=256)
    ap.add_argument("--chunks", default="0,1,2,4")
    ap.add_argument("--scene", default="cornell_box")
    a = ap.parse_args()
    sc = fb.Scene(scenes.override_scene(a.scene, "share", dict(pathsPerPixel=a.spp)))
    ctx = fb.GpuContext(sc, 0, fb.FP32)
    _, full = ctx.render()
    _, full = ctx.render()
    print(f"full frame: {full.device_ms:.3f} ms -> ideal share {full.device_ms / a.world:.3f} ms")
    for mode, kw in (("samples", dict(sample_begin=0, sample_end=a.spp // a.world)), ("tiles", dict(bucket_rank=min(1, a.world - 1), bucket_count=a.world))):
        for c in [int(x) for x in a.chunks.split(",")]:
            if c:
                os.environ["FRAY_GPU_CHUNK"] = str(c)
            else:
                os.environ.pop("FRAY_GPU_CHUNK", None)
            best = None
            for _ in range(6):
                _, s = ctx.render(flags=fb.FRAME_SUM, **kw)
                best = s.device_ms if best is None else min(best, s.device_ms)
            print(f"{mode:8s} chunk {c if c else 'auto':>4}: {best:.3f} ms ({s.kernel_launches} launches, {s.rays} rays)")
    ctx.close()


if __name__ == "__main__":
    main()
