#!/usr/bin/env python3
"""Mirror the reference's bundled scenes/assets into oracle/_ref/data and decode its EXR files.

TEST INFRASTRUCTURE. Runs only where /root/reference exists (the build container); the result
(oracle/_ref/data, git-ignored) travels to the GPU box with the snapshot. Scene files and assets are
input DATA, not source code: they are what both the reference binary and the product render.

For every *.exr a side-car "<name>.exr.f32" (int32 w, int32 h, float32 rgba[h][w][4]) is written,
decoded with OpenCV's bundled OpenEXR. The headless reference build reads the side-cars through
oracle/ref_shim/OpenEXR/ImfRgbaFile.h; tests/test_exr.py uses them to check the product's own
PIZ decoder texel for texel.
"""
import os
import shutil
import sys

os.environ["OPENCV_IO_ENABLE_OPENEXR"] = "1"
import numpy as np


def main():
    src = sys.argv[1] if len(sys.argv) > 1 else "/root/reference/data"
    dst = sys.argv[2] if len(sys.argv) > 2 else os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "data")
    if not os.path.isdir(src):
        print(f"mirror_data: {src} not present, keeping existing {dst}")
        return 0
    os.makedirs(os.path.dirname(dst), exist_ok=True)
    shutil.copytree(src, dst, dirs_exist_ok=True)
    import cv2
    n = 0
    for root, _, files in os.walk(dst):
        for fn in files:
            if not fn.lower().endswith(".exr"):
                continue
            path = os.path.join(root, fn)
            im = cv2.imread(path, cv2.IMREAD_UNCHANGED)
            if im is None:
                print(f"mirror_data: cannot decode {path}", file=sys.stderr)
                return 1
            if im.ndim == 2:
                im = np.stack([im, im, im, np.ones_like(im)], axis=-1)
            if im.shape[2] == 3:
                im = np.concatenate([im, np.ones_like(im[..., :1])], axis=-1)
            rgba = np.ascontiguousarray(im[..., [2, 1, 0, 3]].astype(np.float32))  # BGRA -> RGBA
            with open(path + ".f32", "wb") as f:
                np.array([rgba.shape[1], rgba.shape[0]], dtype=np.int32).tofile(f)
                rgba.tofile(f)
            n += 1
    print(f"mirror_data: mirrored {src} -> {dst}, decoded {n} EXR files")
    return 0


if __name__ == "__main__":
    sys.exit(main())
