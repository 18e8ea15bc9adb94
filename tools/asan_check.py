#!/usr/bin/env python3
"""Memory-safety check of the per-ray device code without a GPU.

compute-sanitizer is closed on the GPU pool, so the templated __host__ __device__ core (fray_b200/csrc/core.cuh, flat.cuh,
scene_image.h) is compiled for the host with AddressSanitizer + UBSan (tests/emul/kernel_emul.cpp) and every golden scene
plus the synthetic flat-table scene is rendered in both precisions, beauty and AOV. Re-executes itself with the sanitizer
run-times preloaded.

    python tools/asan_check.py
"""
import ctypes as C
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = "/tmp/libfray_emul_asan.so"


def main():
    if os.environ.get("FRAY_ASAN_CHILD") != "1":
        subprocess.check_call(["g++", "-O1", "-g", "-fsanitize=address,undefined", "-fno-omit-frame-pointer", "-ffp-contract=off", "-std=c++17", "-fPIC",
                               "-shared", "-Wno-unknown-pragmas", "-x", "c++", os.path.join(ROOT, "tests", "emul", "kernel_emul.cpp"), "-o", SO, "-lpthread"])
        libs = [subprocess.check_output(["g++", f"-print-file-name={n}"], text=True).strip() for n in ("libasan.so", "libubsan.so")]
        env = dict(os.environ, FRAY_ASAN_CHILD="1", LD_PRELOAD=":".join(libs), ASAN_OPTIONS="detect_leaks=0:verify_asan_link_order=0")
        sys.exit(subprocess.call([sys.executable, os.path.abspath(__file__)], env=env))
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import numpy as np

    import fray_b200 as fb
    from fray_b200 import scenes
    from conftest import golden_scene
    import shutil
    cases = json.load(open(os.path.join(ROOT, "tests", "golden", "cases.json")))
    lib = C.CDLL(SO)
    lib.fray_emul_render.argtypes = [C.c_void_p, C.POINTER(fb.FrayFrame), C.c_void_p, C.POINTER(fb.FrayStats), C.c_int, C.c_int]

    def render(scene, precision, **kw):
        out = np.empty((scene.height, scene.width, 3), np.float32)
        frame, stats = fb.make_frame(**kw), fb.FrayStats()
        assert lib.fray_emul_render(scene.flat, C.byref(frame), out.ctypes.data, C.byref(stats), precision, 4) == 0
        assert kw.get("mode", fb.RENDER_BEAUTY) != fb.RENDER_BEAUTY or np.isfinite(out).all()  # a miss has distance inf in the AOV

    lib.fray_emul_render_wave.argtypes = [C.c_void_p, C.POINTER(fb.FrayFrame), C.c_void_p, C.POINTER(fb.FrayStats), C.c_int]

    def render_wave(scene, kd_short, **kw):
        """the wavefront stages (wave.cuh: short-stack KD walk with leaf boxes, origin mask, light loop); -2: not a wavefront scene"""
        out = np.empty((scene.height, scene.width, 3), np.float32)
        frame, stats = fb.make_frame(**kw), fb.FrayStats()
        rc = lib.fray_emul_render_wave(scene.flat, C.byref(frame), out.ctypes.data, C.byref(stats), kd_short)
        assert rc in (0, -2) and (rc != 0 or np.isfinite(out).all())

    todo = [golden_scene(cases, name) for name in cases]
    extra = os.path.join(scenes.DATA_DIR, "flat_transforms__test.fray")
    shutil.copyfile(os.path.join(ROOT, "tests", "scenes", "flat_transforms.fray"), extra)
    todo.append((extra, 7))
    # the open tubes (convex solids with two cap planes) of tests/test_flat_table.py
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from test_flat_table import OPEN_TUBE_OBJ
    with open(os.path.join(scenes.DATA_DIR, "geom", "open_tube__test.obj"), "w") as f:
        f.write(OPEN_TUBE_OBJ)
    tube = os.path.join(scenes.DATA_DIR, "open_tube__test.fray")
    shutil.copyfile(os.path.join(ROOT, "tests", "scenes", "open_tube.fray"), tube)
    todo.append((tube, 5))
    for path, seed in todo:
        sc = fb.Scene(path)
        for precision in (fb.FP32, fb.FP64):
            render(sc, precision, seed=seed)
            render(sc, precision, mode=fb.RENDER_AOV)
            render(sc, precision, seed=seed, bucket_rank=1, bucket_count=3)
        render_wave(sc, 0, seed=seed)
        render_wave(sc, 2, seed=seed)  # a two-entry short stack: kd-restarts on every deep walk
        print("clean:", os.path.basename(path), flush=True)
    print("asan/ubsan: no findings")


if __name__ == "__main__":
    main()
