"""Early-warning check that needs no GPU: the per-ray device code (fray_b200/csrc/core.cuh), compiled for the host by
tests/emul/kernel_emul.cpp, against the oracle. The real parity tests are the `-m gpu` ones (test_gpu_parity.py), which go
through the C ABI of libfray_gpu.so; this file only guards the shared templated core while developing without a GPU."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import fray_b200 as fb
import oracle_util as ou
from conftest import golden_scene, load_golden, local_scene

EMUL_DIR = os.path.join(fb.REPO_ROOT, "tests", "emul")


@pytest.fixture(scope="module")
def emul():
    so = os.path.join(EMUL_DIR, "libfray_emul.so")
    src = os.path.join(EMUL_DIR, "kernel_emul.cpp")
    deps = [src] + [os.path.join(fb.REPO_ROOT, "fray_b200", "csrc", f) for f in ("core.cuh", "flat.cuh", "rng.cuh", "scene_image.h", "wave.cuh")]
    if not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
        subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-std=c++17", "-fPIC", "-shared", "-Wno-unknown-pragmas", "-x", "c++", src,
                               "-o", so, "-lpthread"])
    lib = C.CDLL(so)
    lib.fray_emul_render.argtypes = [C.c_void_p, C.POINTER(fb.FrayFrame), C.c_void_p, C.POINTER(fb.FrayStats), C.c_int, C.c_int]

    def render(scene, precision, **kw):
        out = np.empty((scene.height, scene.width, 3), np.float32)
        frame, stats = fb.make_frame(**kw), fb.FrayStats()
        assert lib.fray_emul_render(scene.flat, C.byref(frame), out.ctypes.data, C.byref(stats), precision, 0) == 0
        return out, fb.RenderStats.of(stats)
    return render


@pytest.mark.parametrize("name", ["cornell_box", "smallpt", "boxed", "forest", "forest_stereo_dof", "bokeh", "dragon"])
def test_device_core_fp64_matches_golden(name, emul, golden_cases, data_dir):
    path, seed = golden_scene(golden_cases, name)
    sc = fb.Scene(path)
    want, ostats = ou.oracle_render(sc, seed=seed)
    got, stats = emul(sc, fb.FP64, seed=seed)
    frac, rmse, mx = ou.compare(want, got, 1e-5)
    assert frac == 1.0 and mx < 2e-5, (frac, rmse, mx)
    assert stats.rays == ostats.rays  # the same rays, one for one


@pytest.mark.parametrize("name", ["cornell_box", "smallpt", "boxed", "forest"])
def test_device_core_fp32_within_tolerance(name, emul, golden_cases, data_dir):
    path, seed = golden_scene(golden_cases, name)
    sc = fb.Scene(path)
    want, _, _ = load_golden(name)
    got, _ = emul(sc, fb.FP32, seed=seed)
    frac, rmse, mx = ou.compare(want, got, 1e-3)
    assert frac >= 0.997, (frac, rmse, mx)


def test_device_core_csg_layered_scene(emul, data_dir):
    """tests/scenes/csg_layered.fray -- Cube, CsgMinus / CsgAnd / nested CsgPlus, Layered with constant and Fresnel opacities,
    no bitmap -- against the image the reference code rendered (tests/golden/csg_layered.npz): parity precision exact, fast
    precision >= 99.9 % of pixels within 1e-3 (north_star's Whitted bar)."""
    sc = fb.Scene(local_scene("csg_layered"))
    ref, node, _ = load_golden("csg_layered")
    assert {0, 1, 2, 3, 4, 5, 6} <= set(np.unique(node).tolist())  # every solid of the scene is seen by some primary ray
    want, ostats = ou.oracle_render(sc, seed=42)
    assert np.array_equal(want, ref)  # the restatement reproduces the reference bit for bit
    got, stats = emul(sc, fb.FP64, seed=42)
    assert ou.compare(ref, got, 2e-5)[0] == 1.0 and stats.rays == ostats.rays
    got, _ = emul(sc, fb.FP32, seed=42)
    frac, rmse, mx = ou.compare(ref, got, 1e-3)
    assert frac >= 0.999, (frac, rmse, mx)
    aov, _ = emul(sc, fb.FP32, mode=fb.RENDER_AOV)
    assert (aov[..., 0].astype(int) == node).mean() >= 0.998


@pytest.fixture(scope="module")
def emul_wave(emul):
    lib = C.CDLL(os.path.join(EMUL_DIR, "libfray_emul.so"))
    lib.fray_emul_render_wave.argtypes = [C.c_void_p, C.POINTER(fb.FrayFrame), C.c_void_p, C.POINTER(fb.FrayStats), C.c_int]

    def render(scene, kd_short=0, **kw):
        out = np.empty((scene.height, scene.width, 3), np.float32)
        frame, stats = fb.make_frame(**kw), fb.FrayStats()
        assert lib.fray_emul_render_wave(scene.flat, C.byref(frame), out.ctypes.data, C.byref(stats), kd_short) == 0
        return out, fb.RenderStats.of(stats)
    return render


@pytest.mark.parametrize("name", ["boxed", "forest", "forest_aa", "forest_stereo_dof", "dragon"])
def test_wavefront_stages_match_the_megakernel(name, emul, emul_wave, golden_cases, data_dir):
    """The wavefront stages of the fast precision (csrc/wave.cuh: trace / shade / shadow over queues, the KD short stack, the
    light loop with its per-point record mask) against the recursion-as-a-task-stack form of core.cuh on the same scenes: the
    same rays one for one, the same frame up to the order of FP32 additions -- also with a two-entry short stack, which forces
    kd-restarts on every deep walk."""
    path, seed = golden_scene(golden_cases, name)
    sc = fb.Scene(path)
    want, wstats = emul(sc, fb.FP32, seed=seed)
    for kd_short in (0, 2):
        got, stats = emul_wave(sc, kd_short, seed=seed)
        frac, rmse, mx = ou.compare(want, got, 1e-5)
        assert frac >= 0.9999 and rmse < 1e-5, (kd_short, frac, rmse, mx)
        assert stats.rays == wstats.rays and stats.shadow_rays == wstats.shadow_rays
