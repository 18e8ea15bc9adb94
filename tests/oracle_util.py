"""Test-side access to the CPU checkers under oracle/ (never imported by the product).

* ``oracle_render``      -- our FP64 restatement (oracle/libfray_oracle.so) on a flattened scene
* ``reference_render``   -- the real reference code with the counter RNG (oracle/_ref/fray_ref_ctr)
* ``override_scene``     -- re-exported from fray_b200.scenes (variants of bundled scenes)
"""
from __future__ import annotations

import ctypes as C
import os
import re
import subprocess
import tempfile

import numpy as np

import fray_b200 as fb

ROOT = fb.REPO_ROOT
ORACLE_DIR = os.path.join(ROOT, "oracle")
REF_DIR = os.path.join(ORACLE_DIR, "_ref")
from fray_b200.scenes import DATA_DIR  # noqa: E402
GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")
REF_CTR = os.path.join(REF_DIR, "fray_ref_ctr")
REF_BIN = os.path.join(REF_DIR, "fray_ref")
REF_STRICT = os.path.join(REF_DIR, "fray_ref_strict")

_oracle = None


def oracle_lib() -> C.CDLL:
    global _oracle
    if _oracle is None:
        path = os.path.join(ORACLE_DIR, "libfray_oracle.so")
        if not os.path.exists(path):
            subprocess.check_call(["make", "-C", ORACLE_DIR, "oracle"])
        L = C.CDLL(path)
        L.fray_oracle_render.argtypes = [C.c_void_p, C.POINTER(fb.FrayFrame), C.c_void_p, C.POINTER(fb.FrayStats), C.c_int]
        L.fray_oracle_samples_per_pixel.argtypes = [C.c_void_p]
        L.fray_oracle_rng_draws.argtypes = [C.c_uint32] * 4 + [C.c_int, C.c_void_p]
        L.fray_oracle_rng_child.restype = C.c_uint32
        L.fray_oracle_rng_child.argtypes = [C.c_uint32] * 3
        L.fray_oracle_philox.argtypes = [C.c_void_p] * 3
        L.fray_oracle_screen_ray.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_void_p]
        L.fray_oracle_intersect.argtypes = [C.c_void_p] * 4
        L.fray_oracle_environment.argtypes = [C.c_void_p] * 3
        L.fray_oracle_work_profile.argtypes = [C.c_void_p]
        _oracle = L
    return _oracle


def oracle_render(scene: fb.Scene, threads: int = 0, **frame_kw):
    """Render `scene` with the FP64 restatement; same frame semantics as GpuContext.render."""
    L = oracle_lib()
    out = np.empty((scene.height, scene.width, 3), dtype=np.float32)
    frame = fb.make_frame(**frame_kw)
    stats = fb.FrayStats()
    rc = L.fray_oracle_render(scene.flat, C.byref(frame), out.ctypes.data, C.byref(stats), threads)
    if rc != 0:
        raise RuntimeError("fray_oracle_render failed")
    return out, fb.RenderStats.of(stats)


PROFILE_KEYS = ["node_tests", "node_hits", "root_box_tests", "kd_inner", "kd_leaf", "tri_calls", "tri_tests", "plane_tests", "sphere_tests",
                "cube_tests", "light_tests", "light_samples", "hemi_samples", "brdf_evals", "texture_lookups", "leaf_refs"]

# algorithmic cost model of SURVEY.md section 8(d): FP32 flops (FMA = 2) and bytes per occurrence
FLOP_COST = dict(node_tests=43, node_hits=52, root_box_tests=20, kd_inner=4, kd_leaf=0, tri_calls=6, tri_tests=40, plane_tests=22, sphere_tests=40,
                 cube_tests=60, light_tests=55, light_samples=45, hemi_samples=60, brdf_evals=12, texture_lookups=15, leaf_refs=0)
BYTE_COST = dict(node_tests=96, node_hits=0, root_box_tests=24, kd_inner=16, kd_leaf=16, tri_calls=16, tri_tests=48, plane_tests=8, sphere_tests=16,
                 cube_tests=16, light_tests=96, light_samples=32, hemi_samples=0, brdf_evals=32, texture_lookups=12, leaf_refs=4)


def oracle_work_profile() -> dict:
    """Counters of the last oracle_render call (how often each piece of the hot path ran)."""
    buf = (C.c_uint64 * 16)()
    oracle_lib().fray_oracle_work_profile(buf)
    return dict(zip(PROFILE_KEYS, (int(v) for v in buf)))


def algorithmic_work_per_ray(profile: dict, rays: int) -> tuple[float, float]:
    """(flop per ray, bytes per ray) under the cost model above."""
    flop = sum(profile[k] * FLOP_COST[k] for k in PROFILE_KEYS) / rays
    byts = sum(profile[k] * BYTE_COST[k] for k in PROFILE_KEYS) / rays
    return flop, byts


def read_dump(path: str) -> np.ndarray:
    with open(path, "rb") as f:
        w, h = np.fromfile(f, np.int32, 2)
        return np.fromfile(f, np.float32, int(w) * int(h) * 3).reshape(int(h), int(w), 3)


def read_aov(path: str):
    with open(path, "rb") as f:
        w, h = np.fromfile(f, np.int32, 2)
        rec = np.fromfile(f, np.dtype([("node", "<i4"), ("dist", "<f8")]), int(w) * int(h))
    return rec["node"].reshape(int(h), int(w)), rec["dist"].reshape(int(h), int(w))


def have_reference() -> bool:
    return os.path.exists(REF_CTR) and os.path.isdir(DATA_DIR)


from fray_b200.scenes import override_scene, scene_path  # noqa: E402,F401  (the scene helpers are product code)


def reference_render(scene_file: str, seed: int = 42, threads: int = 0, aov: bool = False):
    """Run the real reference code on the counter-RNG contract. Returns (rgb, seconds[, node, dist])."""
    with tempfile.TemporaryDirectory() as td:
        out = os.path.join(td, "out.f32")
        cmd = [REF_CTR, scene_file, out, "--seed", str(seed)]
        if threads:
            cmd += ["--threads", str(threads)]
        if aov:
            cmd += ["--aov", os.path.join(td, "out.aov")]
        txt = subprocess.run(cmd, check=True, capture_output=True, text=True).stdout
        m = re.search(r"Render took ([0-9.]+)s", txt)
        sec = float(m.group(1)) if m else float("nan")
        rgb = read_dump(out)
        if aov:
            node, dist = read_aov(os.path.join(td, "out.aov"))
            return rgb, sec, node, dist
        return rgb, sec


def reference_render_unmodified(scene_file: str) -> np.ndarray:
    """The UNMODIFIED reference (oracle/_ref/fray_ref_strict: all of src/*.cpp incl. main.cpp's RendMT::entry bucket loop and
    libstdc++'s thread-keyed generators, strict IEEE flags) on `scene_file`; returns its framebuffer `vfb`, which the SDL
    stand-in dumps at exit ($FRAY_DUMP). Only meaningful for scenes that draw no random numbers (parity tier T0)."""
    with tempfile.TemporaryDirectory() as td:
        out = os.path.join(td, "vfb.f32")
        subprocess.run([REF_STRICT, scene_file], check=True, capture_output=True, text=True, env=dict(os.environ, FRAY_DUMP=out))
        return read_dump(out)


def compare(a: np.ndarray, b: np.ndarray, tol: float = 1e-3):
    """(fraction of pixels whose max channel difference <= tol, RMSE, max abs difference)."""
    d = np.abs(a.astype(np.float64) - b.astype(np.float64))
    per_px = d.max(axis=-1)
    return float((per_px <= tol).mean()), float(np.sqrt((d ** 2).mean())), float(per_px.max())
