"""fray_b200 -- B200-native back end for fray's per-pixel render loop.

Python is only plumbing here: ctypes bindings over the two C-ABI libraries

* ``libfray_host.so``  (include/fray_host.h) -- the host scene layer: `.fray` parser, scene object model, OBJ loader,
  parity-exact KD builder, BMP/EXR IO, flattening;
* ``libfray_gpu.so``   (include/fray_gpu.h)  -- the hand-written sm_100a CUDA renderer behind the drop-in boundary
  that replaces ``pool.run(&RendMT)`` of /root/reference/src/main.cpp:402-404,

plus the ``torch.distributed`` sharding used by ``bench.py`` (fray_b200/dist.py). There is no CPU render path in
this package: without the CUDA library or a CUDA device every render call raises.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
REPO_ROOT = os.path.dirname(_HERE)

FP32 = 0
FP64 = 1
RENDER_BEAUTY = 0
RENDER_AOV = 1
RENDER_PREPASS = 2
FRAME_SUM = 1
FRAME_OWNED_ONLY = 2
FRAME_SAMPLE_RANGE = 4


class FraySettings(C.Structure):  # FrayGpuSettings
    _fields_ = [
        ("frame_width", C.c_int32), ("frame_height", C.c_int32), ("max_trace_depth", C.c_int32), ("gi", C.c_int32),
        ("num_paths", C.c_int32), ("want_aa", C.c_int32), ("ambient", C.c_float * 3), ("saturation", C.c_float),
    ]


class FrayCamera(C.Structure):  # FrayGpuCamera
    _fields_ = [
        ("pos", C.c_double * 3), ("top_left", C.c_double * 3), ("top_right", C.c_double * 3), ("bottom_left", C.c_double * 3),
        ("front", C.c_double * 3), ("up", C.c_double * 3), ("right", C.c_double * 3), ("w", C.c_double), ("h", C.c_double),
        ("aperture_size", C.c_double), ("focal_plane_dist", C.c_double), ("stereo_separation", C.c_double),
        ("left_mask", C.c_float * 3), ("right_mask", C.c_float * 3), ("dof", C.c_int32), ("num_dof_samples", C.c_int32),
    ]


class FraySceneHead(C.Structure):  # leading members of FrayGpuScene (the tables that follow are opaque to Python)
    _fields_ = [("abi_version", C.c_uint32), ("reserved", C.c_uint32), ("settings", FraySettings), ("camera", FrayCamera),
                ("num_nodes", C.c_int32)]


class FrayFrame(C.Structure):  # FrayGpuFrame
    _fields_ = [
        ("spp", C.c_int32), ("seed", C.c_uint32), ("sample_begin", C.c_int32), ("sample_end", C.c_int32),
        ("bucket_rank", C.c_int32), ("bucket_count", C.c_int32), ("mode", C.c_int32), ("flags", C.c_uint32),
    ]


class FrayStats(C.Structure):  # FrayGpuStats
    _fields_ = [
        ("rays", C.c_uint64), ("primary_rays", C.c_uint64), ("shadow_rays", C.c_uint64), ("device_ms", C.c_double),
        ("kernel_launches", C.c_int32), ("reserved", C.c_int32),
    ]


@dataclass
class RenderStats:
    rays: int
    primary_rays: int
    shadow_rays: int
    device_ms: float
    kernel_launches: int

    @staticmethod
    def of(s: FrayStats) -> "RenderStats":
        return RenderStats(int(s.rays), int(s.primary_rays), int(s.shadow_rays), float(s.device_ms), int(s.kernel_launches))


class FrayError(RuntimeError):
    pass


_host = None
_gpu = None


def host_lib() -> C.CDLL:
    """libfray_host.so (built by __graft_entry__.build() / fray_b200.build)."""
    global _host
    if _host is None:
        path = os.path.join(_HERE, "libfray_host.so")
        if not os.path.exists(path):
            raise FrayError(f"{path} is missing: run `python -c 'import __graft_entry__ as g; g.build()'`")
        L = C.CDLL(path)
        L.fray_host_load_scene.restype = C.c_void_p
        L.fray_host_load_scene.argtypes = [C.c_char_p]
        L.fray_host_free_scene.argtypes = [C.c_void_p]
        L.fray_host_last_error.restype = C.c_char_p
        L.fray_host_flat_scene.restype = C.c_void_p
        L.fray_host_flat_scene.argtypes = [C.c_void_p]
        L.fray_host_samples_per_pixel.argtypes = [C.c_void_p]
        L.fray_host_set_int.argtypes = [C.c_void_p, C.c_char_p, C.c_int]
        L.fray_host_move_camera.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_double, C.c_double, C.POINTER(FrayCamera)]
        L.fray_host_mesh_stats.argtypes = [C.c_void_p, C.c_int] + [C.POINTER(C.c_int)] * 4
        L.fray_host_save_image.argtypes = [C.c_char_p, C.c_void_p, C.c_int, C.c_int]
        L.fray_host_load_image.argtypes = [C.c_char_p, C.POINTER(C.POINTER(C.c_float)), C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.fray_host_free_pixels.argtypes = [C.POINTER(C.c_float)]
        L.fray_host_rng_draws.argtypes = [C.c_uint32] * 4 + [C.c_int, C.c_void_p]
        L.fray_host_rng_child.restype = C.c_uint32
        L.fray_host_rng_child.argtypes = [C.c_uint32] * 3
        _host = L
    return _host


def gpu_lib() -> C.CDLL:
    """libfray_gpu.so -- the CUDA renderer. Raises if it was not built (there is no fallback)."""
    global _gpu
    if _gpu is None:
        path = os.environ.get("FRAY_GPU_LIB") or os.path.join(_HERE, "libfray_gpu.so")  # FRAY_GPU_LIB: another build of the same library (A/B timing of kernel variants, tools/build_variant.py)
        if not os.path.exists(path):
            raise FrayError(f"{path} is missing: the CUDA back end must be built with nvcc (see __graft_entry__.build); "
                            "fray_b200 has no CPU render path")
        L = C.CDLL(path)
        L.fray_gpu_abi_version.restype = C.c_uint32
        L.fray_gpu_last_error.restype = C.c_char_p
        L.fray_gpu_samples_per_pixel.argtypes = [C.c_void_p]
        L.fray_gpu_create.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_void_p)]
        L.fray_gpu_update_camera.argtypes = [C.c_void_p, C.POINTER(FrayCamera)]
        L.fray_gpu_render.argtypes = [C.c_void_p, C.POINTER(FrayFrame), C.c_void_p, C.POINTER(FrayStats)]
        L.fray_gpu_render_device.argtypes = [C.c_void_p, C.POINTER(FrayFrame), C.c_void_p, C.c_void_p]
        L.fray_gpu_render_to_host.argtypes = [C.c_void_p, C.POINTER(FrayFrame), C.c_void_p, C.c_void_p]
        L.fray_gpu_host_register.argtypes = [C.c_void_p, C.c_size_t]
        L.fray_gpu_host_unregister.argtypes = [C.c_void_p]
        L.fray_gpu_resolve_device.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]
        L.fray_gpu_resolve_to_host.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]
        L.fray_gpu_frame_export.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.c_char_p]
        L.fray_gpu_frame_import.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(C.c_void_p)]
        L.fray_gpu_frame_close.argtypes = [C.c_void_p, C.c_void_p]
        L.fray_gpu_sync.argtypes = [C.c_void_p, C.POINTER(FrayStats)]
        L.fray_gpu_destroy.argtypes = [C.c_void_p]
        L.fray_gpu_measure_peaks.argtypes = [C.c_int, C.c_double, C.POINTER(C.c_double), C.POINTER(C.c_double)]
        L.fray_gpu_multi_create.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int), C.c_int, C.POINTER(C.c_void_p)]
        L.fray_gpu_multi_device_count.argtypes = [C.c_void_p]
        L.fray_gpu_multi_update_camera.argtypes = [C.c_void_p, C.POINTER(FrayCamera)]
        L.fray_gpu_multi_render.argtypes = [C.c_void_p, C.POINTER(FrayFrame), C.c_int, C.c_void_p, C.POINTER(FrayStats)]
        L.fray_gpu_multi_destroy.argtypes = [C.c_void_p]
        _gpu = L
    return _gpu


class Scene:
    """A parsed + prepared scene (mirrors ``scene.parseScene(); scene.beginRender(); scene.beginFrame()`` of
    /root/reference/src/main.cpp:503-514) together with its flattened tables."""

    def __init__(self, path: str):
        self._lib = host_lib()
        self.path = path
        self._h = self._lib.fray_host_load_scene(os.fsencode(path))
        if not self._h:
            raise FrayError(self._lib.fray_host_last_error().decode(errors="replace"))

    def close(self):
        if getattr(self, "_h", None):
            self._lib.fray_host_free_scene(self._h)
            self._h = None

    __del__ = close

    @property
    def flat(self) -> int:
        """Address of the FrayGpuScene (const FrayGpuScene*)."""
        return self._lib.fray_host_flat_scene(self._h)

    @property
    def head(self) -> FraySceneHead:
        return FraySceneHead.from_address(self.flat)

    @property
    def width(self) -> int:
        return self.head.settings.frame_width

    @property
    def height(self) -> int:
        return self.head.settings.frame_height

    @property
    def spp(self) -> int:
        return self._lib.fray_host_samples_per_pixel(self._h)

    def set(self, **kv: int) -> "Scene":
        for k, v in kv.items():
            if self._lib.fray_host_set_int(self._h, k.encode(), int(v)) != 0:
                raise FrayError(f"unknown scene override {k}")
        return self

    def move_camera(self, dx=0.0, dz=0.0, dyaw=0.0, dpitch=0.0) -> FrayCamera:
        cam = FrayCamera()
        self._lib.fray_host_move_camera(self._h, dx, dz, dyaw, dpitch, C.byref(cam))
        return cam

    def mesh_stats(self):
        out = []
        i = 0
        while True:
            v = [C.c_int() for _ in range(4)]
            if self._lib.fray_host_mesh_stats(self._h, i, *[C.byref(x) for x in v]) != 0:
                return out
            out.append(dict(nodes=v[0].value, leaf_refs=v[1].value, max_depth=v[2].value, triangles=v[3].value))
            i += 1


def make_frame(spp=0, seed=42, sample_begin=None, sample_end=None, bucket_rank=0, bucket_count=0, mode=RENDER_BEAUTY, flags=0) -> FrayFrame:
    """A FrayGpuFrame. A sample range given explicitly is literal (FRAME_SAMPLE_RANGE): begin == end is an empty share."""
    if sample_begin is not None or sample_end is not None:
        flags |= FRAME_SAMPLE_RANGE
    return FrayFrame(spp, seed, sample_begin or 0, sample_end or 0, bucket_rank, bucket_count, mode, flags)


class GpuContext:
    """``fray_gpu_create`` .. ``fray_gpu_destroy`` for one scene on one CUDA device."""

    def __init__(self, scene: Scene, device: int = 0, precision: int = FP32):
        self._lib = gpu_lib()
        self.scene = scene
        self.device = device
        self.precision = precision
        ctx = C.c_void_p()
        rc = self._lib.fray_gpu_create(scene.flat, device, precision, C.byref(ctx))
        if rc != 0:
            raise FrayError(f"fray_gpu_create failed ({rc}): {self._lib.fray_gpu_last_error().decode(errors='replace')}")
        self._ctx = ctx

    def close(self):
        if getattr(self, "_ctx", None):
            self._lib.fray_gpu_destroy(self._ctx)
            self._ctx = None

    __del__ = close

    def _check(self, rc, what):
        if rc != 0:
            raise FrayError(f"{what} failed ({rc}): {self._lib.fray_gpu_last_error().decode(errors='replace')}")

    def update_camera(self, cam: FrayCamera):
        self._check(self._lib.fray_gpu_update_camera(self._ctx, C.byref(cam)), "fray_gpu_update_camera")

    def render(self, out: np.ndarray | None = None, **frame_kw):
        """Render to a host float32 array [h, w, 3] (the end-to-end call: kernels + device->host copy)."""
        h, w = self.scene.height, self.scene.width
        if out is None:
            out = np.empty((h, w, 3), dtype=np.float32)
        assert out.dtype == np.float32 and out.flags.c_contiguous and out.size == h * w * 3
        frame = make_frame(**frame_kw)
        stats = FrayStats()
        self._check(self._lib.fray_gpu_render(self._ctx, C.byref(frame), out.ctypes.data, C.byref(stats)), "fray_gpu_render")
        return out, RenderStats.of(stats)

    def render_device(self, d_rgb: int, stream: int = 0, **frame_kw):
        """Render asynchronously into device memory (address `d_rgb`, h*w*3 floats)."""
        frame = make_frame(**frame_kw)
        self._check(self._lib.fray_gpu_render_device(self._ctx, C.byref(frame), d_rgb, stream), "fray_gpu_render_device")

    def render_to_host(self, pinned_host: int, stream: int = 0, **frame_kw):
        """Render this share (bucket_rank / bucket_count) and store the finished pixels of its own tiles straight into the
        page-locked host frame at `pinned_host` (fray_gpu_render_to_host); asynchronous, wait with sync()."""
        frame = make_frame(**frame_kw)
        self._check(self._lib.fray_gpu_render_to_host(self._ctx, C.byref(frame), pinned_host, stream), "fray_gpu_render_to_host")

    def resolve_device(self, d_sum: int, d_rgb: int, spp: int, stream: int = 0):
        self._check(self._lib.fray_gpu_resolve_device(self._ctx, d_sum, d_rgb, spp, stream), "fray_gpu_resolve_device")

    def resolve_to_host(self, d_sum: int, pinned_host: int, spp: int, stream: int = 0):
        """sum / spp stored straight into page-locked host memory (no separate device-to-host copy)."""
        self._check(self._lib.fray_gpu_resolve_to_host(self._ctx, d_sum, pinned_host, spp, stream), "fray_gpu_resolve_to_host")

    def frame_export(self) -> tuple[int, bytes]:
        """(device address, 64-byte IPC handle) of a context-owned frame other processes can map (fray_gpu_frame_export)."""
        ptr = C.c_void_p()
        handle = C.create_string_buffer(64)
        self._check(self._lib.fray_gpu_frame_export(self._ctx, C.byref(ptr), handle), "fray_gpu_frame_export")
        return int(ptr.value), handle.raw

    def frame_import(self, handle: bytes) -> int:
        """Map a frame exported by another process's context on another GPU of this node; returns the device address."""
        ptr = C.c_void_p()
        self._check(self._lib.fray_gpu_frame_import(self._ctx, handle, C.byref(ptr)), "fray_gpu_frame_import")
        return int(ptr.value)

    def frame_close(self, d_frame: int):
        self._check(self._lib.fray_gpu_frame_close(self._ctx, d_frame), "fray_gpu_frame_close")

    def sync(self) -> RenderStats:
        stats = FrayStats()
        self._check(self._lib.fray_gpu_sync(self._ctx, C.byref(stats)), "fray_gpu_sync")
        return RenderStats.of(stats)


SPLIT_AUTO, SPLIT_TILES, SPLIT_SAMPLES = 0, 1, 2
MULTI_FAST = 0x100


class MultiGpuContext:
    """``fray_gpu_multi_create`` .. ``fray_gpu_multi_destroy``: one scene on several CUDA devices of this node, driven by this
    one process (what ``pool.run(&worker, numThreads)`` is to the reference's host threads, src/main.cpp:402-404)."""

    def __init__(self, scene: Scene, devices, precision: int = FP32):
        self._lib = gpu_lib()
        self.scene = scene
        devs = list(range(devices)) if isinstance(devices, int) else list(devices)
        arr = (C.c_int * len(devs))(*devs)
        ctx = C.c_void_p()
        rc = self._lib.fray_gpu_multi_create(scene.flat, len(devs), arr, precision, C.byref(ctx))
        if rc != 0:
            raise FrayError(f"fray_gpu_multi_create failed ({rc}): {self._lib.fray_gpu_last_error().decode(errors='replace')}")
        self._ctx = ctx
        self.devices = devs

    def close(self):
        if getattr(self, "_ctx", None):
            self._lib.fray_gpu_multi_destroy(self._ctx)
            self._ctx = None

    __del__ = close

    def update_camera(self, cam: FrayCamera):
        if self._lib.fray_gpu_multi_update_camera(self._ctx, C.byref(cam)) != 0:
            raise FrayError(f"fray_gpu_multi_update_camera failed: {self._lib.fray_gpu_last_error().decode(errors='replace')}")

    def render(self, out: np.ndarray | None = None, split: int = SPLIT_AUTO, **frame_kw):
        h, w = self.scene.height, self.scene.width
        if out is None:
            out = np.empty((h, w, 3), dtype=np.float32)
        assert out.dtype == np.float32 and out.flags.c_contiguous and out.size == h * w * 3
        frame = make_frame(**frame_kw)
        stats = FrayStats()
        rc = self._lib.fray_gpu_multi_render(self._ctx, C.byref(frame), split, out.ctypes.data, C.byref(stats))
        if rc != 0:
            raise FrayError(f"fray_gpu_multi_render failed ({rc}): {self._lib.fray_gpu_last_error().decode(errors='replace')}")
        return out, RenderStats.of(stats)


def host_register(address: int, nbytes: int):
    """Page-lock foreign host memory (a shared-memory mapping ...) so that kernels can store into it (fray_gpu_host_register)."""
    if gpu_lib().fray_gpu_host_register(address, nbytes) != 0:
        raise FrayError("fray_gpu_host_register: " + gpu_lib().fray_gpu_last_error().decode())


def host_unregister(address: int):
    gpu_lib().fray_gpu_host_unregister(address)


def measure_peaks(device: int = 0, ms: float = 20.0) -> tuple[float, float]:
    """(sustained FP32 FFMA TFLOP/s, L2-resident read GB/s) measured on `device` right now."""
    L = gpu_lib()
    f, l2 = C.c_double(), C.c_double()
    rc = L.fray_gpu_measure_peaks(device, ms, C.byref(f), C.byref(l2))
    if rc != 0:
        raise FrayError(f"fray_gpu_measure_peaks failed ({rc}): {L.fray_gpu_last_error().decode(errors='replace')}")
    return f.value, l2.value


def save_image(path: str, rgb: np.ndarray):
    """BMP (8-bit, clamped, no gamma) or EXR (HALF RGBA) like the reference's F12 screenshots (src/sdl.cpp:101-140)."""
    rgb = np.ascontiguousarray(rgb, dtype=np.float32)
    if host_lib().fray_host_save_image(os.fsencode(path), rgb.ctypes.data, rgb.shape[1], rgb.shape[0]) != 0:
        raise FrayError(f"cannot write {path}")


def load_image(path: str) -> np.ndarray:
    L = host_lib()
    p = C.POINTER(C.c_float)()
    w, h = C.c_int(), C.c_int()
    if L.fray_host_load_image(os.fsencode(path), C.byref(p), C.byref(w), C.byref(h)) != 0:
        raise FrayError(f"cannot read {path}")
    try:
        return np.ctypeslib.as_array(p, shape=(h.value, w.value, 3)).copy()
    finally:
        L.fray_host_free_pixels(p)
