"""The C-ABI libraries load and export every symbol the headers declare (no compute, no GPU needed)."""
import ctypes as C
import os
import re

import pytest

import fray_b200 as fb

INCLUDE = os.path.join(fb.REPO_ROOT, "include")


def declared_functions(header):
    text = open(os.path.join(INCLUDE, header)).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(fray_(?:gpu|host)_[a-z0-9_]+)\s*\(", text)))


def test_host_library_exports_header():
    lib = fb.host_lib()
    names = declared_functions("fray_host.h")
    assert len(names) >= 10
    for n in names:
        assert hasattr(lib, n), f"libfray_host.so lacks {n}"


def test_gpu_library_exports_header():
    path = os.path.join(fb.REPO_ROOT, "fray_b200", "libfray_gpu.so")
    if not os.path.exists(path):
        pytest.fail("fray_b200/libfray_gpu.so is not built: run __graft_entry__.build() (nvcc cross-compiles without a GPU)")
    lib = fb.gpu_lib()
    names = declared_functions("fray_gpu.h")
    assert {"fray_gpu_create", "fray_gpu_render", "fray_gpu_render_device", "fray_gpu_destroy", "fray_gpu_last_error"} <= set(names)
    for n in names:
        assert hasattr(lib, n), f"libfray_gpu.so lacks {n}"
    assert lib.fray_gpu_abi_version() == 1


def test_struct_layouts_match_header():
    # sizes the ctypes mirrors must have for the by-pointer structs (checked against the C compiler once, here by arithmetic)
    assert C.sizeof(fb.FrayFrame) == 32
    assert C.sizeof(fb.FrayStats) == 40
    assert C.sizeof(fb.FraySettings) == 40
    assert C.sizeof(fb.FrayCamera) == 21 * 8 + 5 * 8 + 6 * 4 + 8


def test_no_cpu_fallback_without_device():
    """Without a CUDA device the product refuses to render (there is no CPU path)."""
    lib = fb.gpu_lib()
    if lib.fray_gpu_device_count() > 0:
        pytest.skip("a CUDA device is present")
    import oracle_util as ou
    if not os.path.isdir(ou.DATA_DIR):
        pytest.skip("no scene data")
    sc = fb.Scene(ou.scene_path("cornell_box"))
    with pytest.raises(fb.FrayError, match="no CUDA device|CPU fallback"):
        fb.GpuContext(sc)
