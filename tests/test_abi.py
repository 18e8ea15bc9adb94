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


def test_struct_layouts_match_header(tmp_path):
    """sizeof / offsetof of the by-pointer structs as the C COMPILER lays them out from include/fray_gpu.h, against the ctypes
    mirrors in fray_b200/__init__.py (a C program is compiled and run; it also proves the header is plain C)."""
    import subprocess
    src = tmp_path / "layout.c"
    src.write_text("""
#include <stdio.h>
#include <stddef.h>
#include "fray_gpu.h"
#define S(t) printf(#t " %zu\\n", sizeof(t))
#define O(t, m) printf(#t "." #m " %zu\\n", offsetof(t, m))
int main(void) {
  S(FrayGpuFrame); S(FrayGpuStats); S(FrayGpuSettings); S(FrayGpuCamera); S(FrayGpuTransform); S(FrayGpuNode); S(FrayGpuGeometry);
  S(FrayGpuMesh); S(FrayGpuKdNode); S(FrayGpuShader); S(FrayGpuLayer); S(FrayGpuTexture); S(FrayGpuBitmap); S(FrayGpuLight); S(FrayGpuScene);
  O(FrayGpuFrame, flags); O(FrayGpuStats, device_ms); O(FrayGpuCamera, dof); O(FrayGpuCamera, left_mask); O(FrayGpuSettings, saturation);
  return 0; }
""")
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-I", INCLUDE, str(src), "-o", str(exe)])
    got = dict(line.split() for line in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.splitlines())
    got = {k: int(v) for k, v in got.items()}
    assert got["FrayGpuFrame"] == C.sizeof(fb.FrayFrame) == 32
    assert got["FrayGpuStats"] == C.sizeof(fb.FrayStats) == 40
    assert got["FrayGpuSettings"] == C.sizeof(fb.FraySettings) == 40
    assert got["FrayGpuCamera"] == C.sizeof(fb.FrayCamera)
    assert got["FrayGpuFrame.flags"] == fb.FrayFrame.flags.offset
    assert got["FrayGpuStats.device_ms"] == fb.FrayStats.device_ms.offset
    assert got["FrayGpuCamera.dof"] == fb.FrayCamera.dof.offset
    assert got["FrayGpuCamera.left_mask"] == fb.FrayCamera.left_mask.offset
    assert got["FrayGpuSettings.saturation"] == fb.FraySettings.saturation.offset
    assert got["FrayGpuTransform"] == 21 * 8 and got["FrayGpuKdNode"] == 24 and got["FrayGpuBitmap"] == 16


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
