"""The reference-side binding of INTEGRATION.md, compiled and run (oracle/_ref/fray_ref_gpu, built by oracle/Makefile from the
UNMODIFIED reference sources + oracle/ref_binding/ref_gpu_main.cpp).

Its FlatBuilder walks the reference's own prepared `Scene` (/root/reference/src/scene.h:280-299, main.cpp:502-514) and fills
the tables of include/fray_gpu.h. (1) Those tables must equal, byte for byte, the ones this repository's host layer
(fray_b200/host/*.cpp: its own parser, OBJ loader, KD builder, EXR reader, flatten.cpp) produces for the same scene file --
on every bundled scene. (2) On a GPU box the binary then renders through the C ABI (fray_gpu_create / update_camera / render),
i.e. the patched render() of src/main.cpp:373-405, and must produce the frame `fray --gpu` produces."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import fray_b200 as fb
import oracle_util as ou
from conftest import local_scene

REF_GPU = os.path.join(ou.REF_DIR, "fray_ref_gpu")
HOST_DUMP = os.path.join(ou.REF_DIR, "host_flat_dump")
BUNDLED = ["boxed", "zaphod", "cornell_box", "smallpt", "forest", "hw9/axe_test", "hw9/nonconvex", "hw9/dragon", "hw10/bokeh", "hw12/sphtri"]


def need_binaries():
    if not (os.path.exists(REF_GPU) and os.path.exists(HOST_DUMP)):
        pytest.skip("oracle/_ref/fray_ref_gpu is not built (make -C oracle ref, where /root/reference exists)")


def sections(path):
    out = {}
    with open(path, "rb") as f:
        while True:
            tag = f.read(16)
            if not tag:
                return out
            n = int(np.frombuffer(f.read(8), np.uint64)[0])
            out[tag.rstrip(b"\0").decode()] = f.read(n)


@pytest.mark.parametrize("name", BUNDLED + ["local:csg_layered"])
def test_reference_scene_flattens_to_the_same_tables(name, tmp_path, data_dir):
    need_binaries()
    path = local_scene(name[6:]) if name.startswith("local:") else ou.scene_path(name)
    a, b = str(tmp_path / "ref.flat"), str(tmp_path / "host.flat")
    subprocess.run([REF_GPU, path, "--dump", a], check=True, capture_output=True)
    subprocess.run([HOST_DUMP, path, b], check=True, capture_output=True)
    sa, sb = sections(a), sections(b)
    assert list(sa) == list(sb) and len(sa) == 26
    for tag in sa:
        assert sa[tag] == sb[tag], f"{name}: section {tag} differs ({len(sa[tag])} vs {len(sb[tag])} bytes)"
    assert len(sa["nodes"]) > 0 and len(sa["camera"]) == C.sizeof(fb.FrayCamera)


@pytest.mark.gpu
@pytest.mark.parametrize("name,flag,precision", [("cornell_box", None, fb.FP32), ("boxed", None, fb.FP32), ("forest", "--fp64", fb.FP64)])
def test_reference_binary_renders_through_the_c_abi(name, flag, precision, tmp_path, data_dir):
    need_binaries()
    settings = dict(frameWidth=96, frameHeight=72, interactive="off")
    if name == "cornell_box":
        settings["pathsPerPixel"] = 8
    path = ou.override_scene(name, "refgpu", settings)
    out = str(tmp_path / "frame.f32")
    cmd = [REF_GPU, path, "--render", out, "--lib", os.path.join(fb.REPO_ROOT, "fray_b200", "libfray_gpu.so")] + ([flag] if flag else [])
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    got = ou.read_dump(out)
    sc = fb.Scene(path)
    ctx = fb.GpuContext(sc, 0, precision)
    want, stats = ctx.render(seed=42)
    ctx.close()
    assert np.array_equal(got, want)
    assert f"{stats.rays} rays" in r.stdout
