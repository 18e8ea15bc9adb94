"""fray_gpu_multi_*: several GPUs driven by one process (the drop-in for pool.run(&worker, numThreads),
/root/reference/src/main.cpp:402-404). On a box with one GPU the same code path runs with the device listed twice: two
contexts, two host threads, shares stored into one frame; with more GPUs the stores cross NVLink."""
import numpy as np
import pytest

import fray_b200 as fb
import oracle_util as ou
from conftest import golden_scene

pytestmark = pytest.mark.gpu


def device_list(n):
    have = fb.gpu_lib().fray_gpu_device_count()
    return [d % have for d in range(n)]


@pytest.mark.parametrize("name", ["cornell_box", "forest_aa", "boxed"])
@pytest.mark.parametrize("n", [2, 3])
def test_tile_split_is_bit_identical(name, n, golden_cases, data_dir):
    path, seed = golden_scene(golden_cases, name)
    sc = fb.Scene(path)
    ctx = fb.GpuContext(sc, 0, fb.FP32)
    want, ws = ctx.render(seed=seed)
    ctx.close()
    multi = fb.MultiGpuContext(sc, device_list(n), fb.FP32)
    got, gs = multi.render(seed=seed, split=fb.SPLIT_TILES)
    assert np.array_equal(got, want)
    assert gs.rays == ws.rays and gs.primary_rays == ws.primary_rays
    again, _ = multi.render(seed=seed, split=fb.SPLIT_TILES)
    assert np.array_equal(again, want)
    multi.close()


@pytest.mark.parametrize("name", ["cornell_box", "forest_aa"])
def test_sample_split_adds_up(name, golden_cases, data_dir):
    path, seed = golden_scene(golden_cases, name)
    sc = fb.Scene(path)
    ctx = fb.GpuContext(sc, 0, fb.FP32)
    want, ws = ctx.render(seed=seed)
    ctx.close()
    multi = fb.MultiGpuContext(sc, device_list(2), fb.FP32)
    got, gs = multi.render(seed=seed, split=fb.SPLIT_SAMPLES)
    np.testing.assert_allclose(got, want, rtol=2e-6, atol=1e-7)  # FP32 regrouping of a pixel's partial sums
    assert gs.rays == ws.rays
    auto, _ = multi.render(seed=seed)  # SPLIT_AUTO: samples only while every GPU keeps >= 8 of them
    np.testing.assert_allclose(auto, want, rtol=2e-6, atol=1e-7)
    multi.close()


def test_multi_matches_the_oracle_in_parity_precision(golden_cases, data_dir):
    path, seed = golden_scene(golden_cases, "cornell_box")
    sc = fb.Scene(path)
    want, ostats = ou.oracle_render(sc, seed=seed)
    multi = fb.MultiGpuContext(sc, device_list(2), fb.FP64)
    got, gs = multi.render(seed=seed, split=fb.SPLIT_TILES)
    assert ou.compare(want, got, 2e-5)[0] == 1.0 and gs.rays == ostats.rays
    multi.close()


def test_camera_moves_reach_every_share(golden_cases, data_dir):
    path, seed = golden_scene(golden_cases, "forest")
    sc = fb.Scene(path)
    multi = fb.MultiGpuContext(sc, device_list(2), fb.FP32)
    ctx = fb.GpuContext(sc, 0, fb.FP32)
    for step in range(3):
        if step:
            cam = sc.move_camera(dx=2.0, dz=-1.0, dyaw=7.0)
            multi.update_camera(cam)
            ctx.update_camera(cam)
        a, _ = multi.render(seed=seed, split=fb.SPLIT_TILES)
        b, _ = ctx.render(seed=seed)
        assert np.array_equal(a, b)
    multi.close()
    ctx.close()


def test_prepass_paints_the_centre_sample_over_16x16_squares(golden_cases, data_dir):
    """FRAY_RENDER_PREPASS: the preview of render(), /root/reference/src/main.cpp:376-391. In a scene without random pixel
    offsets the colour of a square is the colour of its centre pixel in the one-sample frame."""
    path, seed = golden_scene(golden_cases, "forest")
    sc = fb.Scene(path)
    for precision in (fb.FP64, fb.FP32):
        ctx = fb.GpuContext(sc, 0, precision)
        frame, _ = ctx.render(seed=seed)
        pre, ps = ctx.render(seed=seed, mode=fb.RENDER_PREPASS)
        ctx.close()
        h, w = frame.shape[:2]
        assert ps.rays > 0 and ps.primary_rays == ((w + 15) // 16) * ((h + 15) // 16)
        for y0 in range(0, h, 16):
            for x0 in range(0, w, 16):
                y1, x1 = min(h, y0 + 16), min(w, x0 + 16)
                cy, cx = (y0 + y1) // 2, (x0 + x1) // 2
                square = pre[y0:y1, x0:x1]
                assert (square == square[0, 0]).all()
                np.testing.assert_allclose(square[0, 0], frame[cy, cx], rtol=1e-5, atol=1e-6)


def test_cli_devices_and_prepass(tmp_path, data_dir):
    """`fray --gpu --devices 1 --prepass`: the multi-GPU entry of the command line tool (with the one GPU every box has) and the
    preview image next to the frame."""
    import subprocess
    import fray_b200.build as fbuild
    exe = fbuild.build_cli()
    scene_file = ou.override_scene("cornell_box", "climulti", dict(frameWidth=64, frameHeight=48, pathsPerPixel=8))
    out = str(tmp_path / "frame.bmp")
    r = subprocess.run([exe, "--gpu", "-v", "--devices", "1", "--split", "tiles", "--prepass", "--out", out, scene_file], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert "Render took" in r.stdout and "prepass:" in r.stdout and "on 1 GPUs" in r.stdout
    sc = fb.Scene(scene_file)
    ctx = fb.GpuContext(sc, 0, fb.FP32)
    want, _ = ctx.render(seed=42)
    ctx.close()
    got = fb.load_image(out)
    np.testing.assert_allclose(got, np.floor(np.clip(want, 0, 1) * np.float32(255) + np.float32(0.5)) / 255, atol=1e-6)
    pre = fb.load_image(str(tmp_path / "frame.prepass.bmp"))
    assert pre.shape == got.shape and (pre[:16, :16] == pre[0, 0]).all()


@pytest.mark.parametrize("name", ["cornell_box", "boxed", "forest_aa"])
def test_shares_gathered_straight_into_shared_host_memory(name, golden_cases, data_dir):
    """fray_gpu_render_to_host + fray_gpu_host_register (the `host` mode of fray_b200/dist.py): three shares of a tile split
    store their own tiles into one page-locked frame that lives in a shared-memory mapping -- the frame a single render call
    produces, pixel for pixel within FP32 regrouping (a share may cut a pixel's samples into other chunks), every pixel
    written exactly once. Path-traced (the kernels' own final stores go over PCIe) and wavefront (owned tiles copied out)."""
    import mmap
    path, seed = golden_scene(golden_cases, name)
    sc = fb.Scene(path)
    ctx = fb.GpuContext(sc, 0, fb.FP32)
    want, ws = ctx.render(seed=seed)
    nbytes = sc.height * sc.width * 3 * 4
    shm = mmap.mmap(-1, nbytes)  # anonymous shared mapping: what N rank processes would map by name
    view = np.frombuffer(shm, dtype=np.float32).reshape(sc.height, sc.width, 3)
    addr = view.ctypes.data
    fb.host_register(addr, nbytes)
    try:
        view[:] = np.nan
        rays = 0
        for rank in range(3):
            ctx.render_to_host(addr, 0, seed=seed, bucket_rank=rank, bucket_count=3)
            rays += ctx.sync().rays
            assert np.isnan(view).any() == (rank < 2)  # pixels of the other shares are not touched
        np.testing.assert_allclose(view, want, rtol=2e-6, atol=1e-7)
        assert rays == ws.rays
        with pytest.raises(fb.FrayError):
            ctx.render_to_host(addr, 0, seed=seed, flags=fb.FRAME_SUM)
    finally:
        fb.host_unregister(addr)
        ctx.close()
        del view
        shm.close()


def test_render_to_host_rejects_pageable_memory(golden_cases, data_dir):
    path, seed = golden_scene(golden_cases, "cornell_box")
    sc = fb.Scene(path)
    ctx = fb.GpuContext(sc, 0, fb.FP32)
    out = np.zeros((sc.height, sc.width, 3), np.float32)
    with pytest.raises(fb.FrayError):
        ctx.render_to_host(out.ctypes.data, 0, seed=seed)
    ctx.close()
