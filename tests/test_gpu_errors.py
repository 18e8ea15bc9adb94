"""Error behaviour of the C ABI on a GPU box."""
import ctypes as C

import numpy as np
import pytest

import fray_b200 as fb
import oracle_util as ou
from conftest import golden_scene

pytestmark = pytest.mark.gpu


def test_bad_arguments(golden_cases, data_dir):
    lib = fb.gpu_lib()
    assert lib.fray_gpu_device_count() >= 1
    path, _ = golden_scene(golden_cases, "forest")
    sc = fb.Scene(path)
    with pytest.raises(fb.FrayError, match="device ordinal"):
        fb.GpuContext(sc, 99, fb.FP32)
    with pytest.raises(fb.FrayError, match="precision"):
        fb.GpuContext(sc, 0, 7)
    ctx = fb.GpuContext(sc, 0, fb.FP32)
    with pytest.raises(fb.FrayError, match="sample range"):
        ctx.render(sample_begin=3, sample_end=2)
    with pytest.raises(fb.FrayError, match="more than 5 samples"):
        ctx.render(spp=9)
    with pytest.raises(fb.FrayError, match="bucket_rank"):
        ctx.render(bucket_rank=4, bucket_count=4)
    with pytest.raises(fb.FrayError, match="render mode"):
        ctx.render(mode=5)
    # a context survives errors
    img, stats = ctx.render()
    assert np.isfinite(img).all() and stats.rays > 0
    ctx.close()


def test_malformed_scene_is_rejected(golden_cases, data_dir):
    path, _ = golden_scene(golden_cases, "forest")
    sc = fb.Scene(path)
    head = sc.head
    saved = head.abi_version
    head.abi_version = 999
    try:
        with pytest.raises(fb.FrayError, match="abi_version"):
            fb.GpuContext(sc, 0, fb.FP32)
    finally:
        head.abi_version = saved


def test_an_empty_sample_share_renders_nothing(golden_cases, data_dir):
    """FRAME_SAMPLE_RANGE makes [begin, end) literal: begin == end is an empty share (zeros, no rays), not "all samples" --
    what a launcher computing r * spp / world gets when spp < world."""
    path, seed = golden_scene(golden_cases, "cornell_box")
    sc = fb.Scene(path)
    ctx = fb.GpuContext(sc, 0, fb.FP32)
    full, fs = ctx.render(seed=seed, flags=fb.FRAME_SUM)
    empty, es = ctx.render(seed=seed, flags=fb.FRAME_SUM, sample_begin=3, sample_end=3)
    assert not empty.any() and es.rays == 0
    zero, zs = ctx.render(seed=seed, flags=fb.FRAME_SUM, sample_begin=0, sample_end=0)
    assert not zero.any() and zs.rays == 0
    again, _ = ctx.render(seed=seed, flags=fb.FRAME_SUM)  # without the flag 0,0 still means everything
    assert np.array_equal(again, full) and fs.rays > 0
    ctx.close()
    want, _ = ou.oracle_render(sc, seed=seed, flags=fb.FRAME_SUM, sample_begin=3, sample_end=3)
    assert not want.any()
