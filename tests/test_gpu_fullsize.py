"""BASELINE.json's configurations at their FULL sizes on the GPU.

The oracle and the reference need minutes to hours for these, so the small-frame tests (test_gpu_parity.py) pin the device
code to them, and here the two device precisions check each other at full size: the parity precision (FP64, bit-for-bit the
oracle on every small frame) is the stand-in for the CPU image, the fast precision (FP32, the benchmarked path) must stay
within the tolerances of BASELINE.json's north_star against it. Where the real reference finishes in seconds (C1, C2) it is
run live as well. Plus the size-independent properties: shards add up, ray statistics match, images are finite."""
import numpy as np
import pytest

import fray_b200 as fb
import oracle_util as ou

pytestmark = pytest.mark.gpu


def both(scene, **kw):
    out = {}
    for precision in (fb.FP64, fb.FP32):
        ctx = fb.GpuContext(scene, 0, precision)
        out[precision] = ctx.render(**kw)
        ctx.close()
    return out


def test_c1_boxed_full_size(data_dir):
    """configs[0]: data/boxed.fray as shipped (1024x768, Whitted, 2 RectLights x 16 samples per hit, three KD meshes)."""
    sc = fb.Scene(ou.override_scene("boxed", "full", dict(wantPrepass="off")))
    assert (sc.width, sc.height, sc.spp) == (1024, 768, 1)
    r = both(sc)
    (i64, s64), (i32, s32) = r[fb.FP64], r[fb.FP32]
    frac, rmse, mx = ou.compare(i64, i32, 1e-3)
    assert frac >= 0.999, (frac, rmse, mx)
    assert abs(s32.rays - s64.rays) <= 1e-3 * s64.rays
    assert 25e6 < s64.rays < 27e6 and s64.shadow_rays > 0.95 * s64.rays  # 25.85 M rays, 96.9 % shadow rays
    if ou.have_reference():
        ref, _ = ou.reference_render(sc.path, seed=42)
        assert ou.compare(ref, i64, 2e-5)[0] == 1.0
        assert ou.compare(ref, i32, 1e-3)[0] >= 0.999


def test_c2_zaphod_full_size(data_dir):
    """configs[1]: data/zaphod.fray as shipped (645x430, depth of field, 100 samples per pixel, bitmap texture)."""
    sc = fb.Scene(ou.override_scene("zaphod", "full", dict(wantPrepass="off")))
    assert (sc.width, sc.height, sc.spp) == (645, 430, 100)
    r = both(sc)
    (i64, s64), (i32, s32) = r[fb.FP64], r[fb.FP32]
    frac, rmse, mx = ou.compare(i64, i32, 1e-3)
    assert frac >= 0.995 and rmse < 1e-3, (frac, rmse, mx)
    assert s32.rays == s64.rays and s64.primary_rays == 645 * 430 * 100
    assert 55.4e6 < s64.rays < 55.5e6                            # 55.46 M rays measured on the reference
    if ou.have_reference():
        ref, _ = ou.reference_render(sc.path, seed=42)
        assert ou.compare(ref, i64, 2e-5)[0] == 1.0


def test_c4_smallpt_1024spp(data_dir):
    """configs[3]: data/smallpt.fray at 1024 paths per pixel; the sample split used for it on several GPUs adds up."""
    sc = fb.Scene(ou.override_scene("smallpt", "full1024", dict(pathsPerPixel=1024)))
    assert (sc.width, sc.height, sc.spp) == (640, 480, 1024)
    r = both(sc)
    (i64, s64), (i32, s32) = r[fb.FP64], r[fb.FP32]
    frac, rmse, mx = ou.compare(i64, i32, 1e-2)
    assert rmse < 5e-3 and frac >= 0.99, (frac, rmse, mx)      # same seeds: only paths that flip at a silhouette differ
    # the fast precision does not shoot the shadow rays whose BRDF factor is zero anyway (mirror / glass hits, light behind the
    # surface) and counts only what it traces; the closest-hit queries are the same
    assert s32.shadow_rays <= s64.shadow_rays
    assert abs((s32.rays - s32.shadow_rays) - (s64.rays - s64.shadow_rays)) <= 2e-3 * s64.rays
    assert 9.0 < s64.rays / s64.primary_rays < 11.5                 # 10.2 rays per path on the reference
    ctx = fb.GpuContext(sc, 0, fb.FP32)
    parts = [ctx.render(flags=fb.FRAME_SUM, sample_begin=k * 128, sample_end=(k + 1) * 128) for k in range(8)]
    np.testing.assert_allclose(sum(p[0] for p in parts) / 1024.0, i32, rtol=1e-4, atol=1e-5)
    assert sum(p[1].rays for p in parts) == s32.rays
    ctx.close()


def test_c5_forest_4k(data_dir):
    """configs[4]: data/forest.fray at 3840x2160 (beyond the reference's VFB_MAX_SIZE of 3000, src/constants.h:27), with the
    tile split over 8 shares adding up exactly."""
    sc = fb.Scene(ou.override_scene("forest", "4k", dict(frameWidth=3840, frameHeight=2160, interactive="off")))
    assert (sc.width, sc.height) == (3840, 2160)
    r = both(sc)
    (i64, s64), (i32, s32) = r[fb.FP64], r[fb.FP32]
    frac, rmse, mx = ou.compare(i64, i32, 1e-3)
    assert frac >= 0.999, (frac, rmse, mx)
    ctx = fb.GpuContext(sc, 0, fb.FP32)
    a32, _ = ctx.render(mode=fb.RENDER_AOV)
    parts = [ctx.render(flags=fb.FRAME_SUM, bucket_rank=k, bucket_count=8) for k in range(8)]
    assert np.array_equal(sum(p[0] for p in parts), i32)
    assert sum(p[1].rays for p in parts) == s32.rays
    ctx.close()
    ctx = fb.GpuContext(sc, 0, fb.FP64)
    a64, _ = ctx.render(mode=fb.RENDER_AOV)
    ctx.close()
    assert (a32[..., 0] == a64[..., 0]).mean() >= 0.9995          # primary hit ids, except silhouette ties
    assert abs(s32.rays - s64.rays) <= 1e-3 * s64.rays
