"""The flat polygon table of the fast-precision path (fray_b200/csrc/flat.cuh, built by scene_image.h) on a scene that
exercises what the bundled scenes do not: rotated, non-uniformly scaled and mirrored nodes over brute-force meshes, a two-sided
mesh, transformed planes (one textured: object-space uv), a translated textured sphere (sphere list, spherical uv) next to a
squashed one (node loop), a point light beside a rectangular light (per-light shadow sets).
The CPU half runs the very same device code compiled for the host (tests/emul); the GPU half goes through the C ABI."""
import os
import shutil

import numpy as np
import pytest

import fray_b200 as fb
import oracle_util as ou

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def scene(data_dir):
    dst = os.path.join(data_dir, "flat_transforms__test.fray")
    shutil.copyfile(os.path.join(HERE, "scenes", "flat_transforms.fray"), dst)
    return fb.Scene(dst)


def check(scene, render):
    want, ostats = ou.oracle_render(scene, seed=7)
    waov, _ = ou.oracle_render(scene, mode=fb.RENDER_AOV)
    # the scene really contains what it is meant to test
    nodes = set(np.unique(waov[..., 0]).astype(int))
    assert {0, 1, 2, 3, 4, 5, 6, 7} <= nodes, nodes
    got, stats = render(scene, fb.FP32, seed=7)
    frac, rmse, mx = ou.compare(want, got, 1e-3)
    assert frac >= 0.995 and rmse < 5e-3, (frac, rmse, mx)
    assert abs(stats.rays - ostats.rays) <= 0.002 * ostats.rays
    gaov, _ = render(scene, fb.FP32, mode=fb.RENDER_AOV)
    assert (gaov[..., 0] == waov[..., 0]).mean() >= 0.998
    hit = (waov[..., 0] >= 0) & (gaov[..., 0] == waov[..., 0])
    np.testing.assert_allclose(gaov[..., 2][hit], waov[..., 2][hit], rtol=2e-4)
    return got


def test_flat_table_on_the_host_emulator(scene):
    from test_emul_vs_oracle import emul as emul_fixture
    render = emul_fixture.__wrapped__()
    check(scene, render)
    # parity precision never uses the table and must agree with the oracle exactly as everywhere else
    want, ostats = ou.oracle_render(scene, seed=7)
    got, stats = render(scene, fb.FP64, seed=7)
    assert ou.compare(want, got, 1e-5)[0] == 1.0 and stats.rays == ostats.rays


@pytest.mark.gpu
def test_flat_table_on_the_gpu(scene):
    def render(sc, precision, **kw):
        ctx = fb.GpuContext(sc, 0, precision)
        out = ctx.render(**kw)
        ctx.close()
        return out
    check(scene, render)
    want, ostats = ou.oracle_render(scene, seed=7)
    got, stats = render(scene, fb.FP64, seed=7)
    assert ou.compare(want, got, 2e-5)[0] == 1.0 and stats.rays == ostats.rays


# ---- convex hexahedra with two cap planes: an open tube ---------------------------------------------------------------
OPEN_TUBE_OBJ = """# open square tube: four side faces seen from outside, no top, no bottom
v -1 0 -1
v  1 0 -1
v  1 0  1
v -1 0  1
v -1 1 -1
v  1 1 -1
v  1 1  1
v -1 1  1
f 1 5 6 2
f 2 6 7 3
f 3 7 8 4
f 4 8 5 1
"""


@pytest.fixture(scope="module")
def tube_scene(data_dir):
    with open(os.path.join(data_dir, "geom", "open_tube__test.obj"), "w") as f:
        f.write(OPEN_TUBE_OBJ)
    dst = os.path.join(data_dir, "open_tube__test.fray")
    shutil.copyfile(os.path.join(HERE, "scenes", "open_tube.fray"), dst)
    return fb.Scene(dst)


def check_tube(scene, render):
    want, ostats = ou.oracle_render(scene, seed=5)
    waov, _ = ou.oracle_render(scene, mode=fb.RENDER_AOV)
    assert {0, 1, 2} <= set(np.unique(waov[..., 0]).astype(int))
    # the camera does look into the standing tube: floor pixels surrounded by tube pixels
    got, stats = render(scene, fb.FP32, seed=5)
    frac, rmse, mx = ou.compare(want, got, 1e-3)
    assert frac >= 0.995 and rmse < 5e-3, (frac, rmse, mx)
    assert abs(stats.rays - ostats.rays) <= 0.002 * ostats.rays
    gaov, _ = render(scene, fb.FP32, mode=fb.RENDER_AOV)
    assert (gaov[..., 0] == waov[..., 0]).mean() >= 0.998


def test_open_tube_on_the_host_emulator(tube_scene, capfd):
    from test_emul_vs_oracle import emul as emul_fixture
    render = emul_fixture.__wrapped__()
    os.environ["FRAY_GPU_VERBOSE"] = "1"
    try:
        check_tube(tube_scene, render)
    finally:
        del os.environ["FRAY_GPU_VERBOSE"]
    assert "2 convex hexahedra" in capfd.readouterr().err  # both tubes became six-plane solids (four faces + two caps)


@pytest.mark.gpu
def test_open_tube_on_the_gpu(tube_scene):
    def render(sc, precision, **kw):
        ctx = fb.GpuContext(sc, 0, precision)
        out = ctx.render(**kw)
        ctx.close()
        return out
    check_tube(tube_scene, render)


# ---- degenerate inputs: nothing to hit, nothing to light with, frames that are not a multiple of the 8x4 tile ----
EMPTY = """
GlobalSettings {
	frameWidth 37
	frameHeight 21
	wantAA on
}
Camera camera {
	position (0, 1, -5)
	fov 60
}
"""

NO_LIGHTS_GI = """
GlobalSettings {
	frameWidth 13
	frameHeight 9
	gi on
	pathsPerPixel 6
	maxTraceDepth 3
}
Camera camera {
	position (0, 2, -6)
	pitch -10
	fov 60
}
Plane floor {
	y 0
	limit 30
}
Sphere ball {
	O (0, 1, 0)
	R 1
}
Lambert grey {
	color (0.6, 0.6, 0.6)
}
Refl mirror {
	multiplier 0.8
}
Node floor {
	geometry floor
	shader grey
}
Node ball {
	geometry ball
	shader mirror
}
"""


def _edge_cases(tmp_path, render, exact_tol):
    for name, text in (("empty", EMPTY), ("nolights", NO_LIGHTS_GI)):
        p = tmp_path / f"{name}.fray"
        p.write_text(text)
        sc = fb.Scene(str(p))
        want, ostats = ou.oracle_render(sc, seed=3)
        for precision, tol in ((fb.FP64, exact_tol), (fb.FP32, 1e-3)):
            got, stats = render(sc, precision, seed=3)
            assert got.shape == (sc.height, sc.width, 3) and np.isfinite(got).all()
            frac, rmse, mx = ou.compare(want, got, tol)
            assert frac >= (1.0 if precision == fb.FP64 else 0.98), (name, precision, frac, rmse, mx)
            assert stats.primary_rays == ostats.primary_rays == sc.width * sc.height * sc.spp
            if precision == fb.FP64:
                assert stats.rays == ostats.rays
        if name == "empty":
            assert not want.any()  # no environment: black (src/main.cpp:272-276)
        # shards of a frame whose size is not a multiple of the tile still add up
        full, fs = render(sc, fb.FP32, seed=3, flags=fb.FRAME_SUM)
        parts = [render(sc, fb.FP32, seed=3, flags=fb.FRAME_SUM, bucket_rank=r, bucket_count=3) for r in range(3)]
        assert np.array_equal(sum(q[0] for q in parts), full) and sum(q[1].rays for q in parts) == fs.rays


def test_edge_cases_on_the_host_emulator(tmp_path):
    from test_emul_vs_oracle import emul as emul_fixture
    _edge_cases(tmp_path, emul_fixture.__wrapped__(), 1e-5)


@pytest.mark.gpu
def test_edge_cases_on_the_gpu(tmp_path):
    def render(sc, precision, **kw):
        ctx = fb.GpuContext(sc, 0, precision)
        out = ctx.render(**kw)
        ctx.close()
        return out
    _edge_cases(tmp_path, render, 2e-5)
