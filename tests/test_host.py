"""Host scene layer: the `.fray` language, OBJ loading, the parity-exact KD builder, image IO (no GPU needed)."""
import os
import struct

import numpy as np
import pytest

import fray_b200 as fb
import oracle_util as ou


def write(tmp_path, name, text):
    p = tmp_path / name
    p.write_text(text)
    return str(p)


MINIMAL = """
// line comment
GlobalSettings {
	frameWidth 32   # trailing comment
	frameHeight 24
	wantAA off
}
/* block comment
Camera ignored { }
*/
Camera camera {
	position (0, 1, -5)
	fov 60
}
Plane floor {
	y 0
}
Lambert gray {
	color (0.5 0.5 0.5)
}
Node n {
	geometry floor
	shader gray
}
PointLight {
	pos (0, 10, 0)
	power 100
}
"""


def test_parse_minimal_scene(tmp_path):
    sc = fb.Scene(write(tmp_path, "a.fray", MINIMAL))
    assert (sc.width, sc.height, sc.spp) == (32, 24, 1)
    head = sc.head
    assert head.num_nodes == 1 and head.settings.want_aa == 0
    cam = head.camera
    assert tuple(cam.pos) == (0.0, 1.0, -5.0) and cam.w == 32 and cam.h == 24
    # no rotation: the screen is symmetric around the +z axis (camera.cpp:34-57)
    assert cam.top_left[0] == -cam.top_right[0] and cam.top_left[1] == cam.top_right[1] and cam.top_left[2] == 1.0


def test_samples_per_pixel_rule(tmp_path):
    # src/main.cpp:395-400
    assert fb.Scene(write(tmp_path, "a.fray", MINIMAL.replace("wantAA off", "wantAA on"))).spp == 5
    assert fb.Scene(write(tmp_path, "b.fray", MINIMAL.replace("wantAA off", "gi on\n\tpathsPerPixel 33"))).spp == 33
    assert fb.Scene(write(tmp_path, "c.fray", MINIMAL.replace("fov 60", "fov 60\n\tdof on\n\tnumSamples 3"))).spp == 3  # max(1, 3)
    assert fb.Scene(write(tmp_path, "e.fray", MINIMAL.replace("wantAA off", "wantAA on").replace("fov 60", "fov 60\n\tdof on\n\tnumSamples 3"))).spp == 5
    sc = fb.Scene(write(tmp_path, "d.fray", MINIMAL))
    assert sc.set(gi=1, pathsPerPixel=17).spp == 17


@pytest.mark.parametrize("bad, msg", [
    (MINIMAL.replace("geometry floor", "geometry nothere"), "Geometry not defined"),
    (MINIMAL.replace("Plane floor {", "Plank floor {"), "Unknown object class"),
    (MINIMAL.replace("position (0, 1, -5)", ""), "Required property"),
    (MINIMAL.replace("fov 60", "fov 500"), "outside the allowed bounds"),
    (MINIMAL + "\nNode open {\n", "Unfinished object"),
])
def test_parse_errors(tmp_path, bad, msg):
    with pytest.raises(fb.FrayError, match=msg):
        fb.Scene(write(tmp_path, "bad.fray", bad))


def test_missing_file(tmp_path):
    with pytest.raises(fb.FrayError, match="Cannot open"):
        fb.Scene(str(tmp_path / "nope.fray"))
    with pytest.raises(fb.FrayError, match="Required file not found"):
        fb.Scene(write(tmp_path, "m.fray", MINIMAL + '\nMesh m {\n\tfile "nope.obj"\n}\n'))


def test_random_macros_are_deterministic(tmp_path):
    text = MINIMAL.replace("power 100", "power randfloat(50, 60)")
    a = fb.Scene(write(tmp_path, "r1.fray", text))
    b = fb.Scene(write(tmp_path, "r2.fray", text))
    assert bytes(a.head.camera) == bytes(b.head.camera)


def test_obj_loader_and_small_mesh(tmp_path):
    # a quad as one 4-gon: fan-triangulated into 2 triangles, no normals => faceted, <= 20 triangles => no KD tree
    (tmp_path / "q.obj").write_text("# quad\r\nv -1 0 -1 \r\nv 1 0 -1\r\nv 1 0 1\r\nv -1 0 1\r\nvt 0 0\r\nvt 1 0\r\nvt 1 1\r\nvt 0 1\r\nf 1/1 2/2 3/3 4/4\r\n")
    sc = fb.Scene(write(tmp_path, "q.fray", MINIMAL.replace("Plane floor {\n\ty 0", 'Mesh floor {\n\tfile "q.obj"')))
    st = sc.mesh_stats()
    assert st == [dict(nodes=0, leaf_refs=0, max_depth=0, triangles=2)]


def test_kd_builder_matches_reference_statistics(data_dir):
    """Node / leaf-reference counts and depths printed by the reference's own builder."""
    expect = {
        "boxed": [(5491, 32578, 65, 9120), (1383, 5528, 65, 1140), (13, 106, 3, 44)],
        "hw9/axe_test": [(3033, 15814, 65, 552), (1383, 5528, 65, 1140)],
        "hw10/bokeh": [(7053, 39924, 26, 9759)],
    }
    for name, meshes in expect.items():
        st = fb.Scene(ou.scene_path(name)).mesh_stats()
        assert [(m["nodes"], m["leaf_refs"], m["max_depth"], m["triangles"]) for m in st] == meshes


def test_bmp_writer_is_byte_exact(tmp_path):
    """54-byte header, 24-bit BGR, bottom-up, rows padded to 4 bytes, floor(clamp01(x)*255 + 0.5) (src/bitmap.cpp:197-236)."""
    rgb = np.zeros((2, 3, 3), np.float32)
    rgb[0, 0] = (1.0, 0.5, 0.0)
    rgb[1, 2] = (2.0, -1.0, 0.25)
    p = str(tmp_path / "o.bmp")
    fb.save_image(p, rgb)
    raw = open(p, "rb").read()
    row = 3 * 3 + 3  # padded to 12
    assert len(raw) == 54 + 2 * row
    assert raw[:2] == b"BM" and struct.unpack("<iii", raw[2:14]) == (54 + 2 * row, 0, 54)
    assert struct.unpack("<iiiHHiiiiii", raw[14:54]) == (40, 3, 2, 1, 24, 0, 0, 0, 0, 0, 0)
    bottom, top = raw[54:54 + row], raw[54 + row:54 + 2 * row]
    assert top[:3] == bytes([0, 128, 255])          # (1.0, 0.5, 0.0) as B, G, R; 0.5*255+0.5 = 128
    assert bottom[6:9] == bytes([64, 0, 255])       # (2.0, -1.0, 0.25) clamped
    back = fb.load_image(p)
    assert back.shape == (2, 3, 3) and abs(back[0, 0, 0] - 1.0) < 1e-6 and abs(back[0, 0, 1] - 128 / 255) < 1e-6


def test_bmp_reader_on_bundled_textures(data_dir):
    for name, shape in (("texture/zaphod.bmp", (768, 768)), ("texture/zar-bump.bmp", (128, 768)), ("texture/lava.bmp", (512, 512))):
        im = fb.load_image(os.path.join(data_dir, name))
        assert im.shape[:2] == shape and 0.0 <= im.min() and im.max() <= 1.0 and im.std() > 0.01
        # every channel value is k/255 (src/bitmap.cpp:186-188)
        assert np.abs(im * 255 - np.round(im * 255)).max() < 1e-4


def test_exr_piz_decoder_matches_openexr(data_dir):
    """Our PIZ/HALF decoder vs. OpenCV's bundled OpenEXR decode of the same files (side-cars written by oracle/mirror_data.py)."""
    folder = os.path.join(data_dir, "env", "forest")
    for face in ("negx", "negy", "negz", "posx", "posy", "posz"):
        path = os.path.join(folder, face + ".exr")
        side = path + ".f32"
        if not os.path.exists(side):
            pytest.skip("no EXR side-cars")
        with open(side, "rb") as f:
            w, h = np.fromfile(f, np.int32, 2)
            want = np.fromfile(f, np.float32).reshape(h, w, 4)[..., :3]
        got = fb.load_image(path)
        assert got.shape == (256, 256, 3)
        assert np.array_equal(got, want)


def test_exr_writer_roundtrip(tmp_path):
    rs = np.random.RandomState(1)
    rgb = (rs.rand(5, 7, 3) * 40).astype(np.float32)
    p = str(tmp_path / "o.exr")
    fb.save_image(p, rgb)
    back = fb.load_image(p)
    assert np.array_equal(back, rgb.astype(np.float16).astype(np.float32))  # HALF storage, round to nearest even


# ---- the `fray` command line tool (fray_b200/host/main.cpp), the parts that need no GPU ----
def _cli():
    import fray_b200.build as fbuild
    fbuild.build_host()
    exe = fbuild.build_cli()
    assert exe and os.path.exists(exe)
    return exe


def test_cli_exit_codes_follow_the_reference(tmp_path, data_dir):
    """src/main.cpp:494-506: -1 usage error, -2 missing scene file, -3 parse error (as 8-bit exit statuses)."""
    import subprocess
    exe = _cli()
    scene = os.path.join(data_dir, "cornell_box.fray")
    r = subprocess.run([exe, scene], capture_output=True, text=True)          # no --gpu: this build has no CPU renderer
    assert r.returncode == 255 and "--gpu" in r.stderr
    r = subprocess.run([exe, "--gpu", "--bogus", scene], capture_output=True, text=True)
    assert r.returncode == 255 and "Usage" in r.stderr
    r = subprocess.run([exe, "--gpu", str(tmp_path / "missing.fray")], capture_output=True, text=True)
    assert r.returncode == 254 and "does not exist" in r.stderr
    bad = write(tmp_path, "bad.fray", "Camera camera {\n\tposition (0, 0, 0)\n")   # unterminated block
    r = subprocess.run([exe, "--gpu", bad], capture_output=True, text=True)
    assert r.returncode == 253
    if fb.gpu_lib().fray_gpu_device_count() == 0:
        r = subprocess.run([exe, "--gpu", scene], capture_output=True, text=True)
        assert r.returncode == 252 and "no CPU fallback" in r.stderr


def test_exr_writer_is_piz_and_openexr_reads_it(tmp_path):
    """The writer emits what Imf::RgbaOutputFile writes by default (/root/reference/src/bitmap.cpp:266-284): HALF A,B,G,R,
    PIZ-compressed chunks of 32 scan lines. Decoded here by OUR reader and by OpenCV's bundled OpenEXR library; sizes that
    exercise odd widths / heights of the wavelet, a partial last chunk, flat regions (run-length symbol) and the 16-bit wavelet
    mode (more than 2^14 distinct values)."""
    os.environ["OPENCV_IO_ENABLE_OPENEXR"] = "1"
    cv2 = pytest.importorskip("cv2")
    rs = np.random.RandomState(7)
    for h, w, scale in ((70, 90, 4.0), (33, 17, 1.0), (1, 1, 1.0), (64, 64, 1e4), (150, 333, 60000.0)):
        rgb = (rs.rand(h, w, 3) * scale).astype(np.float32)
        rgb[: h // 2] = 0.25
        rgb[0, 0] = [0.0, 1e-3, 1000.0]
        p = str(tmp_path / f"piz_{h}x{w}.exr")
        fb.save_image(p, rgb)
        want = rgb.astype(np.float16).astype(np.float32)
        with open(p, "rb") as f:
            head = f.read(400)
        i = head.index(b"compression\0compression\0")
        assert head[i + 28] == 4  # PIZ_COMPRESSION
        assert np.array_equal(fb.load_image(p), want)
        other = cv2.imread(p, cv2.IMREAD_UNCHANGED)
        if other is None:
            pytest.skip("this OpenCV build has no OpenEXR codec")
        assert other.shape == (h, w, 4) and np.all(other[..., 3] == 1)
        assert np.array_equal(other[..., :3][..., ::-1], want)
    yy, xx = np.mgrid[0:256, 0:256].astype(np.float32)
    smooth = np.stack([xx / 256, yy / 256, (xx + yy) / 512], -1).astype(np.float32)
    p = str(tmp_path / "smooth.exr")
    fb.save_image(p, smooth)
    assert os.path.getsize(p) < 256 * 256 * 8 // 4  # it does compress
