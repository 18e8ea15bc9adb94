#!/usr/bin/env python3
"""Regenerate tests/golden/*.npz from the REAL reference code.

Runs oracle/_ref/fray_ref_ctr (the unmodified reference translation units + the counter-RNG contract, see
oracle/Makefile) on reduced-size variants of every bundled scene and stores the float32 image plus the primary-hit
AOV (node index, world distance). Needs /root/reference to have been built into oracle/_ref (build container only);
the .npz files and cases.json are committed so that the tests run anywhere.

    python tests/golden/make_goldens.py
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))
import numpy as np

import oracle_util as ou

CASES = {
    # name: (scene, GlobalSettings overrides, Camera overrides)
    "cornell_box": ("cornell_box", dict(frameWidth=64, frameHeight=64, pathsPerPixel=8), None),
    "smallpt": ("smallpt", dict(frameWidth=80, frameHeight=60, pathsPerPixel=8), None),
    "boxed": ("boxed", dict(frameWidth=96, frameHeight=72), None),
    "zaphod": ("zaphod", dict(frameWidth=96, frameHeight=64), dict(numSamples=6)),
    "forest": ("forest", dict(frameWidth=128, frameHeight=96, interactive="off"), None),
    "forest_aa": ("forest", dict(frameWidth=64, frameHeight=48, interactive="off", wantAA="on"), None),
    "forest_stereo_dof": ("forest", dict(frameWidth=64, frameHeight=48, interactive="off"), dict(stereoSeparation=0.25, dof="on", numSamples=3)),
    "axe_test": ("hw9/axe_test", dict(frameWidth=96, frameHeight=72), None),
    "nonconvex": ("hw9/nonconvex", dict(frameWidth=96, frameHeight=72), None),
    "bokeh": ("hw10/bokeh", dict(frameWidth=64, frameHeight=48), dict(numSamples=3)),
    "sphtri": ("hw12/sphtri", dict(frameWidth=64, frameHeight=48, pathsPerPixel=8), None),
    "dragon": ("hw9/dragon", dict(frameWidth=64, frameHeight=48), None),
}

# Parity tier T0: the UNMODIFIED reference binary, main.cpp's own RendMT::entry loop included, on the scenes
# that draw no random numbers (point lights only, fixed anti-aliasing table). Stored as t0_<name>.npz; must equal the
# fray_ref_ctr image of the same case bit for bit, which pins ctr_driver.cpp's restatement of the sample loop.
T0_CASES = ["forest", "forest_aa", "axe_test", "nonconvex"]

# Convergence (north_star: "both must converge to a high-spp ground truth"): cornell_box at 100x100, 2048 paths per pixel,
# rendered by the reference code under ANOTHER seed -> cornell_truth.npz, together with the reference's own RMSE against it at
# 64 and 256 paths per pixel under the test seed (what the GPU must reproduce: same seeds, same samples).
TRUTH = dict(scene="cornell_box", size=100, spp=2048, seed=7, test_seed=42, test_spp=(64, 256))

# scenes of our own (tests/scenes/*.fray, copied into the data mirror as <name>__test.fray): rendered as they are
LOCAL_CASES = ["csg_layered"]


def local_scene(name: str) -> str:
    """Copy tests/scenes/<name>.fray into the data mirror (asset paths are relative to it) and return the copy's path."""
    import shutil
    dst = os.path.join(ou.DATA_DIR, name + "__test.fray")
    shutil.copyfile(os.path.join(os.path.dirname(HERE), "scenes", name + ".fray"), dst)
    return dst


def truth_scene(spp: int) -> str:
    return ou.override_scene(TRUTH["scene"], f"truth{spp}", dict(frameWidth=TRUTH["size"], frameHeight=TRUTH["size"], pathsPerPixel=spp))


def make_truth():
    truth, sec = ou.reference_render(truth_scene(TRUTH["spp"]), seed=TRUTH["seed"])
    out = dict(rgb=truth, spp=TRUTH["spp"], seed=TRUTH["seed"], test_seed=TRUTH["test_seed"])
    for spp in TRUTH["test_spp"]:
        img, _ = ou.reference_render(truth_scene(spp), seed=TRUTH["test_seed"])
        out[f"ref_rmse_{spp}"] = ou.compare(truth, img)[1]
    np.savez_compressed(os.path.join(HERE, "cornell_truth.npz"), **out)
    print(f"cornell_truth        {truth.shape[1]}x{truth.shape[0]} at {TRUTH['spp']} spp ({sec:.1f}s); reference RMSE at "
          + ", ".join(f"{spp} spp: {out[f'ref_rmse_{spp}']:.5f}" for spp in TRUTH["test_spp"]))


def main():
    if not ou.have_reference():
        sys.exit("oracle/_ref is not built: run `make -C oracle ref` where /root/reference exists")
    manifest = {}
    for name, (scene, st, cam) in CASES.items():
        f = ou.override_scene(scene, "golden_" + name, st, cam)
        rgb, sec, node, dist = ou.reference_render(f, seed=42, aov=True)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), rgb=rgb, node=node.astype(np.int16), dist=np.minimum(dist, 3e38).astype(np.float32))
        manifest[name] = dict(scene=scene, settings=st, camera=cam, seed=42)
        print(f"{name:20s} {rgb.shape[1]}x{rgb.shape[0]} mean {rgb.mean():.5f} ({sec:.2f}s)")
    for name in T0_CASES:
        scene, st, cam = CASES[name]
        rgb = ou.reference_render_unmodified(ou.override_scene(scene, "golden_" + name, st, cam))
        ctr = np.load(os.path.join(HERE, name + ".npz"))["rgb"]
        assert np.array_equal(rgb, ctr), f"{name}: RendMT::entry and ctr_driver.cpp disagree"
        np.savez_compressed(os.path.join(HERE, "t0_" + name + ".npz"), rgb=rgb)
        print(f"t0_{name:17s} {rgb.shape[1]}x{rgb.shape[0]} mean {rgb.mean():.5f} == fray_ref_ctr")
    make_truth()
    for name in LOCAL_CASES:
        rgb, sec, node, dist = ou.reference_render(local_scene(name), seed=42, aov=True)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), rgb=rgb, node=node.astype(np.int16), dist=np.minimum(dist, 3e38).astype(np.float32))
        print(f"{name:20s} {rgb.shape[1]}x{rgb.shape[0]} mean {rgb.mean():.5f} ({sec:.2f}s)")
    with open(os.path.join(HERE, "cases.json"), "w") as fp:
        json.dump(manifest, fp, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
