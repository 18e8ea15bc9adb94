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
    with open(os.path.join(HERE, "cases.json"), "w") as fp:
        json.dump(manifest, fp, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
