"""pytest configuration: the `gpu` marker, import paths, and fixtures that build / locate the native pieces."""
import json
import os
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
for p in (ROOT, HERE):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")


@pytest.fixture(scope="session", autouse=True)
def native_build():
    """Host library and the oracle are plain g++ builds (seconds); the CUDA library is built by __graft_entry__.build()."""
    import fray_b200.build as fbuild
    fbuild.build_host()
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "oracle"])


@pytest.fixture(scope="session")
def golden_cases():
    with open(os.path.join(HERE, "golden", "cases.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def data_dir():
    import oracle_util as ou
    if not os.path.isdir(ou.DATA_DIR):
        pytest.skip(f"scene data mirror {ou.DATA_DIR} not present (run `make -C oracle data` where /root/reference exists)")
    return ou.DATA_DIR


def load_golden(name):
    import numpy as np
    z = np.load(os.path.join(HERE, "golden", name + ".npz"))
    return z["rgb"], z["node"].astype(int), z["dist"]


def golden_scene(cases, name):
    """(path of the override .fray for golden case `name`, seed)."""
    import oracle_util as ou
    c = cases[name]
    return ou.override_scene(c["scene"], "golden_" + name, c["settings"], c["camera"]), c["seed"]


def local_scene(name):
    """tests/scenes/<name>.fray copied into the data mirror (asset paths are relative to it); the path of the copy."""
    import shutil
    import oracle_util as ou
    dst = os.path.join(ou.DATA_DIR, name + "__test.fray")
    shutil.copyfile(os.path.join(HERE, "scenes", name + ".fray"), dst)
    return dst
