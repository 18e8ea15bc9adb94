"""Where the bundled scene files live and how to derive variants of them (benchmarks and tests change sample counts and
frame sizes without touching the originals).

The scene files and their assets are the reference's `data/` directory. It is not copied into git (GPL, and not ours): the
build (`__graft_entry__.build()` -> `make -C oracle data`) mirrors it to `oracle/_ref/data/`, which travels to the GPU box
with the snapshot. `FRAY_DATA` overrides the location.
"""
from __future__ import annotations

import os
import re

REPO_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DATA_DIR = os.environ.get("FRAY_DATA", os.path.join(REPO_ROOT, "oracle", "_ref", "data"))

_BLOCK_RE = {"GlobalSettings": re.compile(r"^\s*GlobalSettings\b[^{]*\{", re.M), "Camera": re.compile(r"^\s*Camera\b[^{]*\{", re.M)}


def scene_path(name: str) -> str:
    """'cornell_box' or 'hw9/dragon' -> path of the bundled .fray file in the data mirror."""
    return os.path.join(DATA_DIR, name + ".fray")


def override_scene(name: str, tag: str, settings: dict | None = None, camera: dict | None = None) -> str:
    """Write `<name>__<tag>.fray` beside the original (asset paths are relative to the scene file,
    /root/reference/src/scene.cpp:710-721) with the given properties put FIRST in the block, so they win
    (ParsedBlockImpl::findProperty returns the first match, src/scene.cpp:112-122)."""
    src = scene_path(name)
    text = open(src).read()
    for block, props in (("GlobalSettings", settings), ("Camera", camera)):
        if not props:
            continue
        m = _BLOCK_RE[block].search(text)
        if not m:
            raise RuntimeError(f"{src} has no {block} block")
        ins = "".join(f"\n\t{k} {v}" for k, v in props.items())
        text = text[:m.end()] + ins + text[m.end():]
    dst = os.path.join(os.path.dirname(src), f"{os.path.basename(name)}__{tag}.fray")
    # several ranks of one torchrun job ask for the same file at the same time: never expose a half-written one
    if os.path.exists(dst):
        with open(dst) as f:
            if f.read() == text:
                return dst
    tmp = f"{dst}.{os.getpid()}.tmp"
    with open(tmp, "w") as f:
        f.write(text)
    os.replace(tmp, dst)
    return dst
