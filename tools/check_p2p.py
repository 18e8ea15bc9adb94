#!/usr/bin/env python3
"""Multi-GPU check (run under torchrun on >= 2 GPUs): the three decompositions of fray_b200.dist give the same frame.

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/check_p2p.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import torch.distributed as dist

import fray_b200 as fb
import fray_b200.dist as fdist
from fray_b200 import scenes


def main():
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    rank, world = dist.get_rank(), dist.get_world_size()
    scene = fb.Scene(scenes.override_scene("cornell_box", "p2pcheck", dict(frameWidth=200, frameHeight=152, pathsPerPixel=16)))
    frames = {}
    for mode in ("tiles", "p2p", "samples"):
        r = fdist.DistributedRenderer(scene, mode=mode, device=local)
        for _ in range(3):  # repeated frames: the peer frame is overwritten in place
            out = r.render()
        if rank == 0:
            frames[mode] = (r.mode, out.copy())
        dist.barrier()
        r.close()
    if rank == 0:
        single = fb.GpuContext(scene, local, fb.FP32)
        want, _ = single.render()
        single.close()
        for mode, (used, img) in frames.items():
            d = np.abs(img - want).max()
            print(f"{mode:8s} (ran as {used:7s}) max |difference| to the single-GPU frame: {d:.3e}")
            assert d <= (1e-5 if mode == "samples" else 0.0) or (mode != "samples" and d < 1e-6), (mode, d)
    # a moving camera in p2p mode: every frame differs, rank 0 copies frame k to the host while the peers already store frame
    # k+1 into the other half of the exported pair -- each frame must equal the single-GPU frame of the same camera
    r = fdist.DistributedRenderer(scene, mode="p2p", device=local)
    single = fb.GpuContext(scene, local, fb.FP32) if rank == 0 else None
    worst = 0.0
    for k in range(12):
        cam = scene.move_camera(dx=0.0 if k == 0 else 3.0, dz=0.0 if k == 0 else 1.5, dyaw=0.0 if k == 0 else 2.0)  # same parse on every rank: same cameras
        r.ctx.update_camera(cam)
        out = r.render()
        if rank == 0:
            single.update_camera(cam)
            want, _ = single.render()
            worst = max(worst, float(np.abs(out - want).max()))
    if rank == 0:
        single.close()
        print(f"p2p, 12 frames with a moving camera (ran as {r.mode}): max |difference| to the single-GPU frames: {worst:.3e}")
        assert worst < 1e-6, worst
    dist.barrier()
    r.close()
    if rank == 0:
        print(f"check_p2p: OK on {world} GPUs")
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
