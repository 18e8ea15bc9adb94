"""The multi-GPU decomposition on CPU: world_size 2 over gloo. The per-rank partial frames come from the oracle with the
same shard arguments the GPU ranks use (fray_b200.dist.shard); the reduce + resolve plumbing is the product's."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import fray_b200 as fb
import fray_b200.dist as fdist
import oracle_util as ou
from conftest import golden_scene


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, path, mode, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        sc = fb.Scene(path)
        kw = fdist.shard(rank, world, sc.spp, mode)
        part, _ = ou.oracle_render(sc, threads=2, flags=fb.FRAME_SUM, **kw)
        res = fdist.reduce_partials(torch.from_numpy(part), sc.spp)
        if rank == 0:
            np.save(os.path.join(out_dir, f"{mode}.npy"), res.numpy())
        else:
            assert res is None
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("mode", ["tiles", "samples"])
def test_two_ranks_reproduce_the_frame(mode, golden_cases, data_dir, tmp_path):
    path, _ = golden_scene(golden_cases, "cornell_box")
    port = _free_port()
    mp.spawn(_worker, args=(2, port, path, mode, str(tmp_path)), nprocs=2, join=True)
    got = np.load(tmp_path / f"{mode}.npy")
    full, _ = ou.oracle_render(fb.Scene(path))
    if mode == "tiles":
        assert np.array_equal(got, full)
    else:
        np.testing.assert_allclose(got, full, rtol=2e-6, atol=1e-7)


def test_shard_arguments():
    assert fdist.shard(0, 1, 40, "tiles") == {}
    assert fdist.shard(3, 8, 256, "samples") == dict(sample_begin=96, sample_end=128)
    assert fdist.shard(1, 4, 5, "tiles") == dict(bucket_rank=1, bucket_count=4)
    covered = []
    for r in range(3):
        k = fdist.shard(r, 3, 40, "samples")
        covered += list(range(k["sample_begin"], k["sample_end"]))
    assert covered == list(range(40))
    with pytest.raises(ValueError):
        fdist.shard(0, 8, 5, "samples")
    assert fdist.choose_mode(256, 8) == "samples" and fdist.choose_mode(1, 8) == "tiles"
