"""The counter-based RNG contract: published Philox known answers, and product (csrc/rng.cuh) == oracle (fray_rng.h)."""
import ctypes as C

import numpy as np
import pytest

import fray_b200 as fb
import oracle_util as ou


def _philox(ctr, key):
    out = (C.c_uint32 * 4)()
    ou.oracle_lib().fray_oracle_philox((C.c_uint32 * 4)(*ctr), (C.c_uint32 * 2)(*key), out)
    return list(out)


def test_philox_known_answers():
    # Random123 kat_vectors, philox4x32 10 rounds
    assert _philox([0, 0, 0, 0], [0, 0]) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert _philox([0xffffffff] * 4, [0xffffffff] * 2) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert _philox([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0]) == [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def _draws(lib_fn, seed, pixel, sample, branch, n):
    out = np.zeros(n, dtype=np.uint32)
    lib_fn(seed, pixel, sample, branch, n, out.ctypes.data)
    return out


def test_product_rng_equals_oracle_rng():
    host, oracle = fb.host_lib(), ou.oracle_lib()
    rs = np.random.RandomState(7)
    for _ in range(50):
        seed, pixel, sample, branch = (int(v) for v in rs.randint(0, 2 ** 32, size=4, dtype=np.uint64))
        n = int(rs.randint(1, 70))
        a = _draws(host.fray_host_rng_draws, seed, pixel, sample, branch, n)
        b = _draws(oracle.fray_oracle_rng_draws, seed, pixel, sample, branch, n)
        assert np.array_equal(a, b)
        d, k = int(rs.randint(0, 1000)), int(rs.randint(0, 40))
        assert host.fray_host_rng_child(branch, d, k) == oracle.fray_oracle_rng_child(branch, d, k)
        assert host.fray_host_rng_child(branch, d, k) != 0  # branch 0 is reserved for the primary stream


def test_streams_differ_by_pixel_and_sample():
    host = fb.host_lib()
    base = _draws(host.fray_host_rng_draws, 42, 10, 3, 0, 16)
    for args in ((42, 11, 3, 0), (42, 10, 4, 0), (43, 10, 3, 0), (42, 10, 3, 5)):
        assert not np.array_equal(base, _draws(host.fray_host_rng_draws, *args, 16))
    # draw i is word (i & 3) of block (i >> 2): a longer request starts with the shorter one
    assert np.array_equal(base[:7], _draws(host.fray_host_rng_draws, 42, 10, 3, 0, 7))


@pytest.mark.gpu
@pytest.mark.parametrize("mode", [0, 1, 2])
def test_device_streams_equal_the_contract(mode):
    """The forms the kernels generate their streams in -- on-demand blocks with round keys from the kernel parameters, and the
    path tracer's shared-memory ring, sequentially and under pathSegment's skip / draw pattern -- against the oracle's RNG."""
    gpu, oracle = fb.gpu_lib(), ou.oracle_lib()
    gpu.fray_gpu_rng_probe.argtypes = [C.c_int] + [C.c_uint32] * 4 + [C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    rs = np.random.RandomState(11 + mode)
    for _ in range(12):
        seed, pixel, sample, branch = (int(v) for v in rs.randint(0, 2 ** 32, size=4, dtype=np.uint64))
        n = int(rs.randint(1, 200))
        out, drawn = np.zeros(n, dtype=np.uint32), np.zeros(n, dtype=np.uint8)
        assert gpu.fray_gpu_rng_probe(0, seed, pixel, sample, branch, mode, n, out.ctypes.data, drawn.ctypes.data) == 0
        want = _draws(oracle.fray_oracle_rng_draws, seed, pixel, sample, branch, n)
        got = drawn.astype(bool)
        if mode < 2:
            assert got.all()
        else:  # 2 draws, then 6 of every 8
            assert got.sum() == min(n, 2) + 6 * ((n - 2) // 8 if n >= 2 else 0)
        assert np.array_equal(out[got], want[got])
