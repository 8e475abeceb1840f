"""Row T, host side: avl_host_gather (native staging of per-env observation arrays, common/utils.py:129-156) — pure host
code of the C-ABI library, so it is exercised on the CPU: ragged pieces, every thread count, the batch_obs path that uses
it and the numpy fallback for pieces the native path does not take."""
import ctypes

import numpy as np
import pytest
import torch

from avlen_b200 import _lib
from avlen_b200.common import utils as U


def _gather(srcs, dsts):
    n = len(srcs)
    src = (ctypes.c_void_p * n)(*[a.ctypes.data for a in srcs])
    dst = (ctypes.c_void_p * n)(*[a.ctypes.data for a in dsts])
    nb = (ctypes.c_longlong * n)(*[a.nbytes for a in srcs])
    return _lib.lib().avl_host_gather(src, dst, nb, n)


@pytest.mark.parametrize("threads", [0, 1, 2, 3, 8])
def test_gather_ragged_pieces(threads):
    rng = np.random.default_rng(threads)
    sizes = [0, 1, 4095, 4096, 4097, 300000, 7, 1 << 20, 123457, 0, 65536]
    srcs = [rng.integers(0, 256, size=s, dtype=np.uint8) for s in sizes]
    dsts = [np.zeros(s, np.uint8) for s in sizes]
    old = _lib.lib().avl_set_host_gather_threads(threads)
    try:
        assert _gather(srcs, dsts) == 0
    finally:
        _lib.lib().avl_set_host_gather_threads(old)
    for a, b in zip(srcs, dsts):
        assert np.array_equal(a, b)


def test_gather_argument_errors():
    assert _lib.lib().avl_host_gather(None, None, None, 0) == 0
    assert _lib.lib().avl_host_gather(None, None, None, 3) == -1
    assert _lib.lib().avl_host_gather(None, None, None, -1) == -1


def test_batch_obs_stacks_through_the_native_gather_and_falls_back():
    rng = np.random.default_rng(0)
    n = 16
    obs = [{"rgb": rng.integers(0, 256, size=(128, 128, 3), dtype=np.uint8),
            "depth": rng.random((128, 128, 1), dtype=np.float32),
            "pose": rng.random(4).astype(np.float32),
            "odd": np.asfortranarray(rng.random((200, 200)).astype(np.float32))} for _ in range(n)]
    for sensor, shape, dt in (("rgb", (128, 128, 3), np.uint8), ("depth", (128, 128, 1), np.float32), ("pose", (4,), np.float32),
                              ("odd", (200, 200), np.float32)):
        out = np.empty((n,) + shape, dt)
        U._stack_into(obs, sensor, out)
        assert np.array_equal(out, np.stack([o[sensor] for o in obs])), sensor
    b = U.batch_obs(obs, device=torch.device("cpu"))
    assert b["rgb"].dtype == torch.float32 and torch.equal(b["rgb"], torch.from_numpy(np.stack([o["rgb"] for o in obs])).float())
