"""GPU: the lossless int16 transport of the host API (hd_pack_i16 + hd_host_widen_i16) through the C ABI."""
import ctypes

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    from hydrodem_b200 import _lib, device as dev


def _pack(arr):
    lib = _lib.load()
    r = dev.upload(arr)
    ny, nx = arr.shape
    dense = torch.empty(ny * nx, dtype=torch.int16, device="cuda")
    flag = torch.full((1,), 7, dtype=torch.int32, device="cuda")
    _lib.check(lib.hd_pack_i16(r.ptr, r.pitch, ctypes.c_void_p(dense.data_ptr()), ny, nx, ctypes.c_void_p(flag.data_ptr()),
                               dev.stream_ptr()))
    torch.cuda.synchronize()
    return dense.cpu().numpy().reshape(ny, nx), int(flag.item())


@pytest.mark.parametrize("shape", [(1, 1), (5, 3), (64, 128), (257, 1001)])
def test_pack_i16_exact_and_flag_clear(shape):
    rng = np.random.default_rng(3)
    a = rng.integers(-32768, 32768, shape).astype(np.float32)
    a.flat[0], a.flat[-1] = -32768.0, 32767.0
    got, flag = _pack(a)
    assert flag == 0
    np.testing.assert_array_equal(got, a.astype(np.int16))


@pytest.mark.parametrize("bad", [np.nan, 0.5, 32768.0, -32769.0, np.inf])
def test_pack_i16_flags_unrepresentable_values(bad):
    a = np.zeros((33, 70), dtype=np.float32)
    a[17, 69] = bad
    _, flag = _pack(a)
    assert flag == 1
