"""The device sort / scan primitives of the query pipeline (csrc/engine.cu: sort_pairs, exclusive_scan_u64):
small inputs take single-CTA shared-memory kernels, larger ones the CUB pipelines. Both must be a STABLE sort
on the selected key bits / an exact exclusive sum — at every size around the switch-over points."""
import ctypes as C

import numpy as np
import pytest

import impg_b200 as ix

pytestmark = pytest.mark.gpu

SIZES = [1, 2, 3, 31, 32, 33, 100, 1023, 1024, 1025, 2047, 2048, 2049, 4095, 4096, 4097, 50000]


def dev_sort(keys, vals, b0, b1, key_bytes):
    k = np.ascontiguousarray(keys, np.uint64).copy()
    v = np.ascontiguousarray(vals, np.uint32).copy()
    code = ix.lib().impgx_debug_sort_pairs(0, k.ctypes.data_as(C.c_void_p), v.ctypes.data_as(C.c_void_p),
                                           C.c_uint64(len(k)), C.c_int(b0), C.c_int(b1), C.c_int(key_bytes))
    assert code == 0, ix.lib().impgx_last_error()
    return k, v


@pytest.mark.parametrize("n", SIZES)
def test_sort_pairs_is_a_stable_sort_on_the_selected_bits(n):
    rng = np.random.default_rng(n)
    for key_bytes, b0, b1, spread in ((8, 0, 33, 50), (8, 0, 53, 1 << 52), (8, 0, 54, 1 << 53), (8, 0, 64, 7),
                                      (8, 5, 20, 1 << 40), (4, 0, 32, 1 << 31), (4, 0, 9, 300), (4, 3, 17, 1 << 20)):
        keys = rng.integers(0, spread, n, dtype=np.uint64)  # small spreads: many equal keys (stability matters)
        if key_bytes == 8 and b1 == 64:
            keys |= rng.integers(0, 2, n, dtype=np.uint64) << np.uint64(63)
        vals = np.arange(n, dtype=np.uint32)
        k, v = dev_sort(keys, vals, b0, b1, key_bytes)
        field = (keys >> np.uint64(b0)) & np.uint64((1 << (b1 - b0)) - 1)
        order = np.argsort(field, kind="stable")
        assert (v == order.astype(np.uint32)).all(), (n, key_bytes, b0, b1)
        assert (k == keys[order]).all()


@pytest.mark.parametrize("n", SIZES)
def test_exclusive_scan(n):
    rng = np.random.default_rng(n)
    for hi in (2, 1 << 40):
        a = rng.integers(0, hi, n, dtype=np.uint64)
        a[-1] = 0  # the pipeline's convention: n values + one slot that receives the total
        d = a.copy()
        assert ix.lib().impgx_debug_exclusive_scan(0, d.ctypes.data_as(C.c_void_p), C.c_uint64(n)) == 0
        want = np.zeros(n, np.uint64)
        want[1:] = np.cumsum(a[:-1], dtype=np.uint64)
        assert (d == want).all(), n
