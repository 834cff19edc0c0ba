"""Host-only C++ pieces of libb2k.so against their Python twins."""
import ctypes as C

import numpy as np
import pytest


def test_ivector_online_schedule_cpp_equals_python():
    try:
        from kaldi_b200 import _lib
        L = _lib.lib()
    except Exception as e:
        pytest.skip(str(e))
    from kaldi_b200.ivector import online_ivector_schedule
    L.b2k_ivec_online_schedule.argtypes = [C.c_int64] + [C.c_int32] * 8 + [C.c_void_p, C.c_int32, C.c_void_p]
    rng = np.random.default_rng(0)
    for _ in range(200):
        ns = int(rng.integers(400, 400000))
        chunk = int(rng.choice([160, 2880, 8000, 16000, 10**9 // 1000]))
        fl, fs = 400, 160
        T = 0 if ns < fl else 1 + (ns - fl) // fs
        if T == 0:
            continue
        rc, fpc, sub, spl = int(rng.integers(0, 45)), int(rng.choice([21, 51, 150])), 3, int(rng.integers(0, 4))
        want = online_ivector_schedule(ns, chunk, fl, fs, T, rc, fpc, sub, spl)
        got = np.zeros(len(want) + 4, np.int32)
        n = C.c_int32()
        assert L.b2k_ivec_online_schedule(ns, chunk, fl, fs, T, rc, fpc, sub, spl, got.ctypes.data, got.size, C.addressof(n)) == 0
        assert n.value == len(want)
        np.testing.assert_array_equal(got[:n.value], want)
