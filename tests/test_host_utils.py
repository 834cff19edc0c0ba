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


def test_schedule_marks_chunks_that_run_before_any_ivector_frame_is_ready():
    """A model without right context, three input frames per chunk and audio fed 720 samples at a time: the first chunk is
    computed when three feature frames exist and the i-vector pipeline (splice right context 3) has none ready -- the
    reference leaves that chunk's i-vector zero (decodable-online-looped.cc:188-197); the schedule says -1 (it used to say
    frame 0: ADVICE r01).  Both twins."""
    try:
        from kaldi_b200 import _lib
        L = _lib.lib()
    except Exception as e:
        pytest.skip(str(e))
    from kaldi_b200.ivector import online_ivector_schedule
    ns, chunk, fl, fs, rc, fpc, sub, spl = 16000, 720, 400, 160, 0, 3, 3, 3
    T = 1 + (ns - fl) // fs
    want = online_ivector_schedule(ns, chunk, fl, fs, T, rc, fpc, sub, spl)
    assert want[0] == -1 and want[1] >= 0 and np.all(np.diff(want) >= 0) and want[-1] == T - 1
    L.b2k_ivec_online_schedule.argtypes = [C.c_int64] + [C.c_int32] * 8 + [C.c_void_p, C.c_int32, C.c_void_p]
    got = np.zeros(len(want), np.int32)
    n = C.c_int32()
    assert L.b2k_ivec_online_schedule(ns, chunk, fl, fs, T, rc, fpc, sub, spl, got.ctypes.data, got.size, C.addressof(n)) == 0
    np.testing.assert_array_equal(got, want)
    # the reference-side restatement (oracle/ref_wrap/ivector_wrap.cc) leaves such a chunk's i-vector at zero as well
    from kaldi_b200 import ivector as IVM
    from oracle import ivector_oracle as IO
    try:
        R = IO.RefIvector(IVM.make_synthetic_extractor(1, num_gauss=16, ivector_dim=10))
    except Exception as e:
        pytest.skip(str(e))
    feats = np.random.default_rng(0).standard_normal((T, 40)).astype(np.float32)
    out = R.run(feats, want)
    assert not out[0].any() and out[1].any()
