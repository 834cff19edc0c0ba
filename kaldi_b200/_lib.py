"""ctypes loader for the b2k C-ABI shared library (include/b2k.h).

The product has no CPU path: if the CUDA extension is missing this raises, and
every compute entry point returns B2K_ERR_NO_DEVICE without an sm_100 GPU.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libb2k.so")

B2K_OK = 0
B2K_ERR_INVALID, B2K_ERR_NO_DEVICE, B2K_ERR_CUDA, B2K_ERR_OVERFLOW, B2K_ERR_STATE = 1, 2, 3, 4, 5


class B2kError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"b2k error {code}: {msg}")
        self.code = code


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "or `make -C kaldi_b200`. kaldi_b200 has no CPU fallback.")
        _lib = C.CDLL(LIB_PATH)
        _lib.b2k_last_error.restype = C.c_char_p
        _lib.b2k_kernel_launch_count.restype = C.c_int64
        _declare(_lib)
    return _lib


def _declare(L) -> None:
    """argtypes for every entry point of include/b2k.h (64-bit and pointer
    arguments must not go through ctypes' default int conversion)."""
    vp, i32, i64, f32 = C.c_void_p, C.c_int32, C.c_int64, C.c_float
    P = C.POINTER
    sig = {
        "b2k_fst_create": [vp, P(vp)], "b2k_fst_destroy": [vp],
        "b2k_fst_num_states": [vp], "b2k_fst_start": [vp],
        "b2k_dec_cfg_default": [vp],
        "b2k_dec_create": [vp, vp, i32, i32, P(vp)], "b2k_dec_destroy": [vp],
        "b2k_dec_init_decoding": [vp, P(i32), i32, vp],
        "b2k_dec_advance_decoding": [vp, P(i32), vp, i32, vp],
        "b2k_dec_advance_decoding_frames": [vp, P(i32), vp, P(i32), i32, i32, vp],
        "b2k_dec_num_frames_decoded": [vp, i32, P(i32)],
        "b2k_dec_finalize_decoding": [vp, P(i32), i32, vp],
        "b2k_dec_channel_info": [vp, i32, P(i64)],
        "b2k_dec_get_raw_lattice": [vp, i32, vp, vp],
        "b2k_dec_get_raw_lattices": [vp, P(i32), i32, vp, P(i64), P(i64), P(i64), vp],
        "b2k_dec_debug_frame": [vp, i32, i32, P(i32), P(f32), P(i64), P(i32), P(i64), i64, i64],
        "b2k_dec_frame_info": [vp, i32, P(f32), P(f32), P(i32), i32],
        "b2k_feat_cfg_default": [vp], "b2k_feat_create": [vp, P(vp)], "b2k_feat_destroy": [vp],
        "b2k_feat_dim": [vp], "b2k_feat_num_frames": [vp, i64, i32],
        "b2k_feat_compute_batched": [vp, i32, vp, P(i32), P(i32), P(i32), vp, i32, vp],
        "b2k_cmvn_apply_batched": [vp, vp, i32, vp, vp, i32, i32, P(i32), P(i32), vp, vp, vp, vp],
        "b2k_ivec_create": [vp, P(f32), P(f32), P(f32), P(f32), P(C.c_double), P(C.c_double), P(C.c_double), P(vp)],
        "b2k_ivec_destroy": [vp],
        "b2k_ivec_compute_batched": [vp, i32, vp, i32, i32, P(i32), i32, vp, i32, vp],
        "b2k_nnet_create": [vp, i32, vp, i32, P(f32), i64, i32, P(vp)], "b2k_nnet_destroy": [vp],
        "b2k_nnet_num_output_frames": [vp], "b2k_nnet_output_dim": [vp],
        "b2k_nnet_num_launches_per_run": [vp],
        "b2k_nnet_run": [vp, i32, vp, i32, vp, i32, vp, i32, vp],
    }
    L.b2k_nnet_flops_per_lane.argtypes = [vp]
    L.b2k_nnet_flops_per_lane.restype = C.c_double
    if False:
        pass
    for name, argtypes in sig.items():
        fn = getattr(L, name)
        fn.argtypes = argtypes
        fn.restype = C.c_int32


def check(rc: int) -> None:
    if rc != B2K_OK:
        raise B2kError(rc, lib().b2k_last_error().decode("utf-8", "replace"))


def kernel_launch_count() -> int:
    return int(lib().b2k_kernel_launch_count())
