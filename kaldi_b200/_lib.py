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
    return _lib


def check(rc: int) -> None:
    if rc != B2K_OK:
        raise B2kError(rc, lib().b2k_last_error().decode("utf-8", "replace"))


def kernel_launch_count() -> int:
    return int(lib().b2k_kernel_launch_count())
