"""Endpointing, the host mirror of online2/online-endpoint.h over the C ABI (b2k_endpoint_*; kaldi_b200/csrc/host_utils.cu).

    cfg = OnlineEndpointConfig.from_conf("conf/online.conf")          # or OnlineEndpointConfig(); cfg.apply_options("--endpoint...")
    if EndpointDetected(cfg, model.tid2phone, best_path_ilabels, num_frames_decoded, frame_shift, final_relative_cost): ...

The five rules, their defaults and the option names are the reference's (online-endpoint.h:146-166); the best path is the
one `lattice.best_path` / `CudaDecoder.GetBestPath` give, the final relative cost is the decoder's."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib


class _Rule(C.Structure):
    _fields_ = [("must_contain_nonsilence", C.c_int32), ("min_trailing_silence", C.c_float), ("max_relative_cost", C.c_float),
                ("min_utterance_length", C.c_float)]


class _Cfg(C.Structure):
    _fields_ = [("rule", _Rule * 5), ("silence_phones", C.c_char * 512)]


class OnlineEndpointConfig:
    """online-endpoint.h:128 (rule1..rule5 as `rules[0..4]`, `silence_phones` as the colon-separated string)."""

    def __init__(self):
        self.c = _Cfg()
        _lib.check(_lib.lib().b2k_endpoint_cfg_default(C.byref(self.c)))

    @classmethod
    def from_conf(cls, path: str) -> "OnlineEndpointConfig":
        self = cls()
        L = _lib.lib()
        L.b2k_endpoint_cfg_from_conf.argtypes = [C.c_char_p, C.c_void_p]
        _lib.check(L.b2k_endpoint_cfg_from_conf(str(path).encode(), C.byref(self.c)))
        return self

    def apply_options(self, text: str) -> "OnlineEndpointConfig":
        L = _lib.lib()
        L.b2k_endpoint_cfg_apply_options.argtypes = [C.c_char_p, C.c_void_p]
        _lib.check(L.b2k_endpoint_cfg_apply_options(text.encode(), C.byref(self.c)))
        return self

    @property
    def silence_phones(self) -> str:
        return self.c.silence_phones.decode()

    @silence_phones.setter
    def silence_phones(self, v: str):
        self.c.silence_phones = v.encode()

    @property
    def rules(self):
        return self.c.rule


def TrailingSilenceLength(tid2phone: np.ndarray, silence_phones: str, ilabels) -> int:
    """online-endpoint.cc:78: silence frames at the end of a best path (input labels in time order, epsilons allowed)."""
    L = _lib.lib()
    t = np.ascontiguousarray(tid2phone, np.int32)
    a = np.ascontiguousarray(ilabels, np.int32)
    out = C.c_int32()
    L.b2k_trailing_silence_frames.argtypes = [C.c_void_p, C.c_int32, C.c_char_p, C.c_void_p, C.c_int64, C.c_void_p]
    _lib.check(L.b2k_trailing_silence_frames(t.ctypes.data, t.size, silence_phones.encode(), a.ctypes.data, a.size, C.byref(out)))
    return out.value


def EndpointDetected(config: OnlineEndpointConfig, *args) -> bool:
    """Both forms of the reference (online-endpoint.h:171-190):
    EndpointDetected(config, num_frames_decoded, trailing_silence_frames, frame_shift_in_seconds, final_relative_cost)
    EndpointDetected(config, tid2phone, best_path_ilabels, num_frames_decoded, frame_shift_in_seconds, final_relative_cost)"""
    L = _lib.lib()
    out = C.c_int32()
    if len(args) == 4:
        n, sil, shift, cost = args
        L.b2k_endpoint_detected.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_float, C.c_float, C.c_void_p]
        _lib.check(L.b2k_endpoint_detected(C.byref(config.c), int(n), int(sil), float(shift), float(cost), C.byref(out)))
    elif len(args) == 5:
        tid2phone, ilabels, n, shift, cost = args
        t = np.ascontiguousarray(tid2phone, np.int32)
        a = np.ascontiguousarray(ilabels, np.int32)
        L.b2k_endpoint_detected_on_path.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_int64, C.c_int32, C.c_float,
                                                    C.c_float, C.c_void_p, C.c_void_p]
        _lib.check(L.b2k_endpoint_detected_on_path(C.byref(config.c), t.ctypes.data, t.size, a.ctypes.data, a.size, int(n), float(shift),
                                                   float(cost), C.byref(out), None))
    else:
        raise TypeError("EndpointDetected takes 4 or 5 arguments after the config")
    return bool(out.value)
