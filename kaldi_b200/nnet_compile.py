"""ctypes view of the C++ nnet3 program compiler (kaldi_b200/csrc/nnet_compile.cu, include/b2k.h
b2k_nnet_compile*): the host-side, device-free counterpart of nnet_model.compile_program, which remains its
test oracle (tests/test_nnet_compile_cpp.py requires identical programs)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from .nnet import _Node, _Op


class _Layer(C.Structure):
    _fields_ = [("type", C.c_char * 24), ("name", C.c_char * 48), ("side", C.c_char * 48),
                ("dim", C.c_int32), ("bottleneck", C.c_int32), ("stride", C.c_int32), ("big", C.c_int32),
                ("small", C.c_int32), ("log_softmax", C.c_int32),
                ("bypass", C.c_float), ("append_ivector", C.c_float), ("target_rms", C.c_float),
                ("height", C.c_int32), ("filters1", C.c_int32), ("filters2", C.c_int32),
                ("height_in", C.c_int32), ("height_out", C.c_int32), ("height_subsample_out", C.c_int32),
                ("filters_in", C.c_int32), ("filters_out", C.c_int32),
                ("n_time_offsets", C.c_int32), ("time_offsets", C.c_int32 * 8),
                ("n_height_offsets", C.c_int32), ("height_offsets", C.c_int32 * 8)]


class _Weight(C.Structure):
    _fields_ = [("name", C.c_char_p), ("data", C.c_void_p), ("size", C.c_int64), ("rows", C.c_int32), ("cols", C.c_int32)]


class _Cfg(C.Structure):
    _fields_ = [("feat_dim", C.c_int32), ("ivector_dim", C.c_int32), ("num_pdfs", C.c_int32),
                ("frame_subsampling_factor", C.c_int32), ("num_frames", C.c_int32), ("frames_per_chunk", C.c_int32),
                ("use_priors", C.c_int32), ("conv_dense", C.c_int32), ("acoustic_scale", C.c_float)]


def _layer(L: dict) -> _Layer:
    x = _Layer()
    x.type, x.name, x.side = L["type"].encode(), L["name"].encode(), L.get("side", "").encode()
    x.target_rms = 1.0
    for k in ("dim", "bottleneck", "stride", "big", "small", "height", "filters1", "filters2", "height_in", "height_out",
              "height_subsample_out", "filters_in", "filters_out"):
        if k in L:
            setattr(x, k, int(L[k]))
    x.log_softmax = int(bool(L.get("log_softmax", False)))
    x.bypass = float(L.get("bypass", 0.0))
    x.append_ivector = float(L.get("append_ivector", 0.0) or 0.0)
    if "target_rms" in L:
        x.target_rms = float(L["target_rms"])
    for key, n_attr, arr_attr in (("time_offsets", "n_time_offsets", "time_offsets"),
                                  ("height_offsets", "n_height_offsets", "height_offsets")):
        v = L.get(key, [])
        assert len(v) <= 8
        setattr(x, n_attr, len(v))
        for i, o in enumerate(v):
            getattr(x, arr_attr)[i] = int(o)
    return x


class CompiledProgram:
    """Owns a b2k_nnet_program; exposes nodes/ops as ctypes arrays (the ABI structs b2k_nnet_create takes) and the blob."""

    def __init__(self, arch: dict, W: dict, num_frames: int, frames_per_chunk: int = 21, acoustic_scale: float = 1.0,
                 use_priors: bool = True, conv_mode: str | None = None, window: tuple | None = None):
        """window = (first_output_t, num_outputs[, ivector_rows]): b2k_nnet_compile_window, the computation request of
        BatchedStaticNnet3 for a chunk with its context (num_frames = the window's input frames); ivector_rows > 1: chunk n
        of a looped run with the i-vectors of chunks n-(rows-1) .. n."""
        L = _lib.lib()
        layers = (_Layer * len(arch["layers"]))(*[_layer(x) for x in arch["layers"]])
        self._keep = {k: np.ascontiguousarray(v, np.float32) for k, v in W.items()}
        ws = (_Weight * len(self._keep))()
        for i, (k, v) in enumerate(self._keep.items()):
            rows, cols = (v.shape if v.ndim == 2 else (v.shape[0], 1))
            ws[i] = _Weight(k.encode(), v.ctypes.data, v.size, int(rows), int(cols))
        cfg = _Cfg(arch["feat_dim"], arch["ivector_dim"], arch["num_pdfs"], arch["frame_subsampling_factor"],
                   int(num_frames), int(frames_per_chunk), int(use_priors), int((conv_mode or "patch") == "dense"),
                   float(acoustic_scale))
        self.h = C.c_void_p()
        L.b2k_nnet_compile.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p]
        L.b2k_nnet_compile_window.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p]
        if window is None:
            _lib.check(L.b2k_nnet_compile(C.byref(cfg), layers, len(layers), ws, len(ws), C.byref(self.h)))
        else:
            _lib.check(L.b2k_nnet_compile_window(C.byref(cfg), int(window[0]), int(window[1]), int(window[2]) if len(window) > 2 else 1,
                                                 layers, len(layers), ws, len(ws), C.byref(self.h)))
        nn, no, bl = C.c_int32(), C.c_int32(), C.c_int64()
        L.b2k_nnet_program_sizes.argtypes = [C.c_void_p] * 4
        _lib.check(L.b2k_nnet_program_sizes(self.h, C.byref(nn), C.byref(no), C.byref(bl)))
        for f, rt in (("b2k_nnet_program_nodes", C.POINTER(_Node)), ("b2k_nnet_program_ops", C.POINTER(_Op)),
                      ("b2k_nnet_program_blob", C.POINTER(C.c_float))):
            getattr(L, f).restype = rt
            getattr(L, f).argtypes = [C.c_void_p]
        self.nodes = [L.b2k_nnet_program_nodes(self.h)[i] for i in range(nn.value)]
        self.ops = [L.b2k_nnet_program_ops(self.h)[i] for i in range(no.value)]
        self.blob = np.ctypeslib.as_array(L.b2k_nnet_program_blob(self.h), shape=(bl.value,)).copy()
        info = (C.c_int64 * 8)()
        L.b2k_nnet_program_info.argtypes = [C.c_void_p, C.c_void_p]
        _lib.check(L.b2k_nnet_program_info(self.h, info))
        (self.n_out, self.n_chunks, self.left_context, self.right_context, self.model_left, self.model_right,
         self.ivector_m, self.arena_size) = [int(x) for x in info]

    def __del__(self):
        if getattr(self, "h", None):
            _lib.lib().b2k_nnet_program_destroy.argtypes = [C.c_void_p]
            _lib.lib().b2k_nnet_program_destroy(self.h)
            self.h = None


def abi_arrays(arch: dict, W: dict):
    """(layers, weights, keep-alive) as the ABI arrays b2k_nnet_compile / b2k_nnet_stream_create take."""
    layers = (_Layer * len(arch["layers"]))(*[_layer(x) for x in arch["layers"]])
    keep = {k: np.ascontiguousarray(v, np.float32) for k, v in W.items()}
    ws = (_Weight * len(keep))()
    for i, (k, v) in enumerate(keep.items()):
        rows, cols = (v.shape if v.ndim == 2 else (v.shape[0], 1))
        ws[i] = _Weight(k.encode(), v.ctypes.data, v.size, int(rows), int(cols))
    return layers, ws, keep


def model_context(arch: dict) -> tuple:
    """ComputeSimpleNnetContext of the layer list (b2k_nnet_model_context)."""
    L = _lib.lib()
    layers = (_Layer * len(arch["layers"]))(*[_layer(x) for x in arch["layers"]])
    cfg = _Cfg(arch["feat_dim"], arch["ivector_dim"], arch["num_pdfs"], arch["frame_subsampling_factor"], 1,
               arch["frame_subsampling_factor"], 0, 0, 1.0)
    l, r = C.c_int32(), C.c_int32()
    L.b2k_nnet_model_context.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]
    _lib.check(L.b2k_nnet_model_context(C.byref(cfg), layers, len(layers), C.byref(l), C.byref(r)))
    return l.value, r.value


def stream_account(left: int, right: int, sub: int, in_ctx: int, n_new: int, flush: bool) -> tuple:
    """b2k_nnet_stream_account: (frames in context afterwards, output frames) of one BatchContextSwitch."""
    L = _lib.lib()
    a, o = C.c_int32(), C.c_int32()
    L.b2k_nnet_stream_account.argtypes = [C.c_int32] * 6 + [C.c_void_p, C.c_void_p]
    _lib.check(L.b2k_nnet_stream_account(left, right, sub, in_ctx, n_new, int(flush), C.byref(a), C.byref(o)))
    return a.value, o.value


def looped_ivector_rows(arch: dict, frames_per_chunk: int) -> int:
    """b2k_nnet_looped_ivector_rows: i-vector rows a looped window program reads (ceil(left / C) + lag + 1)."""
    L = _lib.lib()
    layers = (_Layer * len(arch["layers"]))(*[_layer(x) for x in arch["layers"]])
    cfg = _Cfg(arch["feat_dim"], arch["ivector_dim"], arch["num_pdfs"], arch["frame_subsampling_factor"], 1,
               int(frames_per_chunk), 0, 0, 1.0)
    r = C.c_int32()
    L.b2k_nnet_looped_ivector_rows.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]
    _lib.check(L.b2k_nnet_looped_ivector_rows(C.byref(cfg), layers, len(layers), C.byref(r)))
    return r.value
