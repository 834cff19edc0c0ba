"""Model files through the C++ reader of libb2k.so (kaldi_b200/csrc/model_io.cu, include/b2k.h b2k_model_*):
`final.mdl` (TransitionModel + AmNnetSimple, hmm/transition-model.cc:394, nnet3/am-nnet-simple.cc:34) or a raw
nnet3 model (nnet3/nnet-nnet.cc:630), binary or text → the layer list, named weights and transition-id → pdf table
that b2k_nnet_compile / b2k_fst_create take.  kaldi_io.py holds the same readers in Python as the test oracle."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib


class KaldiModel:
    """Owns a b2k_model handle.  `NnetComputer.from_model(m, ...)` compiles and uploads it without going through
    the Python compiler; `m.tid2pdf` feeds `CudaFst`."""

    def __init__(self, path: str, is_mdl: bool | None = None, frame_subsampling_factor: int | None = None):
        L = _lib.lib()
        if is_mdl is None:
            is_mdl = str(path).endswith(".mdl")
        self.h = C.c_void_p()
        L.b2k_model_read.argtypes = [C.c_char_p, C.c_int32, C.c_void_p]
        _lib.check(L.b2k_model_read(str(path).encode(), int(bool(is_mdl)), C.byref(self.h)))
        if frame_subsampling_factor is not None:               # the tool's --frame-subsampling-factor (not stored in the file)
            L.b2k_model_set_frame_subsampling_factor.argtypes = [C.c_void_p, C.c_int32]
            _lib.check(L.b2k_model_set_frame_subsampling_factor(self.h, int(frame_subsampling_factor)))
        self._finish_init(L)

    def _finish_init(self, L):
        L.b2k_model_frame_subsampling_ambiguous.argtypes = [C.c_void_p]
        self.frame_subsampling_ambiguous = bool(L.b2k_model_frame_subsampling_ambiguous(self.h))
        info = (C.c_int32 * 8)()
        L.b2k_model_info.argtypes = [C.c_void_p, C.c_void_p]
        _lib.check(L.b2k_model_info(self.h, info))
        (self.feat_dim, self.ivector_dim, self.num_pdfs, self.frame_subsampling_factor, self.n_layers, self.n_weights,
         n_tid, has_priors) = [int(x) for x in info]
        self.has_priors = bool(has_priors)
        for f in ("b2k_model_layers", "b2k_model_weights", "b2k_model_tid2pdf", "b2k_model_tid2phone"):
            getattr(L, f).restype = C.c_void_p
            getattr(L, f).argtypes = [C.c_void_p]
        self.layers_ptr = L.b2k_model_layers(self.h)
        self.weights_ptr = L.b2k_model_weights(self.h)
        self.tid2pdf = self.tid2phone = None
        if n_tid:
            p = C.cast(L.b2k_model_tid2pdf(self.h), C.POINTER(C.c_int32))
            self.tid2pdf = np.ctypeslib.as_array(p, shape=(n_tid,)).copy()
            p = C.cast(L.b2k_model_tid2phone(self.h), C.POINTER(C.c_int32))
            self.tid2phone = np.ctypeslib.as_array(p, shape=(n_tid,)).copy()       # TransitionIdToPhone, for endpointing

    @classmethod
    def from_arch(cls, arch: dict, W: dict, tid2pdf=None) -> "KaldiModel":
        """b2k_model_from_arrays: a model that never was a file (synthetic weights of bench.py / tests)."""
        from .nnet_compile import _Layer, _Weight, _layer
        L = _lib.lib()
        self = cls.__new__(cls)
        layers = (_Layer * len(arch["layers"]))(*[_layer(x) for x in arch["layers"]])
        keep = {k: np.ascontiguousarray(v, np.float32) for k, v in W.items()}
        ws = (_Weight * len(keep))()
        for i, (k, v) in enumerate(keep.items()):
            rows, cols = (v.shape if v.ndim == 2 else (v.shape[0], 1))
            ws[i] = _Weight(k.encode(), v.ctypes.data, v.size, int(rows), int(cols))
        t2p = None if tid2pdf is None else np.ascontiguousarray(tid2pdf, np.int32)
        self.h = C.c_void_p()
        L.b2k_model_from_arrays.argtypes = [C.c_int32] * 4 + [C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p]
        _lib.check(L.b2k_model_from_arrays(arch["feat_dim"], arch["ivector_dim"], arch["num_pdfs"], arch["frame_subsampling_factor"],
                                           layers, len(layers), ws, len(ws), None if t2p is None else t2p.ctypes.data,
                                           0 if t2p is None else len(t2p), C.byref(self.h)))
        self._finish_init(L)
        return self

    def layer_types(self) -> list[tuple[str, str]]:
        from .nnet_compile import _Layer
        a = C.cast(self.layers_ptr, C.POINTER(_Layer))
        return [(a[i].type.decode(), a[i].name.decode()) for i in range(self.n_layers)]

    def weights(self) -> dict[str, np.ndarray]:
        from .nnet_compile import _Weight
        a = C.cast(self.weights_ptr, C.POINTER(_Weight))
        out = {}
        for i in range(self.n_weights):
            w = a[i]
            v = np.ctypeslib.as_array(C.cast(w.data, C.POINTER(C.c_float)), shape=(w.size,)).copy()
            out[w.name.decode()] = v.reshape(w.rows, w.cols) if w.cols > 1 else v
        return out

    def compile(self, num_frames: int, frames_per_chunk: int = 21, acoustic_scale: float = 1.0, use_priors: bool = True,
                conv_mode: str | None = None) -> "C.c_void_p":
        """b2k_nnet_compile on the model's own arrays → a b2k_nnet_program handle (caller destroys it)."""
        from .nnet_compile import _Cfg
        L = _lib.lib()
        if self.frame_subsampling_ambiguous:
            raise _lib.B2kError(_lib.B2K_ERR_INVALID,
                                "the model's layers do not decide the frame subsampling factor: construct KaldiModel with "
                                "frame_subsampling_factor (3 for chain models, 1 otherwise)")
        cfg = _Cfg(self.feat_dim, self.ivector_dim, self.num_pdfs, self.frame_subsampling_factor, int(num_frames),
                   int(frames_per_chunk), int(use_priors), int((conv_mode or "patch") == "dense"), float(acoustic_scale))
        prog = C.c_void_p()
        L.b2k_nnet_compile.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p]
        _lib.check(L.b2k_nnet_compile(C.byref(cfg), C.c_void_p(self.layers_ptr), self.n_layers,
                                      C.c_void_p(self.weights_ptr), self.n_weights, C.byref(prog)))
        return prog

    def close(self):
        if getattr(self, "h", None):
            L = _lib.lib()
            L.b2k_model_destroy.argtypes = [C.c_void_p]
            L.b2k_model_destroy(self.h)
            self.h = None

    __del__ = close
