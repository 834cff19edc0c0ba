"""nnet3-stage oracle — TEST INFRASTRUCTURE ONLY.

* `RefNnet`: ctypes view of oracle/_ref/libkaldi_ref_nnet3.so = the reference's
  OWN nnet3 CPU forward (Nnet::ReadConfig -> SetBatchnormTestMode ->
  CollapseModel -> DecodableNnetSimpleLooped, i.e. the looped computation the
  online2 decodable runs) compiled from /root/reference/src (oracle/ref_nnet.py).
* `forward_dense`: numpy restatement of the components' Propagate for the
  TDNN-F family, pinned against RefNnet (tests/test_nnet_oracle.py):
    TdnnComponent::Propagate            nnet3/nnet-tdnn-component.cc:181-211
    AffineComponent/LinearComponent/FixedAffineComponent::Propagate
                                        nnet3/nnet-simple-component.cc:1242,3224,3392
    RectifiedLinearComponent::Propagate nnet3/nnet-simple-component.cc:964-972
    BatchNormComponent (test mode)      nnet3/nnet-normalize-component.cc:455-466
    LogSoftmaxComponent::Propagate      nnet3/nnet-simple-component.cc:3618-3625
    output post-processing              nnet3/decodable-online-looped.cc:218-223
    i-vector per chunk / Round()        nnet3/nnet-compile-looped.cc:179-205,
                                        decodable-simple-looped.cc:262-279
PARITY STATUS: pinned to the compiled reference (no golden outputs exist in the
reference tree for nnet3; its own tests are self-consistency only).
GEMM association is BLAS-defined, so comparisons are relative-tolerance based
(1e-4 of the output scale, the north-star figure).
"""
from __future__ import annotations

import ctypes as C
import os
import tempfile

import numpy as np

from kaldi_b200 import nnet_model as NM

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_ref", "libkaldi_ref_nnet3.so")


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


class RefNnet:
    def __init__(self, arch: dict, W: dict, frames_per_chunk: int = 20, acoustic_scale: float = 1.0,
                 use_priors: bool = True, collapse: bool = True):
        if not os.path.exists(_SO):
            from . import ref_nnet
            ref_nnet.build()
        L = self.lib = C.CDLL(_SO)
        L.ref_nnet_create.restype = C.c_void_p
        L.ref_nnet_component_name.restype = C.c_char_p
        L.ref_nnet_component_type.restype = C.c_char_p
        for f in ("ref_nnet_destroy", "ref_nnet_num_components", "ref_nnet_num_params", "ref_nnet_get_params",
                  "ref_nnet_set_params", "ref_nnet_set_batchnorm", "ref_nnet_prepare", "ref_nnet_info",
                  "ref_nnet_forward", "ref_nnet_component_name", "ref_nnet_component_type"):
            getattr(L, f).argtypes = None
        self.arch = arch
        with tempfile.TemporaryDirectory() as td:
            cfg = NM.to_nnet3_config(arch, W, td)
            self.h = C.c_void_p(L.ref_nnet_create(cfg.encode()))
        if not self.h:
            raise RuntimeError("reference Nnet::ReadConfig failed")
        # BatchNorm components whose dim differs from their statistics (block-dim) or with a target rms
        bn_shape = {}
        for Ly in arch["layers"]:
            if Ly["type"] == "conv":
                bn_shape[Ly["name"] + ".batchnorm"] = (Ly["height_out"] * Ly["filters_out"], 1.0)
            elif Ly["type"] == "ivector-linear-bn":
                bn_shape[Ly["name"] + "-batchnorm"] = (Ly["dim"], float(Ly["target_rms"]))
        n = L.ref_nnet_num_components(self.h)
        for i in range(n):
            name = L.ref_nnet_component_name(self.h, i).decode()
            typ = L.ref_nnet_component_type(self.h, i).decode()
            if typ == "BatchNormComponent":
                if name + ".mean" not in W and arch.get("recipe_extras"):
                    mean, var = np.zeros(32, np.float32), np.ones(32, np.float32)      # prefinal-xent.batchnorm1
                else:
                    mean, var = W[name + ".mean"], W[name + ".var"]
                dim, target_rms = bn_shape.get(name, (mean.size, 1.0))
                r = L.ref_nnet_set_batchnorm(self.h, i, C.c_int(dim), C.c_int(mean.size), C.c_float(NM.BN_EPS),
                                             C.c_float(target_rms), C.c_float(1000.0), _p(mean, C.c_float), _p(var, C.c_float))
                assert r == 0, name
            elif typ in ("NaturalGradientAffineComponent", "AffineComponent", "TdnnComponent", "LinearComponent",
                         "TimeHeightConvolutionComponent"):
                if name + ".w" not in W and arch.get("recipe_extras"):
                    continue                                  # the xent branch keeps the reference's own initialisation
                w = W[name + ".w"]
                vec = w.reshape(-1)
                if name + ".b" in W:
                    vec = np.concatenate([vec, W[name + ".b"]])
                vec = np.ascontiguousarray(vec, np.float32)
                assert L.ref_nnet_num_params(self.h, i) == vec.size, (name, typ, L.ref_nnet_num_params(self.h, i), vec.size)
                assert L.ref_nnet_set_params(self.h, i, _p(vec, C.c_float)) == 0
        pri = np.ascontiguousarray(W["priors"], np.float32)
        r = L.ref_nnet_prepare(self.h, C.c_int(frames_per_chunk), C.c_int(arch["frame_subsampling_factor"]),
                               C.c_float(acoustic_scale), _p(pri, C.c_float) if use_priors else None,
                               C.c_int(pri.size if use_priors else 0), C.c_int(int(collapse)))
        if r != 0:
            raise RuntimeError("reference looped compilation failed")
        info = (C.c_int * 4)()
        L.ref_nnet_info(self.h, info)
        self.left_context, self.right_context, self.frames_per_chunk, self.output_dim = list(info)

    def __del__(self):
        try:
            self.lib.ref_nnet_destroy(self.h)
        except Exception:
            pass

    def forward(self, feats: np.ndarray, ivectors: np.ndarray | None = None, period: int = 1) -> np.ndarray:
        f = np.ascontiguousarray(feats, np.float32)
        T, D = f.shape
        n_out = (T + self.arch["frame_subsampling_factor"] - 1) // self.arch["frame_subsampling_factor"]
        out = np.zeros((n_out, self.output_dim), np.float32)
        if ivectors is not None:
            iv = np.ascontiguousarray(ivectors, np.float32)
            r = self.lib.ref_nnet_forward(self.h, _p(f, C.c_float), T, D, _p(iv, C.c_float), iv.shape[0], iv.shape[1],
                                          int(period), _p(out, C.c_float), n_out)
        else:
            r = self.lib.ref_nnet_forward(self.h, _p(f, C.c_float), T, D, None, 0, 0, 1, _p(out, C.c_float), n_out)
        if r != n_out:
            raise RuntimeError(f"reference forward failed ({r})")
        return out

    def forward_simple(self, feats: np.ndarray, online_ivectors: np.ndarray | None, period: int, frames_per_chunk: int) -> np.ndarray:
        """DecodableNnetSimple (the offline tools' decodable, nnet3/nnet-am-decodable-simple.cc) over the utterance."""
        f = np.ascontiguousarray(feats, np.float32)
        T, D = f.shape
        sub = self.arch["frame_subsampling_factor"]
        n_out = (T + sub - 1) // sub
        out = np.zeros((n_out, self.output_dim), np.float32)
        self.lib.ref_nnet_forward_simple.argtypes = None
        if online_ivectors is not None:
            iv = np.ascontiguousarray(online_ivectors, np.float32)
            r = self.lib.ref_nnet_forward_simple(self.h, _p(f, C.c_float), T, D, _p(iv, C.c_float), iv.shape[0], iv.shape[1], int(period),
                                                 int(frames_per_chunk), _p(out, C.c_float), n_out)
        else:
            r = self.lib.ref_nnet_forward_simple(self.h, _p(f, C.c_float), T, D, None, 0, 0, 1, int(frames_per_chunk), _p(out, C.c_float), n_out)
        if r != n_out:
            raise RuntimeError(f"reference simple forward failed ({r})")
        return out

    def chunk_ivector_rows(self, num_frames: int, num_ivector_rows: int, period: int):
        """Row of the online_ivectors matrix chunk n reads: GetCurrentIvector(end_input_frame)
        (decodable-simple-looped.cc:190,262-279)."""
        sub = self.arch["frame_subsampling_factor"]
        n_out = (num_frames + sub - 1) // sub
        C_ = self.frames_per_chunk
        n_chunks = (n_out * sub + C_ - 1) // C_
        rows = []
        for n in range(n_chunks):
            end_input = (n + 1) * C_ + self.right_context
            rows.append(min(end_input // period, num_ivector_rows - 1))
        return rows


# ----------------------------------------------------------------------------- numpy restatement

def forward_dense(arch: dict, W: dict, feats: np.ndarray, chunk_ivectors: np.ndarray | None,
                  frames_per_chunk: int = 21, acoustic_scale: float = 1.0, use_priors: bool = True) -> np.ndarray:
    """Dense-in-time evaluation on t in [-pad, T+pad) with edge-clamped input
    (decodable-simple-looped.cc:150-163); returns rows for t = 0, 3, 6, ...
    chunk_ivectors: [n_chunks, ivector_dim], the i-vector chunk n supplied."""
    f32 = np.float32
    sub = arch["frame_subsampling_factor"]
    x = np.asarray(feats, f32)
    T = x.shape[0]
    n_out = (T + sub - 1) // sub
    Lc, Rc = NM.model_context(arch)
    pad_l, pad_r = Lc + 4, Rc + 4 + sub * n_out - T
    tt = np.arange(-pad_l, T + pad_r)
    cur = x[np.clip(tt, 0, T - 1)]                      # cur[i] is time tt[i]
    Cc = frames_per_chunk
    m = (Cc + Rc - 1) // Cc
    iv_t = None
    if chunk_ivectors is not None:
        civ = np.asarray(chunk_ivectors, f32)
        rows = np.clip(np.floor_divide(tt, Cc) - m, 0, civ.shape[0] - 1)
        iv_t = civ[rows]

    def shift(a, o):          # a'[i] = a[i + o] (time t + o); edges filled by edge value (never used: padding is ample)
        idx = np.clip(np.arange(a.shape[0]) + o, 0, a.shape[0] - 1)
        return a[idx]

    def bn(a, name):
        s, o = NM.bn_scale_offset(W[name + ".mean"], W[name + ".var"])
        return (a * s[None, :]).astype(f32) + o[None, :]

    def affine(a, name, bias=True):
        y = (a @ W[name + ".w"].T).astype(f32)
        if bias and name + ".b" in W:
            y = y + W[name + ".b"][None, :]
        return y.astype(f32)

    fd = arch["feat_dim"]
    for L in arch["layers"]:
        t, n = L["type"], L["name"]
        if t == "idct":
            cur = affine(cur, n)
        elif t == "batchnorm":
            cur = bn(cur, n)
        elif t == "delta":
            b0 = cur
            b1 = (f32(-1.0) * shift(cur, -1)).astype(f32) + shift(cur, 1)
            b2 = (shift(cur, -2) + shift(cur, 2)).astype(f32) + (f32(-2.0) * cur).astype(f32)
            cur = bn(np.concatenate([b0, b1, b2], 1).astype(f32), n)
        elif t == "lda":
            cur = affine(np.concatenate([shift(cur, o) for o in NM.splice_of(L)] + [iv_t], 1).astype(f32), n)
        elif t == "relu-batchnorm":
            sp = NM.splice_of(L)
            a = cur if sp == [0] else np.concatenate([shift(cur, o) for o in sp], 1)
            if L.get("append_ivector"):
                a = np.concatenate([a, (f32(L["append_ivector"]) * iv_t).astype(f32)], 1)
            cur = bn(np.maximum(affine(a, n + ".affine"), 0), n + ".batchnorm")
        elif t == "tdnnf":
            s = L["stride"]
            a = np.concatenate([shift(cur, -s), cur], 1) if s else cur
            lin = affine(a, n + ".linear", bias=False)
            a2 = np.concatenate([lin, shift(lin, s)], 1) if s else lin
            y = bn(np.maximum(affine(a2, n + ".affine"), 0), n + ".batchnorm")
            cur = ((f32(L["bypass"]) * cur).astype(f32) + y).astype(f32)
        elif t == "linear":
            cur = affine(cur, n, bias=False)
        elif t == "prefinal":
            y = bn(np.maximum(affine(cur, n + ".affine"), 0), n + ".batchnorm1")
            cur = bn(affine(y, n + ".linear", bias=False), n + ".batchnorm2")
        elif t == "output":
            cur = affine(cur, n + ".affine")
            if L.get("log_softmax"):
                mx = cur.max(1, keepdims=True)
                cur = (cur - mx - np.log(np.exp(cur - mx).sum(1, keepdims=True))).astype(f32)
    out = cur[pad_l + sub * np.arange(n_out)]
    if use_priors:
        out = out - np.log(W["priors"]).astype(f32)[None, :]
    return (out * f32(acoustic_scale)).astype(f32)
