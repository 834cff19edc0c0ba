"""Online i-vector oracle — TEST INFRASTRUCTURE ONLY.

`RefIvector`: ctypes view of the reference's own i-vector code compiled in
oracle/_ref (see oracle/ref_wrap/ivector_wrap.cc for exactly which reference
functions run and which ~60 lines of OnlineIvectorFeature glue are restated).
`make_cpu_extractor` builds it from the same synthetic extractor dict the
product uses (kaldi_b200.ivector.make_synthetic_extractor) by writing Kaldi
text-format objects for IvectorExtractor::Read / DiagGmm::Read.

PARITY STATUS: pinned to the compiled reference components (the reference's
ivector-extractor-test only checks self-consistency; no golden vectors exist)."""
from __future__ import annotations

import ctypes as C
import io
import os

import numpy as np

from kaldi_b200 import ivector as IVM

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_ref", "libkaldi_ref_nnet3.so")


def _vec_text(v):
    return " [ " + " ".join(repr(float(x)) for x in v) + " ]\n"


def _mat_text(m):
    buf = io.StringIO()
    buf.write(" [\n")
    for r in m:
        buf.write("  " + " ".join(repr(float(x)) for x in r) + "\n")
    buf.write(" ]\n")
    return buf.getvalue()


def _sp_text(m):
    buf = io.StringIO()
    buf.write(" [\n")
    for i in range(m.shape[0]):
        buf.write("  " + " ".join(repr(float(x)) for x in m[i, :i + 1]) + "\n")
    buf.write(" ]\n")
    return buf.getvalue()


def extractor_text(ex: dict) -> str:
    """IvectorExtractor::Write text layout (ivector/ivector-extractor.cc:807-825); w_ empty."""
    G = ex["num_gauss"]
    parts = ["<IvectorExtractor> <w>  [ ]\n", "<w_vec> " + _vec_text(np.log(ex["ubm_weights"].astype(np.float64))), f"<M> {G}\n"]
    parts += [_mat_text(ex["M"][g]) for g in range(G)]
    parts.append("<SigmaInv>\n")
    parts += [_sp_text(ex["sigma_inv"][g]) for g in range(G)]
    parts.append(f"<IvectorOffset> {ex['prior_offset']!r} </IvectorExtractor>\n")
    return "".join(parts)


def ubm_text(ex: dict) -> str:
    """DiagGmm::Write text layout (gmm/diag-gmm.cc:728-756)."""
    return ("<DiagGMM> <GCONSTS> " + _vec_text(ex["gconsts"]) + "<WEIGHTS> " + _vec_text(ex["ubm_weights"]) +
            "<MEANS_INVVARS> " + _mat_text(ex["means_invvars"]) + "<INV_VARS> " + _mat_text(ex["inv_vars"]) + "</DiagGMM>\n")


class RefIvector:
    def __init__(self, ex: dict):
        self.ex = ex
        if not os.path.exists(_SO):
            from . import ref_nnet
            ref_nnet.build()
        L = self.lib = C.CDLL(_SO)
        L.ref_ivector_create.restype = C.c_void_p
        lda = np.ascontiguousarray(ex["lda_mat"], np.float32)
        self.h = C.c_void_p(L.ref_ivector_create(extractor_text(ex).encode(), ubm_text(ex).encode(),
                                                 lda.ctypes.data_as(C.POINTER(C.c_float)), lda.shape[0], lda.shape[1]))
        if not self.h:
            raise RuntimeError("reference IvectorExtractor/DiagGmm Read failed")

    def __del__(self):
        try:
            self.lib.ref_ivector_destroy(self.h)
        except Exception:
            pass

    def new_speaker(self):
        """An opaque speaker (OnlineIvectorExtractorAdaptationState on the reference side) for run(..., speaker=)."""
        self.lib.ref_ivector_speaker_create.restype = C.c_void_p
        return C.c_void_p(self.lib.ref_ivector_speaker_create())

    def speaker_state(self, speaker) -> np.ndarray:
        """The speaker's adaptation state as doubles in b2k's layout (b2k_ivec_compute_batched_adapt)."""
        D, iv = self.ex["base_dim"], self.ex["ivector_dim"]
        out = np.zeros(2 * (D + 1) + 1 + iv + iv * (iv + 1) // 2, np.float64)
        self.lib.ref_ivector_speaker_state.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        if self.lib.ref_ivector_speaker_state(speaker, D, iv, out.ctypes.data) != 0:
            raise RuntimeError("speaker without a state")
        return out

    def run(self, feats: np.ndarray, schedule, online_cmvn_iextractor: bool = False, debug: bool = False, speaker=None,
            max_remembered_frames: float = 1000.0):
        ex = self.ex
        f = np.ascontiguousarray(feats, np.float32)
        T, D = f.shape
        sched = np.ascontiguousarray(schedule, np.int32)
        out = np.zeros((len(sched), ex["ivector_dim"]), np.float32)
        g = np.ascontiguousarray(ex["global_cmvn_stats"], np.float64)
        dr = np.zeros((T, ex["feat_dim"]), np.float32) if debug else None
        dn = np.zeros((T, ex["feat_dim"]), np.float32) if debug else None
        fp, dp, ip = C.POINTER(C.c_float), C.POINTER(C.c_double), C.POINTER(C.c_int32)
        r = self.lib.ref_ivector_run_speaker(self.h, f.ctypes.data_as(fp), T, D, g.ctypes.data_as(dp), ex["cmn_window"],
                                             ex["speaker_frames"], ex["global_frames"], ex["splice"], ex["splice"],
                                             ex["num_gselect"], C.c_float(ex["min_post"]), C.c_float(ex["posterior_scale"]),
                                             C.c_float(ex["max_count"]), ex["num_cg_iters"], int(online_cmvn_iextractor),
                                             sched.ctypes.data_as(ip), len(sched), out.ctypes.data_as(fp),
                                             dr.ctypes.data_as(fp) if debug else None, dn.ctypes.data_as(fp) if debug else None,
                                             speaker, C.c_float(max_remembered_frames))
        if r != 0:
            raise RuntimeError("reference i-vector extraction failed")
        return (out, dr, dn) if debug else out

    def run_real(self, feats: np.ndarray, schedule, online_cmvn_iextractor: bool = False, speaker=None,
                 max_remembered_frames: float = 1000.0, delta_weights=None):
        """The same through the reference's OWN OnlineIvectorFeature (oracle/ref_wrap/silence_wrap.cc: ref_ivector_run_real), which
        run()'s glue restates.  delta_weights: per chunk a list of (frame, weight difference) handed to UpdateFrameWeights before that
        chunk's GetFrame (the silence weighting of the online2 tools); None = unweighted."""
        ex = self.ex
        if not hasattr(self.lib, "ref_ivector_run_real"):
            raise RuntimeError("oracle/_ref predates silence_wrap.cc")
        f = np.ascontiguousarray(feats, np.float32)
        T, D = f.shape
        sched = np.ascontiguousarray(schedule, np.int32)
        out = np.zeros((len(sched), ex["ivector_dim"]), np.float32)
        g = np.ascontiguousarray(ex["global_cmvn_stats"], np.float64)
        fp, dp, ip = C.POINTER(C.c_float), C.POINTER(C.c_double), C.POINTER(C.c_int32)
        off = fr = w = None
        if delta_weights is not None:
            assert len(delta_weights) == len(sched)
            off = np.zeros(len(sched) + 1, np.int32)
            off[1:] = np.cumsum([len(d) for d in delta_weights])
            flat = [p for d in delta_weights for p in d]
            fr = np.array([p[0] for p in flat] + [0], np.int32)
            w = np.array([p[1] for p in flat] + [0.0], np.float32)
        self.lib.ref_ivector_run_real.argtypes = None
        r = self.lib.ref_ivector_run_real(self.h, f.ctypes.data_as(fp), T, D, g.ctypes.data_as(dp), ex["cmn_window"],
                                          ex["speaker_frames"], ex["global_frames"], ex["splice"], ex["splice"],
                                          ex["num_gselect"], C.c_float(ex["min_post"]), C.c_float(ex["posterior_scale"]),
                                          C.c_float(ex["max_count"]), ex["num_cg_iters"], int(online_cmvn_iextractor),
                                          sched.ctypes.data_as(ip), len(sched), out.ctypes.data_as(fp),
                                          speaker, C.c_float(max_remembered_frames),
                                          off.ctypes.data_as(ip) if off is not None else None,
                                          fr.ctypes.data_as(ip) if fr is not None else None,
                                          w.ctypes.data_as(fp) if w is not None else None)
        if r != 0:
            raise RuntimeError("reference OnlineIvectorFeature failed")
        return out

    # used by bench.py's CPU arm
    def chunk_ivectors(self, feats, n_chunks, frames_per_chunk, right_context):
        sched = IVM.online_ivector_schedule(160000 if feats.shape[0] == 998 else (feats.shape[0] - 1) * 160 + 400, 2880, 400, 160,
                                            feats.shape[0], right_context, frames_per_chunk, 3, self.ex["splice"])
        assert len(sched) == n_chunks
        return self.run(feats, sched)


def make_cpu_extractor(seed: int = 0) -> RefIvector:
    return RefIvector(IVM.make_synthetic_extractor(seed))
