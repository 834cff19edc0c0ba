"""Host-side mirror of the reference's batched GPU feature surface over the
b2k C-ABI: cudafeat/online-batched-feature-pipeline-cuda.h:44-134
(OnlineBatchedFeaturePipelineCuda::ComputeFeaturesBatched) and the options of
feat/feature-mfcc.h / feature-fbank.h / online-feature.h (OnlineCmvnOptions)."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, fields

import numpy as np

from . import _lib


class _FeatCfg(C.Structure):
    _fields_ = [("feature_type", C.c_int32), ("samp_freq", C.c_float), ("frame_shift_ms", C.c_float),
                ("frame_length_ms", C.c_float), ("dither", C.c_float), ("preemph_coeff", C.c_float),
                ("remove_dc_offset", C.c_int32), ("round_to_power_of_two", C.c_int32), ("snip_edges", C.c_int32),
                ("window_type", C.c_int32), ("num_bins", C.c_int32), ("low_freq", C.c_float),
                ("high_freq", C.c_float), ("num_ceps", C.c_int32), ("use_energy", C.c_int32),
                ("energy_floor", C.c_float), ("raw_energy", C.c_int32), ("cepstral_lifter", C.c_float),
                ("htk_compat", C.c_int32), ("use_log_fbank", C.c_int32), ("use_power", C.c_int32),
                ("htk_mode", C.c_int32), ("max_lanes", C.c_int32),
                ("lpc_order", C.c_int32), ("compress_factor", C.c_float), ("cepstral_scale", C.c_float)]


class _CmvnCfg(C.Structure):
    _fields_ = [("cmn_window", C.c_int32), ("speaker_frames", C.c_int32), ("global_frames", C.c_int32),
                ("normalize_mean", C.c_int32), ("normalize_variance", C.c_int32)]


@dataclass
class FeatureOptions:
    """MfccOptions / FbankOptions with their FrameExtractionOptions and
    MelBanksOptions; defaults = mfcc_hires.conf of the named recipes, dither 0."""
    feature_type: int = 0
    samp_freq: float = 16000.0
    frame_shift_ms: float = 10.0
    frame_length_ms: float = 25.0
    dither: float = 0.0
    preemph_coeff: float = 0.97
    remove_dc_offset: int = 1
    round_to_power_of_two: int = 1
    snip_edges: int = 1
    window_type: int = 0
    num_bins: int = 40
    low_freq: float = 20.0
    high_freq: float = -400.0
    num_ceps: int = 40
    use_energy: int = 0
    energy_floor: float = 0.0
    raw_energy: int = 1
    cepstral_lifter: float = 22.0
    htk_compat: int = 0
    use_log_fbank: int = 1
    use_power: int = 1
    htk_mode: int = 0
    max_lanes: int = 1024
    lpc_order: int = 12              # PlpOptions (feature_type 2), feat/feature-plp.h:38-66
    compress_factor: float = 0.33333
    cepstral_scale: float = 1.0


@dataclass
class OnlineCmvnOptions:
    """feat/online-feature.h:203-227"""
    cmn_window: int = 600
    speaker_frames: int = 600
    global_frames: int = 200
    normalize_mean: bool = True
    normalize_variance: bool = False


def _ptr_array(ptrs):
    arr = (C.c_void_p * len(ptrs))(*[int(p) for p in ptrs])
    return C.cast(arr, C.c_void_p), arr


class BatchedFeatures:
    """Batched spectral features + online CMVN on the GPU."""

    def __init__(self, opts: FeatureOptions):
        L = _lib.lib()
        self.opts = opts
        c = _FeatCfg(**{f.name: getattr(opts, f.name) for f in fields(opts)})
        self.h = C.c_void_p()
        _lib.check(L.b2k_feat_create(C.cast(C.byref(c), C.c_void_p), C.byref(self.h)))
        self.dim = int(L.b2k_feat_dim(self.h))

    def __del__(self):
        try:
            if self.h:
                _lib.lib().b2k_feat_destroy(self.h)
        except Exception:
            pass

    def Dim(self) -> int:
        return self.dim

    def NumFrames(self, num_samples: int, flush: bool = True) -> int:
        return int(_lib.lib().b2k_feat_num_frames(self.h, int(num_samples), int(flush)))

    def ComputeFeaturesBatched(self, wave_ptrs, num_samples, first_frame, num_frames, out_ptrs,
                               row_stride: int, stream: int = 0, int16: bool = False):
        """int16 = True: wave_ptrs are device pointers to 16-bit PCM (b2k_feat_compute_batched_i16)."""
        n = len(wave_ptrs)
        ns = np.ascontiguousarray(num_samples, np.int32)
        ff = np.ascontiguousarray(first_frame, np.int32)
        nf = np.ascontiguousarray(num_frames, np.int32)
        wp, _k1 = _ptr_array(wave_ptrs)
        op, _k2 = _ptr_array(out_ptrs)
        i32p = C.POINTER(C.c_int32)
        L = _lib.lib()
        L.b2k_feat_compute_batched_i16.argtypes = L.b2k_feat_compute_batched.argtypes
        fn = L.b2k_feat_compute_batched_i16 if int16 else L.b2k_feat_compute_batched
        _lib.check(fn(self.h, n, wp, ns.ctypes.data_as(i32p), ff.ctypes.data_as(i32p), nf.ctypes.data_as(i32p),
                      op, int(row_stride), C.c_void_p(stream)))

    def ApplyCmvnBatched(self, cmvn: OnlineCmvnOptions, in_ptrs, out_ptrs, in_stride, out_stride,
                         first_frame, num_frames, state_ptrs, global_stats_ptr, speaker_stats_ptr=0,
                         stream: int = 0):
        n = len(in_ptrs)
        ff = np.ascontiguousarray(first_frame, np.int32)
        nf = np.ascontiguousarray(num_frames, np.int32)
        c = _CmvnCfg(cmvn.cmn_window, cmvn.speaker_frames, cmvn.global_frames, int(cmvn.normalize_mean),
                     int(cmvn.normalize_variance))
        ip, _k1 = _ptr_array(in_ptrs)
        op, _k2 = _ptr_array(out_ptrs)
        sp, _k3 = _ptr_array(state_ptrs)
        i32p = C.POINTER(C.c_int32)
        _lib.check(_lib.lib().b2k_cmvn_apply_batched(
            self.h, C.cast(C.byref(c), C.c_void_p), n, ip, op, int(in_stride), int(out_stride),
            ff.ctypes.data_as(i32p), nf.ctypes.data_as(i32p), sp, C.c_void_p(int(global_stats_ptr)),
            C.c_void_p(int(speaker_stats_ptr)) if speaker_stats_ptr else None, C.c_void_p(stream)))

    # convenience for tests: whole utterances given as numpy arrays
    def compute(self, waves, chunk_frames: int | None = None):
        import torch
        d_w = [torch.from_numpy(np.ascontiguousarray(w, np.float32)).cuda() for w in waves]
        ns = [int(w.numel()) for w in d_w]
        T = [self.NumFrames(n) for n in ns]
        outs = [torch.zeros(max(t, 1), self.dim, device="cuda") for t in T]
        if chunk_frames is None:
            self.ComputeFeaturesBatched([w.data_ptr() for w in d_w], ns, [0] * len(ns), T,
                                        [o.data_ptr() for o in outs], self.dim)
        else:
            done = [0] * len(ns)
            while any(d < t for d, t in zip(done, T)):
                nf = [min(chunk_frames, t - d) for d, t in zip(done, T)]
                self.ComputeFeaturesBatched([w.data_ptr() for w in d_w], ns, done, nf,
                                            [o.data_ptr() for o in outs], self.dim)
                done = [d + k for d, k in zip(done, nf)]
        torch.cuda.synchronize()
        return [o[:t].cpu().numpy() for o, t in zip(outs, T)]
