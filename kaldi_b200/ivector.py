"""Online i-vector extraction on the GPU (A12-A16 of SURVEY.md §8a) behind the
semantics of online2/online-ivector-feature.cc (OnlineIvectorFeature) with the
defaults of OnlineIvectorExtractionConfig (online-ivector-feature.h:105-111):
use_most_recent_ivector=true, greedy=false, num_gselect=5, min_post=0.025,
posterior_scale=0.1, num_cg_iters=15; max_count as in the online recipes' conf.

No trained extractor exists in the reference tree, so `make_synthetic_extractor`
builds a seeded one with the recipe shapes: 512-Gaussian diagonal UBM over
40-dim LDA features, splice +-3 (280 -> 40 LDA with offset), 100-dim i-vector;
derived quantities as IvectorExtractor::ComputeDerivedVars
(ivector/ivector-extractor.cc:182-218)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib


def make_synthetic_extractor(seed: int = 0, num_gauss: int = 512, feat_dim: int = 40, ivector_dim: int = 100,
                             splice: int = 3, base_dim: int = 40, max_count: float = 100.0) -> dict:
    rng = np.random.default_rng(seed + 4242)
    K = base_dim * (2 * splice + 1)
    # LDA-like transform with offset column; inputs are MFCCs of magnitude ~10-100
    lda = (rng.standard_normal((feat_dim, K)) / np.sqrt(K) * 0.1).astype(np.float32)
    lda_off = (rng.standard_normal(feat_dim) * 0.1).astype(np.float32)
    lda_mat = np.concatenate([lda, lda_off[:, None]], axis=1).astype(np.float32)       # [40 x 281]
    means = (rng.standard_normal((num_gauss, feat_dim)) * 1.0).astype(np.float64)
    var = rng.uniform(0.5, 2.0, (num_gauss, feat_dim)).astype(np.float64)
    weights = rng.uniform(0.5, 1.5, num_gauss)
    weights /= weights.sum()
    inv_vars = (1.0 / var)
    means_invvars = means * inv_vars
    # DiagGmm::ComputeGconsts (gmm/diag-gmm.cc): log w - 0.5*(D log 2pi + sum log var + sum mean^2/var)
    gconsts = np.log(weights) - 0.5 * (feat_dim * np.log(2 * np.pi) + np.log(var).sum(1) + (means * means * inv_vars).sum(1))
    # extractor: M_g [feat_dim x ivector_dim], Sigma_inv_g SPD [feat_dim x feat_dim]
    M = (rng.standard_normal((num_gauss, feat_dim, ivector_dim)) * 0.1).astype(np.float64)
    M[:, :, 0] += means / 10.0 * 0 + rng.standard_normal((num_gauss, feat_dim)) * 0.05
    A = rng.standard_normal((num_gauss, feat_dim, feat_dim)) * 0.1
    sigma_inv = np.einsum("gij,gkj->gik", A, A) + np.eye(feat_dim)[None] * rng.uniform(0.5, 1.5, (num_gauss, 1, 1))
    prior_offset = 10.0
    sigma_inv_m = np.einsum("gij,gjk->gik", sigma_inv, M)                                 # [G, 40, 100]
    U_full = np.einsum("gji,gjk->gik", M, sigma_inv_m)                                    # M^T Sigma_inv M  [G,100,100]
    il = np.tril_indices(ivector_dim)
    U = np.ascontiguousarray(U_full[:, il[0], il[1]])                                     # packed lower triangle [G, 5050]
    return dict(num_gauss=num_gauss, feat_dim=feat_dim, ivector_dim=ivector_dim, splice=splice, base_dim=base_dim,
                lda_mat=lda_mat, ubm_weights=weights.astype(np.float32), gconsts=gconsts.astype(np.float32),
                means_invvars=means_invvars.astype(np.float32), inv_vars=inv_vars.astype(np.float32),
                M=M, sigma_inv=sigma_inv, sigma_inv_m=np.ascontiguousarray(sigma_inv_m), U=U,
                prior_offset=prior_offset, max_count=max_count, num_gselect=5, min_post=0.025, posterior_scale=0.1,
                num_cg_iters=15, cmn_window=600, speaker_frames=600, global_frames=200,
                global_cmvn_stats=_synthetic_global_cmvn(seed, base_dim))


def _synthetic_global_cmvn(seed: int, dim: int) -> np.ndarray:
    rng = np.random.default_rng(seed + 99)
    cnt = 10000.0
    mean = rng.standard_normal(dim) * 3.0
    mean[0] += 60.0
    var = rng.uniform(20.0, 80.0, dim)
    g = np.zeros((2, dim + 1), np.float64)
    g[0, :dim] = mean * cnt
    g[1, :dim] = (var + mean * mean) * cnt
    g[0, dim] = cnt
    return g


def chunk_samples_of(samp_freq: float, chunk_length_secs: float, num_samples: int) -> int:
    """Samples per AcceptWaveform call of online2-wav-nnet3-latgen-faster (:234-240): int32(samp_freq *
    chunk_length_secs) in float32 arithmetic (truncated), 0 becomes 1, chunk_length_secs <= 0 = the whole file."""
    if chunk_length_secs <= 0:
        return int(num_samples)
    c = int(np.float32(samp_freq) * np.float32(chunk_length_secs))
    return 1 if c == 0 else c


def online_ivector_schedule(num_samples: int, chunk_samples: int, frame_length: int, frame_shift: int,
                            num_feature_frames: int, nnet_right_context: int, frames_per_chunk: int, subsampling: int,
                            splice_right: int = 3):
    """For each nnet chunk n, the frame index OnlineIvectorFeature::GetFrame is
    called with when online2-wav-nnet3-latgen-faster feeds `chunk_samples` at a
    time (online2-wav-nnet3-latgen-faster.cc:245-268): the chunk is computed by
    the first AdvanceDecoding after which DecodableNnetLoopedOnlineBase::
    NumFramesReady (decodable-online-looped.cc:56-84) covers it, and it asks for
    frame min(features_ready - 1, ivector_frames_ready - 1)
    (decodable-online-looped.cc:185-193), ivector_frames_ready being smaller by
    the splice right context until the input is finished."""
    T = num_feature_frames
    n_out = (T + subsampling - 1) // subsampling
    n_chunks = (n_out * subsampling + frames_per_chunk - 1) // frames_per_chunk
    sched = []
    fed, done_chunks = 0, 0
    while done_chunks < n_chunks:
        fed = min(fed + chunk_samples, num_samples)
        finished = fed >= num_samples
        ready = 0 if fed < frame_length else 1 + (fed - frame_length) // frame_shift       # snip-edges
        if finished:
            ready = T
            chunks_ready = n_chunks
            iv_frame = T - 1
        else:
            chunks_ready = max(0, ready - nnet_right_context) // frames_per_chunk
            iv_ready = max(0, ready - splice_right)
            iv_frame = min(ready - 1, iv_ready - 1)
        while done_chunks < min(chunks_ready, n_chunks):
            sched.append(iv_frame if iv_frame >= 0 else -1)   # -1: no i-vector frame ready, the chunk's i-vector stays zero
            done_chunks += 1
    return np.asarray(sched, np.int32)


class _IvecCfg(C.Structure):
    _fields_ = [("base_dim", C.c_int32), ("splice_left", C.c_int32), ("splice_right", C.c_int32),
                ("feat_dim", C.c_int32), ("num_gauss", C.c_int32), ("ivector_dim", C.c_int32),
                ("num_gselect", C.c_int32), ("min_post", C.c_float), ("posterior_scale", C.c_float),
                ("max_count", C.c_float), ("prior_offset", C.c_float), ("num_cg_iters", C.c_int32),
                ("cmn_window", C.c_int32), ("speaker_frames", C.c_int32), ("global_frames", C.c_int32),
                ("max_lanes", C.c_int32), ("max_frames", C.c_int32)]


class IvectorExtractorGpu:
    """GPU OnlineIvectorFeature for batches of equal-length utterances."""

    def __init__(self, ex: dict, max_lanes: int, max_frames: int):
        self.ex = ex
        L = _lib.lib()
        c = _IvecCfg(ex["base_dim"], ex["splice"], ex["splice"], ex["feat_dim"], ex["num_gauss"], ex["ivector_dim"],
                     ex["num_gselect"], ex["min_post"], ex["posterior_scale"], ex["max_count"], ex["prior_offset"],
                     ex["num_cg_iters"], ex["cmn_window"], ex["speaker_frames"], ex["global_frames"], max_lanes, max_frames)
        f32p, f64p = C.POINTER(C.c_float), C.POINTER(C.c_double)
        self._keep = [np.ascontiguousarray(ex["lda_mat"], np.float32), np.ascontiguousarray(ex["gconsts"], np.float32),
                      np.ascontiguousarray(ex["means_invvars"], np.float32), np.ascontiguousarray(ex["inv_vars"], np.float32),
                      np.ascontiguousarray(ex["sigma_inv_m"], np.float64), np.ascontiguousarray(ex["U"], np.float64),
                      np.ascontiguousarray(ex["global_cmvn_stats"], np.float64)]
        k = self._keep
        self.h = C.c_void_p()
        _lib.check(L.b2k_ivec_create(C.cast(C.byref(c), C.c_void_p), k[0].ctypes.data_as(f32p), k[1].ctypes.data_as(f32p),
                                     k[2].ctypes.data_as(f32p), k[3].ctypes.data_as(f32p), k[4].ctypes.data_as(f64p),
                                     k[5].ctypes.data_as(f64p), k[6].ctypes.data_as(f64p), C.byref(self.h)))
        self.ivector_dim = ex["ivector_dim"]

    def __del__(self):
        try:
            if self.h:
                _lib.lib().b2k_ivec_destroy(self.h)
        except Exception:
            pass

    def Compute(self, feat_ptrs, feat_stride: int, num_frames: int, schedule, out_ptrs, out_stride: int, stream: int = 0):
        n = len(feat_ptrs)
        sched = np.ascontiguousarray(schedule, np.int32)
        fa = (C.c_void_p * n)(*[int(p) for p in feat_ptrs])
        oa = (C.c_void_p * n)(*[int(p) for p in out_ptrs])
        _lib.check(_lib.lib().b2k_ivec_compute_batched(
            self.h, n, C.cast(fa, C.c_void_p), int(feat_stride), int(num_frames),
            sched.ctypes.data_as(C.POINTER(C.c_int32)), len(sched), C.cast(oa, C.c_void_p), int(out_stride),
            C.c_void_p(stream)))

    def AdaptationStateDoubles(self) -> int:
        """Doubles in one speaker's adaptation state (speaker CMVN stats, then the i-vector stats)."""
        L = _lib.lib()
        L.b2k_ivec_adaptation_state_doubles.restype = C.c_int64
        L.b2k_ivec_adaptation_state_doubles.argtypes = [C.c_void_p]
        return int(L.b2k_ivec_adaptation_state_doubles(self.h))

    def ComputeAdapt(self, feat_ptrs, feat_stride: int, num_frames: int, schedule, out_ptrs, out_stride: int,
                     state_in_ptrs, state_out_ptrs, max_remembered_frames: float = 1000.0, stream: int = 0):
        """Compute with speaker adaptation: state_in_ptrs[i] (0 / None = a new speaker) is what SetAdaptationState would be
        given before the utterance, state_out_ptrs[i] receives GetAdaptationState after it (device pointers to
        AdaptationStateDoubles() doubles; in and out may coincide).  online2-wav-nnet3-latgen-faster.cc:199-221,287."""
        n = len(feat_ptrs)
        sched = np.ascontiguousarray(schedule, np.int32)
        fa = (C.c_void_p * n)(*[int(p) for p in feat_ptrs])
        oa = (C.c_void_p * n)(*[int(p) for p in out_ptrs])
        si = (C.c_void_p * n)(*[int(p) if p else None for p in state_in_ptrs]) if state_in_ptrs is not None else None
        so = (C.c_void_p * n)(*[int(p) if p else None for p in state_out_ptrs]) if state_out_ptrs is not None else None
        L = _lib.lib()
        L.b2k_ivec_compute_batched_adapt.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_int32,
                                                     C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_float, C.c_void_p]
        _lib.check(L.b2k_ivec_compute_batched_adapt(
            self.h, n, C.cast(fa, C.c_void_p), int(feat_stride), int(num_frames), sched.ctypes.data, len(sched),
            C.cast(oa, C.c_void_p), int(out_stride), si, so, float(max_remembered_frames), C.c_void_p(stream)))

    # pipeline hook
    def compute_chunk_ivectors(self, pipe, n: int, stream: int = 0):
        T, D = pipe.T, pipe.feat.dim
        nc, ivd = pipe.nnet.n_chunks, self.ivector_dim
        if getattr(self, "_sched_key", None) != (T, nc):
            fo = pipe.cfg.feature_opts
            self._sched = online_ivector_schedule(
                pipe.cfg.num_samples, chunk_samples_of(fo.samp_freq, pipe.cfg.chunk_length_secs, pipe.cfg.num_samples),
                int(fo.samp_freq * 0.001 * fo.frame_length_ms), int(fo.samp_freq * 0.001 * fo.frame_shift_ms), T,
                pipe.nnet.prog["model_right"], pipe.cfg.frames_per_chunk, pipe.arch["frame_subsampling_factor"],
                self.ex["splice"])
            assert len(self._sched) == nc
            self._sched_key = (T, nc)
        fp = [pipe.d_feats.data_ptr() + 4 * T * D * i for i in range(n)]
        op = [pipe.d_ivec.data_ptr() + 4 * nc * ivd * i for i in range(n)]
        self.Compute(fp, D, T, self._sched, op, ivd, stream)


def extractor_from_kaldi_files(ie_path: str, dubm_path: str, lda_mat_path: str, global_cmvn_path: str,
                               splice: int = 3, base_dim: int = 40, num_gselect: int = 5, min_post: float = 0.025,
                               posterior_scale: float = 0.1, max_count: float = 100.0, num_cg_iters: int = 15,
                               cmn_window: int = 600, speaker_frames: int = 600, global_frames: int = 200) -> dict:
    """The extractor description IvectorExtractorGpu consumes, from the files of an
    ivector_extractor directory (final.ie, final.dubm, final.mat, global_cmvn.stats;
    steps/online/nnet2/train_ivector_extractor.sh) and the options of its
    conf/ivector_extractor.conf (defaults of OnlineIvectorExtractionConfig,
    online2/online-ivector-feature.h:55-140)."""
    from . import kaldi_io as KIO
    ie = KIO.read_ivector_extractor(ie_path)
    ubm = KIO.read_diag_gmm(dubm_path)
    lda_mat = KIO.read_matrix(lda_mat_path).astype(np.float32)
    if ubm["num_gauss"] != ie["num_gauss"] or ubm["feat_dim"] != ie["feat_dim"]:
        raise ValueError("final.dubm and final.ie disagree on the number of Gaussians / feature dimension")
    if lda_mat.shape != (ie["feat_dim"], base_dim * (2 * splice + 1) + 1):
        raise ValueError(f"final.mat is {lda_mat.shape}, expected ({ie['feat_dim']}, {base_dim * (2 * splice + 1) + 1})")
    return dict(num_gauss=ie["num_gauss"], feat_dim=ie["feat_dim"], ivector_dim=ie["ivector_dim"], splice=splice,
                base_dim=base_dim, lda_mat=lda_mat, ubm_weights=ubm["ubm_weights"], gconsts=ubm["gconsts"],
                means_invvars=ubm["means_invvars"], inv_vars=ubm["inv_vars"], M=ie["M"], sigma_inv=ie["sigma_inv"],
                sigma_inv_m=ie["sigma_inv_m"], U=ie["U"], prior_offset=ie["prior_offset"], max_count=max_count,
                num_gselect=num_gselect, min_post=min_post, posterior_scale=posterior_scale, num_cg_iters=num_cg_iters,
                cmn_window=cmn_window, speaker_frames=speaker_frames, global_frames=global_frames,
                global_cmvn_stats=KIO.read_cmvn_stats(global_cmvn_path))


class IvectorFiles:
    """The same four files through the C++ reader of libb2k.so (kaldi_b200/csrc/model_io.cu, b2k_ivec_files_*);
    `IvectorExtractorGpu.from_files` uploads them without going through numpy."""

    def __init__(self, ie_path: str, dubm_path: str, lda_mat_path: str, global_cmvn_path: str):
        L = _lib.lib()
        self.h = C.c_void_p()
        L.b2k_ivec_files_read.argtypes = [C.c_char_p] * 4 + [C.c_void_p]
        _lib.check(L.b2k_ivec_files_read(str(ie_path).encode(), str(dubm_path).encode(), str(lda_mat_path).encode(),
                                         str(global_cmvn_path).encode(), C.byref(self.h)))
        info, po = (C.c_int32 * 8)(), C.c_float()
        L.b2k_ivec_files_info.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        _lib.check(L.b2k_ivec_files_info(self.h, info, C.byref(po)))
        (self.num_gauss, self.feat_dim, self.ivector_dim, self.lda_rows, self.lda_cols, self.cmvn_dim,
         self.num_weights) = [int(x) for x in info][:7]
        self.prior_offset = float(po.value)

    def arrays(self) -> dict:
        L = _lib.lib()
        L.b2k_ivec_files_f32.restype = C.POINTER(C.c_float)
        L.b2k_ivec_files_f32.argtypes = [C.c_void_p, C.c_int32]
        L.b2k_ivec_files_f64.restype = C.POINTER(C.c_double)
        L.b2k_ivec_files_f64.argtypes = [C.c_void_p, C.c_int32]
        G, F, D = self.num_gauss, self.feat_dim, self.ivector_dim

        def f32(i, shape):
            return np.ctypeslib.as_array(L.b2k_ivec_files_f32(self.h, i), shape=shape).copy()

        def f64(i, shape):
            return np.ctypeslib.as_array(L.b2k_ivec_files_f64(self.h, i), shape=shape).copy()
        out = dict(lda_mat=f32(0, (self.lda_rows, self.lda_cols)), gconsts=f32(1, (G,)), means_invvars=f32(2, (G, F)),
                   inv_vars=f32(3, (G, F)), sigma_inv_m=f64(0, (G, F, D)), U=f64(1, (G, D * (D + 1) // 2)),
                   global_cmvn_stats=f64(2, (2, self.cmvn_dim + 1)))
        if self.num_weights:
            out["ubm_weights"] = f32(4, (self.num_weights,))
        return out

    def close(self):
        if getattr(self, "h", None):
            L = _lib.lib()
            L.b2k_ivec_files_destroy.argtypes = [C.c_void_p]
            L.b2k_ivec_files_destroy(self.h)
            self.h = None

    __del__ = close


def _from_files(cls, files: IvectorFiles, max_lanes: int, max_frames: int, splice: int = 3, base_dim: int = 40,
                num_gselect: int = 5, min_post: float = 0.025, posterior_scale: float = 0.1, max_count: float = 100.0,
                num_cg_iters: int = 15, cmn_window: int = 600, speaker_frames: int = 600, global_frames: int = 200):
    """b2k_ivec_create_from_files: options = defaults of OnlineIvectorExtractionConfig
    (online2/online-ivector-feature.h:55-140), everything else from the files."""
    self = cls.__new__(cls)
    c = _IvecCfg(base_dim, splice, splice, files.feat_dim, files.num_gauss, files.ivector_dim, num_gselect, min_post,
                 posterior_scale, max_count, files.prior_offset, num_cg_iters, cmn_window, speaker_frames, global_frames,
                 max_lanes, max_frames)
    self.h = C.c_void_p()
    self.ex = dict(splice=splice, base_dim=base_dim, feat_dim=files.feat_dim, num_gauss=files.num_gauss,
                   ivector_dim=files.ivector_dim)
    self.ivector_dim = files.ivector_dim
    self._keep = []
    L = _lib.lib()
    L.b2k_ivec_create_from_files.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    _lib.check(L.b2k_ivec_create_from_files(C.byref(c), files.h, C.byref(self.h)))
    return self


IvectorExtractorGpu.from_files = classmethod(_from_files)
