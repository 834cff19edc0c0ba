"""Batched online2 inference pipeline: features -> [i-vector] -> nnet3 -> decoder.

Mirrors cuda_decoder::BatchedThreadedNnet3CudaOnlinePipeline::DecodeBatch
(cudadecoder/batched-threaded-nnet3-cuda-online-pipeline.cc:316,377:
ComputeGPUFeatureExtraction -> RunNnet3 -> RunDecoder -> finalize) but keeps the
numerical semantics of the CPU tool online2-wav-nnet3-latgen-faster
(online2bin/online2-wav-nnet3-latgen-faster.cc:199-299): per-chunk i-vectors as
DecodableNnetLoopedOnlineBase::AdvanceChunk would have received them, CPU
decoder search semantics, finalized raw lattice per utterance.

Everything between the host waveform buffers and the host lattice arrays runs
on the GPU through the b2k C-ABI; torch is used only for device buffers, pinned
staging and streams.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

from . import nnet_model as NM
from . import synth
from .decoder import CudaDecoder, CudaDecoderConfig, CudaFst
from .feat import BatchedFeatures, FeatureOptions
from .nnet import NnetComputer


@dataclass
class PipelineConfig:
    feature_opts: FeatureOptions = field(default_factory=FeatureOptions)
    decoder_cfg: dict = field(default_factory=lambda: dict(synth.DEFAULT_DECODER_CFG))
    frames_per_chunk: int = 21          # --frames-per-chunk=20 rounded up to a multiple of 3 (GetChunkSize)
    acoustic_scale: float = 1.0         # chain models decode with --acwt 1.0
    max_batch: int = 64
    num_samples: int = 160000           # utterances are padded/truncated to this many samples per batch
    reference_order: bool = True
    chunk_length_secs: float = 0.18     # the CPU tool's --chunk-length (drives which i-vector a chunk sees)
    frame_subsampling_factor: int = 0   # --frame-subsampling-factor; 0 = the model's own when its layers decide it
    extract_ivectors: bool = True
    max_tokens: int = 0                 # 0 = sized from the utterance length
    max_links: int = 0
    max_tokens_per_frame: int = 32768   # sizes the per-lane token hash (2x this many slots) and per-frame scratch


class BatchedPipeline:
    def __init__(self, cfg: PipelineConfig, arch: dict, weights: dict, graph: dict, ivector_extractor=None):
        import torch
        self.torch = torch
        self.cfg = cfg
        self.arch = arch
        self.weights = weights
        self.feat = BatchedFeatures(cfg.feature_opts)
        self.T = self.feat.NumFrames(cfg.num_samples)
        self.nnet = NnetComputer(arch, weights, self.T, cfg.max_batch, cfg.frames_per_chunk, cfg.acoustic_scale)
        self.fst = CudaFst(graph)
        nf = self.nnet.n_out
        dc = CudaDecoderConfig.from_dict(
            cfg.decoder_cfg, max_frames=nf + 2,
            max_tokens=cfg.max_tokens or int(nf * 9000), max_links=cfg.max_links or int(nf * 16000),
            reference_order=cfg.reference_order, max_tokens_per_frame=cfg.max_tokens_per_frame)
        self.dec = CudaDecoder(self.fst, dc, cfg.max_batch)
        B = cfg.max_batch
        self.h_wave = torch.empty(B, cfg.num_samples, dtype=torch.float32).pin_memory()
        self.d_wave = torch.empty(B, cfg.num_samples, dtype=torch.float32, device="cuda")
        self.d_feats = torch.empty(B, self.T, self.feat.dim, dtype=torch.float32, device="cuda")
        self.d_ivec = torch.zeros(B, self.nnet.n_chunks, arch["ivector_dim"], dtype=torch.float32, device="cuda")
        self.d_loglikes = torch.empty(B, nf, arch["num_pdfs"], dtype=torch.float32, device="cuda")
        if isinstance(ivector_extractor, dict):      # a synthetic/loaded extractor description
            from .ivector import IvectorExtractorGpu
            ivector_extractor = IvectorExtractorGpu(ivector_extractor, cfg.max_batch, self.T)
        self.ivector_extractor = ivector_extractor
        self.audio_seconds_per_utt = cfg.num_samples / cfg.feature_opts.samp_freq

    # ---- stages (device resident) ------------------------------------------------
    def compute_features(self, n: int, stream: int = 0):
        S, T, D = self.cfg.num_samples, self.T, self.feat.dim
        wp = [self.d_wave.data_ptr() + 4 * S * i for i in range(n)]
        op = [self.d_feats.data_ptr() + 4 * T * D * i for i in range(n)]
        self.feat.ComputeFeaturesBatched(wp, [S] * n, [0] * n, [T] * n, op, D, stream)

    def compute_ivectors(self, n: int, stream: int = 0):
        if self.ivector_extractor is not None and self.cfg.extract_ivectors:
            self.ivector_extractor.compute_chunk_ivectors(self, n, stream)

    def compute_nnet(self, n: int, stream: int = 0):
        T, D, P, nf = self.T, self.feat.dim, self.arch["num_pdfs"], self.nnet.n_out
        ivd, nc = self.arch["ivector_dim"], self.nnet.n_chunks
        ip = [self.d_feats.data_ptr() + 4 * T * D * i for i in range(n)]
        vp = [self.d_ivec.data_ptr() + 4 * nc * ivd * i for i in range(n)]
        op = [self.d_loglikes.data_ptr() + 4 * nf * P * i for i in range(n)]
        self.nnet.Run(ip, D, vp, ivd, op, P, stream)

    def decode(self, n: int, stream: int = 0):
        P, nf = self.arch["num_pdfs"], self.nnet.n_out
        ch = list(range(n))
        self.dec.InitDecoding(ch, stream)
        lp = [self.d_loglikes.data_ptr() + 4 * nf * P * i for i in range(n)]
        self.dec.AdvanceDecodingFrames(ch, lp, [nf] * n, P, stream)
        self.dec.FinalizeDecoding(ch, stream)

    def run_device(self, n: int, stream: int = 0):
        """All four stages with inputs already resident in d_wave."""
        self.compute_features(n, stream)
        self.compute_ivectors(n, stream)
        self.compute_nnet(n, stream)
        self.decode(n, stream)

    def output_calibration(self, waves, target_std: float = 1.0):
        """(per-pdf mean of the raw nnet output, scale) over a calibration batch,
        for nnet_model.apply_output_calibration (synthetic models only)."""
        torch = self.torch
        n = len(waves)
        for i, w in enumerate(waves):
            self.h_wave[i].copy_(torch.from_numpy(np.ascontiguousarray(w, dtype=np.float32)))
        self.d_wave[:n].copy_(self.h_wave[:n])
        self.compute_features(n)
        self.compute_ivectors(n)
        self.compute_nnet(n)
        torch.cuda.synchronize()
        ll = self.d_loglikes[:n].double()
        raw = ll / self.cfg.acoustic_scale + torch.log(torch.from_numpy(self.weights["priors"]).double().cuda())[None, None, :]
        mean = raw.mean(dim=(0, 1))
        resid_std = float((raw - mean[None, None, :]).std(dim=2).mean())
        return mean.float().cpu().numpy(), target_std / max(resid_std, 1e-6)

    # ---- public entry: host buffers in, host lattices out -------------------------
    def decode_batch(self, waves, want_lattices: bool = True):
        """waves: list of 1-D float32/int16 arrays (Kaldi int16-range convention),
        each exactly cfg.num_samples long.  Returns the finalized raw lattices of
        the batch packed in one dict (see CudaDecoder.GetRawLattices / SplitLattices)."""
        import time
        torch = self.torch
        n = len(waves)
        assert 0 < n <= self.cfg.max_batch
        t0 = time.perf_counter()
        hw = self.h_wave.numpy()
        for i, w in enumerate(waves):
            assert len(w) == self.cfg.num_samples, "utterances are bucketed by length before batching"
            hw[i] = w                                   # numpy converts int16/float64 input to float32 in place
        t1 = time.perf_counter()
        self.d_wave[:n].copy_(self.h_wave[:n], non_blocking=True)
        self.run_device(n)
        t2 = time.perf_counter()
        if want_lattices:
            # one pack kernel + one D2H for the whole batch; CudaDecoder.SplitLattices gives per-utterance views
            out = self.dec.GetRawLattices(list(range(n)))
        else:
            out = [self.dec.ChannelInfo(c) for c in range(n)]
        t3 = time.perf_counter()
        # host-side wall time of the three parts (the last one includes waiting for the GPU)
        self.last_host_ms = dict(stage_to_pinned=(t1 - t0) * 1e3, submit=(t2 - t1) * 1e3, wait_and_readback=(t3 - t2) * 1e3)
        return out


# ---- the same pipeline in C++ (kaldi_b200/csrc/pipeline.cu, include/b2k.h b2k_pipeline_*) ---------------------------

def _native_structs():
    import ctypes as C
    from .decoder import _DecCfg
    from .feat import _CmvnCfg, _FeatCfg

    class _PipelineCfg(C.Structure):
        _fields_ = [("feat", _FeatCfg), ("dec", _DecCfg), ("frames_per_chunk", C.c_int32), ("acoustic_scale", C.c_float),
                    ("max_batch", C.c_int32), ("num_samples", C.c_int64), ("chunk_length_secs", C.c_float),
                    ("ivector_splice_right", C.c_int32), ("use_priors", C.c_int32), ("conv_dense", C.c_int32),
                    ("use_cmvn", C.c_int32), ("cmvn", _CmvnCfg), ("global_cmvn_stats", C.c_void_p),
                    ("frame_subsampling_factor", C.c_int32)]

    class _PipelinePlan(C.Structure):
        _fields_ = [("num_feature_frames", C.c_int32), ("feat_dim", C.c_int32), ("num_output_frames", C.c_int32),
                    ("num_chunks", C.c_int32), ("num_pdfs", C.c_int32), ("ivector_dim", C.c_int32),
                    ("model_right_context", C.c_int32), ("chunk_samples", C.c_int32), ("dec", _DecCfg),
                    ("device_bytes", C.c_int64), ("pinned_bytes", C.c_int64)]
    return _PipelineCfg, _PipelinePlan


def native_cfg(cfg: PipelineConfig, ivector_splice_right: int = 3):
    """PipelineConfig -> b2k_pipeline_cfg (defaults from b2k_pipeline_cfg_default, then every field of cfg)."""
    import ctypes as C
    from dataclasses import fields
    from . import _lib
    PC, _ = _native_structs()
    L = _lib.lib()
    c = PC()
    L.b2k_pipeline_cfg_default.argtypes = [C.c_void_p]
    L.b2k_pipeline_cfg_default.restype = None
    L.b2k_pipeline_cfg_default(C.byref(c))
    for f in fields(cfg.feature_opts):
        setattr(c.feat, f.name, getattr(cfg.feature_opts, f.name))
    d = cfg.decoder_cfg
    c.dec.beam, c.dec.lattice_beam, c.dec.max_active, c.dec.min_active = d["beam"], d["lattice_beam"], d["max_active"], d["min_active"]
    c.dec.beam_delta, c.dec.prune_interval, c.dec.prune_scale = d["beam_delta"], d["prune_interval"], d["prune_scale"]
    c.dec.hash_ratio = d.get("hash_ratio", 2.0)
    c.dec.reference_order = int(cfg.reference_order)
    c.dec.max_tokens_per_frame = cfg.max_tokens_per_frame
    c.dec.max_tokens, c.dec.max_links, c.dec.max_frames = cfg.max_tokens, cfg.max_links, 0
    c.frames_per_chunk, c.acoustic_scale, c.max_batch = cfg.frames_per_chunk, cfg.acoustic_scale, cfg.max_batch
    c.num_samples, c.chunk_length_secs, c.ivector_splice_right = cfg.num_samples, cfg.chunk_length_secs, ivector_splice_right
    c.frame_subsampling_factor = cfg.frame_subsampling_factor
    return c


def native_plan(cfg: PipelineConfig, model, ivector_splice_right: int = 3) -> dict:
    """b2k_pipeline_plan_for: the sizes the C++ pipeline derives for (cfg, model); needs no device."""
    import ctypes as C
    from . import _lib
    _, PP = _native_structs()
    L = _lib.lib()
    c, pl = native_cfg(cfg, ivector_splice_right), PP()
    L.b2k_pipeline_plan_for.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    _lib.check(L.b2k_pipeline_plan_for(C.byref(c), model.h, C.byref(pl)))
    out = {f[0]: getattr(pl, f[0]) for f in pl._fields_ if f[0] != "dec"}
    out["dec"] = {f[0]: getattr(pl.dec, f[0]) for f in pl.dec._fields_}
    return out


class NativeBatchedPipeline:
    """BatchedPipeline with the orchestration in C++: model.KaldiModel (a model file) + CudaFst (+ IvectorExtractorGpu)
    -> b2k_pipeline_create; decode_batch(waves) -> b2k_pipeline_decode_batch + b2k_pipeline_get_raw_lattices.  Returns
    the same packed dictionary as BatchedPipeline.decode_batch."""

    def __init__(self, cfg: PipelineConfig, model, fst: CudaFst, ivector_extractor=None):
        import ctypes as C
        from . import _lib
        self.cfg, self.model, self.fst, self.ivector_extractor = cfg, model, fst, ivector_extractor
        L = self._L = _lib.lib()
        splice = ivector_extractor.ex["splice"] if ivector_extractor is not None else 3
        c = native_cfg(cfg, splice)
        self.h = C.c_void_p()
        L.b2k_pipeline_create.argtypes = [C.c_void_p] * 5
        _lib.check(L.b2k_pipeline_create(C.byref(c), model.h, fst.h,
                                         ivector_extractor.h if ivector_extractor is not None else None, C.byref(self.h)))
        _, PP = _native_structs()
        pl = PP()
        L.b2k_pipeline_get_plan.argtypes = [C.c_void_p, C.c_void_p]
        _lib.check(L.b2k_pipeline_get_plan(self.h, C.byref(pl)))
        self.plan = pl
        self.audio_seconds_per_utt = cfg.num_samples / cfg.feature_opts.samp_freq

    def __del__(self):
        try:
            if self.h:
                import ctypes as C
                self._L.b2k_pipeline_destroy.argtypes = [C.c_void_p]
                self._L.b2k_pipeline_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def set_speaker_states(self, state_in_ptrs, state_out_ptrs, max_remembered_frames: float = 1000.0):
        """Speaker adaptation of the NEXT batch (b2k_pipeline_set_speaker_states): device pointers (0 / None = a new speaker /
        state not kept) to IvectorExtractorGpu.AdaptationStateDoubles() doubles per batch slot."""
        import ctypes as C
        from . import _lib
        n = len(state_in_ptrs)
        assert len(state_out_ptrs) == n
        si = (C.c_void_p * n)(*[int(p) if p else None for p in state_in_ptrs])
        so = (C.c_void_p * n)(*[int(p) if p else None for p in state_out_ptrs])
        self._L.b2k_pipeline_set_speaker_states.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_float]
        _lib.check(self._L.b2k_pipeline_set_speaker_states(self.h, n, si, so, float(max_remembered_frames)))

    def decode_batch(self, waves, stream: int = 0):
        import ctypes as C
        from . import _lib
        from .decoder import _RawLattice, _p
        L = self._L
        n = len(waves)
        i16 = all(getattr(w, "dtype", None) == np.int16 for w in waves)
        keep = [np.ascontiguousarray(w, np.int16 if i16 else np.float32) for w in waves]
        for w in keep:
            assert w.ndim == 1 and len(w) == self.cfg.num_samples, "utterances are bucketed by length before batching"
        ptrs = (C.c_void_p * n)(*[w.ctypes.data for w in keep])
        fn = L.b2k_pipeline_decode_batch_i16 if i16 else L.b2k_pipeline_decode_batch
        fn.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]
        _lib.check(fn(self.h, n, ptrs, C.c_void_p(stream)))
        so = np.zeros(n + 1, np.int64); ao = np.zeros(n + 1, np.int64); fo = np.zeros(n + 1, np.int64)
        i64p = C.POINTER(C.c_int64)
        r = _RawLattice()
        L.b2k_pipeline_get_raw_lattices.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, i64p, i64p, i64p, C.c_void_p]
        args = (self.h, n, C.cast(C.byref(r), C.c_void_p), so.ctypes.data_as(i64p), ao.ctypes.data_as(i64p),
                fo.ctypes.data_as(i64p), C.c_void_p(stream))
        _lib.check(L.b2k_pipeline_get_raw_lattices(*args))          # sizes
        ns, na, nf = r.num_states, r.num_arcs, r.num_finals
        out = dict(
            state_frame=np.zeros(ns, np.int32), state_hclg=np.zeros(ns, np.int32),
            state_tot_cost=np.zeros(ns, np.float32), state_extra_cost=np.zeros(ns, np.float32),
            arc_src=np.zeros(na, np.int32), arc_dst=np.zeros(na, np.int32),
            arc_ilabel=np.zeros(na, np.int32), arc_olabel=np.zeros(na, np.int32),
            arc_graph_cost=np.zeros(na, np.float32), arc_acoustic_cost=np.zeros(na, np.float32),
            final_state=np.zeros(nf, np.int32), final_cost=np.zeros(nf, np.float32))
        for k, v in out.items():
            setattr(r, k, _p(v, C.c_float if v.dtype == np.float32 else C.c_int32))
        _lib.check(L.b2k_pipeline_get_raw_lattices(*args))
        out.update(state_offs=so, arc_offs=ao, final_offs=fo)
        return out

    # ---- pipelined form: submit(k+1) before collect(k) overlaps batch k+1's staging / copy with batch k's kernels and
    #      batch k's lattice read-back with batch k+1's kernels (b2k_pipeline_submit_i16 / b2k_pipeline_collect)
    def submit(self, waves_i16, ptrs=None):
        """waves_i16: a [n x num_samples] int16 array (or a list of int16 rows).  Returns immediately."""
        import ctypes as C
        from . import _lib
        L = self._L
        if ptrs is None:
            rows = [np.ascontiguousarray(w, np.int16) for w in waves_i16]
            for w in rows:
                assert w.ndim == 1 and len(w) == self.cfg.num_samples, "utterances are bucketed by length before batching"
            self._keep = rows
            ptrs = (C.c_void_p * len(rows))(*[w.ctypes.data for w in rows])
        L.b2k_pipeline_submit_i16.argtypes = [C.c_void_p, C.c_int32, C.c_void_p]
        _lib.check(L.b2k_pipeline_submit_i16(self.h, len(ptrs), ptrs))

    @staticmethod
    def row_pointers(waves_i16: np.ndarray):
        """ctypes pointer table of a C-contiguous [n x num_samples] int16 array (build once, reuse per step)."""
        import ctypes as C
        assert waves_i16.dtype == np.int16 and waves_i16.flags["C_CONTIGUOUS"] and waves_i16.ndim == 2
        base, stride = waves_i16.ctypes.data, waves_i16.strides[0]
        return (C.c_void_p * waves_i16.shape[0])(*[base + i * stride for i in range(waves_i16.shape[0])])

    def collect(self, copy: bool = True):
        """The oldest outstanding batch's raw lattices (same dictionary as decode_batch).  copy=False returns views of
        the pipeline's own arrays, valid until the next collect."""
        import ctypes as C
        from . import _lib
        from .decoder import _RawLattice
        L = self._L
        r = _RawLattice()
        n = C.c_int32()
        i64p = C.POINTER(C.c_int64)
        so, ao, fo = i64p(), i64p(), i64p()
        L.b2k_pipeline_collect.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        _lib.check(L.b2k_pipeline_collect(self.h, C.byref(n), C.byref(r), C.byref(so), C.byref(ao), C.byref(fo)))
        ns, na, nf, nn = r.num_states, r.num_arcs, r.num_finals, n.value

        def arr(ptr, count, dt):
            if count == 0:
                return np.zeros(0, dt)
            a = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_float if dt == np.float32 else C.c_int64 if dt == np.int64 else C.c_int32)), shape=(count,))
            return a.copy() if copy else a
        out = dict(
            state_frame=arr(r.state_frame, ns, np.int32), state_hclg=arr(r.state_hclg, ns, np.int32),
            state_tot_cost=arr(r.state_tot_cost, ns, np.float32), state_extra_cost=arr(r.state_extra_cost, ns, np.float32),
            arc_src=arr(r.arc_src, na, np.int32), arc_dst=arr(r.arc_dst, na, np.int32),
            arc_ilabel=arr(r.arc_ilabel, na, np.int32), arc_olabel=arr(r.arc_olabel, na, np.int32),
            arc_graph_cost=arr(r.arc_graph_cost, na, np.float32), arc_acoustic_cost=arr(r.arc_acoustic_cost, na, np.float32),
            final_state=arr(r.final_state, nf, np.int32), final_cost=arr(r.final_cost, nf, np.float32),
            state_offs=arr(so, nn + 1, np.int64), arc_offs=arr(ao, nn + 1, np.int64), final_offs=arr(fo, nn + 1, np.int64))
        return out

    # ---- multi-GPU shards (kaldi_b200/ingest.py): device int16 in, device-packed lattices out
    def run_device_i16(self, d_ptr: int, n: int, stream: int = 0):
        import ctypes as C
        from . import _lib
        L = self._L
        L.b2k_pipeline_run_device_i16.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]
        _lib.check(L.b2k_pipeline_run_device_i16(self.h, n, C.c_void_p(d_ptr), C.c_void_p(stream)))

    def pack_device(self, n: int, d_buf_ptr: int, cap_bytes: int, stream: int = 0):
        import ctypes as C
        from . import _lib
        L = self._L
        L.b2k_pipeline_pack_device.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_int64, C.c_void_p]
        _lib.check(L.b2k_pipeline_pack_device(self.h, n, C.c_void_p(d_buf_ptr), cap_bytes, C.c_void_p(stream)))

    @staticmethod
    def packed_bytes_needed(header: np.ndarray, n: int) -> tuple[int, int]:
        """(status, bytes) from a host copy of a packed buffer's header (int64 view)."""
        tail = header[3 * (n + 1):]
        return int(tail[0]), int(tail[1])

    @staticmethod
    def unpack_lattices(h_buf: np.ndarray, n: int) -> dict:
        """b2k_dec_unpack_lattices on a host copy (uint8 array) of a packed buffer: the same dictionary as decode_batch."""
        import ctypes as C
        from . import _lib
        from .decoder import _RawLattice, _p
        L = _lib.lib()
        h_buf = np.ascontiguousarray(h_buf, np.uint8)
        r = _RawLattice()
        so = np.zeros(n + 1, np.int64); ao = np.zeros(n + 1, np.int64); fo = np.zeros(n + 1, np.int64)
        i64p = C.POINTER(C.c_int64)
        L.b2k_dec_unpack_lattices.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, i64p, i64p, i64p]
        args = (h_buf.ctypes.data, n, C.cast(C.byref(r), C.c_void_p), so.ctypes.data_as(i64p), ao.ctypes.data_as(i64p), fo.ctypes.data_as(i64p))
        _lib.check(L.b2k_dec_unpack_lattices(*args))            # sizes
        ns, na, nf = r.num_states, r.num_arcs, r.num_finals
        out = dict(
            state_frame=np.zeros(ns, np.int32), state_hclg=np.zeros(ns, np.int32),
            state_tot_cost=np.zeros(ns, np.float32), state_extra_cost=np.zeros(ns, np.float32),
            arc_src=np.zeros(na, np.int32), arc_dst=np.zeros(na, np.int32),
            arc_ilabel=np.zeros(na, np.int32), arc_olabel=np.zeros(na, np.int32),
            arc_graph_cost=np.zeros(na, np.float32), arc_acoustic_cost=np.zeros(na, np.float32),
            final_state=np.zeros(nf, np.int32), final_cost=np.zeros(nf, np.float32))
        for k, v in out.items():
            setattr(r, k, _p(v, C.c_float if v.dtype == np.float32 else C.c_int32))
        _lib.check(L.b2k_dec_unpack_lattices(*args))
        out.update(state_offs=so, arc_offs=ao, final_offs=fo)
        return out

    def read(self, what: str, n: int) -> np.ndarray:
        """Stage outputs of batch slots 0..n-1 on the host: 'features', 'ivectors' or 'loglikes'."""
        import ctypes as C
        from . import _lib
        pl = self.plan
        idx, shape = {"features": (0, (n, pl.num_feature_frames, pl.feat_dim)),
                      "ivectors": (1, (n, pl.num_chunks, max(1, pl.ivector_dim))),
                      "loglikes": (2, (n, pl.num_output_frames, pl.num_pdfs))}[what]
        out = np.zeros(shape, np.float32)
        self._L.b2k_pipeline_read.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]
        _lib.check(self._L.b2k_pipeline_read(self.h, idx, n, out.ctypes.data, None))
        return out
