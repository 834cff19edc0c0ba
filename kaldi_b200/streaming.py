"""Chunk-by-chunk decoding of several audio streams at once: the call structure of
cuda_decoder::BatchedThreadedNnet3CudaOnlinePipeline::DecodeBatch(corr_ids, wave_samples, is_first_chunk, is_last_chunk)
(cudadecoder/batched-threaded-nnet3-cuda-online-pipeline.cc:316-377) over the b2k stages:

  samples of this call  ->  b2k_feat_compute_batched on the frames that became computable (online == offline, bit-exact)
                        ->  b2k_nnet_stream_run_batch (BatchedStaticNnet3::RunBatch: context restored / saved per channel,
                            right context flushed on the last chunk)
                        ->  b2k_dec_advance_decoding_frames on the output frames of this call
                        ->  b2k_dec_best_path(use_final_probs = 0): the partial hypothesis
  last chunk            ->  b2k_dec_finalize_decoding + the raw lattice.

Nothing is recomputed from the start of the utterance; per channel the device keeps the waveform so far (the feature kernel
reads frames that straddle two chunks from it), the context frames of the network and the decoder's arenas.  Host side only;
PyTorch holds the buffers."""
from __future__ import annotations

import numpy as np

from .decoder import CudaDecoder, CudaDecoderConfig, CudaFst
from .feat import BatchedFeatures, FeatureOptions
from .nnet import BatchedStaticNnet3


class StreamingBatchedDecoder:
    def __init__(self, arch: dict, weights: dict, graph: dict, decoder_cfg: dict, nchannels: int, max_seconds: float = 30.0,
                 frames_per_chunk: int = 51, feature_opts: FeatureOptions | None = None, acoustic_scale: float = 1.0):
        import torch
        self.torch = torch
        self.arch, self.nchannels, self.fpc = arch, nchannels, frames_per_chunk
        fo = feature_opts or FeatureOptions(max_lanes=max(nchannels, 8))
        self.feat = BatchedFeatures(fo)
        self.D, self.P, self.ivd = self.feat.Dim(), arch["num_pdfs"], arch["ivector_dim"]
        assert self.D == arch["feat_dim"]
        self.max_samples = int(max_seconds * fo.samp_freq)
        self.max_frames = self.feat.NumFrames(self.max_samples, True) + 1
        self.nnet = BatchedStaticNnet3(arch, weights, max_batch=nchannels, nchannels=nchannels, frames_per_chunk=frames_per_chunk,
                                       acoustic_scale=acoustic_scale)
        self.opc = self.nnet.output_frames_per_chunk
        sub = arch["frame_subsampling_factor"]
        c = CudaDecoderConfig.from_dict(decoder_cfg, max_frames=(self.max_frames + sub - 1) // sub + 8)
        self.fst = CudaFst(graph)
        self.dec = CudaDecoder(self.fst, c, nchannels, nchannels)
        dev = "cuda"
        self.d_wave = torch.zeros(nchannels, self.max_samples, device=dev)
        self.d_feats = torch.zeros(nchannels, self.max_frames, self.D, device=dev)
        self.d_zero_iv = torch.zeros(max(1, self.ivd), device=dev)
        self.d_out = torch.zeros(nchannels * self.opc, self.P, device=dev)
        self.d_eos = torch.zeros(nchannels * self.opc, self.P, device=dev)
        self.samples = [0] * nchannels
        self.frames = [0] * nchannels
        self.out_frames = [0] * nchannels
        self.started = [False] * nchannels

    def DecodeBatch(self, channels, wave_chunks, is_first_chunk, is_last_chunk, ivector_ptrs=None, want_partial: bool = True,
                    keep_loglikes: list | None = None):
        """One chunk of samples (float, int16 range) per listed channel.  Returns per channel a dictionary: the output frames
        this call decoded, the partial best path (olabels / words so far, use_final_probs = False) and, for a last chunk, the
        finalized raw lattice.  keep_loglikes: a list that receives (channel, [frames x pdfs]) of what the decoder consumed."""
        torch = self.torch
        n = len(channels)
        assert len(set(channels)) == n
        first_frame, n_new = [], []
        for ch, w, first, last in zip(channels, wave_chunks, is_first_chunk, is_last_chunk):
            if first:
                self.samples[ch] = self.frames[ch] = self.out_frames[ch] = 0
                self.started[ch] = False
            w = np.ascontiguousarray(w, np.float32)
            assert self.samples[ch] + len(w) <= self.max_samples, "stream longer than max_seconds"
            self.d_wave[ch, self.samples[ch]:self.samples[ch] + len(w)] = torch.from_numpy(w).cuda()
            self.samples[ch] += len(w)
            ready = self.feat.NumFrames(self.samples[ch], bool(last)) if self.samples[ch] > 0 else 0
            first_frame.append(self.frames[ch])
            n_new.append(max(0, ready - self.frames[ch]))
            assert n_new[-1] <= self.fpc, "feed at most frames_per_chunk frames of audio per call"
        wp = [self.d_wave[ch].data_ptr() for ch in channels]
        fp = [self.d_feats[ch].data_ptr() for ch in channels]
        live = [i for i in range(n) if n_new[i] > 0]
        if live:
            self.feat.ComputeFeaturesBatched([wp[i] for i in live], [self.samples[channels[i]] for i in live],
                                             [first_frame[i] for i in live], [n_new[i] for i in live], [fp[i] for i in live], self.D)
        new_ptr = [self.d_feats[ch, f0].data_ptr() if k > 0 else self.d_feats[ch].data_ptr() for ch, f0, k in zip(channels, first_frame, n_new)]
        iv = None
        if self.ivd > 0:
            iv = [int(p) for p in ivector_ptrs] if ivector_ptrs is not None else [self.d_zero_iv.data_ptr()] * n
        no, ne = self.nnet.RunBatch(channels, new_ptr, self.D, iv, n_new, is_first_chunk, is_last_chunk, self.d_out.data_ptr(),
                                    self.d_eos.data_ptr(), self.P)
        for ch, k in zip(channels, n_new):
            self.frames[ch] += k
        fresh = [ch for ch in channels if not self.started[ch]]
        if fresh:
            self.dec.InitDecoding(fresh)
            for ch in fresh:
                self.started[ch] = True
        stride = self.d_out.stride(0)
        act = [(i, ch) for i, ch in enumerate(channels) if no[i] > 0]
        if act:
            self.dec.AdvanceDecodingFrames([ch for _, ch in act], [self.d_out[i * self.opc].data_ptr() for i, _ in act],
                                           [no[i] for i, _ in act], stride)
        act2 = [(i, ch) for i, ch in enumerate(channels) if ne[i] > 0]
        if act2:
            self.dec.AdvanceDecodingFrames([ch for _, ch in act2], [self.d_eos[i * self.opc].data_ptr() for i, _ in act2],
                                           [ne[i] for i, _ in act2], stride)
        if keep_loglikes is not None:
            torch.cuda.synchronize()
            o, e = self.d_out.cpu().numpy(), self.d_eos.cpu().numpy()
            for i, ch in enumerate(channels):
                keep_loglikes.append((ch, np.concatenate([o[i * self.opc:i * self.opc + no[i]], e[i * self.opc:i * self.opc + ne[i]]], 0)))
        res = []
        partial = self.dec.GetBestPath(channels, use_final_probs=False) if want_partial else [None] * n
        for i, ch in enumerate(channels):
            self.out_frames[ch] += no[i] + ne[i]
            r = dict(channel=ch, new_output_frames=no[i] + ne[i], frames_decoded=self.out_frames[ch])
            if want_partial:
                r["partial_words"] = partial[i]["olabels"][partial[i]["olabels"] != 0]
                r["partial_cost"] = partial[i]["best_cost"]
            res.append(r)
        done = [ch for ch, last in zip(channels, is_last_chunk) if last]
        if done:
            self.dec.FinalizeDecoding(done)
            for r, last in zip(res, is_last_chunk):
                if last:
                    r["lattice"] = self.dec.GetRawLattice(r["channel"])
        return res


class NativeStreamingDecoder:
    """The same through the C ABI alone (b2k_stream_*, kaldi_b200/csrc/stream_pipeline.cu): model.KaldiModel + CudaFst ->
    b2k_stream_create; DecodeBatch(channels, int16 chunks, is_first_chunk, is_last_chunk) -> b2k_stream_decode_batch_i16;
    partial hypotheses and lattices come from the stream's decoder handle (b2k_dec_best_path, b2k_dec_get_raw_lattice)."""

    def __init__(self, model, fst: CudaFst, decoder_cfg: dict, nchannels: int, max_seconds: float = 30.0, frames_per_chunk: int = 51,
                 feature_opts: FeatureOptions | None = None, acoustic_scale: float = 1.0, use_priors: bool = True,
                 decoder_kwargs: dict | None = None):
        import ctypes as C
        from dataclasses import fields
        from . import _lib
        from .decoder import _DecCfg
        from .feat import _FeatCfg
        self._C, self._lib_mod = C, _lib
        L = self._L = _lib.lib()
        fo = feature_opts or FeatureOptions(max_lanes=max(nchannels, 8))

        class _StreamCfg(C.Structure):
            _fields_ = [("feat", _FeatCfg), ("dec", _DecCfg), ("nchannels", C.c_int32), ("max_seconds", C.c_float),
                        ("frames_per_chunk", C.c_int32), ("acoustic_scale", C.c_float), ("use_priors", C.c_int32)]
        c = _StreamCfg()
        L.b2k_stream_cfg_default.argtypes = [C.c_void_p]
        L.b2k_stream_cfg_default.restype = None
        L.b2k_stream_cfg_default(C.byref(c))
        c.feat = _FeatCfg(**{f.name: getattr(fo, f.name) for f in fields(fo)})
        dc = CudaDecoderConfig.from_dict(decoder_cfg, **(decoder_kwargs or {}))
        c.dec = _DecCfg(dc.default_beam, dc.lattice_beam, dc.max_active, dc.min_active, dc.beam_delta, dc.prune_interval,
                        dc.prune_scale, dc.max_tokens_per_frame, dc.max_frames, dc.max_tokens, dc.max_links,
                        int(dc.reference_order), dc.hash_ratio, dc.max_arcs_per_frame, dc.max_lattice_states, dc.max_lattice_arcs)
        c.nchannels, c.max_seconds, c.frames_per_chunk = int(nchannels), float(max_seconds), int(frames_per_chunk)
        c.acoustic_scale, c.use_priors = float(acoustic_scale), int(use_priors)
        self.model, self.fst = model, fst
        self.h = C.c_void_p()
        L.b2k_stream_create.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        _lib.check(L.b2k_stream_create(C.byref(c), model.h, fst.h, C.byref(self.h)))
        info = (C.c_int64 * 8)()
        L.b2k_stream_info.argtypes = [C.c_void_p, C.c_void_p]
        _lib.check(L.b2k_stream_info(self.h, info))
        (self.nchannels, self.max_samples, self.max_frames, self.D, self.P, self.ivd, self.opc, self.fpc) = [int(x) for x in info]
        # a CudaDecoder view of the stream's own decoder (not owned: the stream destroys it)
        L.b2k_stream_decoder.restype = C.c_void_p
        L.b2k_stream_decoder.argtypes = [C.c_void_p]
        self.dec = CudaDecoder.__new__(CudaDecoder)
        self.dec.h = None
        self._dec_h = C.c_void_p(L.b2k_stream_decoder(self.h))
        self.dec.fst, self.dec.config, self.dec.nlanes, self.dec.nchannels = fst, dc, nchannels, nchannels
        L.b2k_stream_decode_batch_i16.argtypes = [C.c_void_p, C.c_int32] + [C.c_void_p] * 10 + [C.c_void_p]

    def _with_dec(self):
        self.dec.h = self._dec_h
        return self.dec

    def __del__(self):
        try:
            if getattr(self, "dec", None) is not None:
                self.dec.h = None                      # the view must not destroy the stream's decoder
            if self.h:
                self._L.b2k_stream_destroy.argtypes = [self._C.c_void_p]
                self._L.b2k_stream_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def DecodeBatch(self, channels, wave_chunks_i16, is_first_chunk, is_last_chunk, want_partial: bool = True,
                    lattices: str = "each"):
        """lattices: "each" = the raw lattice of every stream that ended in its result; "batched" = one packed read-back of all of
        them (CudaDecoder.GetRawLattices) in the first result under "lattices_packed"; "none"."""
        C = self._C
        n = len(channels)
        keep = [np.ascontiguousarray(w, np.int16) for w in wave_chunks_i16]
        ch = (C.c_int32 * n)(*[int(c) for c in channels])
        hp = (C.c_void_p * n)(*[w.ctypes.data if len(w) else None for w in keep])
        ns = (C.c_int32 * n)(*[len(w) for w in keep])
        fi = (C.c_int32 * n)(*[int(bool(x)) for x in is_first_chunk])
        la = (C.c_int32 * n)(*[int(bool(x)) for x in is_last_chunk])
        new, sofar = (C.c_int32 * n)(), (C.c_int32 * n)()
        pn, pf = (C.c_void_p * n)(), (C.c_void_p * n)()
        self._lib_mod.check(self._L.b2k_stream_decode_batch_i16(self.h, n, ch, hp, ns, fi, la, None, new, sofar, pn, pf, None))
        dec = self._with_dec()
        try:
            res = []
            partial = dec.GetBestPath(list(channels), use_final_probs=False) if want_partial else [None] * n
            for i in range(n):
                r = dict(channel=int(channels[i]), new_output_frames=int(new[i]), frames_decoded=int(sofar[i]))
                if want_partial:
                    r["partial_words"] = partial[i]["olabels"][partial[i]["olabels"] != 0]
                    r["partial_cost"] = partial[i]["best_cost"]
                if is_last_chunk[i] and lattices == "each":
                    r["lattice"] = dec.GetRawLattice(int(channels[i]))
                res.append(r)
            done = [int(c) for c, l in zip(channels, is_last_chunk) if l]
            if done and lattices == "batched":
                res[0]["lattices_packed"] = dec.GetRawLattices(done)
                res[0]["lattices_channels"] = done
            return res
        finally:
            self.dec.h = None
