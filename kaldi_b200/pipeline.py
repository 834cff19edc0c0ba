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
