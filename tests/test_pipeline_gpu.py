"""End-to-end GPU test: waveform -> features -> nnet3 -> decoder through the
public pipeline API, compared stage by stage with the CPU reference path
(compiled reference features + compiled reference nnet3 + decoder oracle)."""
import numpy as np
import pytest

from kaldi_b200 import nnet_model as NM, synth

pytestmark = pytest.mark.gpu


def test_end_to_end_small_model_matches_cpu_reference_path():
    import torch
    from kaldi_b200.decoder import CudaDecoder, lattice_to_canonical
    from kaldi_b200.pipeline import BatchedPipeline, PipelineConfig
    from oracle import dec_oracle as D, feat_oracle as F, nnet_oracle as NO
    P = 300
    arch = NM.arch_tiny(P)
    W = NM.random_weights(arch, seed=2)
    g = synth.make_hclg(200_000, num_pdfs=P, seed=4)
    S = 32000
    cfg = PipelineConfig(max_batch=3, num_samples=S, extract_ivectors=False)
    pipe = BatchedPipeline(cfg, arch, W, g)
    waves = [synth.make_audio(S, seed=i) for i in range(3)]
    rng = np.random.default_rng(0)
    civ = rng.standard_normal((3, pipe.nnet.n_chunks, 100)).astype(np.float32)
    pipe.d_ivec[:3].copy_(torch.from_numpy(civ).cuda())
    packed = pipe.decode_batch(waves)
    lats = CudaDecoder.SplitLattices(packed)
    R = F.RefFeat()
    RN = NO.RefNnet(arch, W)
    for i in range(3):
        feats = R.compute(waves[i], F.FeatOpts(), online_chunk=2880)
        got_feats = pipe.d_feats[i].cpu().numpy()
        np.testing.assert_allclose(got_feats, feats, atol=2e-3, rtol=0)
        # log-likes: GPU nnet on GPU features vs reference nnet on reference features
        ends = [(n + 1) * RN.frames_per_chunk + RN.right_context for n in range(pipe.nnet.n_chunks)]
        mat = np.zeros((ends[-1] + 1, 100), np.float32)
        prev = 0
        for n, e in enumerate(ends):
            mat[prev:e + 1] = civ[i, n]; prev = e + 1
        ll_ref = RN.forward(feats, mat, period=1)
        ll_gpu = pipe.d_loglikes[i].cpu().numpy()
        assert np.abs(ll_gpu - ll_ref).max() <= 1e-4 * np.abs(ll_ref).max()
        # decoder: bit-exact given identical log-likes (the GPU's own)
        o = D.DecoderOracle(g, cfg.decoder_cfg)
        o.decode(ll_gpu, mode=D.MODE_REFERENCE_ORDER)
        want = o.lattice()
        got = lattice_to_canonical(lats[i])
        assert all(np.array_equal(got[k], want[k]) for k in got)


def test_end_to_end_with_online_ivectors():
    """Full chain including the GPU i-vector stage, against the CPU path built
    from the compiled reference pieces (features, i-vector, nnet3) + decoder oracle."""
    from kaldi_b200 import ivector as IVM
    from kaldi_b200.decoder import CudaDecoder, lattice_to_canonical
    from kaldi_b200.pipeline import BatchedPipeline, PipelineConfig
    from oracle import dec_oracle as D, feat_oracle as F, ivector_oracle as IO, nnet_oracle as NO
    P = 300
    arch = NM.arch_tiny(P)
    W = NM.random_weights(arch, seed=2)
    g = synth.make_hclg(200_000, num_pdfs=P, seed=4)
    S = 48000
    ex = IVM.make_synthetic_extractor(3, num_gauss=64, ivector_dim=100)
    cfg = PipelineConfig(max_batch=2, num_samples=S)
    pipe = BatchedPipeline(cfg, arch, W, g, ivector_extractor=ex)
    waves = [synth.make_audio(S, seed=10 + i) for i in range(2)]
    lats = CudaDecoder.SplitLattices(pipe.decode_batch(waves))
    R, RN, RI = F.RefFeat(), NO.RefNnet(arch, W), IO.RefIvector(ex)
    for i in range(2):
        feats = R.compute(waves[i], F.FeatOpts(), online_chunk=2880)
        sched = IVM.online_ivector_schedule(S, 2880, 400, 160, feats.shape[0], RN.right_context, RN.frames_per_chunk, 3)
        civ = RI.run(feats, sched)
        got_iv = pipe.d_ivec[i].cpu().numpy()
        assert np.abs(got_iv - civ).max() <= 2e-4 * np.linalg.norm(civ, axis=1).max()
        ends = [(n + 1) * RN.frames_per_chunk + RN.right_context for n in range(len(sched))]
        mat = np.zeros((ends[-1] + 1, 100), np.float32)
        prev = 0
        for n, e in enumerate(ends):
            mat[prev:e + 1] = civ[n]; prev = e + 1
        ll_ref = RN.forward(feats, mat, period=1)
        ll_gpu = pipe.d_loglikes[i].cpu().numpy()
        assert np.abs(ll_gpu - ll_ref).max() <= 1e-4 * np.abs(ll_ref).max()
        o = D.DecoderOracle(g, cfg.decoder_cfg)
        o.decode(ll_gpu, mode=D.MODE_REFERENCE_ORDER)
        got, want = lattice_to_canonical(lats[i]), o.lattice()
        assert all(np.array_equal(got[k], want[k]) for k in got)
