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


@pytest.mark.gpu
def test_native_pipeline_carries_speaker_adaptation_across_waves():
    """b2k_pipeline_set_speaker_states + ingest.speaker_waves: five utterances of three speakers decoded in waves through the
    C++ pipeline; every utterance's chunk i-vectors equal the reference-side run that keeps one OnlineCmvnState /
    OnlineIvectorEstimationStats per speaker in the tool's order (online2-wav-nnet3-latgen-faster.cc:199-221,287), and they
    differ from the run without adaptation."""
    import torch
    from kaldi_b200 import ivector as IVM, nnet_model as NM, synth
    from kaldi_b200.decoder import CudaFst
    from kaldi_b200.ingest import speaker_waves
    from kaldi_b200.model import KaldiModel
    from kaldi_b200.pipeline import NativeBatchedPipeline, PipelineConfig
    from oracle import ivector_oracle as IO
    S, B, P = 32000, 3, 64
    arch = NM.arch_tiny(P)
    W = NM.random_weights(arch, seed=2)
    g = synth.make_hclg(50_000, num_pdfs=P, seed=4)
    cfg = PipelineConfig(max_batch=B, num_samples=S)
    T = 1 + (S - 400) // 160
    ex = IVM.make_synthetic_extractor(3, num_gauss=64, ivector_dim=100)
    ivx = IVM.IvectorExtractorGpu(ex, B, T)
    nat = NativeBatchedPipeline(cfg, KaldiModel.from_arch(arch, W), CudaFst(g), ivx)
    speakers = ["a", "b", "a", "c", "a"]
    waves_audio = [synth.make_audio(S, seed=40 + i) for i in range(len(speakers))]
    names = sorted(set(speakers))
    d_state = torch.zeros(len(names), ivx.AdaptationStateDoubles(), dtype=torch.float64, device="cuda")
    has_state = {s: False for s in names}
    got, feats = {}, {}
    for wave in speaker_waves(speakers):
        ptr = [d_state[names.index(speakers[u])].data_ptr() for u in wave]
        nat.set_speaker_states([p if has_state[speakers[u]] else 0 for p, u in zip(ptr, wave)], ptr, 1000.0)
        nat.decode_batch([waves_audio[u] for u in wave])
        iv, fe = nat.read("ivectors", len(wave)), nat.read("features", len(wave))
        for i, u in enumerate(wave):
            got[u], feats[u] = iv[i].copy(), fe[i].copy()
            has_state[speakers[u]] = True
    R = IO.RefIvector(ex)
    ref_spk = {s: R.new_speaker() for s in names}
    sched = IVM.online_ivector_schedule(S, 2880, 400, 160, T, NM.model_context(arch)[1], cfg.frames_per_chunk, 3)
    differs = 0
    for u, s in enumerate(speakers):            # the tool's order: a speaker's utterances one after the other
        want = R.run(feats[u], sched, speaker=ref_spk[s], max_remembered_frames=1000.0)
        scale = np.linalg.norm(want, axis=1).max()
        assert got[u].shape == want.shape and np.abs(got[u] - want).max() <= 2e-4 * scale, (u, np.abs(got[u] - want).max(), scale)
        differs += np.abs(R.run(feats[u], sched) - want).max() > 1e-3 * scale
    assert differs == 2                          # the second and third utterance of speaker a
    nat.decode_batch([waves_audio[0]])           # the states are one-shot: this batch runs as new speakers again
    assert np.abs(nat.read("ivectors", 1)[0] - got[0]).max() <= 1e-6 * np.abs(got[0]).max() + 1e-7
