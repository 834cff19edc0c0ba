"""GPU parity tests of the online i-vector stage against the reference's own
i-vector code compiled in oracle/_ref (see oracle/ref_wrap/ivector_wrap.cc).
Tolerance: the i-vector is the result of <=15 CG iterations in double on
float-derived statistics; posterior pruning is discrete, so the comparison is
2e-4 of the i-vector norm (observed ~1e-6) -- far inside the effect it has on
log-likelihoods (north star 1e-4)."""
import numpy as np
import pytest

from kaldi_b200 import ivector as IVM, synth

pytestmark = pytest.mark.gpu


def _feats(seed, n=160000):
    from oracle import feat_oracle as F
    return F.mfcc_fbank(synth.make_audio(n, seed=seed), F.FeatOpts())


@pytest.mark.parametrize("G,D,max_count", [(64, 20, 0.0), (64, 20, 100.0), (512, 100, 100.0)])
def test_chunk_ivectors_match_reference(G, D, max_count):
    import torch
    from oracle import ivector_oracle as IO
    ex = IVM.make_synthetic_extractor(1, num_gauss=G, ivector_dim=D, max_count=max_count)
    R = IO.RefIvector(ex)
    feats = [_feats(3), _feats(4)]
    T = feats[0].shape[0]
    sched = IVM.online_ivector_schedule(160000, 2880, 400, 160, T, 29, 21, 3)
    gpu = IVM.IvectorExtractorGpu(ex, max_lanes=2, max_frames=T)
    d_f = [torch.from_numpy(f).cuda() for f in feats]
    d_o = [torch.zeros(len(sched), D, device="cuda") for _ in feats]
    gpu.Compute([x.data_ptr() for x in d_f], 40, T, sched, [x.data_ptr() for x in d_o], D)
    torch.cuda.synchronize()
    for f, o in zip(feats, d_o):
        ref = R.run(f, sched)
        got = o.cpu().numpy()
        scale = np.linalg.norm(ref, axis=1).max()
        assert np.abs(got - ref).max() <= 2e-4 * scale, (np.abs(got - ref).max(), scale)


def test_schedule_matches_online_tool_bookkeeping():
    # 10 s utterance, 0.18 s chunks, nnet right context 29, chunk 21: chunk 0 is computed after the
    # 3rd audio chunk (52 frames ready >= 21 + 29) and asks for frame min(51, 52-3-1) = 48
    s = IVM.online_ivector_schedule(160000, 2880, 400, 160, 998, 29, 21, 3)
    assert len(s) == 48 and s[0] == 48 and s[-1] == 997 and np.all(np.diff(s) >= 0)
    # short utterance: everything is flushed at InputFinished
    s2 = IVM.online_ivector_schedule(8000, 2880, 400, 160, 48, 29, 21, 3)
    assert list(s2) == [47] * len(s2)


def test_chunks_before_the_first_ivector_frame_get_zeros():
    """sched[n] = -1 (b2k_ivec_online_schedule for a chunk that runs before any i-vector frame is ready): all zeros for that
    chunk, and the later chunks are what they are without it (decodable-online-looped.cc:188-197)."""
    import torch
    from oracle import ivector_oracle as IO
    G, D = 64, 20
    ex = IVM.make_synthetic_extractor(1, num_gauss=G, ivector_dim=D, max_count=100.0)
    R = IO.RefIvector(ex)
    f = _feats(5, 32000)
    T = f.shape[0]
    sched = np.array([-1, -1, 10, 50, T - 1], np.int32)
    gpu = IVM.IvectorExtractorGpu(ex, max_lanes=1, max_frames=T)
    d_f = torch.from_numpy(f).cuda()
    d_o = torch.full((len(sched), D), 7.0, device="cuda")
    gpu.Compute([d_f.data_ptr()], 40, T, sched, [d_o.data_ptr()], D)
    torch.cuda.synchronize()
    got, ref = d_o.cpu().numpy(), R.run(f, sched)
    assert not got[:2].any() and not ref[:2].any()
    assert np.abs(got - ref).max() <= 2e-4 * np.linalg.norm(ref, axis=1).max()


@pytest.mark.parametrize("max_remembered", [1000.0, 150.0])
def test_speaker_adaptation_state_carries_over_like_the_reference(max_remembered):
    """Three utterances of one speaker and one of another, run as the tool runs a speaker's utterances (SetAdaptationState
    before, GetAdaptationState after, online2-wav-nnet3-latgen-faster.cc:199-221,287): every chunk's i-vector and the state
    after every utterance against the reference's own OnlineCmvn::GetState / OnlineIvectorEstimationStats (oracle/_ref;
    LimitFrames restated over their Scale()), and the carried state must matter (the second utterance's first i-vector is not
    what a new speaker gets).  max_remembered 150 makes LimitFrames fire on both halves."""
    import torch
    from oracle import ivector_oracle as IO
    G, D = 64, 20
    ex = IVM.make_synthetic_extractor(2, num_gauss=G, ivector_dim=D, max_count=100.0)
    R = IO.RefIvector(ex)
    utts = [_feats(11, 48000), _feats(12, 48000), _feats(13, 48000)]
    other = _feats(14, 48000)
    T = utts[0].shape[0]
    sched = IVM.online_ivector_schedule(48000, 2880, 400, 160, T, 29, 21, 3)
    gpu = IVM.IvectorExtractorGpu(ex, max_lanes=2, max_frames=T)
    S = gpu.AdaptationStateDoubles()
    assert S == 2 * 41 + 1 + D + D * (D + 1) // 2
    d_state = torch.zeros(2, S, dtype=torch.float64, device="cuda")
    spk_a, spk_b = R.new_speaker(), R.new_speaker()
    fresh = None
    for k, f in enumerate(utts):
        lanes = [f, other] if k == 0 else [f]
        d_f = [torch.from_numpy(x).cuda() for x in lanes]
        d_o = [torch.zeros(len(sched), D, device="cuda") for _ in lanes]
        sin = [0 if k == 0 else d_state[0].data_ptr()] + ([0] if k == 0 else [])
        sout = [d_state[0].data_ptr()] + ([d_state[1].data_ptr()] if k == 0 else [])
        gpu.ComputeAdapt([x.data_ptr() for x in d_f], 40, T, sched, [x.data_ptr() for x in d_o], D, sin, sout, max_remembered)
        torch.cuda.synchronize()
        ref = R.run(f, sched, speaker=spk_a, max_remembered_frames=max_remembered)
        got = d_o[0].cpu().numpy()
        scale = np.linalg.norm(ref, axis=1).max()
        assert np.abs(got - ref).max() <= 2e-4 * scale, (k, np.abs(got - ref).max(), scale)
        st_ref, st_got = R.speaker_state(spk_a), d_state[0].cpu().numpy()
        np.testing.assert_allclose(st_got, st_ref, rtol=1e-5, atol=1e-5 * np.abs(st_ref).max(), err_msg=f"state after utterance {k}")
        if k == 0:
            ref_b = R.run(other, sched, speaker=spk_b, max_remembered_frames=max_remembered)
            assert np.abs(d_o[1].cpu().numpy() - ref_b).max() <= 2e-4 * np.linalg.norm(ref_b, axis=1).max()
            np.testing.assert_allclose(d_state[1].cpu().numpy(), R.speaker_state(spk_b), rtol=1e-5,
                                       atol=1e-5 * np.abs(R.speaker_state(spk_b)).max())
        if k == 1:
            fresh = R.run(f, sched)                   # the same utterance as a new speaker
            assert np.abs(fresh[0] - ref[0]).max() > 1e-3 * scale, "the carried state had no effect"
