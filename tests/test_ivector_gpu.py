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
