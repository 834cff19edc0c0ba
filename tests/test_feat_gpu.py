"""GPU parity tests of the fused MFCC/fbank kernel and the online CMVN kernel
(through the C-ABI) against the compiled reference (oracle/_ref), its golden
outputs and the numpy restatement.  Tolerance: float32 chain with different
FFT/dot association -> 2e-3 absolute on coefficients of magnitude ~100
(2e-5 of scale); CMVN mean normalisation is bit-exact."""
import os

import numpy as np
import pytest

from kaldi_b200 import synth

pytestmark = pytest.mark.gpu
ATOL = 2e-3
G = np.load(os.path.join(os.path.dirname(__file__), "golden", "feat_golden.npz"))


def _opts(o):
    from kaldi_b200.feat import FeatureOptions
    from dataclasses import asdict
    d = asdict(o)
    return FeatureOptions(**d)


@pytest.fixture(scope="module")
def cfgs():
    from oracle import feat_oracle as F
    return dict(
        mfcc_hires=F.FeatOpts(),
        fbank40=F.FeatOpts(feature_type=1),
        mfcc_hires_nosnip=F.FeatOpts(snip_edges=0),
        mfcc13_energy=F.FeatOpts(num_bins=23, num_ceps=13, use_energy=1, high_freq=0.0, energy_floor=1.0),
        htk1=F.FeatOpts(num_bins=23, num_ceps=13, preemph_coeff=0.0, window_type=1, remove_dc_offset=0,
                        low_freq=0.0, high_freq=0.0, htk_mode=1, htk_compat=1, use_energy=0),
        fbank_mag_energy=F.FeatOpts(feature_type=1, use_power=0, use_energy=1, raw_energy=0, htk_compat=1),
    )


@pytest.mark.parametrize("name", ["mfcc_hires", "fbank40", "mfcc_hires_nosnip", "mfcc13_energy", "htk1"])
def test_kernel_matches_reference_golden(cfgs, name):
    from kaldi_b200.feat import BatchedFeatures
    bf = BatchedFeatures(_opts(cfgs[name]))
    waves = [G["test_wav"].astype(np.float32), synth.make_audio(16000, seed=42)]
    got = bf.compute(waves)
    for g, key in zip(got, ("testwav", "synth42")):
        want = G[f"{key}_{name}"]
        assert g.shape == want.shape
        np.testing.assert_allclose(g, want, atol=ATOL, rtol=0)
    if name == "htk1":   # the reference's own golden vectors, its own tolerance (feature-mfcc-test.cc:163)
        assert np.abs(got[0][10:-10] - G["htk_fea_1_static"][10:-10]).max() < 1.0


def test_kernel_vs_compiled_reference_ragged_batch_and_chunked(cfgs):
    from kaldi_b200.feat import BatchedFeatures
    from oracle import feat_oracle as F
    R = F.RefFeat()
    lens = [160000, 400, 559, 16001, 48017, 2880 * 5]
    waves = [synth.make_audio(n, seed=i) for i, n in enumerate(lens)]
    for name in ("mfcc_hires", "fbank_mag_energy", "mfcc_hires_nosnip"):
        o = cfgs[name]
        bf = BatchedFeatures(_opts(o))
        got = bf.compute(waves)
        got_chunked = bf.compute(waves, chunk_frames=18)     # 0.18 s chunks as online2 feeds
        for w, g, gc in zip(waves, got, got_chunked):
            ref = R.compute(w, o)
            assert g.shape == ref.shape
            np.testing.assert_allclose(g, ref, atol=ATOL, rtol=0)
            assert np.array_equal(g, gc)                     # online == offline, bit for bit


def test_too_short_utterance_gives_zero_frames(cfgs):
    from kaldi_b200.feat import BatchedFeatures
    bf = BatchedFeatures(_opts(cfgs["mfcc_hires"]))
    assert bf.NumFrames(399) == 0 and bf.NumFrames(400) == 1 and bf.NumFrames(160000) == 998
    got = bf.compute([synth.make_audio(399, seed=1), synth.make_audio(1000, seed=2)])
    assert got[0].shape == (0, 40) and got[1].shape == (4, 40)


def test_dither_is_rejected():
    from kaldi_b200.feat import BatchedFeatures, FeatureOptions
    from kaldi_b200._lib import B2kError
    with pytest.raises(B2kError):
        BatchedFeatures(FeatureOptions(dither=1.0))


def test_online_cmvn_bit_exact(cfgs):
    import torch
    from kaldi_b200.feat import BatchedFeatures, OnlineCmvnOptions
    from oracle import feat_oracle as F
    R = F.RefFeat()
    bf = BatchedFeatures(_opts(cfgs["mfcc_hires"]))
    waves = [synth.make_audio(160000, seed=3), synth.make_audio(30000, seed=4)]
    feats = [R.compute(w, cfgs["mfcc_hires"]) for w in waves]
    gs = np.zeros((2, 41)); gs[0, :40] = feats[0].sum(0); gs[1, :40] = (feats[0].astype(np.float64) ** 2).sum(0); gs[0, 40] = len(feats[0])
    d_gs = torch.from_numpy(gs.reshape(-1)).cuda()
    for opts, exact in ((OnlineCmvnOptions(), True), (OnlineCmvnOptions(cmn_window=100, speaker_frames=100, global_frames=30), True),
                        (OnlineCmvnOptions(normalize_variance=True), False)):
        d_in = [torch.from_numpy(f).cuda() for f in feats]
        d_out = [torch.zeros_like(x) for x in d_in]
        st = [torch.zeros(2 * 41, dtype=torch.float64, device="cuda") for _ in feats]
        # two calls (chunked) must carry the state
        T = [f.shape[0] for f in feats]
        half = [t // 2 for t in T]
        bf.ApplyCmvnBatched(opts, [x.data_ptr() for x in d_in], [x.data_ptr() for x in d_out], 40, 40,
                            [0, 0], half, [s.data_ptr() for s in st], d_gs.data_ptr())
        bf.ApplyCmvnBatched(opts, [x.data_ptr() for x in d_in], [x.data_ptr() for x in d_out], 40, 40,
                            half, [t - h for t, h in zip(T, half)], [s.data_ptr() for s in st], d_gs.data_ptr())
        torch.cuda.synchronize()
        for f, o in zip(feats, d_out):
            ref = R.online_cmvn(f, cmn_window=opts.cmn_window, speaker_frames=opts.speaker_frames,
                                global_frames=opts.global_frames, normalize_variance=opts.normalize_variance,
                                global_stats=gs)
            if exact:
                assert np.array_equal(o.cpu().numpy(), ref)
            else:
                np.testing.assert_allclose(o.cpu().numpy(), ref, rtol=2e-6, atol=2e-6)


def test_full_size_batch_properties(cfgs):
    """BASELINE-size: 64 x 10 s utterances; shift invariance (frame f of a wave
    delayed by 160 samples == frame f+1) and batch/lane independence."""
    from kaldi_b200.feat import BatchedFeatures
    bf = BatchedFeatures(_opts(cfgs["mfcc_hires"]))
    base = [synth.make_audio(160000, seed=100 + i) for i in range(8)]
    waves = base * 8
    got = bf.compute(waves)
    for i in range(8):
        for r in range(1, 8):
            assert np.array_equal(got[i], got[i + 8 * r])
    shifted = bf.compute([base[0][160:]])[0]
    assert np.array_equal(shifted, got[0][1:])
    assert all(g.shape == (998, 40) and np.isfinite(g).all() for g in got)


@pytest.mark.parametrize("name", ["mfcc_hires", "mfcc_hires_nosnip", "fbank_mag_energy"])
def test_int16_input_and_unaligned_sources_give_the_same_bits(cfgs, name):
    """The kernel stages the samples of a CTA's frames with one bulk copy cut to 16-byte boundaries of the source and reads the
    ragged ends directly: (a) 16-bit PCM in (b2k_feat_compute_batched_i16) gives exactly the features of its float form,
    (b) a waveform that starts 1, 2 or 3 samples past a 16-byte boundary gives exactly the features of the aligned copy,
    (c) frames requested in pieces that start in the middle of a CTA's span do too; short and ragged lengths included."""
    import torch
    from kaldi_b200.feat import BatchedFeatures
    bf = BatchedFeatures(_opts(cfgs[name]))
    D = bf.Dim()
    for n_samples, seed in ((48017, 1), (16000, 2), (559, 3), (400, 4), (10480 + 160 * 64, 5)):
        pcm = np.clip(np.round(synth.make_audio(n_samples, seed=seed)), -32768, 32767).astype(np.int16)
        want = bf.compute([pcm.astype(np.float32)])[0]
        T = want.shape[0]
        if T == 0:
            continue
        for shift in (0, 1, 2, 3, 5):
            for i16 in (False, True):
                buf = torch.zeros(n_samples + 16, dtype=torch.int16 if i16 else torch.float32, device="cuda")
                src = torch.from_numpy(pcm if i16 else pcm.astype(np.float32)).cuda()
                buf[shift:shift + n_samples] = src
                out = torch.zeros(T, D, device="cuda")
                ptr = buf.data_ptr() + shift * (2 if i16 else 4)
                pieces = [(0, T)] if shift != 5 else [(0, min(7, T)), (min(7, T), T - min(7, T))]
                for f0, k in pieces:
                    if k > 0:
                        bf.ComputeFeaturesBatched([ptr], [n_samples], [f0], [k], [out.data_ptr()], D, int16=i16)
                torch.cuda.synchronize()
                np.testing.assert_array_equal(out.cpu().numpy(), want, err_msg=f"{name} n={n_samples} shift={shift} int16={i16}")


@pytest.mark.parametrize("kw", [dict(use_energy=1), dict(use_energy=0), dict(use_energy=0, htk_compat=1, cepstral_scale=10.0),
                                dict(use_energy=1, raw_energy=0, num_ceps=9, lpc_order=10, cepstral_lifter=0.0)])
def test_plp_kernel_vs_compiled_reference(kw):
    """feature_type 2: PlpComputer::Compute (feat/feature-plp.cc:112-182, compiled in oracle/_ref) -- the kernel's tail after the
    mel bank (equal loudness, cube-root compression, inverse-DFT autocorrelation, Durbin, LPC -> cepstrum, lifter, scale, C0 or
    energy, HTK order).  Durbin's recursion amplifies the ~1e-6 differences between the two FFTs about 40-fold (measured on the
    reference itself), so the tolerance is 2e-3 per unit of cepstral scale on cepstra of magnitude <= 6 (C0 ~ 20); int16 input
    gives the same bits as float input."""
    import torch
    from kaldi_b200.feat import BatchedFeatures
    from oracle import feat_oracle as F
    o = F.FeatOpts(**{**dict(feature_type=2, num_bins=23, num_ceps=13, low_freq=20.0, high_freq=0.0), **kw})
    bf = BatchedFeatures(_opts(o))
    assert bf.Dim() == o.num_ceps
    R = F.RefFeat()
    waves = [np.clip(np.round(synth.make_audio(n, seed=60 + i)), -32768, 32767).astype(np.float32) for i, n in enumerate((48000, 16000, 559))]
    got = bf.compute(waves)
    for w, g in zip(waves, got):
        want = R.compute(w, o)
        assert g.shape == want.shape
        assert np.abs(g - want).max() <= 2e-3 * o.cepstral_scale, (np.abs(g - want).max(), np.abs(want).max())
    pcm = waves[0].astype(np.int16)
    d = torch.from_numpy(pcm).cuda()
    T = got[0].shape[0]
    out = torch.zeros(T, bf.Dim(), device="cuda")
    bf.ComputeFeaturesBatched([d.data_ptr()], [len(pcm)], [0], [T], [out.data_ptr()], bf.Dim(), int16=True)
    torch.cuda.synchronize()
    np.testing.assert_array_equal(out.cpu().numpy(), got[0])
