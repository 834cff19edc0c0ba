"""Pins the i-vector oracle.  oracle/ref_wrap/ivector_wrap.cc runs the reference's own arithmetic (extractor, UBM, CMVN, splice, LDA,
posteriors, CG) but RESTATES the glue of OnlineIvectorFeature (which frames go into the statistics when, the adaptation state);
every GPU parity test of the i-vector stage compares against that.  Here the same inputs go through the reference's OWN
OnlineIvectorFeature (online2/online-ivector-feature.cc compiled unmodified over a replay decoder: oracle/ref_wrap/silence_wrap.cc)
and the two must agree bit for bit: i-vectors per chunk, and the speaker's state carried over three utterances."""
import os

import numpy as np
import pytest

from kaldi_b200 import ivector as IVM


@pytest.fixture(scope="module")
def ref():
    from oracle import ivector_oracle as IO
    if os.path.isdir("/root/reference"):
        from oracle import ref_nnet
        ref_nnet.build(quiet=True)
    if not os.path.exists(IO._SO):
        pytest.skip("oracle/_ref/libkaldi_ref_nnet3.so not built")
    r = IO.make_cpu_extractor(3)
    if not hasattr(r.lib, "ref_ivector_run_real"):
        pytest.skip("oracle/_ref predates silence_wrap.cc")
    return r


def _feats(rng, T, D):
    return (rng.standard_normal((T, D)) * 3.0 + rng.standard_normal(D)).astype(np.float32)


@pytest.mark.parametrize("online_cmvn", [False, True])
def test_restated_glue_equals_the_references_class(ref, online_cmvn):
    rng = np.random.default_rng(21)
    D = ref.ex["base_dim"]
    for T in (37, 160, 421):
        f = _feats(rng, T, D)
        # GetFrame at non-decreasing frames, the "no i-vector frame ready" sentinel in front, a repeated frame, the last frame
        sched = [-1] + sorted(rng.integers(0, T, size=9).tolist()) + [T - 1, T - 1]
        a = ref.run(f, sched, online_cmvn_iextractor=online_cmvn)
        b = ref.run_real(f, sched, online_cmvn_iextractor=online_cmvn)
        assert np.array_equal(a, b)
        assert np.abs(a[1:]).max() > 0 and not a[0].any()


@pytest.mark.parametrize("max_remembered", [1000.0, 120.0, -1.0])
def test_speaker_state_over_utterances_equals_the_references_class(ref, max_remembered):
    rng = np.random.default_rng(8)
    D = ref.ex["base_dim"]
    sa, sb = ref.new_speaker(), ref.new_speaker()
    for T in (150, 90, 260):
        f = _feats(rng, T, D)
        sched = list(range(9, T, 30)) + [T - 1]
        a = ref.run(f, sched, speaker=sa, max_remembered_frames=max_remembered)
        b = ref.run_real(f, sched, speaker=sb, max_remembered_frames=max_remembered)
        assert np.array_equal(a, b)
        assert np.array_equal(ref.speaker_state(sa), ref.speaker_state(sb))
    ref.lib.ref_ivector_speaker_destroy(sa); ref.lib.ref_ivector_speaker_destroy(sb)


def test_zero_delta_weights_change_nothing_and_silence_weights_do(ref):
    """UpdateFrameWeights with weight 1 on every frame is the unweighted extraction; down-weighting a stretch moves the i-vector."""
    rng = np.random.default_rng(4)
    D, T = ref.ex["base_dim"], 200
    f = _feats(rng, T, D)
    sched = [49, 99, 149, 199]
    plain = ref.run_real(f, sched)
    ones = [[(t, 1.0) for t in range(lo, hi + 1)] for lo, hi in ((0, 49), (50, 99), (100, 149), (150, 199))]
    assert np.array_equal(ref.run_real(f, sched, delta_weights=ones), plain)
    # frames 50..99 counted as silence (weight 0.001) when they arrive, and 0..20 taken back afterwards
    w = [list(d) for d in ones]
    w[1] = [(t, 0.001) for t in range(50, 100)]
    w[2] = w[2] + [(t, 0.001 - 1.0) for t in range(0, 21)]
    got = ref.run_real(f, sched, delta_weights=w)
    assert np.array_equal(got[0], plain[0]) and np.abs(got[1] - plain[1]).max() > 1e-4 and np.abs(got[3] - plain[3]).max() > 1e-4
