"""Chunk-by-chunk decoding of several streams at once (kaldi_b200/streaming.py: the call structure of
BatchedThreadedNnet3CudaOnlinePipeline::DecodeBatch over b2k_feat_compute_batched / b2k_nnet_stream_run_batch /
b2k_dec_advance_decoding_frames / b2k_dec_best_path): streams of different lengths that start and end at different calls."""
import numpy as np
import pytest

from kaldi_b200 import nnet_model as NM, synth

pytestmark = pytest.mark.gpu


def test_streams_decode_chunk_by_chunk():
    from kaldi_b200.decoder import lattice_to_canonical
    from kaldi_b200.feat import BatchedFeatures, FeatureOptions
    from kaldi_b200 import lattice as LAT
    from kaldi_b200.streaming import StreamingBatchedDecoder
    from oracle import dec_oracle as D
    P, fpc = 200, 21
    arch = NM.arch_tiny(P)
    W = NM.random_weights(arch, seed=2)
    g = synth.make_hclg(150_000, num_pdfs=P, seed=4)
    cfg = dict(synth.DEFAULT_DECODER_CFG)
    sd = StreamingBatchedDecoder(arch, W, g, cfg, nchannels=4, max_seconds=6.0, frames_per_chunk=fpc)
    lens = [48000, 30000, 16000]
    audio = [synth.make_audio(n, seed=70 + i) for i, n in enumerate(lens)]
    chunk = fpc * 160                                   # at most 21 new frames per call
    chan = [3, 0, 2]
    start_call = [0, 2, 1]                              # the streams do not start together
    pos = [0, 0, 0]
    ll = {c: [] for c in chan}
    lattices, partial_frames, last_partial = {}, {c: [] for c in chan}, {}
    call = 0
    while len(lattices) < 3:
        act = [u for u in range(3) if call >= start_call[u] and pos[u] < lens[u]]
        if act:
            pieces = [audio[u][pos[u]:pos[u] + chunk] for u in act]
            first = [pos[u] == 0 for u in act]
            last = [pos[u] + chunk >= lens[u] for u in act]
            keep = []
            res = sd.DecodeBatch([chan[u] for u in act], pieces, first, last, keep_loglikes=keep)
            for (c, x) in keep:
                ll[c].append(x)
            for u, r in zip(act, res):
                pos[u] += chunk
                partial_frames[chan[u]].append(r["frames_decoded"])
                last_partial[chan[u]] = r["partial_words"]
                if "lattice" in r:
                    lattices[chan[u]] = r["lattice"]
        call += 1
    feat = BatchedFeatures(FeatureOptions(max_lanes=8))
    offline = feat.compute(audio)
    for u, c in enumerate(chan):
        T = offline[u].shape[0]
        # the features computed as the samples arrived are the offline features (bit-exact), read from the channel's buffer
        np.testing.assert_array_equal(sd.d_feats[c, :T].cpu().numpy(), offline[u])
        x = np.concatenate(ll[c], 0)
        # every call rounds its own output count up (batched-static-nnet3.cc:181-186): at least ceil(T / 3) frames in total
        n_calls = len(ll[c])
        assert x.shape[1] == P and (T + 2) // 3 <= x.shape[0] <= (T + 2) // 3 + n_calls and np.isfinite(x).all()
        assert partial_frames[c] == sorted(partial_frames[c]) and partial_frames[c][-1] == x.shape[0]
        # the lattice of the stream = the reference-order decoder on exactly the frames the stream was given
        o = D.DecoderOracle(g, cfg)
        o.decode(x, mode=D.MODE_REFERENCE_ORDER)
        got, want = lattice_to_canonical(lattices[c]), o.lattice()
        assert all(np.array_equal(got[k], want[k]) for k in got), f"stream on channel {c}"
        sp = LAT.best_path(lattices[c])
        assert np.isfinite(sp["total_cost"]) and last_partial[c] is not None


def test_native_stream_pipeline_equals_the_python_composition():
    """b2k_stream_* (C++ orchestration, 16-bit PCM in, the feature kernel reading it directly) against StreamingBatchedDecoder
    (the same stage calls from Python, validated against the reference-order decoder above) on the same chunks: the same number
    of output frames per call, the same partial hypotheses, bit-identical lattices; and its misuse is refused."""
    from kaldi_b200 import _lib
    from kaldi_b200.decoder import CudaFst, lattice_to_canonical
    from kaldi_b200.model import KaldiModel
    from kaldi_b200.streaming import NativeStreamingDecoder, StreamingBatchedDecoder
    P, fpc = 200, 21
    arch = NM.arch_tiny(P)
    W = NM.random_weights(arch, seed=2)
    g = synth.make_hclg(150_000, num_pdfs=P, seed=4)
    cfg = dict(synth.DEFAULT_DECODER_CFG)
    py = StreamingBatchedDecoder(arch, W, g, cfg, nchannels=4, max_seconds=6.0, frames_per_chunk=fpc)
    nat = NativeStreamingDecoder(KaldiModel.from_arch(arch, W), CudaFst(g), cfg, nchannels=4, max_seconds=6.0, frames_per_chunk=fpc)
    assert (nat.opc, nat.fpc, nat.D, nat.P) == (py.opc, fpc, py.D, P)
    lens = [40000, 25000]
    pcm = [np.clip(np.round(synth.make_audio(n, seed=90 + i)), -32768, 32767).astype(np.int16) for i, n in enumerate(lens)]
    chan, chunk, pos = [2, 0], fpc * 160, [0, 0]
    lat_py, lat_nat = {}, {}
    while len(lat_nat) < 2:
        act = [u for u in range(2) if pos[u] < lens[u]]
        pieces = [pcm[u][pos[u]:pos[u] + chunk] for u in act]
        first = [pos[u] == 0 for u in act]
        last = [pos[u] + chunk >= lens[u] for u in act]
        ch = [chan[u] for u in act]
        a = py.DecodeBatch(ch, [p.astype(np.float32) for p in pieces], first, last)
        b = nat.DecodeBatch(ch, pieces, first, last)
        for u, ra, rb in zip(act, a, b):
            assert (ra["new_output_frames"], ra["frames_decoded"]) == (rb["new_output_frames"], rb["frames_decoded"])
            assert np.array_equal(ra["partial_words"], rb["partial_words"]) and ra["partial_cost"] == rb["partial_cost"]
            if "lattice" in ra:
                lat_py[chan[u]], lat_nat[chan[u]] = ra["lattice"], rb["lattice"]
            pos[u] += chunk
    for c in chan:
        x, y = lattice_to_canonical(lat_py[c]), lattice_to_canonical(lat_nat[c])
        assert all(np.array_equal(x[k], y[k]) for k in x) and len(x["states"]) > 0
    with pytest.raises(_lib.B2kError):       # a stream must start with is_first_chunk
        nat.DecodeBatch([1], [pcm[0][:chunk]], [False], [False])
    with pytest.raises(_lib.B2kError):       # more audio than frames_per_chunk frames in one call
        nat.DecodeBatch([1], [pcm[0][:chunk * 3]], [True], [False])
    with pytest.raises(_lib.B2kError):       # the same channel twice
        nat.DecodeBatch([1, 1], [pcm[0][:chunk]] * 2, [True, True], [False, False])
