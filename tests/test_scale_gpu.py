"""Parity at BASELINE scale, on the device, against the reference's own code compiled in oracle/_ref:
  (a) the full-width librispeech tdnn_1d (1536/160 x 17 layers, 6024 pdfs) and cnn_tdnn_1a networks: log-likelihoods within
      1e-4 of the output scale of the reference's looped CPU forward (nnet3/decodable-simple-looped.cc);
  (b) the decoder on a 5 M-arc graph x 333 frames at beams 10 / 15 / 20: every frame's token list IN HASHLIST ORDER and the
      finalized raw lattice bit-identical to decoder/lattice-faster-decoder.cc.  The 50 M-arc graph of BASELINE configs[2]
      (minutes of graph drawing) runs with B2K_BIG_TESTS=1; its log is kept under profiles/."""
import os

import numpy as np
import pytest

from kaldi_b200 import nnet_model as NM
from kaldi_b200 import synth

pytestmark = pytest.mark.gpu
RTOL_SCALE = 1e-4


def _ref_forward(arch, W, feats, civ, acoustic_scale=1.0):
    from oracle import nnet_oracle as NO
    if not (os.path.exists(NO._SO) or os.path.isdir("/root/reference")):
        pytest.skip("oracle/_ref nnet3 library not present")
    R = NO.RefNnet(arch, W, frames_per_chunk=20, acoustic_scale=acoustic_scale)
    n_chunks = civ.shape[0]
    ends = [(n + 1) * R.frames_per_chunk + R.right_context for n in range(n_chunks)]
    mat = np.zeros((ends[-1] + 1, civ.shape[1]), np.float32)
    prev = 0
    for n, e in enumerate(ends):
        mat[prev:e + 1] = civ[n]
        prev = e + 1
    return R.forward(feats, mat, period=1)


@pytest.mark.parametrize("which,T", [("librispeech_1d", 150), ("librispeech_cnn_tdnn_1a", 90)])
def test_full_width_network_vs_compiled_reference(which, T):
    from kaldi_b200.nnet import NnetComputer
    arch = getattr(NM, "arch_" + which)(6024)
    W = NM.random_weights(arch, seed=5)
    nc = NnetComputer(arch, W, num_frames=T, max_batch=2)
    rng = np.random.default_rng(T)
    batch = [((rng.standard_normal((T, 40)) * 10).astype(np.float32), rng.standard_normal((nc.n_chunks, 100)).astype(np.float32))
             for _ in range(2)]
    outs = nc.forward([b[0] for b in batch], [b[1] for b in batch])
    for (feats, civ), o in zip(batch, outs):
        ref = _ref_forward(arch, W, feats, civ)
        assert o.shape == ref.shape
        scale = np.abs(ref).max()
        err = np.abs(o - ref).max()
        assert err <= RTOL_SCALE * scale, (which, err / scale)


def _decode_and_compare(g, beam, T, seed):
    from test_decoder_gpu import _mk, _run_gpu, _check_against_compiled_reference, REF
    ll = synth.make_loglikes(g, T, seed=seed)
    cfg = dict(synth.DEFAULT_DECODER_CFG, beam=beam)
    fst, dec = _mk(g, cfg, T=T, ref=REF, max_tokens=T * 12000, max_links=T * 24000)
    _run_gpu(dec, [ll])
    info = dec.ChannelInfo(0)
    assert info["status"] == 0, info
    _check_against_compiled_reference(g, cfg, ll, dec, 0)
    return info


_G5 = {}


@pytest.mark.parametrize("beam", [10.0, 15.0, 20.0])
def test_decoder_5m_arcs_full_utterance_vs_compiled_reference(beam):
    if "g" not in _G5:
        _G5["g"] = synth.make_hclg(5_000_000, num_pdfs=2336, seed=1)
    info = _decode_and_compare(_G5["g"], beam, 333, seed=4242)
    assert info["frames_decoded"] == 333


@pytest.mark.skipif(not os.environ.get("B2K_BIG_TESTS"), reason="50 M-arc graph: minutes of graph drawing; run with B2K_BIG_TESTS=1")
@pytest.mark.parametrize("beam", [13.0, 17.0])
def test_decoder_50m_arcs_full_utterance_vs_compiled_reference(beam):
    if "g50" not in _G5:
        _G5["g50"] = synth.make_hclg(50_000_000, num_pdfs=6024, seed=1)
    info = _decode_and_compare(_G5["g50"], beam, 333, seed=777)
    assert info["frames_decoded"] == 333


def test_full_width_looped_stream_vs_compiled_reference():
    """The chunked executor at BASELINE width: librispeech tdnn_1d (1536 / 160 x 17 layers, 6024 pdfs, context 40 / 40) evaluated
    chunk by chunk through b2k_nnet_stream in looped mode (windows of 20 + 80 frames, five chunks of i-vector history), an
    i-vector that changes every chunk, against the reference's looped CPU forward over the whole utterance."""
    import torch
    from kaldi_b200.nnet import BatchedStaticNnet3
    from test_nnet_stream import _looped_windows
    arch = NM.arch_librispeech_1d(6024)
    W = NM.random_weights(arch, seed=5)
    T, C = 150, 21
    rng = np.random.default_rng(T)
    feats = (rng.standard_normal((T, 40)) * 10).astype(np.float32)
    n_chunks = (((T + 2) // 3) * 3 + C - 1) // C
    civ = rng.standard_normal((n_chunks, 100)).astype(np.float32)
    from oracle import nnet_oracle as NO
    if not (os.path.exists(NO._SO) or os.path.isdir("/root/reference")):
        pytest.skip("oracle/_ref nnet3 library not present")
    R = NO.RefNnet(arch, W, frames_per_chunk=C)
    assert R.frames_per_chunk == C
    ends = [(n + 1) * C + R.right_context for n in range(n_chunks)]
    mat = np.zeros((ends[-1] + 1, 100), np.float32)
    prev = 0
    for n, e in enumerate(ends):
        mat[prev:e + 1] = civ[n]
        prev = e + 1
    ref = R.forward(feats, mat, period=1)
    nn = BatchedStaticNnet3(arch, W, max_batch=1, frames_per_chunk=C, looped=True)
    L, Rc, k1, opc, P = nn.left_context, nn.right_context, nn.ivector_rows, nn.output_frames_per_chunk, nn.output_dim
    assert (L, Rc) == (R.left_context, R.right_context)
    d_out = torch.zeros(opc, P, device="cuda")
    z = torch.zeros(k1, 100, device="cuda")
    for b in range(0, Rc, C):            # the right context (40 frames) is longer than a chunk (21): it arrives in two calls
        piece = torch.from_numpy(feats[np.clip(np.arange(b, min(b + C, Rc)), 0, T - 1)]).cuda()
        assert nn.RunBatch([0], [piece.data_ptr()], 40, [z.data_ptr()], [piece.shape[0]], [b == 0], [False], d_out.data_ptr(), 0, P) == ([0], [0])
    outs = []
    for n, win, ivr, keep in _looped_windows(arch, feats, civ, C, L, Rc, k1):
        new = torch.from_numpy(np.ascontiguousarray(win[L + Rc:])).cuda()
        d_iv = torch.from_numpy(ivr).cuda()
        no, _ = nn.RunBatch([0], [new.data_ptr()], 40, [d_iv.data_ptr()], [C], [False], [False], d_out.data_ptr(), 0, P)
        torch.cuda.synchronize()
        assert no == [opc]
        outs.append(d_out.cpu().numpy()[:keep].copy())
    got = np.concatenate(outs, 0)
    assert got.shape == ref.shape
    assert np.abs(got - ref).max() <= RTOL_SCALE * np.abs(ref).max(), np.abs(got - ref).max() / np.abs(ref).max()


def test_live_best_path_on_the_5m_arc_graph():
    """b2k_dec_best_path at BASELINE configs[1] scale (5 M arcs, 333 frames, beam 15, 16 lanes): mid-utterance and at the end the
    path costs what the cheapest token of the last frame costs, one transition-id per frame; after FinalizeDecoding it is the
    shortest path of the raw lattice (whose bit-identity with the reference is the test above)."""
    import torch
    from kaldi_b200 import lattice as LAT
    from kaldi_b200.decoder import CudaDecoder, CudaDecoderConfig, CudaFst
    if "g" not in _G5:
        _G5["g"] = synth.make_hclg(5_000_000, num_pdfs=2336, seed=1)
    g, T, B = _G5["g"], 333, 16
    cfg = dict(synth.DEFAULT_DECODER_CFG, beam=15.0)
    dec = CudaDecoder(CudaFst(g), CudaDecoderConfig.from_dict(cfg, max_frames=T + 2, max_tokens=T * 12000, max_links=T * 24000), B)
    ll = torch.from_numpy(np.stack([synth.make_loglikes(g, T, seed=500 + i) for i in range(4)])).cuda()
    ll = ll.repeat(B // 4, 1, 1).contiguous()
    ch = list(range(B))
    dec.InitDecoding(ch)
    done = 0
    for step in (111, 222):
        dec.AdvanceDecodingFrames(ch, [ll[c, done].data_ptr() for c in ch], [step] * B, ll.stride(1))
        done += step
        bps = dec.GetBestPath(ch, use_final_probs=False, cap=4096)
        for c in (0, 5, B - 1):
            ts, tc, _ = dec.DebugFrame(c, done)
            bp = bps[c]
            assert bp["num_frames"] == done and int((bp["ilabels"] != 0).sum()) == done
            assert np.float32(bp["best_cost"]) == tc.min()
            offs = dec.FrameInfo(c, done)["cost_offset"].astype(np.float64).sum()
            tot = bp["graph_costs"].astype(np.float64).sum() + bp["acoustic_costs"].astype(np.float64).sum() + offs
            assert abs(tot - float(tc.min())) <= 1e-4 * abs(float(tc.min())) + 1e-2
        assert np.array_equal(bps[0]["ilabels"], bps[4]["ilabels"])            # lanes 0 and 4 decode the same frames
    dec.FinalizeDecoding(ch)
    torch.cuda.synchronize()
    for c in (0, B - 1):
        assert dec.ChannelInfo(c)["status"] == 0
        bp = dec.GetBestPath([c], use_final_probs=True, cap=4096)[0]
        sp = LAT.best_path(dec.GetRawLattice(c))
        assert np.array_equal(bp["olabels"][bp["olabels"] != 0], sp["olabels"])
        tot = bp["graph_costs"].astype(np.float64).sum() + bp["acoustic_costs"].astype(np.float64).sum() + bp["final_cost"]
        assert abs(tot - sp["total_cost"]) <= 1e-4 * abs(sp["total_cost"]) + 1e-2
