"""b2k_dec_best_path: the best path of a channel that is still decoding -- LatticeFasterOnlineDecoderTpl::BestPathEnd +
TraceBackBestPath (decoder/lattice-faster-online-decoder.cc:78-167), FinalRelativeCost (lattice-faster-decoder.cc:545-586).
Checked through what defines it: the path is a path of HCLG from the start state, it consumes one transition-id per decoded
frame, its acoustic costs are the frames' log-likelihoods, its cost is the cost of the cheapest token of the last frame (the
token lists themselves are compared bit for bit with the reference in test_decoder_gpu.py), and after FinalizeDecoding it is
the shortest path of the raw lattice."""
import numpy as np
import pytest

from kaldi_b200 import synth

pytestmark = pytest.mark.gpu


def _mk(g, cfg, T, nlanes=1):
    from kaldi_b200.decoder import CudaFst, CudaDecoder, CudaDecoderConfig
    fst = CudaFst(g)
    c = CudaDecoderConfig.from_dict(cfg, max_frames=T + 2, max_tokens=1_500_000, max_links=3_000_000)
    return fst, CudaDecoder(fst, c, nlanes)


def _check_path(g, ll, dec, ch, bp, use_final, acwt_tol=5e-3):
    n_frames = bp["num_frames"]
    ts, tc, _ = dec.DebugFrame(ch, n_frames)
    fin = g["final"][ts]
    any_final = np.isfinite(fin).any()
    # BestPathEnd
    if use_final and any_final:
        want_best = (tc.astype(np.float32) + fin.astype(np.float32)).min()
    else:
        want_best = tc.min()
    assert np.float32(bp["best_cost"]) == np.float32(want_best)
    want_rel = np.float32(np.inf) if not any_final else np.float32((tc + fin).min()) - np.float32(tc.min())
    assert np.float32(bp["final_relative_cost"]) == want_rel or (np.isinf(want_rel) and np.isinf(bp["final_relative_cost"]))
    if use_final and any_final:
        assert np.float32(bp["final_cost"]) == g["final"][bp["end_state"]]
    else:
        assert bp["final_cost"] == 0.0
    # a path of HCLG from the start state that ends in the end token's state
    off, il, ol, w, ns = g["offsets"], g["ilabel"], g["olabel"], g["weight"], g["nextstate"]
    cur, frame = g["start"], 0
    for k in range(len(bp["ilabels"])):
        a = np.arange(off[cur], off[cur + 1])
        hit = a[(il[a] == bp["ilabels"][k]) & (ol[a] == bp["olabels"][k]) & (ns[a] == bp["arc_state"][k]) &
                (w[a].view(np.int32) == bp["graph_costs"][k:k + 1].view(np.int32)[0])]
        assert len(hit) >= 1, f"arc {k} of the path is not an arc of HCLG out of state {cur}"
        if bp["ilabels"][k] != 0:
            frame += 1
            pdf = g["tid2pdf"][bp["ilabels"][k]]
            assert abs(bp["acoustic_costs"][k] + ll[frame - 1, pdf]) <= acwt_tol, (k, bp["acoustic_costs"][k], ll[frame - 1, pdf])
        else:
            assert bp["acoustic_costs"][k] == 0.0
        assert bp["arc_frame"][k] == frame
        cur = bp["arc_state"][k]
    assert frame == n_frames and cur == bp["end_state"]
    # its cost is the end token's cost: the token's cost carries the per-frame offsets the acoustic costs had removed
    offs = dec.FrameInfo(ch, n_frames)["cost_offset"].astype(np.float64)
    total = bp["graph_costs"].astype(np.float64).sum() + bp["acoustic_costs"].astype(np.float64).sum() + offs.sum()
    end_cost = float(bp["best_cost"]) - float(bp["final_cost"])
    assert abs(total - end_cost) <= 1e-4 * max(1.0, abs(end_cost)) + 1e-2, (total, end_cost)


@pytest.mark.parametrize("seed,cfgmod", [(1, {}), (3, {"max_active": 2000}), (6, {"beam": 9.0})])
def test_best_path_while_decoding(seed, cfgmod):
    import torch
    g = synth.make_hclg(400_000, num_pdfs=800, seed=seed)
    T = 60
    lls = [synth.make_loglikes(g, T, seed=seed + 50 + i) for i in range(3)]
    cfg = dict(synth.DEFAULT_DECODER_CFG, **cfgmod)
    fst, dec = _mk(g, cfg, T, nlanes=3)
    d = [torch.from_numpy(np.ascontiguousarray(x)).cuda() for x in lls]
    ch = [0, 1, 2]
    dec.InitDecoding(ch)
    bp0 = dec.GetBestPath([1], use_final_probs=False)[0]             # nothing decoded yet: the start state's closure
    assert bp0["num_frames"] == 0 and (bp0["ilabels"] == 0).all()
    done = 0
    words_so_far = None
    for step in (1, 19, 20, 20):
        dec.AdvanceDecodingFrames(ch, [x[done].data_ptr() for x in d], [step] * 3, d[0].stride(0))
        done += step
        for use_final in (False, True):
            bps = dec.GetBestPath(ch, use_final_probs=use_final)
            for c in ch:
                assert bps[c]["num_frames"] == done
                _check_path(g, lls[c], dec, c, bps[c], use_final)
        # asking did not disturb the decoder: same answer twice, and decoding continues to the reference's result below
        again = dec.GetBestPath([2], use_final_probs=False)[0]
        assert np.array_equal(again["ilabels"], dec.GetBestPath(ch, use_final_probs=False)[2]["ilabels"])
        words_so_far = again["olabels"][again["olabels"] != 0]
    assert words_so_far is not None
    dec.FinalizeDecoding(ch)
    torch.cuda.synchronize()
    from kaldi_b200 import lattice as LAT
    from kaldi_b200.decoder import lattice_to_canonical
    from oracle import dec_oracle as D
    for c in ch:
        assert dec.ChannelInfo(c)["status"] == 0
        o = D.DecoderOracle(g, cfg)
        o.decode(lls[c], mode=D.MODE_REFERENCE_ORDER)
        got, want = lattice_to_canonical(dec.GetRawLattice(c)), o.lattice()
        assert all(np.array_equal(got[k], want[k]) for k in got), "the lattice changed because best paths were requested"
        # after FinalizeDecoding: the shortest path of the raw lattice (GetBestPath of the non-online decoder, :102-108)
        bp = dec.GetBestPath([c], use_final_probs=True)[0]
        _check_path(g, lls[c], dec, c, bp, True)
        sp = LAT.best_path(dec.GetRawLattice(c))
        assert np.array_equal(bp["olabels"][bp["olabels"] != 0], sp["olabels"])
        assert np.array_equal(bp["ilabels"][bp["ilabels"] != 0], sp["ilabels"])
        tot = bp["graph_costs"].astype(np.float64).sum() + bp["acoustic_costs"].astype(np.float64).sum() + bp["final_cost"]
        assert abs(tot - sp["total_cost"]) <= 1e-4 * abs(sp["total_cost"]) + 1e-2


def test_best_path_errors():
    from kaldi_b200 import _lib
    g = synth.make_hclg(50_000, num_pdfs=200, seed=2)
    fst, dec = _mk(g, dict(synth.DEFAULT_DECODER_CFG), 16, nlanes=2)
    with pytest.raises(_lib.B2kError):
        dec.GetBestPath([0])                       # InitDecoding has not run on this channel
    with pytest.raises(_lib.B2kError):
        dec.GetBestPath([5])                       # no such channel
    import torch
    ll = synth.make_loglikes(g, 12, seed=3)
    d = torch.from_numpy(ll).cuda()
    dec.InitDecoding([0])
    dec.AdvanceDecodingFrames([0], [d.data_ptr()], [12], d.stride(0))
    with pytest.raises(_lib.B2kError) as e:
        dec.GetBestPath([0], cap=4)                # longer than the caller's buffers
    assert e.value.code == _lib.B2K_ERR_OVERFLOW
    assert len(dec.GetBestPath([0], use_final_probs=False)[0]["ilabels"]) >= 12
