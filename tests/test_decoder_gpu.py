"""GPU parity tests: CUDA decoder (through the C-ABI) vs the oracle, bit-exact.

Compared objects (SURVEY.md §7 hard part 1): (i) the un-pruned token set
{(state, tot_cost bits)} and link set of EVERY frame, (ii) per-frame
next_cutoff / cost_offset, (iii) the finalized raw lattice as a canonical set.
"""
import numpy as np
import pytest

from kaldi_b200 import synth

pytestmark = pytest.mark.gpu


REF, FREE = True, False      # reference_order=True <-> oracle mode 0; False <-> oracle mode 1


def _mk(g, cfg, nlanes=1, nchannels=None, T=64, ref=REF, **kw):
    from kaldi_b200.decoder import CudaFst, CudaDecoder, CudaDecoderConfig
    fst = CudaFst(g)
    c = CudaDecoderConfig.from_dict(cfg, max_frames=max(T + 2, 16), max_tokens=kw.get("max_tokens", 1_500_000),
                                    max_links=kw.get("max_links", 3_000_000), reference_order=ref,
                                    max_tokens_per_frame=kw.get("max_tpf", 32768))
    return fst, CudaDecoder(fst, c, nlanes, nchannels)


def _run_gpu(dec, ll_list, chunk=None, one_frame_api=False):
    import torch
    n = len(ll_list)
    d = [torch.from_numpy(np.ascontiguousarray(x)).cuda() for x in ll_list]
    ch = list(range(n))
    dec.InitDecoding(ch)
    T = [x.shape[0] for x in ll_list]
    if one_frame_api:
        for f in range(max(T)):
            la = [(c, d[c][f].data_ptr()) for c in ch if f < T[c]]
            dec.AdvanceDecoding(la)
    elif chunk:
        done = [0] * n
        while any(done[c] < T[c] for c in ch):
            act = [c for c in ch if done[c] < T[c]]
            nf = [min(chunk, T[c] - done[c]) for c in act]
            dec.AdvanceDecodingFrames(act, [d[c][done[c]].data_ptr() for c in act], nf, d[0].stride(0))
            for c, k in zip(act, nf):
                done[c] += k
    else:
        dec.AdvanceDecodingFrames(ch, [x.data_ptr() for x in d], T, d[0].stride(0))
    dec.FinalizeDecoding(ch)
    torch.cuda.synchronize()


def _check_against_oracle(g, cfg, ll, dec, channel, frames=True):
    from kaldi_b200.decoder import lattice_to_canonical
    from oracle import dec_oracle as D
    o = D.DecoderOracle(g, cfg)
    ref = dec.config.reference_order
    o.decode(ll, mode=D.MODE_REFERENCE_ORDER if ref else D.MODE_ORDER_FREE, record_frames=frames)
    info = dec.ChannelInfo(channel)
    assert info["status"] == 0, info
    assert info["frames_decoded"] == ll.shape[0]
    T = ll.shape[0]
    fi_o = o.frame_info()
    fi_g = dec.FrameInfo(channel, T)
    np.testing.assert_array_equal(fi_g["cost_offset"].view(np.int32), fi_o["cost_offset"].view(np.int32))
    np.testing.assert_array_equal(fi_g["cutoff"].view(np.int32), fi_o["cutoff"].view(np.int32))
    np.testing.assert_array_equal(fi_g["ntoks"], fi_o["ntoks"])
    if frames:
        for f in range(0, T + 1):
            ts, tc, lk = dec.DebugFrame(channel, f)
            got = D.canonical_raw_frame(ts, tc, lk)
            want = o.raw_frame(f)
            assert np.array_equal(got["toks"], want["toks"]), f"token set differs on frame {f}"
            assert np.array_equal(got["links"], want["links"]), f"link set differs on frame {f}"
    got = lattice_to_canonical(dec.GetRawLattice(channel))
    want = o.lattice()
    for k in ("states", "arcs", "finals"):
        assert got[k].shape == want[k].shape, (k, got[k].shape, want[k].shape)
        assert np.array_equal(got[k], want[k]), f"finalized lattice differs in {k}"
    st = o.stats()
    assert info["arcs_emitting"] == st["arcs_emitting"]
    # eps arcs: the GPU counts the arcs its parallel closure examined, the CPU its LIFO re-visits;
    # only "did any eps arc get examined" is comparable
    assert (info["arcs_nonemitting"] > 0) == (st["arcs_nonemitting"] > 0)
    return st


def _check_against_compiled_reference(g, cfg, ll, dec, channel):
    """Directly against the reference's own lattice-faster-decoder.cc (oracle/_ref, built by
    oracle/ref_decoder.py): every frame's token list IN HASHLIST ORDER (= the order of the
    channel's token arena) and the finalized raw lattice, bit for bit."""
    from kaldi_b200.decoder import lattice_to_canonical
    from oracle import dec_oracle as D
    from oracle import ref_decoder as R
    if not R.available():
        pytest.skip("oracle/_ref decoder library not present")
    r = R.RefDecoder(g, cfg)
    r.decode(ll)
    for f in range(ll.shape[0] + 1):
        ts, tc, _ = dec.DebugFrame(channel, f)
        st, co = r.frame_tokens(f)
        assert np.array_equal(ts, st), f"token list order of frame {f} differs from the reference's HashList order"
        assert np.array_equal(tc.view(np.int32), co.view(np.int32)), f"token costs of frame {f} differ"
    assert D.lattices_equal(lattice_to_canonical(dec.GetRawLattice(channel)), r.lattice())


@pytest.mark.parametrize("seed,cfgmod", [(0, {}), (2, {"max_active": 3000}), (4, {"beam": 8.0, "min_active": 2000}),
                                         (5, {"prune_interval": 7})])   # the reference prunes every 7 frames; the GPU only at the end
def test_gpu_equals_compiled_reference_decoder(seed, cfgmod):
    g = synth.make_hclg(400_000, num_pdfs=800, seed=seed)
    T = 50
    ll = synth.make_loglikes(g, T, seed=seed + 100)
    cfg = dict(synth.DEFAULT_DECODER_CFG, **cfgmod)
    fst, dec = _mk(g, cfg, T=T, ref=REF)
    _run_gpu(dec, [ll])
    assert dec.ChannelInfo(0)["status"] == 0
    _check_against_compiled_reference(g, cfg, ll, dec, 0)


@pytest.mark.parametrize("ref", [REF, FREE])
def test_tiny_graph(ref):
    g = synth.tiny_graph()
    cfg = dict(synth.DEFAULT_DECODER_CFG, min_active=0)
    ll = np.array([[0.0, 3.0], [1.0, 0.0]], np.float32)
    fst, dec = _mk(g, cfg, T=2, max_tokens=1000, max_links=1000, ref=ref)
    _run_gpu(dec, [ll])
    _check_against_oracle(g, cfg, ll, dec, 0)


@pytest.mark.parametrize("ref", [REF, FREE])
@pytest.mark.parametrize("seed,cfgmod", [
    (0, {}),                                            # recipe settings: beam 15, max-active 7000, min-active 200
    (1, {"max_active": 2**31 - 1, "min_active": 0, "beam": 9.0}),    # GetCutoff fast path (up to 61 k tokens/frame)
    (2, {"max_active": 3000}),                          # max_active fires on most frames
    (3, {"beam": 10.0, "lattice_beam": 6.0}),
    (4, {"beam": 8.0, "min_active": 2000}),             # min_active branch
])
def test_frames_and_lattice_bit_exact(seed, cfgmod, ref):
    g = synth.make_hclg(400_000, num_pdfs=800, seed=seed)
    T = 50
    ll = synth.make_loglikes(g, T, seed=seed + 100)
    cfg = dict(synth.DEFAULT_DECODER_CFG, **cfgmod)
    fst, dec = _mk(g, cfg, T=T, ref=ref, max_tpf=131072 if cfgmod.get("min_active") == 0 else 32768)
    _run_gpu(dec, [ll])
    st = _check_against_oracle(g, cfg, ll, dec, 0)
    if "max_active" in cfgmod and cfgmod["max_active"] == 3000:
        assert st["max_active_branch"] > 5
    if cfgmod.get("min_active") == 2000:
        assert st["min_active_branch"] > 0


@pytest.mark.parametrize("ref", [REF, FREE])
def test_multi_lane_chunked_and_one_frame_api(ref):
    g = synth.make_hclg(300_000, num_pdfs=500, seed=7)
    cfg = dict(synth.DEFAULT_DECODER_CFG)
    lls = [synth.make_loglikes(g, T, seed=20 + i) for i, T in enumerate([30, 17, 25, 30])]
    # whole-utterance, ragged lengths
    fst, dec = _mk(g, cfg, nlanes=4, T=32, max_tokens=600_000, max_links=1_200_000, ref=ref)
    _run_gpu(dec, lls)
    for c in range(4):
        _check_against_oracle(g, cfg, lls[c], dec, c, frames=False)
    # batched packed read-back == per-channel read-back
    from kaldi_b200.decoder import CudaDecoder, lattice_to_canonical
    packed = dec.GetRawLattices([0, 1, 2, 3])
    for c, lat in enumerate(CudaDecoder.SplitLattices(packed)):
        a, b = lattice_to_canonical(lat), lattice_to_canonical(dec.GetRawLattice(c))
        assert all(np.array_equal(a[k], b[k]) for k in a)
        assert lat["state_frame"][0] == 0                     # lattice state 0 is the start state
    # chunked (7 frames per call) must give the same thing; channels are reusable
    _run_gpu(dec, lls, chunk=7)
    for c in range(4):
        _check_against_oracle(g, cfg, lls[c], dec, c, frames=False)
    # reference-style one-frame AdvanceDecoding(lanes_assignments)
    _run_gpu(dec, lls, one_frame_api=True)
    for c in range(4):
        _check_against_oracle(g, cfg, lls[c], dec, c, frames=(c == 1))


def test_dead_end_graph_no_surviving_tokens():
    g = dict(num_states=2, start=0, num_pdfs=1, offsets=np.array([0, 1, 1], np.int32),
             ilabel=np.array([1], np.int32), olabel=np.array([0], np.int32),
             weight=np.array([0.5], np.float32), nextstate=np.array([1], np.int32),
             final=np.array([np.inf, np.inf], np.float32), tid2pdf=np.array([0, 0], np.int32))
    cfg = dict(synth.DEFAULT_DECODER_CFG)
    ll = np.zeros((3, 1), np.float32)
    for ref in (REF, FREE):
        fst, dec = _mk(g, cfg, T=3, max_tokens=100, max_links=100, ref=ref)
        _run_gpu(dec, [ll])
        info = dec.ChannelInfo(0)
        assert info["status"] == 0 and info["frames_decoded"] == 3
        fi = dec.FrameInfo(0, 3)
        assert fi["ntoks"].tolist() == [1, 0, 0]


def test_overflow_is_reported_not_silent():
    g = synth.make_hclg(300_000, num_pdfs=500, seed=9)
    cfg = dict(synth.DEFAULT_DECODER_CFG)
    ll = synth.make_loglikes(g, 20, seed=1)
    fst, dec = _mk(g, cfg, T=20, max_tokens=5_000, max_links=10_000)   # far too small
    _run_gpu(dec, [ll])
    info = dec.ChannelInfo(0)
    assert info["status"] == 4                    # B2K_ERR_OVERFLOW
    assert info["err_line"] > 0                   # which capacity check fired (source line of decoder.cu)
    from kaldi_b200._lib import B2kError
    with pytest.raises(B2kError, match=r"decoder\.cu:\d+"):
        dec.GetRawLattice(0)


@pytest.mark.parametrize("env", [
    {"B2K_DEC_THREADS": "512"},                    # the CTA shape used for batches larger than one wave
    {"B2K_DEC_THREADS": "256"},
    {"B2K_DEC_RS_CAPS": "0,0,0"},                  # replay walk over the global records, full first list order
    {"B2K_DEC_RS_CAPS": "64,64,64"},               # small frames in shared memory, the others fall back mid-utterance
    {"B2K_FIN_THREADS": "256"},
], ids=lambda e: ",".join(f"{k}={v}" for k, v in e.items()))
def test_tuning_knobs_do_not_change_results(env, monkeypatch):
    """Every alternative code path behind a tuning knob (CTA width, shared-memory replay
    capacities and their fallbacks) must stay bit-exact."""
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    g = synth.make_hclg(400_000, num_pdfs=800, seed=0)
    T = 40
    ll = synth.make_loglikes(g, T, seed=100)
    cfg = dict(synth.DEFAULT_DECODER_CFG)
    fst, dec = _mk(g, cfg, T=T)
    _run_gpu(dec, [ll])
    _check_against_oracle(g, cfg, ll, dec, 0)


def test_advance_before_init_is_a_state_error():
    import torch
    g = synth.make_hclg(50_000, num_pdfs=100, seed=1)
    fst, dec = _mk(g, dict(synth.DEFAULT_DECODER_CFG), T=4, max_tokens=10_000, max_links=10_000)
    ll = torch.zeros(4, 100, device="cuda")
    dec.AdvanceDecodingFrames([0], [ll.data_ptr()], [4], 100)
    assert dec.ChannelInfo(0)["status"] == 5      # B2K_ERR_STATE


def test_full_length_utterance_properties():
    """BASELINE-size case (333 decoder frames): lattice invariants that do not
    need the oracle to finish quickly: every arc's endpoints exist, extra_cost
    in [0, lattice_beam], a zero-extra-cost path exists from start to a final."""
    g = synth.make_hclg(2_000_000, num_pdfs=2336, seed=11)
    cfg = dict(synth.DEFAULT_DECODER_CFG)
    T = 333
    ll = synth.make_loglikes(g, T, seed=5)
    fst, dec = _mk(g, cfg, T=T, max_tokens=4_000_000, max_links=8_000_000)
    _run_gpu(dec, [ll])
    info = dec.ChannelInfo(0)
    assert info["status"] == 0 and info["frames_decoded"] == T
    lat = dec.GetRawLattice(0)
    ex = lat["state_extra_cost"]
    assert np.all(ex >= 0) and np.all(ex <= cfg["lattice_beam"])
    assert lat["arc_src"].min() >= 0 and lat["arc_dst"].min() >= 0
    assert lat["state_frame"].max() == T and lat["state_frame"].min() == 0
    assert (lat["state_frame"] == 0).sum() >= 1
    # frames are non-decreasing along arcs, +1 exactly for emitting arcs
    df = lat["state_frame"][lat["arc_dst"]] - lat["state_frame"][lat["arc_src"]]
    assert np.array_equal(df, (lat["arc_ilabel"] != 0).astype(np.int32))
    assert (ex == 0).sum() >= T + 1          # the best path has extra_cost 0 on every frame
    # and the whole thing equals the oracle too
    _check_against_oracle(g, cfg, ll, dec, 0, frames=False)


@pytest.mark.parametrize("seed", [3, 11])
def test_state_zero_is_the_start_state_and_determinization_is_repeatable(seed):
    """include/b2k.h promises "state 0 = start" for a raw lattice and b2k_lat_determinize_pruned starts there.  Lattice-state
    ids are handed out by atomics while the sweep walks backwards: when the first token list keeps an eps-successor of the
    start beside the start token, either of them used to end up as state 0, and the compact lattice of the same utterance
    changed from run to run (found with tools/determinism_probe.py; the order-free comparisons above cannot see it)."""
    from kaldi_b200 import synth
    from kaldi_b200.lattice import determinize_pruned
    g = synth.make_hclg(50_000, num_pdfs=64, seed=4)
    cfg = dict(synth.DEFAULT_DECODER_CFG)
    ll = synth.make_loglikes(g, 60, seed=seed)
    beam = float(cfg["lattice_beam"])
    first = None
    for rep in range(12):
        fst, dec = _mk(g, cfg, T=60)
        _run_gpu(dec, [ll])
        lat = dec.GetRawLattice(0)
        assert lat["state_frame"][0] == 0 and lat["state_hclg"][0] == fst.Start()
        det = determinize_pruned(lat, beam)
        # (compared as multisets: the numbering of the compact states may follow the order of the raw arrays)
        cur = {k: np.sort(det[k]) for k in ("arc_word", "arc_graph_cost", "arc_acoustic_cost", "final_graph_cost", "final_acoustic_cost")}
        cur["num_states"] = np.asarray(det["num_states"])
        if first is None:
            first = cur
        for k in cur:
            np.testing.assert_array_equal(cur[k], first[k], err_msg=f"run {rep}: {k}")
