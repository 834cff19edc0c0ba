"""The C/C++ route from a model file to log-likelihoods — b2k_model_read → b2k_nnet_compile →
b2k_nnet_create_from_program → b2k_nnet_run — on tests/golden/tiny_final.mdl, a binary final.mdl written by the
reference's own TransitionModel::Write + AmNnetSimple::Write (tests/golden/make_model_golden.py), against
tests/golden/nnet_golden.npz = the reference's own forward of that model.  CPU: reader, compiler and the ABI program
(interpreted in numpy); GPU: the same through the device.  Tolerance 1e-4 of the output scale (north star)."""
import ctypes as C
import os

import numpy as np
import pytest

from kaldi_b200 import nnet_model as NM

HERE = os.path.dirname(__file__)
MDL = os.path.join(HERE, "golden", "tiny_final.mdl")
GOLD = os.path.join(HERE, "golden", "nnet_golden.npz")
T2P = os.path.join(HERE, "golden", "tiny_final_tid2pdf.npy")
RTOL_SCALE = 1e-4


def _model():
    try:
        from kaldi_b200.model import KaldiModel
        return KaldiModel(MDL)
    except OSError as e:                       # libb2k.so not built
        pytest.skip(str(e))


def test_reader_returns_the_model_that_was_written():
    m = _model()
    arch = NM.arch_tiny(64)
    W = NM.random_weights(arch, seed=11)
    assert (m.feat_dim, m.ivector_dim, m.num_pdfs, m.frame_subsampling_factor) == (
        arch["feat_dim"], arch["ivector_dim"], arch["num_pdfs"], arch["frame_subsampling_factor"])
    assert m.has_priors
    assert [t for t, _ in m.layer_types()] == [L["type"] for L in arch["layers"]]
    got = m.weights()
    for k, v in W.items():
        tol = dict(rtol=2e-5, atol=1e-6) if k.endswith((".mean", ".var")) else dict(rtol=0, atol=0)
        np.testing.assert_allclose(got[k].reshape(v.shape), v, err_msg=k, **tol)
    np.testing.assert_array_equal(m.tid2pdf, np.load(T2P))
    # and the Python reader agrees
    arch2, W2, t2p2 = NM.load_kaldi_mdl(MDL)
    np.testing.assert_array_equal(t2p2, m.tid2pdf)
    for k in W2:
        np.testing.assert_array_equal(np.asarray(W2[k], np.float32).reshape(-1), got[k].reshape(-1), err_msg=k)


def test_compiled_abi_program_reproduces_the_reference_forward_on_cpu():
    from kaldi_b200 import _lib
    from kaldi_b200.nnet import _Node, _Op
    from oracle import program_interp as PI
    m = _model()
    g = np.load(GOLD)
    L = _lib.lib()
    prog = m.compile(num_frames=g["feats"].shape[0])
    try:
        nn, no, bl = C.c_int32(), C.c_int32(), C.c_int64()
        L.b2k_nnet_program_sizes.argtypes = [C.c_void_p] * 4
        assert L.b2k_nnet_program_sizes(prog, C.byref(nn), C.byref(no), C.byref(bl)) == 0
        for f, rt in (("b2k_nnet_program_nodes", C.POINTER(_Node)), ("b2k_nnet_program_ops", C.POINTER(_Op)),
                      ("b2k_nnet_program_blob", C.POINTER(C.c_float))):
            getattr(L, f).restype = rt
            getattr(L, f).argtypes = [C.c_void_p]
        nodes = [L.b2k_nnet_program_nodes(prog)[i] for i in range(nn.value)]
        ops = [L.b2k_nnet_program_ops(prog)[i] for i in range(no.value)]
        blob = np.ctypeslib.as_array(L.b2k_nnet_program_blob(prog), shape=(bl.value,)).copy()
        out = PI.run_program(PI.program_from_abi(nodes, ops, blob), g["feats"], g["chunk_ivectors"])
    finally:
        L.b2k_nnet_program_destroy.argtypes = [C.c_void_p]
        L.b2k_nnet_program_destroy(prog)
    assert out.shape == g["ref_out"].shape
    assert np.abs(out - g["ref_out"]).max() <= RTOL_SCALE * np.abs(g["ref_out"]).max()


def test_device_creation_fails_loudly_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from kaldi_b200.nnet import NnetComputer
    with pytest.raises(Exception):
        NnetComputer.from_model(_model(), num_frames=100, max_batch=1)


@pytest.mark.gpu
def test_model_file_to_loglikes_on_the_device():
    from kaldi_b200.model import KaldiModel
    from kaldi_b200.nnet import NnetComputer
    g = np.load(GOLD)
    m = KaldiModel(MDL)
    nc = NnetComputer.from_model(m, num_frames=g["feats"].shape[0], max_batch=2)
    assert (nc.n_out, nc.n_chunks) == (g["ref_out"].shape[0], g["chunk_ivectors"].shape[0])
    out = nc.forward([g["feats"], g["feats"]], [g["chunk_ivectors"], g["chunk_ivectors"]])
    for o in out:
        assert np.abs(o - g["ref_out"]).max() <= RTOL_SCALE * np.abs(g["ref_out"]).max()
    # the Python-compiled program of the same model gives the same numbers
    arch = NM.arch_tiny(64)
    ref = NnetComputer(arch, NM.random_weights(arch, seed=11), num_frames=g["feats"].shape[0], max_batch=2)
    out2 = ref.forward([g["feats"], g["feats"]], [g["chunk_ivectors"], g["chunk_ivectors"]])
    np.testing.assert_allclose(out[0], out2[0], rtol=1e-5, atol=1e-5)


@pytest.mark.gpu
@pytest.mark.parametrize("with_ivectors,int16", [(True, False), (False, True)])
def test_cpp_pipeline_equals_python_pipeline(with_ivectors, int16):
    """b2k_pipeline_* (model file -> C++ orchestration) against BatchedPipeline (Python orchestration of the same
    stage calls) on the same waveforms: identical features and i-vectors, log-likelihoods equal up to the one-ulp
    differences of the two compilers' derived BatchNorm scales, and the lattice bit-exact with the decoder oracle
    run on the C++ pipeline's own log-likelihoods."""
    from kaldi_b200 import ivector as IVM, synth
    from kaldi_b200.decoder import CudaDecoder, CudaFst, lattice_to_canonical
    from kaldi_b200.model import KaldiModel
    from kaldi_b200.pipeline import BatchedPipeline, NativeBatchedPipeline, PipelineConfig, native_plan
    from oracle import dec_oracle as D
    S, B = 32000, 3
    g = synth.make_hclg(50_000, num_pdfs=64, seed=4)
    cfg = PipelineConfig(max_batch=B, num_samples=S, extract_ivectors=with_ivectors)
    m = KaldiModel(MDL)
    T = native_plan(cfg, m)["num_feature_frames"]
    ex = IVM.make_synthetic_extractor(3, num_gauss=64, ivector_dim=100) if with_ivectors else None
    ivx = IVM.IvectorExtractorGpu(ex, B, T) if with_ivectors else None
    nat = NativeBatchedPipeline(cfg, m, CudaFst(g), ivx)
    waves = [synth.make_audio(S, seed=20 + i) for i in range(B)]
    if int16:
        waves = [np.clip(np.round(w), -32768, 32767).astype(np.int16) for w in waves]
    lats = CudaDecoder.SplitLattices(nat.decode_batch(waves))
    arch = NM.arch_tiny(64)
    py = BatchedPipeline(cfg, arch, NM.random_weights(arch, seed=11), g, ivector_extractor=ex)
    py.decode_batch(waves)
    np.testing.assert_array_equal(nat.read("features", B), py.d_feats[:B].cpu().numpy())
    np.testing.assert_array_equal(nat.read("ivectors", B), py.d_ivec[:B].cpu().numpy())
    ll = nat.read("loglikes", B)
    np.testing.assert_allclose(ll, py.d_loglikes[:B].cpu().numpy(), rtol=1e-4, atol=1e-4)
    for i in range(B):
        o = D.DecoderOracle(g, cfg.decoder_cfg)
        o.decode(ll[i], mode=D.MODE_REFERENCE_ORDER)
        want, got = o.lattice(), lattice_to_canonical(lats[i])
        assert all(np.array_equal(got[k], want[k]) for k in got)


@pytest.mark.gpu
def test_graph_file_route_gives_the_same_decoder(tmp_path):
    """HCLG.fst -> b2k_fst_file_read -> b2k_fst_create_from_file against b2k_fst_create on the in-memory CSR: the
    decoder's raw lattice is identical."""
    import torch
    from kaldi_b200 import kaldi_io as KIO, synth
    from kaldi_b200.decoder import CudaDecoder, CudaDecoderConfig, CudaFst, lattice_to_canonical
    g = synth.make_hclg(50_000, num_pdfs=64, seed=4)
    p = str(tmp_path / "HCLG.fst")
    KIO.write_openfst(p, g, "const")
    rng = np.random.default_rng(3)
    ll = torch.from_numpy((rng.standard_normal((40, 64)) * 2.0).astype(np.float32)).cuda()
    lats = []
    for fst in (CudaFst(g), CudaFst.from_file(p, g["tid2pdf"], num_pdfs=64)):
        dc = CudaDecoderConfig.from_dict(synth.DEFAULT_DECODER_CFG, max_frames=64, max_tokens=1_500_000, max_links=3_000_000)
        dec = CudaDecoder(fst, dc, 1)
        dec.InitDecoding([0])
        dec.AdvanceDecodingFrames([0], [ll.data_ptr()], [40], 64)
        dec.FinalizeDecoding([0])
        assert dec.ChannelInfo(0)["status"] == 0
        lats.append(lattice_to_canonical(dec.GetRawLattice(0)))
    assert all(np.array_equal(lats[0][k], lats[1][k]) for k in lats[0])
    assert lats[0]["states"].shape[0] > 0


@pytest.mark.gpu
def test_pure_cpp_route_from_files_to_compact_lattices(tmp_path):
    """tests/cabi/pipeline_device_route.cc: model file + graph file + chunked audio -> UtteranceBatcher ->
    B2kPipelineBackend -> compact lattices, no Python in between; against the ctypes view of the same C++ pipeline
    (NativeBatchedPipeline) followed by lattice.determinize_pruned."""
    import shutil
    import subprocess
    from kaldi_b200 import kaldi_io as KIO, synth
    from kaldi_b200.decoder import CudaDecoder, CudaFst
    from kaldi_b200.lattice import determinize_pruned
    from kaldi_b200.model import KaldiModel
    from kaldi_b200.pipeline import NativeBatchedPipeline, PipelineConfig
    if not shutil.which("g++"):
        pytest.skip("g++ missing")
    root = os.path.dirname(HERE)
    S, n = 32000, 3
    g = synth.make_hclg(50_000, num_pdfs=64, seed=4)
    fst_path, wav_path, t2p_path, out_path = (str(tmp_path / x) for x in ("HCLG.fst", "waves.f32", "tid2pdf.i32", "out.bin"))
    KIO.write_openfst(fst_path, g, "const")
    waves = [synth.make_audio(S, seed=40 + i).astype(np.float32) for i in range(n)]
    np.concatenate(waves).astype("<f4").tofile(wav_path)
    np.asarray(g["tid2pdf"], "<i4").tofile(t2p_path)
    exe = str(tmp_path / "pipeline_device_route")
    so_dir = os.path.join(root, "kaldi_b200")
    r = subprocess.run(["g++", "-std=c++17", "-I" + os.path.join(root, "include"), "-I" + os.path.join(so_dir, "host"),
                        os.path.join(HERE, "cabi", "pipeline_device_route.cc"), "-o", exe, "-L" + so_dir, "-lb2k",
                        "-Wl,-rpath," + so_dir], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([exe, MDL, fst_path, wav_path, str(n), str(S), out_path, t2p_path], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    # read back: per utterance in callback order
    d = open(out_path, "rb").read()
    p, got = 0, {}
    while p < len(d):
        cid, ns, na, nf, nt = np.frombuffer(d, "<i8", 5, p)
        p += 40

        def take(dt, k):
            nonlocal p
            a = np.frombuffer(d, dt, int(k), p).copy()
            p += a.nbytes
            return a
        rec = dict(num_states=int(ns), arc_src=take("<i4", na), arc_dst=take("<i4", na), arc_word=take("<i4", na),
                   arc_graph_cost=take("<f4", na), arc_acoustic_cost=take("<f4", na), arc_off=take("<i8", na + 1),
                   final_state=take("<i4", nf), final_graph_cost=take("<f4", nf), final_acoustic_cost=take("<f4", nf),
                   final_off=take("<i8", nf + 1), tids=take("<i4", nt))
        got[int(cid)] = rec
    assert sorted(got) == [100, 101, 102]
    # the same C++ pipeline through ctypes, then the determinizer
    cfg = PipelineConfig(max_batch=2, num_samples=S, extract_ivectors=False)
    nat = NativeBatchedPipeline(cfg, KaldiModel(MDL), CudaFst(g), None)
    beam = float(cfg.decoder_cfg["lattice_beam"])
    for first in (0, 2):                                     # the batcher decoded utterances {0,1} and then {2}
        batch = waves[first:first + 2]
        lats = CudaDecoder.SplitLattices(nat.decode_batch(batch))
        for j, lat in enumerate(lats):
            want, rec = determinize_pruned(lat, beam), got[100 + first + j]
            assert rec["num_states"] == want["num_states"]
            for k in ("arc_src", "arc_dst", "arc_word", "arc_graph_cost", "arc_acoustic_cost", "final_state",
                      "final_graph_cost", "final_acoustic_cost"):
                np.testing.assert_array_equal(rec[k], want[k], err_msg="%s (utterance %d; finals here %s / %s, there %s / %s)" % (
                    k, first + j, rec["final_graph_cost"], rec["final_acoustic_cost"], want["final_graph_cost"], want["final_acoustic_cost"]))
            tids = np.concatenate([np.concatenate(want["arc_tids"]) if want["arc_tids"] else np.zeros(0, np.int32),
                                   np.concatenate(want["final_tids"]) if want["final_tids"] else np.zeros(0, np.int32)])
            np.testing.assert_array_equal(rec["tids"], tids)


@pytest.mark.gpu
def test_c99_program_decodes_a_directory_on_the_device(tmp_path):
    """The complete C route (option files, final.mdl, HCLG.fst, 8 kHz WAVE resampled, extractor files) down to a binary
    CompactLattice archive that reads back."""
    import shutil
    import struct
    import subprocess
    sys_path_added = os.path.dirname(__file__)
    import sys
    sys.path.insert(0, sys_path_added)
    from test_experiment_dir import _build
    from kaldi_b200.lattice import compact_best_path, read_lattice_archive
    if not shutil.which("gcc"):
        pytest.skip("gcc missing")
    root = os.path.dirname(HERE)
    d = str(tmp_path / "exp")
    hclg, _ = _build(d)
    x = (3000 * np.sin(2 * np.pi * 300 * np.arange(16000) / 8000) + np.random.default_rng(0).normal(0, 300, 16000)).astype("<i2").tobytes()
    body = b"WAVE" + b"fmt " + struct.pack("<IHHIIHH", 16, 1, 1, 8000, 16000, 2, 16) + b"data" + struct.pack("<I", len(x)) + x
    wav, ark = str(tmp_path / "utt.wav"), str(tmp_path / "out.ark")
    open(wav, "wb").write(b"RIFF" + struct.pack("<I", len(body)) + body)
    so_dir = os.path.join(root, "kaldi_b200")
    exe = str(tmp_path / "experiment_route")
    r = subprocess.run(["gcc", "-std=c99", "-I" + os.path.join(root, "include"), os.path.join(HERE, "cabi", "experiment_route.c"), "-o", exe,
                        "-L" + so_dir, "-lb2k", "-Wl,-rpath," + so_dir], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([exe, os.path.join(d, "conf", "online.conf"), os.path.join(d, "final.mdl"), hclg, wav, ark], capture_output=True,
                       text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    (key, kind, clat), = read_lattice_archive(open(ark, "rb").read())
    assert key == "utt" and kind == "compact" and clat["num_states"] > 0 and len(clat["final_state"]) > 0
    bp = compact_best_path(clat)
    assert np.isfinite(bp["total_cost"]) and len(bp["tids"]) == (1 + (32000 - 400) // 160 + 2) // 3      # one transition-id per decoder frame


@pytest.mark.gpu
def test_c99_streaming_program_decodes_chunk_by_chunk_on_the_device(tmp_path):
    """tests/cabi/stream_route.c: the same directory, the utterance fed 0.51 s at a time through b2k_stream_*, a partial hypothesis
    and the end-point rules after every chunk, the compact lattice at the end: one transition-id per decoder frame."""
    import shutil
    import struct
    import subprocess
    import sys
    sys.path.insert(0, os.path.dirname(__file__))
    from test_experiment_dir import _build
    from kaldi_b200.lattice import compact_best_path, read_lattice_archive
    if not shutil.which("gcc"):
        pytest.skip("gcc missing")
    root = os.path.dirname(HERE)
    d = str(tmp_path / "exp")
    hclg, _ = _build(d)
    n = 40000
    x = (3000 * np.sin(2 * np.pi * 300 * np.arange(n) / 16000) + np.random.default_rng(1).normal(0, 300, n)).astype("<i2").tobytes()
    body = b"WAVE" + b"fmt " + struct.pack("<IHHIIHH", 16, 1, 1, 16000, 32000, 2, 16) + b"data" + struct.pack("<I", len(x)) + x
    wav, ark = str(tmp_path / "utt.wav"), str(tmp_path / "out.ark")
    open(wav, "wb").write(b"RIFF" + struct.pack("<I", len(body)) + body)
    so_dir = os.path.join(root, "kaldi_b200")
    exe = str(tmp_path / "stream_route")
    r = subprocess.run(["gcc", "-std=c99", "-I" + os.path.join(root, "include"), os.path.join(HERE, "cabi", "stream_route.c"), "-o", exe,
                        "-L" + so_dir, "-lb2k", "-Wl,-rpath," + so_dir], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([exe, os.path.join(d, "conf", "online.conf"), os.path.join(d, "final.mdl"), hclg, wav, ark], capture_output=True,
                       text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.count("frames decoded") >= 3 and "streamed in" in r.stdout, r.stdout
    (key, kind, clat), = read_lattice_archive(open(ark, "rb").read())
    assert key == "utt" and kind == "compact" and clat["num_states"] > 0 and len(clat["final_state"]) > 0
    bp = compact_best_path(clat)
    frames = 1 + (n - 400) // 160
    if "ended at an end point" not in r.stdout:
        # every call rounds its own output count up: at least ceil(frames / 3) decoder frames, at most one more per call
        assert (frames + 2) // 3 <= len(bp["tids"]) <= (frames + 2) // 3 + 6 and np.isfinite(bp["total_cost"])


@pytest.mark.gpu
@pytest.mark.parametrize("T", [64, 23])
def test_chain_tdnn_without_factorisation_on_the_device(T):
    from kaldi_b200.nnet import NnetComputer
    from oracle import nnet_oracle as NO
    arch = NM.arch_tiny_tdnn()
    W = NM.random_weights(arch, seed=7)
    nc = NnetComputer(arch, W, num_frames=T, max_batch=3, acoustic_scale=0.9)
    rng = np.random.default_rng(T)
    batch = [((rng.standard_normal((T, 40)) * 10).astype(np.float32), rng.standard_normal((nc.n_chunks, 100)).astype(np.float32))
             for _ in range(3)]
    outs = nc.forward([b[0] for b in batch], [b[1] for b in batch])
    for (feats, civ), o in zip(batch, outs):
        mine = NO.forward_dense(arch, W, feats, civ, frames_per_chunk=21, acoustic_scale=0.9)
        assert o.shape == mine.shape
        assert np.abs(o - mine).max() <= RTOL_SCALE * np.abs(mine).max()
