"""The C++ model-file reader of libb2k.so (kaldi_b200/csrc/model_io.cu: b2k_model_read) on files written by the
reference's own Write() methods (oracle/_ref), against the architecture and weights that were put into the
reference model and against the Python reader; then straight into the C++ program compiler."""
import ctypes as C
import os

import numpy as np
import pytest

from kaldi_b200 import nnet_model as NM


def _lib():
    try:
        from kaldi_b200 import _lib as L
        return L.lib()
    except Exception as e:
        pytest.skip(str(e))


def _read(L, path, is_mdl):
    from kaldi_b200.nnet_compile import _Layer, _Weight
    h = C.c_void_p()
    L.b2k_model_read.argtypes = [C.c_char_p, C.c_int32, C.c_void_p]
    rc = L.b2k_model_read(path.encode(), int(is_mdl), C.byref(h))
    if rc:
        L.b2k_last_error.restype = C.c_char_p
        raise RuntimeError(L.b2k_last_error().decode())
    info = (C.c_int32 * 8)()
    L.b2k_model_info.argtypes = [C.c_void_p, C.c_void_p]
    assert L.b2k_model_info(h, info) == 0
    L.b2k_model_layers.restype = C.POINTER(_Layer)
    L.b2k_model_layers.argtypes = [C.c_void_p]
    L.b2k_model_weights.restype = C.POINTER(_Weight)
    L.b2k_model_weights.argtypes = [C.c_void_p]
    L.b2k_model_tid2pdf.restype = C.POINTER(C.c_int32)
    L.b2k_model_tid2pdf.argtypes = [C.c_void_p]
    layers = [L.b2k_model_layers(h)[i] for i in range(info[4])]
    W = {}
    for i in range(info[5]):
        w = L.b2k_model_weights(h)[i]
        a = np.ctypeslib.as_array(C.cast(w.data, C.POINTER(C.c_float)), shape=(w.size,)).copy()
        W[w.name.decode()] = a.reshape(w.rows, w.cols) if w.cols > 1 else a
    t2p = np.ctypeslib.as_array(L.b2k_model_tid2pdf(h), shape=(info[6],)).copy() if info[6] else None
    return h, list(info), layers, W, t2p


def _layer_dict(x):
    d = {f[0]: getattr(x, f[0]) for f in x._fields_}
    d["type"], d["name"], d["side"] = d["type"].decode(), d["name"].decode(), d["side"].decode()
    d["time_offsets"] = list(d["time_offsets"])[:d["n_time_offsets"]]
    d["height_offsets"] = list(d["height_offsets"])[:d["n_height_offsets"]]
    return d


CHAIN_TOPO = """<Topology>
<TopologyEntry>
<ForPhones> 1 2 3 4 5 </ForPhones>
<State> 0 <ForwardPdfClass> 0 <SelfLoopPdfClass> 1 <Transition> 0 0.5 <Transition> 1 0.5 </State>
<State> 1 </State>
</TopologyEntry>
</Topology>
"""


@pytest.mark.parametrize("which", ["idct-delta", "lda", "cnn", "tdnn"])
@pytest.mark.parametrize("binary", [1, 0])
def test_cpp_reader_on_reference_written_models(tmp_path, which, binary):
    L = _lib()
    from kaldi_b200.nnet_compile import _layer, CompiledProgram, _Cfg
    from oracle import nnet_oracle as NO
    if not (os.path.exists(NO._SO) or os.path.isdir("/root/reference")):
        pytest.skip("oracle/_ref nnet3 library not present")
    arch = NM.arch_tiny_cnn() if which == "cnn" else NM.arch_tiny_tdnn() if which == "tdnn" else NM.arch_tiny(front=which)
    Wt = NM.random_weights(arch, seed=3)
    R = NO.RefNnet(arch, Wt, collapse=False)
    if not hasattr(R.lib, "ref_write_final_mdl"):
        pytest.skip("oracle/_ref library predates the writers")
    R.lib.ref_nnet_write.argtypes = [C.c_void_p, C.c_char_p, C.c_int]
    R.lib.ref_write_final_mdl.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.c_char_p, C.c_int, C.c_int,
                                          C.c_void_p, C.c_int, C.c_void_p, C.c_int]
    raw, mdl = str(tmp_path / "final.raw"), str(tmp_path / "final.mdl")
    assert R.lib.ref_nnet_write(R.h, raw.encode(), binary) == 0
    pri = np.ascontiguousarray(Wt["priors"], np.float32)
    tid2pdf = np.zeros(64, np.int32)
    n_tids = R.lib.ref_write_final_mdl(R.h, mdl.encode(), binary, CHAIN_TOPO.encode(), 5, 2, pri.ctypes.data, pri.size,
                                       tid2pdf.ctypes.data, tid2pdf.size)
    assert n_tids > 0
    for path, is_mdl in ((raw, 0), (mdl, 1)):
        h, info, layers, W, t2p = _read(L, path, is_mdl)
        try:
            assert info[:4] == [arch["feat_dim"], arch["ivector_dim"], arch["num_pdfs"], arch["frame_subsampling_factor"]]
            # +-3 splices without a stride-3 TDNN-F layer do not decide --frame-subsampling-factor: flagged, and settled by the caller
            L.b2k_model_frame_subsampling_ambiguous.argtypes = [C.c_void_p]
            L.b2k_model_set_frame_subsampling_factor.argtypes = [C.c_void_p, C.c_int32]
            assert L.b2k_model_frame_subsampling_ambiguous(h) == (1 if which == "tdnn" else 0)
            if which == "tdnn":
                assert L.b2k_model_set_frame_subsampling_factor(h, 1) == 0 and L.b2k_model_frame_subsampling_ambiguous(h) == 0
                i2 = (C.c_int32 * 8)(); L.b2k_model_info(h, i2); assert i2[3] == 1
                assert L.b2k_model_set_frame_subsampling_factor(h, 3) == 0
            elif any(x["type"] == "tdnnf" and x.get("stride") == 3 for x in arch["layers"]):
                assert L.b2k_model_set_frame_subsampling_factor(h, 1) != 0           # contradicts the stride-3 TDNN-F layers
            want = [_layer_dict(_layer(x)) for x in arch["layers"]]
            got = [_layer_dict(x) for x in layers]
            assert len(got) == len(want)
            for a, b in zip(got, want):
                for k in b:
                    if isinstance(b[k], float):
                        assert abs(a[k] - b[k]) < 1e-6, (b["name"], k, a[k], b[k])
                    else:
                        assert a[k] == b[k], (b["name"], k, a[k], b[k])
            for k, v in Wt.items():
                if k == "priors" and not is_mdl:
                    np.testing.assert_array_equal(W[k], np.ones_like(v))
                    continue
                tol = dict(rtol=0, atol=0) if (binary and not k.endswith((".mean", ".var"))) else dict(rtol=2e-5, atol=1e-6)
                np.testing.assert_allclose(W[k].reshape(v.shape), v, err_msg=k, **tol)
            if is_mdl:
                np.testing.assert_array_equal(t2p, tid2pdf[:n_tids + 1])
            else:
                assert t2p is None
            if binary and is_mdl:
                # straight into the C++ compiler: same program as compiling the original arch/weights
                cfg = _Cfg(info[0], info[1], info[2], info[3], 120, 21, 1, 0, 1.0)
                prog = C.c_void_p()
                L.b2k_nnet_compile.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p]
                L.b2k_model_layers.restype = C.c_void_p
                L.b2k_model_weights.restype = C.c_void_p
                assert L.b2k_nnet_compile(C.byref(cfg), L.b2k_model_layers(h), info[4], L.b2k_model_weights(h), info[5], C.byref(prog)) == 0
                nn, no, bl = C.c_int32(), C.c_int32(), C.c_int64()
                L.b2k_nnet_program_sizes.argtypes = [C.c_void_p] * 4
                assert L.b2k_nnet_program_sizes(prog, C.byref(nn), C.byref(no), C.byref(bl)) == 0
                ref = CompiledProgram(arch, Wt, 120, 21)
                assert (nn.value, no.value, bl.value) == (len(ref.nodes), len(ref.ops), ref.blob.size)
                L.b2k_nnet_program_blob.restype = C.POINTER(C.c_float)
                L.b2k_nnet_program_blob.argtypes = [C.c_void_p]
                blob = np.ctypeslib.as_array(L.b2k_nnet_program_blob(prog), shape=(bl.value,))
                np.testing.assert_allclose(blob, ref.blob, rtol=2e-6, atol=1e-7)
                L.b2k_nnet_program_destroy.argtypes = [C.c_void_p]
                L.b2k_nnet_program_destroy(prog)
        finally:
            L.b2k_model_destroy.argtypes = [C.c_void_p]
            L.b2k_model_destroy(h)


def test_cpp_reader_rejects_garbage(tmp_path):
    L = _lib()
    p = str(tmp_path / "junk")
    open(p, "wb").write(b"\0B<Nnet3> \nnot a model\n\n")
    h = C.c_void_p()
    L.b2k_model_read.argtypes = [C.c_char_p, C.c_int32, C.c_void_p]
    assert L.b2k_model_read(p.encode(), 0, C.byref(h)) != 0
    assert L.b2k_model_read(str(tmp_path / "missing").encode(), 0, C.byref(h)) != 0


@pytest.mark.parametrize("which", ["idct-delta", "lda", "cnn"])
@pytest.mark.parametrize("binary", [1, 0])
def test_trained_recipe_model_with_dropout_and_xent_branch(tmp_path, binary, which):
    """What final.mdl of a chain recipe really contains besides the layers: GeneralDropoutComponent nodes (identity in test
    mode) after every batchnorm, and the cross-entropy branch (prefinal-xent / output-xent) next to the chain output.
    Both readers must see the same network as without them, and the program compiled from what the C++ reader returns
    must reproduce the reference's forward of that very model."""
    L = _lib()
    from kaldi_b200 import kaldi_io as KIO
    from kaldi_b200.model import KaldiModel
    from kaldi_b200.nnet import _Node, _Op
    from oracle import nnet_oracle as NO
    from oracle import program_interp as PI
    if not (os.path.exists(NO._SO) or os.path.isdir("/root/reference")):
        pytest.skip("oracle/_ref nnet3 library not present")
    plain = NM.arch_tiny_cnn() if which == "cnn" else NM.arch_tiny(64, front=which)
    arch = dict(plain, recipe_extras=True)
    Wt = NM.random_weights(plain, seed=5)
    R = NO.RefNnet(arch, Wt, collapse=False)
    if not hasattr(R.lib, "ref_nnet_write"):
        pytest.skip("oracle/_ref library predates the writers")
    R.lib.ref_nnet_write.argtypes = [C.c_void_p, C.c_char_p, C.c_int]
    raw = str(tmp_path / "final.raw")
    assert R.lib.ref_nnet_write(R.h, raw.encode(), binary) == 0
    text = open(raw, "rb").read()
    assert b"GeneralDropoutComponent" in text and b"output-xent" in text
    assert (b"SpecAugmentTimeMaskComponent" in text) == (which != "lda")
    # Python reader
    arch2, W2 = NM.load_kaldi_raw(raw)
    assert [(x["type"], x["name"]) for x in arch2["layers"]] == [(x["type"], x["name"]) for x in plain["layers"]]
    assert not any("xent" in k for k in W2)
    # C++ reader -> C++ compiler -> numpy interpreter of the ABI program == the reference's forward of this model
    m = KaldiModel(raw, is_mdl=False)
    assert [t for t, _ in m.layer_types()] == [x["type"] for x in plain["layers"]]
    got = m.weights()
    for k, v in Wt.items():
        if k == "priors":
            continue
        tol = dict(rtol=0, atol=0) if (binary and not k.endswith((".mean", ".var"))) else dict(rtol=2e-5, atol=1e-6)
        np.testing.assert_allclose(got[k].reshape(v.shape), v, err_msg=k, **tol)
    T = 60
    rng = np.random.default_rng(0)
    feats = (rng.standard_normal((T, plain["feat_dim"])) * 10).astype(np.float32)
    iv = rng.standard_normal((T, plain["ivector_dim"])).astype(np.float32)
    Rp = NO.RefNnet(arch, Wt, use_priors=False)
    ref = Rp.forward(feats, iv, period=1)
    rows = Rp.chunk_ivector_rows(T, T, 1)
    prog = m.compile(num_frames=T, frames_per_chunk=21, use_priors=False)
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
        out = PI.run_program(PI.program_from_abi(nodes, ops, blob), feats, iv[rows])
    finally:
        L.b2k_nnet_program_destroy.argtypes = [C.c_void_p]
        L.b2k_nnet_program_destroy(prog)
    assert out.shape == ref.shape
    assert np.abs(out - ref).max() <= 1e-4 * np.abs(ref).max()


@pytest.mark.parametrize("which", ["idct-delta", "cnn"])
def test_model_from_memory_as_the_nnet3_shims_serialise_it(tmp_path, which):
    """b2k_model_read_memory on the bytes Nnet::Write / AmNnetSimple::Write / final.mdl produce (kaldi_b200/host/
    b2k_nnet3_shims.h: ModelB2k writes the reference's object into a string stream): the same layers and weights as the file
    reader, with and without the binary marker in front of a text stream, and AmNnetSimple alone (kind 2) keeps its priors."""
    L = _lib()
    from oracle import nnet_oracle as NO
    if not (os.path.exists(NO._SO) or os.path.isdir("/root/reference")):
        pytest.skip("oracle/_ref nnet3 library not present")
    arch = NM.arch_tiny_cnn() if which == "cnn" else NM.arch_tiny(front=which)
    Wt = NM.random_weights(arch, seed=5)
    R = NO.RefNnet(arch, Wt, collapse=False)
    if not hasattr(R.lib, "ref_write_am_nnet"):
        pytest.skip("oracle/_ref library predates ref_write_am_nnet")
    R.lib.ref_nnet_write.argtypes = [C.c_void_p, C.c_char_p, C.c_int]
    R.lib.ref_write_am_nnet.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.c_int, C.c_void_p, C.c_int]
    R.lib.ref_write_final_mdl.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.c_char_p, C.c_int, C.c_int,
                                          C.c_void_p, C.c_int, C.c_void_p, C.c_int]
    pri = np.ascontiguousarray(Wt["priors"], np.float32)
    L.b2k_model_read_memory.argtypes = [C.c_char_p, C.c_int64, C.c_int32, C.c_void_p]
    L.b2k_model_destroy.argtypes = [C.c_void_p]

    def from_memory(data, kind):
        h = C.c_void_p()
        rc = L.b2k_model_read_memory(data, len(data), kind, C.byref(h))
        assert rc == 0, L.b2k_last_error()
        try:
            info = (C.c_int32 * 8)()
            L.b2k_model_info.argtypes = [C.c_void_p, C.c_void_p]
            assert L.b2k_model_info(h, info) == 0
            L.b2k_model_weights.restype = C.c_void_p
            from kaldi_b200.nnet_compile import _Weight
            ws = C.cast(L.b2k_model_weights(h), C.POINTER(_Weight))
            W = {ws[i].name.decode(): np.ctypeslib.as_array(C.cast(ws[i].data, C.POINTER(C.c_float)), shape=(ws[i].size,)).copy()
                 for i in range(info[5])}
            return list(info), W
        finally:
            L.b2k_model_destroy(h)

    for binary in (1, 0):
        raw, am, am_nohdr, mdl = (str(tmp_path / f"{n}{binary}") for n in ("raw", "am", "am_nohdr", "mdl"))
        assert R.lib.ref_nnet_write(R.h, raw.encode(), binary) == 0
        assert R.lib.ref_write_am_nnet(R.h, am.encode(), binary, 1, pri.ctypes.data, pri.size) == 0
        assert R.lib.ref_write_am_nnet(R.h, am_nohdr.encode(), 0, 0, pri.ctypes.data, pri.size) == 0
        t2p = np.zeros(64, np.int32)
        assert R.lib.ref_write_final_mdl(R.h, mdl.encode(), binary, CHAIN_TOPO.encode(), 5, 2, pri.ctypes.data, pri.size,
                                         t2p.ctypes.data, t2p.size) > 0
        _h, finfo, _layers, fW, _t = _read(L, mdl, 1)
        L.b2k_model_destroy(_h)
        for path, kind, has_priors in ((raw, 0, False), (am, 2, True), (am_nohdr, 2, True), (mdl, 1, True)):
            info, W = from_memory(open(path, "rb").read(), kind)
            assert info[:6] == finfo[:6] and info[7] == (1 if has_priors else 0)
            for k, v in fW.items():
                if k == "priors" and not has_priors:
                    continue
                tol = dict(rtol=0, atol=0) if (binary and path != am_nohdr) else dict(rtol=2e-5, atol=1e-6)
                np.testing.assert_allclose(W[k], np.asarray(v).ravel(), err_msg=f"{path} {k}", **tol)
    assert L.b2k_model_read_memory(b"\0Bgarbage", 9, 2, C.byref(C.c_void_p())) != 0
    assert L.b2k_model_read_memory(None, 0, 2, C.byref(C.c_void_p())) != 0


@pytest.mark.parametrize("which", ["lda", "tdnn"])
def test_collapsed_network_is_refused_not_misread(tmp_path, which):
    """nnet3::CollapseModel merges the LDA front into the first affine ("lda.tdnn1.affine"): such a network is not the model as
    trained, and both readers say so instead of matching their patterns against it (the tools built with the drop-in headers leave
    CollapseModel out: kaldi_b200/host/b2k_dropin_common.h).  The same model uncollapsed reads fine."""
    L = _lib()
    from oracle import nnet_oracle as NO
    if not (os.path.exists(NO._SO) or os.path.isdir("/root/reference")):
        pytest.skip("oracle/_ref nnet3 library not present")
    arch = NM.arch_tiny_tdnn() if which == "tdnn" else NM.arch_tiny(front=which)
    Wt = NM.random_weights(arch, seed=3)
    for collapse in (False, True):
        R = NO.RefNnet(arch, Wt, collapse=collapse)
        R.lib.ref_nnet_write.argtypes = [C.c_void_p, C.c_char_p, C.c_int]
        p = str(tmp_path / ("c.raw" if collapse else "p.raw"))
        assert R.lib.ref_nnet_write(R.h, p.encode(), 1) == 0
        h = C.c_void_p()
        L.b2k_model_read.argtypes = [C.c_char_p, C.c_int32, C.c_void_p]
        rc = L.b2k_model_read(p.encode(), 0, C.byref(h))
        L.b2k_last_error.restype = C.c_char_p
        if collapse:
            assert rc != 0 and "CollapseModel" in L.b2k_last_error().decode()
            from kaldi_b200.kaldi_io import KaldiFormatError
            with pytest.raises(KaldiFormatError, match="CollapseModel"):
                NM.load_kaldi_raw(p, frame_subsampling_factor=arch["frame_subsampling_factor"])
        else:
            assert rc == 0, L.b2k_last_error().decode()
            L.b2k_model_destroy.argtypes = [C.c_void_p]
            L.b2k_model_destroy(h)


def test_network_as_the_looped_info_leaves_it(tmp_path):
    """DecodableNnetSimpleLoopedInfo rewrites the i-vector term of the network it is given (nnet-compile-looped.cc:36-75:
    ReplaceIndex(ivector, t, 0) -> Round(ivector, period)); the looped decodable shim serialises THAT network
    (b2k_nnet3_shims.h: ModelB2k(info.nnet)).  Both readers take it and see the same layers as before the rewrite."""
    import re
    L = _lib()
    from oracle import nnet_oracle as NO
    if not (os.path.exists(NO._SO) or os.path.isdir("/root/reference")):
        pytest.skip("oracle/_ref nnet3 library not present")
    arch = NM.arch_tiny(front="lda")
    R = NO.RefNnet(arch, NM.random_weights(arch, seed=3), collapse=False)
    R.lib.ref_nnet_write.argtypes = [C.c_void_p, C.c_char_p, C.c_int]
    p, q = str(tmp_path / "a.raw"), str(tmp_path / "b.raw")
    assert R.lib.ref_nnet_write(R.h, p.encode(), 0) == 0
    txt = open(p, errors="replace").read()
    assert "ReplaceIndex(ivector, t, 0)" in txt
    open(q, "w").write(re.sub(r"ReplaceIndex\(ivector, t, 0\)", "Round(ivector, 30)", txt))
    ha, ia, la, _, _ = _read(L, p, 0)
    hb, ib, lb, _, _ = _read(L, q, 0)
    assert ia == ib and [_layer_dict(x) for x in la] == [_layer_dict(x) for x in lb]
    L.b2k_model_destroy.argtypes = [C.c_void_p]
    L.b2k_model_destroy(ha); L.b2k_model_destroy(hb)
    a1, _ = NM.load_kaldi_raw(p, frame_subsampling_factor=3)
    a2, _ = NM.load_kaldi_raw(q, frame_subsampling_factor=3)
    assert a1["layers"] == a2["layers"]
