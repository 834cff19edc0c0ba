"""kaldi_b200/kaldi_io.py against files written by the reference's OWN writers
(oracle/_ref: Nnet::Write, Matrix/Vector::Write): binary and text mode must give
back exactly the parameters that were put into the reference model."""
import ctypes as C
import os

import numpy as np
import pytest

from kaldi_b200 import kaldi_io as KIO
from kaldi_b200 import nnet_model as NM


def _ref_model_files(tmp_path, arch, W):
    from oracle import nnet_oracle as NO
    if not (os.path.exists(NO._SO) or os.path.isdir("/root/reference")):
        pytest.skip("oracle/_ref nnet3 library not present")
    R = NO.RefNnet(arch, W, collapse=False)     # the model as trained (CollapseModel renames and merges components)
    if not hasattr(R.lib, "ref_nnet_write"):
        pytest.skip("oracle/_ref nnet3 library predates ref_nnet_write")
    R.lib.ref_nnet_write.argtypes = [C.c_void_p, C.c_char_p, C.c_int]
    out = {}
    for mode, binary in (("bin", 1), ("txt", 0)):
        p = str(tmp_path / f"model.{mode}")
        assert R.lib.ref_nnet_write(R.h, p.encode(), binary) == 0
        out[mode] = p
    return out


@pytest.mark.parametrize("front", ["idct-delta", "lda", "cnn", "tdnn"])
def test_nnet3_raw_model_round_trip_through_the_reference_writer(tmp_path, front):
    arch = NM.arch_tiny_cnn() if front == "cnn" else NM.arch_tiny_tdnn() if front == "tdnn" else NM.arch_tiny(front=front)
    W = NM.random_weights(arch, seed=3)
    files = _ref_model_files(tmp_path, arch, W)
    sub = arch["frame_subsampling_factor"] if front == "tdnn" else None      # +-3 splices alone do not decide it: stated by the caller
    a3, W3 = NM.load_kaldi_raw(files["bin"], frame_subsampling_factor=sub)
    assert a3["num_pdfs"] == arch["num_pdfs"] and W3["priors"].shape == (arch["num_pdfs"],)
    for mode, path in files.items():
        parsed = KIO.read_nnet3_raw(path)
        arch2, W2 = KIO.nnet3_to_arch(parsed, name=arch["name"], frame_subsampling_factor=sub)
        assert [(L["type"], L["name"]) for L in arch2["layers"]] == [(L["type"], L["name"]) for L in arch["layers"]]
        for L, L2 in zip(arch["layers"], arch2["layers"]):
            for k, v in L.items():
                if isinstance(v, float):
                    assert abs(L2[k] - v) < 1e-6, (L["name"], k)
                else:
                    assert L2.get(k) == v, (L["name"], k, L2.get(k), v)
        assert arch2["feat_dim"] == arch["feat_dim"] and arch2["ivector_dim"] == arch["ivector_dim"]
        assert arch2["num_pdfs"] == arch["num_pdfs"]
        assert arch2["frame_subsampling_factor"] == arch["frame_subsampling_factor"]
        for k, v in W.items():
            if k == "priors":
                continue
            assert k in W2, k
            if mode == "bin" and not k.endswith((".mean", ".var")):
                np.testing.assert_array_equal(W2[k], v, err_msg=k)
            elif mode == "bin":
                # BatchNormComponent keeps sum and sum-of-squares (x count) and Write() divides again
                # (nnet-normalize-component.cc:591-650): the stats come back within an ulp or two
                np.testing.assert_allclose(W2[k], v, rtol=1e-6, atol=1e-7, err_msg=k)
            else:   # text mode prints ~6 significant digits
                np.testing.assert_allclose(W2[k], v, rtol=2e-5, atol=1e-6, err_msg=k)


def test_matrix_and_vector_files(tmp_path):
    rng = np.random.default_rng(0)
    m = rng.standard_normal((7, 5)).astype(np.float32)
    v = rng.standard_normal(9).astype(np.float32)
    for binary in (True, False):
        KIO.write_matrix(str(tmp_path / "m"), m, binary)
        KIO.write_vector(str(tmp_path / "v"), v, binary)
        np.testing.assert_array_equal(KIO.read_matrix(str(tmp_path / "m")), m)
        np.testing.assert_array_equal(KIO.read_vector(str(tmp_path / "v")), v)
    md = rng.standard_normal((3, 4))
    KIO.write_matrix(str(tmp_path / "md"), md, True)
    np.testing.assert_array_equal(KIO.read_matrix(str(tmp_path / "md")), md)


def test_ivector_extractor_and_ubm_files_written_by_the_reference(tmp_path):
    from kaldi_b200.ivector import make_synthetic_extractor
    from oracle import ivector_oracle as IV
    from oracle import nnet_oracle as NO
    if not (os.path.exists(NO._SO) or os.path.isdir("/root/reference")):
        pytest.skip("oracle/_ref library not present")
    ex = make_synthetic_extractor(seed=1, num_gauss=6, feat_dim=8, ivector_dim=5, splice=1, base_dim=8)
    R = IV.RefIvector(ex)
    if not hasattr(R.lib, "ref_ivector_write"):
        pytest.skip("oracle/_ref library predates ref_ivector_write")
    R.lib.ref_ivector_write.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p, C.c_int]
    for binary in (1, 0):
        pe, pu = str(tmp_path / f"final{binary}.ie"), str(tmp_path / f"final{binary}.dubm")
        assert R.lib.ref_ivector_write(R.h, pe.encode(), pu.encode(), binary) == 0
        got = KIO.read_ivector_extractor(pe)
        ubm = KIO.read_diag_gmm(pu)
        tol = dict(rtol=0, atol=0) if binary else dict(rtol=2e-5, atol=1e-6)
        np.testing.assert_allclose(got["M"], ex["M"], **tol)
        np.testing.assert_allclose(got["sigma_inv"], ex["sigma_inv"], **tol)
        assert abs(got["prior_offset"] - ex["prior_offset"]) < 1e-6
        np.testing.assert_allclose(got["sigma_inv_m"], ex["sigma_inv_m"], rtol=1e-4 if not binary else 1e-12, atol=1e-6 if not binary else 1e-12)
        np.testing.assert_allclose(got["U"], ex["U"], rtol=1e-4 if not binary else 1e-12, atol=1e-5 if not binary else 1e-12)
        for k in ("gconsts", "ubm_weights", "means_invvars", "inv_vars"):
            np.testing.assert_allclose(ubm[k], ex[k], rtol=1e-5 if not binary else 1e-6, atol=1e-6, err_msg=k)


CHAIN_TOPO = """<Topology>
<TopologyEntry>
<ForPhones> 1 2 3 4 5 </ForPhones>
<State> 0 <ForwardPdfClass> 0 <SelfLoopPdfClass> 1 <Transition> 0 0.5 <Transition> 1 0.5 </State>
<State> 1 </State>
</TopologyEntry>
</Topology>
"""
HMM3_TOPO = """<Topology>
<TopologyEntry>
<ForPhones> 1 2 3 4 5 </ForPhones>
<State> 0 <PdfClass> 0 <Transition> 0 0.75 <Transition> 1 0.25 </State>
<State> 1 <PdfClass> 1 <Transition> 1 0.75 <Transition> 2 0.25 </State>
<State> 2 <PdfClass> 2 <Transition> 2 0.75 <Transition> 3 0.25 </State>
<State> 3 </State>
</TopologyEntry>
</Topology>
"""


@pytest.mark.parametrize("topo,classes", [(CHAIN_TOPO, 2), (HMM3_TOPO, 3)], ids=["chain", "hmm3"])
def test_final_mdl_written_by_the_reference(tmp_path, topo, classes):
    from oracle import nnet_oracle as NO
    if not (os.path.exists(NO._SO) or os.path.isdir("/root/reference")):
        pytest.skip("oracle/_ref nnet3 library not present")
    arch = NM.arch_tiny()
    W = NM.random_weights(arch, seed=5)
    R = NO.RefNnet(arch, W, collapse=False)
    if not hasattr(R.lib, "ref_write_final_mdl"):
        pytest.skip("oracle/_ref nnet3 library predates ref_write_final_mdl")
    R.lib.ref_write_final_mdl.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.c_char_p, C.c_int, C.c_int,
                                          C.c_void_p, C.c_int, C.c_void_p, C.c_int]
    pri = np.ascontiguousarray(W["priors"], np.float32)
    for binary in (1, 0):
        path = str(tmp_path / f"final{binary}.mdl")
        tid2pdf = np.zeros(256, np.int32)
        n = R.lib.ref_write_final_mdl(R.h, path.encode(), binary, topo.encode(), 5, classes, pri.ctypes.data, pri.size,
                                      tid2pdf.ctypes.data, tid2pdf.size)
        assert n > 0
        m = KIO.read_final_mdl(path)
        tm = m["transition_model"]
        np.testing.assert_array_equal(tm["tid2pdf"], tid2pdf[:n + 1])      # TransitionIdToPdf of the reference
        assert tm["num_pdfs"] == 5 * classes
        np.testing.assert_allclose(m["priors"], pri, rtol=0 if binary else 2e-5, atol=0 if binary else 1e-7)
        arch2, W2 = KIO.nnet3_to_arch(m["nnet"])
        assert [(L["type"], L["name"]) for L in arch2["layers"]] == [(L["type"], L["name"]) for L in arch["layers"]]
        np.testing.assert_allclose(W2["output.affine.w"], W["output.affine.w"], rtol=0 if binary else 2e-5, atol=0 if binary else 1e-6)


@pytest.mark.parametrize("fst_type,aligned", [("const", False), ("const", True), ("vector", False)])
def test_openfst_binary_layout_self_consistency(tmp_path, fst_type, aligned):
    """PARITY UNPINNED (no OpenFst in this image): reader and writer agree with each other on the published
    layout; the arrays come back bit for bit, arc order preserved."""
    from kaldi_b200 import synth
    g = synth.make_hclg(20_000, num_pdfs=50, seed=3)
    p = str(tmp_path / "HCLG.fst")
    KIO.write_openfst(p, g, fst_type, aligned)
    h = KIO.read_openfst(p)
    assert h["num_states"] == g["num_states"] and h["start"] == g["start"] and h["fst_type"] == fst_type
    for k in ("offsets", "ilabel", "olabel", "nextstate"):
        np.testing.assert_array_equal(h[k], np.asarray(g[k], np.int32)[:len(h[k])], err_msg=k)
    np.testing.assert_array_equal(h["weight"].view(np.int32), np.asarray(g["weight"], np.float32).view(np.int32))
    np.testing.assert_array_equal(h["final"].view(np.int32), np.asarray(g["final"], np.float32).view(np.int32))


def test_written_matrices_are_read_by_the_reference(tmp_path):
    """The other direction: what kaldi_io writes, the reference's Matrix/Vector::Read must accept, bit for bit."""
    from oracle import nnet_oracle as NO
    if not (os.path.exists(NO._SO) or os.path.isdir("/root/reference")):
        pytest.skip("oracle/_ref library not present")
    L = C.CDLL(NO._SO)
    if not hasattr(L, "ref_read_matrix"):
        pytest.skip("oracle/_ref library predates ref_read_matrix")
    L.ref_read_matrix.argtypes = [C.c_char_p, C.c_void_p, C.c_int, C.c_void_p]
    L.ref_read_vector.argtypes = [C.c_char_p, C.c_void_p, C.c_int]
    rng = np.random.default_rng(4)
    m = rng.standard_normal((6, 9)).astype(np.float32)
    v = rng.standard_normal(11).astype(np.float32)
    for binary in (True, False):
        pm, pv = str(tmp_path / f"m{int(binary)}"), str(tmp_path / f"v{int(binary)}")
        KIO.write_matrix(pm, m, binary)
        KIO.write_vector(pv, v, binary)
        out = np.zeros(m.size, np.float32)
        cols = C.c_int()
        rows = L.ref_read_matrix(pm.encode(), out.ctypes.data, out.size, C.addressof(cols))
        assert (rows, cols.value) == m.shape
        np.testing.assert_array_equal(out.reshape(m.shape), m)
        outv = np.zeros(v.size, np.float32)
        assert L.ref_read_vector(pv.encode(), outv.ctypes.data, outv.size) == v.size
        np.testing.assert_array_equal(outv, v)
