"""The C++ readers of the i-vector extractor directory (kaldi_b200/csrc/model_io.cu: b2k_ivec_files_*) on final.ie /
final.dubm written by the reference's own IvectorExtractor::Write / DiagGmm::Write (oracle/_ref), final.mat and
global_cmvn.stats in Kaldi matrix format, against the extractor that was put in and against the Python readers."""
import ctypes as C
import os

import numpy as np
import pytest


@pytest.mark.parametrize("binary", [1, 0])
def test_cpp_reader_on_reference_written_extractor(tmp_path, binary):
    from kaldi_b200 import kaldi_io as KIO
    from kaldi_b200.ivector import IvectorFiles, make_synthetic_extractor
    from oracle import ivector_oracle as IV
    from oracle import nnet_oracle as NO
    if not (os.path.exists(NO._SO) or os.path.isdir("/root/reference")):
        pytest.skip("oracle/_ref library not present")
    ex = make_synthetic_extractor(seed=1, num_gauss=6, feat_dim=8, ivector_dim=5, splice=1, base_dim=8)
    R = IV.RefIvector(ex)
    if not hasattr(R.lib, "ref_ivector_write"):
        pytest.skip("oracle/_ref library predates ref_ivector_write")
    R.lib.ref_ivector_write.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p, C.c_int]
    pe, pu = str(tmp_path / "final.ie"), str(tmp_path / "final.dubm")
    pm, pc = str(tmp_path / "final.mat"), str(tmp_path / "global_cmvn.stats")
    assert R.lib.ref_ivector_write(R.h, pe.encode(), pu.encode(), binary) == 0
    KIO.write_matrix(pm, ex["lda_mat"], binary=bool(binary))
    KIO.write_matrix(pc, np.asarray(ex["global_cmvn_stats"], np.float64), binary=bool(binary))
    try:
        f = IvectorFiles(pe, pu, pm, pc)
    except OSError as e:
        pytest.skip(str(e))
    assert (f.num_gauss, f.feat_dim, f.ivector_dim, f.lda_rows, f.lda_cols, f.cmvn_dim) == (6, 8, 5, 8, 8 * 3 + 1, 8)
    assert abs(f.prior_offset - ex["prior_offset"]) < 1e-6
    a = f.arrays()
    exact = dict(rtol=0, atol=0)
    loose = dict(rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(a["sigma_inv_m"], ex["sigma_inv_m"], **(dict(rtol=1e-12, atol=1e-12) if binary else loose))
    np.testing.assert_allclose(a["U"], ex["U"], **(dict(rtol=1e-12, atol=1e-12) if binary else loose))
    for k in ("gconsts", "means_invvars", "inv_vars", "ubm_weights"):
        np.testing.assert_allclose(a[k], ex[k], rtol=1e-6 if binary else 1e-5, atol=1e-6, err_msg=k)
    np.testing.assert_allclose(a["lda_mat"], ex["lda_mat"], **(exact if binary else dict(rtol=1e-6, atol=1e-7)))
    np.testing.assert_allclose(a["global_cmvn_stats"], ex["global_cmvn_stats"], **(exact if binary else dict(rtol=1e-6, atol=1e-6)))
    # the Python readers see the same numbers
    ie, ubm = KIO.read_ivector_extractor(pe), KIO.read_diag_gmm(pu)
    np.testing.assert_allclose(a["sigma_inv_m"], ie["sigma_inv_m"], rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(a["U"], ie["U"], rtol=1e-12, atol=1e-12)
    for k in ("gconsts", "means_invvars", "inv_vars"):
        np.testing.assert_array_equal(a[k], ubm[k])
    np.testing.assert_array_equal(a["lda_mat"], KIO.read_matrix(pm).astype(np.float32))


def test_cpp_reader_checks_consistency(tmp_path):
    from kaldi_b200 import kaldi_io as KIO
    from kaldi_b200.ivector import IvectorFiles
    p = str(tmp_path / "junk")
    open(p, "wb").write(b"\0B<Nope> ")
    with pytest.raises((RuntimeError, OSError)):
        IvectorFiles(p, p, p, p)
    with pytest.raises((RuntimeError, OSError)):
        IvectorFiles(str(tmp_path / "missing"), p, p, p)


def test_device_creation_fails_loudly_without_a_gpu(tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from kaldi_b200 import kaldi_io as KIO
    from kaldi_b200.ivector import IvectorExtractorGpu, IvectorFiles, make_synthetic_extractor
    from oracle import ivector_oracle as IV
    from oracle import nnet_oracle as NO
    if not (os.path.exists(NO._SO) or os.path.isdir("/root/reference")):
        pytest.skip("oracle/_ref library not present")
    ex = make_synthetic_extractor(seed=1, num_gauss=6, feat_dim=8, ivector_dim=5, splice=1, base_dim=8)
    R = IV.RefIvector(ex)
    R.lib.ref_ivector_write.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p, C.c_int]
    pe, pu = str(tmp_path / "final.ie"), str(tmp_path / "final.dubm")
    pm, pc = str(tmp_path / "final.mat"), str(tmp_path / "global_cmvn.stats")
    assert R.lib.ref_ivector_write(R.h, pe.encode(), pu.encode(), 1) == 0
    KIO.write_matrix(pm, ex["lda_mat"])
    KIO.write_matrix(pc, np.asarray(ex["global_cmvn_stats"], np.float64))
    f = IvectorFiles(pe, pu, pm, pc)
    with pytest.raises(RuntimeError):                      # wrong splice for this final.mat
        IvectorExtractorGpu.from_files(f, 2, 100, splice=3, base_dim=8)
    with pytest.raises(RuntimeError):                      # right shapes, no device
        IvectorExtractorGpu.from_files(f, 2, 100, splice=1, base_dim=8)
