"""A network written the way the recipes write it — the reference's OWN xconfig library (egs/wsj/s5/steps/libs/nnet3/xconfig,
imported in the build container) turns a network.xconfig shaped like run_tdnn_1k.sh / run_tdnn_1d.sh into nnet3 config lines, the
reference's Nnet::ReadConfig builds it, Nnet::Write stores it — must load through both model readers (dropout components, the
cross-entropy branch and all) and the compiled program must reproduce the reference's looped forward of that very model."""
import ctypes as C
import os
import sys
import warnings

import numpy as np
import pytest

from kaldi_b200 import kaldi_io as KIO, nnet_model as NM

STEPS = "/root/reference/egs/wsj/s5/steps"

def _randomise_parameters(L, h, rng):
    """Kaldi initialises output layers to zero: give every updatable component seeded random parameters so that the forward
    comparison is not 0 == 0."""
    for i in range(L.ref_nnet_num_components(h)):
        n = L.ref_nnet_num_params(h, i)
        if n > 0:
            v = (rng.standard_normal(n) * 0.1).astype(np.float32)
            assert L.ref_nnet_set_params(h, i, v.ctypes.data_as(C.POINTER(C.c_float))) == 0


XCONFIG = """input dim=100 name=ivector
input dim=40 name=input
fixed-affine-layer name=lda input=Append(-1,0,1,ReplaceIndex(ivector, t, 0)) affine-transform-file={lda}
relu-batchnorm-dropout-layer name=tdnn1 l2-regularize=0.01 dropout-proportion=0.0 dropout-per-dim-continuous=true dim=64
tdnnf-layer name=tdnnf2 l2-regularize=0.01 dropout-proportion=0.0 bypass-scale=0.66 dim=64 bottleneck-dim=16 time-stride=1
tdnnf-layer name=tdnnf3 l2-regularize=0.01 dropout-proportion=0.0 bypass-scale=0.66 dim=64 bottleneck-dim=16 time-stride=0
tdnnf-layer name=tdnnf4 l2-regularize=0.01 dropout-proportion=0.0 bypass-scale=0.66 dim=64 bottleneck-dim=16 time-stride=3
tdnnf-layer name=tdnnf5 l2-regularize=0.01 dropout-proportion=0.0 bypass-scale=0.66 dim=64 bottleneck-dim=16 time-stride=3
linear-component name=prefinal-l dim=32 l2-regularize=0.01 orthonormal-constraint=-1.0
prefinal-layer name=prefinal-chain input=prefinal-l l2-regularize=0.01 big-dim=64 small-dim=32
output-layer name=output include-log-softmax=false dim=48 l2-regularize=0.002
prefinal-layer name=prefinal-xent input=prefinal-l l2-regularize=0.01 big-dim=64 small-dim=32
output-layer name=output-xent dim=48 learning-rate-factor=5.0 l2-regularize=0.002
"""
BN_DIMS = {"tdnn1.batchnorm": 64, "tdnnf2.batchnorm": 64, "tdnnf3.batchnorm": 64, "tdnnf4.batchnorm": 64, "tdnnf5.batchnorm": 64,
           "prefinal-chain.batchnorm1": 64, "prefinal-chain.batchnorm2": 32, "prefinal-xent.batchnorm1": 64, "prefinal-xent.batchnorm2": 32}


@pytest.mark.parametrize("binary", [1, 0])
def test_model_built_by_the_references_xconfig_library(tmp_path, binary):
    if not os.path.isdir(STEPS):
        pytest.skip("the reference's xconfig library exists in the build container only")
    from oracle import nnet_oracle as NO
    from oracle import program_interp as PI
    sys.path.insert(0, STEPS)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        import libs.nnet3.xconfig.parser as xparser
    rng = np.random.default_rng(5)
    lda = str(tmp_path / "lda.mat")
    KIO.write_matrix(lda, (rng.standard_normal((220, 221)) / 15).astype(np.float32), binary=False)
    xc = str(tmp_path / "network.xconfig")
    open(xc, "w").write(XCONFIG.format(lda=lda))
    lines = [line for layer in xparser.read_xconfig_file(xc) for base, line in layer.get_full_config() if base == "final"]
    config = "\n".join(lines) + "\n"
    assert "GeneralDropoutComponent" in config and "output-xent" in config
    # the reference builds the network from those lines (random parameters of its own), we give the batchnorms statistics
    R = NO.RefNnet.__new__(NO.RefNnet)
    L = R.lib = C.CDLL(NO._SO)
    L.ref_nnet_create.restype = C.c_void_p
    L.ref_nnet_component_name.restype = C.c_char_p
    L.ref_nnet_component_type.restype = C.c_char_p
    R.h = C.c_void_p(L.ref_nnet_create(config.encode()))
    assert R.h, "the reference rejected its own xconfig output"
    R.arch = {"frame_subsampling_factor": 3}
    _randomise_parameters(L, R.h, rng)
    f32p = C.POINTER(C.c_float)
    for i in range(L.ref_nnet_num_components(R.h)):
        name, typ = L.ref_nnet_component_name(R.h, i).decode(), L.ref_nnet_component_type(R.h, i).decode()
        if typ == "BatchNormComponent":
            d = BN_DIMS[name]
            mean, var = (rng.standard_normal(d) * 0.1).astype(np.float32), rng.uniform(0.5, 1.5, d).astype(np.float32)
            assert L.ref_nnet_set_batchnorm(R.h, i, C.c_int(d), C.c_int(d), C.c_float(1e-3), C.c_float(1.0), C.c_float(1000.0),
                                            mean.ctypes.data_as(f32p), var.ctypes.data_as(f32p)) == 0, name
    raw = str(tmp_path / "final.raw")
    L.ref_nnet_write.argtypes = [C.c_void_p, C.c_char_p, C.c_int]
    assert L.ref_nnet_write(R.h, raw.encode(), binary) == 0
    # both readers see the recipe's layers and nothing of the dropout / xent machinery
    arch, W = NM.load_kaldi_raw(raw)
    want_layers = [("lda", "lda"), ("relu-batchnorm", "tdnn1"), ("tdnnf", "tdnnf2"), ("tdnnf", "tdnnf3"), ("tdnnf", "tdnnf4"), ("tdnnf", "tdnnf5"),
                   ("linear", "prefinal-l"), ("prefinal", "prefinal-chain"), ("output", "output")]
    assert [(x["type"], x["name"]) for x in arch["layers"]] == want_layers
    assert [x["stride"] for x in arch["layers"] if x["type"] == "tdnnf"] == [1, 0, 3, 3]
    assert all(abs(x["bypass"] - 0.66) < 1e-6 for x in arch["layers"] if x["type"] == "tdnnf")
    assert (arch["feat_dim"], arch["ivector_dim"], arch["num_pdfs"], arch["frame_subsampling_factor"]) == (40, 100, 48, 3)
    try:
        from kaldi_b200.model import KaldiModel
        m = KaldiModel(raw, is_mdl=False)
    except OSError as e:
        pytest.skip(str(e))
    assert m.layer_types() == want_layers
    got = m.weights()
    for k, v in W.items():
        np.testing.assert_array_equal(got[k].reshape(-1), np.asarray(v, np.float32).reshape(-1), err_msg=k)
    # forward: the reference's looped computation of ITS model against the program compiled from what we read
    pri = np.ones(48, np.float32)
    assert L.ref_nnet_prepare(R.h, C.c_int(20), C.c_int(3), C.c_float(1.0), None, C.c_int(0), C.c_int(1)) == 0
    info = (C.c_int * 4)()
    L.ref_nnet_info(R.h, info)
    R.left_context, R.right_context, R.frames_per_chunk, R.output_dim = list(info)
    T = 70
    feats = (rng.standard_normal((T, 40)) * 10).astype(np.float32)
    iv = rng.standard_normal((T, 100)).astype(np.float32)
    ref = R.forward(feats, iv, period=1)
    prog = NM.compile_program(arch, dict(W, priors=pri), T, 21, use_priors=False)
    out = PI.run_program(prog, feats, iv[R.chunk_ivector_rows(T, T, 1)])
    assert np.abs(ref).max() > 1e-2 and ref.std() > 1e-3                                  # a real comparison, not 0 == 0
    assert out.shape == ref.shape and np.abs(out - ref).max() <= 1e-4 * np.abs(ref).max()
    assert (R.left_context, R.right_context) == (prog["model_left"], prog["model_right"])


CNN_XCONFIG = """input dim=100 name=ivector
input dim=40 name=input
idct-layer name=idct input=input dim=40 cepstral-lifter=22 affine-transform-file={idct}
linear-component name=ivector-linear l2-regularize=0.01 dim=200 input=ReplaceIndex(ivector, t, 0)
batchnorm-component name=ivector-batchnorm target-rms=0.025
batchnorm-component name=idct-batchnorm input=idct
combine-feature-maps-layer name=combine_inputs input=Append(idct-batchnorm, ivector-batchnorm) num-filters1=1 num-filters2=5 height=40
conv-relu-batchnorm-layer name=cnn1 l2-regularize=0.01 height-in=40 height-out=40 time-offsets=-1,0,1 height-offsets=-1,0,1 num-filters-out=8
conv-relu-batchnorm-layer name=cnn2 l2-regularize=0.01 height-in=40 height-out=20 height-subsample-out=2 time-offsets=-1,0,1 height-offsets=-1,0,1 num-filters-out=16
conv-relu-batchnorm-layer name=cnn3 l2-regularize=0.01 height-in=20 height-out=10 height-subsample-out=2 time-offsets=-1,0,1 height-offsets=-1,0,1 num-filters-out=16
tdnnf-layer name=tdnnf4 l2-regularize=0.01 dropout-proportion=0.0 bypass-scale=0.0 dim=64 bottleneck-dim=32 time-stride=0
tdnnf-layer name=tdnnf5 l2-regularize=0.01 dropout-proportion=0.0 bypass-scale=0.66 dim=64 bottleneck-dim=16 time-stride=3
linear-component name=prefinal-l dim=32 l2-regularize=0.01 orthonormal-constraint=-1.0
prefinal-layer name=prefinal-chain input=prefinal-l l2-regularize=0.01 big-dim=64 small-dim=32
output-layer name=output include-log-softmax=false dim=48 l2-regularize=0.002
prefinal-layer name=prefinal-xent input=prefinal-l l2-regularize=0.01 big-dim=64 small-dim=32
output-layer name=output-xent dim=48 learning-rate-factor=5.0 l2-regularize=0.002
"""
CNN_BN = {"ivector-batchnorm": (200, 200), "idct-batchnorm": (40, 40), "cnn1.batchnorm": (320, 8), "cnn2.batchnorm": (320, 16), "cnn3.batchnorm": (160, 16),
          "tdnnf4.batchnorm": (64, 64), "tdnnf5.batchnorm": (64, 64), "prefinal-chain.batchnorm1": (64, 64), "prefinal-chain.batchnorm2": (32, 32),
          "prefinal-xent.batchnorm1": (64, 64), "prefinal-xent.batchnorm2": (32, 32)}


def test_cnn_tdnnf_built_by_the_references_xconfig_library(tmp_path):
    """The run_cnn_tdnn_1a.sh shape (idct, i-vector branch, combine-feature-maps, convolutions with height subsampling, a first
    TDNN-F without bypass) from the reference's xconfig library through the readers and the compiler."""
    if not os.path.isdir(STEPS):
        pytest.skip("the reference's xconfig library exists in the build container only")
    from oracle import nnet_oracle as NO
    from oracle import program_interp as PI
    sys.path.insert(0, STEPS)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        import libs.nnet3.xconfig.parser as xparser
    rng = np.random.default_rng(6)
    idct = str(tmp_path / "idct.mat")
    KIO.write_matrix(idct, np.concatenate([np.linalg.qr(rng.standard_normal((40, 40)))[0], np.zeros((40, 1))], 1).astype(np.float32), binary=False)
    xc = str(tmp_path / "network.xconfig")
    open(xc, "w").write(CNN_XCONFIG.format(idct=idct))
    lines = [line for layer in xparser.read_xconfig_file(xc) for base, line in layer.get_full_config() if base == "final"]
    R = NO.RefNnet.__new__(NO.RefNnet)
    L = R.lib = C.CDLL(NO._SO)
    L.ref_nnet_create.restype = C.c_void_p
    L.ref_nnet_component_name.restype = C.c_char_p
    L.ref_nnet_component_type.restype = C.c_char_p
    R.h = C.c_void_p(L.ref_nnet_create(("\n".join(lines) + "\n").encode()))
    assert R.h
    R.arch = {"frame_subsampling_factor": 3}
    _randomise_parameters(L, R.h, rng)
    f32p = C.POINTER(C.c_float)
    for i in range(L.ref_nnet_num_components(R.h)):
        name, typ = L.ref_nnet_component_name(R.h, i).decode(), L.ref_nnet_component_type(R.h, i).decode()
        if typ == "BatchNormComponent":
            d, bd = CNN_BN[name]
            mean, var = (rng.standard_normal(bd) * 0.1).astype(np.float32), rng.uniform(0.5, 1.5, bd).astype(np.float32)
            rms = 0.025 if name == "ivector-batchnorm" else 1.0
            assert L.ref_nnet_set_batchnorm(R.h, i, C.c_int(d), C.c_int(bd), C.c_float(1e-3), C.c_float(rms), C.c_float(1000.0),
                                            mean.ctypes.data_as(f32p), var.ctypes.data_as(f32p)) == 0, name
    raw = str(tmp_path / "final.raw")
    L.ref_nnet_write.argtypes = [C.c_void_p, C.c_char_p, C.c_int]
    assert L.ref_nnet_write(R.h, raw.encode(), 1) == 0
    arch, W = NM.load_kaldi_raw(raw)
    assert [(x["type"], x["name"]) for x in arch["layers"]] == [
        ("idct", "idct"), ("ivector-linear-bn", "ivector"), ("batchnorm", "idct-batchnorm"), ("combine", "combine_inputs"), ("conv", "cnn1"),
        ("conv", "cnn2"), ("conv", "cnn3"), ("tdnnf", "tdnnf4"), ("tdnnf", "tdnnf5"), ("linear", "prefinal-l"), ("prefinal", "prefinal-chain"),
        ("output", "output")]
    assert [x["bypass"] for x in arch["layers"] if x["type"] == "tdnnf"] == [0.0, pytest.approx(0.66)]
    try:
        from kaldi_b200.model import KaldiModel
        assert KaldiModel(raw, is_mdl=False).layer_types() == [(x["type"], x["name"]) for x in arch["layers"]]
    except OSError:
        pass
    assert L.ref_nnet_prepare(R.h, C.c_int(20), C.c_int(3), C.c_float(1.0), None, C.c_int(0), C.c_int(1)) == 0
    info = (C.c_int * 4)()
    L.ref_nnet_info(R.h, info)
    R.left_context, R.right_context, R.frames_per_chunk, R.output_dim = list(info)
    T = 50
    feats = (rng.standard_normal((T, 40)) * 10).astype(np.float32)
    iv = rng.standard_normal((T, 100)).astype(np.float32)
    ref = R.forward(feats, iv, period=1)
    for mode in ("patch", "dense"):
        prog = NM.compile_program(arch, dict(W, priors=np.ones(48, np.float32)), T, 21, use_priors=False, conv_mode=mode)
        out = PI.run_program(prog, feats, iv[R.chunk_ivector_rows(T, T, 1)])
        assert np.abs(ref).max() > 1e-2 and ref.std() > 1e-3
        assert out.shape == ref.shape and np.abs(out - ref).max() <= 1e-4 * np.abs(ref).max(), mode


TDNN_1K_XCONFIG = """input dim=100 name=ivector
input dim=40 name=input
idct-layer name=idct input=input dim=40 cepstral-lifter=22 affine-transform-file={idct}
batchnorm-component name=batchnorm0 input=idct
spec-augment-layer name=spec-augment freq-max-proportion=0.5 time-zeroed-proportion=0.2 time-mask-max-frames=20
delta-layer name=delta input=spec-augment
no-op-component name=input2 input=Append(delta, Scale(0.4, ReplaceIndex(ivector, t, 0)))
relu-batchnorm-layer name=tdnn1 l2-regularize=0.03 dim=64 input=input2
tdnnf-layer name=tdnnf2 l2-regularize=0.03 bypass-scale=0.66 dim=64 bottleneck-dim=16 time-stride=1
tdnnf-layer name=tdnnf3 l2-regularize=0.03 bypass-scale=0.66 dim=64 bottleneck-dim=16 time-stride=0
tdnnf-layer name=tdnnf4 l2-regularize=0.03 bypass-scale=0.66 dim=64 bottleneck-dim=16 time-stride=3
linear-component name=prefinal-l dim=32 l2-regularize=0.03 orthonormal-constraint=-1.0
prefinal-layer name=prefinal-chain input=prefinal-l l2-regularize=0.03 small-dim=32 big-dim=64
output-layer name=output include-log-softmax=false dim=48 l2-regularize=0.015
prefinal-layer name=prefinal-xent input=prefinal-l l2-regularize=0.03 small-dim=32 big-dim=64
output-layer name=output-xent dim=48 learning-rate-factor=5.0 l2-regularize=0.015
"""
TDNN_1K_BN = {"batchnorm0": 40, "delta": 120, "tdnn1.batchnorm": 64, "tdnnf2.batchnorm": 64, "tdnnf3.batchnorm": 64, "tdnnf4.batchnorm": 64,
              "prefinal-chain.batchnorm1": 64, "prefinal-chain.batchnorm2": 32, "prefinal-xent.batchnorm1": 64, "prefinal-xent.batchnorm2": 32}


def test_mini_librispeech_tdnn_1k_shape_from_the_references_xconfig_library(tmp_path):
    """The network of the headline configuration (egs/mini_librispeech/s5/local/chain/tuning/run_tdnn_1k.sh:170-207): idct,
    batchnorm, spec-augment, delta-layer, the no-op node that appends the scaled i-vector, relu-batchnorm, TDNN-F, both branches."""
    if not os.path.isdir(STEPS):
        pytest.skip("the reference's xconfig library exists in the build container only")
    from oracle import nnet_oracle as NO
    from oracle import program_interp as PI
    sys.path.insert(0, STEPS)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        import libs.nnet3.xconfig.parser as xparser
    rng = np.random.default_rng(7)
    idct = str(tmp_path / "idct.mat")
    KIO.write_matrix(idct, np.concatenate([np.linalg.qr(rng.standard_normal((40, 40)))[0], np.zeros((40, 1))], 1).astype(np.float32), binary=False)
    xc = str(tmp_path / "network.xconfig")
    open(xc, "w").write(TDNN_1K_XCONFIG.format(idct=idct))
    lines = [line for layer in xparser.read_xconfig_file(xc) for base, line in layer.get_full_config() if base == "final"]
    R = NO.RefNnet.__new__(NO.RefNnet)
    L = R.lib = C.CDLL(NO._SO)
    L.ref_nnet_create.restype = C.c_void_p
    L.ref_nnet_component_name.restype = C.c_char_p
    L.ref_nnet_component_type.restype = C.c_char_p
    R.h = C.c_void_p(L.ref_nnet_create(("\n".join(lines) + "\n").encode()))
    assert R.h
    R.arch = {"frame_subsampling_factor": 3}
    _randomise_parameters(L, R.h, rng)
    f32p = C.POINTER(C.c_float)
    for i in range(L.ref_nnet_num_components(R.h)):
        name, typ = L.ref_nnet_component_name(R.h, i).decode(), L.ref_nnet_component_type(R.h, i).decode()
        if typ == "BatchNormComponent":
            d = TDNN_1K_BN[name]
            mean, var = (rng.standard_normal(d) * 0.1).astype(np.float32), rng.uniform(0.5, 1.5, d).astype(np.float32)
            assert L.ref_nnet_set_batchnorm(R.h, i, C.c_int(d), C.c_int(d), C.c_float(1e-3), C.c_float(1.0), C.c_float(1000.0),
                                            mean.ctypes.data_as(f32p), var.ctypes.data_as(f32p)) == 0, name
    raw = str(tmp_path / "final.raw")
    L.ref_nnet_write.argtypes = [C.c_void_p, C.c_char_p, C.c_int]
    assert L.ref_nnet_write(R.h, raw.encode(), 1) == 0
    arch, W = NM.load_kaldi_raw(raw)
    assert [(x["type"], x["name"]) for x in arch["layers"]] == [
        ("idct", "idct"), ("batchnorm", "batchnorm0"), ("delta", "delta"), ("relu-batchnorm", "tdnn1"), ("tdnnf", "tdnnf2"), ("tdnnf", "tdnnf3"),
        ("tdnnf", "tdnnf4"), ("linear", "prefinal-l"), ("prefinal", "prefinal-chain"), ("output", "output")]
    assert arch["layers"][3]["append_ivector"] == pytest.approx(0.4)
    try:
        from kaldi_b200.model import KaldiModel
        assert KaldiModel(raw, is_mdl=False).layer_types() == [(x["type"], x["name"]) for x in arch["layers"]]
    except OSError:
        pass
    assert L.ref_nnet_prepare(R.h, C.c_int(20), C.c_int(3), C.c_float(1.0), None, C.c_int(0), C.c_int(1)) == 0
    info = (C.c_int * 4)()
    L.ref_nnet_info(R.h, info)
    R.left_context, R.right_context, R.frames_per_chunk, R.output_dim = list(info)
    T = 64
    feats = (rng.standard_normal((T, 40)) * 10).astype(np.float32)
    iv = rng.standard_normal((T, 100)).astype(np.float32)
    ref = R.forward(feats, iv, period=1)
    prog = NM.compile_program(arch, dict(W, priors=np.ones(48, np.float32)), T, 21, use_priors=False)
    out = PI.run_program(prog, feats, iv[R.chunk_ivector_rows(T, T, 1)])
    assert np.abs(ref).max() > 1e-2 and ref.std() > 1e-3                                  # a real comparison, not 0 == 0
    assert out.shape == ref.shape and np.abs(out - ref).max() <= 1e-4 * np.abs(ref).max()


WSJ_TDNN_1F_XCONFIG = """input dim=100 name=ivector
input dim=40 name=input
fixed-affine-layer name=lda input=Append(-2,-1,0,1,2,ReplaceIndex(ivector, t, 0)) affine-transform-file={lda}
relu-batchnorm-layer name=tdnn1 l2-regularize=0.01 dim=56
relu-batchnorm-layer name=tdnn2 l2-regularize=0.01 dim=56 input=Append(-1,0,1)
relu-batchnorm-layer name=tdnn3 l2-regularize=0.01 dim=56
relu-batchnorm-layer name=tdnn4 l2-regularize=0.01 dim=56 input=Append(-1,0,1)
relu-batchnorm-layer name=tdnn5 l2-regularize=0.01 dim=56
relu-batchnorm-layer name=tdnn6 l2-regularize=0.01 dim=56 input=Append(-3,0,3)
relu-batchnorm-layer name=tdnn7 l2-regularize=0.01 dim=56 input=Append(-3,0,3)
relu-batchnorm-layer name=tdnn8 l2-regularize=0.01 dim=56 input=Append(-6,-3,0)
relu-batchnorm-layer name=prefinal-chain l2-regularize=0.01 dim=56
output-layer name=output l2-regularize=0.005 include-log-softmax=false dim=48
relu-batchnorm-layer name=prefinal-xent l2-regularize=0.01 input=tdnn8 dim=56
output-layer name=output-xent l2-regularize=0.005 dim=48 learning-rate-factor=5.0
"""


@pytest.mark.parametrize("binary", [1, 0])
def test_wsj_tdnn_1f_shape_from_the_references_xconfig_library(tmp_path, binary):
    """The chain TDNN without factorisation (egs/wsj/s5/local/chain/tuning/run_tdnn_1f.sh:165-186; 90 tuning scripts of the
    reference use this layer form): LDA over Append(-2..2, ivector), relu-batchnorm-layers over Append(-1,0,1), Append(-3,0,3),
    Append(-6,-3,0), both output branches; narrower layers, same structure."""
    if not os.path.isdir(STEPS):
        pytest.skip("the reference's xconfig library exists in the build container only")
    from oracle import nnet_oracle as NO
    from oracle import program_interp as PI
    sys.path.insert(0, STEPS)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        import libs.nnet3.xconfig.parser as xparser
    rng = np.random.default_rng(11)
    lda = str(tmp_path / "lda.mat")
    KIO.write_matrix(lda, (rng.standard_normal((300, 301)) / 17).astype(np.float32), binary=False)
    xc = str(tmp_path / "network.xconfig")
    open(xc, "w").write(WSJ_TDNN_1F_XCONFIG.format(lda=lda))
    lines = [line for layer in xparser.read_xconfig_file(xc) for base, line in layer.get_full_config() if base == "final"]
    config = "\n".join(lines) + "\n"
    assert "Append(Offset(tdnn7.batchnorm, -6), Offset(tdnn7.batchnorm, -3), tdnn7.batchnorm)" in config
    R = NO.RefNnet.__new__(NO.RefNnet)
    L = R.lib = C.CDLL(NO._SO)
    L.ref_nnet_create.restype = C.c_void_p
    L.ref_nnet_component_name.restype = C.c_char_p
    L.ref_nnet_component_type.restype = C.c_char_p
    R.h = C.c_void_p(L.ref_nnet_create(config.encode()))
    assert R.h, "the reference rejected its own xconfig output"
    R.arch = {"frame_subsampling_factor": 3}
    _randomise_parameters(L, R.h, rng)
    f32p = C.POINTER(C.c_float)
    for i in range(L.ref_nnet_num_components(R.h)):
        name, typ = L.ref_nnet_component_name(R.h, i).decode(), L.ref_nnet_component_type(R.h, i).decode()
        if typ == "BatchNormComponent":
            mean, var = (rng.standard_normal(56) * 0.1).astype(np.float32), rng.uniform(0.5, 1.5, 56).astype(np.float32)
            assert L.ref_nnet_set_batchnorm(R.h, i, C.c_int(56), C.c_int(56), C.c_float(1e-3), C.c_float(1.0), C.c_float(1000.0),
                                            mean.ctypes.data_as(f32p), var.ctypes.data_as(f32p)) == 0, name
    raw = str(tmp_path / "final.raw")
    L.ref_nnet_write.argtypes = [C.c_void_p, C.c_char_p, C.c_int]
    assert L.ref_nnet_write(R.h, raw.encode(), binary) == 0
    # +-3 splices without a stride-3 TDNN-F layer do not decide --frame-subsampling-factor (chain run_tdnn_1f.sh: 3, the plain
    # nnet3 aishell run_tdnn_1a.sh with the same Append(-3,0,3): 1): both readers refuse to guess, and take what the caller states
    from kaldi_b200.kaldi_io import KaldiFormatError
    with pytest.raises(KaldiFormatError, match="frame subsampling factor"):
        NM.load_kaldi_raw(raw)
    assert NM.load_kaldi_raw(raw, frame_subsampling_factor=1)[0]["frame_subsampling_factor"] == 1
    arch, W = NM.load_kaldi_raw(raw, frame_subsampling_factor=3)
    want = NM.arch_wsj_tdnn_1f(48, dim=56)
    assert [(x["type"], x["name"], x.get("time_offsets")) for x in arch["layers"]] == \
           [(x["type"], x["name"], x.get("time_offsets")) for x in want["layers"]]
    assert (arch["feat_dim"], arch["ivector_dim"], arch["num_pdfs"], arch["frame_subsampling_factor"]) == (40, 100, 48, 3)
    try:
        from kaldi_b200.model import KaldiModel
        from kaldi_b200.nnet_compile import _Layer
        m0 = KaldiModel(raw, is_mdl=False)
        assert m0.frame_subsampling_ambiguous
        from kaldi_b200._lib import B2kError
        with pytest.raises(B2kError, match="frame subsampling factor"):
            m0.compile(70)
        assert KaldiModel(raw, is_mdl=False, frame_subsampling_factor=1).frame_subsampling_factor == 1
        m = KaldiModel(raw, is_mdl=False, frame_subsampling_factor=3)
        assert not m.frame_subsampling_ambiguous and m.frame_subsampling_factor == 3
    except OSError as e:
        pytest.skip(str(e))
    assert m.layer_types() == [(x["type"], x["name"]) for x in want["layers"]]
    a = C.cast(m.layers_ptr, C.POINTER(_Layer))
    assert [list(a[i].time_offsets[:a[i].n_time_offsets]) for i in range(m.n_layers)] == [x.get("time_offsets", []) for x in want["layers"]]
    got = m.weights()
    for k, v in W.items():
        np.testing.assert_array_equal(got[k].reshape(-1), np.asarray(v, np.float32).reshape(-1), err_msg=k)
    # forward: the reference's looped computation of ITS model against the programs compiled (Python and C++) from what was read
    assert L.ref_nnet_prepare(R.h, C.c_int(20), C.c_int(3), C.c_float(1.0), None, C.c_int(0), C.c_int(1)) == 0
    info = (C.c_int * 4)()
    L.ref_nnet_info(R.h, info)
    R.left_context, R.right_context, R.frames_per_chunk, R.output_dim = list(info)
    assert (R.left_context, R.right_context) == NM.model_context(arch) == (16, 10)        # 2+1+1+3+3+6 / 2+1+1+3+3
    T = 70
    feats = (rng.standard_normal((T, 40)) * 10).astype(np.float32)
    iv = rng.standard_normal((T, 100)).astype(np.float32)
    ref = R.forward(feats, iv, period=1)
    assert np.abs(ref).max() > 1e-2 and ref.std() > 1e-3
    civ = iv[R.chunk_ivector_rows(T, T, 1)]
    prog = NM.compile_program(arch, dict(W, priors=np.ones(48, np.float32)), T, 21, use_priors=False)
    out = PI.run_program(prog, feats, civ)
    assert out.shape == ref.shape and np.abs(out - ref).max() <= 1e-4 * np.abs(ref).max()
    from kaldi_b200 import _lib as LB
    from kaldi_b200.nnet_compile import _Node, _Op
    B = LB.lib()
    hp = m.compile(T, 21, use_priors=False)
    try:
        nn, no, bl = C.c_int32(), C.c_int32(), C.c_int64()
        B.b2k_nnet_program_sizes.argtypes = [C.c_void_p] * 4
        assert B.b2k_nnet_program_sizes(hp, C.byref(nn), C.byref(no), C.byref(bl)) == 0
        for f, rt in (("b2k_nnet_program_nodes", C.POINTER(_Node)), ("b2k_nnet_program_ops", C.POINTER(_Op)),
                      ("b2k_nnet_program_blob", C.POINTER(C.c_float))):
            getattr(B, f).restype = rt
            getattr(B, f).argtypes = [C.c_void_p]
        nodes = [B.b2k_nnet_program_nodes(hp)[i] for i in range(nn.value)]
        ops = [B.b2k_nnet_program_ops(hp)[i] for i in range(no.value)]
        blob = np.ctypeslib.as_array(B.b2k_nnet_program_blob(hp), shape=(bl.value,)).copy()
        out2 = PI.run_program(PI.program_from_abi(nodes, ops, blob), feats, civ)
    finally:
        B.b2k_nnet_program_destroy.argtypes = [C.c_void_p]
        B.b2k_nnet_program_destroy(hp)
    assert out2.shape == ref.shape and np.abs(out2 - ref).max() <= 1e-4 * np.abs(ref).max()


@pytest.mark.parametrize("edit", ["dead-layer", "two-sources", "skip"])
def test_skip_connections_are_an_error_not_a_different_network(tmp_path, edit):
    """A layer that reads a layer other than the one before it (dense / skip TDNNs) is outside the supported families: both
    readers must refuse the file instead of returning the chain without the skip.  A layer nothing reads any more is not a
    skip: it is dropped like the cross-entropy branch, and what remains is the network the reference computes."""
    from oracle import nnet_oracle as NO
    if not os.path.exists(NO._SO):
        pytest.skip("oracle/_ref nnet3 library not present")
    arch = NM.arch_tiny_tdnn()
    W = NM.random_weights(arch, seed=1)
    config = NM.to_nnet3_config(arch, W, str(tmp_path))
    old = "input=Append(Offset(tdnn2.batchnorm, -3), tdnn2.batchnorm, Offset(tdnn2.batchnorm, 3))"
    assert config.count(old) == 1
    new = {"dead-layer": "input=Append(Offset(tdnn1.batchnorm, -3), tdnn1.batchnorm, Offset(tdnn1.batchnorm, 3))",
           "two-sources": "input=Append(Offset(tdnn1.batchnorm, -3), tdnn2.batchnorm, Offset(tdnn2.batchnorm, 3))",
           "skip": old}[edit]
    config = config.replace(old, new)
    if edit == "skip":           # tdnn4 reads tdnn2 past tdnn3, and tdnn3 stays alive through the xent-like second output
        old4 = "input=Append(Offset(tdnn3.batchnorm, -6), Offset(tdnn3.batchnorm, -3), tdnn3.batchnorm)"
        assert config.count(old4) == 1
        config = config.replace(old4, "input=Append(Offset(tdnn2.batchnorm, -6), Offset(tdnn2.batchnorm, -3), tdnn2.batchnorm)")
        config = config.replace("component-node name=prefinal-chain.affine component=prefinal-chain.affine input=tdnn4.batchnorm",
                                "component-node name=prefinal-chain.affine component=prefinal-chain.affine input=Sum(tdnn4.batchnorm, tdnn3.batchnorm)")
        assert "Sum(tdnn4.batchnorm, tdnn3.batchnorm)" in config
    L = C.CDLL(NO._SO)
    L.ref_nnet_create.restype = C.c_void_p
    h = C.c_void_p(L.ref_nnet_create(config.encode()))
    assert h, "the reference accepts the network"
    raw = str(tmp_path / "skip.raw")
    L.ref_nnet_write.argtypes = [C.c_void_p, C.c_char_p, C.c_int]
    assert L.ref_nnet_write(h, raw.encode(), 1) == 0
    try:
        from kaldi_b200 import _lib as LB
        from kaldi_b200.model import KaldiModel
        LB.lib()
    except OSError as e:
        pytest.skip(str(e))
    if edit == "dead-layer":
        want = [x["name"] for x in arch["layers"] if x["name"] != "tdnn2"]
        assert [x["name"] for x in NM.load_kaldi_raw(raw, frame_subsampling_factor=3)[0]["layers"]] == want
        assert [n for _, n in KaldiModel(raw, is_mdl=False).layer_types()] == want
        return
    with pytest.raises(KIO.KaldiFormatError, match="unsupported|skip|source"):
        NM.load_kaldi_raw(raw, frame_subsampling_factor=3)
    with pytest.raises(LB.B2kError):
        KaldiModel(raw, is_mdl=False)
