"""CPU tests of the nnet3 oracle: the numpy restatement (oracle/nnet_oracle.py
forward_dense) against the reference's own nnet3 CPU forward compiled in
oracle/_ref (DecodableNnetSimpleLooped over the looped, collapsed computation),
and golden outputs generated from it (tests/golden/nnet_golden.npz)."""
import os

import numpy as np
import pytest

from kaldi_b200 import nnet_model as NM
from oracle import nnet_oracle as NO

HAVE_REF = os.path.exists(NO._SO) or os.path.isdir("/root/reference/src")
GOLD = os.path.join(os.path.dirname(__file__), "golden", "nnet_golden.npz")
RTOL_SCALE = 1e-4      # north star: log-likelihoods within 1e-4 relative (of the output scale)


def _inputs(T, seed):
    rng = np.random.default_rng(seed)
    feats = (rng.standard_normal((T, 40)) * 10).astype(np.float32)
    iv = rng.standard_normal((T, 100)).astype(np.float32)
    return feats, iv


@pytest.mark.skipif(not HAVE_REF, reason="oracle/_ref not available")
@pytest.mark.parametrize("front,log_softmax", [("idct-delta", False), ("lda", False), ("idct-delta", True)])
def test_restatement_vs_compiled_reference(front, log_softmax):
    arch = NM.arch_tiny(64, front=front)
    arch["layers"][-1]["log_softmax"] = log_softmax
    W = NM.random_weights(arch, seed=3)
    R = NO.RefNnet(arch, W, frames_per_chunk=20, acoustic_scale=0.9)
    assert (R.left_context, R.right_context) == NM.model_context(arch)   # ComputeSimpleNnetContext
    assert R.frames_per_chunk == 21                                       # GetChunkSize rounds 20 up to a multiple of 3
    for T in (130, 21, 7, 64):
        feats, iv = _inputs(T, T)
        ref = R.forward(feats, iv, period=1)                              # a different i-vector every chunk
        rows = R.chunk_ivector_rows(T, T, 1)
        mine = NO.forward_dense(arch, W, feats, iv[rows], frames_per_chunk=21, acoustic_scale=0.9)
        assert mine.shape == ref.shape == ((T + 2) // 3, 64)
        assert np.abs(mine - ref).max() <= RTOL_SCALE * np.abs(ref).max()


@pytest.mark.skipif(not HAVE_REF, reason="oracle/_ref not available")
def test_mini_librispeech_shape_vs_reference():
    arch = NM.arch_mini_librispeech_1k(num_pdfs=200)     # full layer shapes, fewer pdfs to keep it quick
    W = NM.random_weights(arch, seed=5)
    R = NO.RefNnet(arch, W)
    assert (R.left_context, R.right_context) == (29, 29)
    feats, iv = _inputs(100, 1)
    ref = R.forward(feats, iv, period=1)
    mine = NO.forward_dense(arch, W, feats, iv[R.chunk_ivector_rows(100, 100, 1)])
    assert np.abs(mine - ref).max() <= RTOL_SCALE * np.abs(ref).max()


def test_restatement_vs_golden_fixture():
    g = np.load(GOLD)
    arch = NM.arch_tiny(64)
    W = NM.random_weights(arch, seed=11)
    mine = NO.forward_dense(arch, W, g["feats"], g["chunk_ivectors"], frames_per_chunk=21)
    assert np.abs(mine - g["ref_out"]).max() <= RTOL_SCALE * np.abs(g["ref_out"]).max()


def test_compile_program_shapes():
    arch = NM.arch_mini_librispeech_1k()
    W = NM.random_weights(arch, 0)
    p = NM.compile_program(arch, W, 998)
    assert p["n_out"] == 333 and p["model_left"] == 29 and p["model_right"] == 29
    # algorithmic FLOPs per output frame (~12 MFLOP for mini_librispeech 1k)
    assert 11e6 < NM.flops_per_output_frame(p) < 14e6
    names = {n[0]: n for n in p["nodes"]}
    assert names["tdnnf6.noop"][4] == 3 and names["tdnnf3.noop"][4] == 1    # time grids: step 3 after tdnnf4
    assert NM.num_parameters(arch, W) == 4466056


def _ref_forward(R, feats, civ, n_chunks):
    """The reference's looped forward with chunk n reading civ[n] (online_ivectors matrix of period 1)."""
    ends = [(n + 1) * R.frames_per_chunk + R.right_context for n in range(n_chunks)]
    mat = np.zeros((ends[-1] + 1, civ.shape[1]), np.float32)
    prev = 0
    for n, e in enumerate(ends):
        mat[prev:e + 1] = civ[n]
        prev = e + 1
    return R.forward(feats, mat, period=1)


@pytest.mark.parametrize("which,T", [("tdnnf", 64), ("cnn", 90), ("cnn", 23), ("cnn-patch", 90), ("cnn-patch", 23), ("tdnn", 64), ("tdnn", 23)])
def test_compiled_program_vs_compiled_reference(which, T):
    """The op program itself (what the CUDA executor runs), interpreted in numpy, against the reference's
    compiled nnet3: covers the CNN-TDNN-F front end (TimeHeightConvolutionComponent as a dense map per
    time offset, combine-feature-maps permutation folded into the weights, per-chunk i-vector branch)."""
    from oracle import nnet_oracle as NO
    from oracle.program_interp import run_program
    arch = NM.arch_tiny_cnn() if which.startswith("cnn") else NM.arch_tiny_tdnn() if which == "tdnn" else NM.arch_tiny(64)
    W = NM.random_weights(arch, seed=7)
    prog = NM.compile_program(arch, W, T, 21, acoustic_scale=0.9, conv_mode="patch" if which == "cnn-patch" else "dense")
    R = NO.RefNnet(arch, W, frames_per_chunk=20, acoustic_scale=0.9)
    assert (prog["model_left"], prog["model_right"]) == (R.left_context, R.right_context)
    rng = np.random.default_rng(5)
    feats = (rng.standard_normal((T, 40)) * 10).astype(np.float32)
    civ = rng.standard_normal((prog["n_chunks"], 100)).astype(np.float32)
    mine = run_program(prog, feats, civ)
    ref = _ref_forward(R, feats, civ, prog["n_chunks"])
    assert mine.shape == ref.shape
    assert np.abs(mine - ref).max() <= 1e-4 * np.abs(ref).max()


@pytest.mark.parametrize("which", ["librispeech_tdnn_1d", "librispeech_cnn_tdnn_1a", "wsj_tdnn_1f"])
def test_baseline_config_architectures_vs_compiled_reference(which):
    """BASELINE.json configs 2 and 3 at full width (fewer pdfs to keep it quick): the compiled op program vs the
    reference's compiled nnet3, and ComputeSimpleNnetContext."""
    from oracle import nnet_oracle as NO
    from oracle.program_interp import run_program
    arch = {"librispeech_tdnn_1d": NM.arch_librispeech_1d, "librispeech_cnn_tdnn_1a": NM.arch_librispeech_cnn_tdnn_1a,
            "wsj_tdnn_1f": NM.arch_wsj_tdnn_1f}[which](num_pdfs=512)
    W = NM.random_weights(arch, seed=1)
    T = 50
    prog = NM.compile_program(arch, W, T, 21)
    R = NO.RefNnet(arch, W, frames_per_chunk=20)
    assert (prog["model_left"], prog["model_right"]) == (R.left_context, R.right_context)
    rng = np.random.default_rng(2)
    feats = (rng.standard_normal((T, 40)) * 10).astype(np.float32)
    civ = rng.standard_normal((prog["n_chunks"], 100)).astype(np.float32)
    ref = _ref_forward(R, feats, civ, prog["n_chunks"])
    assert np.abs(run_program(prog, feats, civ) - ref).max() <= 1e-4 * np.abs(ref).max()
