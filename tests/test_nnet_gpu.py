"""GPU parity tests of the nnet3 executor (through the C-ABI) against the numpy
restatement, the compiled reference (oracle/_ref) and its golden output.
Tolerance: 1e-4 of the output scale (north star); observed ~1e-6."""
import os

import numpy as np
import pytest

from kaldi_b200 import nnet_model as NM

pytestmark = pytest.mark.gpu
RTOL_SCALE = 1e-4
GOLD = os.path.join(os.path.dirname(__file__), "golden", "nnet_golden.npz")


def _inputs(T, seed, n_chunks):
    rng = np.random.default_rng(seed)
    return ((rng.standard_normal((T, 40)) * 10).astype(np.float32),
            rng.standard_normal((n_chunks, 100)).astype(np.float32))


def test_golden_from_reference():
    from kaldi_b200.nnet import NnetComputer
    g = np.load(GOLD)
    arch = NM.arch_tiny(64)
    W = NM.random_weights(arch, seed=11)
    nc = NnetComputer(arch, W, num_frames=g["feats"].shape[0], max_batch=2)
    out = nc.forward([g["feats"], g["feats"]], [g["chunk_ivectors"], g["chunk_ivectors"]])
    for o in out:
        assert np.abs(o - g["ref_out"]).max() <= RTOL_SCALE * np.abs(g["ref_out"]).max()


@pytest.mark.parametrize("front,log_softmax,T", [("idct-delta", False, 130), ("lda", False, 64), ("idct-delta", True, 21),
                                                 ("lda", False, 7)])
def test_vs_compiled_reference_and_restatement(front, log_softmax, T):
    from kaldi_b200.nnet import NnetComputer
    from oracle import nnet_oracle as NO
    arch = NM.arch_tiny(64, front=front)
    arch["layers"][-1]["log_softmax"] = log_softmax
    W = NM.random_weights(arch, seed=3)
    nc = NnetComputer(arch, W, num_frames=T, max_batch=3, acoustic_scale=0.9)
    R = NO.RefNnet(arch, W, frames_per_chunk=20, acoustic_scale=0.9)
    batch = [_inputs(T, s, nc.n_chunks) for s in range(3)]
    outs = nc.forward([b[0] for b in batch], [b[1] for b in batch])
    for (feats, civ), o in zip(batch, outs):
        mine = NO.forward_dense(arch, W, feats, civ, frames_per_chunk=21, acoustic_scale=0.9)
        # reference: i-vector matrix with one row per chunk, read at row end_input_frame/period
        period = 10 ** 6
        ref_rows = []
        iv_full = np.zeros((0, 100), np.float32)
        # build an online_ivectors matrix such that chunk n reads civ[n]: period 1, row index = end_input_frame(n)
        ends = [(n + 1) * R.frames_per_chunk + R.right_context for n in range(nc.n_chunks)]
        mat = np.zeros((ends[-1] + 1, 100), np.float32)
        prev = 0
        for n, e in enumerate(ends):
            mat[prev:e + 1] = civ[n]
            prev = e + 1
        ref = R.forward(feats, mat, period=1)
        scale = np.abs(ref).max()
        assert np.abs(mine - ref).max() <= RTOL_SCALE * scale
        assert o.shape == ref.shape
        assert np.abs(o - ref).max() <= RTOL_SCALE * scale, np.abs(o - ref).max() / scale


def test_mini_librispeech_full_utterance():
    """config[1] shape: 998 feature frames -> 333 output frames x 2336 pdfs."""
    from kaldi_b200.nnet import NnetComputer
    from oracle import nnet_oracle as NO
    arch = NM.arch_mini_librispeech_1k()
    W = NM.random_weights(arch, seed=0)
    T = 998
    nc = NnetComputer(arch, W, num_frames=T, max_batch=4)
    assert nc.n_out == 333 and nc.n_chunks == 48
    batch = [_inputs(T, s, nc.n_chunks) for s in range(4)]
    outs = nc.forward([b[0] for b in batch], [b[1] for b in batch])
    feats, civ = batch[2]
    mine = NO.forward_dense(arch, W, feats, civ)
    scale = np.abs(mine).max()
    assert outs[2].shape == (333, 2336)
    assert np.abs(outs[2] - mine).max() <= RTOL_SCALE * scale
    # batch independence: same input in another slot gives the same bits
    outs2 = nc.forward([batch[2][0], batch[0][0]], [batch[2][1], batch[0][1]])
    assert np.array_equal(outs2[0], outs[2]) and np.array_equal(outs2[1], outs[0])


def test_missing_ivectors_is_an_error():
    from kaldi_b200.nnet import NnetComputer
    from kaldi_b200._lib import B2kError
    import torch
    arch = NM.arch_tiny(64)
    nc = NnetComputer(arch, NM.random_weights(arch, 0), num_frames=30, max_batch=1)
    x = torch.zeros(30, 40, device="cuda"); o = torch.zeros(10, 64, device="cuda")
    with pytest.raises(B2kError):
        nc.Run([x.data_ptr()], 40, None, 0, [o.data_ptr()], 64)


@pytest.mark.parametrize("conv_mode", ["dense", "patch"])
@pytest.mark.parametrize("T", [90, 23])
def test_cnn_tdnnf_front_end_vs_compiled_reference(T, conv_mode):
    """CNN-TDNN-F (BASELINE config 3 family): TimeHeightConvolutionComponent + ReLU + block BatchNorm layers,
    combine-feature-maps, the per-chunk i-vector linear+batchnorm branch, a no-bypass TDNN-F layer."""
    from kaldi_b200.nnet import NnetComputer
    from oracle import nnet_oracle as NO
    from oracle.program_interp import run_program
    arch = NM.arch_tiny_cnn()
    W = NM.random_weights(arch, seed=7)
    nc = NnetComputer(arch, W, num_frames=T, max_batch=2, acoustic_scale=0.9, conv_mode=conv_mode)
    R = NO.RefNnet(arch, W, frames_per_chunk=20, acoustic_scale=0.9)
    batch = [_inputs(T, s, nc.n_chunks) for s in range(2)]
    outs = nc.forward([b[0] for b in batch], [b[1] for b in batch])
    prog = NM.compile_program(arch, W, T, 21, acoustic_scale=0.9, conv_mode=conv_mode)
    for (feats, civ), o in zip(batch, outs):
        ends = [(n + 1) * R.frames_per_chunk + R.right_context for n in range(nc.n_chunks)]
        mat = np.zeros((ends[-1] + 1, 100), np.float32)
        prev = 0
        for n, e in enumerate(ends):
            mat[prev:e + 1] = civ[n]
            prev = e + 1
        ref = R.forward(feats, mat, period=1)
        scale = np.abs(ref).max()
        assert o.shape == ref.shape
        assert np.abs(o - ref).max() <= RTOL_SCALE * scale
        assert np.abs(o - run_program(prog, feats, civ)).max() <= RTOL_SCALE * scale
