"""The C++ nnet3 program compiler (kaldi_b200/csrc/nnet_compile.cu, host only) against its Python oracle
nnet_model.compile_program: nodes, ops (every ABI field) and parameter blob must be identical."""
import ctypes as C

import numpy as np
import pytest

from kaldi_b200 import nnet_model as NM


def _python_structs(prog):
    """The ABI structs NnetComputer builds from a Python-compiled program (kaldi_b200/nnet.py)."""
    from kaldi_b200.nnet import _Node, _Op, _term
    nodes = []
    for i, (name, dim, rows, _t0, _step) in enumerate(prog["nodes"]):
        kind = {"input": 1, "ivector": 2, "output": 3}.get(name, 0)
        nodes.append(_Node(dim, rows, kind, prog["arena_off"].get(i, 0)))
    ops = []
    for o in prog["ops"]:
        op = _Op()
        op.out, op.rows = o["out"], o["rows"]
        op.w = op.bias = op.sub_vec = -1
        op.bn_scale, op.bn_offset = o.get("bn_scale", -1), o.get("bn_offset", -1)
        op.out_scale, op.block_dim = 1.0, 1
        if o["type"] == "gemm":
            op.type, op.N, op.K = 0, o["N"], o["K"]
            op.n_terms = len(o["terms"])
            op.hsplit = o.get("hsplit", 1)
            for j, t in enumerate(o["terms"]):
                op.terms[j] = _term(t)
            op.w, op.bias, op.sub_vec = o["w"], o["bias"], o.get("sub_vec", -1)
            op.relu, op.log_softmax, op.out_scale = o["relu"], o["log_softmax"], o["out_scale"]
            if o.get("res"):
                op.has_res, op.res, op.res_alpha = 1, _term(o["res"]), o["res_alpha"]
        else:
            op.type, op.block_dim = 1, o["block_dim"]
            flat = [(b, t) for b, blk in enumerate(o["blocks"]) for t in blk]
            op.n_terms = len(flat)
            for j, (b, t) in enumerate(flat):
                op.terms[j] = _term(t, block=b)
        ops.append(op)
    return nodes, ops


def _fields(s, skip=()):
    return {f[0]: getattr(s, f[0]) for f in s._fields_ if f[0] not in skip}


def _term_tuple(t):
    return tuple(getattr(t, f[0]) for f in t._fields_)


CASES = [("tiny-idct", {}), ("tiny-lda", {}), ("tiny-logsoftmax", {}), ("cnn-patch", {"conv_mode": "patch"}),
         ("cnn-dense", {"conv_mode": "dense"}), ("mini", {}), ("libri-1d", {}), ("libri-cnn", {}), ("tiny-tdnn", {}), ("wsj-1f", {})]


@pytest.mark.parametrize("which,kw", CASES, ids=[c[0] for c in CASES])
@pytest.mark.parametrize("T", [998, 23])
def test_cpp_compiler_equals_python_compiler(which, kw, T):
    try:
        from kaldi_b200.nnet_compile import CompiledProgram
    except Exception as e:       # libb2k.so not built
        pytest.skip(str(e))
    if which == "tiny-idct":
        arch = NM.arch_tiny(64)
    elif which == "tiny-lda":
        arch = NM.arch_tiny(64, front="lda")
    elif which == "tiny-logsoftmax":
        arch = NM.arch_tiny(64)
        arch["layers"][-1]["log_softmax"] = True
    elif which == "tiny-tdnn":
        arch = NM.arch_tiny_tdnn()
    elif which == "wsj-1f":
        arch = NM.arch_wsj_tdnn_1f(256)
    elif which == "mini":
        arch = NM.arch_mini_librispeech_1k(256)
    elif which == "libri-1d":
        arch = NM.arch_librispeech_1d(256)
    elif which == "libri-cnn":
        arch = NM.arch_librispeech_cnn_tdnn_1a(256)
    else:
        arch = NM.arch_tiny_cnn()
    if which.startswith("libri") and T != 998:
        pytest.skip("one length is enough for the full-width models")
    W = NM.random_weights(arch, seed=9)
    prog = NM.compile_program(arch, W, T, 21, acoustic_scale=0.9, **kw)
    cp = CompiledProgram(arch, W, T, 21, acoustic_scale=0.9, **kw)
    assert (cp.n_out, cp.n_chunks, cp.left_context, cp.right_context, cp.model_left, cp.model_right, cp.ivector_m,
            cp.arena_size) == (prog["n_out"], prog["n_chunks"], prog["left_context"], prog["right_context"],
                               prog["model_left"], prog["model_right"], prog["ivector_m"], prog["arena_size"])
    pn, po = _python_structs(prog)
    assert len(cp.nodes) == len(pn) and len(cp.ops) == len(po)
    for i, (a, b) in enumerate(zip(cp.nodes, pn)):
        assert _fields(a) == _fields(b), (i, prog["nodes"][i][0])
    for i, (a, b) in enumerate(zip(cp.ops, po)):
        fa, fb = _fields(a, skip=("terms", "res", "hsplit")), _fields(b, skip=("terms", "res", "hsplit"))
        for k in fa:
            if isinstance(fa[k], float):
                assert abs(fa[k] - fb[k]) <= 1e-7 * max(1.0, abs(fb[k])), (i, k, fa[k], fb[k])
            else:
                assert fa[k] == fb[k], (i, prog["nodes"][po[i].out][0], k, fa[k], fb[k])
        assert max(a.hsplit, 1) == max(b.hsplit, 1)
        for j in range(a.n_terms):
            ta, tb = _term_tuple(a.terms[j]), _term_tuple(b.terms[j])
            assert ta == pytest.approx(tb, rel=1e-7, abs=0), (i, j, ta, tb)
        if a.has_res:
            assert _term_tuple(a.res) == pytest.approx(_term_tuple(b.res), rel=1e-7, abs=0)
    assert cp.blob.shape == prog["blob"].shape
    # weights are copied bit for bit; the derived BatchNorm scales (powf) and log priors (logf) may differ from numpy by an ulp
    np.testing.assert_allclose(cp.blob, prog["blob"], rtol=1e-6, atol=0)
    assert (cp.blob != prog["blob"]).mean() < 0.05


@pytest.mark.parametrize("T", [1, 2, 3, 4, 7, 20, 21, 22, 63, 64])
def test_very_short_utterances_against_the_reference_forward(T):
    """Utterances shorter than the model context and around the chunk boundaries: both compilers agree, and the compiled
    program (numpy interpreter) reproduces the reference's looped forward (north-star tolerance 1e-4; observed ~1e-6)."""
    try:
        from kaldi_b200.nnet_compile import CompiledProgram
    except Exception as e:
        pytest.skip(str(e))
    import os
    from oracle import nnet_oracle as NO
    from oracle import program_interp as PI
    if not (os.path.exists(NO._SO) or os.path.isdir("/root/reference")):
        pytest.skip("oracle/_ref nnet3 library not present")
    arch = NM.arch_tiny(64)
    W = NM.random_weights(arch, seed=9)
    prog = NM.compile_program(arch, W, T, 21, use_priors=False)
    cp = CompiledProgram(arch, W, T, 21, use_priors=False)
    assert (cp.n_out, cp.n_chunks) == (prog["n_out"], prog["n_chunks"]) == ((T + 2) // 3, (((T + 2) // 3) * 3 + 20) // 21)
    np.testing.assert_allclose(cp.blob, prog["blob"], rtol=1e-6, atol=0)
    rng = np.random.default_rng(T)
    feats = (rng.standard_normal((T, 40)) * 10).astype(np.float32)
    iv = rng.standard_normal((T, 100)).astype(np.float32)
    R = NO.RefNnet(arch, W, use_priors=False)
    ref = R.forward(feats, iv, period=1)
    out = PI.run_program(prog, feats, iv[R.chunk_ivector_rows(T, T, 1)])
    assert out.shape == ref.shape and np.abs(out - ref).max() <= 1e-4 * np.abs(ref).max()
