"""Chunked nnet3 with carried context (b2k_nnet_stream_*, the counterpart of cuda_decoder::BatchedStaticNnet3,
cudadecoder/batched-static-nnet3.{h,cc}): frame bookkeeping against a restatement of BatchContextSwitch, the window program
against the reference's own compiled forward (oracle/_ref nnet3) on the CPU, and the device path against both."""
import os

import numpy as np
import pytest

from kaldi_b200 import nnet_model as NM


def _lib_or_skip():
    try:
        from kaldi_b200 import nnet_compile as NC
        NC._lib.lib()
        return NC
    except Exception as e:       # libb2k.so not built
        pytest.skip(str(e))


def _account_restated(L, R, sub, in_ctx, n_new, flush):
    """batched-static-nnet3.cc:151-190, line by line."""
    nframes_in_batch = n_new
    if in_ctx == 0:
        nframes_in_batch += L
    else:
        nframes_in_batch += in_ctx
    if flush:
        nframes_in_batch += R
    after = min(nframes_in_batch, L + R)
    minus = max(0, nframes_in_batch - (L + R))
    return after, (minus + sub - 1) // sub


def test_frame_bookkeeping_equals_batch_context_switch():
    NC = _lib_or_skip()
    rng = np.random.default_rng(0)
    for _ in range(2000):
        L, R, sub = int(rng.integers(0, 45)), int(rng.integers(0, 45)), int(rng.integers(1, 5))
        in_ctx = int(rng.integers(0, L + R + 1))
        flush = bool(rng.integers(0, 2))
        n_new = 0 if flush else int(rng.integers(0, 80))
        assert NC.stream_account(L, R, sub, in_ctx, n_new, flush) == _account_restated(L, R, sub, in_ctx, n_new, flush)
    with pytest.raises(Exception):
        NC.stream_account(3, 3, 3, 0, 5, True)          # the flush carries no new frames (:163)


class _StreamSim:
    """numpy restatement of RunBatch for ONE channel over the C++-compiled window program: build_batch_with_context /
    save_context_from_batch / the flush variant (batched-static-nnet3-kernels.cu:29-205), rows past the valid ones zero."""

    def __init__(self, arch, W, fpc, use_priors=False, acoustic_scale=1.0):
        from oracle import program_interp as PI
        NC = _lib_or_skip()
        self.PI = PI
        self.L, self.R = NC.model_context(arch)
        self.sub = arch["frame_subsampling_factor"]
        self.fpc, self.opc, self.W = fpc, (self.sub - 1 + fpc) // self.sub, fpc + self.L + self.R
        cp = NC.CompiledProgram(arch, W, self.W, self.sub, acoustic_scale=acoustic_scale, use_priors=use_priors,
                                window=(self.L, self.opc))
        assert (cp.n_out, cp.n_chunks, cp.left_context, cp.right_context) == (self.opc, 1, self.L, self.R)
        self.prog = PI.program_from_abi(cp.nodes, cp.ops, cp.blob)
        self.dim = arch["feat_dim"]
        self.ctx = np.zeros((self.L + self.R, self.dim), np.float32)
        self.in_ctx = 0
        self.NC = NC

    def chunk(self, new, ivector):
        n = new.shape[0]
        rows = ([new[0]] * self.L if self.in_ctx == 0 else []) + list(self.ctx[:self.in_ctx]) + list(new)
        win = np.zeros((self.W, self.dim), np.float32)
        win[:len(rows)] = np.asarray(rows, np.float32).reshape(len(rows), self.dim)
        after, n_out = self.NC.stream_account(self.L, self.R, self.sub, self.in_ctx, n, False)
        out = self.PI.run_program(self.prog, win, None if ivector is None else ivector[None, :])
        n_copy = min(len(rows), self.L + self.R)
        self.ctx[:n_copy] = win[len(rows) - n_copy:len(rows)]
        self.in_ctx = after
        return out[:n_out]

    def flush(self, ivector):
        rows = list(self.ctx[:self.in_ctx]) + [self.ctx[self.in_ctx - 1]] * self.R
        win = np.zeros((self.W, self.dim), np.float32)
        win[:len(rows)] = np.asarray(rows, np.float32)
        after, n_out = self.NC.stream_account(self.L, self.R, self.sub, self.in_ctx, 0, True)
        self.in_ctx = after
        return self.PI.run_program(self.prog, win, None if ivector is None else ivector[None, :])[:n_out]


def _ref_or_skip(arch, W):
    from oracle import nnet_oracle as NO
    if not (os.path.exists(NO._SO) or os.path.isdir("/root/reference")):
        pytest.skip("oracle/_ref nnet3 library not present")
    return NO.RefNnet(arch, W, use_priors=False)


ARCHS = {"tiny": lambda: NM.arch_tiny(64), "tiny-lda": lambda: NM.arch_tiny(64, front="lda"), "tdnn": NM.arch_tiny_tdnn,
         "cnn": NM.arch_tiny_cnn}


@pytest.mark.parametrize("which,fpc,T", [("tiny", 21, 99), ("tiny", 51, 159), ("tiny", 9, 36), ("tiny-lda", 21, 95),
                                         ("tdnn", 21, 75), ("cnn", 21, 66), ("tiny", 21, 5), ("tiny", 21, 24)])
def test_chunks_reassemble_the_reference_whole_utterance_forward(which, fpc, T):
    """A constant i-vector, chunks whose first output lies on the subsampling grid ((first chunk - right context), the
    chunk length and the LAST chunk's length multiples of 3: the flush continues at (frames so far - right context) too):
    the concatenated chunk outputs + the flush are the reference's whole-utterance forward (DecodableNnetSimpleLooped,
    compiled in oracle/_ref), tolerance 1e-4 of the output scale as in the other nnet3 tests."""
    arch = ARCHS[which]()
    W = NM.random_weights(arch, seed=9)
    R = _ref_or_skip(arch, W)
    rng = np.random.default_rng(T)
    feats = (rng.standard_normal((T, arch["feat_dim"])) * 10).astype(np.float32)
    iv = rng.standard_normal(100).astype(np.float32)
    ref = R.forward(feats, np.tile(iv, (T, 1)), period=1)
    sim = _StreamSim(arch, W, fpc)
    first = fpc if (fpc - sim.R) % 3 == 0 else fpc - ((fpc - sim.R) % 3)      # keeps the second chunk on the grid
    outs, pos = [], 0
    while pos < T:
        n = min(first if pos == 0 else fpc, T - pos)
        outs.append(sim.chunk(feats[pos:pos + n], iv))
        pos += n
    outs.append(sim.flush(iv))
    got = np.concatenate(outs, 0)
    assert got.shape == ref.shape, (got.shape, ref.shape)
    assert np.abs(got - ref).max() <= 1e-4 * np.abs(ref).max()


def test_off_grid_chunks_follow_the_reference_arithmetic():
    """frames_per_chunk = 20 with a right context of 9: the second call's first output sits at input time 11, off the 3-grid --
    the reference's own arithmetic (batched-static-nnet3.cc:123-152: outputs at window time k*3 after the context).  Every
    output must equal the whole-utterance forward of the utterance cut so that this time lies on ITS grid."""
    arch = NM.arch_tiny(64)
    W = NM.random_weights(arch, seed=3)
    R = _ref_or_skip(arch, W)
    rng = np.random.default_rng(5)
    T, fpc = 130, 20
    feats = (rng.standard_normal((T, 40)) * 10).astype(np.float32)
    iv = rng.standard_normal(100).astype(np.float32)
    sim = _StreamSim(arch, W, fpc)
    shifted = {s: R.forward(feats[s:], np.tile(iv, (T - s, 1)), period=1) for s in range(3)}
    pos, t_next, checked = 0, None, 0
    while pos < T:
        n = min(fpc, T - pos)
        in_ctx = sim.in_ctx
        out = sim.chunk(feats[pos:pos + n], iv)
        t0 = 0 if in_ctx == 0 else pos - sim.R                 # input time of this call's first output
        for k in range(out.shape[0]):
            t = t0 + 3 * k
            s = t % 3
            if t - s >= sim.L + 3:                              # away from the cut's own left padding
                want = shifted[s][(t - s) // 3]
                assert np.abs(out[k] - want).max() <= 1e-4 * np.abs(want).max()
                checked += 1
        pos += n
    assert checked > 20


def _looped_windows(arch, feats, chunk_iv, C, L, Rc, k1):
    """The inputs of the looped schedule, chunk by chunk: window n = input frames [n*C - L, (n+1)*C + Rc) with indices clamped
    to the utterance (decodable-online-looped.cc:150-160), i-vector rows = the i-vectors of chunks n-(k1-1) .. n (chunk 0's
    in front of the utterance)."""
    T = feats.shape[0]
    n_out = (T + 2) // 3
    n_chunks = (n_out * 3 + C - 1) // C
    for n in range(n_chunks):
        idx = np.clip(np.arange(n * C - L, (n + 1) * C + Rc), 0, T - 1)
        ivr = np.stack([chunk_iv[max(0, n - (k1 - 1) + r)] for r in range(k1)])
        yield n, feats[idx], ivr, min(C // 3, n_out - n * (C // 3))


@pytest.mark.parametrize("which,T", [("tiny", 100), ("tiny", 21), ("tiny", 5), ("tiny-lda", 77), ("cnn", 64), ("tdnn", 90)])
def test_looped_windows_reproduce_the_reference_looped_forward(which, T):
    """Window programs with the looped i-vector arithmetic (ivector_rows > 1): an i-vector that CHANGES with every chunk, the
    chunks evaluated one window at a time, equal the reference's DecodableNnetSimpleLooped run over the whole utterance
    (Round(ivector, C) with its lag, nnet-compile-looped.cc:179-205)."""
    from oracle import program_interp as PI
    NC = _lib_or_skip()
    arch = ARCHS[which]()
    W = NM.random_weights(arch, seed=4)
    R = _ref_or_skip(arch, W)
    C = R.frames_per_chunk
    rng = np.random.default_rng(T + 1)
    feats = (rng.standard_normal((T, arch["feat_dim"])) * 10).astype(np.float32)
    iv = rng.standard_normal((T, 100)).astype(np.float32)
    ref = R.forward(feats, iv, period=1)
    chunk_iv = iv[R.chunk_ivector_rows(T, T, 1)]
    L, Rc = NC.model_context(arch)
    k1 = NC.looped_ivector_rows(arch, C)
    assert k1 == -(-L // C) + (C + Rc - 1) // C + 1
    cp = NC.CompiledProgram(arch, W, C + L + Rc, C, use_priors=False, window=(L, C // 3, k1))
    assert cp.n_chunks == k1
    prog = PI.program_from_abi(cp.nodes, cp.ops, cp.blob)
    outs = [PI.run_program(prog, win, ivr)[:keep] for _n, win, ivr, keep in _looped_windows(arch, feats, chunk_iv, C, L, Rc, k1)]
    got = np.concatenate(outs, 0)
    assert got.shape == ref.shape
    assert np.abs(got - ref).max() <= 1e-4 * np.abs(ref).max()


@pytest.mark.parametrize("which,T,fpc,period", [("tiny", 100, 50, 10), ("tiny", 37, 21, 1), ("tiny-lda", 77, 50, 7), ("cnn", 64, 30, 10),
                                                 ("tdnn", 90, 50, 10), ("tiny", 4, 50, 10)])
def test_offline_chunks_reproduce_the_reference_simple_decodable(which, T, fpc, period):
    """What NnetSimpleComputerB2k::Compute does (kaldi_b200/host/b2k_nnet3_shims.h: the decodable of nnet3-latgen-faster), restated
    over the C++-compiled window program: every chunk of --frames-per-chunk outputs from its own clamped window with the online
    i-vector nearest its middle frame, against the reference's own DecodableNnetSimple (nnet3/nnet-am-decodable-simple.cc,
    compiled in oracle/_ref)."""
    from oracle import program_interp as PI
    NC = _lib_or_skip()
    arch = ARCHS[which]()
    W = NM.random_weights(arch, seed=8)
    R = _ref_or_skip(arch, W)
    if not hasattr(R.lib, "ref_nnet_forward_simple"):
        pytest.skip("oracle/_ref library predates ref_nnet_forward_simple")
    sf = arch["frame_subsampling_factor"]
    rng = np.random.default_rng(T)
    feats = (rng.standard_normal((T, arch["feat_dim"])) * 10).astype(np.float32)
    online_iv = rng.standard_normal(((T + period - 1) // period, 100)).astype(np.float32)
    ref = R.forward_simple(feats, online_iv, period, fpc)
    fpc_r = sf * ((fpc + sf - 1) // sf)
    opc = fpc_r // sf
    L, Rc = NC.model_context(arch)
    window = (opc - 1) * sf + 1 + L + Rc
    cp = NC.CompiledProgram(arch, W, window, sf, use_priors=False, window=(L, opc, 1))
    prog = PI.program_from_abi(cp.nodes, cp.ops, cp.blob)
    n_sub = (T + sf - 1) // sf
    outs = []
    for c in range((n_sub + opc - 1) // opc):
        first_out = c * opc * sf
        n = min(n_sub - c * opc, opc)
        last_out = first_out + (n - 1) * sf
        win = feats[np.clip(first_out - L + np.arange(window), 0, T - 1)]
        iv_frame = min((first_out + (last_out - first_out) // 2) // period, online_iv.shape[0] - 1)
        outs.append(PI.run_program(prog, win, online_iv[iv_frame][None, :])[:n])
    got = np.concatenate(outs, 0)
    assert got.shape == ref.shape
    assert np.abs(got - ref).max() <= 1e-4 * np.abs(ref).max()


@pytest.mark.gpu
@pytest.mark.parametrize("which,T", [("tiny", 100), ("cnn", 64)])
def test_device_looped_stream_reproduces_the_whole_utterance_run(which, T):
    """b2k_nnet_stream in looped mode, fed as DecodableNnetLoopedOnlineBase::AdvanceChunk reads its input (right-context frames
    first, then a chunk per call, indices clamped): the device's own whole-utterance forward with per-chunk i-vectors
    (b2k_nnet_run, itself pinned to the reference's looped forward) chunk by chunk."""
    import torch
    from kaldi_b200.nnet import BatchedStaticNnet3, NnetComputer
    NC = _lib_or_skip()
    arch = ARCHS[which]()
    W = NM.random_weights(arch, seed=4)
    C = 21
    rng = np.random.default_rng(T)
    feats = (rng.standard_normal((T, arch["feat_dim"])) * 10).astype(np.float32)
    whole = NnetComputer(arch, W, T, 1, frames_per_chunk=C, use_priors=False)
    chunk_iv = rng.standard_normal((whole.n_chunks, 100)).astype(np.float32)
    ref = whole.forward([feats], [chunk_iv])[0]
    nn = BatchedStaticNnet3(arch, W, max_batch=1, frames_per_chunk=C, use_priors=False, looped=True)
    L, Rc, k1, opc, P = nn.left_context, nn.right_context, nn.ivector_rows, nn.output_frames_per_chunk, nn.output_dim
    assert k1 == NC.looped_ivector_rows(arch, C) and opc == C // 3
    d_out = torch.zeros(opc, P, device="cuda")
    outs = []
    zeros_iv = torch.zeros(k1, 100, device="cuda")
    for b in range(0, Rc, C):            # the right-context frames first (in pieces of at most a chunk): no output yet
        piece = torch.from_numpy(feats[np.clip(np.arange(b, min(b + C, Rc)), 0, T - 1)]).cuda()
        no, ne = nn.RunBatch([0], [piece.data_ptr()], arch["feat_dim"], [zeros_iv.data_ptr()], [piece.shape[0]], [b == 0], [False],
                             d_out.data_ptr(), 0, P)
        assert no == [0] and ne == [0]
    for n, win, ivr, keep in _looped_windows(arch, feats, chunk_iv, C, L, Rc, k1):
        new = torch.from_numpy(np.ascontiguousarray(win[L + Rc:])).cuda()          # the C frames this chunk adds
        d_iv = torch.from_numpy(ivr).cuda()
        no, ne = nn.RunBatch([0], [new.data_ptr()], arch["feat_dim"], [d_iv.data_ptr()], [C], [False], [False], d_out.data_ptr(), 0, P)
        torch.cuda.synchronize()
        assert no == [opc]
        outs.append(d_out.cpu().numpy()[:keep].copy())
    got = np.concatenate(outs, 0)
    assert got.shape == ref.shape and np.abs(got - ref).max() <= 1e-4 * np.abs(ref).max()


@pytest.mark.gpu
@pytest.mark.parametrize("which,fpc", [("tiny", 21), ("tiny", 51), ("cnn", 21), ("tdnn", 30)])
def test_device_stream_equals_the_window_program_and_the_whole_utterance(which, fpc):
    """Four channels of different lengths through b2k_nnet_stream_run_batch in one batch (channels end at different calls, one
    utterance shorter than a chunk): every call's frame counts and outputs equal the numpy restatement over the same window
    program, and a grid-aligned channel reassembles the device's own whole-utterance forward (b2k_nnet_run)."""
    import torch
    from kaldi_b200.nnet import BatchedStaticNnet3, NnetComputer
    arch = ARCHS[which]()
    W = NM.random_weights(arch, seed=11)
    rng = np.random.default_rng(7)
    lens = [150, 97, 40, 7]
    dim, ivd = arch["feat_dim"], arch["ivector_dim"]
    feats = [(rng.standard_normal((T, dim)) * 10).astype(np.float32) for T in lens]
    ivs = [rng.standard_normal(ivd).astype(np.float32) for _ in lens]
    nn = BatchedStaticNnet3(arch, W, max_batch=4, nchannels=6, frames_per_chunk=fpc, use_priors=False)
    sims = [_StreamSim(arch, W, fpc) for _ in lens]
    assert (nn.left_context, nn.right_context, nn.output_frames_per_chunk) == (sims[0].L, sims[0].R, sims[0].opc)
    opc, P = nn.output_frames_per_chunk, nn.output_dim
    d_feats = [torch.from_numpy(f).cuda() for f in feats]
    d_ivs = [torch.from_numpy(v).cuda() for v in ivs]
    d_out = torch.full((4 * opc, P), float("nan"), device="cuda")
    d_eos = torch.full((4 * opc, P), float("nan"), device="cuda")
    chan_of = [5, 0, 3, 1]
    pos = [0] * 4
    got = [[] for _ in lens]
    want = [[] for _ in lens]
    scale = 0.0
    while any(p < T for p, T in zip(pos, lens)):
        live = [u for u in range(4) if pos[u] < lens[u]]
        n_new = [min(fpc, lens[u] - pos[u]) for u in live]
        first = [pos[u] == 0 for u in live]
        last = [pos[u] + n == lens[u] for u, n in zip(live, n_new)]
        no, ne = nn.RunBatch([chan_of[u] for u in live], [d_feats[u][pos[u]:].data_ptr() for u in live], dim,
                             [d_ivs[u].data_ptr() for u in live], n_new, first, last, d_out.data_ptr(), d_eos.data_ptr(), P)
        torch.cuda.synchronize()
        o, e = d_out.cpu().numpy(), d_eos.cpu().numpy()
        for i, u in enumerate(live):
            w = sims[u].chunk(feats[u][pos[u]:pos[u] + n_new[i]], ivs[u])
            assert no[i] == w.shape[0]
            got[u].append(o[i * opc:i * opc + no[i]]); want[u].append(w)
            if last[i]:
                wf = sims[u].flush(ivs[u])
                assert ne[i] == wf.shape[0]
                got[u].append(e[i * opc:i * opc + ne[i]]); want[u].append(wf)
            else:
                assert ne[i] == 0
            pos[u] += n_new[i]
    for u in range(4):
        g, w = np.concatenate(got[u], 0), np.concatenate(want[u], 0)
        scale = np.abs(w).max()
        assert g.shape == w.shape and np.isfinite(g).all()
        assert np.abs(g - w).max() <= 1e-4 * scale, (u, np.abs(g - w).max(), scale)
    if (fpc - nn.right_context) % 3 == 0 and fpc % 3 == 0:
        for u in (0, 3):
            T = lens[u]
            if T > nn.right_context and (T - nn.right_context) % 3 != 0:
                continue        # the flush continues at (frames so far - right context): off the grid, as in the reference
            whole = NnetComputer(arch, W, T, 1, frames_per_chunk=3 * ((T + 2) // 3) + 3, use_priors=False)
            ref = whole.forward([feats[u]], [ivs[u][None, :]])[0]
            g = np.concatenate(got[u], 0)
            assert g.shape == ref.shape and np.abs(g - ref).max() <= 1e-4 * np.abs(ref).max()


@pytest.mark.gpu
def test_device_stream_rejects_what_the_reference_asserts():
    from kaldi_b200 import _lib
    from kaldi_b200.nnet import BatchedStaticNnet3
    arch = NM.arch_tiny(64)
    W = NM.random_weights(arch, seed=1)
    with pytest.raises(_lib.B2kError):
        BatchedStaticNnet3(arch, W, max_batch=2, frames_per_chunk=6)            # < right context (:172-175)
    with pytest.raises(_lib.B2kError):
        BatchedStaticNnet3(arch, W, max_batch=4, nchannels=2, frames_per_chunk=21)   # nchannels < max_batch (h:62)
    import torch
    nn = BatchedStaticNnet3(arch, W, max_batch=2, frames_per_chunk=21)
    f = torch.zeros(30, 40, device="cuda"); v = torch.zeros(100, device="cuda"); o = torch.zeros(2 * 7, 64, device="cuda")
    with pytest.raises(_lib.B2kError):       # more frames than a chunk (:158)
        nn.RunBatch([0], [f.data_ptr()], 40, [v.data_ptr()], [22], [True], [False], o.data_ptr(), o.data_ptr(), 64)
    with pytest.raises(_lib.B2kError):       # a channel that never had a first chunk
        nn.RunBatch([1], [f.data_ptr()], 40, [v.data_ptr()], [21], [False], [False], o.data_ptr(), o.data_ptr(), 64)
    with pytest.raises(_lib.B2kError):       # the same channel twice
        nn.RunBatch([0, 0], [f.data_ptr()] * 2, 40, [v.data_ptr()] * 2, [21, 21], [True, True], [False, False], o.data_ptr(), o.data_ptr(), 64)
