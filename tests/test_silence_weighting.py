"""Silence weighting of the online i-vector (kaldi_b200/host/b2k_silence_weighting.h: b2k_host::SilenceWeighting, the core of the
OnlineSilenceWeightingB2k shim) against the reference's OWN OnlineSilenceWeighting (online2/online-ivector-feature.cc:465-750,
compiled unmodified in oracle/_ref by oracle/ref_wrap/silence_wrap.cc).

Both are driven by the same "decoder": a random trellis of tokens -- per time index a few tokens with distinct HCLG states, each
with ONE fixed predecessor and incoming label (epsilon tokens included), as the lattice decoder keeps them -- from which the best
path after every chunk is the walk back from a randomly chosen token of the newest frame.  Paths therefore share their history
exactly as the reference assumes (same token => same history), diverge tens of frames back now and then, and hit the case where
only the arc LEAVING an unchanged token changes.  The reference's class sees the walk through the BestPathEnd / TraceBackBestPath
interface (tokens = ids), b2k's the arrays b2k_dec_best_path returns (ilabels + the state each arc enters).  The lists of
(feature frame, weight difference) must be identical, float for float, call after call."""
import ctypes as C
import os
import shutil
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MDL = os.path.join(ROOT, "tests", "golden", "tiny_final.mdl")          # 5 phones, written by the reference (make_model_golden.py)
START = 1 << 20                                                         # the reference side's id of the start token


@pytest.fixture(scope="module")
def ref():
    from oracle import nnet_oracle as NO
    if os.path.isdir("/root/reference"):
        from oracle import ref_nnet
        ref_nnet.build(quiet=True)
    if not os.path.exists(NO._SO):
        pytest.skip("oracle/_ref/libkaldi_ref_nnet3.so not built")
    R = C.CDLL(NO._SO)
    if not hasattr(R, "ref_silw_create"):
        pytest.skip("oracle/_ref predates silence_wrap.cc")
    R.ref_silw_create.restype = C.c_void_p
    R.ref_silw_create.argtypes = [C.c_char_p, C.c_char_p, C.c_float, C.c_float, C.c_int]
    for f in ("ref_silw_destroy", "ref_silw_active"):
        getattr(R, f).argtypes = [C.c_void_p]
    R.ref_silw_traceback.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int]
    R.ref_silw_delta_weights.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int]
    R.ref_silw_nonsilence_frames.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int]
    return R


@pytest.fixture(scope="module")
def mine(tmp_path_factory):
    if not shutil.which("g++"):
        pytest.skip("g++ missing")
    so = str(tmp_path_factory.mktemp("silw") / "libsilw.so")
    r = subprocess.run(["g++", "-std=c++17", "-O1", "-Wall", "-Wextra", "-Werror", "-shared", "-fPIC",
                        "-I" + os.path.join(ROOT, "kaldi_b200", "host"),
                        os.path.join(ROOT, "tests", "cabi", "silence_weighting_capi.cc"), "-o", so], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    M = C.CDLL(so)
    M.silw_create.restype = C.c_void_p
    M.silw_create.argtypes = [C.c_void_p, C.c_int, C.c_char_p, C.c_float, C.c_float, C.c_int]
    for f in ("silw_destroy", "silw_active", "silw_list_ok"):
        getattr(M, f).argtypes = [C.c_void_p]
    M.silw_traceback_from_path.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int]
    M.silw_delta_weights.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int]
    M.silw_nonsilence_frames.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int]
    return M


@pytest.fixture(scope="module")
def tid2phone(ref):
    out = np.zeros(4096, np.int32)
    n = ref.ref_tid2phone(MDL.encode(), out.ctypes.data_as(C.c_void_p), 4096)
    assert n > 0
    return out[:n + 1].copy()


class Trellis:
    """tokens[t] = list of (state, predecessor (t', index) or None, ilabel into the token); lanes keep a history apart for a while"""

    def __init__(self, rng, num_tids, lanes=3, stay=0.93, repeat=0.8, eps=0.15):
        self.rng, self.num_tids, self.lanes, self.stay, self.repeat, self.eps = rng, num_tids, lanes, stay, repeat, eps
        self.tokens = [[(0, None, 0)]]                                   # time 0: the start token
        if rng.random() < 0.5:                                           # ... and perhaps an epsilon successor
            self.tokens[0].append((7, (0, 0), 0))

    def _tid_into(self, t, i):
        while True:
            st, pred, lab = self.tokens[t][i]
            if lab != 0 or pred is None:
                return lab
            t, i = pred

    def grow(self, n):
        rng = self.rng
        for _ in range(n):
            prev, t = self.tokens[-1], len(self.tokens)
            states = rng.choice(40, size=self.lanes + 2, replace=False) + 1
            cur = []
            for lane in range(self.lanes):
                src = lane if (rng.random() < self.stay and lane < len(prev)) else int(rng.integers(len(prev)))
                last = self._tid_into(t - 1, src)
                tid = last if (last > 0 and rng.random() < self.repeat) else int(rng.integers(1, self.num_tids + 1))
                cur.append((int(states[lane]), (t - 1, src), tid))
            k = self.lanes
            while rng.random() < self.eps and k < len(states):           # epsilon tokens of the same time index
                cur.append((int(states[k]), (t, int(rng.integers(len(cur)))), 0))
                k += 1
            self.tokens.append(cur)

    def best_path(self, end_index=None):
        """arcs start first: (ilabels, the state each arc enters, the id of the token each arc leaves)"""
        t = len(self.tokens) - 1
        i = int(self.rng.integers(len(self.tokens[t]))) if end_index is None else end_index
        il, dst, src = [], [], []
        while True:
            st, pred, lab = self.tokens[t][i]
            if pred is None:
                break
            il.append(lab); dst.append(st)
            t, i = pred
            src.append(START if self.tokens[t][i][1] is None else self.tokens[t][i][0])
        return (np.array(il[::-1], np.int32), np.array(dst[::-1], np.int32), np.array(src[::-1], np.int32))


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _delta(fn, h, ready, first):
    fr, w = np.zeros(1 << 14, np.int32), np.zeros(1 << 14, np.float32)
    n = fn(h, ready, first, _p(fr), _p(w), fr.size)
    return n, fr[:max(n, 0)].copy(), w[:max(n, 0)].copy()


def _pair(ref, mine, tid2phone, sil, weight, max_dur, fs):
    r = ref.ref_silw_create(MDL.encode(), sil.encode(), weight, max_dur, fs)
    m = mine.silw_create(_p(tid2phone), tid2phone.size, sil.encode(), weight, max_dur, fs)
    assert r and m
    return r, m


@pytest.mark.parametrize("fs", [1, 3])
@pytest.mark.parametrize("max_dur", [-1.0, 4.0, 9.5])
@pytest.mark.parametrize("weight", [0.0, 0.001])
def test_delta_weights_follow_the_reference_call_after_call(ref, mine, tid2phone, fs, max_dur, weight):
    rng = np.random.default_rng(1000 * fs + int(10 * max_dur) + int(1000 * weight) + 7)
    for trial in range(6):
        r, m = _pair(ref, mine, tid2phone, "1:2", weight, max_dur, fs)
        first = 0 if trial % 2 == 0 else int(rng.integers(1, 4)) * fs
        tr = Trellis(rng, tid2phone.size - 1, stay=0.93 if trial < 4 else 0.995)
        frames, ready, changed, quirk_seen = 0, 0, 0, 0
        last = None
        while frames < 420:
            n = int(rng.integers(1, 31))
            tr.grow(n)
            frames += n
            il, dst, src = tr.best_path()
            assert ref.ref_silw_traceback(r, _p(il), _p(src), il.size, frames) == 0
            assert mine.silw_traceback_from_path(m, _p(il), _p(dst), il.size, frames) == 0
            tids = il[il != 0]
            if last is not None:
                k = min(last.size, tids.size)
                changed += int((last[:k] != tids[:k]).sum())
            last = tids
            # the features run ahead of the decoder by some frames, never backwards
            ready = max(ready, first + fs * frames + int(rng.integers(0, 5 * fs)))
            if rng.random() < 0.85:
                a, b = _delta(ref.ref_silw_delta_weights, r, ready, first), _delta(mine.silw_delta_weights, m, ready, first)
                assert a[0] == b[0] and a[0] >= fs
                assert np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2])
                quirk_seen += int((a[2] < 0).any())
            if rng.random() < 0.3:
                fa, fb = np.zeros(1 << 12, np.int32), np.zeros(1 << 12, np.int32)
                na = ref.ref_silw_nonsilence_frames(r, ready, first, _p(fa), fa.size)
                nb = mine.silw_nonsilence_frames(m, ready, first, _p(fb), fb.size)
                assert na == nb and na >= 0 and np.array_equal(fa[:na], fb[:nb])
        assert changed > 0                      # the tracebacks did revise earlier frames ...
        if weight != 1.0:
            assert quirk_seen > 0               # ... and weights that were out already were taken back
        ref.ref_silw_destroy(r); mine.silw_destroy(m)


def test_no_traceback_yet_and_weights_before_any_decoding(ref, mine, tid2phone):
    r, m = _pair(ref, mine, tid2phone, "1", 0.25, -1.0, 3)
    # nothing decoded: every frame gets the silence weight
    for ready in (0, 7, 7, 40):
        a, b = _delta(ref.ref_silw_delta_weights, r, ready, 0), _delta(mine.silw_delta_weights, m, ready, 0)
        assert a[0] == b[0] and np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2])
    # 150 frames later still no traceback: the range starts 100 frames back and repeats the weight that went out
    a, b = _delta(ref.ref_silw_delta_weights, r, 3 * 200, 0), _delta(mine.silw_delta_weights, m, 3 * 200, 0)
    assert a[0] == b[0] > 0 and np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2])
    # the precondition num_frames_ready > first_decoder_frame (or 0): a KALDI_ASSERT (abort) there, an error here
    assert mine.silw_delta_weights(m, 3, 3, None, None, 0) == -1
    ref.ref_silw_destroy(r); mine.silw_destroy(m)


def test_frames_decoded_must_not_decrease(ref, mine, tid2phone):
    rng = np.random.default_rng(5)
    r, m = _pair(ref, mine, tid2phone, "1:2", 0.0, -1.0, 1)
    tr = Trellis(rng, tid2phone.size - 1)
    tr.grow(30)
    il, dst, src = tr.best_path()
    assert ref.ref_silw_traceback(r, _p(il), _p(src), il.size, 30) == 0
    assert mine.silw_traceback_from_path(m, _p(il), _p(dst), il.size, 30) == 0
    # a path of 20 frames afterwards: "Number of frames decoded decreased"
    keep = np.flatnonzero(il != 0)[20]
    assert ref.ref_silw_traceback(r, _p(il), _p(src), int(keep), 20) == -1
    assert mine.silw_traceback_from_path(m, _p(il), _p(dst), int(keep), 20) == -1
    # a path that does not hold one emitting arc per frame is refused before anything changes (b2k only: the reference asserts)
    assert mine.silw_traceback_from_path(m, _p(il), _p(dst), il.size, 31) == -3
    ref.ref_silw_destroy(r); mine.silw_destroy(m)


@pytest.mark.parametrize("sil,ok", [("1:2", True), ("1,2", True), ("3", True), (" 1: 2", True), ("1:2 ", False), ("1:2:", False), ("1::2", False),
                                    ("1:x", False), ("", True), ("99999999999", False)])
def test_silence_list_parsing_is_the_references(ref, mine, tid2phone, sil, ok):
    """SplitStringToIntegers' result is not checked by the reference: a list with one bad field is an EMPTY list (and the
    weighting still reports Active()).  Same here, visible through the weights; b2k also says so (SilencePhonesParsed)."""
    rng = np.random.default_rng(11)
    r, m = _pair(ref, mine, tid2phone, sil, 0.125, -1.0, 1)
    assert ref.ref_silw_active(r) == mine.silw_active(m) == (1 if sil else 0)
    assert mine.silw_list_ok(m) == (1 if ok else 0)
    tr = Trellis(rng, tid2phone.size - 1, repeat=0.3)
    tr.grow(90)
    il, dst, src = tr.best_path()
    assert ref.ref_silw_traceback(r, _p(il), _p(src), il.size, 90) == 0
    assert mine.silw_traceback_from_path(m, _p(il), _p(dst), il.size, 90) == 0
    a, b = _delta(ref.ref_silw_delta_weights, r, 90, 0), _delta(mine.silw_delta_weights, m, 90, 0)
    assert a[0] == b[0] and np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2])
    if not ok or not sil:
        assert np.all(a[2] == 1.0)              # nothing is silence
    ref.ref_silw_destroy(r); mine.silw_destroy(m)


def test_kaldi_typed_wrapper_beside_the_references_class(tmp_path, ref):
    """OnlineSilenceWeightingB2k (the Kaldi-typed shim) constructed from the same TransitionModel and OnlineSilenceWeightingConfig as
    the reference's class, in one C++ program (tests/cabi/silence_weighting_kaldi_test.cc) linked against oracle/_ref: 300 tracebacks,
    delta weights and non-silence frames equal call after call; a decoder that goes backwards raises KaldiFatalError in both.
    Needs the reference's headers: this container only."""
    if not os.path.isdir("/root/reference/src"):
        pytest.skip("/root/reference not present")
    so_b2k = os.path.join(ROOT, "kaldi_b200", "libb2k.so")
    if not os.path.exists(so_b2k):
        pytest.skip("libb2k.so not built")
    from oracle import nnet_oracle as NO, ref_feat as RF
    blas = os.path.dirname(RF.find_openblas())
    flags = RF.cxxflags(["-DHAVE_CUDA=0", "-I" + os.path.join(ROOT, "include"), "-I" + os.path.join(ROOT, "kaldi_b200", "host"),
                         "-I/usr/local/cuda/include", "-I" + os.path.join(ROOT, "oracle", "_ref", "inc"),
                         "-I" + os.path.join(ROOT, "oracle", "ref_wrap")])
    exe = str(tmp_path / "swk")
    r = subprocess.run(["g++"] + flags + [os.path.join(ROOT, "tests", "cabi", "silence_weighting_kaldi_test.cc"), "-o", exe, NO._SO,
                        "-Wl,-rpath," + os.path.dirname(NO._SO), "-Wl,-rpath-link," + blas, "-Wl,-rpath," + blas,
                        "-L" + os.path.dirname(so_b2k), "-lb2k", "-Wl,-rpath," + os.path.dirname(so_b2k), "-Wl,--allow-shlib-undefined"],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    r = subprocess.run([exe, MDL], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "silence weighting wrapper ok" in r.stdout, r.stdout + r.stderr[-2000:]
