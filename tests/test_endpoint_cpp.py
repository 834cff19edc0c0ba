"""Endpointing through the C ABI (kaldi_b200/csrc/host_utils.cu: b2k_endpoint_*, b2k_trailing_silence_frames) against the
reference's own online2/online-endpoint.cc compiled in oracle/_ref (oracle/ref_wrap/endpoint_wrap.cc: the rules, the option
group and TrailingSilenceLength run unmodified; the decoder they query is a replay of a given best path)."""
import ctypes as C
import itertools
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MDL = os.path.join(ROOT, "tests", "golden", "tiny_final.mdl")          # 5 phones, written by the reference (make_model_golden.py)


class Rule(C.Structure):
    _fields_ = [("must_contain_nonsilence", C.c_int32), ("min_trailing_silence", C.c_float), ("max_relative_cost", C.c_float),
                ("min_utterance_length", C.c_float)]


class Cfg(C.Structure):
    _fields_ = [("rule", Rule * 5), ("silence_phones", C.c_char * 512)]


def _lib():
    try:
        from kaldi_b200 import _lib as L
        return L.lib()
    except OSError as e:
        pytest.skip(str(e))


def _ref():
    from oracle import nnet_oracle as NO
    if os.path.isdir("/root/reference"):
        from oracle import ref_nnet
        ref_nnet.build(quiet=True)
    if not os.path.exists(NO._SO):
        pytest.skip("oracle/_ref/libkaldi_ref_nnet3.so not built")
    R = C.CDLL(NO._SO)
    if not hasattr(R, "ref_endpoint_detected"):
        pytest.skip("oracle/_ref predates endpoint_wrap.cc")
    return R


def _err(L):
    L.b2k_last_error.restype = C.c_char_p
    return L.b2k_last_error().decode()


def _flat(c):
    return [v for r in c.rule for v in (float(r.must_contain_nonsilence), r.min_trailing_silence, r.max_relative_cost, r.min_utterance_length)]


CONFS = {
    "empty": "",
    "recipe": "--endpoint.silence-phones=1:2:3:4:5\n",
    "rules": """# a tuned rule set
--endpoint.silence-phones=1:2
--endpoint.rule1.min-trailing-silence=3.5
--endpoint.rule2.must-contain-nonsilence=false
--endpoint.rule2.max-relative-cost=2.5   # comment
--endpoint.rule3.min-trailing-silence=0.75
--endpoint.rule3.max_relative_cost=inf
--endpoint.rule4.min-utterance-length=1.25
--endpoint.rule5.min-utterance-length=12
--endpoint.rule5.must-contain-nonsilence
""",
}


@pytest.mark.parametrize("which", sorted(CONFS))
def test_option_group_equals_the_references_registration(tmp_path, which):
    L, R = _lib(), _ref()
    p = tmp_path / "endpoint.conf"
    p.write_text(CONFS[which])
    c = Cfg()
    assert L.b2k_endpoint_cfg_from_conf(str(p).encode(), C.byref(c)) == 0, _err(L)
    out = (C.c_float * 20)()
    sil = C.create_string_buffer(512)
    assert R.ref_endpoint_config(CONFS[which].encode(), out, sil, 512) == 0
    assert _flat(c) == list(out)
    assert c.silence_phones == sil.value
    # the text form on top of the defaults gives the same
    d = Cfg()
    assert L.b2k_endpoint_cfg_default(C.byref(d)) == 0
    text = " ".join(l.split("#")[0].strip() for l in CONFS[which].splitlines() if l.split("#")[0].strip())
    assert L.b2k_endpoint_cfg_apply_options(text.encode(), C.byref(d)) == 0, _err(L)
    assert _flat(d) == list(out) and d.silence_phones == sil.value


def test_other_option_groups_pass_and_unknown_endpoint_names_do_not(tmp_path):
    L, R = _lib(), _ref()
    p = tmp_path / "online.conf"
    p.write_text("--feature-type=mfcc\n--endpoint.silence-phones=3\n--beam=13\n")
    c = Cfg()
    assert L.b2k_endpoint_cfg_from_conf(str(p).encode(), C.byref(c)) == 0 and c.silence_phones == b"3"
    for bad in ("--endpoint.rule6.min-trailing-silence=1\n", "--endpoint.rule1.min-trailing-silence=abc\n",
                "--endpoint.rule1.must-contain-nonsilence=maybe\n", "--endpoint.silence=1\n"):
        p.write_text(bad)
        keep = _flat(c)
        assert L.b2k_endpoint_cfg_from_conf(str(p).encode(), C.byref(c)) != 0
        assert _flat(c) == keep                                          # untouched on failure
        out = (C.c_float * 20)()
        assert R.ref_endpoint_config(bad.encode(), out, C.create_string_buffer(8), 8) == -1       # the reference rejects them too
    assert L.b2k_endpoint_cfg_from_conf(str(tmp_path / "absent.conf").encode(), C.byref(c)) != 0
    assert L.b2k_endpoint_cfg_from_conf(None, C.byref(c)) != 0 and L.b2k_endpoint_cfg_default(None) != 0


def test_rules_equal_the_reference_on_a_grid(tmp_path):
    """Every combination of (frames, trailing silence, relative cost) around each rule's thresholds, for three rule sets."""
    L, R = _lib(), _ref()
    L.b2k_endpoint_detected.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_float, C.c_float, C.c_void_p]
    R.ref_endpoint_detected.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_float, C.c_float]
    frames = [0, 1, 49, 50, 51, 99, 100, 101, 200, 499, 500, 501, 1199, 1200, 1999, 2000, 2001, 3000]
    costs = [0.0, 1.999, 2.0, 2.001, 2.5, 7.99, 8.0, 8.01, 1e30, float("inf")]
    n = fired = 0
    for which, shift in itertools.product(sorted(CONFS), (0.01, 0.03)):
        p = tmp_path / "e.conf"
        p.write_text(CONFS[which])
        c = Cfg()
        assert L.b2k_endpoint_cfg_from_conf(str(p).encode(), C.byref(c)) == 0
        for f in frames:
            for s in sorted({0, 1, f // 3, f - 1, f, 16, 17, 25, 33, 34, 50, 66, 67, 116, 117, 166, 167, 350, 500} & set(range(f + 1))):
                for cost in costs:
                    got = C.c_int32(-7)
                    assert L.b2k_endpoint_detected(C.byref(c), f, s, shift, cost, C.byref(got)) == 0
                    ref = R.ref_endpoint_detected(CONFS[which].encode(), f, s, shift, cost)
                    assert ref in (0, 1) and got.value == ref, (which, shift, f, s, cost)
                    n += 1
                    fired += ref
    assert n > 5000 and 0.1 < fired / n < 0.9                            # both outcomes are well represented
    got = C.c_int32(-7)
    assert L.b2k_endpoint_detected(C.byref(c), 10, 11, 0.01, 0.0, C.byref(got)) != 0 and "trailing" in _err(L)       # the reference asserts


def _model(L):
    h = C.c_void_p()
    L.b2k_model_read.argtypes = [C.c_char_p, C.c_int32, C.c_void_p]
    assert L.b2k_model_read(MDL.encode(), 1, C.byref(h)) == 0, _err(L)
    info = (C.c_int32 * 8)()
    L.b2k_model_info.argtypes = [C.c_void_p, C.c_void_p]
    assert L.b2k_model_info(h, info) == 0
    L.b2k_model_tid2phone.restype = C.POINTER(C.c_int32)
    L.b2k_model_tid2phone.argtypes = [C.c_void_p]
    return h, np.ctypeslib.as_array(L.b2k_model_tid2phone(h), shape=(info[6],)).copy()


def test_tid2phone_equals_the_references_transition_model():
    L, R = _lib(), _ref()
    h, mine = _model(L)
    ref = np.zeros(64, np.int32)
    n = R.ref_tid2phone(MDL.encode(), ref.ctypes.data_as(C.c_void_p), 64)
    assert n == len(mine) - 1 and np.array_equal(mine, ref[:n + 1]) and set(mine[1:]) == {1, 2, 3, 4, 5}
    L.b2k_model_destroy.argtypes = [C.c_void_p]
    L.b2k_model_destroy(h)
    assert L.b2k_model_tid2phone(None) is None or not L.b2k_model_tid2phone(None)


def test_trailing_silence_and_decoder_form_equal_the_reference(tmp_path):
    L, R = _lib(), _ref()
    h, t2p = _model(L)
    T = len(t2p)
    L.b2k_trailing_silence_frames.argtypes = [C.c_void_p, C.c_int32, C.c_char_p, C.c_void_p, C.c_int64, C.c_void_p]
    L.b2k_endpoint_detected_on_path.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_int64, C.c_int32, C.c_float, C.c_float,
                                                C.c_void_p, C.c_void_p]
    R.ref_endpoint_detected_on_path.argtypes = [C.c_char_p, C.c_char_p, C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_float, C.c_void_p]
    rng = np.random.default_rng(5)
    sil_tids = [t for t in range(1, T) if t2p[t] in (1, 2)]
    sp_tids = [t for t in range(1, T) if t2p[t] not in (1, 2)]
    conf = CONFS["rules"]                                                # silence phones 1:2
    p = tmp_path / "e.conf"
    p.write_text(conf)
    c = Cfg()
    assert L.b2k_endpoint_cfg_from_conf(str(p).encode(), C.byref(c)) == 0
    cases = [np.zeros(0, np.int32), np.zeros(7, np.int32), np.array(sil_tids * 3, np.int32), np.array(sp_tids * 3, np.int32)]
    for _ in range(200):
        n_sp, n_sil = int(rng.integers(0, 400)), int(rng.integers(0, 400))
        path = np.concatenate([rng.choice(sil_tids, int(rng.integers(0, 30))), rng.choice(sp_tids, n_sp), rng.choice(sil_tids, n_sil)])
        eps = rng.random(len(path) + 40) < 0.15                           # epsilon arcs anywhere, as a raw best path has them
        full = np.zeros(len(eps), np.int32)
        idx = np.flatnonzero(~eps)[:len(path)]
        full[idx] = path[:len(idx)]
        cases.append(full)
    fired = 0
    for full in cases:
        frames = int((full != 0).sum())
        for cost in (0.0, 2.2, 9.0, float("inf")):
            got, trail, rtrail = C.c_int32(-7), C.c_int32(-7), C.c_int32(-7)
            assert L.b2k_endpoint_detected_on_path(C.byref(c), t2p.ctypes.data, T, full.ctypes.data, len(full), frames, 0.03, cost,
                                                   C.byref(got), C.byref(trail)) == 0, _err(L)
            ref = R.ref_endpoint_detected_on_path(conf.encode(), MDL.encode(), full.ctypes.data, len(full), frames, 0.03, cost, C.byref(rtrail))
            assert ref in (0, 1) and (got.value, trail.value) == (ref, rtrail.value)
            fired += ref
        alone = C.c_int32(-7)
        assert L.b2k_trailing_silence_frames(t2p.ctypes.data, T, b"1:2", full.ctypes.data, len(full), C.byref(alone)) == 0
        assert alone.value == (rtrail.value if frames else 0)
    assert 0 < fired < 4 * len(cases)
    # bad silence lists and labels: errors here, KALDI_ERR / KALDI_ASSERT there
    one = np.array([1], np.int32)
    out = C.c_int32()
    for bad in (b"", b"1::2", b"1:2:", b"1:x", b"2:2", b"1 :2", b"99999999999"):
        assert L.b2k_trailing_silence_frames(t2p.ctypes.data, T, bad, one.ctypes.data, 1, C.byref(out)) != 0, bad
    for ok in (b"1: 2", b"5", b"+1:-3"):
        assert L.b2k_trailing_silence_frames(t2p.ctypes.data, T, ok, one.ctypes.data, 1, C.byref(out)) == 0, ok
    for lab in (T, -1, 1 << 30):
        bad = np.array([lab], np.int32)
        assert L.b2k_trailing_silence_frames(t2p.ctypes.data, T, b"1", bad.ctypes.data, 1, C.byref(out)) != 0
    assert L.b2k_trailing_silence_frames(None, T, b"1", one.ctypes.data, 1, C.byref(out)) != 0
    L.b2k_model_destroy.argtypes = [C.c_void_p]
    L.b2k_model_destroy(h)


def test_python_mirror_reads_like_the_reference_interface(tmp_path):
    _lib()
    from kaldi_b200 import _lib as LB
    from kaldi_b200.endpoint import EndpointDetected, OnlineEndpointConfig, TrailingSilenceLength
    from kaldi_b200.model import KaldiModel
    m = KaldiModel(MDL)
    assert list(m.tid2phone) == [0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5]
    cfg = OnlineEndpointConfig()
    assert cfg.silence_phones == "" and [r.min_trailing_silence for r in cfg.rules] == [5.0, 0.5, 1.0, 2.0, 0.0]
    cfg.apply_options("--endpoint.silence-phones=1 --endpoint.rule2.min-trailing-silence=0.3")
    path = [5, 6, 0, 9, 10, 1, 0, 2, 2, 1] + [0]                         # three non-silence frames, four of phone 1
    assert TrailingSilenceLength(m.tid2phone, cfg.silence_phones, path) == 4
    assert not EndpointDetected(cfg, m.tid2phone, path, 9, 0.03, 0.0)    # 0.12 s of silence
    assert EndpointDetected(cfg, m.tid2phone, path[:5] + [1] * 11, 16, 0.03, 0.0)       # 0.33 s: rule2 (10 frames are 0.29999998 s in float, as in the reference)
    assert not EndpointDetected(cfg, m.tid2phone, path[:5] + [1] * 11, 16, 0.03, 2.5)   # relative cost too high for rule2
    assert EndpointDetected(cfg, 700, 0, 0.03, float("inf")) and not EndpointDetected(cfg, 600, 0, 0.03, float("inf"))   # rule5: 20 s
    p = tmp_path / "c.conf"
    p.write_text("--endpoint.rule5.min-utterance-length=10\n")
    assert EndpointDetected(OnlineEndpointConfig.from_conf(p), 600, 0, 0.03, float("inf"))
    with pytest.raises(LB.B2kError):
        TrailingSilenceLength(m.tid2phone, "", [1])
    with pytest.raises(TypeError):
        EndpointDetected(cfg, 1, 2)
