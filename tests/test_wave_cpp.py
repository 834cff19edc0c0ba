"""RIFF/WAVE input through the C++ reader of libb2k.so (kaldi_b200/csrc/host_utils.cu: b2k_wave_read) against the
reference's own WaveData::Read (feat/wave-reader.cc, compiled in oracle/_ref) on the same files: plain PCM, stereo, extra
chunks, WAVE_FORMAT_EXTENSIBLE, RIFX, stream-mode sizes, truncated data, and the files both must reject."""
import ctypes as C
import struct

import numpy as np
import pytest


def _riff(fmt_body: bytes, data: bytes, pre_fmt=b"", pre_data=b"", big=False, riff_size=None, data_size=None, pad=b""):
    e = ">" if big else "<"
    body = b"WAVE" + pre_fmt + b"fmt " + struct.pack(e + "I", len(fmt_body)) + fmt_body + pre_data + b"data" + \
        struct.pack(e + "I", len(data) if data_size is None else data_size) + data + pad
    return (b"RIFX" if big else b"RIFF") + struct.pack(e + "I", len(body) if riff_size is None else riff_size) + body


def _fmt(channels=1, rate=16000, bits=16, fmt=1, big=False, byte_rate=None, block_align=None, extensible_guid=None):
    e = ">" if big else "<"
    br = rate * bits // 8 * channels if byte_rate is None else byte_rate
    ba = channels * bits // 8 if block_align is None else block_align
    out = struct.pack(e + "HHIIHH", fmt, channels, rate, br, ba, bits)
    if fmt == 0xFFFE:
        g = extensible_guid or (0x00000001, 0x00100000, 0xAA000080, 0x719B3800)
        out += struct.pack(e + "HHI", 22, bits, 3) + struct.pack(e + "IIII", *g)
    return out


def _chunk(tag, payload, big=False):
    return tag + struct.pack((">" if big else "<") + "I", len(payload)) + payload


def _samples(n, channels=1, seed=0, big=False):
    x = np.random.default_rng(seed).integers(-32768, 32768, (n, channels)).astype(">i2" if big else "<i2")
    return x, x.tobytes()


def _mine(path):
    try:
        from kaldi_b200 import _lib
        L = _lib.lib()
    except OSError as e:
        pytest.skip(str(e))
    h = C.c_void_p()
    L.b2k_wave_read.argtypes = [C.c_char_p, C.c_void_p]
    if L.b2k_wave_read(path.encode(), C.byref(h)) != 0:
        return None
    f, ch, n = C.c_float(), C.c_int32(), C.c_int64()
    L.b2k_wave_info.argtypes = [C.c_void_p] * 4
    L.b2k_wave_info(h, C.byref(f), C.byref(ch), C.byref(n))
    L.b2k_wave_data.restype = C.POINTER(C.c_float)
    L.b2k_wave_data.argtypes = [C.c_void_p]
    a = np.ctypeslib.as_array(L.b2k_wave_data(h), shape=(ch.value, n.value)).copy() if n.value else np.zeros((ch.value, 0), np.float32)
    L.b2k_wave_destroy.argtypes = [C.c_void_p]
    L.b2k_wave_destroy(h)
    return f.value, a


def _reference(path):
    from oracle import feat_oracle as F
    R = F.RefFeat()
    if not hasattr(R.lib, "ref_wave_read"):
        pytest.skip("oracle/_ref feature library predates ref_wave_read")
    out = np.zeros(1 << 20, np.float32)
    ch, n, f = C.c_int(), C.c_longlong(), C.c_float()
    R.lib.ref_wave_read.argtypes = [C.c_char_p, C.c_void_p, C.c_longlong, C.c_void_p, C.c_void_p, C.c_void_p]
    rc = R.lib.ref_wave_read(path.encode(), out.ctypes.data, out.size, C.byref(ch), C.byref(n), C.byref(f))
    if rc != 0:
        return None
    return f.value, out[:ch.value * n.value].reshape(ch.value, n.value).copy()


x1, b1 = _samples(1000)
x2, b2 = _samples(777, channels=2, seed=1)
xb, bb = _samples(300, channels=2, seed=2, big=True)
GOOD = {
    "mono": _riff(_fmt(), b1),
    "stereo-8k": _riff(_fmt(channels=2, rate=8000), b2),
    "junk-before-fmt": _riff(_fmt(), b1, pre_fmt=_chunk(b"JUNK", b"\0" * 28)),
    "fact-and-list-before-data": _riff(_fmt(), b1, pre_data=_chunk(b"fact", struct.pack("<I", 1000)) + _chunk(b"LIST", b"INFOxxxxyyyy")),
    "long-fmt-chunk": _riff(_fmt() + b"\0\0", b1),
    "extensible": _riff(_fmt(channels=2, fmt=0xFFFE), b2),
    "rifx": _riff(_fmt(channels=2, big=True), bb, big=True),
    "stream-riff0": _riff(_fmt(), b1, riff_size=0),
    "stream-data-ffffffff": _riff(_fmt(), b1, data_size=0xFFFFFFFF),
    "stream-sox": _riff(_fmt(), b1, data_size=0x7FFFF000),
    "truncated": _riff(_fmt(), b1[:1500], data_size=2000),
    "odd-trailing-byte": _riff(_fmt(channels=2), b2 + b"\x7f", data_size=len(b2) + 1),
    "padding-after-data": _riff(_fmt(), b1, pad=b"\0"),
    "one-sample": _riff(_fmt(), b1[:2]),
}
BAD = {
    "not-riff": b"RIFY" + _riff(_fmt(), b1)[4:],
    "not-wave": _riff(_fmt(), b1).replace(b"WAVE", b"WAVX", 1),
    "8-bit": _riff(_fmt(bits=8), b1),
    "float-format": _riff(_fmt(fmt=3), b1),
    "no-channels": _riff(_fmt(channels=0), b1),
    "bad-byte-rate": _riff(_fmt(byte_rate=12345), b1),
    "bad-block-align": _riff(_fmt(block_align=3), b1),
    "extensible-float": _riff(_fmt(fmt=0xFFFE, extensible_guid=(3, 0x00100000, 0xAA000080, 0x719B3800)), b1),
    "no-data": _riff(_fmt(), b""),
    "header-only": _riff(_fmt(), b1)[:30],
    "empty": b"",
}


@pytest.mark.parametrize("name", sorted(GOOD))
def test_reader_equals_the_references_on_valid_files(tmp_path, name):
    p = str(tmp_path / "a.wav")
    open(p, "wb").write(GOOD[name])
    ref, mine = _reference(p), _mine(p)
    assert ref is not None, "the reference itself rejects this file: fix the test"
    assert mine is not None
    assert mine[0] == ref[0] and mine[1].shape == ref[1].shape
    np.testing.assert_array_equal(mine[1], ref[1])
    if name == "mono":
        np.testing.assert_array_equal(mine[1][0], x1[:, 0].astype(np.float32))
    if name == "rifx":
        np.testing.assert_array_equal(mine[1], xb.astype(np.float32).T)


@pytest.mark.parametrize("name", sorted(BAD))
def test_both_reject_invalid_files(tmp_path, name):
    p = str(tmp_path / "a.wav")
    open(p, "wb").write(BAD[name])
    assert _reference(p) is None, "the reference accepts this file: fix the test"
    assert _mine(p) is None


@pytest.mark.parametrize("rin,rout,n", [(8000, 16000, 4000), (44100, 16000, 9000), (48000, 16000, 4801), (16000, 8000, 3000),
                                        (22050, 16000, 5000), (16000, 16000, 1000), (8000, 16000, 1), (44100, 16000, 2), (11025, 16000, 0)])
def test_resampler_equals_the_references(rin, rout, n):
    """ResampleWaveform (feat/resample.cc): same number of output samples, values within 1e-5 of the output scale (the
    reference's dot products are BLAS calls, ours a plain left-to-right loop)."""
    try:
        from kaldi_b200 import _lib
        L = _lib.lib()
    except OSError as e:
        pytest.skip(str(e))
    from oracle import feat_oracle as F
    R = F.RefFeat()
    if not hasattr(R.lib, "ref_resample_waveform"):
        pytest.skip("oracle/_ref feature library predates ref_resample_waveform")
    rng = np.random.default_rng(rin + rout + n)
    t = np.arange(n) / rin
    x = (3000 * np.sin(2 * np.pi * 440 * t) + 1000 * np.sin(2 * np.pi * 3100 * t) + rng.normal(0, 300, n)).astype(np.float32)
    ref = np.zeros(int(n * rout / rin) + 16, np.float32)
    R.lib.ref_resample_waveform.restype = C.c_longlong
    R.lib.ref_resample_waveform.argtypes = [C.c_float, C.c_void_p, C.c_longlong, C.c_float, C.c_void_p, C.c_longlong]
    nr = R.lib.ref_resample_waveform(float(rin), x.ctypes.data, n, float(rout), ref.ctypes.data, ref.size)
    assert 0 <= nr <= ref.size
    L.b2k_resample_waveform.argtypes = [C.c_float, C.c_void_p, C.c_int64, C.c_float, C.c_void_p, C.c_int64, C.c_void_p]
    no = C.c_int64()
    rc = L.b2k_resample_waveform(float(rin), x.ctypes.data, n, float(rout), None, 0, C.byref(no))
    assert no.value == nr and rc == (0 if nr == 0 else 4)      # size query: B2K_ERR_OVERFLOW with the size set
    out = np.zeros(max(nr, 1), np.float32)
    assert L.b2k_resample_waveform(float(rin), x.ctypes.data, n, float(rout), out.ctypes.data, out.size, C.byref(no)) == 0
    if nr:
        scale = max(1.0, float(np.abs(ref[:nr]).max()))
        assert np.abs(out[:nr] - ref[:nr]).max() <= 1e-5 * scale
    if rin == rout and n:
        np.testing.assert_allclose(out[:nr], x[:nr], rtol=0, atol=1e-5 * 4000 + 0.05 * 4000)   # an (almost) identity low-pass


def test_resampler_rejects_bad_arguments():
    try:
        from kaldi_b200 import _lib
        L = _lib.lib()
    except OSError as e:
        pytest.skip(str(e))
    L.b2k_resample_waveform.argtypes = [C.c_float, C.c_void_p, C.c_int64, C.c_float, C.c_void_p, C.c_int64, C.c_void_p]
    x = np.zeros(10, np.float32)
    n = C.c_int64()
    assert L.b2k_resample_waveform(0.0, x.ctypes.data, 10, 16000.0, None, 0, C.byref(n)) == 1
    assert L.b2k_resample_waveform(8000.5, x.ctypes.data, 10, 16000.0, None, 0, C.byref(n)) == 1
    assert L.b2k_resample_waveform(8000.0, None, 10, 16000.0, None, 0, C.byref(n)) == 1


def test_channel_count_is_read_as_the_reference_reads_it(tmp_path):
    """WaveInfo keeps the channel count in 8 bits (wave-reader.h:101): a header that says 258 channels with the byte rate and block size
    of 2 is a two-channel file to the reference, and one that says 256 has no channels.  Found by a soak run against WaveData::Read."""
    x, b = _samples(40, channels=2, seed=3)
    p = str(tmp_path / "c258.wav")
    open(p, "wb").write(_riff(_fmt(channels=258, byte_rate=16000 * 4, block_align=4), b))
    mine, ref = _mine(p), _reference(p)
    assert mine is not None and ref is not None
    assert mine[0] == ref[0] and mine[1].shape == ref[1].shape == (2, 40) and np.array_equal(mine[1], ref[1])
    open(p, "wb").write(_riff(_fmt(channels=256, byte_rate=0, block_align=0), b))
    assert _mine(p) is None and _reference(p) is None
    # less than one sample block of data: the reference asserts (it cannot be called on it), this reader reports an error
    open(p, "wb").write(_riff(_fmt(channels=2), b[:3], data_size=3))
    assert _mine(p) is None
