"""Feature-stage oracle — TEST INFRASTRUCTURE ONLY (tests/, smoke(), bench.py's
cpu_baseline / --impl reference legs).

Two checkers:

* `RefFeat`: ctypes view of oracle/_ref/libkaldi_ref_feat.so = the reference's
  OWN feature sources compiled where they lie (oracle/ref_feat.py).  This is the
  real reference; it also travels to the GPU box (same image => same OpenBLAS).
* numpy restatement (`mfcc_fbank`, `online_cmvn`) of the same algorithm, each
  function citing the file:line it follows.  It is pinned (tests/test_feat_
  oracle.py) against (a) RefFeat, (b) the reference's HTK golden vectors
  src/feat/test_data/test.wav.fea_htk.* through the committed fixtures in
  tests/golden/ (tolerance 1.0 abs as feature-mfcc-test.cc:163 does), and (c)
  online == offline (online-feature-test.cc:146-196).

PARITY STATUS: pinned (reference itself compiled + the reference's own golden
vectors).  Float association of BLAS dot/gemv is unspecified, so comparisons
are tolerance based (written in each test).
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass, asdict

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_REF_SO = os.path.join(_HERE, "_ref", "libkaldi_ref_feat.so")

F32 = np.float32
FLT_EPS = np.finfo(np.float32).eps


@dataclass
class FeatOpts:
    """Union of MfccOptions / FbankOptions / FrameExtractionOptions /
    MelBanksOptions (feat/feature-mfcc.h:38-60, feature-fbank.h:41-60,
    feature-window.h:38-66, mel-computations.h:43-58).  Defaults = the
    mfcc_hires.conf of the named recipes with dither off."""
    feature_type: int = 0          # 0 mfcc, 1 fbank
    samp_freq: float = 16000.0
    frame_shift_ms: float = 10.0
    frame_length_ms: float = 25.0
    dither: float = 0.0
    preemph_coeff: float = 0.97
    remove_dc_offset: int = 1
    round_to_power_of_two: int = 1
    snip_edges: int = 1
    num_bins: int = 40
    low_freq: float = 20.0
    high_freq: float = -400.0
    num_ceps: int = 40
    use_energy: int = 0
    energy_floor: float = 0.0
    raw_energy: int = 1
    cepstral_lifter: float = 22.0
    htk_compat: int = 0
    use_log_fbank: int = 1
    use_power: int = 1
    window_type: int = 0           # 0 povey, 1 hamming, 2 hanning, 3 rectangular
    htk_mode: int = 0              # MelBanksOptions::htk_mode (mel-computations.h:52-55)
    lpc_order: int = 12            # PlpOptions (feature_type 2), feat/feature-plp.h:38-66
    compress_factor: float = 0.33333
    cepstral_scale: float = 1.0

    @property
    def window_shift(self):       # WindowShift  feature-window.h:106
        return int(self.samp_freq * 0.001 * self.frame_shift_ms)

    @property
    def window_size(self):        # WindowSize   feature-window.h:109
        return int(self.samp_freq * 0.001 * self.frame_length_ms)

    @property
    def padded_window_size(self):  # PaddedWindowSize feature-window.h:112
        n = self.window_size
        if not self.round_to_power_of_two:
            return n
        p = 1
        while p < n:
            p *= 2
        return p

    @property
    def dim(self):
        if self.feature_type in (0, 2):
            return self.num_ceps
        return self.num_bins + (1 if self.use_energy else 0)


class _COpts(C.Structure):
    _fields_ = [("feature_type", C.c_int), ("samp_freq", C.c_float), ("frame_shift_ms", C.c_float),
                ("frame_length_ms", C.c_float), ("dither", C.c_float), ("preemph_coeff", C.c_float),
                ("remove_dc_offset", C.c_int), ("round_to_power_of_two", C.c_int), ("snip_edges", C.c_int),
                ("num_bins", C.c_int), ("low_freq", C.c_float), ("high_freq", C.c_float),
                ("num_ceps", C.c_int), ("use_energy", C.c_int), ("energy_floor", C.c_float),
                ("raw_energy", C.c_int), ("cepstral_lifter", C.c_float), ("htk_compat", C.c_int),
                ("use_log_fbank", C.c_int), ("use_power", C.c_int),
                ("window_type", C.c_int), ("htk_mode", C.c_int),
                ("lpc_order", C.c_int), ("compress_factor", C.c_float), ("cepstral_scale", C.c_float)]


def _copts(o: FeatOpts) -> _COpts:
    return _COpts(**asdict(o))


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


class RefFeat:
    """The compiled reference (oracle/_ref)."""

    def __init__(self):
        if not os.path.exists(_REF_SO):
            from . import ref_feat
            ref_feat.build()
        self.lib = C.CDLL(_REF_SO)

    @staticmethod
    def available() -> bool:
        return os.path.exists(_REF_SO) or os.path.isdir("/root/reference/src")

    def compute(self, wave: np.ndarray, o: FeatOpts, online_chunk: int | None = None) -> np.ndarray:
        w = np.ascontiguousarray(wave, dtype=np.float32)
        max_rows = w.size // max(o.window_shift, 1) + 8
        out = np.zeros((max_rows, o.dim), np.float32)
        dim = C.c_int()
        co = _copts(o)
        if online_chunk is None:
            n = self.lib.ref_feat_compute(C.byref(co), _p(w, C.c_float), w.size, _p(out, C.c_float), max_rows, C.byref(dim))
        else:
            n = self.lib.ref_feat_online(C.byref(co), _p(w, C.c_float), w.size, int(online_chunk),
                                         _p(out, C.c_float), max_rows, C.byref(dim))
        if n < 0:
            raise RuntimeError(f"reference feature computation failed ({n})")
        assert dim.value == o.dim
        return out[:n].copy()

    def online_cmvn(self, feats: np.ndarray, cmn_window=600, speaker_frames=600, global_frames=200,
                    normalize_mean=True, normalize_variance=False, global_stats=None, order=None) -> np.ndarray:
        f = np.ascontiguousarray(feats, np.float32)
        T, D = f.shape
        n = T if order is None else len(order)
        out = np.zeros((n, D), np.float32)
        gs = None if global_stats is None else np.ascontiguousarray(global_stats, np.float64)
        od = None if order is None else np.ascontiguousarray(order, np.int32)
        r = self.lib.ref_cmvn_online(_p(f, C.c_float), T, D, cmn_window, speaker_frames, global_frames,
                                     int(normalize_mean), int(normalize_variance),
                                     None if gs is None else _p(gs, C.c_double),
                                     None if od is None else _p(od, C.c_int), n, _p(out, C.c_float))
        if r < 0:
            raise RuntimeError("reference OnlineCmvn failed")
        return out


# ----------------------------------------------------------------------------- restatement

def num_frames(num_samples: int, o: FeatOpts, flush: bool = True) -> int:
    """NumFrames, feat/feature-window.cc:42-87."""
    shift, length = o.window_shift, o.window_size
    if o.snip_edges:
        return 0 if num_samples < length else 1 + (num_samples - length) // shift
    n = (num_samples + shift // 2) // shift
    if flush:
        return n
    end = first_sample_of_frame(n - 1, o) + length
    while n > 0 and end > num_samples:
        n -= 1
        end -= shift
    return n


def first_sample_of_frame(frame: int, o: FeatOpts) -> int:
    """FirstSampleOfFrame, feat/feature-window.cc:30-40."""
    shift = o.window_shift
    if o.snip_edges:
        return frame * shift
    return shift * frame + shift // 2 - o.window_size // 2


def window_function(o: FeatOpts) -> np.ndarray:
    """FeatureWindowFunction, feat/feature-window.cc:109-135: double math, stored f32."""
    n = o.window_size
    a = 2.0 * np.pi / (n - 1)
    i = np.arange(n, dtype=np.float64)
    if o.window_type == 1:
        return (0.54 - 0.46 * np.cos(a * i)).astype(np.float32)
    if o.window_type == 2:
        return (0.5 - 0.5 * np.cos(a * i)).astype(np.float32)
    if o.window_type == 3:
        return np.ones(n, np.float32)
    return np.power(0.5 - 0.5 * np.cos(a * i), 0.85).astype(np.float32)


def mel_scale(f):
    """MelScale, feat/mel-computations.h:85: 1127.0f * logf(1.0f + f / 700.0f)."""
    f = np.asarray(f, dtype=np.float32)
    return (F32(1127.0) * np.log(F32(1.0) + f / F32(700.0))).astype(np.float32)


def mel_banks(o: FeatOpts):
    """MelBanks::MelBanks (vtln_warp = 1), feat/mel-computations.cc:33-142.
    Returns [(first_index, weights f32)] per bin; all arithmetic in float32."""
    nfft_bins = o.padded_window_size // 2
    nyquist = F32(0.5) * F32(o.samp_freq)
    low = F32(o.low_freq)
    high = F32(o.high_freq) if o.high_freq > 0.0 else F32(nyquist + F32(o.high_freq))
    fft_bin_width = F32(o.samp_freq) / F32(o.padded_window_size)
    mel_low, mel_high = mel_scale(low), mel_scale(high)
    delta = F32((mel_high - mel_low) / F32(o.num_bins + 1))
    mel = mel_scale(fft_bin_width * np.arange(nfft_bins, dtype=np.float32))
    bins = []
    for b in range(o.num_bins):
        left = F32(mel_low + F32(b) * delta)
        center = F32(mel_low + F32(b + 1) * delta)
        right = F32(mel_low + F32(b + 2) * delta)
        idx = np.nonzero((mel > left) & (mel < right))[0]
        first, last = int(idx[0]), int(idx[-1])
        m = mel[first:last + 1]
        up = ((m - left) / F32(center - left)).astype(np.float32)
        down = ((right - m) / F32(right - center)).astype(np.float32)
        w = np.where(m <= center, up, down).astype(np.float32)
        if o.htk_mode and b == 0 and mel_low != 0.0:          # :133-134 (replicates an HTK bug)
            w[0] = 0.0
        bins.append((first, w))
    return bins


def dct_matrix(o: FeatOpts) -> np.ndarray:
    """ComputeDctMatrix rows [0, num_ceps), matrix/matrix-functions.cc:592-608 (Real = float)."""
    N = o.num_bins
    M = np.zeros((N, N), np.float32)
    M[0, :] = F32(np.sqrt(F32(1.0) / F32(N)))
    norm = F32(np.sqrt(F32(2.0) / F32(N)))
    n = np.arange(N, dtype=np.float64)
    for k in range(1, N):
        M[k, :] = (F32(norm) * np.cos(np.pi / N * (n + 0.5) * k)).astype(np.float32)
    return M[:o.num_ceps].copy()


def lifter_coeffs(o: FeatOpts) -> np.ndarray:
    """ComputeLifterCoeffs, feat/mel-computations.cc:253-259."""
    Q = float(F32(o.cepstral_lifter))
    i = np.arange(o.num_ceps, dtype=np.float64)
    return (1.0 + 0.5 * Q * np.sin(np.pi * i / Q)).astype(np.float32)


def extract_frames(wave: np.ndarray, o: FeatOpts) -> np.ndarray:
    """ExtractWindow's gather with reflection, feat/feature-window.cc:166-214 (offline: sample_offset 0)."""
    n = len(wave)
    T = num_frames(n, o, True)
    L = o.window_size
    idx = np.array([first_sample_of_frame(f, o) for f in range(T)], dtype=np.int64)[:, None] + np.arange(L)[None, :]
    for _ in range(4):   # repeated reflection (:202-211)
        idx = np.where(idx < 0, -idx - 1, idx)
        idx = np.where(idx >= n, 2 * n - 1 - idx, idx)
    return np.asarray(wave, np.float32)[idx]


def plp_tables(o: FeatOpts):
    """Equal-loudness weights at the bands' centre frequencies (GetEqualLoudnessVector, mel-computations.cc:301-313; centres as
    MelBanks::MelBanks leaves them, :99-102) and the inverse-DFT bases (InitIdftBases, feature-functions.cc:188-203), float
    arithmetic where the reference's is."""
    nyq = F32(0.5) * F32(o.samp_freq)
    high = F32(o.high_freq) if o.high_freq > 0.0 else nyq + F32(o.high_freq)

    def mel_scale(f):
        return F32(1127.0) * np.log(F32(1.0) + F32(f) / F32(700.0)).astype(np.float32)
    mel_low, mel_high = mel_scale(o.low_freq), mel_scale(high)
    delta = F32((mel_high - mel_low) / F32(o.num_bins + 1))
    eql = np.zeros(o.num_bins, np.float32)
    for b in range(o.num_bins):
        center_mel = F32(mel_low + F32(b + 1) * delta)
        f0 = F32(F32(700.0) * (np.exp(F32(center_mel / F32(1127.0)), dtype=np.float32) - F32(1.0)))
        fsq = F32(f0 * f0)
        fsub = F32(float(fsq) / (float(fsq) + 1.6e5))
        eql[b] = F32(float(fsub) * float(fsub) * ((float(fsq) + 1.44e6) / (float(fsq) + 9.61e6)))
    nb, dm = o.lpc_order + 1, o.num_bins + 2
    angle = F32(np.pi / float(F32(dm - 1)))
    scale = F32(1.0 / (2.0 * float(F32(dm - 1))))
    idft = np.zeros((nb, dm), np.float32)
    for i in range(nb):
        idft[i, 0] = F32(1.0 * float(scale))
        for j in range(1, dm - 1):
            idft[i, j] = F32(2.0 * float(scale) * np.cos(float(F32(F32(angle * F32(i)) * F32(j)))))
        idft[i, dm - 1] = F32(float(scale) * np.cos(float(F32(F32(angle * F32(i)) * F32(dm - 1)))))
    return eql, idft


def _plp_tail(mel: np.ndarray, raw_log_energy, o: FeatOpts) -> np.ndarray:
    """PlpComputer::Compute after the mel bank (feature-plp.cc:140-182): equal loudness, compression, autocorrelation by the
    inverse-DFT bases, Durbin's recursion (mel-computations.cc:266-297), LPC -> cepstrum (:300-309), lifter, scale, C0 / energy."""
    eql, idft = plp_tables(o)
    T = mel.shape[0]
    comp = np.power((mel * eql[None, :]).astype(np.float32), F32(o.compress_factor)).astype(np.float32)
    dup = np.concatenate([comp[:, :1], comp, comp[:, -1:]], 1)
    ac = (dup @ idft.T).astype(np.float32)
    LO = o.lpc_order
    out = np.zeros((T, o.num_ceps), np.float32)
    lift = lifter_coeffs(o) if o.cepstral_lifter != 0.0 else np.ones(o.num_ceps, np.float32)
    for t in range(T):
        a = ac[t]
        lpc = np.zeros(LO, np.float32); tmp = np.zeros(LO, np.float32)
        E = F32(a[0])
        for i in range(LO):
            ki = F32(a[i + 1])
            for j in range(i):
                ki = F32(ki + F32(lpc[j] * a[i - j]))
            ki = F32(ki / E)
            c = F32(F32(1.0) - F32(ki * ki))
            if c < F32(1.0e-5):
                c = F32(1.0e-5)
            E = F32(E * c)
            tmp[i] = -ki
            for j in range(i):
                tmp[j] = F32(lpc[j] - F32(ki * lpc[i - j - 1]))
            lpc[:i + 1] = tmp[:i + 1]
        res = max(F32(-np.log(F32(F32(1.0) / E))), np.finfo(np.float32).tiny)
        cep = np.zeros(LO, np.float32)
        for i in range(LO):
            sm = 0.0
            for j in range(i):
                sm += float(F32(F32(F32(i - j) * lpc[j]) * cep[i - j - 1]))
            cep[i] = F32(float(-lpc[i]) - sm / float(F32(i + 1)))
        out[t, 0] = res
        out[t, 1:] = cep[:o.num_ceps - 1]
    out = (out * lift[None, :]).astype(np.float32)
    if o.cepstral_scale != 1.0:
        out = (out * F32(o.cepstral_scale)).astype(np.float32)
    if o.use_energy:
        e = raw_log_energy
        if o.energy_floor > 0.0:
            e = np.maximum(e, F32(np.log(F32(o.energy_floor))))
        out[:, 0] = e
    if o.htk_compat:
        out = np.concatenate([out[:, 1:], out[:, :1]], 1)
    return out


def mfcc_fbank(wave: np.ndarray, o: FeatOpts) -> np.ndarray:
    """Mfcc/Fbank::ComputeFeatures = ExtractWindow + ProcessWindow
    (feature-window.cc:137-160) + MfccComputer::Compute (feature-mfcc.cc:28-80)
    or FbankComputer::Compute (feature-fbank.cc:72-123).  dither must be 0."""
    assert o.dither == 0.0, "parity runs use --dither=0 (SURVEY.md §7 hard part 6)"
    fr = extract_frames(wave, o).astype(np.float32)          # [T, L]
    T, L = fr.shape
    P = o.padded_window_size
    if o.remove_dc_offset:                                    # window->Add(-window->Sum() / frame_length)
        s = fr.sum(axis=1, dtype=np.float32)
        fr = (fr + (-s / F32(L))[:, None]).astype(np.float32)
    raw_log_energy = None
    need_raw = bool(o.use_energy and o.raw_energy)
    if need_raw:
        e = np.maximum((fr * fr).sum(axis=1, dtype=np.float32), FLT_EPS)
        raw_log_energy = np.log(e).astype(np.float32)
    if o.preemph_coeff != 0.0:                                # Preemphasize :100-107
        c = F32(o.preemph_coeff)
        out = fr.copy()
        out[:, 1:] = fr[:, 1:] - c * fr[:, :-1]
        out[:, 0] = fr[:, 0] - c * fr[:, 0]
        fr = out.astype(np.float32)
    fr = (fr * window_function(o)[None, :]).astype(np.float32)
    if o.use_energy and not o.raw_energy:
        e = np.maximum((fr * fr).sum(axis=1, dtype=np.float32), FLT_EPS)
        raw_log_energy = np.log(e).astype(np.float32)
    pad = np.zeros((T, P), np.float32)
    pad[:, :L] = fr
    # SplitRadixRealFft::Compute (matrix/srfft.cc:355-432) computes the forward
    # real DFT; its packing [re0, reN/2, re1, im1, ...] is only a layout.  The
    # restatement uses a double-precision DFT rounded to f32 (rounding of the
    # split-radix butterflies is reproduced to ~1e-6 relative, not bit-exact).
    spec = np.fft.rfft(pad.astype(np.float64), axis=1)
    re, im = spec.real.astype(np.float32), spec.imag.astype(np.float32)
    power = (re * re + im * im).astype(np.float32)            # ComputePowerSpectrum feature-functions.cc:29-51
    if o.feature_type == 1 and not o.use_power:
        power = np.sqrt(power).astype(np.float32)             # ApplyPow(0.5)
    bins = mel_banks(o)
    mel = np.zeros((T, o.num_bins), np.float32)
    for b, (first, w) in enumerate(bins):                     # MelBanks::Compute mel-computations.cc:226-251
        mel[:, b] = (power[:, first:first + len(w)] * w[None, :]).sum(axis=1, dtype=np.float32)
    if o.htk_mode:
        mel = np.maximum(mel, F32(1.0))                       # :237 HTK-like flooring
    if o.feature_type == 2:
        return _plp_tail(mel, raw_log_energy, o)
    if o.feature_type == 0 or o.use_log_fbank:
        mel = np.log(np.maximum(mel, FLT_EPS)).astype(np.float32)
    if o.feature_type == 1:
        if not o.use_energy:
            return mel
        e = raw_log_energy
        if o.energy_floor > 0.0:
            e = np.maximum(e, F32(np.log(F32(o.energy_floor))))
        return np.concatenate([mel, e[:, None]], 1) if o.htk_compat else np.concatenate([e[:, None], mel], 1)
    feat = (mel @ dct_matrix(o).T).astype(np.float32)         # AddMatVec(dct_matrix_, mel_energies_)
    if o.cepstral_lifter != 0.0:
        feat = (feat * lifter_coeffs(o)[None, :]).astype(np.float32)
    if o.use_energy:
        e = raw_log_energy
        if o.energy_floor > 0.0:
            e = np.maximum(e, F32(np.log(F32(o.energy_floor))))
        feat[:, 0] = e
    if o.htk_compat:
        energy = feat[:, 0].copy()
        feat[:, :-1] = feat[:, 1:].copy()
        if not o.use_energy:
            energy = (energy * F32(np.sqrt(2.0))).astype(np.float32)
        feat[:, -1] = energy
    return feat


def online_cmvn(feats: np.ndarray, cmn_window=600, speaker_frames=600, global_frames=200,
                normalize_mean=True, normalize_variance=False, global_stats=None,
                speaker_stats=None) -> np.ndarray:
    """OnlineCmvn::GetFrame for frames 0..T-1 in order
    (feat/online-feature.cc:421-452): sliding-window stats in double
    (ComputeStatsForFrame :337-368), SmoothOnlineCmvnStats (:372-419),
    ApplyCmvn (transform/cmvn.cc:64-115)."""
    x = np.asarray(feats, np.float32)
    T, D = x.shape
    out = np.zeros_like(x)
    s0 = np.zeros(D + 1, np.float64)
    s1 = np.zeros(D + 1, np.float64)
    g = None if global_stats is None else np.asarray(global_stats, np.float64)
    sp = None if speaker_stats is None else np.asarray(speaker_stats, np.float64)
    for t in range(T):
        xd = x[t].astype(np.float64)
        s0[:D] += xd
        if normalize_variance:
            s1[:D] += xd * xd
        s0[D] += 1.0
        p = t - cmn_window
        if p >= 0:
            pd = x[p].astype(np.float64)
            s0[:D] -= pd
            if normalize_variance:
                s1[:D] -= pd * pd
            s0[D] -= 1.0
        st = np.stack([s0.copy(), s1.copy()])
        cur = st[0, D]
        if cur < cmn_window:
            if sp is not None and sp.size:
                cfs = min(cmn_window - cur, speaker_frames, sp[0, D])
                if cfs > 0.0:
                    st = st + (cfs / sp[0, D]) * sp
                cur = st[0, D]
            if cur < cmn_window:
                if g is None:
                    raise RuntimeError("Global CMN stats are required")   # :417
                cfg = min(cmn_window - cur, global_frames)
                if cfg > 0.0:
                    st = st + (cfg / g[0, D]) * g
        count = st[0, D]
        if not normalize_mean:
            out[t] = x[t]
        elif not normalize_variance:
            # offset.AddVec(-1.0 / count, mean_stats): alpha is BaseFloat (kaldi-vector.cc:1044-1052)
            alpha = np.float64(np.float32(-1.0 / count))
            offset = (alpha * st[0, :D]).astype(np.float32)
            out[t] = x[t] + offset
        else:
            mean = st[0, :D] / count
            var = np.maximum(st[1, :D] / count - mean * mean, 1.0e-20)
            scale = 1.0 / np.sqrt(var)
            out[t] = (x[t] * scale.astype(np.float32)) + (-(mean * scale)).astype(np.float32)
    return out
