"""Generates tests/golden/feat_*.npz from the reference itself (run in the build
container, where /root/reference exists).  Inputs: the reference's own
src/feat/test_data/test.wav (+ HTK golden test.wav.fea_htk.1 produced by HTK
HCopy, src/feat/test_data/README) and one seeded synthetic utterance.  Outputs
are produced by the reference's feature code compiled in oracle/_ref."""
import os, struct, sys, wave
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from kaldi_b200 import synth
from oracle import feat_oracle as F

TD = "/root/reference/src/feat/test_data"
w = wave.open(os.path.join(TD, "test.wav"))
x = np.frombuffer(w.readframes(w.getnframes()), dtype="<i2").copy()


def read_htk(path):
    d = open(path, "rb").read()
    n, period, sz, kind = struct.unpack(">iihh", d[:12])
    return np.frombuffer(d[12:12 + n * sz], dtype=">f4").reshape(n, sz // 4).astype(np.float32)


R = F.RefFeat()
out = dict(test_wav=x, htk_fea_1_static=read_htk(os.path.join(TD, "test.wav.fea_htk.1"))[:, :13])
cfgs = dict(
    mfcc_hires=F.FeatOpts(),
    fbank40=F.FeatOpts(feature_type=1),
    mfcc_hires_nosnip=F.FeatOpts(snip_edges=0),
    mfcc13_energy=F.FeatOpts(num_bins=23, num_ceps=13, use_energy=1, high_freq=0.0, energy_floor=1.0),
    htk1=F.FeatOpts(num_bins=23, num_ceps=13, preemph_coeff=0.0, window_type=1, remove_dc_offset=0,
                    low_freq=0.0, high_freq=0.0, htk_mode=1, htk_compat=1, use_energy=0),
)
xf = x.astype(np.float32)
syn = synth.make_audio(16000, seed=42)
for k, o in cfgs.items():
    out["testwav_" + k] = R.compute(xf, o)
    out["synth42_" + k] = R.compute(syn, o)
m = out["synth42_mfcc_hires"]
gs = np.zeros((2, 41)); gs[0, :40] = m.sum(0) * 3; gs[1, :40] = (m.astype(np.float64) ** 2).sum(0) * 3; gs[0, 40] = 3 * len(m)
out["cmvn_global_stats"] = gs
out["cmvn_mean_w50"] = R.online_cmvn(m, cmn_window=50, speaker_frames=50, global_frames=20, global_stats=gs)
out["cmvn_var_w50"] = R.online_cmvn(m, cmn_window=50, speaker_frames=50, global_frames=20, global_stats=gs, normalize_variance=True)
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "feat_golden.npz"), **out)
print({k: v.shape for k, v in out.items()})
