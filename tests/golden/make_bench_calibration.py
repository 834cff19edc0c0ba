"""Generates tests/golden/bench_calibration.npz: the prior-style output
calibration (per-pdf mean of the raw nnet output + scale to per-frame std 1.0)
of the synthetic mini_librispeech model used by bench.py, so that our arm and
the reference arm decode exactly the same model.  Computed on the CPU with the
nnet3 restatement over 4 seeded calibration utterances (i-vectors zero)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from kaldi_b200 import nnet_model as NM, synth
from oracle import nnet_oracle as NO, feat_oracle as F
arch = NM.arch_mini_librispeech_1k(2336)
W = NM.random_weights(arch, seed=0)
raws = []
for i in range(4):
    feats = F.mfcc_fbank(synth.make_audio(160000, seed=777_000 + i), F.FeatOpts())
    raws.append(NO.forward_dense(arch, W, feats, np.zeros((48, 100), np.float32), use_priors=False).astype(np.float64))
raw = np.concatenate(raws)
mean = raw.mean(0)
scale = 1.0 / (raw - mean[None, :]).std(1).mean()
np.savez(os.path.join(ROOT, "tests", "golden", "bench_calibration.npz"), mean=mean.astype(np.float32), scale=np.float32(scale))
print("scale", scale, "static std", mean.std())
