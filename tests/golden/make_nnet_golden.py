"""Generates tests/golden/nnet_golden.npz from the reference's own nnet3 CPU
forward (oracle/_ref; run in the build container)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from kaldi_b200 import nnet_model as NM
from oracle import nnet_oracle as NO
arch = NM.arch_tiny(64)
W = NM.random_weights(arch, seed=11)
R = NO.RefNnet(arch, W)
rng = np.random.default_rng(0)
T = 100
feats = (rng.standard_normal((T, 40)) * 10).astype(np.float32)
iv = rng.standard_normal((T, 100)).astype(np.float32)
ref = R.forward(feats, iv, period=1)
rows = R.chunk_ivector_rows(T, T, 1)
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "nnet_golden.npz"), feats=feats,
                    chunk_ivectors=iv[rows], ref_out=ref)
print(ref.shape, rows)
