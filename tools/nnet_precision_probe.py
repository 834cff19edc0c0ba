"""Error of the nnet3 executor's GEMM kernels against the compiled reference and against a float64 evaluation, full-width
networks (GPU box; B2K_NNET_GEMM selects the kernel).  Usage: python tools/nnet_precision_probe.py <arch> <T> <out.npy>"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from kaldi_b200 import nnet_model as NM
from kaldi_b200.nnet import NnetComputer

which, T, out = sys.argv[1], int(sys.argv[2]), sys.argv[3]
scale_in = float(os.environ.get("PROBE_INPUT_SCALE", "10"))
arch = getattr(NM, "arch_" + which)(6024)
W = NM.random_weights(arch, seed=5)
nc = NnetComputer(arch, W, num_frames=T, max_batch=2)
rng = np.random.default_rng(T)
feats = (rng.standard_normal((T, 40)) * scale_in).astype(np.float32)
civ = rng.standard_normal((nc.n_chunks, 100)).astype(np.float32)
o = nc.forward([feats], [civ])[0]
np.save(out, o)
line = {"mode": os.environ.get("B2K_NNET_GEMM", "ts"), "arch": which, "T": T, "scale": float(np.abs(o).max())}
try:
    from test_scale_gpu import _ref_forward
    ref = _ref_forward(arch, W, feats, civ)
    np.save(out.replace(".npy", "_ref.npy"), ref)
    line["err_vs_ref_over_scale"] = float(np.abs(o - ref).max() / np.abs(ref).max())
    line["rms_err_vs_ref_over_scale"] = float(np.sqrt(np.mean((o - ref) ** 2)) / np.abs(ref).max())
    line["mean_signed_err_over_scale"] = float(np.mean(o - ref) / np.abs(ref).max())
except BaseException as e:   # pytest.skip raises BaseException
    line["ref"] = repr(e)[:100]
print(line)
