"""Generates tests/golden/tiny_final.mdl: the model of nnet_golden.npz (arch_tiny(64), weights seed 11) written as a
binary final.mdl by the reference's own TransitionModel::Write + AmNnetSimple::Write (oracle/_ref; run in the build
container), plus tiny_final_tid2pdf.npy = the reference TransitionModel's TransitionIdToPdf table."""
import ctypes as C
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from kaldi_b200 import nnet_model as NM
from oracle import nnet_oracle as NO

TOPO = """<Topology>
<TopologyEntry>
<ForPhones> 1 2 3 4 5 </ForPhones>
<State> 0 <ForwardPdfClass> 0 <SelfLoopPdfClass> 1 <Transition> 0 0.5 <Transition> 1 0.5 </State>
<State> 1 </State>
</TopologyEntry>
</Topology>
"""
arch = NM.arch_tiny(64)
W = NM.random_weights(arch, seed=11)
R = NO.RefNnet(arch, W, collapse=False)
R.lib.ref_write_final_mdl.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.c_char_p, C.c_int, C.c_int,
                                      C.c_void_p, C.c_int, C.c_void_p, C.c_int]
pri = np.ascontiguousarray(W["priors"], np.float32)
tid2pdf = np.zeros(64, np.int32)
out = os.path.join(ROOT, "tests", "golden", "tiny_final.mdl")
n = R.lib.ref_write_final_mdl(R.h, out.encode(), 1, TOPO.encode(), 5, 2, pri.ctypes.data, pri.size,
                              tid2pdf.ctypes.data, tid2pdf.size)
assert n > 0
np.save(os.path.join(ROOT, "tests", "golden", "tiny_final_tid2pdf.npy"), tid2pdf[:n + 1])
print(os.path.getsize(out), n)
