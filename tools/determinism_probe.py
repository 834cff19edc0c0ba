"""Is the pipeline deterministic run to run on the inputs of tests/test_zz_model_route.py::test_pure_cpp_route...?  Decodes the
same batch repeatedly (one pipeline object and fresh ones), compares log-likelihoods, canonical raw lattices (order-free) and
the determinized compact lattices.  GPU box only."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kaldi_b200 import synth
from kaldi_b200.decoder import CudaDecoder, CudaFst, lattice_to_canonical
from kaldi_b200.lattice import determinize_pruned
from kaldi_b200.model import KaldiModel
from kaldi_b200.pipeline import NativeBatchedPipeline, PipelineConfig

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MDL = os.path.join(ROOT, "tests", "golden", "tiny_final.mdl")
S, n = 32000, 3
g = synth.make_hclg(50_000, num_pdfs=64, seed=4)
waves = [synth.make_audio(S, seed=40 + i).astype(np.float32) for i in range(n)]
cfg = PipelineConfig(max_batch=2, num_samples=S, extract_ivectors=False)
beam = float(cfg.decoder_cfg["lattice_beam"])
first = None
nat = None
for rep in range(int(sys.argv[1]) if len(sys.argv) > 1 else 12):
    if nat is None or rep % 4 == 0:
        nat = NativeBatchedPipeline(cfg, KaldiModel(MDL), CudaFst(g), None)
    raw = nat.decode_batch(waves[0:2])
    ll = nat.read("loglikes", 2).copy()
    lats = CudaDecoder.SplitLattices(raw)
    can = [lattice_to_canonical(l) for l in lats]
    det = [determinize_pruned(l, beam) for l in lats]
    cur = dict(ll=ll, can=can, det=det)
    if first is None:
        first = cur
        print("rep 0: states", [len(c["states"]) for c in can], "finals", [(d["final_graph_cost"], d["final_acoustic_cost"]) for d in det])
        continue
    msgs = []
    if not np.array_equal(first["ll"], ll):
        msgs.append("loglikes differ (max abs %g)" % np.abs(first["ll"] - ll).max())
    for u in range(2):
        for k in can[u]:
            if not np.array_equal(first["can"][u][k], can[u][k]):
                msgs.append(f"utt {u} canonical raw lattice field {k} differs")
        for k in ("arc_src", "arc_dst", "arc_word", "arc_graph_cost", "arc_acoustic_cost", "final_state", "final_graph_cost", "final_acoustic_cost"):
            if not np.array_equal(first["det"][u][k], det[u][k]):
                msgs.append(f"utt {u} compact lattice field {k} differs: {first['det'][u][k][:3]} vs {det[u][k][:3]}")
    print("rep", rep, "OK" if not msgs else "; ".join(msgs[:6]))
