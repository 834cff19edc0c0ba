"""CPU timing of the raw-lattice -> compact-lattice step (kaldi_b200/csrc/lattice_det.cu) on lattices of the bench
workload: utterances go through the reference's own CPU path compiled in oracle/_ref (as bench.py --impl reference
does), the finalized raw lattice is determinized with beam = lattice_beam.  Host only; prints one JSON line."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from kaldi_b200 import synth  # noqa: E402
from kaldi_b200.lattice import best_path, compact_best_path, determinize_pruned, raw_lattice_from_canonical  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 3
    rows = []
    bench._CPU_STATE["record"] = True          # keep the keyed lattice of each utterance
    for i in range(n):
        bench.cpu_reference_one((1000 + i, 0))
        dec = bench._CPU_STATE["dec"]
        lat = raw_lattice_from_canonical(dec.lattice())
        beam = float(synth.DEFAULT_DECODER_CFG["lattice_beam"])
        t0 = time.perf_counter()
        c = determinize_pruned(lat, beam)
        dt = time.perf_counter() - t0
        a, b = best_path(lat), compact_best_path(c)
        assert abs(a["total_cost"] - b["total_cost"]) < 1e-2 and a["olabels"].tolist() == b["words"].tolist()
        rows.append(dict(raw_states=len(lat["state_frame"]), raw_arcs=len(lat["arc_src"]), det_states=c["num_states"],
                         det_arcs=len(c["arc_src"]), ms=dt * 1e3, words=len(b["words"]), **c["stats"]))
    print(json.dumps(dict(workload="mini_librispeech_tdnn_1k/10s-utts/beam15/hclg5M", lattice_beam=beam, utterances=rows,
                          mean_ms=float(np.mean([r["ms"] for r in rows])))))


if __name__ == "__main__":
    main()
