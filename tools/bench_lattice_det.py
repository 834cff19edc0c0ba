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
    try:
        from oracle import ref_det as RD
        RD.lib()
    except Exception:
        RD = None
    bench._CPU_STATE["record"] = True          # keep the keyed lattice of each utterance
    for i in range(n):
        bench.cpu_reference_one((1000 + i, 0))
        dec = bench._CPU_STATE["dec"]
        lat = raw_lattice_from_canonical(dec.lattice())
        beam = float(synth.DEFAULT_DECODER_CFG["lattice_beam"])
        t0 = time.perf_counter()
        c = determinize_pruned(lat, beam)
        dt = time.perf_counter() - t0
        ref_ms = None
        if RD is not None:                    # the reference's own determinizer on the same lattice (phone + word passes)
            ntid = int(lat["arc_ilabel"].max()) + 2
            t = np.arange(ntid)
            t1 = time.perf_counter()
            r = RD.determinize(lat, beam, phone_determinize=True, phone_of=(1 + np.maximum(t - 1, 0) // 2 % 40).astype(np.int32),
                               self_loop=((t % 2 == 0) & (t > 0)).astype(np.uint8), phone_start=(t % 2 == 1).astype(np.uint8))
            ref_ms = (time.perf_counter() - t1) * 1e3
            rb = compact_best_path(r)
            assert rb["words"].tolist() == compact_best_path(c)["words"].tolist()
        a, b = best_path(lat), compact_best_path(c)
        assert abs(a["total_cost"] - b["total_cost"]) < 1e-2 and a["olabels"].tolist() == b["words"].tolist()
        rows.append(dict(raw_states=len(lat["state_frame"]), raw_arcs=len(lat["arc_src"]), det_states=c["num_states"],
                         det_arcs=len(c["arc_src"]), ms=dt * 1e3, reference_ms=ref_ms, words=len(b["words"]), **c["stats"]))
    print(json.dumps(dict(workload="mini_librispeech_tdnn_1k/10s-utts/beam15/hclg5M", lattice_beam=beam, utterances=rows,
                          mean_ms=float(np.mean([r["ms"] for r in rows])),
                          reference_mean_ms=(float(np.mean([r["reference_ms"] for r in rows])) if RD is not None else None))))


if __name__ == "__main__":
    main()
