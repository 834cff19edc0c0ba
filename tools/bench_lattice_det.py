"""CPU timing of the raw-lattice -> compact-lattice step (kaldi_b200/csrc/lattice_det.cu) on lattices of the bench workload:
utterances go through the reference's own CPU path compiled in oracle/_ref (as bench.py --impl reference does), the finalized raw
lattice is determinized with beam = lattice_beam.  Times are those of the C calls alone (no Python marshalling): ours =
b2k_lat_determinize_pruned, reference = DeterminizeLatticePhonePrunedWrapper of lat/determinize-lattice-pruned.cc in oracle/_ref
(compiled against the container-only OpenFst stand-in, so its containers are not OpenFst's), in its default configuration (phone
pass + word pass) and word pass only.  Host only, one core; prints one JSON line."""
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from kaldi_b200 import _lib, synth  # noqa: E402
from kaldi_b200.decoder import _RawLattice, _p  # noqa: E402
from kaldi_b200.lattice import best_path, compact_best_path, determinize_pruned, raw_lattice_from_canonical  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 4
    bench._CPU_STATE["record"] = True          # keep the keyed lattice of each utterance
    bench._CPU_STATE["waves"] = [synth.make_audio(bench.NUM_SAMPLES, seed=1000 + i) for i in range(n)]
    L = _lib.lib()
    L.b2k_lat_determinize_pruned.argtypes = [C.c_void_p, C.c_float, C.c_int64, C.c_void_p]
    L.b2k_clat_destroy.argtypes = [C.c_void_p]
    try:
        from oracle import ref_det as RD
        RL = RD.lib()
        RL.ref_det_last_ms.restype = C.c_double
    except Exception:
        RD = None
    beam = float(synth.DEFAULT_DECODER_CFG["lattice_beam"])
    rows = []
    for i in range(n):
        bench.cpu_reference_one((i, bench.DEFAULT_WORKLOAD))
        lat = raw_lattice_from_canonical(bench._CPU_STATE["dec"].lattice())
        keep = {k: np.ascontiguousarray(lat[k], np.float32 if lat[k].dtype.kind == "f" else np.int32) for k in lat}
        r = _RawLattice()
        r.num_states, r.num_arcs, r.num_finals = len(keep["state_frame"]), len(keep["arc_src"]), len(keep["final_state"])
        for k, v in keep.items():
            setattr(r, k, _p(v, C.c_float if v.dtype == np.float32 else C.c_int32))
        t0 = time.perf_counter()
        for _ in range(20):
            h = C.c_void_p()
            L.b2k_lat_determinize_pruned(C.byref(r), beam, 0, C.byref(h))
            L.b2k_clat_destroy(h)
        ours = (time.perf_counter() - t0) / 20 * 1e3
        c = determinize_pruned(lat, beam)
        a, b = best_path(lat), compact_best_path(c)
        assert abs(a["total_cost"] - b["total_cost"]) < 1e-2 and a["olabels"].tolist() == b["words"].tolist()
        row = dict(raw_states=len(lat["state_frame"]), raw_arcs=len(lat["arc_src"]), det_states=c["num_states"], det_arcs=len(c["arc_src"]),
                   words=len(b["words"]), ours_ms=ours, **c["stats"])
        if RD is not None:
            ntid = int(lat["arc_ilabel"].max()) + 2
            t = np.arange(ntid)
            ph = dict(phone_of=(1 + np.maximum(t - 1, 0) // 2 % 40).astype(np.int32), self_loop=((t % 2 == 0) & (t > 0)).astype(np.uint8),
                      phone_start=(t % 2 == 1).astype(np.uint8))
            for name, pd in (("reference_phone_and_word_ms", True), ("reference_word_only_ms", False)):
                ms = []
                for _ in range(8):
                    rr = RD.determinize(lat, beam, phone_determinize=pd, **ph)
                    ms.append(RL.ref_det_last_ms())
                row[name] = float(np.mean(ms))
                assert compact_best_path(rr)["words"].tolist() == b["words"].tolist()
        rows.append(row)
    out = dict(workload="mini_librispeech_tdnn_1k/10s-utts/beam15/hclg5M", lattice_beam=beam, cores=1, utterances=rows,
               ours_mean_ms=float(np.mean([r["ours_ms"] for r in rows])))
    if RD is not None:
        out["reference_phone_and_word_mean_ms"] = float(np.mean([r["reference_phone_and_word_ms"] for r in rows]))
        out["reference_word_only_mean_ms"] = float(np.mean([r["reference_word_only_ms"] for r in rows]))
    print(json.dumps(out))


if __name__ == "__main__":
    main()
