"""kaldi_b200/lattice.py on lattices produced by the reference's own decoder (oracle/_ref) and the restatement."""
import io
import itertools

import numpy as np
import pytest

from kaldi_b200 import lattice as LT
from kaldi_b200 import synth
from oracle import dec_oracle as D


def _lattice(seed, cfgmod=None, T=40):
    g = synth.make_hclg(200_000, num_pdfs=400, seed=seed)
    ll = synth.make_loglikes(g, T, seed=seed + 100)
    cfg = dict(synth.DEFAULT_DECODER_CFG, **(cfgmod or {}))
    from oracle import ref_decoder as R
    if R.available():
        d = R.RefDecoder(g, cfg)
        d.decode(ll)
        return LT.raw_lattice_from_canonical(d.lattice())
    o = D.DecoderOracle(g, cfg)
    o.decode(ll, mode=D.MODE_REFERENCE_ORDER)
    return LT.raw_lattice_from_canonical(o.lattice())


def _all_paths_cost(lat, limit=200000):
    """Brute force: cheapest start->final path by exhaustive DFS (small lattices only)."""
    out = {}
    for a in range(len(lat["arc_src"])):
        out.setdefault(int(lat["arc_src"][a]), []).append(a)
    fin = dict(zip(lat["final_state"].tolist(), lat["final_cost"].astype(np.float64).tolist()))
    best = [np.inf]
    count = [0]

    def rec(s, c):
        count[0] += 1
        assert count[0] < limit
        if s in fin:
            best[0] = min(best[0], c + fin[s])
        for a in out.get(s, []):
            rec(int(lat["arc_dst"][a]), c + float(lat["arc_graph_cost"][a]) + float(lat["arc_acoustic_cost"][a]))
    rec(0, 0.0)
    return best[0]


@pytest.mark.parametrize("seed,cfgmod", [(0, {}), (3, {"lattice_beam": 4.0}), (5, {"beam": 10.0, "lattice_beam": 3.0})])
def test_best_path_is_the_cheapest_path(seed, cfgmod):
    lat = _lattice(seed, cfgmod)
    bp = LT.best_path(lat)
    assert bp["states"][0] == 0 and bp["states"][-1] in set(lat["final_state"].tolist())
    assert abs(bp["total_cost"] - (bp["graph_cost"] + bp["acoustic_cost"])) < 1e-6 * max(1.0, abs(bp["total_cost"]))
    assert abs(bp["total_cost"] - _all_paths_cost(lat)) < 1e-6 * max(1.0, abs(bp["total_cost"]))
    # the best path runs through tokens whose extra_cost is (numerically) zero: that is what the pruning sweep computes
    assert np.all(lat["state_extra_cost"][bp["states"]] <= 1e-3)
    # one transition-id per frame
    assert len(bp["ilabels"]) == int(lat["state_frame"].max())


def test_lattice_text_round_trip():
    lat = _lattice(1)
    buf = io.StringIO()
    LT.write_lattice_text(buf, "utt1", lat)
    lines = buf.getvalue().split("\n")
    assert lines[0] == "utt1" and lines[-2] == "" and lines[-1] == ""
    arcs = [l.split("\t") for l in lines[1:] if l.count("\t") == 4]
    finals = [l.split("\t") for l in lines[1:] if l.count("\t") == 1]
    assert len(arcs) == len(lat["arc_src"]) and len(finals) == len(lat["final_state"])
    got = sorted((int(a[0]), int(a[1]), int(a[2]), int(a[3]), np.float32(a[4].split(",")[0]).view(np.int32).item(),
                  np.float32(a[4].split(",")[1]).view(np.int32).item()) for a in arcs)
    want = sorted(zip(lat["arc_src"].tolist(), lat["arc_dst"].tolist(), lat["arc_ilabel"].tolist(), lat["arc_olabel"].tolist(),
                      lat["arc_graph_cost"].view(np.int32).tolist(), lat["arc_acoustic_cost"].view(np.int32).tolist()))
    assert got == want                     # repr(float32) round-trips bit for bit
