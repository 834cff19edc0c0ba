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
    LT.write_lattice_text(buf, "utt1", lat, exact=True)
    lines = buf.getvalue().split("\n")
    assert lines[0] == "utt1" and lines[-2] == "" and lines[-1] == ""
    arcs = [l.split("\t") for l in lines[1:] if l.count("\t") == 4]
    finals = [l.split("\t") for l in lines[1:] if l.count("\t") == 1]
    assert len(arcs) == len(lat["arc_src"]) and len(finals) == len(lat["final_state"])
    got = sorted((int(a[0]), int(a[1]), int(a[2]), int(a[3]), np.float32(a[4].split(",")[0]).view(np.int32).item(),
                  np.float32(a[4].split(",")[1]).view(np.int32).item()) for a in arcs)
    want = sorted(zip(lat["arc_src"].tolist(), lat["arc_dst"].tolist(), lat["arc_ilabel"].tolist(), lat["arc_olabel"].tolist(),
                      lat["arc_graph_cost"].view(np.int32).tolist(), lat["arc_acoustic_cost"].view(np.int32).tolist()))
    assert got == want                     # exact=True: repr(float32) round-trips bit for bit
    buf = io.StringIO()
    LT.write_lattice_text(buf, "utt1", lat)            # default: the reference's six significant digits
    w = [l.split("\t")[4].split(",") for l in buf.getvalue().split("\n")[1:] if l.count("\t") == 4]
    assert all(len(x.replace("-", "").replace(".", "").split("e")[0].lstrip("0")) <= 6 for pair in w for x in pair)


def _small_raw_and_compact():
    from kaldi_b200.lattice import determinize_pruned
    f32, i32 = np.float32, np.int32
    lat = dict(state_frame=np.zeros(4, i32), state_hclg=np.arange(4, dtype=i32), state_tot_cost=np.zeros(4, f32),
               state_extra_cost=np.zeros(4, f32),
               arc_src=np.array([0, 0, 1, 2, 0], i32), arc_dst=np.array([1, 2, 3, 3, 3], i32),
               arc_ilabel=np.array([11, 12, 0, 14, 15], i32), arc_olabel=np.array([7, 7, 0, 0, 9], i32),
               arc_graph_cost=np.array([1.0, 0.5, 0.25, 0.25, 5.0], f32), arc_acoustic_cost=np.array([1.0, 1.0, 0.0, 0.0, -0.5], f32),
               final_state=np.array([3], i32), final_cost=np.array([0.5], f32))
    try:
        return lat, determinize_pruned(lat, 100.0)
    except OSError as e:
        pytest.skip(str(e))


def test_binary_archives_round_trip():
    import io
    from kaldi_b200.lattice import read_lattice_archive, write_compact_lattice_binary, write_lattice_binary
    lat, clat = _small_raw_and_compact()
    buf = io.BytesIO()
    write_lattice_binary(buf, "utt-a", lat)
    write_compact_lattice_binary(buf, "utt-b", clat)
    write_lattice_binary(buf, "utt-c", lat)
    entries = read_lattice_archive(buf.getvalue())
    assert [(k, kind) for k, kind, _ in entries] == [("utt-a", "lattice"), ("utt-b", "compact"), ("utt-c", "lattice")]
    a = entries[0][2]
    order = np.argsort(lat["arc_src"], kind="stable")
    for k in ("arc_src", "arc_dst", "arc_ilabel", "arc_olabel", "arc_graph_cost", "arc_acoustic_cost"):
        np.testing.assert_array_equal(a[k], lat[k][order], err_msg=k)
    assert a["start"] == 0 and a["final_state"].tolist() == [3] and a["final_cost"].tolist() == [0.5]
    b = entries[1][2]
    order = np.argsort(clat["arc_src"], kind="stable")
    for k in ("arc_src", "arc_dst", "arc_word", "arc_graph_cost", "arc_acoustic_cost"):
        np.testing.assert_array_equal(b[k], clat[k][order], err_msg=k)
    assert [t.tolist() for t in b["arc_tids"]] == [clat["arc_tids"][i].tolist() for i in order]
    assert b["final_state"].tolist() == clat["final_state"].tolist()
    assert [t.tolist() for t in b["final_tids"]] == [t.tolist() for t in clat["final_tids"]]
    # the first byte after the "\0B" marker is 214 = the low byte of the FST magic (kaldi-lattice.cc:381 relies on it)
    raw = buf.getvalue()
    assert raw[raw.index(b"\0B") + 2] == 214


def test_weight_encodings_equal_the_references_own_write():
    """LatticeWeightTpl::Write / CompactLatticeWeightTpl::Write and the arc type strings, from the reference's
    fstext/lattice-weight.h compiled in oracle/_ref."""
    import ctypes as C
    from kaldi_b200.lattice import _fst_header, compact_weight_bytes, lattice_weight_bytes
    try:
        from oracle import ref_det
        if not ref_det.available():
            pytest.skip("oracle/_ref determinizer library not present")
        L = ref_det.lib()
    except (OSError, RuntimeError) as e:
        pytest.skip(str(e))
    if not hasattr(L, "ref_lattice_weight_bytes"):
        pytest.skip("oracle/_ref library predates ref_lattice_weight_bytes")
    L.ref_lattice_weight_bytes.argtypes = [C.c_float, C.c_float, C.c_void_p, C.c_int32, C.c_void_p, C.c_int32]
    buf = (C.c_uint8 * 256)()
    for g, a, tids in [(1.5, -2.25, [3, 1, 4, 1, 5]), (0.0, 0.0, []), (float("inf"), float("inf"), []), (7.0, 0.125, [2 ** 30])]:
        t = np.array(tids, np.int32)
        n = L.ref_lattice_weight_bytes(g, a, t.ctypes.data, len(t), buf, 256)
        assert bytes(buf[:n]) == compact_weight_bytes(g, a, t)
        n = L.ref_lattice_weight_bytes(g, a, None, -1, buf, 256)
        assert bytes(buf[:n]) == lattice_weight_bytes(g, a)
    s = C.create_string_buffer(64)
    L.ref_lattice_type_strings.argtypes = [C.c_void_p, C.c_int32]
    L.ref_lattice_type_strings(s, 64)
    assert s.value.decode().split() == ["lattice4", "compactlattice44"]
    assert b"lattice4" in _fst_header("lattice4", 0, 1) and b"compactlattice44" in _fst_header("compactlattice44", 0, 1)


def test_scale_compact_lattice():
    from kaldi_b200.lattice import compact_best_path, scale_compact_lattice
    _, clat = _small_raw_and_compact()
    s = scale_compact_lattice(clat, acoustic_scale=0.5)
    np.testing.assert_array_equal(s["arc_acoustic_cost"], clat["arc_acoustic_cost"] * np.float32(0.5))
    np.testing.assert_array_equal(s["arc_graph_cost"], clat["arc_graph_cost"])
    assert compact_best_path(s)["acoustic_cost"] == pytest.approx(0.5 * compact_best_path(clat)["acoustic_cost"])


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_cpp_best_path_equals_python_best_path(seed):
    import ctypes as C
    try:
        from kaldi_b200 import _lib
        from kaldi_b200.decoder import _RawLattice, _p
        L = _lib.lib()
    except OSError as e:
        pytest.skip(str(e))
    lat = _lattice(seed)
    want = LT.best_path(lat)
    keep = {k: np.ascontiguousarray(lat[k], np.float32 if lat[k].dtype.kind == "f" else np.int32) for k in
            ("state_frame", "state_hclg", "state_tot_cost", "state_extra_cost", "arc_src", "arc_dst", "arc_ilabel", "arc_olabel",
             "arc_graph_cost", "arc_acoustic_cost", "final_state", "final_cost")}
    r = _RawLattice()
    r.num_states, r.num_arcs, r.num_finals = len(keep["state_frame"]), len(keep["arc_src"]), len(keep["final_state"])
    for k, v in keep.items():
        setattr(r, k, _p(v, C.c_float if v.dtype == np.float32 else C.c_int32))
    L.b2k_lat_best_path.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]
    nw, nt, g, a = C.c_int32(), C.c_int32(), C.c_float(), C.c_float()
    assert L.b2k_lat_best_path(C.byref(r), None, C.byref(nw), None, C.byref(nt), 0, C.byref(g), C.byref(a)) in (0, 4)   # sizes query
    words, tids = np.zeros(max(nw.value, 1), np.int32), np.zeros(max(nt.value, 1), np.int32)
    cap = max(nw.value, nt.value, 1)
    words, tids = np.zeros(cap, np.int32), np.zeros(cap, np.int32)
    assert L.b2k_lat_best_path(C.byref(r), words.ctypes.data, C.byref(nw), tids.ctypes.data, C.byref(nt), cap, C.byref(g), C.byref(a)) == 0
    assert words[:nw.value].tolist() == want["olabels"].tolist()
    assert tids[:nt.value].tolist() == want["ilabels"].tolist()
    assert g.value == pytest.approx(want["graph_cost"], abs=1e-3) and a.value == pytest.approx(want["acoustic_cost"], abs=1e-3)


def test_cpp_best_path_arcs_trace_the_same_path():
    import ctypes as C
    try:
        from kaldi_b200 import _lib
        from kaldi_b200.decoder import _RawLattice, _p
        L = _lib.lib()
    except OSError as e:
        pytest.skip(str(e))
    lat = _lattice(1)
    want = LT.best_path(lat)
    keep = {k: np.ascontiguousarray(lat[k], np.float32 if lat[k].dtype.kind == "f" else np.int32) for k in
            ("state_frame", "state_hclg", "state_tot_cost", "state_extra_cost", "arc_src", "arc_dst", "arc_ilabel", "arc_olabel",
             "arc_graph_cost", "arc_acoustic_cost", "final_state", "final_cost")}
    r = _RawLattice()
    r.num_states, r.num_arcs, r.num_finals = len(keep["state_frame"]), len(keep["arc_src"]), len(keep["final_state"])
    for k, v in keep.items():
        setattr(r, k, _p(v, C.c_float if v.dtype == np.float32 else C.c_int32))
    L.b2k_lat_best_path_arcs.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]
    n, fi = C.c_int64(), C.c_int64()
    assert L.b2k_lat_best_path_arcs(C.byref(r), None, C.byref(n), 0, C.byref(fi)) == 4       # B2K_ERR_OVERFLOW: size returned
    arcs = np.zeros(n.value, np.int64)
    assert L.b2k_lat_best_path_arcs(C.byref(r), arcs.ctypes.data, C.byref(n), n.value, C.byref(fi)) == 0
    states = [0] + [int(lat["arc_dst"][a]) for a in arcs]
    assert states == want["states"].tolist()
    assert all(int(lat["arc_src"][a]) == s for a, s in zip(arcs, states[:-1]))
    assert int(lat["final_state"][fi.value]) == states[-1]
    # empty lattice: nothing, no error
    e = _RawLattice()
    assert L.b2k_lat_best_path_arcs(C.byref(e), None, C.byref(n), 0, C.byref(fi)) == 0 and n.value == 0 and fi.value == -1


def test_text_weights_equal_the_references_printer():
    """The weight column of text-mode lattices: the reference's own operator<< for LatticeWeight / CompactLatticeWeight."""
    import ctypes as C
    try:
        from oracle import ref_det
        if not ref_det.available():
            pytest.skip("oracle/_ref determinizer library not present")
        L = ref_det.lib()
    except (OSError, RuntimeError) as e:
        pytest.skip(str(e))
    if not hasattr(L, "ref_lattice_weight_text"):
        pytest.skip("oracle/_ref library predates ref_lattice_weight_text")
    L.ref_lattice_weight_text.argtypes = [C.c_float, C.c_float, C.c_void_p, C.c_int32, C.c_void_p, C.c_int32]
    buf = C.create_string_buffer(512)
    rng = np.random.default_rng(0)
    cases = [(1.5, -2.25, [3, 1, 4]), (0.1, 1e-7, []), (123456.789, 0.30000001, [7]), (float("inf"), float("inf"), []),
             (0.0, 0.0, [2 ** 30]), (1e10, -1e-10, [1, 2]), (float("-inf"), 3.0, [5])]
    cases += [(float(np.float32(rng.normal() * 10.0 ** int(rng.integers(-6, 7)))), float(np.float32(rng.normal())), rng.integers(1, 999, rng.integers(0, 5)).tolist())
              for _ in range(200)]
    for g, a, tids in cases:
        t = np.array(tids, np.int32)
        L.ref_lattice_weight_text(g, a, t.ctypes.data, len(t), buf, 512)
        assert buf.value.decode() == f"{LT._num(g)},{LT._num(a)},{'_'.join(str(int(x)) for x in tids)}"
        L.ref_lattice_weight_text(g, a, None, -1, buf, 512)
        assert buf.value.decode() == f"{LT._num(g)},{LT._num(a)}"


def test_text_tables_leave_out_unit_weights():
    """FstPrinter with show_weight_one = false: an arc / final weight equal to One() has no weight column."""
    f32, i32 = np.float32, np.int32
    lat = dict(state_frame=np.zeros(3, i32), state_hclg=np.arange(3, dtype=i32), state_tot_cost=np.zeros(3, f32),
               state_extra_cost=np.zeros(3, f32), arc_src=np.array([0, 1], i32), arc_dst=np.array([1, 2], i32),
               arc_ilabel=np.array([4, 5], i32), arc_olabel=np.array([0, 9], i32), arc_graph_cost=np.array([0.0, 1.25], f32),
               arc_acoustic_cost=np.array([0.0, -3.0], f32), final_state=np.array([2], i32), final_cost=np.array([0.0], f32))
    buf = io.StringIO()
    LT.write_lattice_text(buf, "u", lat)
    assert buf.getvalue() == "u\n0\t1\t4\t0\n1\t2\t5\t9\t1.25,-3\n2\n\n"


@pytest.mark.parametrize("binary", [0, 1])
def test_cpp_table_writers_write_the_same_bytes_as_the_python_ones(tmp_path, binary):
    import ctypes as C
    try:
        from kaldi_b200 import _lib
        from kaldi_b200.decoder import _RawLattice, _p
        L = _lib.lib()
    except OSError as e:
        pytest.skip(str(e))
    lat = _lattice(2)
    lat["arc_graph_cost"][:3] = 0.0
    lat["arc_acoustic_cost"][:3] = 0.0                      # a few unit weights
    keep = {k: np.ascontiguousarray(lat[k], np.float32 if lat[k].dtype.kind == "f" else np.int32) for k in
            ("state_frame", "state_hclg", "state_tot_cost", "state_extra_cost", "arc_src", "arc_dst", "arc_ilabel", "arc_olabel",
             "arc_graph_cost", "arc_acoustic_cost", "final_state", "final_cost")}
    r = _RawLattice()
    r.num_states, r.num_arcs, r.num_finals = len(keep["state_frame"]), len(keep["arc_src"]), len(keep["final_state"])
    for k, v in keep.items():
        setattr(r, k, _p(v, C.c_float if v.dtype == np.float32 else C.c_int32))
    p = str(tmp_path / "lat.ark")
    L.b2k_lat_write.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p, C.c_int32, C.c_int32]
    L.b2k_clat_write.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p, C.c_int32, C.c_int32]
    assert L.b2k_lat_write(C.byref(r), b"utt-1", p.encode(), binary, 0) == 0
    h = C.c_void_p()
    L.b2k_lat_determinize_pruned.argtypes = [C.c_void_p, C.c_float, C.c_int64, C.c_void_p]
    assert L.b2k_lat_determinize_pruned(C.byref(r), 8.0, 0, C.byref(h)) == 0
    assert L.b2k_clat_write(h, b"utt-1", p.encode(), binary, 1) == 0            # appended
    L.b2k_clat_destroy.argtypes = [C.c_void_p]
    L.b2k_clat_destroy(h)
    clat = LT.determinize_pruned(lat, 8.0)
    buf = io.BytesIO() if binary else io.StringIO()
    if binary:
        LT.write_lattice_binary(buf, "utt-1", lat)
        LT.write_compact_lattice_binary(buf, "utt-1", clat)
        want = buf.getvalue()
    else:
        LT.write_lattice_text(buf, "utt-1", lat)
        LT.write_compact_lattice_text(buf, "utt-1", clat)
        want = buf.getvalue().encode()
    assert open(p, "rb").read() == want
    assert L.b2k_lat_write(C.byref(r), b"k", str(tmp_path / "no" / "such" / "dir").encode(), binary, 0) != 0


def test_symbol_table_and_transcript(tmp_path):
    p = str(tmp_path / "words.txt")
    open(p, "w", encoding="utf-8").write("<eps> 0\nhello 1\nwörld\t2\n\n#0 3\n")
    st = LT.read_symbol_table(p)
    assert st == {0: "<eps>", 1: "hello", 2: "wörld", 3: "#0"}
    assert LT.transcript(np.array([1, 2, 1], np.int32), st) == "hello wörld hello" and LT.transcript([], st) == ""
    with pytest.raises(KeyError):
        LT.transcript([7], st)
    open(p, "w").write("only-one-column\n")
    with pytest.raises(ValueError):
        LT.read_symbol_table(p)


@pytest.mark.parametrize("seed", [0, 3])
def test_cpp_compact_best_path_equals_python(seed):
    import ctypes as C
    try:
        from kaldi_b200 import _lib
        from kaldi_b200.decoder import _RawLattice, _p
        L = _lib.lib()
    except OSError as e:
        pytest.skip(str(e))
    lat = _lattice(seed)
    want = LT.compact_best_path(LT.determinize_pruned(lat, 8.0))
    keep = {k: np.ascontiguousarray(lat[k], np.float32 if lat[k].dtype.kind == "f" else np.int32) for k in
            ("state_frame", "state_hclg", "state_tot_cost", "state_extra_cost", "arc_src", "arc_dst", "arc_ilabel", "arc_olabel",
             "arc_graph_cost", "arc_acoustic_cost", "final_state", "final_cost")}
    r = _RawLattice()
    r.num_states, r.num_arcs, r.num_finals = len(keep["state_frame"]), len(keep["arc_src"]), len(keep["final_state"])
    for k, v in keep.items():
        setattr(r, k, _p(v, C.c_float if v.dtype == np.float32 else C.c_int32))
    h = C.c_void_p()
    L.b2k_lat_determinize_pruned.argtypes = [C.c_void_p, C.c_float, C.c_int64, C.c_void_p]
    assert L.b2k_lat_determinize_pruned(C.byref(r), 8.0, 0, C.byref(h)) == 0
    L.b2k_clat_best_path.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]
    nw, nt, g, a = C.c_int32(), C.c_int32(), C.c_float(), C.c_float()
    assert L.b2k_clat_best_path(h, None, C.byref(nw), None, C.byref(nt), 0, 0, C.byref(g), C.byref(a)) in (0, 4)
    words, tids = np.zeros(max(nw.value, 1), np.int32), np.zeros(max(nt.value, 1), np.int32)
    assert L.b2k_clat_best_path(h, words.ctypes.data, C.byref(nw), tids.ctypes.data, C.byref(nt), words.size, tids.size, C.byref(g), C.byref(a)) == 0
    L.b2k_clat_destroy.argtypes = [C.c_void_p]
    L.b2k_clat_destroy(h)
    assert words[:nw.value].tolist() == want["words"].tolist() and tids[:nt.value].tolist() == want["tids"].tolist()
    assert g.value == pytest.approx(want["graph_cost"], abs=1e-3) and a.value == pytest.approx(want["acoustic_cost"], abs=1e-3)
    bp = LT.best_path(lat)                                   # and it is the raw lattice's best path
    assert words[:nw.value].tolist() == bp["olabels"].tolist() and tids[:nt.value].tolist() == bp["ilabels"].tolist()
