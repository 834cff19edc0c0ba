"""Host-side utilities on finalized raw lattices (the dictionaries CudaDecoder.GetRawLattice /
SplitLattices return) — SURVEY.md §8(f) row 1, the part that needs no determinization:

* best_path(): LatticeFasterDecoderTpl::GetBestPath (decoder/lattice-faster-decoder.cc:102-108 =
  GetRawLattice + fst::ShortestPath over the tropical sum graph + acoustic): 1-best transition-id and
  word sequence with its graph / acoustic cost;
* write_lattice_text(): the text form of a Kaldi `Lattice` table entry (lat/kaldi-lattice.cc WriteLattice,
  text mode: key line, one `src dst ilabel olabel graph,acoustic` line per arc, `state graph,acoustic`
  per final state, blank line), which `lattice-copy ark,t:- ark:out` style tools read.

Plain numpy on the host; nothing here touches the GPU.
"""
from __future__ import annotations

import numpy as np


def _topological_order(num_states: int, src: np.ndarray, dst: np.ndarray) -> np.ndarray:
    indeg = np.bincount(dst, minlength=num_states).astype(np.int64)
    order = np.argsort(src, kind="stable")
    starts = np.searchsorted(src[order], np.arange(num_states + 1))
    ready = [int(s) for s in np.flatnonzero(indeg == 0)]
    out = []
    while ready:
        s = ready.pop()
        out.append(s)
        for a in order[starts[s]:starts[s + 1]]:
            d = int(dst[a])
            indeg[d] -= 1
            if indeg[d] == 0:
                ready.append(d)
    if len(out) != num_states:
        raise ValueError("the raw lattice has a cycle (epsilon cycle in the decoding graph?)")
    return np.array(out, np.int64)


def best_path(lat: dict) -> dict:
    """Lowest-cost path from lattice state 0 (the start state) to a final state, cost = sum of graph and
    acoustic costs of the arcs plus the final cost.  Returns ilabels (transition-ids, epsilons removed),
    olabels (word ids, epsilons removed), graph_cost, acoustic_cost (final cost added to the graph part as
    LatticeWeight(final, 0) does, :188), total_cost and the state sequence."""
    ns = len(lat["state_frame"])
    if ns == 0:
        return dict(ilabels=np.zeros(0, np.int32), olabels=np.zeros(0, np.int32), graph_cost=np.inf,
                    acoustic_cost=np.inf, total_cost=np.inf, states=np.zeros(0, np.int32))
    src, dst = lat["arc_src"].astype(np.int64), lat["arc_dst"].astype(np.int64)
    w = lat["arc_graph_cost"].astype(np.float64) + lat["arc_acoustic_cost"].astype(np.float64)
    dist = np.full(ns, np.inf)
    back = np.full(ns, -1, np.int64)
    dist[0] = 0.0
    order = np.argsort(src, kind="stable")
    starts = np.searchsorted(src[order], np.arange(ns + 1))
    for s in _topological_order(ns, src, dst):
        if not np.isfinite(dist[s]):
            continue
        for a in order[starts[s]:starts[s + 1]]:
            nd = dist[s] + w[a]
            if nd < dist[dst[a]]:
                dist[dst[a]] = nd
                back[dst[a]] = a
    fs, fc = lat["final_state"].astype(np.int64), lat["final_cost"].astype(np.float64)
    if len(fs) == 0:
        raise ValueError("lattice without final states")
    tot = dist[fs] + fc
    k = int(np.argmin(tot))
    if not np.isfinite(tot[k]):
        raise ValueError("no final state is reachable from the start state")
    arcs = []
    s = int(fs[k])
    states = [s]
    while s != 0:
        a = int(back[s])
        arcs.append(a)
        s = int(src[a])
        states.append(s)
    arcs = np.array(arcs[::-1], np.int64)
    il, ol = lat["arc_ilabel"][arcs], lat["arc_olabel"][arcs]
    return dict(ilabels=il[il != 0].astype(np.int32), olabels=ol[ol != 0].astype(np.int32),
                graph_cost=float(lat["arc_graph_cost"][arcs].astype(np.float64).sum() + fc[k]),
                acoustic_cost=float(lat["arc_acoustic_cost"][arcs].astype(np.float64).sum()),
                total_cost=float(tot[k]), states=np.array(states[::-1], np.int32))


def _num(x: float, exact: bool = False) -> str:
    """A float as the reference's weight printer writes it (LatticeWeightTpl::WriteFloatType, fstext/lattice-weight.h:148-160,
    on a default-precision ostream): six significant digits, "Infinity" / "-Infinity" / "BadNumber"."""
    x = float(np.float32(x))
    if exact and x == x and abs(x) != float("inf"):
        return repr(x)                                       # shortest text that reads back to the same float32
    if x == float("inf"):
        return "Infinity"
    if x == float("-inf"):
        return "-Infinity"
    if x != x:
        return "BadNumber"
    return "%g" % x


def write_lattice_text(f, key: str, lat: dict, exact: bool = False) -> None:
    """One entry of a text-mode Lattice table (`ark,t`): arcs grouped by source state as an FST printer emits them.  Numbers are
    printed as the reference prints them (six significant digits); exact=True prints every float32 in full instead (the
    reference's reader accepts both), for archives that must read back bit for bit."""
    num = lambda x: _num(x, exact)
    f.write(key + "\n")
    src = lat["arc_src"]
    order = np.argsort(src, kind="stable")
    finals = dict(zip(lat["final_state"].tolist(), lat["final_cost"].tolist()))
    starts = np.searchsorted(src[order], np.arange(len(lat["state_frame"]) + 1))
    for s in range(len(lat["state_frame"])):
        for a in order[starts[s]:starts[s + 1]]:
            g, ac = float(lat["arc_graph_cost"][a]), float(lat["arc_acoustic_cost"][a])
            # fst::FstPrinter with show_weight_one = false (kaldi-lattice.cc:406-409) leaves the weight column out when the
            # weight is One(); the reference's reader accepts both forms (PARITY of this detail unpinned: no OpenFst here)
            w = "" if (g == 0.0 and ac == 0.0) else f"\t{num(g)},{num(ac)}"
            f.write(f"{s}\t{int(lat['arc_dst'][a])}\t{int(lat['arc_ilabel'][a])}\t{int(lat['arc_olabel'][a])}{w}\n")
        if s in finals:
            f.write(f"{s}\n" if finals[s] == 0.0 else f"{s}\t{num(finals[s])},0\n")
    f.write("\n")


def raw_lattice_from_canonical(c: dict) -> dict:
    """The flat-array form from the canonical rows of oracle.dec_oracle / ref_decoder (tests): states are
    numbered frame by frame with the (unique) frame-0 start token first."""
    st = c["states"]
    key = {(int(f), int(s)): i for i, (f, s) in enumerate(st[:, :2])}
    # state 0 must be the start state: the frame-0 token that no arc enters
    arcs = c["arcs"]
    entered = {(int(a[2]), int(a[3])) for a in arcs}
    start = [i for i, (f, s) in enumerate(st[:, :2]) if f == 0 and (int(f), int(s)) not in entered]
    perm = list(range(len(st)))
    if start and start[0] != 0:
        perm[0], perm[start[0]] = perm[start[0]], perm[0]
    inv = {old: new for new, old in enumerate(perm)}
    sid = {k: inv[v] for k, v in key.items()}
    st = st[perm]
    last = int(st[:, 0].max()) if len(st) else 0
    return dict(state_frame=st[:, 0].astype(np.int32), state_hclg=st[:, 1].astype(np.int32),
                state_tot_cost=st[:, 2].view(np.float32).copy(), state_extra_cost=st[:, 3].view(np.float32).copy(),
                arc_src=np.array([sid[(int(a[0]), int(a[1]))] for a in arcs], np.int32),
                arc_dst=np.array([sid[(int(a[2]), int(a[3]))] for a in arcs], np.int32),
                arc_ilabel=arcs[:, 4].astype(np.int32), arc_olabel=arcs[:, 5].astype(np.int32),
                arc_graph_cost=arcs[:, 6].view(np.float32).copy(), arc_acoustic_cost=arcs[:, 7].view(np.float32).copy(),
                final_state=np.array([sid[(last, int(r[0]))] for r in c["finals"]], np.int32),
                final_cost=c["finals"][:, 1].view(np.float32).copy())


# ---- raw lattice -> compact lattice (kaldi_b200/csrc/lattice_det.cu through the C ABI; host only) -------------------

def determinize_pruned(lat: dict, beam: float, max_states: int = 0, phones: dict | None = None, word_determinize: bool = True,
                       minimize: bool = False, delta: float = 1.0 / 1024.0) -> dict:
    """DeterminizeLatticePhonePrunedWrapper's role (lat/determinize-lattice-pruned.h:284) for one finalized raw
    lattice: returns the compact lattice as flat arrays — arc_src/arc_dst/arc_word/arc_graph_cost/arc_acoustic_cost,
    arc_tids (list of int32 arrays), final_state/final_graph_cost/final_acoustic_cost/final_tids, num_states (state 0
    = start) — plus `stats` (subsets expanded, elements).  See include/b2k.h b2k_lat_determinize_pruned.
    phones = dict(phone_of, self_loop, phone_start) over transition-ids: the two-pass form with the phone-level first pass
    (b2k_lat_determinize_phone_pruned).  minimize: push strings and weights, then merge equivalent states (b2k_clat_minimize:
    DeterminizeLatticePhonePrunedOptions::minimize, off by default in the reference too)."""
    import ctypes as C
    from . import _lib
    from .decoder import _RawLattice, _p
    L = _lib.lib()
    keep = {k: np.ascontiguousarray(lat[k], np.float32 if lat[k].dtype.kind == "f" else np.int32) for k in
            ("state_frame", "state_hclg", "state_tot_cost", "state_extra_cost", "arc_src", "arc_dst", "arc_ilabel",
             "arc_olabel", "arc_graph_cost", "arc_acoustic_cost", "final_state", "final_cost")}
    r = _RawLattice()
    r.num_states, r.num_arcs, r.num_finals = len(keep["state_frame"]), len(keep["arc_src"]), len(keep["final_state"])
    for k, v in keep.items():
        setattr(r, k, _p(v, C.c_float if v.dtype == np.float32 else C.c_int32))
    h = C.c_void_p()
    L.b2k_lat_determinize_pruned.argtypes = [C.c_void_p, C.c_float, C.c_int64, C.c_void_p]
    if phones is None:
        _lib.check(L.b2k_lat_determinize_pruned(C.byref(r), float(beam), int(max_states), C.byref(h)))
    else:
        po = np.ascontiguousarray(phones["phone_of"], np.int32)
        sl = np.ascontiguousarray(phones["self_loop"], np.uint8)
        ps = np.ascontiguousarray(phones["phone_start"], np.uint8)
        L.b2k_lat_determinize_phone_pruned.argtypes = [C.c_void_p, C.c_float, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32,
                                                       C.c_int32, C.c_int32, C.c_void_p]
        _lib.check(L.b2k_lat_determinize_phone_pruned(C.byref(r), float(beam), int(max_states), po.ctypes.data, sl.ctypes.data,
                                                      ps.ctypes.data, len(po), 1, int(word_determinize), C.byref(h)))
    if minimize:
        L.b2k_clat_minimize.argtypes = [C.c_void_p, C.c_float]
        rc = L.b2k_clat_minimize(h, float(delta))
        if rc:
            L.b2k_clat_destroy.argtypes = [C.c_void_p]
            L.b2k_clat_destroy(h)
            _lib.check(rc)
    L.b2k_clat_effective_beam.restype = C.c_float
    L.b2k_clat_effective_beam.argtypes = [C.c_void_p]
    eff = float(L.b2k_clat_effective_beam(h))
    try:
        sz = (C.c_int64 * 6)()
        L.b2k_clat_sizes.argtypes = [C.c_void_p, C.c_void_p]
        _lib.check(L.b2k_clat_sizes(h, sz))
        ns, na, nf, nt = int(sz[0]), int(sz[1]), int(sz[2]), int(sz[3])

        class _CL(C.Structure):
            _fields_ = [("num_states", C.c_int64), ("num_arcs", C.c_int64), ("num_finals", C.c_int64), ("num_tids", C.c_int64),
                        ("arc_src", C.c_void_p), ("arc_dst", C.c_void_p), ("arc_word", C.c_void_p),
                        ("arc_graph_cost", C.c_void_p), ("arc_acoustic_cost", C.c_void_p), ("arc_tids_off", C.c_void_p),
                        ("final_state", C.c_void_p), ("final_graph_cost", C.c_void_p), ("final_acoustic_cost", C.c_void_p),
                        ("final_tids_off", C.c_void_p), ("tids", C.c_void_p)]
        out = dict(arc_src=np.zeros(na, np.int32), arc_dst=np.zeros(na, np.int32), arc_word=np.zeros(na, np.int32),
                   arc_graph_cost=np.zeros(na, np.float32), arc_acoustic_cost=np.zeros(na, np.float32),
                   arc_tids_off=np.zeros(na + 1, np.int64), final_state=np.zeros(nf, np.int32),
                   final_graph_cost=np.zeros(nf, np.float32), final_acoustic_cost=np.zeros(nf, np.float32),
                   final_tids_off=np.zeros(nf + 1, np.int64), tids=np.zeros(nt, np.int32))
        cl = _CL()
        for k, v in out.items():
            setattr(cl, k, v.ctypes.data)
        L.b2k_clat_copy.argtypes = [C.c_void_p, C.c_void_p]
        _lib.check(L.b2k_clat_copy(h, C.byref(cl)))
    finally:
        L.b2k_clat_destroy.argtypes = [C.c_void_p]
        L.b2k_clat_destroy(h)
    t, ao, fo = out.pop("tids"), out.pop("arc_tids_off"), out.pop("final_tids_off")
    out["arc_tids"] = [t[ao[i]:ao[i + 1]] for i in range(na)]
    out["final_tids"] = [t[fo[i]:fo[i + 1]] for i in range(nf)]
    out["num_states"] = ns
    out["stats"] = dict(subsets_expanded=int(sz[4]), elements_total=int(sz[5]))
    out["effective_beam"] = eff
    return out


def compact_best_path(clat: dict) -> dict:
    """CompactLatticeShortestPath + the word / transition-id read-out of online2-wav-nnet3-latgen-faster.cc:43-54."""
    ns = clat["num_states"]
    if ns == 0 or len(clat["final_state"]) == 0:
        return dict(words=np.zeros(0, np.int32), tids=np.zeros(0, np.int32), graph_cost=np.inf, acoustic_cost=np.inf,
                    total_cost=np.inf)
    src, dst = clat["arc_src"].astype(np.int64), clat["arc_dst"].astype(np.int64)
    w = clat["arc_graph_cost"].astype(np.float64) + clat["arc_acoustic_cost"].astype(np.float64)
    dist, back = np.full(ns, np.inf), np.full(ns, -1, np.int64)
    dist[0] = 0.0
    order = np.argsort(src, kind="stable")
    starts = np.searchsorted(src[order], np.arange(ns + 1))
    for s in _topological_order(ns, src, dst):
        if np.isfinite(dist[s]):
            for a in order[starts[s]:starts[s + 1]]:
                if dist[s] + w[a] < dist[dst[a]]:
                    dist[dst[a]], back[dst[a]] = dist[s] + w[a], a
    fs = clat["final_state"].astype(np.int64)
    tot = dist[fs] + clat["final_graph_cost"].astype(np.float64) + clat["final_acoustic_cost"].astype(np.float64)
    k = int(np.argmin(tot))
    arcs, s = [], int(fs[k])
    while s != 0:
        arcs.append(int(back[s]))
        s = int(src[arcs[-1]])
    arcs = arcs[::-1]
    tids = [clat["arc_tids"][a] for a in arcs] + [clat["final_tids"][k]]
    return dict(words=clat["arc_word"][arcs].astype(np.int32),
                tids=np.concatenate(tids).astype(np.int32) if tids else np.zeros(0, np.int32),
                graph_cost=float(clat["arc_graph_cost"][arcs].astype(np.float64).sum() + clat["final_graph_cost"][k]),
                acoustic_cost=float(clat["arc_acoustic_cost"][arcs].astype(np.float64).sum() + clat["final_acoustic_cost"][k]),
                total_cost=float(tot[k]))


def write_compact_lattice_text(f, key: str, clat: dict, exact: bool = False) -> None:
    """One entry of a text-mode CompactLattice table (lat/kaldi-lattice.cc WriteCompactLattice, text mode): key line,
    `src dst word graph,acoustic,tid_tid_...` per arc, `state graph,acoustic,tid_...` per final state, blank line."""
    num = lambda x: _num(x, exact)
    f.write(key + "\n")
    order = np.argsort(clat["arc_src"], kind="stable")
    starts = np.searchsorted(clat["arc_src"][order], np.arange(clat["num_states"] + 1))
    fin = {int(s): i for i, s in enumerate(clat["final_state"])}
    for s in range(clat["num_states"]):
        for a in order[starts[s]:starts[s + 1]]:
            g, ac, t = float(clat["arc_graph_cost"][a]), float(clat["arc_acoustic_cost"][a]), clat["arc_tids"][a]
            w = "" if (g == 0.0 and ac == 0.0 and len(t) == 0) else f"\t{num(g)},{num(ac)},{'_'.join(str(int(x)) for x in t)}"   # One() omitted
            f.write(f"{s}\t{int(clat['arc_dst'][a])}\t{int(clat['arc_word'][a])}{w}\n")
        if s in fin:
            i = fin[s]
            if float(clat["final_graph_cost"][i]) == 0.0 and float(clat["final_acoustic_cost"][i]) == 0.0 and len(clat["final_tids"][i]) == 0:
                f.write(f"{s}\n")
                continue
            f.write(f"{s}\t{num(clat['final_graph_cost'][i])},{num(clat['final_acoustic_cost'][i])},"
                    f"{'_'.join(str(int(t)) for t in clat['final_tids'][i])}\n")
    f.write("\n")


# ---- binary table entries (lat/kaldi-lattice.cc WriteLattice / WriteCompactLattice with binary = true) ---------------
#
# `key` + ' ' + "\0B" + VectorFst<Arc>::Write: FstHeader {int32 magic 2125659606, "vector", arc type ("lattice4" =
# LatticeWeightTpl<float>::Type(), lattice-weight.h:86-89; "compactlattice44", :470-474), int32 version 2, int32 flags 0,
# uint64 properties, int64 start, int64 numstates, int64 numarcs 0}, then per state: final weight, int64 narcs, arcs
# {int32 ilabel, int32 olabel, weight, int32 nextstate}.  LatticeWeight = 2 floats (lattice-weight.h:141-146);
# CompactLatticeWeight = LatticeWeight + int32 length + int32 transition-ids (:531-540).  The weight encodings are
# checked against the reference's own Write() (oracle/_ref, tests/test_lattice_utils.py); the container layout is
# OpenFst's published one and, OpenFst being absent here, PARITY UNPINNED like the HCLG reader.

import struct as _struct

_FST_MAGIC = 2125659606
_INF32 = _struct.pack("<f", float("inf"))


def _fst_header(arctype: str, start: int, num_states: int) -> bytes:
    def s(x):
        return _struct.pack("<i", len(x)) + x.encode("ascii")
    return (_struct.pack("<i", _FST_MAGIC) + s("vector") + s(arctype) + _struct.pack("<ii", 2, 0) +
            _struct.pack("<Qqqq", 0x3, start, num_states, 0))          # properties: kExpanded | kMutable


def lattice_weight_bytes(g: float, a: float) -> bytes:
    return _struct.pack("<ff", g, a)


def compact_weight_bytes(g: float, a: float, tids) -> bytes:
    t = np.ascontiguousarray(tids, np.int32)
    return _struct.pack("<ff", g, a) + _struct.pack("<i", len(t)) + t.astype("<i4").tobytes()


def write_lattice_binary(f, key: str, lat: dict) -> None:
    """One binary `Lattice` table entry (a raw lattice: state 0 = start)."""
    ns = len(lat["state_frame"])
    f.write(key.encode() + b" \0B" + _fst_header("lattice4", 0 if ns else -1, ns))
    order = np.argsort(lat["arc_src"], kind="stable")
    starts = np.searchsorted(lat["arc_src"][order], np.arange(ns + 1))
    fin = dict(zip(lat["final_state"].tolist(), lat["final_cost"].tolist()))
    arc_dt = np.dtype([("i", "<i4"), ("o", "<i4"), ("g", "<f4"), ("a", "<f4"), ("n", "<i4")])
    for s in range(ns):
        f.write(lattice_weight_bytes(fin[s], 0.0) if s in fin else _INF32 * 2)
        idx = order[starts[s]:starts[s + 1]]
        f.write(_struct.pack("<q", len(idx)))
        rec = np.zeros(len(idx), arc_dt)
        rec["i"], rec["o"] = lat["arc_ilabel"][idx], lat["arc_olabel"][idx]
        rec["g"], rec["a"], rec["n"] = lat["arc_graph_cost"][idx], lat["arc_acoustic_cost"][idx], lat["arc_dst"][idx]
        f.write(rec.tobytes())


def write_compact_lattice_binary(f, key: str, clat: dict) -> None:
    """One binary `CompactLattice` table entry (acceptor: ilabel = olabel = word)."""
    ns = clat["num_states"]
    f.write(key.encode() + b" \0B" + _fst_header("compactlattice44", 0 if ns else -1, ns))
    order = np.argsort(clat["arc_src"], kind="stable")
    starts = np.searchsorted(clat["arc_src"][order], np.arange(ns + 1))
    fin = {int(s): i for i, s in enumerate(clat["final_state"])}
    for s in range(ns):
        if s in fin:
            i = fin[s]
            f.write(compact_weight_bytes(clat["final_graph_cost"][i], clat["final_acoustic_cost"][i], clat["final_tids"][i]))
        else:
            f.write(_INF32 * 2 + _struct.pack("<i", 0))
        idx = order[starts[s]:starts[s + 1]]
        f.write(_struct.pack("<q", len(idx)))
        for a in idx:
            w = int(clat["arc_word"][a])
            f.write(_struct.pack("<ii", w, w) + compact_weight_bytes(clat["arc_graph_cost"][a], clat["arc_acoustic_cost"][a],
                                                                     clat["arc_tids"][a]) + _struct.pack("<i", int(clat["arc_dst"][a])))


def read_lattice_archive(data: bytes) -> list:
    """Entries of a binary Lattice / CompactLattice archive as written above: [(key, 'lattice' | 'compact', dict)]."""
    out, p = [], 0
    while p < len(data):
        e = data.index(b" ", p)
        key = data[p:e].decode()
        if data[e + 1:e + 3] != b"\0B":
            raise ValueError("not a binary table entry")
        p = e + 3

        def rd(fmt):
            nonlocal p
            v = _struct.unpack_from(fmt, data, p)
            p += _struct.calcsize(fmt)
            return v

        def rs():
            nonlocal p
            (n,) = rd("<i")
            s = data[p:p + n].decode("ascii")
            p += n
            return s
        if rd("<i")[0] != _FST_MAGIC:
            raise ValueError("bad FST magic")
        fsttype, arctype = rs(), rs()
        _ver, _flags = rd("<ii")
        _props, start, ns, _na = rd("<Qqqq")
        if fsttype != "vector" or arctype not in ("lattice4", "compactlattice44"):
            raise ValueError(f"unsupported FST {fsttype}/{arctype}")
        compact = arctype.startswith("compact")

        def rw():
            g, a = rd("<ff")
            if not compact:
                return g, a, None
            (n,) = rd("<i")
            t = np.frombuffer(data, "<i4", n, p).copy()
            nonlocal_p_advance(4 * n)
            return g, a, t

        def nonlocal_p_advance(k):
            nonlocal p
            p += k
        src, dst, il, ol, g_, a_, tids, fs, fg, fa, ft = [], [], [], [], [], [], [], [], [], [], []
        for s in range(ns):
            g, a, t = rw()
            if np.isfinite(g) or np.isfinite(a):
                fs.append(s); fg.append(g); fa.append(a); ft.append(t)
            (na,) = rd("<q")
            for _ in range(na):
                i, o = rd("<ii")
                g, a, t = rw()
                (n,) = rd("<i")
                src.append(s); dst.append(n); il.append(i); ol.append(o); g_.append(g); a_.append(a); tids.append(t)
        i32, f32 = np.int32, np.float32
        if compact:
            d = dict(num_states=int(ns), arc_src=np.array(src, i32), arc_dst=np.array(dst, i32), arc_word=np.array(il, i32),
                     arc_graph_cost=np.array(g_, f32), arc_acoustic_cost=np.array(a_, f32), arc_tids=tids,
                     final_state=np.array(fs, i32), final_graph_cost=np.array(fg, f32), final_acoustic_cost=np.array(fa, f32),
                     final_tids=ft)
        else:
            d = dict(num_states=int(ns), arc_src=np.array(src, i32), arc_dst=np.array(dst, i32), arc_ilabel=np.array(il, i32),
                     arc_olabel=np.array(ol, i32), arc_graph_cost=np.array(g_, f32), arc_acoustic_cost=np.array(a_, f32),
                     final_state=np.array(fs, i32), final_cost=np.array(fg, f32))
        d["start"] = int(start)
        out.append((key, "compact" if compact else "lattice", d))
    return out


def scale_compact_lattice(clat: dict, graph_scale: float = 1.0, acoustic_scale: float = 1.0) -> dict:
    """fst::ScaleLattice with a diagonal scale (AcousticLatticeScale / GraphLatticeScale, fstext/lattice-utils.h): what
    online2-wav-nnet3-latgen-faster does with 1/acoustic_scale before writing (online2-wav-nnet3-latgen-faster.cc:289-291)."""
    out = dict(clat)
    for k, s in (("arc_graph_cost", graph_scale), ("final_graph_cost", graph_scale), ("arc_acoustic_cost", acoustic_scale),
                 ("final_acoustic_cost", acoustic_scale)):
        out[k] = (clat[k].astype(np.float32) * np.float32(s)).astype(np.float32)
    return out


def read_symbol_table(path: str) -> dict:
    """words.txt / phones.txt (fst::SymbolTable::ReadText: one `symbol id` pair per line, whitespace separated) as {id: symbol}:
    what online2-wav-nnet3-latgen-faster uses to print the transcript of the best path (online2-wav-nnet3-latgen-faster.cc:66-76)."""
    out = {}
    with open(path, encoding="utf-8") as f:
        for ln, line in enumerate(f, 1):
            parts = line.split()
            if not parts:
                continue
            if len(parts) != 2:
                raise ValueError(f"{path}:{ln}: expected `symbol id`")
            try:
                out[int(parts[1])] = parts[0]
            except ValueError:
                raise ValueError(f"{path}:{ln}: the id is not an integer") from None
    return out


def transcript(word_ids, symbols: dict) -> str:
    """The words of a best path, separated by single spaces, as the tool prints them after the utterance id."""
    missing = [int(w) for w in word_ids if int(w) not in symbols]
    if missing:
        raise KeyError(f"Word-id {missing[0]} not in symbol table.")              # the tool's own message (:73)
    return " ".join(symbols[int(w)] for w in word_ids)
