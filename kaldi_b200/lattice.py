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


def _num(x: float) -> str:
    return repr(float(np.float32(x))) if np.isfinite(x) else "Infinity"


def write_lattice_text(f, key: str, lat: dict) -> None:
    """One entry of a text-mode Lattice table (`ark,t`): arcs grouped by source state as an FST printer emits them."""
    f.write(key + "\n")
    src = lat["arc_src"]
    order = np.argsort(src, kind="stable")
    finals = dict(zip(lat["final_state"].tolist(), lat["final_cost"].tolist()))
    starts = np.searchsorted(src[order], np.arange(len(lat["state_frame"]) + 1))
    for s in range(len(lat["state_frame"])):
        for a in order[starts[s]:starts[s + 1]]:
            f.write(f"{s}\t{int(lat['arc_dst'][a])}\t{int(lat['arc_ilabel'][a])}\t{int(lat['arc_olabel'][a])}\t"
                    f"{_num(lat['arc_graph_cost'][a])},{_num(lat['arc_acoustic_cost'][a])}\n")
        if s in finals:
            f.write(f"{s}\t{_num(finals[s])},0\n")
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
