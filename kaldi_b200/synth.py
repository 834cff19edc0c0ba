"""Seeded synthetic inputs for the online2 hot path.

No trained models, HCLG graphs or audio exist in the reference tree (SURVEY.md
§7 hard part 8), so every config is driven by generators with the shapes the
recipes name.  Everything here is plain numpy and deterministic in `seed`;
it is shared by the tests, bench.py and the oracle harness and contains no
algorithm from the hot path itself.
"""
from __future__ import annotations

import numpy as np

__all__ = ["make_hclg", "make_loglikes", "make_audio", "tiny_graph", "DEFAULT_DECODER_CFG"]

# recipe decode settings: egs/wsj/s5/steps/online/nnet3/decode.sh:13-16 and
# LatticeFasterDecoderConfig defaults (decoder/lattice-faster-decoder.h:38-106)
DEFAULT_DECODER_CFG = dict(
    beam=15.0, max_active=7000, min_active=200, lattice_beam=8.0,
    prune_interval=25, beam_delta=0.5, hash_ratio=2.0, prune_scale=0.1,
)


def make_hclg(num_arcs: int, num_pdfs: int = 2336, seed: int = 0,
              eps_frac: float = 0.12, selfloop_frac: float = 0.9,
              olabel_frac: float = 0.05, final_frac: float = 0.02,
              weight_max: float = 1.5):
    """HCLG-like decoding graph in ConstFst order (arcs of a state contiguous,
    in 'file order').

    states = arcs/2.7; ~90 % of states carry an emitting self-loop whose
    transition-id shares the pdf of the arcs entering the state (as in a real
    HCLG); other arcs have geometric out-degree plus a few hub states with
    1e3-1e4 arcs (word-start fan-out); `eps_frac` of the non-loop arcs are
    input-epsilon and always point to a higher-numbered state (no eps cycles,
    lattice-faster-decoder.cc:995).  transition-ids: 2*pdf+1 = forward,
    2*pdf+2 = self-loop; tid2pdf[0] is unused.

    Defaults (weight_max=1.5 with make_loglikes sigma=0.7, boost=1.5) were
    tuned with the oracle so that beam-15 decoding keeps 2-8 k tokens and
    ~10-15 k arcs per frame with max_active=7000 firing on some frames -- the
    "typical magnitudes" of SURVEY.md §8a.  SURVEY §8d's first guess
    (weights U(0,8), sigma 2, boost 8) collapses to <500 tokens/frame.
    """
    rng = np.random.default_rng(seed)
    N = max(int(num_arcs / 2.7), 16)
    pdf_of_state = rng.integers(0, num_pdfs, size=N, dtype=np.int64)
    has_self = rng.random(N) < selfloop_frac
    n_self = int(has_self.sum())
    # hubs
    n_hubs = max(2, N // 150_000)
    hub_ids = np.unique(np.concatenate([[0], rng.integers(0, N, size=n_hubs)]))
    hub_deg = rng.integers(1000, 10_000, size=hub_ids.size)
    hub_deg = np.minimum(hub_deg, max(4, num_arcs // (4 * hub_ids.size)))
    remaining = max(num_arcs - n_self - int(hub_deg.sum()), N)
    mean_extra = remaining / N
    p = 1.0 / (1.0 + mean_extra)
    deg = rng.geometric(p, size=N).astype(np.int64) - 1
    # make sure no state without a self loop is a dead end
    deg = np.where(~has_self & (deg == 0), 1, deg)
    deg[hub_ids] += hub_deg
    tot_deg = deg + has_self.astype(np.int64)
    offsets = np.zeros(N + 1, dtype=np.int64)
    np.cumsum(tot_deg, out=offsets[1:])
    A = int(offsets[-1])
    assert A < 2**31 - 1
    src = np.repeat(np.arange(N, dtype=np.int64), tot_deg)
    pos = np.arange(A, dtype=np.int64) - offsets[src]
    is_self = has_self[src] & (pos == 0)
    # destinations
    dst = rng.integers(0, N, size=A, dtype=np.int64)
    is_eps = (rng.random(A) < eps_frac) & ~is_self
    # eps arcs go strictly forward in state numbering
    span = (N - 1 - src)
    can_eps = span > 0
    is_eps &= can_eps
    fwd = src + 1 + (rng.random(A) * np.maximum(span, 1)).astype(np.int64)
    fwd = np.minimum(fwd, N - 1)
    dst = np.where(is_eps, fwd, dst)
    dst = np.where(is_self, src, dst)
    ilabel = np.where(is_self, 2 * pdf_of_state[src] + 2, 2 * pdf_of_state[dst] + 1)
    ilabel = np.where(is_eps, 0, ilabel)
    weight = rng.random(A, dtype=np.float32) * np.float32(weight_max)
    # self-loops are cheap (-log of a self-loop prob ~0.5-0.9)
    weight = np.where(is_self, rng.random(A, dtype=np.float32) * np.float32(0.7) + np.float32(0.1),
                      weight).astype(np.float32)
    olabel = np.where((rng.random(A) < olabel_frac) & ~is_self,
                      rng.integers(1, 200_000, size=A), 0)
    final = np.full(N, np.inf, dtype=np.float32)
    fin = rng.random(N) < final_frac
    final[fin] = (rng.random(int(fin.sum()), dtype=np.float32) * np.float32(5.0))
    tid2pdf = np.zeros(2 * num_pdfs + 1, dtype=np.int32)
    tid2pdf[1:] = (np.arange(1, 2 * num_pdfs + 1) - 1) // 2
    return dict(
        num_states=N, start=0, num_pdfs=num_pdfs,
        offsets=offsets.astype(np.int32),
        ilabel=ilabel.astype(np.int32), olabel=olabel.astype(np.int32),
        weight=weight, nextstate=dst.astype(np.int32), final=final,
        tid2pdf=tid2pdf,
    )


def make_loglikes(graph: dict, num_frames: int, seed: int = 0, sigma: float = 0.7,
                  boost: float = 1.5) -> np.ndarray:
    """[T, num_pdfs] float32 pseudo log-likelihoods: N(0, sigma^2) with the pdf
    of one random path through the graph boosted by `boost` on each frame so a
    best path exists."""
    rng = np.random.default_rng(seed + 7919)
    P = int(graph["num_pdfs"])
    ll = (rng.standard_normal((num_frames, P), dtype=np.float32) * np.float32(sigma))
    off, il, ns, t2p = graph["offsets"], graph["ilabel"], graph["nextstate"], graph["tid2pdf"]
    s = int(graph["start"])
    for t in range(num_frames):
        # follow eps arcs (bounded) until an emitting arc is chosen
        for _ in range(8):
            a0, a1 = int(off[s]), int(off[s + 1])
            if a1 == a0:
                s = int(graph["start"])
                continue
            a = int(rng.integers(a0, a1))
            if il[a] == 0:
                s = int(ns[a])
                continue
            ll[t, t2p[il[a]]] += np.float32(boost)
            s = int(ns[a])
            break
    return ll


def make_audio(num_samples: int, seed: int = 0, sample_rate: float = 16000.0) -> np.ndarray:
    """Synthetic 16 kHz utterance in Kaldi's int16-range float convention
    (feat/wave-reader.h:60-62).  Speech-like non-stationarity: a sequence of
    40-160 ms segments, each with its own 5 partials in [100, 4000] Hz (phase
    continuous), 4 Hz amplitude modulation, N(0, 300^2) noise (SURVEY.md §8d asks
    for frames that differ; a stationary tone mixture makes the acoustic scores
    constant in time and the beam collapses onto self-loops)."""
    rng = np.random.default_rng(seed + 104729)
    n = int(num_samples)
    bounds = [0]
    while bounds[-1] < n:
        bounds.append(bounds[-1] + int(rng.uniform(0.04, 0.16) * sample_rate))
    nseg = len(bounds) - 1
    f = rng.uniform(100.0, 4000.0, size=(nseg, 5))
    a = rng.uniform(0.2, 1.0, size=(nseg, 5))
    seg = np.searchsorted(np.asarray(bounds[1:]), np.arange(n), side="right")
    x = np.zeros(n, dtype=np.float64)
    for k in range(5):
        phase = 2 * np.pi * np.cumsum(f[seg, k]) / sample_rate + rng.uniform(0, 2 * np.pi)
        x += a[seg, k] * np.sin(phase)
    t = np.arange(n, dtype=np.float64) / sample_rate
    env = 0.6 + 0.4 * np.sin(2 * np.pi * 4.0 * t + rng.uniform(0, 2 * np.pi))
    x = 3000.0 * env * x + rng.standard_normal(n) * 300.0
    x = np.clip(np.round(x), -32768, 32767)
    return x.astype(np.float32)


def tiny_graph() -> dict:
    """3-state hand-built graph with a known best path (config 1 style plumbing):
    0 --(tid1/pdf0, w=1, o=7)--> 1 ; 0 --(eps, w=0.5)--> 2 ; 2 --(tid2/pdf1, w=0.25, o=9)--> 1 ;
    1 --(tid1 self-loop, w=0.1)--> 1 ; state 1 final with cost 0."""
    return dict(num_states=3, start=0, num_pdfs=2,
                offsets=np.array([0, 2, 3, 4], np.int32),
                ilabel=np.array([1, 0, 1, 2], np.int32), olabel=np.array([7, 0, 0, 9], np.int32),
                weight=np.array([1.0, 0.5, 0.1, 0.25], np.float32),
                nextstate=np.array([1, 2, 1, 1], np.int32),
                final=np.array([np.inf, 0.0, np.inf], np.float32),
                tid2pdf=np.array([0, 0, 1], np.int32))
