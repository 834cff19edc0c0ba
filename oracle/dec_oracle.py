"""ctypes wrapper over oracle/decoder_oracle.cc — TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import this module (see oracle/README.md).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libb2k_oracle_dec.so")

MODE_REFERENCE_ORDER = 0
MODE_ORDER_FREE = 1


class _Cfg(C.Structure):
    _fields_ = [("beam", C.c_float), ("max_active", C.c_int32), ("min_active", C.c_int32),
                ("lattice_beam", C.c_float), ("prune_interval", C.c_int32),
                ("beam_delta", C.c_float), ("hash_ratio", C.c_float), ("prune_scale", C.c_float)]


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "decoder_oracle.cc")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        os.makedirs(os.path.dirname(_SO), exist_ok=True)
        subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-std=c++17", "-fPIC",
                               "-shared", "-o", _SO, src])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.b2k_oracle_dec_create.restype = C.c_void_p
    return _lib


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


class DecoderOracle:
    """CPU LatticeFasterDecoder restatement over a CSR graph dict (synth.make_hclg)."""

    def __init__(self, graph: dict, cfg: dict):
        L = lib()
        self.g = graph
        c = _Cfg(cfg["beam"], cfg["max_active"], cfg["min_active"], cfg["lattice_beam"],
                 cfg["prune_interval"], cfg["beam_delta"], cfg["hash_ratio"], cfg["prune_scale"])
        g = graph
        self._keep = [np.ascontiguousarray(g[k]) for k in
                      ("offsets", "ilabel", "olabel", "weight", "nextstate", "final", "tid2pdf")]
        off, il, ol, w, ns, fin, t2p = self._keep
        self.h = C.c_void_p(L.b2k_oracle_dec_create(
            C.c_int32(g["num_states"]), C.c_int32(g["start"]), _p(off, C.c_int32),
            _p(il, C.c_int32), _p(ol, C.c_int32), _p(w, C.c_float), _p(ns, C.c_int32),
            _p(fin, C.c_float), _p(t2p, C.c_int32), C.c_int32(t2p.size), C.byref(c)))
        self.T = 0

    def __del__(self):
        try:
            lib().b2k_oracle_dec_destroy(self.h)
        except Exception:
            pass

    def decode(self, loglikes: np.ndarray, mode: int = MODE_REFERENCE_ORDER,
               record_frames: bool = False, finalize: bool = True):
        ll = np.ascontiguousarray(loglikes, dtype=np.float32)
        self.T = ll.shape[0]
        lib().b2k_oracle_dec_decode(self.h, _p(ll, C.c_float), C.c_int32(ll.shape[0]),
                                    C.c_int32(ll.shape[1]), C.c_int32(mode),
                                    C.c_int32(int(record_frames)), C.c_int32(int(finalize)))
        return self.stats()

    def stats(self) -> dict:
        out = np.zeros(12, dtype=np.int64)
        lib().b2k_oracle_dec_stats(self.h, _p(out, C.c_int64))
        keys = ["arcs_emitting", "arcs_nonemitting", "tokens_expanded", "extra_links",
                "best_ties", "min_active_branch", "max_active_branch", "links_admitted",
                "toks_created", "lat_states", "lat_arcs", "lat_finals"]
        return dict(zip(keys, out.tolist()))

    def frame_info(self):
        cut = np.zeros(self.T, dtype=np.float32)
        nt = np.zeros(self.T, dtype=np.int32)
        co = np.zeros(self.T, dtype=np.float32)
        lib().b2k_oracle_dec_frame_info(self.h, _p(cut, C.c_float), _p(nt, C.c_int32),
                                        _p(co, C.c_float))
        return dict(cutoff=cut, ntoks=nt, cost_offset=co)

    def lattice(self) -> dict:
        st = self.stats()
        ns, na, nf = st["lat_states"], st["lat_arcs"], st["lat_finals"]
        sf = np.zeros(ns, np.int32); ss = np.zeros(ns, np.int32)
        tot = np.zeros(ns, np.float32); ext = np.zeros(ns, np.float32)
        arcs = np.zeros((na, 8), np.int32)
        fs = np.zeros(nf, np.int32); fc = np.zeros(nf, np.float32)
        lib().b2k_oracle_dec_lattice(self.h, _p(sf, C.c_int32), _p(ss, C.c_int32),
                                     _p(tot, C.c_float), _p(ext, C.c_float),
                                     _p(arcs, C.c_int32), _p(fs, C.c_int32), _p(fc, C.c_float))
        return canonical_lattice(sf, ss, tot, ext, arcs, fs, fc)

    def raw_frame(self, frame_plus_one: int) -> dict:
        nt = C.c_int64(); nl = C.c_int64()
        lib().b2k_oracle_dec_raw_sizes(self.h, C.c_int32(frame_plus_one), C.byref(nt), C.byref(nl))
        ts = np.zeros(nt.value, np.int32); tc = np.zeros(nt.value, np.float32)
        lk = np.zeros((nl.value, 7), np.int32)
        lib().b2k_oracle_dec_raw_copy(self.h, C.c_int32(frame_plus_one), _p(ts, C.c_int32),
                                      _p(tc, C.c_float), _p(lk, C.c_int32))
        return canonical_raw_frame(ts, tc, lk)


def _sort_rows(m: np.ndarray) -> np.ndarray:
    if m.shape[0] == 0:
        return m
    order = np.lexsort(m.T[::-1])
    return m[order]


def canonical_lattice(state_frame, state_state, state_tot, state_extra, arcs8, final_state,
                      final_cost) -> dict:
    """Canonical (order-free) form of a finalized raw lattice (SURVEY.md §7 hard
    part 1): sorted integer rows with float fields as raw bits."""
    states = np.stack([state_frame.astype(np.int32), state_state.astype(np.int32),
                       np.ascontiguousarray(state_tot, np.float32).view(np.int32),
                       np.ascontiguousarray(state_extra, np.float32).view(np.int32)], axis=1) \
        if len(state_frame) else np.zeros((0, 4), np.int32)
    finals = np.stack([final_state.astype(np.int32),
                       np.ascontiguousarray(final_cost, np.float32).view(np.int32)], axis=1) \
        if len(final_state) else np.zeros((0, 2), np.int32)
    return dict(states=_sort_rows(states), arcs=_sort_rows(np.asarray(arcs8, np.int32).reshape(-1, 8)),
                finals=_sort_rows(finals))


def canonical_raw_frame(tok_state, tok_cost, links7) -> dict:
    toks = np.stack([tok_state.astype(np.int32),
                     np.ascontiguousarray(tok_cost, np.float32).view(np.int32)], axis=1) \
        if len(tok_state) else np.zeros((0, 2), np.int32)
    return dict(toks=_sort_rows(toks), links=_sort_rows(np.asarray(links7, np.int32).reshape(-1, 7)))


def lattices_equal(a: dict, b: dict) -> bool:
    return all(a[k].shape == b[k].shape and np.array_equal(a[k], b[k]) for k in ("states", "arcs", "finals"))
