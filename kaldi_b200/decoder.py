"""Host-side mirror of the reference's GPU decoder surface over the b2k C-ABI.

Names and call order follow cudadecoder/cuda-fst.h:75-82 (CudaFst) and
cudadecoder/cuda-decoder.h:58-346 (CudaDecoderConfig, CudaDecoder); the search
semantics follow the CPU decoder (decoder/lattice-faster-decoder.cc), see
kaldi_b200/csrc/decoder.cu.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import _lib


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


class _FstCsr(C.Structure):
    _fields_ = [("num_states", C.c_int32), ("start", C.c_int32),
                ("offsets", C.POINTER(C.c_int32)), ("ilabel", C.POINTER(C.c_int32)),
                ("olabel", C.POINTER(C.c_int32)), ("weight", C.POINTER(C.c_float)),
                ("nextstate", C.POINTER(C.c_int32)), ("final_cost", C.POINTER(C.c_float)),
                ("tid2pdf", C.POINTER(C.c_int32)), ("num_tids", C.c_int32)]


class _DecCfg(C.Structure):
    _fields_ = [("beam", C.c_float), ("lattice_beam", C.c_float), ("max_active", C.c_int32),
                ("min_active", C.c_int32), ("beam_delta", C.c_float), ("prune_interval", C.c_int32),
                ("prune_scale", C.c_float), ("max_tokens_per_frame", C.c_int32),
                ("max_frames", C.c_int32), ("max_tokens", C.c_int64), ("max_links", C.c_int64),
                ("reference_order", C.c_int32), ("hash_ratio", C.c_float),
                ("max_arcs_per_frame", C.c_int32), ("max_lattice_states", C.c_int32),
                ("max_lattice_arcs", C.c_int32)]


class _RawLattice(C.Structure):
    _fields_ = [("num_states", C.c_int64), ("num_arcs", C.c_int64), ("num_finals", C.c_int64),
                ("state_frame", C.POINTER(C.c_int32)), ("state_hclg", C.POINTER(C.c_int32)),
                ("state_tot_cost", C.POINTER(C.c_float)), ("state_extra_cost", C.POINTER(C.c_float)),
                ("arc_src", C.POINTER(C.c_int32)), ("arc_dst", C.POINTER(C.c_int32)),
                ("arc_ilabel", C.POINTER(C.c_int32)), ("arc_olabel", C.POINTER(C.c_int32)),
                ("arc_graph_cost", C.POINTER(C.c_float)), ("arc_acoustic_cost", C.POINTER(C.c_float)),
                ("final_state", C.POINTER(C.c_int32)), ("final_cost", C.POINTER(C.c_float))]


class CudaFst:
    """cuda_decoder::CudaFst (cuda-fst.h:75-82): device CSR of the HCLG."""

    def __init__(self, graph: dict):
        L = _lib.lib()
        keep = [np.ascontiguousarray(graph[k]) for k in
                ("offsets", "ilabel", "olabel", "weight", "nextstate", "final", "tid2pdf")]
        off, il, ol, w, ns, fin, t2p = keep
        csr = _FstCsr(int(graph["num_states"]), int(graph["start"]), _p(off, C.c_int32),
                      _p(il, C.c_int32), _p(ol, C.c_int32), _p(w, C.c_float), _p(ns, C.c_int32),
                      _p(fin, C.c_float), _p(t2p, C.c_int32), int(t2p.size))
        self.h = C.c_void_p()
        _lib.check(L.b2k_fst_create(C.cast(C.byref(csr), C.c_void_p), C.byref(self.h)))
        self.num_pdfs = int(graph["num_pdfs"])

    @classmethod
    def from_file(cls, path: str, tid2pdf=None, num_pdfs: int | None = None):
        """HCLG.fst through the C++ reader (kaldi_b200/csrc/fst_io.cu): b2k_fst_file_read -> b2k_fst_create_from_file."""
        L = _lib.lib()
        fh = C.c_void_p()
        L.b2k_fst_file_read.argtypes = [C.c_char_p, C.c_void_p]
        _lib.check(L.b2k_fst_file_read(str(path).encode(), C.byref(fh)))
        self = cls.__new__(cls)
        self.h = C.c_void_p()
        try:
            t2p = None if tid2pdf is None else np.ascontiguousarray(tid2pdf, np.int32)
            L.b2k_fst_create_from_file.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]
            _lib.check(L.b2k_fst_create_from_file(fh, None if t2p is None else t2p.ctypes.data,
                                                  0 if t2p is None else int(t2p.size), C.byref(self.h)))
        finally:
            L.b2k_fst_file_destroy.argtypes = [C.c_void_p]
            L.b2k_fst_file_destroy(fh)
        self.num_pdfs = int(num_pdfs) if num_pdfs is not None else (int(t2p.max()) + 1 if t2p is not None else 0)
        return self

    def NumStates(self) -> int:
        return int(_lib.lib().b2k_fst_num_states(self.h))

    def Start(self) -> int:
        return int(_lib.lib().b2k_fst_start(self.h))

    def __del__(self):
        try:
            if self.h:
                _lib.lib().b2k_fst_destroy(self.h)
        except Exception:
            pass


@dataclass
class CudaDecoderConfig:
    """cuda_decoder::CudaDecoderConfig (cuda-decoder.h:58-163) + the
    LatticeFasterDecoderConfig fields that define the parity semantics."""
    default_beam: float = 15.0
    lattice_beam: float = 8.0
    max_active: int = 7000
    min_active: int = 200
    beam_delta: float = 0.5
    prune_interval: int = 25
    prune_scale: float = 0.1
    max_tokens_per_frame: int = 32768
    max_frames: int = 1024
    max_tokens: int = 3_000_000
    max_links: int = 6_000_000
    reference_order: bool = True     # bit-exact HashList-order emulation
    hash_ratio: float = 2.0
    max_arcs_per_frame: int = 1 << 20
    max_lattice_states: int = 131072
    max_lattice_arcs: int = 262144

    @classmethod
    def from_dict(cls, d: dict, **kw):
        return cls(default_beam=d["beam"], lattice_beam=d["lattice_beam"], max_active=d["max_active"],
                   min_active=d["min_active"], beam_delta=d["beam_delta"],
                   prune_interval=d["prune_interval"], prune_scale=d["prune_scale"],
                   hash_ratio=d.get("hash_ratio", 2.0), **kw)


class CudaDecoder:
    """cuda_decoder::CudaDecoder (cuda-decoder.h:171-346)."""

    def __init__(self, fst: CudaFst, config: CudaDecoderConfig, nlanes: int, nchannels: int | None = None):
        L = _lib.lib()
        nchannels = nlanes if nchannels is None else nchannels
        self.fst, self.config, self.nlanes, self.nchannels = fst, config, nlanes, nchannels
        c = _DecCfg(config.default_beam, config.lattice_beam, config.max_active, config.min_active,
                    config.beam_delta, config.prune_interval, config.prune_scale,
                    config.max_tokens_per_frame, config.max_frames, config.max_tokens, config.max_links,
                    int(config.reference_order), config.hash_ratio, config.max_arcs_per_frame,
                    config.max_lattice_states, config.max_lattice_arcs)
        self.h = C.c_void_p()
        _lib.check(L.b2k_dec_create(fst.h, C.cast(C.byref(c), C.c_void_p), nlanes, nchannels, C.byref(self.h)))

    def __del__(self):
        try:
            if self.h:
                _lib.lib().b2k_dec_destroy(self.h)
        except Exception:
            pass

    @staticmethod
    def _chan(channels):
        return np.ascontiguousarray(channels, dtype=np.int32)

    def InitDecoding(self, channels, stream: int = 0):
        ch = self._chan(channels)
        _lib.check(_lib.lib().b2k_dec_init_decoding(self.h, _p(ch, C.c_int32), len(ch), C.c_void_p(stream)))

    def AdvanceDecoding(self, lanes_assignments, stream: int = 0):
        """One frame: lanes_assignments = [(channel, device_ptr_to_loglike_row), ...]
        (cuda-decoder.h:264-265)."""
        ch = self._chan([c for c, _ in lanes_assignments])
        ptrs = (C.c_void_p * len(ch))(*[int(ptr) for _, ptr in lanes_assignments])
        _lib.check(_lib.lib().b2k_dec_advance_decoding(self.h, _p(ch, C.c_int32), C.cast(ptrs, C.c_void_p), len(ch),
                                                     C.c_void_p(stream)))

    def AdvanceDecodingFrames(self, channels, loglike_ptrs, num_frames, row_stride: int, stream: int = 0):
        ch = self._chan(channels)
        nf = np.ascontiguousarray(num_frames, dtype=np.int32)
        ptrs = (C.c_void_p * len(ch))(*[int(x) for x in loglike_ptrs])
        _lib.check(_lib.lib().b2k_dec_advance_decoding_frames(
            self.h, _p(ch, C.c_int32), C.cast(ptrs, C.c_void_p), _p(nf, C.c_int32), int(row_stride), len(ch),
            C.c_void_p(stream)))

    def FinalizeDecoding(self, channels, stream: int = 0):
        ch = self._chan(channels)
        _lib.check(_lib.lib().b2k_dec_finalize_decoding(self.h, _p(ch, C.c_int32), len(ch), C.c_void_p(stream)))

    def NumFramesDecoded(self, channel: int) -> int:
        out = C.c_int32()
        _lib.check(_lib.lib().b2k_dec_num_frames_decoded(self.h, int(channel), C.byref(out)))
        return out.value

    def ChannelInfo(self, channel: int) -> dict:
        info = (C.c_int64 * 32)()
        _lib.check(_lib.lib().b2k_dec_channel_info(self.h, int(channel), C.cast(info, C.POINTER(C.c_int64))))
        keys = ["status", "frames_decoded", "ntok", "nlink", "arcs_emitting", "arcs_nonemitting",
                "lat_states", "lat_arcs", "lat_finals", "finalized", "any_final", "err_line"]
        d = {k: int(info[i]) for i, k in enumerate(keys)}
        d["prof_cycles"] = [int(info[16 + k]) for k in range(16)]
        return d

    def GetRawLattice(self, channel: int, stream: int = 0) -> dict:
        """Finalized raw lattice as flat arrays (content of GetRawLattice,
        lattice-faster-decoder.cc:114-197)."""
        L = _lib.lib()
        r = _RawLattice()
        _lib.check(L.b2k_dec_get_raw_lattice(self.h, int(channel), C.cast(C.byref(r), C.c_void_p), C.c_void_p(stream)))
        ns, na, nf = r.num_states, r.num_arcs, r.num_finals
        out = dict(
            state_frame=np.zeros(ns, np.int32), state_hclg=np.zeros(ns, np.int32),
            state_tot_cost=np.zeros(ns, np.float32), state_extra_cost=np.zeros(ns, np.float32),
            arc_src=np.zeros(na, np.int32), arc_dst=np.zeros(na, np.int32),
            arc_ilabel=np.zeros(na, np.int32), arc_olabel=np.zeros(na, np.int32),
            arc_graph_cost=np.zeros(na, np.float32), arc_acoustic_cost=np.zeros(na, np.float32),
            final_state=np.zeros(nf, np.int32), final_cost=np.zeros(nf, np.float32))
        for k, v in out.items():
            setattr(r, k, _p(v, C.c_float if v.dtype == np.float32 else C.c_int32))
        _lib.check(L.b2k_dec_get_raw_lattice(self.h, int(channel), C.cast(C.byref(r), C.c_void_p), C.c_void_p(stream)))
        return out

    def GetRawLattices(self, channels, stream: int = 0):
        """Batched read-back of finalized lattices: one pack kernel + one D2H."""
        L = _lib.lib()
        ch = self._chan(channels)
        n = len(ch)
        so = np.zeros(n + 1, np.int64); ao = np.zeros(n + 1, np.int64); fo = np.zeros(n + 1, np.int64)
        i64p = C.POINTER(C.c_int64)
        r = _RawLattice()
        args = (self.h, _p(ch, C.c_int32), n, C.cast(C.byref(r), C.c_void_p), so.ctypes.data_as(i64p),
                ao.ctypes.data_as(i64p), fo.ctypes.data_as(i64p), C.c_void_p(stream))
        _lib.check(L.b2k_dec_get_raw_lattices(*args))
        ns, na, nf = r.num_states, r.num_arcs, r.num_finals
        out = dict(
            state_frame=np.zeros(ns, np.int32), state_hclg=np.zeros(ns, np.int32),
            state_tot_cost=np.zeros(ns, np.float32), state_extra_cost=np.zeros(ns, np.float32),
            arc_src=np.zeros(na, np.int32), arc_dst=np.zeros(na, np.int32),
            arc_ilabel=np.zeros(na, np.int32), arc_olabel=np.zeros(na, np.int32),
            arc_graph_cost=np.zeros(na, np.float32), arc_acoustic_cost=np.zeros(na, np.float32),
            final_state=np.zeros(nf, np.int32), final_cost=np.zeros(nf, np.float32))
        for k, v in out.items():
            setattr(r, k, _p(v, C.c_float if v.dtype == np.float32 else C.c_int32))
        _lib.check(L.b2k_dec_get_raw_lattices(*args))
        out.update(state_offs=so, arc_offs=ao, final_offs=fo)
        return out

    @staticmethod
    def SplitLattices(packed: dict):
        """Views of the individual lattices inside a GetRawLattices result."""
        so, ao, fo = packed["state_offs"], packed["arc_offs"], packed["final_offs"]
        res = []
        for i in range(len(so) - 1):
            d = {}
            for k, v in packed.items():
                if k.startswith("state_") and not k.endswith("offs"):
                    d[k] = v[so[i]:so[i + 1]]
                elif k.startswith("arc_") and not k.endswith("offs"):
                    d[k] = v[ao[i]:ao[i + 1]]
                elif k.startswith("final_") and not k.endswith("offs"):
                    d[k] = v[fo[i]:fo[i + 1]]
            res.append(d)
        return res

    def GetBestPath(self, channels, use_final_probs: bool = True, cap: int = 8192, stream: int = 0):
        """Best path of channels that are still decoding (or finalized), without changing them:
        LatticeFasterOnlineDecoderTpl::GetBestPath (lattice-faster-online-decoder.cc:54-75) / CudaDecoder::GetBestPath
        (cuda-decoder.h:279).  One dictionary per channel: ilabels, olabels, graph / acoustic costs, the frame and HCLG
        state each arc leads to, final_cost, best_cost, final_relative_cost, end_state, num_frames."""
        L = _lib.lib()
        ch = self._chan(channels)
        n = len(ch)

        class _Info(C.Structure):
            _fields_ = [("status", C.c_int32), ("n_arcs", C.c_int32), ("end_state", C.c_int32), ("num_frames", C.c_int32),
                        ("final_cost", C.c_float), ("best_cost", C.c_float), ("final_relative_cost", C.c_float)]
        info = (_Info * n)()
        il = np.zeros((n, cap), np.int32); ol = np.zeros((n, cap), np.int32)
        gc = np.zeros((n, cap), np.float32); ac = np.zeros((n, cap), np.float32)
        fr = np.zeros((n, cap), np.int32); st = np.zeros((n, cap), np.int32)
        L.b2k_dec_best_path.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32] + [C.c_void_p] * 8
        _lib.check(L.b2k_dec_best_path(self.h, ch.ctypes.data, n, int(bool(use_final_probs)), int(cap), il.ctypes.data,
                                       ol.ctypes.data, gc.ctypes.data, ac.ctypes.data, fr.ctypes.data, st.ctypes.data,
                                       C.cast(info, C.c_void_p), C.c_void_p(stream)))
        out = []
        for i in range(n):
            k = info[i].n_arcs
            out.append(dict(ilabels=il[i, :k].copy(), olabels=ol[i, :k].copy(), graph_costs=gc[i, :k].copy(),
                            acoustic_costs=ac[i, :k].copy(), arc_frame=fr[i, :k].copy(), arc_state=st[i, :k].copy(),
                            final_cost=float(info[i].final_cost), best_cost=float(info[i].best_cost),
                            final_relative_cost=float(info[i].final_relative_cost), end_state=int(info[i].end_state),
                            num_frames=int(info[i].num_frames)))
        return out

    def DebugFrame(self, channel: int, frame_plus_one: int):
        L = _lib.lib()
        nt, nl = C.c_int64(), C.c_int64()
        _lib.check(L.b2k_dec_debug_frame(self.h, channel, frame_plus_one, None, None, C.byref(nt),
                                         None, C.byref(nl), 0, 0))
        ts = np.zeros(nt.value, np.int32); tc = np.zeros(nt.value, np.float32)
        lk = np.zeros((nl.value, 7), np.int32)
        _lib.check(L.b2k_dec_debug_frame(self.h, channel, frame_plus_one, _p(ts, C.c_int32),
                                         _p(tc, C.c_float), C.byref(nt), _p(lk, C.c_int32),
                                         C.byref(nl), nt.value, nl.value))
        return ts, tc, lk

    def FrameInfo(self, channel: int, T: int):
        cut = np.zeros(T, np.float32); co = np.zeros(T, np.float32); nt = np.zeros(T, np.int32)
        _lib.check(_lib.lib().b2k_dec_frame_info(self.h, channel, _p(cut, C.c_float), _p(co, C.c_float),
                                                 _p(nt, C.c_int32), T))
        return dict(cutoff=cut, cost_offset=co, ntoks=nt)


def lattice_to_canonical(lat: dict) -> dict:
    """Order-free canonical rows of a raw lattice (same form as
    oracle.dec_oracle.canonical_lattice; float fields as raw bits)."""
    sf, sh = lat["state_frame"], lat["state_hclg"]
    states = np.stack([sf, sh, lat["state_tot_cost"].view(np.int32),
                       lat["state_extra_cost"].view(np.int32)], axis=1) if len(sf) else np.zeros((0, 4), np.int32)
    src, dst = lat["arc_src"], lat["arc_dst"]
    arcs = np.stack([sf[src], sh[src], sf[dst], sh[dst], lat["arc_ilabel"], lat["arc_olabel"],
                     lat["arc_graph_cost"].view(np.int32), lat["arc_acoustic_cost"].view(np.int32)],
                    axis=1) if len(src) else np.zeros((0, 8), np.int32)
    fs = lat["final_state"]
    finals = np.stack([sh[fs], lat["final_cost"].view(np.int32)], axis=1) if len(fs) else np.zeros((0, 2), np.int32)

    def srt(m):
        return m[np.lexsort(m.T[::-1])] if m.shape[0] else m
    return dict(states=srt(states.astype(np.int32)), arcs=srt(arcs.astype(np.int32)),
                finals=srt(finals.astype(np.int32)))


def read_fst_file(path: str) -> dict:
    """The CSR arrays of an OpenFst binary file as the C++ reader sees them (b2k_fst_file_read / b2k_fst_file_csr);
    same dictionary as kaldi_io.read_openfst.  Host only."""
    L = _lib.lib()
    fh = C.c_void_p()
    L.b2k_fst_file_read.argtypes = [C.c_char_p, C.c_void_p]
    _lib.check(L.b2k_fst_file_read(str(path).encode(), C.byref(fh)))
    try:
        csr, is_const = _FstCsr(), C.c_int32()
        L.b2k_fst_file_csr.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        _lib.check(L.b2k_fst_file_csr(fh, C.byref(csr), C.byref(is_const)))
        S = csr.num_states
        offsets = np.ctypeslib.as_array(csr.offsets, shape=(S + 1,)).copy()
        A = int(offsets[-1])

        def arr(ptr, n):
            return np.ctypeslib.as_array(ptr, shape=(n,)).copy() if n else np.zeros(0, np.int32)
        return dict(num_states=int(S), start=int(csr.start), offsets=offsets, ilabel=arr(csr.ilabel, A),
                    olabel=arr(csr.olabel, A), weight=arr(csr.weight, A).astype(np.float32, copy=False),
                    nextstate=arr(csr.nextstate, A), final=arr(csr.final_cost, S).astype(np.float32, copy=False),
                    fst_type="const" if is_const.value else "vector")
    finally:
        L.b2k_fst_file_destroy.argtypes = [C.c_void_p]
        L.b2k_fst_file_destroy(fh)
