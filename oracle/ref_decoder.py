"""oracle/_ref decoder library: the reference's OWN decoder search code
(decoder/lattice-faster-decoder.{h,cc}, util/hash-list{,-inl}.h, itf/decodable-itf.h,
base/*) compiled from where it lies under /root/reference/src, against a
container-only stand-in for the absent OpenFst (oracle/ref_wrap/fst_stub/: Fst /
ArcIterator / MemoryPool / Lattice containers, no search logic), plus our wrapper
oracle/ref_wrap/decoder_wrap.cc.  TEST INFRASTRUCTURE ONLY.

This is what pins oracle/decoder_oracle.cc (the restatement) to the reference:
tests/test_decoder_oracle.py compares, bit for bit, the token list of every frame
IN HASHLIST ORDER and the finalized raw lattice of the two on seeded inputs.

Python side: RefDecoder (same calling convention as dec_oracle.DecoderOracle).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from . import ref_feat

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(ref_feat.OUT_DIR, "libkaldi_ref_decoder.so")
BASE_SOURCES = ["base/" + f for f in ("kaldi-error.cc", "kaldi-math.cc", "kaldi-utils.cc", "io-funcs.cc", "timer.cc")]


def build(quiet: bool = False, force: bool = False) -> str:
    wrap = os.path.join(HERE, "ref_wrap", "decoder_wrap.cc")
    stub = os.path.join(HERE, "ref_wrap", "fst_stub")
    newest = max([os.path.getmtime(wrap)] + [os.path.getmtime(os.path.join(dp, f))
                                              for dp, _, fs in os.walk(stub) for f in fs])
    if not force and os.path.exists(SO) and os.path.getmtime(SO) >= newest:
        return SO
    if not os.path.isdir(ref_feat.SRC):
        raise RuntimeError("/root/reference not present: cannot (re)build oracle/_ref")
    # the stand-in headers must shadow <fst/...>, lat/..., fstext/..., decoder/grammar-fst.h: first on the include path
    flags = ref_feat.cxxflags()
    flags = [f for f in flags if not f.startswith("-I")] + ["-I" + stub] + [f for f in flags if f.startswith("-I")]
    objdir = os.path.join(ref_feat.OUT_DIR, "obj_decoder")
    objs = ref_feat.compile_objects(BASE_SOURCES + ["decoder/lattice-faster-decoder.cc"], objdir, flags, quiet)
    wobj = os.path.join(objdir, "decoder_wrap.o")
    subprocess.check_call(["g++"] + flags + ["-c", wrap, "-o", wobj])
    subprocess.check_call(["g++", "-shared", "-o", SO] + objs + [wobj, "-lpthread", "-lm", "-ldl"])
    return SO


class _Cfg(C.Structure):
    _fields_ = [("beam", C.c_float), ("max_active", C.c_int32), ("min_active", C.c_int32),
                ("lattice_beam", C.c_float), ("prune_interval", C.c_int32), ("beam_delta", C.c_float),
                ("hash_ratio", C.c_float), ("prune_scale", C.c_float)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        path = SO if (os.path.exists(SO) and not os.path.isdir(ref_feat.SRC)) else build(quiet=True)
        _lib = C.CDLL(path)
        _lib.b2k_refdec_create.restype = C.c_void_p
        _lib.b2k_refdec_create.argtypes = [C.c_int32, C.c_int32] + [C.c_void_p] * 7 + [C.c_int32, C.c_void_p]
        _lib.b2k_refdec_destroy.argtypes = [C.c_void_p]
        _lib.b2k_refdec_decode.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32]
        _lib.b2k_refdec_decode.restype = C.c_int
        _lib.b2k_refdec_frame_size.argtypes = [C.c_void_p, C.c_int32, C.c_void_p]
        _lib.b2k_refdec_frame_copy.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]
        _lib.b2k_refdec_lattice_sizes.argtypes = [C.c_void_p] * 4
        _lib.b2k_refdec_lattice.argtypes = [C.c_void_p] * 8
    return _lib


def available() -> bool:
    return os.path.exists(SO) or os.path.isdir(ref_feat.SRC)


class RefDecoder:
    """kaldi::LatticeFasterDecoderTpl<ConstFst<StdArc>> (the reference's code) on a CSR graph dict
    (same keys as kaldi_b200.synth.make_hclg) and a [T, num_pdfs] log-likelihood matrix."""

    def __init__(self, g: dict, cfg: dict):
        a = {k: np.ascontiguousarray(g[k], dt) for k, dt in
             [("offsets", np.int32), ("ilabel", np.int32), ("olabel", np.int32), ("weight", np.float32),
              ("nextstate", np.int32), ("final", np.float32), ("tid2pdf", np.int32)]}
        self._keep = a
        c = _Cfg(cfg["beam"], cfg["max_active"], cfg["min_active"], cfg["lattice_beam"], cfg["prune_interval"],
                 cfg["beam_delta"], cfg.get("hash_ratio", 2.0), cfg["prune_scale"])
        self.h = lib().b2k_refdec_create(int(g["num_states"]), int(g["start"]), a["offsets"].ctypes.data,
                                         a["ilabel"].ctypes.data, a["olabel"].ctypes.data, a["weight"].ctypes.data,
                                         a["nextstate"].ctypes.data, a["final"].ctypes.data,
                                         a["tid2pdf"].ctypes.data, len(a["tid2pdf"]), C.addressof(c))
        self.T = 0

    def __del__(self):
        if getattr(self, "h", None):
            lib().b2k_refdec_destroy(self.h)
            self.h = None

    def decode(self, loglikes: np.ndarray, record: bool = True) -> None:
        """record=True keeps every frame's token list and the keyed lattice (parity tests);
        record=False is the plain InitDecoding / AdvanceDecoding / FinalizeDecoding sequence (timing)."""
        ll = np.ascontiguousarray(loglikes, np.float32)
        self.T = ll.shape[0]
        rc = lib().b2k_refdec_decode(self.h, ll.ctypes.data, ll.shape[0], ll.shape[1], int(record))
        if rc:
            raise RuntimeError("the reference decoder raised")

    def lattice_sizes(self):
        ns, na, nf = C.c_int64(), C.c_int64(), C.c_int64()
        lib().b2k_refdec_lattice_sizes(self.h, C.addressof(ns), C.addressof(na), C.addressof(nf))
        return ns.value, na.value, nf.value

    def frame_tokens(self, frame_plus_one: int):
        """(states, tot_costs) of the frame's token list in HashList iteration order."""
        n = C.c_int64()
        lib().b2k_refdec_frame_size(self.h, frame_plus_one, C.addressof(n))
        st = np.zeros(n.value, np.int32); co = np.zeros(n.value, np.float32)
        lib().b2k_refdec_frame_copy(self.h, frame_plus_one, st.ctypes.data, co.ctypes.data)
        return st, co

    def lattice(self) -> dict:
        from .dec_oracle import canonical_lattice
        ns, na, nf = C.c_int64(), C.c_int64(), C.c_int64()
        lib().b2k_refdec_lattice_sizes(self.h, C.addressof(ns), C.addressof(na), C.addressof(nf))
        sf = np.zeros(ns.value, np.int32); ss = np.zeros(ns.value, np.int32)
        tot = np.zeros(ns.value, np.float32); ext = np.zeros(ns.value, np.float32)
        arcs = np.zeros((na.value, 8), np.int32)
        fs = np.zeros(nf.value, np.int32); fc = np.zeros(nf.value, np.float32)
        lib().b2k_refdec_lattice(self.h, sf.ctypes.data, ss.ctypes.data, tot.ctypes.data, ext.ctypes.data,
                                 arcs.ctypes.data, fs.ctypes.data, fc.ctypes.data)
        return canonical_lattice(sf, ss, tot, ext, arcs, fs, fc)


if __name__ == "__main__":
    print(build(force=True))
