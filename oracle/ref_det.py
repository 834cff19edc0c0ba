"""TEST INFRASTRUCTURE ONLY.  The reference's OWN lattice determinizer (lat/determinize-lattice-pruned.cc, compiled
where it lies under /root/reference against the container-only OpenFst stand-in of oracle/ref_wrap/fst_stub_det/) as a
callable oracle: determinize(raw_lattice, beam, phone_determinize=..., ...) -> compact lattice in the same dictionary
form as kaldi_b200.lattice.determinize_pruned.  Built into oracle/_ref/libkaldi_ref_det.so (travels to the GPU box)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from . import ref_feat

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(ref_feat.OUT_DIR, "libkaldi_ref_det.so")
BASE_SOURCES = ["base/" + f for f in ("kaldi-error.cc", "kaldi-math.cc", "kaldi-utils.cc", "io-funcs.cc", "timer.cc")]


def build(quiet: bool = False, force: bool = False) -> str:
    wrap = os.path.join(HERE, "ref_wrap", "det_wrap.cc")
    stub = os.path.join(HERE, "ref_wrap", "fst_stub_det")
    newest = max([os.path.getmtime(wrap)] + [os.path.getmtime(os.path.join(dp, f)) for dp, _, fs in os.walk(stub) for f in fs])
    if not force and os.path.exists(SO) and os.path.getmtime(SO) >= newest:
        return SO
    if not os.path.isdir(ref_feat.SRC):
        raise RuntimeError("/root/reference not present: cannot (re)build oracle/_ref")
    flags = ref_feat.cxxflags()
    flags = [f for f in flags if not f.startswith("-I")] + ["-I" + stub] + [f for f in flags if f.startswith("-I")]
    objdir = os.path.join(ref_feat.OUT_DIR, "obj_det")
    for f in ("lat_determinize-lattice-pruned.o", "lat_push-lattice.o", "lat_minimize-lattice.o"):
        o = os.path.join(objdir, f)         # they depend on the stand-in headers, which the object cache does not track
        if os.path.exists(o) and os.path.getmtime(o) < newest:
            os.remove(o)
    objs = ref_feat.compile_objects(BASE_SOURCES + ["lat/determinize-lattice-pruned.cc", "lat/push-lattice.cc", "lat/minimize-lattice.cc"],
                                    objdir, flags, quiet)
    wobj = os.path.join(objdir, "det_wrap.o")
    subprocess.check_call(["g++"] + flags + ["-c", wrap, "-o", wobj])
    subprocess.check_call(["g++", "-shared", "-o", SO] + objs + [wobj, "-lpthread", "-lm", "-ldl"])
    return SO


def available() -> bool:
    return os.path.exists(SO) or os.path.isdir(ref_feat.SRC)


_lib = None


def lib():
    global _lib
    if _lib is None:
        path = SO if (os.path.exists(SO) and not os.path.isdir(ref_feat.SRC)) else build(quiet=True)
        _lib = C.CDLL(path)
        _lib.ref_det_run.restype = C.c_void_p
        _lib.ref_det_run.argtypes = [C.c_int32, C.c_int64] + [C.c_void_p] * 6 + [C.c_int64, C.c_void_p, C.c_void_p, C.c_double,
                                                                                     C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32]
        _lib.ref_det_sizes.argtypes = [C.c_void_p, C.c_void_p]
        _lib.ref_det_copy.argtypes = [C.c_void_p] * 12
        _lib.ref_det_free.argtypes = [C.c_void_p]
    return _lib


def determinize(lat: dict, beam: float, phone_determinize: bool = False, phone_of=None, self_loop=None, phone_start=None,
                max_mem: int = 0, minimize: bool = False) -> dict:
    L = lib()
    k = {n: np.ascontiguousarray(lat[n], np.float32 if lat[n].dtype.kind == "f" else np.int32) for n in
         ("arc_src", "arc_dst", "arc_ilabel", "arc_olabel", "arc_graph_cost", "arc_acoustic_cost", "final_state", "final_cost")}
    ph = lo = st = None
    n_tids = 0
    if phone_determinize:
        ph = np.ascontiguousarray(phone_of, np.int32)
        lo = np.ascontiguousarray(self_loop, np.uint8)
        st = np.ascontiguousarray(phone_start, np.uint8)
        n_tids = len(ph)
    p = lambda a: None if a is None else a.ctypes.data
    h = L.ref_det_run(len(lat["state_frame"]), len(k["arc_src"]), p(k["arc_src"]), p(k["arc_dst"]), p(k["arc_ilabel"]),
                      p(k["arc_olabel"]), p(k["arc_graph_cost"]), p(k["arc_acoustic_cost"]), len(k["final_state"]),
                      p(k["final_state"]), p(k["final_cost"]), float(beam), int(phone_determinize), p(ph), p(lo), p(st),
                      n_tids, int(max_mem), int(minimize))
    try:
        sz = (C.c_int64 * 5)()
        L.ref_det_sizes(h, sz)
        ns, na, nf, nt, ok = [int(x) for x in sz]
        out = dict(arc_src=np.zeros(na, np.int32), arc_dst=np.zeros(na, np.int32), arc_word=np.zeros(na, np.int32),
                   arc_graph_cost=np.zeros(na, np.float32), arc_acoustic_cost=np.zeros(na, np.float32))
        ao, fo, t = np.zeros(na + 1, np.int64), np.zeros(nf + 1, np.int64), np.zeros(nt, np.int32)
        fin = dict(final_state=np.zeros(nf, np.int32), final_graph_cost=np.zeros(nf, np.float32),
                   final_acoustic_cost=np.zeros(nf, np.float32))
        L.ref_det_copy(h, p(out["arc_src"]), p(out["arc_dst"]), p(out["arc_word"]), p(out["arc_graph_cost"]),
                       p(out["arc_acoustic_cost"]), p(ao), p(fin["final_state"]), p(fin["final_graph_cost"]),
                       p(fin["final_acoustic_cost"]), p(fo), p(t))
    finally:
        L.ref_det_free(h)
    out.update(fin)
    out["arc_tids"] = [t[ao[i]:ao[i + 1]] for i in range(na)]
    out["final_tids"] = [t[fo[i]:fo[i + 1]] for i in range(nf)]
    out["num_states"] = ns
    out["ok"] = ok
    return out
