"""oracle/_ref nnet3 library: the reference's OWN nnet3 CPU forward compiled from
/root/reference/src (cudamatrix with HAVE_CUDA=0, nnet3 core, hmm, tree, gmm,
+ the base/matrix/util objects of ref_feat.py) with the 4-line fst/fst-decl.h
stub SURVEY.md §8c describes (OpenFst is absent; only forward declarations are
needed by hmm/transition-model.h).  TEST INFRASTRUCTURE ONLY."""
from __future__ import annotations

import os
import subprocess

from . import ref_feat as RF

HERE = os.path.dirname(os.path.abspath(__file__))
OUT_DIR = RF.OUT_DIR
SO = os.path.join(OUT_DIR, "libkaldi_ref_nnet3.so")

NNET_SOURCES = (
    ["cudamatrix/" + f for f in (
        "cu-allocator.cc", "cu-array.cc", "cu-block-matrix.cc", "cu-common.cc", "cu-compressed-matrix.cc",
        "cu-device.cc", "cu-math.cc", "cu-matrix.cc", "cu-packed-matrix.cc", "cu-rand.cc", "cu-sp-matrix.cc",
        "cu-sparse-matrix.cc", "cu-tp-matrix.cc", "cu-vector.cc")]
    + ["nnet3/" + f for f in (
        "nnet-common.cc", "nnet-component-itf.cc", "nnet-simple-component.cc", "nnet-normalize-component.cc",
        "nnet-general-component.cc", "nnet-combined-component.cc", "nnet-tdnn-component.cc",
        "nnet-convolutional-component.cc", "nnet-attention-component.cc", "attention.cc", "convolution.cc",
        "natural-gradient-online.cc", "nnet-parse.cc", "nnet-descriptor.cc", "nnet-nnet.cc", "nnet-graph.cc",
        "nnet-computation.cc", "nnet-computation-graph.cc", "nnet-compile.cc", "nnet-compile-utils.cc",
        "nnet-compile-looped.cc", "nnet-analyze.cc", "nnet-optimize.cc", "nnet-optimize-utils.cc",
        "nnet-compute.cc", "nnet-utils.cc", "am-nnet-simple.cc", "decodable-simple-looped.cc",
        "decodable-online-looped.cc", "nnet-am-decodable-simple.cc")]
    + ["hmm/" + f for f in ("transition-model.cc", "hmm-topology.cc", "posterior.cc")]
    + ["tree/" + f for f in ("build-tree-questions.cc", "build-tree-utils.cc", "build-tree.cc",
                             "cluster-utils.cc", "clusterable-classes.cc", "context-dep.cc",
                             "event-map.cc", "tree-renderer.cc")]
    + ["gmm/" + f for f in ("diag-gmm.cc", "am-diag-gmm.cc", "full-gmm.cc", "diag-gmm-normal.cc",
                            "full-gmm-normal.cc", "model-common.cc")]
    + ["ivector/ivector-extractor.cc"]
)
NNET_SOURCES = [s for s in NNET_SOURCES if s != "tree/tree-renderer.cc"]


def build(quiet: bool = False, force: bool = False) -> str:
    wrap = os.path.join(HERE, "ref_wrap", "nnet_wrap.cc")
    wraps = [wrap, os.path.join(HERE, "ref_wrap", "ivector_wrap.cc"), os.path.join(HERE, "ref_wrap", "nnet_stubs.cc"),
             os.path.join(HERE, "ref_wrap", "endpoint_wrap.cc"), os.path.join(HERE, "ref_wrap", "silence_wrap.cc"),
             os.path.join(HERE, "ref_wrap", "replay_decoder.h"), os.path.join(HERE, "ref_wrap", "ivector_ref_types.h")]
    if not force and os.path.exists(SO) and all(os.path.getmtime(SO) >= os.path.getmtime(w) for w in wraps):
        return SO
    if not os.path.isdir(RF.SRC):
        raise RuntimeError("/root/reference not present: cannot (re)build oracle/_ref")
    blas = RF.find_openblas()
    stub = os.path.join(OUT_DIR, "inc", "fst")
    os.makedirs(stub, exist_ok=True)
    with open(os.path.join(stub, "fst-decl.h"), "w") as f:
        f.write("// forward declarations only (OpenFst is not available in this container)\n"
                "#ifndef B2K_FST_DECL_STUB_H_\n#define B2K_FST_DECL_STUB_H_\n"
                "namespace fst { template <class A> class Fst; template <class W> class ArcTpl; "
                "template <class T> class TropicalWeightTpl; typedef TropicalWeightTpl<float> TropicalWeight; "
                "typedef ArcTpl<TropicalWeight> StdArc; template <class A, class S> class VectorFst; }\n#endif\n")
    flags = RF.cxxflags(["-DHAVE_CUDA=0"])
    base_objs = RF.compile_objects(RF.FEAT_SOURCES, os.path.join(OUT_DIR, "obj_feat"), RF.cxxflags(), quiet)
    objs = RF.compile_objects(NNET_SOURCES, os.path.join(OUT_DIR, "obj_nnet"), flags, quiet)
    wobj = os.path.join(OUT_DIR, "obj_nnet", "nnet_wrap.o")
    sobj = os.path.join(OUT_DIR, "obj_nnet", "nnet_stubs.o")
    iobj = os.path.join(OUT_DIR, "obj_nnet", "ivector_wrap.o")
    subprocess.check_call(["g++"] + flags + ["-c", os.path.join(HERE, "ref_wrap", "ivector_wrap.cc"), "-o", iobj])
    subprocess.check_call(["g++"] + flags + ["-c", wrap, "-o", wobj])
    eobj = os.path.join(OUT_DIR, "obj_nnet", "endpoint_wrap.o")
    subprocess.check_call(["g++"] + flags + ["-c", os.path.join(HERE, "ref_wrap", "endpoint_wrap.cc"), "-o", eobj])
    # OnlineSilenceWeighting: the reference's online-ivector-feature.cc as it lies, over the replay decoder
    zobj = os.path.join(OUT_DIR, "obj_nnet", "silence_wrap.o")
    subprocess.check_call(["g++"] + flags + ["-c", os.path.join(HERE, "ref_wrap", "silence_wrap.cc"), "-o", zobj])
    subprocess.check_call(["g++"] + flags + ["-c", os.path.join(HERE, "ref_wrap", "nnet_stubs.cc"), "-o", sobj])
    subprocess.check_call(["g++", "-shared", "-o", SO] + objs + base_objs + [wobj, sobj, iobj, eobj, zobj, blas,
                          "-Wl,--disable-new-dtags,-rpath," + os.path.dirname(blas), "-lpthread", "-lm", "-ldl"])
    return SO


def compile_only():
    """pre-compile the reference objects (slow part) without the wrapper"""
    stub = os.path.join(OUT_DIR, "inc", "fst")
    os.makedirs(stub, exist_ok=True)
    if not os.path.exists(os.path.join(stub, "fst-decl.h")):
        with open(os.path.join(stub, "fst-decl.h"), "w") as f:
            f.write("#ifndef B2K_FST_DECL_STUB_H_\n#define B2K_FST_DECL_STUB_H_\n"
                    "namespace fst { template <class A> class Fst; template <class W> class ArcTpl; "
                    "template <class T> class TropicalWeightTpl; typedef TropicalWeightTpl<float> TropicalWeight; "
                    "typedef ArcTpl<TropicalWeight> StdArc; template <class A, class S> class VectorFst; }\n#endif\n")
    RF.compile_objects(NNET_SOURCES, os.path.join(OUT_DIR, "obj_nnet"), RF.cxxflags(["-DHAVE_CUDA=0"]))


if __name__ == "__main__":
    import sys
    if "--compile-only" in sys.argv:
        compile_only()
    else:
        print(build(force=True))
