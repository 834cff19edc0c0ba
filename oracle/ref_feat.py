"""oracle/_ref feature library: the reference's OWN feature sources compiled
from where they lie under /root/reference/src (base + matrix + util subset +
feat + transform/cmvn), plus our thin wrapper oracle/ref_wrap/feat_wrap.cc.
TEST INFRASTRUCTURE ONLY.  Recipe follows SURVEY.md §8c: g++ directly on the
files, -DOPENFST_VER=10804 -DHAVE_CLAPACK, the vendored tools/CLAPACK headers,
linked against the OpenBLAS bundled with opencv-python-headless in this image
(present on the GPU box too: same image).  The reference's build system is not
run; nothing is copied into the repo; outputs go to oracle/_ref/ (git-ignored).
"""
from __future__ import annotations

import glob
import os
import subprocess
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"
SRC = os.path.join(REF, "src")
OUT_DIR = os.path.join(HERE, "_ref")
SO = os.path.join(OUT_DIR, "libkaldi_ref_feat.so")


def find_openblas():
    import sysconfig
    sp = sysconfig.get_paths()["purelib"]
    c = glob.glob(os.path.join(sp, "opencv_python_headless.libs", "libopenblasp-r0-*.so"))
    if not c:
        raise RuntimeError("bundled OpenBLAS (opencv_python_headless.libs) not found")
    return c[0]


FEAT_SOURCES = (
    ["base/" + f for f in ("kaldi-error.cc", "kaldi-math.cc", "kaldi-utils.cc", "io-funcs.cc", "timer.cc")]
    + ["matrix/" + f for f in ("compressed-matrix.cc", "kaldi-matrix.cc", "kaldi-vector.cc",
                               "matrix-functions.cc", "optimization.cc", "packed-matrix.cc", "qr.cc",
                               "sp-matrix.cc", "sparse-matrix.cc", "srfft.cc", "tp-matrix.cc")]
    + ["util/" + f for f in ("kaldi-holder.cc", "kaldi-io.cc", "kaldi-semaphore.cc", "kaldi-table.cc",
                             "kaldi-thread.cc", "parse-options.cc", "simple-io-funcs.cc",
                             "simple-options.cc", "text-utils.cc")]
    + ["feat/" + f for f in ("feature-window.cc", "feature-mfcc.cc", "feature-fbank.cc", "feature-plp.cc",
                             "feature-spectrogram.cc", "mel-computations.cc", "feature-functions.cc",
                             "online-feature.cc", "resample.cc", "wave-reader.cc")]
    + ["transform/cmvn.cc"]
)


def cxxflags(inc_extra=()):
    ver = os.path.join(OUT_DIR, "inc")
    os.makedirs(os.path.join(ver, "base"), exist_ok=True)
    vh = os.path.join(ver, "base", "version.h")
    if not os.path.exists(vh):
        with open(vh, "w") as f:
            f.write('#define KALDI_VERSION "5.5-oracle"\n')
    return ["-std=c++17", "-O2", "-fPIC", "-w", "-DOPENFST_VER=10804", "-DHAVE_CLAPACK",
            "-DKALDI_DOUBLEPRECISION=0", "-DHAVE_EXECINFO_H=1", "-I" + SRC,
            "-I" + os.path.join(REF, "tools", "CLAPACK"), "-I" + ver] + list(inc_extra)


def compile_objects(sources, objdir, flags, quiet=False):
    os.makedirs(objdir, exist_ok=True)
    jobs = []
    for rel in sources:
        src = rel if os.path.isabs(rel) else os.path.join(SRC, rel)
        obj = os.path.join(objdir, rel.replace("/", "_").replace(".cc", ".o"))
        if os.path.exists(obj) and os.path.getmtime(obj) >= os.path.getmtime(src):
            jobs.append((None, obj))
            continue
        jobs.append((["g++"] + flags + ["-c", src, "-o", obj], obj))

    def run(j):
        if j[0] is not None:
            subprocess.check_call(j[0])
        return j[1]
    with ThreadPoolExecutor(max_workers=os.cpu_count() or 4) as ex:
        return list(ex.map(run, jobs))


def build(quiet: bool = False, force: bool = False) -> str:
    wrap = os.path.join(HERE, "ref_wrap", "feat_wrap.cc")
    if (not force and os.path.exists(SO) and os.path.getmtime(SO) >= os.path.getmtime(wrap)):
        return SO
    if not os.path.isdir(SRC):
        raise RuntimeError("/root/reference not present: cannot (re)build oracle/_ref")
    blas = find_openblas()
    flags = cxxflags()
    objs = compile_objects(FEAT_SOURCES, os.path.join(OUT_DIR, "obj_feat"), flags, quiet)
    wobj = os.path.join(OUT_DIR, "obj_feat", "feat_wrap.o")
    subprocess.check_call(["g++"] + flags + ["-c", wrap, "-o", wobj])
    subprocess.check_call(["g++", "-shared", "-o", SO] + objs + [wobj, blas,
                          "-Wl,--disable-new-dtags,-rpath," + os.path.dirname(blas), "-lpthread", "-lm", "-ldl"])
    return SO


if __name__ == "__main__":
    print(build(force=True))
