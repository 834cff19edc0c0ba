"""Every ABI entry point called with null handles / null outputs must return an error code (or a sentinel for the
value-returning accessors) instead of crashing, without a device and without any object having been created.  Run in a
child process so that a crash is a test failure, not a dead test session."""
import os
import subprocess
import sys
import textwrap

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CHILD = textwrap.dedent(r'''
    import ctypes as C, re, sys
    sys.path.insert(0, %r)
    from kaldi_b200 import _lib
    L = _lib.lib()
    hdr = open(%r).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    found = re.findall(r"\n\s*(int|int32_t|int64_t|float|double|void|const [a-z0-9_]+(?: [a-z0-9_]+)?|b2k_dec)\s*(\*?)\s*(b2k_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", hdr)
    protos = [(r + (" *" if star else ""), n, a) for r, star, n, a in found]
    checked = 0
    for ret, name, args in protos:
        fn = getattr(L, name)
        params = [a.strip() for a in args.replace("\n", " ").split(",")] if args.strip() not in ("", "void") else []
        argv, types = [], []
        for a in params:
            if "*" in a:
                argv.append(None); types.append(C.c_void_p)
            elif a.startswith(("float", "double")):
                argv.append(1.0); types.append(C.c_float if a.startswith("float") else C.c_double)
            elif "int64_t" in a:
                argv.append(1); types.append(C.c_int64)
            else:
                argv.append(1); types.append(C.c_int32)
        fn.argtypes = types
        if ret in ("int", "int32_t"):
            fn.restype = C.c_int32
        elif ret == "int64_t":
            fn.restype = C.c_int64
        elif ret in ("float", "double"):
            fn.restype = C.c_float if ret == "float" else C.c_double
        elif ret == "void":
            fn.restype = None
        else:
            fn.restype = C.c_void_p
        r = fn(*argv)
        checked += 1
        if ret == "int" and name not in ("b2k_version",) and not name.endswith("_destroy"):
            assert r != 0, name + " accepted null arguments"
        if ret not in ("int", "int32_t", "int64_t", "float", "double", "void") and name != "b2k_last_error":
            assert not r, name + " returned a pointer for a null handle"
    print("checked", checked)
''')


def test_null_arguments_never_crash():
    so = os.path.join(ROOT, "kaldi_b200", "libb2k.so")
    if not os.path.exists(so):
        import pytest
        pytest.skip("libb2k.so not built")
    r = subprocess.run([sys.executable, "-c", CHILD % (ROOT, os.path.join(ROOT, "include", "b2k.h"))], capture_output=True, text=True,
                       timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    n = int(r.stdout.strip().split()[-1])
    exported = subprocess.run("nm -D %s | grep -c ' T b2k_'" % so, shell=True, capture_output=True, text=True).stdout.strip()
    assert n == int(exported), (n, exported)                 # every exported entry point was exercised
