"""include/b2k.h is a C header: a C99 translation unit (tests/cabi/host_route.c) compiles against it with
-Wall -Wextra -pedantic, links against libb2k.so and drives the device-free part of the model route."""
import os
import re
import shutil
import subprocess

import pytest

from kaldi_b200 import nnet_model as NM

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_c99_caller_compiles_links_and_runs(tmp_path):
    so = os.path.join(ROOT, "kaldi_b200", "libb2k.so")
    if not os.path.exists(so) or not shutil.which("gcc"):
        pytest.skip("libb2k.so or gcc missing")
    exe = str(tmp_path / "host_route")
    cmd = ["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I" + os.path.join(ROOT, "include"),
           os.path.join(ROOT, "tests", "cabi", "host_route.c"), "-o", exe, "-L" + os.path.dirname(so), "-lb2k",
           "-Wl,-rpath," + os.path.dirname(so)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([exe, os.path.join(ROOT, "tests", "golden", "tiny_final.mdl")], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    kv = dict(re.findall(r"(\w+)=(-?\d+)", r.stdout))
    T = 1 + (32000 - 400) // 160
    arch = NM.arch_tiny(64)
    prog = NM.compile_program(arch, NM.random_weights(arch, seed=11), T, 21)
    assert int(kv["pdfs"]) == 64 and int(kv["frames"]) == T
    assert (int(kv["out"]), int(kv["chunks"])) == (prog["n_out"], prog["n_chunks"]) == (int(kv["prog_out"]), int(kv["prog_chunks"]))
    assert (int(kv["nodes"]), int(kv["ops"]), int(kv["blob"])) == (len(prog["nodes"]), len(prog["ops"]), prog["blob"].size)
    assert int(kv["tids"]) == 11
    import torch
    assert int(kv["device_rc"]) == (0 if torch.cuda.is_available() else 2)      # B2K_ERR_NO_DEVICE: no CPU path
