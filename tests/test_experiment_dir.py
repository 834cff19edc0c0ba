"""tools/check_experiment_dir.py on a directory laid out like steps/online/nnet3/prepare_online_decoding.sh leaves it, built
from files the reference itself wrote (final.mdl with the recipe extras, final.ie, final.dubm) plus option files in the recipes'
own wording: every C++ reader is exercised on one consistent set of files, host only."""
import ctypes as C
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CHAIN_TOPO = """<Topology>
<TopologyEntry>
<ForPhones> 1 2 3 4 5 </ForPhones>
<State> 0 <ForwardPdfClass> 0 <SelfLoopPdfClass> 1 <Transition> 0 0.5 <Transition> 1 0.5 </State>
<State> 1 </State>
</TopologyEntry>
</Topology>
"""


def _build(d):
    from kaldi_b200 import kaldi_io as KIO, nnet_model as NM, synth
    from kaldi_b200.ivector import make_synthetic_extractor
    from oracle import ivector_oracle as IV, nnet_oracle as NO
    if not (os.path.exists(NO._SO) or os.path.isdir("/root/reference")):
        pytest.skip("oracle/_ref libraries not present")
    os.makedirs(os.path.join(d, "conf"))
    os.makedirs(os.path.join(d, "ivector_extractor"))
    plain = NM.arch_tiny(10)                                  # 10 pdfs = 5 phones x 2 pdf classes of the topology below
    arch = dict(plain, recipe_extras=True)
    W = NM.random_weights(plain, seed=1)
    R = NO.RefNnet(arch, W, collapse=False)
    if not hasattr(R.lib, "ref_write_final_mdl"):
        pytest.skip("oracle/_ref library predates the writers")
    R.lib.ref_write_final_mdl.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.c_char_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int]
    pri, t2p = np.ascontiguousarray(W["priors"], np.float32), np.zeros(64, np.int32)
    n_tids = R.lib.ref_write_final_mdl(R.h, os.path.join(d, "final.mdl").encode(), 1, CHAIN_TOPO.encode(), 5, 2, pri.ctypes.data, pri.size,
                                       t2p.ctypes.data, t2p.size)
    assert n_tids > 0
    ex = make_synthetic_extractor(seed=1, num_gauss=64, feat_dim=40, ivector_dim=100, splice=3, base_dim=40)   # sizes the GPU tests use
    RI = IV.RefIvector(ex)
    RI.lib.ref_ivector_write.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p, C.c_int]
    ie = os.path.join(d, "ivector_extractor")
    assert RI.lib.ref_ivector_write(RI.h, os.path.join(ie, "final.ie").encode(), os.path.join(ie, "final.dubm").encode(), 1) == 0
    KIO.write_matrix(os.path.join(ie, "final.mat"), ex["lda_mat"])
    KIO.write_matrix(os.path.join(ie, "global_cmvn.stats"), np.asarray(ex["global_cmvn_stats"], np.float64))
    open(os.path.join(d, "conf", "splice.conf"), "w").write("--left-context=3\n--right-context=3\n")
    open(os.path.join(d, "conf", "online_cmvn.conf"), "w").write("# configuration file for apply-cmvn-online\n")
    open(os.path.join(d, "conf", "mfcc.conf"), "w").write("--use-energy=false\n--num-mel-bins=40\n--num-ceps=40\n--low-freq=20\n--high-freq=-400\n--dither=0\n")
    open(os.path.join(d, "conf", "ivector_extractor.conf"), "w").write(
        f"--splice-config={d}/conf/splice.conf\n--cmvn-config={d}/conf/online_cmvn.conf\n--lda-matrix={ie}/final.mat\n"
        f"--global-cmvn-stats={ie}/global_cmvn.stats\n--diag-ubm={ie}/final.dubm\n--ivector-extractor={ie}/final.ie\n"
        "--num-gselect=5\n--min-post=0.025\n--posterior-scale=0.1\n--max-remembered-frames=1000\n--max-count=100\n")
    open(os.path.join(d, "conf", "online.conf"), "w").write(
        f"--feature-type=mfcc\n--mfcc-config={d}/conf/mfcc.conf\n--ivector-extraction-config={d}/conf/ivector_extractor.conf\n"
        "--endpoint.silence-phones=1:2\n")
    g = synth.make_hclg(3_000, num_pdfs=5, seed=2)            # transition-ids 1..10 = what the transition model has
    hclg = os.path.join(d, "HCLG.fst")
    KIO.write_openfst(hclg, g, "const")
    return hclg, n_tids


def test_consistent_directory_is_accepted_and_sized(tmp_path):
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    try:
        import check_experiment_dir as CE
        d = str(tmp_path / "exp")
        hclg, n_tids = _build(d)
        r = CE.check(d, hclg, seconds=2.0, batch=8)
    except OSError as e:
        pytest.skip(str(e))
    assert r["ok"] and r["feature_type"] == "mfcc" and r["other_options"] == ["--endpoint.silence-phones=1:2"]
    assert r["features"]["dim"] == 40 and r["features"]["dither"] == 0.0
    assert r["model"]["num_pdfs"] == 10 and r["model"]["ivector_dim"] == 100 and r["model"]["transition_ids"] == n_tids
    assert r["ivector_extractor"]["splice"] == [3, 3] and r["ivector_extractor"]["lda"] == [40, 281] and r["ivector_extractor"]["max_count"] == 100.0
    assert r["graph_info"]["type"] == "const" and r["graph_info"]["max_ilabel"] <= n_tids
    assert r["plan"]["feature_frames"] == 1 + (32000 - 400) // 160 and r["plan"]["output_frames"] == (r["plan"]["feature_frames"] + 2) // 3
    assert r["warnings"] == []


def test_inconsistencies_are_reported(tmp_path):
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    try:
        import check_experiment_dir as CE
        d = str(tmp_path / "exp")
        hclg, _ = _build(d)
    except OSError as e:
        pytest.skip(str(e))
    open(os.path.join(d, "conf", "mfcc.conf"), "w").write("--use-energy=false\n--num-mel-bins=40\n--num-ceps=20\n--dither=0\n")
    with pytest.raises(RuntimeError, match="feature dimension"):
        CE.check(d, hclg)
    open(os.path.join(d, "conf", "mfcc.conf"), "w").write("--use-energy=false\n--num-mel-bins=40\n--num-ceps=40\n--low-freq=20\n--high-freq=-400\n")
    r = CE.check(d, hclg)
    assert any("dither" in w for w in r["warnings"])          # the reference's default dither of 1.0
    open(os.path.join(d, "conf", "splice.conf"), "w").write("--left-context=2\n--right-context=2\n")
    with pytest.raises(RuntimeError, match="final.mat"):
        CE.check(d, hclg)


def test_c99_program_drives_the_directory_through_the_abi(tmp_path):
    """tests/cabi/experiment_route.c: option files, model, graph, waveform (resampled from 8 kHz), extractor — all through the C
    ABI; without a GPU it stops at the first device call with the library's message."""
    import shutil
    import struct
    import subprocess
    so = os.path.join(ROOT, "kaldi_b200", "libb2k.so")
    if not os.path.exists(so) or not shutil.which("gcc"):
        pytest.skip("libb2k.so or gcc missing")
    d = str(tmp_path / "exp")
    hclg, n_tids = _build(d)
    x = (3000 * np.sin(2 * np.pi * 300 * np.arange(8000) / 8000)).astype("<i2").tobytes()          # one second at 8 kHz
    body = b"WAVE" + b"fmt " + struct.pack("<IHHIIHH", 16, 1, 1, 8000, 16000, 2, 16) + b"data" + struct.pack("<I", len(x)) + x
    wav = str(tmp_path / "utt.wav")
    open(wav, "wb").write(b"RIFF" + struct.pack("<I", len(body)) + body)
    exe = str(tmp_path / "experiment_route")
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-I" + os.path.join(ROOT, "include"),
                        os.path.join(ROOT, "tests", "cabi", "experiment_route.c"), "-o", exe, "-L" + os.path.dirname(so), "-lb2k",
                        "-Wl,-rpath," + os.path.dirname(so)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([exe, os.path.join(d, "conf", "online.conf"), os.path.join(d, "final.mdl"), hclg, wav, str(tmp_path / "out.ark")],
                       capture_output=True, text=True, timeout=300)
    frames = 1 + (16000 - 400) // 160                        # 8000 samples at 8 kHz -> 16000 at 16 kHz
    assert f"host side ready: {frames} feature frames -> {(frames + 2) // 3} decoder frames, 10 pdfs, {n_tids} transition-ids" in r.stdout, r.stdout + r.stderr
    assert "endpointing: silence phones 1:2, rule2 fires after 0.6 s of trailing silence: yes" in r.stdout, r.stdout
    import torch
    if torch.cuda.is_available():
        assert r.returncode == 0, r.stdout + r.stderr
    else:
        assert r.returncode == 3 and "no CUDA device" in r.stderr


def test_c99_streaming_program_reaches_the_device_boundary(tmp_path):
    """tests/cabi/stream_route.c (one utterance chunk by chunk through b2k_stream_*, partial hypotheses and the end-point rules per
    chunk): compiles as strict C99 against include/b2k.h and runs its host side; without a GPU it stops at the first device call
    with the library's message."""
    import shutil
    import struct
    import subprocess
    so = os.path.join(ROOT, "kaldi_b200", "libb2k.so")
    if not os.path.exists(so) or not shutil.which("gcc"):
        pytest.skip("libb2k.so or gcc missing")
    d = str(tmp_path / "exp")
    hclg, n_tids = _build(d)
    x = (3000 * np.sin(2 * np.pi * 300 * np.arange(8000) / 8000)).astype("<i2").tobytes()
    body = b"WAVE" + b"fmt " + struct.pack("<IHHIIHH", 16, 1, 1, 8000, 16000, 2, 16) + b"data" + struct.pack("<I", len(x)) + x
    wav = str(tmp_path / "utt.wav")
    open(wav, "wb").write(b"RIFF" + struct.pack("<I", len(body)) + body)
    exe = str(tmp_path / "stream_route")
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I" + os.path.join(ROOT, "include"),
                        os.path.join(ROOT, "tests", "cabi", "stream_route.c"), "-o", exe, "-L" + os.path.dirname(so), "-lb2k",
                        "-Wl,-rpath," + os.path.dirname(so)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([exe, os.path.join(d, "conf", "online.conf"), os.path.join(d, "final.mdl"), hclg, wav, str(tmp_path / "out.ark")],
                       capture_output=True, text=True, timeout=300)
    assert f"host side ready: 16000 samples in chunks of 8160, 10 pdfs, {n_tids} transition-ids, silence phones 1:2" in r.stdout, r.stdout + r.stderr
    import torch
    if torch.cuda.is_available():
        assert r.returncode == 0, r.stdout + r.stderr
    else:
        assert r.returncode == 3 and "no CUDA device" in r.stderr
