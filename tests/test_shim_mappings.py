"""The option mappings of the Kaldi-typed shims, run (not only compiled): tests/cabi/option_mapping_kaldi_test.cc is linked against
oracle/_ref (the reference's own ParseOptions and option structs) and libb2k.so; a config file read by the reference's
MfccOptions / FbankOptions / PlpOptions / OnlineEndpointConfig and mapped by the shim (ToB2kFeatCfg, ToB2kEndpointConfig) must give
the struct that b2k's own readers (b2k_feat_cfg_from_conf, b2k_endpoint_cfg_from_conf) make of the same file.  Host code on both
sides: no device.  Needs the reference's headers: this container only."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CONFS = {
    "mfcc": ["",
             "--use-energy=false\n--num-mel-bins=40\n--num-ceps=40\n--low-freq=20\n--high-freq=-400\n",        # mfcc_hires.conf
             "--sample-frequency=8000\n--frame-length=20\n--frame-shift=5\n--window-type=hamming\n--preemphasis-coefficient=0.9\n"
             "--snip-edges=false\n--cepstral-lifter=0\n--htk-compat=true\n--raw-energy=false\n--energy-floor=1.5\n--dither=0\n"],
    "fbank": ["",
              "--num-mel-bins=80\n--use-log-fbank=false\n--use-power=false\n--use-energy=true\n--htk-compat=true\n"
              "--round-to-power-of-two=true\n--remove-dc-offset=false\n--window-type=hanning\n"],
    "plp": ["",
            "--lpc-order=10\n--num-ceps=11\n--compress-factor=0.25\n--cepstral-scale=2.0\n--cepstral-lifter=11\n--num-mel-bins=20\n"
            "--use-energy=false\n--window-type=rectangular\n"],
    "endpoint": ["",
                 "--endpoint.silence-phones=1:2:3\n--endpoint.rule2.min-trailing-silence=0.3\n--endpoint.rule5.min-utterance-length=9\n"
                 "--endpoint.rule1.must-contain-nonsilence=true\n"],
}


def test_option_mappings_of_the_shims_equal_b2ks_own_readers(tmp_path):
    if not os.path.isdir("/root/reference/src"):
        pytest.skip("/root/reference not present")
    so_b2k = os.path.join(ROOT, "kaldi_b200", "libb2k.so")
    if not os.path.exists(so_b2k):
        pytest.skip("libb2k.so not built")
    from oracle import nnet_oracle as NO, ref_feat as RF, ref_nnet
    ref_nnet.build(quiet=True)
    blas = os.path.dirname(RF.find_openblas())
    flags = RF.cxxflags(["-DHAVE_CUDA=0", "-I" + os.path.join(ROOT, "include"), "-I" + os.path.join(ROOT, "kaldi_b200", "host"),
                         "-I/usr/local/cuda/include", "-I" + os.path.join(ROOT, "oracle", "_ref", "inc"),
                         "-I" + os.path.join(ROOT, "oracle", "ref_wrap")])
    exe = str(tmp_path / "omk")
    r = subprocess.run(["g++"] + flags + [os.path.join(ROOT, "tests", "cabi", "option_mapping_kaldi_test.cc"), "-o", exe, NO._SO,
                        "-Wl,-rpath," + os.path.dirname(NO._SO), "-Wl,-rpath-link," + blas, "-Wl,-rpath," + blas,
                        "-L" + os.path.dirname(so_b2k), "-lb2k", "-Wl,-rpath," + os.path.dirname(so_b2k), "-Wl,--allow-shlib-undefined"],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    args = []
    for kind, texts in CONFS.items():
        for i, text in enumerate(texts):
            p = tmp_path / f"{kind}{i}.conf"
            p.write_text(text)
            args += [kind, str(p)]
    r = subprocess.run([exe] + args, capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "option mappings ok (9 files)" in r.stdout, r.stdout + r.stderr[-2000:]
