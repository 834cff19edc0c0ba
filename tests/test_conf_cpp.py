"""Kaldi option files (conf/mfcc_hires.conf, conf/ivector_extractor.conf, splice.conf, online_cmvn.conf) through the C++
readers of libb2k.so (kaldi_b200/csrc/host_utils.cu: b2k_feat_cfg_from_conf, b2k_ivec_cfg_from_conf), against the
reference's own ParseOptions::ReadConfigFile + MfccOptions / FbankOptions compiled in oracle/_ref, and — for the i-vector
options, whose header needs OpenFst — against the option names the reference's header registers."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from kaldi_b200.feat import _FeatCfg

HIRES = """# config for high-resolution MFCC features, intended for neural network training
--use-energy=false   # use average of log energy, not energy.
--num-mel-bins=40     # similar to Google's setup.
--num-ceps=40     # there is no dimensionality reduction.
--low-freq=20     # low cutoff frequency for mel bins... this is high-bandwidth data, so
                  # there might be some information at the low end.
--high-freq=-400 # high cutoff frequently, relative to Nyquist of 8000 (=7600)
--dither=0
"""
ODD = """
  --Sample_Frequency=8000
--frame-length=20 
--snip-edges=F
--remove_dc_offset
--window-type=hamming
--raw-energy=0   # trailing comment
--energy-floor=1.5
"""
# spellings of numbers the reference's ConvertStringToReal / ConvertStringToInteger take (found by a soak run against its ParseOptions)
NUMBERS = "--energy-floor=INF\n--low-freq=+20\n--high-freq=-.5e3\n--dither=5.\n--num-mel-bins=+30\n--cepstral-lifter=1E1\n--preemphasis-coefficient=-infinity\n"
FBANK = "--num-mel-bins=64\n--use-log-fbank=false\n--use-power=t\n--use-energy=true\n--dither=0.0\n--htk-compat=TRUE\n"


def _lib():
    try:
        from kaldi_b200 import _lib as L
        return L.lib()
    except OSError as e:
        pytest.skip(str(e))


def _mine(L, path, feature_type):
    c = _FeatCfg()
    L.b2k_feat_cfg_from_conf.argtypes = [C.c_char_p, C.c_int32, C.c_void_p]
    rc = L.b2k_feat_cfg_from_conf(path.encode(), feature_type, C.byref(c))
    return rc, c


def _reference(path, feature_type):
    from oracle import feat_oracle as F
    R = F.RefFeat()
    if not hasattr(R.lib, "ref_feat_opts_from_conf"):
        pytest.skip("oracle/_ref feature library predates ref_feat_opts_from_conf")
    out = (C.c_float * 24)()
    win = C.create_string_buffer(32)
    R.lib.ref_feat_opts_from_conf.argtypes = [C.c_char_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int]
    rc = R.lib.ref_feat_opts_from_conf(path.encode(), feature_type, out, win, 32)
    return rc, list(out), win.value.decode()


FIELDS = ["samp_freq", "frame_shift_ms", "frame_length_ms", "dither", "preemph_coeff", "remove_dc_offset", "round_to_power_of_two",
          "snip_edges", "num_bins", "low_freq", "high_freq", "num_ceps", "use_energy", "energy_floor", "raw_energy",
          "cepstral_lifter", "htk_compat", "use_log_fbank", "use_power"]
WINDOWS = {"povey": 0, "hamming": 1, "hanning": 2, "rectangular": 3}


@pytest.mark.parametrize("text,feature_type", [(HIRES, 0), (ODD, 0), ("", 0), (FBANK, 1), (ODD, 1), ("", 1), (NUMBERS, 0)],
                         ids=["hires", "odd-spelling", "defaults-mfcc", "fbank", "odd-fbank", "defaults-fbank", "number-spellings"])
def test_feature_options_equal_the_references_parse(tmp_path, text, feature_type):
    L = _lib()
    p = str(tmp_path / "feat.conf")
    open(p, "w").write(text)
    rc, c = _mine(L, p, feature_type)
    rrc, ref, win = _reference(p, feature_type)
    assert rc == 0 and rrc == 0
    for i, k in enumerate(FIELDS):
        if feature_type == 1 and k in ("num_ceps", "cepstral_lifter"):
            continue                                          # FbankOptions has no such option
        if feature_type == 0 and k in ("use_log_fbank", "use_power"):
            continue
        assert float(getattr(c, k)) == ref[i] or float(getattr(c, k)) == pytest.approx(ref[i], abs=1e-6), k
    assert c.window_type == WINDOWS[win] and c.feature_type == feature_type


PLP = "--lpc-order=10\n--num-ceps=9\n--compress-factor=0.25\n--cepstral-lifter=20\n--cepstral-scale=2.5\n--use-energy=false\n--num-mel-bins=21\n--dither=0\n"


@pytest.mark.parametrize("text", [PLP, "", ODD], ids=["plp", "defaults-plp", "odd-plp"])
def test_plp_options_equal_the_references_parse(tmp_path, text):
    """PlpOptions::Register (feat/feature-plp.h:68-90): the option file of --plp-config through b2k_feat_cfg_from_conf(…, 2, …)
    and through the reference's own ParseOptions."""
    L = _lib()
    p = str(tmp_path / "plp.conf")
    open(p, "w").write(text)
    rc, c = _mine(L, p, 2)
    rrc, ref, win = _reference(p, 2)
    assert rc == 0 and rrc == 0
    for i, k in enumerate(FIELDS + ["lpc_order", "compress_factor", "cepstral_scale"]):
        if k in ("use_log_fbank", "use_power"):
            continue
        assert float(getattr(c, k)) == pytest.approx(ref[i], abs=1e-6), k
    assert c.window_type == WINDOWS[win] and c.feature_type == 2
    for bad in ("--cepstral-lifter=22.5\n", "--use-power=true\n", "--lpc-order=twelve\n"):      # an int32 option; an fbank option; not a number
        open(p, "w").write(bad)
        assert _mine(L, p, 2)[0] != 0 and _reference(p, 2)[0] != 0


@pytest.mark.parametrize("text", ["--no-such-option=3\n", "num-ceps=13\n", "--frame-length = 20\n", "--num-ceps=thirteen\n", "--use-energy=maybe\n",
                                  "--use-log-fbank=true\n", "--htk-compat=\n", "--dither=0x10\n", "--dither=1e400\n", "--num-ceps=2147483648\n",
                                  "--num-ceps=3.0\n", "--dither=1.0f\n", "--frame-shift=\n"])
def test_both_reject_the_same_files(tmp_path, text):
    L = _lib()
    p = str(tmp_path / "bad.conf")
    open(p, "w").write(text)
    rc, _ = _mine(L, p, 0)
    rrc, _, _ = _reference(p, 0)
    assert rc != 0 and rrc != 0


def test_missing_file_is_an_error(tmp_path):
    L = _lib()
    assert _mine(L, str(tmp_path / "nope.conf"), 0)[0] != 0


class _IvecCfg(C.Structure):
    _fields_ = [("base_dim", C.c_int32), ("splice_left", C.c_int32), ("splice_right", C.c_int32), ("feat_dim", C.c_int32),
                ("num_gauss", C.c_int32), ("ivector_dim", C.c_int32), ("num_gselect", C.c_int32), ("min_post", C.c_float),
                ("posterior_scale", C.c_float), ("max_count", C.c_float), ("prior_offset", C.c_float), ("num_cg_iters", C.c_int32),
                ("cmn_window", C.c_int32), ("speaker_frames", C.c_int32), ("global_frames", C.c_int32), ("max_lanes", C.c_int32),
                ("max_frames", C.c_int32)]


class _IvecPaths(C.Structure):
    _fields_ = [(k, C.c_char * 512) for k in ("lda_matrix", "global_cmvn_stats", "splice_config", "cmvn_config", "diag_ubm",
                                               "ivector_extractor")] + \
               [("ivector_period", C.c_int32), ("use_most_recent_ivector", C.c_int32), ("greedy_ivector_extractor", C.c_int32),
                ("online_cmvn_iextractor", C.c_int32), ("max_remembered_frames", C.c_float)]


def _ivec(L, path, base_dim=40):
    c, p = _IvecCfg(), _IvecPaths()
    c.base_dim, c.max_lanes, c.max_frames = base_dim, 8, 1000
    L.b2k_ivec_cfg_from_conf.argtypes = [C.c_char_p, C.c_void_p, C.c_void_p]
    return L.b2k_ivec_cfg_from_conf(path.encode(), C.byref(c), C.byref(p)), c, p


def test_ivector_extraction_config_as_the_recipes_write_it(tmp_path):
    """steps/online/nnet3/prepare_online_decoding.sh writes exactly these lines into conf/ivector_extractor.conf."""
    L = _lib()
    d = str(tmp_path)
    open(os.path.join(d, "splice.conf"), "w").write("--left-context=3\n--right-context=3\n")
    open(os.path.join(d, "online_cmvn.conf"), "w").write("# configuration file for apply-cmvn-online, used in the script ../local/run_online_decoding.sh\n")
    conf = os.path.join(d, "ivector_extractor.conf")
    open(conf, "w").write(f"--splice-config={d}/splice.conf\n--cmvn-config={d}/online_cmvn.conf\n--lda-matrix={d}/final.mat\n"
                          f"--global-cmvn-stats={d}/global_cmvn.stats\n--diag-ubm={d}/final.dubm\n--ivector-extractor={d}/final.ie\n"
                          "--num-gselect=5\n--min-post=0.025\n--posterior-scale=0.1\n--max-remembered-frames=1000\n--max-count=100\n")
    rc, c, p = _ivec(L, conf)
    assert rc == 0
    assert (c.splice_left, c.splice_right, c.num_gselect, c.num_cg_iters) == (3, 3, 5, 15)
    assert (c.cmn_window, c.speaker_frames, c.global_frames) == (600, 600, 200)
    assert c.min_post == pytest.approx(0.025) and c.posterior_scale == pytest.approx(0.1) and c.max_count == pytest.approx(100.0)
    assert (c.base_dim, c.max_lanes, c.max_frames) == (40, 8, 1000)
    assert p.lda_matrix.decode() == f"{d}/final.mat" and p.ivector_extractor.decode() == f"{d}/final.ie"
    assert p.diag_ubm.decode() == f"{d}/final.dubm" and p.global_cmvn_stats.decode() == f"{d}/global_cmvn.stats"
    assert (p.ivector_period, p.use_most_recent_ivector, p.greedy_ivector_extractor) == (10, 1, 0)
    assert p.max_remembered_frames == pytest.approx(1000.0)
    # the reference's defaults when the file says nothing (OnlineIvectorExtractionConfig(), OnlineSpliceOptions(), OnlineCmvnOptions())
    empty = os.path.join(d, "empty.conf")
    open(empty, "w").write("\n")
    rc, c, p = _ivec(L, empty)
    assert rc == 0 and (c.splice_left, c.splice_right, c.num_gselect) == (4, 4, 5) and c.max_count == 0.0
    # rejected: unknown names, unsupported CMVN variants
    open(conf, "a").write("--no-such=1\n")
    assert _ivec(L, conf)[0] != 0
    open(os.path.join(d, "online_cmvn.conf"), "w").write("--norm-vars=true\n")
    open(conf, "w").write(f"--cmvn-config={d}/online_cmvn.conf\n")
    assert _ivec(L, conf)[0] != 0


def test_ivector_option_names_are_the_ones_the_reference_registers():
    """online2/online-ivector-feature.h needs OpenFst to compile, so the name set is checked against its text."""
    src = "/root/reference/src"
    if not os.path.isdir(src):
        pytest.skip("the reference tree exists in the build container only")
    L = _lib()
    import tempfile

    def names(path, start, end):
        text = open(os.path.join(src, path)).read()
        text = text[text.index(start):]
        text = text[:text.index(end)]
        return set(re.findall(r'Register\("([a-z\-]+)"', text))
    groups = {"top": names("online2/online-ivector-feature.h", "struct OnlineIvectorExtractionConfig", "struct OnlineIvectorExtractionInfo"),
              "splice": names("feat/online-feature.h", "struct OnlineSpliceOptions", "class OnlineSpliceFrames"),
              "cmvn": names("feat/online-feature.h", "struct OnlineCmvnOptions", "struct OnlineCmvnState")}
    assert "num-gselect" in groups["top"] and "left-context" in groups["splice"] and "cmn-window" in groups["cmvn"]
    with tempfile.TemporaryDirectory() as d:
        sub = {"splice": os.path.join(d, "s.conf"), "cmvn": os.path.join(d, "c.conf")}
        for g, ns in groups.items():
            for n in sorted(ns):
                for f in sub.values():
                    open(f, "w").write("\n")
                val = {"top": "1", "splice": "2", "cmvn": "5"}[g]
                if n in ("norm-vars", "online-cmvn-iextractor", "greedy-ivector-extractor"):
                    val = "false"                             # accepted by name; true has no kernel behind it and is an error
                if n == "norm-means":
                    val = "true"
                if n == "skip-dims":
                    continue                                  # accepted by name, rejected by value unless empty (unsupported)
                if g == "top":
                    if n in ("cmvn-config", "splice-config"):
                        val = sub["cmvn" if n == "cmvn-config" else "splice"]      # these two files are opened
                    top = f"--{n}={val}\n"
                else:
                    open(sub[g], "w").write(f"--{n}={val}\n")
                    top = f"--{'splice-config' if g == 'splice' else 'cmvn-config'}={sub[g]}\n"
                conf = os.path.join(d, "top.conf")
                open(conf, "w").write(top)
                assert _ivec(L, conf)[0] == 0, (g, n)
        # values that would change what the reference computes are refused, not ignored
        for bad in ("--use-most-recent-ivector=false", "--greedy-ivector-extractor=true", "--num-gselect"):
            open(conf, "w").write(bad + "\n")
            assert _ivec(L, conf)[0] != 0, bad


class _OnlineConf(C.Structure):
    _fields_ = [("feature_type", C.c_int32), ("add_pitch", C.c_int32)] + \
               [(k, C.c_char * 512) for k in ("mfcc_config", "fbank_config", "cmvn_config", "global_cmvn_stats", "ivector_extraction_config")] + \
               [("rest", C.c_char * 4096), ("plp_config", C.c_char * 512)]


def test_online_conf_as_prepare_online_decoding_writes_it(tmp_path):
    L = _lib()
    p = str(tmp_path / "online.conf")
    open(p, "w").write("--feature-type=mfcc\n--mfcc-config=/exp/conf/mfcc.conf\n--ivector-extraction-config=/exp/conf/ivector_extractor.conf\n"
                       "--endpoint.silence-phones=1:2:3:4:5\n--endpoint.rule1.min-trailing-silence=0.5\n")
    c = _OnlineConf()
    L.b2k_online_conf_read.argtypes = [C.c_char_p, C.c_void_p]
    assert L.b2k_online_conf_read(p.encode(), C.byref(c)) == 0
    assert c.feature_type == 0 and c.mfcc_config.decode() == "/exp/conf/mfcc.conf"
    assert c.ivector_extraction_config.decode() == "/exp/conf/ivector_extractor.conf"
    assert sorted(c.rest.decode().split()) == ["--endpoint.rule1.min-trailing-silence=0.5", "--endpoint.silence-phones=1:2:3:4:5"]
    open(p, "w").write("--feature-type=plp\n--plp-config=/exp/conf/plp.conf\n")
    assert L.b2k_online_conf_read(p.encode(), C.byref(c)) == 0 and c.feature_type == 2 and c.plp_config.decode() == "/exp/conf/plp.conf"
    open(p, "w").write("--feature-type=spectrogram\n")
    assert L.b2k_online_conf_read(p.encode(), C.byref(c)) != 0
    open(p, "w").write("--feature-type=fbank\n--add-pitch=true\n")
    assert L.b2k_online_conf_read(p.encode(), C.byref(c)) != 0
    # the feature group's names are the ones the reference registers
    hdr = "/root/reference/src/online2/online-nnet2-feature-pipeline.h"
    if os.path.exists(hdr):
        text = open(hdr).read()
        text = text[text.index("struct OnlineNnet2FeaturePipelineConfig"):text.index("struct OnlineNnet2FeaturePipelineInfo")]
        names = set(re.findall(r'Register\("([a-z\-]+)"', text))
        assert {"feature-type", "mfcc-config", "ivector-extraction-config", "add-pitch"} <= names
        for n in sorted(names):
            val = "false" if n == "add-pitch" else ("mfcc" if n == "feature-type" else "x")
            open(p, "w").write(f"--{n}={val}\n")
            assert L.b2k_online_conf_read(p.encode(), C.byref(c)) == 0, n
            assert c.rest.decode() == "", n                    # recognised, not passed through


def test_decoder_and_decodable_options_in_the_tools_spelling():
    from kaldi_b200.pipeline import PipelineConfig, native_cfg
    L = _lib()
    L.b2k_pipeline_cfg_apply_options.argtypes = [C.c_char_p, C.c_void_p]
    c = native_cfg(PipelineConfig(max_batch=2, num_samples=16000))
    text = b"--beam=13.5 --Max_Active=5000\n--min-active=100 --lattice-beam=6 --beam-delta=0.25 --hash-ratio=3\n--acoustic-scale=0.9 " \
           b"--frames-per-chunk=50 --chunk-length=0.3 --endpoint.silence-phones=1:2:3 --do-endpointing=false --det.max-mem=1000 " \
           b"--ivector-silence-weighting.silence-weight=0.5 --word-symbol-table=/x/words.txt --online=true --prune-interval=10\n"
    assert L.b2k_pipeline_cfg_apply_options(text, C.byref(c)) == 0
    assert (c.dec.beam, c.dec.max_active, c.dec.min_active, c.dec.lattice_beam, c.dec.prune_interval) == (13.5, 5000, 100, 6.0, 10)
    assert c.dec.beam_delta == 0.25 and c.dec.hash_ratio == 3.0 and c.acoustic_scale == pytest.approx(0.9)
    assert c.frames_per_chunk == 51 and c.chunk_length_secs == pytest.approx(0.3)       # 50 rounded up to a multiple of 3
    before = bytes(c)
    # (silence weighting without silence phones is inactive, OnlineSilenceWeightingConfig::Active(): accepted above; active = refused)
    for bad in (b"--no-such-option=1", b"beam=3", b"--beam=wide", b"--extra-left-context-initial=5", b"--frames-per-chunk=0",
                b"--ivector-silence-weighting.silence-weight=0.5 --ivector-silence-weighting.silence-phones=1:2:3",
                b"--do-endpointing=true", b"--word-symbol-table", b"--beam"):
        assert L.b2k_pipeline_cfg_apply_options(bad, C.byref(c)) != 0
        assert bytes(c) == before                             # a rejected text changes nothing
    # the names are the ones the reference registers
    src = "/root/reference/src"
    if os.path.isdir(src):
        def names(path, start, end):
            t = open(os.path.join(src, path)).read()
            t = t[t.index(start):]
            return set(re.findall(r'Register\("([a-z\-]+)"', t[:t.index(end)]))
        dec = names("decoder/lattice-faster-decoder.h", "struct LatticeFasterDecoderConfig", "void Check() const")
        nn = names("nnet3/decodable-simple-looped.h", "struct NnetSimpleLoopedComputationOptions", "class DecodableNnetSimpleLoopedInfo")
        assert {"beam", "lattice-beam", "hash-ratio"} <= dec and {"acoustic-scale", "frames-per-chunk"} <= nn
        for n in sorted(dec | nn):
            val = {"determinize-lattice": "true", "debug-computation": "false", "extra-left-context-initial": "0", "frame-subsampling-factor": "3"}.get(n, "7")
            assert L.b2k_pipeline_cfg_apply_options(f"--{n}={val}".encode(), C.byref(c)) == 0, n
