"""Host side of the C++ pipeline (kaldi_b200/csrc/pipeline.cu): b2k_pipeline_plan_for against the sizing rules of the
Python BatchedPipeline / the Python program compiler, argument checking, and the loud failure without a device."""
import ctypes as C
import os

import numpy as np
import pytest

from kaldi_b200 import nnet_model as NM

MDL = os.path.join(os.path.dirname(__file__), "golden", "tiny_final.mdl")


def _model():
    try:
        from kaldi_b200.model import KaldiModel
        return KaldiModel(MDL)
    except OSError as e:
        pytest.skip(str(e))


@pytest.mark.parametrize("S", [160000, 32000, 48017, 400, 1000])
def test_plan_matches_python_sizing(S):
    from kaldi_b200.pipeline import PipelineConfig, native_plan
    m = _model()
    cfg = PipelineConfig(max_batch=7, num_samples=S)
    pl = native_plan(cfg, m)
    T = 1 + (S - 400) // 160                                  # snip-edges NumFrames (feat/feature-window.cc:42)
    arch = NM.arch_tiny(64)
    prog = NM.compile_program(arch, NM.random_weights(arch, seed=11), T, cfg.frames_per_chunk, cfg.acoustic_scale)
    assert (pl["num_feature_frames"], pl["feat_dim"], pl["num_output_frames"], pl["num_chunks"], pl["num_pdfs"],
            pl["ivector_dim"]) == (T, 40, prog["n_out"], prog["n_chunks"], 64, 100)
    assert pl["chunk_samples"] == 2880
    nf = prog["n_out"]
    d = pl["dec"]
    assert (d["max_frames"], d["max_tokens"], d["max_links"], d["max_tokens_per_frame"], d["reference_order"]) == (
        nf + 2, nf * 9000, nf * 16000, 32768, 1)
    assert abs(d["beam"] - cfg.decoder_cfg["beam"]) < 1e-6 and abs(d["lattice_beam"] - cfg.decoder_cfg["lattice_beam"]) < 1e-6
    assert pl["pinned_bytes"] == 4 * 7 * S
    assert pl["device_bytes"] == 4 * 7 * (S + T * 40 + prog["n_chunks"] * 100 + nf * 64)


def test_plan_keeps_explicit_capacities_and_checks_arguments():
    from kaldi_b200.pipeline import PipelineConfig, native_plan
    from kaldi_b200.feat import FeatureOptions
    m = _model()
    pl = native_plan(PipelineConfig(max_batch=2, num_samples=16000, max_tokens=12345, max_links=54321), m)
    assert (pl["dec"]["max_tokens"], pl["dec"]["max_links"]) == (12345, 54321)
    with pytest.raises(RuntimeError):        # too short for one frame
        native_plan(PipelineConfig(max_batch=2, num_samples=399), m)
    with pytest.raises(RuntimeError):        # feature dimension differs from the model's input
        native_plan(PipelineConfig(max_batch=2, num_samples=16000, feature_opts=FeatureOptions(num_ceps=20)), m)
    with pytest.raises(RuntimeError):        # chunk size not a multiple of the subsampling factor
        native_plan(PipelineConfig(max_batch=2, num_samples=16000, frames_per_chunk=20), m)


def test_pipeline_creation_fails_loudly_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from kaldi_b200 import _lib
    from kaldi_b200.pipeline import PipelineConfig, native_cfg
    m = _model()
    L = _lib.lib()
    c = native_cfg(PipelineConfig(max_batch=2, num_samples=16000))
    h = C.c_void_p()
    L.b2k_pipeline_create.argtypes = [C.c_void_p] * 5
    fake_fst = C.c_void_p(8)                 # never dereferenced: the device check comes first
    assert L.b2k_pipeline_create(C.byref(c), m.h, fake_fst, None, C.byref(h)) == 2      # B2K_ERR_NO_DEVICE
    assert not h.value


def test_cmvn_option_is_validated_and_sized():
    """use_cmvn (OnlineNnet2FeaturePipelineInfo::use_cmvn): global stats are mandatory, one more feature buffer is planned."""
    from kaldi_b200 import _lib
    from kaldi_b200.pipeline import PipelineConfig, native_cfg, _native_structs
    m = _model()
    L = _lib.lib()
    _, PP = _native_structs()
    c, pl = native_cfg(PipelineConfig(max_batch=2, num_samples=16000)), PP()
    L.b2k_pipeline_plan_for.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    assert L.b2k_pipeline_plan_for(C.byref(c), m.h, C.byref(pl)) == 0
    base = pl.device_bytes
    assert (c.use_cmvn, c.cmvn.cmn_window, c.cmvn.speaker_frames, c.cmvn.global_frames, c.cmvn.normalize_mean) == (0, 600, 600, 200, 1)
    c.use_cmvn = 1
    assert L.b2k_pipeline_plan_for(C.byref(c), m.h, C.byref(pl)) == 1                    # no global stats
    stats = np.zeros((2, 41), np.float64)
    c.global_cmvn_stats = stats.ctypes.data
    assert L.b2k_pipeline_plan_for(C.byref(c), m.h, C.byref(pl)) == 0
    T = pl.num_feature_frames
    assert pl.device_bytes == base + 4 * 2 * T * 40 + 8 * 3 * 2 * 41
