"""Host-side C++ above the C ABI (kaldi_b200/host/): the utterance batcher behind the reference's chunked
DecodeBatch(corr_ids, wave_samples, is_first_chunk, is_last_chunk) entry point, unit-tested with a mock backend;
the b2k pipeline backend type- and link-checked against libb2k.so; the Kaldi-typed shims type-checked against the
reference's own headers where those exist (this container only)."""
import os
import shutil
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "kaldi_b200", "host")


def _gxx():
    if not shutil.which("g++"):
        pytest.skip("g++ missing")


def test_batcher_logic_with_mock_backend(tmp_path):
    _gxx()
    exe = str(tmp_path / "batcher_test")
    r = subprocess.run(["g++", "-std=c++17", "-Wall", "-Wextra", "-Werror", "-I" + HOST,
                        os.path.join(ROOT, "tests", "cabi", "batcher_test.cc"), "-o", exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "batcher ok" in r.stdout


def test_dynamic_batcher_with_mock_pipeline(tmp_path):
    """b2k_host::DynamicBatcher (the role of cuda_decoder::CudaOnlinePipelineDynamicBatcher): four producer threads, 40 streams on
    12 channels, batches of 8 -- one chunk per stream and batch, chunks in order, nothing lost, channels respected; the timeout
    path, the full-batch path, empty last chunks, a reused id, an exception from the pipeline."""
    _gxx()
    exe = str(tmp_path / "dynamic_batcher_test")
    r = subprocess.run(["g++", "-std=c++17", "-Wall", "-Wextra", "-Werror", "-pthread", "-I" + HOST,
                        os.path.join(ROOT, "tests", "cabi", "dynamic_batcher_test.cc"), "-o", exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    for _ in range(3):
        r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
        assert r.returncode == 0 and "dynamic batcher ok" in r.stdout, r.stdout + r.stderr


def test_utterance_pump_with_mock_pipeline(tmp_path):
    """b2k_host::UtterancePump (the role of BatchedThreadedNnet3CudaPipeline2's control thread: whole utterances through the
    chunk-at-a-time pipeline): every sample once and in order, flags, batch sizes, eager and draining runs."""
    _gxx()
    exe = str(tmp_path / "utterance_pump_test")
    r = subprocess.run(["g++", "-std=c++17", "-Wall", "-Wextra", "-Werror", "-I" + HOST,
                        os.path.join(ROOT, "tests", "cabi", "utterance_pump_test.cc"), "-o", exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "utterance pump ok" in r.stdout, r.stdout + r.stderr


def test_pipeline_backend_links_and_fails_loudly_without_a_device(tmp_path):
    _gxx()
    so = os.path.join(ROOT, "kaldi_b200", "libb2k.so")
    if not os.path.exists(so):
        pytest.skip("libb2k.so not built")
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present: the device route is covered by the gpu tests")
    exe = str(tmp_path / "pslink")
    r = subprocess.run(["g++", "-std=c++17", "-Wall", "-Wextra", "-I" + os.path.join(ROOT, "include"), "-I" + HOST,
                        os.path.join(ROOT, "tests", "cabi", "pipeline_shim_link.cc"), "-o", exe, "-L" + os.path.dirname(so),
                        "-lb2k", "-Wl,-rpath," + os.path.dirname(so)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([exe, os.path.join(ROOT, "tests", "golden", "tiny_final.mdl")], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "no CUDA device" in r.stdout


def test_kaldi_typed_shims_compile_against_the_reference_headers():
    if not os.path.isdir("/root/reference/src"):
        pytest.skip("the reference tree exists in the build container only")
    _gxx()
    sys.path.insert(0, ROOT)
    from oracle import check_shims
    assert check_shims.check()
