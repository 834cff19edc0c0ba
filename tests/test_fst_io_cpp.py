"""The C++ OpenFst-file reader (kaldi_b200/csrc/fst_io.cu).  PARITY UNPINNED, like kaldi_io.read_openfst: OpenFst is
absent from this image, so the reader is only checked against the Python writer/reader of the published layout."""
import struct

import numpy as np
import pytest

from kaldi_b200 import kaldi_io as KIO, synth


def _read(path):
    try:
        from kaldi_b200.decoder import read_fst_file
        return read_fst_file(path)
    except OSError as e:
        pytest.skip(str(e))


@pytest.mark.parametrize("fst_type,aligned", [("const", False), ("const", True), ("vector", False)])
def test_cpp_reader_returns_the_graph_that_was_written(tmp_path, fst_type, aligned):
    g = synth.make_hclg(20_000, num_pdfs=50, seed=3)
    p = str(tmp_path / "HCLG.fst")
    KIO.write_openfst(p, g, fst_type, aligned)
    h, py = _read(p), KIO.read_openfst(p)
    assert (h["num_states"], h["start"], h["fst_type"]) == (g["num_states"], g["start"], fst_type)
    for k in ("offsets", "ilabel", "olabel", "nextstate"):
        np.testing.assert_array_equal(h[k], py[k], err_msg=k)
        np.testing.assert_array_equal(h[k], np.asarray(g[k], np.int32)[:len(h[k])], err_msg=k)
    for k in ("weight", "final"):
        np.testing.assert_array_equal(h[k].view(np.int32), py[k].view(np.int32), err_msg=k)
    np.testing.assert_array_equal(h["weight"].view(np.int32), np.asarray(g["weight"], np.float32).view(np.int32))


def test_symbol_tables_are_skipped(tmp_path):
    g = synth.make_hclg(2_000, num_pdfs=20, seed=5)
    p = str(tmp_path / "plain.fst")
    KIO.write_openfst(p, g, "vector")
    d = open(p, "rb").read()

    def fstr(x):
        return struct.pack("<i", len(x)) + x

    def symtab(words):
        out = struct.pack("<i", 2125658996) + fstr(b"syms") + struct.pack("<qq", len(words), len(words))
        for i, w in enumerate(words):
            out += fstr(w) + struct.pack("<q", i)
        return out
    hdr_len = 4 + 4 + len("vector") + 4 + len("standard") + 8 + 32
    hdr = bytearray(d[:hdr_len])
    flags_at = 4 + 4 + len("vector") + 4 + len("standard") + 4
    struct.pack_into("<i", hdr, flags_at, 3)                       # HAS_ISYMBOLS | HAS_OSYMBOLS
    q = str(tmp_path / "syms.fst")
    open(q, "wb").write(bytes(hdr) + symtab([b"<eps>", b"a", b"bb"]) + symtab([b"<eps>", b"hello"]) + d[hdr_len:])
    a, b = _read(p), _read(q)
    for k in ("offsets", "ilabel", "olabel", "nextstate", "weight", "final"):
        np.testing.assert_array_equal(a[k], b[k], err_msg=k)
    py = KIO.read_openfst(q)
    np.testing.assert_array_equal(py["ilabel"], b["ilabel"])


def test_malformed_files_are_rejected(tmp_path):
    from kaldi_b200.decoder import read_fst_file
    g = synth.make_hclg(2_000, num_pdfs=20, seed=5)
    p = str(tmp_path / "ok.fst")
    KIO.write_openfst(p, g, "const")
    d = open(p, "rb").read()
    cases = {"magic": b"\0\0\0\0" + d[4:], "truncated": d[:len(d) // 2], "empty": b"",
             "arctype": d[:4 + 4 + 5 + 4] + b"standarX" + d[4 + 4 + 5 + 4 + 8:]}
    for name, blob in cases.items():
        q = str(tmp_path / (name + ".fst"))
        open(q, "wb").write(blob)
        with pytest.raises(RuntimeError):
            read_fst_file(q)
    with pytest.raises(RuntimeError):
        read_fst_file(str(tmp_path / "missing.fst"))


def test_device_creation_fails_loudly_without_a_gpu(tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from kaldi_b200.decoder import CudaFst
    g = synth.make_hclg(2_000, num_pdfs=20, seed=5)
    p = str(tmp_path / "HCLG.fst")
    KIO.write_openfst(p, g, "const")
    with pytest.raises(RuntimeError):
        CudaFst.from_file(p, g["tid2pdf"])
