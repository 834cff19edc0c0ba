"""Readers (and the writers the tests need) for the Kaldi on-disk formats that sit
on the input side of the hot path — SURVEY.md §8(f) row 2, built without
Kaldi/OpenFst:

* basic types, tokens, float/double matrices and vectors in binary and text mode
  (base/io-funcs{,-inl.h}.cc, matrix/kaldi-matrix.cc Read/Write :1375-1545,
  kaldi-vector.cc :1130-1260),
* "raw" nnet3 models as Nnet::Write emits them (nnet3/nnet-nnet.cc:630-656):
  config lines + components, every component parsed generically into
  {token: value}, and the mapping of a TDNN-F chain model (the xconfig layer
  names of steps/libs/nnet3/xconfig) onto the arch/weights dictionaries of
  kaldi_b200.nnet_model,
* DiagGmm (gmm/diag-gmm.cc:835-895), IvectorExtractor (ivector/ivector-extractor.cc:
  807-870), global CMVN stats.

Pinned by tests/test_kaldi_io.py against files written by the reference's own
writers (oracle/_ref).  Host-side only; nothing here touches the GPU.
"""
from __future__ import annotations

import io
import re
import struct

import numpy as np


class KaldiFormatError(ValueError):
    pass


def _format_errors(fn):
    """Readers see files that may be cut short or corrupted: whatever goes wrong while parsing (a struct that cannot be
    unpacked, bytes that are not ASCII, a field that is missing, an index outside a table) is reported as KaldiFormatError."""
    import functools

    @functools.wraps(fn)
    def wrapped(*a, **kw):
        try:
            return fn(*a, **kw)
        except KaldiFormatError:
            raise
        except (ValueError, KeyError, IndexError, struct.error, UnicodeDecodeError, OverflowError, AttributeError, TypeError) as e:
            raise KaldiFormatError(f"{fn.__name__}: malformed file ({type(e).__name__}: {e})") from e
    return wrapped


class Reader:
    """Sequential reader over a Kaldi object file (binary files start with "\\0B")."""

    def __init__(self, data: bytes):
        self.d = data
        self.p = 0
        self.binary = data[:2] == b"\0B"
        if self.binary:
            self.p = 2

    @classmethod
    def open(cls, path: str) -> "Reader":
        with open(path, "rb") as f:
            return cls(f.read())

    # ---- low level
    def _ws(self):
        while self.p < len(self.d) and self.d[self.p:self.p + 1].isspace():
            self.p += 1

    def eof(self) -> bool:
        if not self.binary:
            self._ws()
        return self.p >= len(self.d)

    def peek_byte(self) -> int:
        if not self.binary:
            self._ws()
        if self.p >= len(self.d):
            raise KaldiFormatError("unexpected end of file")
        return self.d[self.p]

    def _count(self, n: int, elem_bytes: int) -> int:
        """A count read from the file: non-negative and no larger than the bytes that are left (np.frombuffer would read
        count = -1 as 'everything', and a position that moves backwards never reaches the end of the file)."""
        if n < 0 or n * elem_bytes > len(self.d) - self.p:
            raise KaldiFormatError(f"implausible element count {n} at byte {self.p}")
        return n

    def read_token(self) -> str:
        """ReadToken (io-funcs.cc:154): whitespace-delimited, one trailing space consumed."""
        self._ws() if not self.binary else None
        # tokens are written as "<tok> " in both modes; in binary mode there is no leading whitespace,
        # except the newline Nnet::Write puts after the config section
        while self.p < len(self.d) and self.d[self.p:self.p + 1].isspace():
            self.p += 1
        e = self.p
        while e < len(self.d) and not self.d[e:e + 1].isspace():
            e += 1
        tok = self.d[self.p:e].decode("ascii")
        self.p = e + 1
        if not tok:
            raise KaldiFormatError("empty token")
        return tok

    def expect_token(self, want: str):
        got = self.read_token()
        if got != want:
            raise KaldiFormatError(f"expected token {want}, got {got} at byte {self.p}")

    def read_line(self) -> str:
        e = self.d.index(b"\n", self.p)
        s = self.d[self.p:e].decode("ascii")
        self.p = e + 1
        return s

    def _text_number(self) -> str:
        self._ws()
        e = self.p
        while e < len(self.d) and not self.d[e:e + 1].isspace():
            e += 1
        if e == self.p:
            raise KaldiFormatError("unexpected end of file (number expected)")
        s = self.d[self.p:e].decode("ascii")
        self.p = e
        return s

    # ---- basic types (io-funcs-inl.h:34-110, io-funcs.cc:26-130)
    def read_int(self) -> int:
        if not self.binary:
            return int(self._text_number())
        n = struct.unpack_from("b", self.d, self.p)[0]
        self.p += 1
        size = abs(n)
        fmt = {1: "b", 2: "h", 4: "i", 8: "q"}[size] if n > 0 else {1: "B", 2: "H", 4: "I", 8: "Q"}[size]
        v = struct.unpack_from("<" + fmt, self.d, self.p)[0]
        self.p += size
        return v

    def read_float(self) -> float:
        if not self.binary:
            s = self._text_number()
            return float({"inf": "inf", "-inf": "-inf", "nan": "nan"}.get(s.lower(), s))
        size = self.d[self.p]
        self.p += 1
        if size == 4:
            v = struct.unpack_from("<f", self.d, self.p)[0]
        elif size == 8:
            v = struct.unpack_from("<d", self.d, self.p)[0]
        else:
            raise KaldiFormatError(f"expected a float, saw size byte {size}")
        self.p += size
        return v

    def read_bool(self) -> bool:
        c = chr(self.peek_byte())
        if c not in "TF":
            raise KaldiFormatError(f"expected T/F, saw {c!r}")
        self.p += 1
        return c == "T"

    def read_int_vector(self) -> np.ndarray:
        """ReadIntegerVector (io-funcs-inl.h:232-290)."""
        if self.binary:
            size = self.d[self.p]
            self.p += 1
            n = struct.unpack_from("<i", self.d, self.p)[0]
            self.p += 4
            dt = {1: np.int8, 2: np.int16, 4: np.int32, 8: np.int64}[size]
            v = np.frombuffer(self.d, dt, self._count(n, size), self.p).copy()
            self.p += n * size
            return v
        self._ws()
        if self.d[self.p:self.p + 1] != b"[":
            raise KaldiFormatError("expected [ at the start of an integer vector")
        e = self.d.index(b"]", self.p)
        v = np.array(self.d[self.p + 1:e].split(), dtype=np.int64)
        self.p = e + 1
        return v

    def read_int_pair_vector(self) -> np.ndarray:
        """ReadIntegerPairVector (io-funcs-inl.h:113-190): [n, 2]; text form "[ a,b c,d ]"."""
        if self.binary:
            size = self.d[self.p]
            self.p += 1
            n = struct.unpack_from("<i", self.d, self.p)[0]
            self.p += 4
            dt = {1: np.int8, 2: np.int16, 4: np.int32, 8: np.int64}[size]
            v = np.frombuffer(self.d, dt, self._count(2 * n, size) if n >= 0 else self._count(n, size), self.p).reshape(n, 2).copy()
            self.p += 2 * n * size
            return v
        self._ws()
        e = self.d.index(b"]", self.p)
        items = self.d[self.p + 1:e].split()
        self.p = e + 1
        return np.array([[int(x) for x in it.split(b",")] for it in items], dtype=np.int64).reshape(-1, 2)

    # ---- matrices and vectors
    def _text_brackets(self) -> np.ndarray:
        self._ws()
        if self.d[self.p:self.p + 1] != b"[":
            raise KaldiFormatError(f"expected [ at byte {self.p}")
        e = self.d.index(b"]", self.p)
        body = self.d[self.p + 1:e].decode("ascii")
        self.p = e + 1
        rows = [r.split() for r in body.split("\n") if r.strip()]
        return rows

    def read_vector(self) -> np.ndarray:
        if self.binary:
            tag = self.d[self.p:self.p + 3]
            if tag not in (b"FV ", b"DV "):
                raise KaldiFormatError(f"expected FV/DV, saw {tag!r}")
            self.p += 3
            n = self.read_int()
            dt = np.float32 if tag == b"FV " else np.float64
            v = np.frombuffer(self.d, dt, self._count(n, np.dtype(dt).itemsize), self.p).copy()
            self.p += n * v.itemsize
            return v
        rows = self._text_brackets()
        flat = [x for r in rows for x in r]
        return np.array(flat, dtype=np.float64).astype(np.float32) if flat else np.zeros(0, np.float32)

    def read_matrix(self, text_dtype=np.float32) -> np.ndarray:
        """text_dtype: a text matrix carries no type tag; the reference parses it into the precision of the object it
        is read into (Matrix<double> for the i-vector extractor), so the caller says which."""
        if self.binary:
            tag = self.d[self.p:self.p + 3]
            if tag[:2] == b"CM":
                raise KaldiFormatError("compressed matrices are not supported")
            if tag not in (b"FM ", b"DM "):
                raise KaldiFormatError(f"expected FM/DM, saw {tag!r}")
            self.p += 3
            r, c = self.read_int(), self.read_int()
            dt = np.float32 if tag == b"FM " else np.float64
            if r < 0 or c < 0:
                raise KaldiFormatError("negative matrix size")
            m = np.frombuffer(self.d, dt, self._count(r * c, np.dtype(dt).itemsize), self.p).reshape(r, c).copy()
            self.p += r * c * m.itemsize
            return m
        rows = self._text_brackets()
        if not rows:
            return np.zeros((0, 0), np.float32)
        return np.array(rows, dtype=np.float64).astype(text_dtype)

    def read_packed(self) -> np.ndarray:
        """SpMatrix/TpMatrix (matrix/packed-matrix.cc:236-330): packed lower triangle -> full symmetric matrix."""
        if self.binary:
            tag = self.d[self.p:self.p + 3]
            if tag not in (b"FP ", b"DP "):
                raise KaldiFormatError(f"expected FP/DP, saw {tag!r}")
            self.p += 3
            n = self.read_int()
            dt = np.float32 if tag == b"FP " else np.float64
            ne = n * (n + 1) // 2
            if n < 0:
                raise KaldiFormatError("negative matrix size")
            flat = np.frombuffer(self.d, dt, self._count(ne, np.dtype(dt).itemsize), self.p).copy()
            self.p += ne * flat.itemsize
        else:
            rows = self._text_brackets()
            n = len(rows)
            flat = np.array([x for r in rows for x in r], dtype=np.float64)
        m = np.zeros((n, n), flat.dtype)
        il = np.tril_indices(n)
        m[il] = flat
        m[(il[1], il[0])] = flat
        return m

    # ---- generic "<Token> value value ..." records
    def _binary_value(self):
        b = self.d[self.p]
        nxt = self.d[self.p:self.p + 3]
        if nxt in (b"FM ", b"DM "):
            return self.read_matrix()
        if nxt in (b"FV ", b"DV "):
            return self.read_vector()
        if nxt[:2] == b"CM":
            raise KaldiFormatError("compressed matrices are not supported")
        if b in (ord("T"), ord("F")):
            return self.read_bool()
        if b in (4, 8):                 # float/double or int32/int64: keep the raw bytes, typed on demand
            raw = self.d[self.p + 1:self.p + 1 + b]
            self.p += 1 + b
            return RawScalar(raw)
        if b in (1, 2, 0xFF, 0xFE, 0xFC):
            return self.read_int()
        raise KaldiFormatError(f"cannot parse a value at byte {self.p} (0x{b:02x})")

    def read_fields(self, end_token: str, int_vector_tokens=("<TimeOffsets>", "<RequiredTimeOffsets>", "<ColumnMap>"),
                    pair_vector_tokens=("<Offsets>",)) -> dict:
        """Parses "<A> v <B> v v ... </End>" into {"<A>": [values], ...} (order kept)."""
        out = {}
        while True:
            tok = self.read_token()
            if tok == end_token:
                return out
            if not tok.startswith("<"):
                raise KaldiFormatError(f"expected a token, got {tok!r}")
            vals = []
            while True:
                if self.binary:
                    if self.d[self.p:self.p + 1] == b"<":
                        break
                    if tok in pair_vector_tokens:
                        vals.append(self.read_int_pair_vector())
                    elif tok in int_vector_tokens:
                        vals.append(self.read_int_vector())
                    else:
                        vals.append(self._binary_value())
                else:
                    self._ws()
                    c = self.d[self.p:self.p + 1]
                    if c == b"<":
                        break
                    if c == b"[" and tok in pair_vector_tokens:
                        vals.append(self.read_int_pair_vector())
                    elif c == b"[":
                        rows = self._text_brackets()
                        if tok in int_vector_tokens:
                            vals.append(np.array([x for r in rows for x in r], dtype=np.int64))
                        elif len(rows) <= 1:
                            vals.append(np.array(rows[0] if rows else [], dtype=np.float64).astype(np.float32))
                        else:
                            vals.append(np.array(rows, dtype=np.float64).astype(np.float32))
                    else:
                        s = self._text_number()
                        vals.append(True if s == "T" else False if s == "F" else RawScalar(text=s))
            out[tok] = vals


class RawScalar:
    """A 4/8-byte binary scalar (or a text number) whose type the format does not reveal."""

    def __init__(self, raw: bytes = b"", text: str | None = None):
        self.raw, self.text = raw, text

    def as_int(self) -> int:
        if self.text is not None:
            return int(self.text)
        return struct.unpack("<i" if len(self.raw) == 4 else "<q", self.raw)[0]

    def as_float(self) -> float:
        if self.text is not None:
            return float(self.text)
        return struct.unpack("<f" if len(self.raw) == 4 else "<d", self.raw)[0]

    def __repr__(self):
        return f"RawScalar({self.text if self.text is not None else self.raw.hex()})"


# ----------------------------------------------------------------------------- files

@_format_errors
def read_matrix(path: str) -> np.ndarray:
    return Reader.open(path).read_matrix()


@_format_errors
def read_vector(path: str) -> np.ndarray:
    return Reader.open(path).read_vector()


def write_matrix(path: str, m: np.ndarray, binary: bool = True) -> None:
    m = np.ascontiguousarray(m)
    with open(path, "wb") as f:
        if binary:
            tag = b"DM " if m.dtype == np.float64 else b"FM "
            mm = m if m.dtype == np.float64 else m.astype(np.float32)
            f.write(b"\0B" + tag + b"\4" + struct.pack("<i", m.shape[0]) + b"\4" + struct.pack("<i", m.shape[1]))
            f.write(mm.tobytes())
        else:
            f.write(b" [\n")
            for r in m:
                f.write(("  " + " ".join(repr(float(np.float32(x))) for x in r) + "\n").encode())
            f.seek(-1, io.SEEK_END)
            f.write(b" ]\n")


def write_vector(path: str, v: np.ndarray, binary: bool = True) -> None:
    v = np.ascontiguousarray(v)
    with open(path, "wb") as f:
        if binary:
            tag = b"DV " if v.dtype == np.float64 else b"FV "
            vv = v if v.dtype == np.float64 else v.astype(np.float32)
            f.write(b"\0B" + tag + b"\4" + struct.pack("<i", v.shape[0]) + vv.tobytes())
        else:
            f.write((" [ " + " ".join(repr(float(x)) for x in v) + " ]\n").encode())


# ----------------------------------------------------------------------------- nnet3 raw models

@_format_errors
def read_nnet3_raw(path: str) -> dict:
    """{"config": [lines], "components": {name: {"type": T, "<Token>": [values], ...}}} (Nnet::Read, nnet-nnet.cc:586)."""
    return _read_nnet3(Reader.open(path))


def _read_nnet3(r: Reader) -> dict:
    r.expect_token("<Nnet3>")
    lines = []
    # config-like section: text lines up to the first blank line, in both modes
    if r.d[r.p:r.p + 1] == b"\n":
        r.p += 1
    while True:
        line = r.read_line()
        if not line.strip():
            if lines:
                break
            continue
        lines.append(line.strip())
    r.expect_token("<NumComponents>")
    n = r.read_int()
    comps = {}
    for _ in range(n):
        r.expect_token("<ComponentName>")
        name = r.read_token()
        typ = r.read_token()
        if not (typ.startswith("<") and typ.endswith(">")):
            raise KaldiFormatError(f"bad component type token {typ!r}")
        fields = r.read_fields("</" + typ[1:])
        fields["type"] = typ[1:-1]
        comps[name] = fields
    r.expect_token("</Nnet3>")
    return {"config": lines, "components": comps}


def _f(fields: dict, tok: str, i: int = 0):
    return fields[tok][i]


def _parse_config_line(line: str) -> tuple[str, dict]:
    kind, rest = line.split(" ", 1)
    kv = {}
    for m in re.finditer(r"(\S+?)=(.*?)(?=\s+\S+?=|$)", rest):
        kv[m.group(1)] = m.group(2).strip()
    return kind, kv


_DESCRIPTOR_WORDS = {"Append", "Offset", "Sum", "Scale", "ReplaceIndex", "Round", "IfDefined", "Failover", "Switch", "Const", "t", "x"}
_IDENTITY_AT_TEST_TIME = ("GeneralDropoutComponent", "DropoutComponent", "SpecAugmentTimeMaskComponent")   # SetDropoutTestMode(true), online2-wav-nnet3-latgen-faster.cc:147


def _descriptor_nodes(desc: str) -> list[str]:
    return [w for w in re.findall(r"[A-Za-z_][A-Za-z0-9_.\-]*", desc) if w not in _DESCRIPTOR_WORDS]


def _splice_offsets(desc: str, expect_src: str | None = None) -> list[int]:
    """The time offsets over which an input descriptor splices ONE source node: `x` -> [0];
    `Append(Offset(x, -1), x, Offset(x, 1))` -> [-1, 0, 1] (what xconfig writes for input=Append(-1,0,1)); a trailing
    term on the ivector input is not part of the splice.  Anything else (two sources, nested expressions, a source other than
    `expect_src` = the layer before, i.e. a skip connection) raises."""
    d = desc.strip()
    if not (d.startswith("Append(") and d.endswith(")")):
        if expect_src is not None and d != expect_src:
            raise KaldiFormatError(f"input {desc} is not the layer before ({expect_src}): skip connections are not supported")
        return [0]
    terms, depth, start = [], 0, 7
    for p in range(7, len(d) - 1):
        depth += d[p] == "("
        depth -= d[p] == ")"
        if d[p] == "," and depth == 0:
            terms.append(d[start:p])
            start = p + 1
    terms.append(d[start:len(d) - 1])
    offs, src = [], None
    for k, term in enumerate(terms):
        term = term.strip()
        if "ivector" in _descriptor_nodes(term):
            if k != len(terms) - 1:
                raise KaldiFormatError(f"unsupported input descriptor {desc}")
            continue
        m = re.fullmatch(r"Offset\(\s*([^,()\s]+)\s*,\s*(-?\d+)\s*\)", term)
        node, o = (m.group(1), int(m.group(2))) if m else (term, 0)
        if not re.fullmatch(r"[A-Za-z_][A-Za-z0-9_.\-]*", node) or (src is not None and node != src):
            raise KaldiFormatError(f"unsupported input descriptor {desc}")
        src = node
        offs.append(o)
    if not offs or len(offs) > 8:
        raise KaldiFormatError(f"unsupported input descriptor {desc}")
    if expect_src is not None and src != expect_src:
        raise KaldiFormatError(f"input {desc} is not the layer before ({expect_src}): skip connections are not supported")
    return offs


def _inference_view(nodes: list, comps: dict) -> list:
    """What a trained recipe model looks like to the decoder: (1) dropout components are the identity in test mode, so
    every reference to such a node is replaced by the node's own input; (2) only what the output node named "output"
    depends on is kept (chain recipes leave their cross-entropy branch, output-xent, in final.mdl)."""
    # name -> descriptor text that replaces it.  Nodes are defined before they are used, so one pass in file order is enough.
    alias = {}

    def subst(desc):
        return re.sub(r"[A-Za-z_][A-Za-z0-9_.\-]*", lambda m: alias.get(m.group(0), m.group(0)), desc)
    out = []
    for kind, kv in nodes:
        kv = dict(kv)
        for key in ("input", "input-node"):
            if key in kv:
                kv[key] = subst(kv[key])
        if kind == "component-node":
            typ = comps[kv["component"]]["type"]
            src = kv["input"].strip()
            if typ in _IDENTITY_AT_TEST_TIME:
                if _descriptor_nodes(src) != [src]:
                    raise KaldiFormatError(f"dropout node {kv['name']} has a compound input descriptor: {src}")
                alias[kv["name"]] = src
                continue
            # no-op-component (trivial_layers.py): a name for a descriptor, e.g. input2 = Append(delta, Scale(0.4, ivector)) in
            # run_tdnn_1k.sh:181.  The NoOp nodes that belong to a layer pattern stay: the delta-layer's "<input>_2" and the
            # tdnnf-layer's "<name>.noop".
            if typ == "NoOpComponent" and not kv["name"].endswith(".noop") and not (kv["name"].endswith("_2") and "_copy1" in src):
                alias[kv["name"]] = src
                continue
        out.append((kind, kv))
    roots = [kv for kind, kv in out if kind == "output-node" and kv["name"] == "output"]
    if not roots:
        return out
    by_name = {kv["name"]: (kind, kv) for kind, kv in out}
    keep, todo = set(), ["output"]
    while todo:
        n = todo.pop()
        if n in keep or n not in by_name:
            continue
        keep.add(n)
        kind, kv = by_name[n]
        todo += _descriptor_nodes(kv.get("input", "")) + _descriptor_nodes(kv.get("input-node", ""))
    return [(kind, kv) for kind, kv in out if kv["name"] in keep]


@_format_errors
def nnet3_to_arch(parsed: dict, name: str = "from_file", frame_subsampling_factor: int | None = None) -> tuple[dict, dict]:
    """Maps a parsed TDNN-F chain model onto (arch, weights) of kaldi_b200.nnet_model.

    Recognises the node patterns that steps/libs/nnet3/xconfig emits for: idct-layer /
    FixedAffine lda, batchnorm-component, spec-augment-free delta-layer, relu-batchnorm-layer
    (optionally with Scale(s, ReplaceIndex(ivector, t, 0)) appended), tdnnf-layer,
    linear-component, prefinal-layer, output-layer; anything else raises."""
    comps = parsed["components"]
    nodes = []
    dims = {}
    for line in parsed["config"]:
        kind, kv = _parse_config_line(line)
        if kind == "input-node":
            dims[kv["name"]] = int(kv["dim"])
        elif kind in ("component-node", "dim-range-node", "output-node"):
            nodes.append((kind, kv))
    W = {}
    layers = []
    arch = {"name": name, "feat_dim": dims["input"], "ivector_dim": dims.get("ivector", 0)}

    def mat(c, tok):
        f = comps[c]
        if tok == "<LinearParams>" and tok not in f:      # LinearComponent calls its matrix <Params>
            tok = "<Params>"
        return np.ascontiguousarray(_f(f, tok), np.float32)

    def bn(c, key):
        f = comps[c]
        assert f["type"] == "BatchNormComponent", (c, f["type"])
        count = _f(f, "<Count>").as_float()
        mean, var = np.asarray(_f(f, "<StatsMean>"), np.float64), np.asarray(_f(f, "<StatsVar>"), np.float64)
        # BatchNormComponent::Write stores mean and uncentered variance times nothing: Read() (nnet-normalize-component.cc
        # :591-614) takes <StatsMean> as the mean and <StatsVar> as the variance when written after ComputeDerived.
        W[key + ".mean"] = mean.astype(np.float32)
        W[key + ".var"] = var.astype(np.float32)
        return count

    node_dim = {"input": dims["input"], "ivector": dims.get("ivector", 0)}
    nodes = _inference_view(nodes, comps)
    cn = [(kv["name"], kv) for kind, kv in nodes if kind == "component-node"]
    for n, kv in cn:     # nnet3::CollapseModel names a merged component "<first>.<second>" and leaves the node its name
        c = kv.get("component", n)
        if c != n and (c.endswith("." + n) or c.startswith(n + ".")):
            raise KaldiFormatError(f"component {c} of node {n} is a merged one: the network has been through nnet3::CollapseModel; "
                                   "b2k folds batch-norm and dropout itself and takes the model as it was trained")
    names = [n for n, _ in cn]
    inputs = {n: kv["input"] for n, kv in cn}
    i = 0
    sub = None
    while i < len(cn):
        n, kv = cn[i]
        c = comps[kv["component"]]
        t = c["type"]
        if t == "LinearComponent" and "ivector" in inputs[n] and i + 1 < len(cn) \
                and comps[cn[i + 1][1]["component"]]["type"] == "BatchNormComponent" and n.endswith("-linear"):
            base = n[:-7]                       # linear-component + batchnorm-component on ReplaceIndex(ivector, t, 0)
            w = mat(n, "<LinearParams>")
            bnn = cn[i + 1][0]
            layers.append({"type": "ivector-linear-bn", "name": base, "dim": int(w.shape[0]),
                           "target_rms": round(_f(comps[bnn], "<TargetRms>").as_float(), 6)})
            W[n + ".w"] = w
            bn(bnn, bnn)
            node_dim[bnn] = int(w.shape[0])
            i += 2
        elif t == "PermuteComponent":
            m = re.match(r"Append\((\S+),\s*(\S+)\)", inputs[n])
            if not m:
                raise KaldiFormatError(f"unsupported PermuteComponent input {inputs[n]}")
            main, side = m.group(1), m.group(2)
            nxt = comps[cn[i + 1][1]["component"]]
            if nxt["type"] != "TimeHeightConvolutionComponent":
                raise KaldiFormatError("PermuteComponent is only supported as combine-feature-maps in front of a convolution")
            h = _f(nxt, "<HeightIn>").as_int()
            f1, f2 = node_dim[main] // h, node_dim[side] // h
            want = []
            for hh in range(h):
                want += [hh * f1 + f for f in range(f1)] + [h * f1 + hh * f2 + f for f in range(f2)]
            if np.asarray(_f(c, "<ColumnMap>")).tolist() != want:
                raise KaldiFormatError("PermuteComponent column map is not a combine-feature-maps interleave")
            layers.append({"type": "combine", "name": n, "side": side, "height": h, "filters1": f1, "filters2": f2})
            node_dim[n] = node_dim[main] + node_dim[side]
            i += 1
        elif t == "TimeHeightConvolutionComponent" and n.endswith(".conv"):
            base = n[:-5]
            offs = np.asarray(_f(c, "<Offsets>"))
            L = {"type": "conv", "name": base, "height_in": _f(c, "<HeightIn>").as_int(),
                 "height_out": _f(c, "<HeightOut>").as_int(), "height_subsample_out": _f(c, "<HeightSubsampleOut>").as_int(),
                 "filters_in": _f(c, "<NumFiltersIn>").as_int(), "filters_out": _f(c, "<NumFiltersOut>").as_int(),
                 "time_offsets": sorted(set(offs[:, 0].tolist())), "height_offsets": sorted(set(offs[:, 1].tolist()))}
            if len(offs) != len(L["time_offsets"]) * len(L["height_offsets"]):
                raise KaldiFormatError("convolution offsets are not a full time x height grid")
            if sorted(np.asarray(_f(c, "<RequiredTimeOffsets>")).tolist()) != L["time_offsets"]:
                raise KaldiFormatError("time zero-padding (required-time-offsets) is not supported")
            if [x[0] for x in cn[i + 1:i + 3]] != [base + ".relu", base + ".batchnorm"]:
                raise KaldiFormatError(f"expected conv-relu-batchnorm at {base}")
            layers.append(L)
            W[n + ".w"], W[n + ".b"] = mat(n, "<LinearParams>"), mat(n, "<BiasParams>")
            bn(base + ".batchnorm", base + ".batchnorm")
            node_dim[base + ".batchnorm"] = L["height_out"] * L["filters_out"]
            i += 3
        elif t == "FixedAffineComponent" and inputs[n] == "input":
            W[n + ".w"], W[n + ".b"] = mat(n, "<LinearParams>"), mat(n, "<BiasParams>")
            layers.append({"type": "idct", "name": n, "dim": int(W[n + ".w"].shape[0])})
            node_dim[n] = int(W[n + ".w"].shape[0])
            i += 1
        elif t == "FixedAffineComponent":
            L = {"type": "lda", "name": n}
            sp = _splice_offsets(inputs[n], cn[i - 1][0] if i else "input")
            if sp != [-1, 0, 1]:
                L["time_offsets"] = sp
            layers.append(L)
            W[n + ".w"], W[n + ".b"] = mat(n, "<LinearParams>"), mat(n, "<BiasParams>")
            i += 1
        elif t == "BatchNormComponent" and i + 1 < len(cn) and comps[cn[i + 1][1]["component"]]["type"] == "NoOpComponent" \
                and cn[i + 1][0].endswith("_2") and "_copy1" in inputs[cn[i + 1][0]]:
            # batchnorm-component followed by delta-layer (its NoOp is named <input descriptor>_2 and reads the
            # dim-range copies <input descriptor>_copy1/2, trivial_layers.py:236-256; the input descriptor is the
            # batchnorm itself or a spec-augment-layer behind it)
            layers.append({"type": "batchnorm", "name": n})
            bn(n, n)
            dn = cn[i + 2][0]
            layers.append({"type": "delta", "name": dn})
            bn(dn, dn)
            i += 3
        elif t == "BatchNormComponent":
            layers.append({"type": "batchnorm", "name": n})
            bn(n, n)
            node_dim[n] = _f(c, "<Dim>").as_int()
            i += 1
        elif t in ("NaturalGradientAffineComponent", "AffineComponent") and n.endswith(".affine") \
                and i + 2 < len(cn) and cn[i + 1][0] == n[:-7] + ".relu":
            base = n[:-7]
            nxt = [x[0] for x in cn[i:i + 8]]
            if base + ".batchnorm1" in nxt:     # prefinal-layer: affine relu batchnorm1 linear batchnorm2
                big = mat(n, "<LinearParams>").shape[0]
                small = mat(base + ".linear", "<LinearParams>").shape[0]
                layers.append({"type": "prefinal", "name": base, "big": int(big), "small": int(small)})
                W[n + ".w"], W[n + ".b"] = mat(n, "<LinearParams>"), mat(n, "<BiasParams>")
                bn(base + ".batchnorm1", base + ".batchnorm1")
                W[base + ".linear.w"] = mat(base + ".linear", "<LinearParams>")
                bn(base + ".batchnorm2", base + ".batchnorm2")
                i += 5
            else:                               # relu-batchnorm-layer
                L = {"type": "relu-batchnorm", "name": base, "dim": int(mat(n, "<LinearParams>").shape[0])}
                m = re.search(r"Scale\(([0-9.eE+-]+),\s*ivector\)", inputs[n]) or \
                    re.search(r"Scale\(([0-9.eE+-]+),\s*ReplaceIndex\(ivector", inputs[n])
                if "ivector" in inputs[n]:
                    L["append_ivector"] = float(m.group(1)) if m else 1.0
                sp = _splice_offsets(inputs[n], cn[i - 1][0] if i else "input")
                if sp != [0]:
                    L["time_offsets"] = sp
                layers.append(L)
                W[n + ".w"], W[n + ".b"] = mat(n, "<LinearParams>"), mat(n, "<BiasParams>")
                bn(base + ".batchnorm", base + ".batchnorm")
                i += 3
        elif t == "TdnnComponent" and n.endswith(".linear"):
            base = n[:-7]
            offs = np.asarray(_f(c, "<TimeOffsets>")).tolist()
            stride = int(max(abs(o) for o in offs))
            if stride == 3 and sub is None:
                sub = 3
            has_noop = base + ".noop" in inputs
            m = re.search(r"Scale\(([0-9.eE+-]+),", inputs[base + ".noop"]) if has_noop else None
            wl = mat(n, "<LinearParams>")
            wa = mat(base + ".affine", "<LinearParams>")
            layers.append({"type": "tdnnf", "name": base, "dim": int(wa.shape[0]), "bottleneck": int(wl.shape[0]),
                           "stride": stride, "bypass": float(m.group(1)) if m else 0.0})
            W[n + ".w"] = wl
            W[base + ".affine.w"], W[base + ".affine.b"] = wa, mat(base + ".affine", "<BiasParams>")
            bn(base + ".batchnorm", base + ".batchnorm")
            i += 5 if has_noop else 4           # bypass-scale 0: no NoOp node
        elif t == "LinearComponent":
            w = mat(n, "<LinearParams>")
            layers.append({"type": "linear", "name": n, "dim": int(w.shape[0])})
            W[n + ".w"] = w
            i += 1
        elif t in ("NaturalGradientAffineComponent", "AffineComponent") and n.endswith(".affine"):
            base = n[:-7]                       # output-layer: <name>.affine (+ log-softmax for xent outputs)
            w = mat(n, "<LinearParams>")
            L = {"type": "output", "name": base, "dim": int(w.shape[0]), "log_softmax": False}
            W[n + ".w"], W[n + ".b"] = w, mat(n, "<BiasParams>")
            i += 1
            if i < len(cn) and comps[cn[i][1]["component"]]["type"] == "LogSoftmaxComponent":
                L["log_softmax"] = True
                i += 1
            layers.append(L)
        else:
            raise KaldiFormatError(f"unsupported node pattern at {n} ({t})")
    arch["layers"] = layers
    out = [L for L in layers if L["type"] == "output"]
    arch["num_pdfs"] = out[0]["dim"] if out else 0
    # The factor is not stored in the file (the tools take --frame-subsampling-factor, default 1).  A TDNN-F layer with
    # time-stride 3 only exists in chain recipes (factor 3); +-3 splices of relu-batchnorm layers do not decide it (chain
    # run_tdnn_1f.sh has them and so does the plain nnet3 aishell run_tdnn_1a.sh): the caller must then state it.
    stride3 = any(L["type"] == "tdnnf" and L["stride"] == 3 for L in layers)
    splice3 = any(L["type"] == "relu-batchnorm" and 3 in map(abs, L.get("time_offsets", [])) for L in layers)
    arch["frame_subsampling_ambiguous"] = bool(splice3 and not stride3)
    if frame_subsampling_factor is not None:
        if int(frame_subsampling_factor) <= 0 or (stride3 and int(frame_subsampling_factor) != 3):
            raise KaldiFormatError("--frame-subsampling-factor disagrees with the model (TDNN-F layers with time-stride 3: factor 3)")
        arch["frame_subsampling_factor"] = int(frame_subsampling_factor)
        arch["frame_subsampling_ambiguous"] = False
    elif arch["frame_subsampling_ambiguous"]:
        raise KaldiFormatError("the layers do not decide the frame subsampling factor (splices at +-3 without a stride-3 TDNN-F "
                               "layer): pass frame_subsampling_factor (3 for chain models, 1 otherwise)")
    else:
        arch["frame_subsampling_factor"] = 3 if stride3 else 1
    return arch, W


# ----------------------------------------------------------------------------- i-vector extractor side

@_format_errors
def read_diag_gmm(path: str) -> dict:
    """DiagGmm::Read (gmm/diag-gmm.cc:758-800): final.dubm."""
    r = Reader.open(path)
    tok = r.read_token()
    if tok not in ("<DiagGMM>", "<DiagGMMBegin>"):
        raise KaldiFormatError(f"not a DiagGmm: {tok}")
    out = {}
    while True:
        tok = r.read_token()
        if tok in ("</DiagGMM>", "<DiagGMMEnd>"):
            break
        if tok == "<GCONSTS>":
            out["gconsts"] = r.read_vector().astype(np.float32)
        elif tok == "<WEIGHTS>":
            out["ubm_weights"] = r.read_vector().astype(np.float32)
        elif tok == "<MEANS_INVVARS>":
            out["means_invvars"] = r.read_matrix().astype(np.float32)
        elif tok == "<INV_VARS>":
            out["inv_vars"] = r.read_matrix().astype(np.float32)
        else:
            raise KaldiFormatError(f"unexpected token {tok} in DiagGmm")
    out["num_gauss"], out["feat_dim"] = out["means_invvars"].shape
    return out


@_format_errors
def read_ivector_extractor(path: str) -> dict:
    """IvectorExtractor::Read (ivector/ivector-extractor.cc:828-848) + ComputeDerivedVars (:182-230): final.ie.
    Returns M [G, F, D], sigma_inv [G, F, F], w_vec, prior_offset and the derived
    sigma_inv_m [G, F, D] and U [G, D(D+1)/2] (packed lower triangles of M^T Sigma^-1 M), all float64."""
    r = Reader.open(path)
    r.expect_token("<IvectorExtractor>")
    r.expect_token("<w>")
    w = r.read_matrix()
    r.expect_token("<w_vec>")
    w_vec = r.read_vector().astype(np.float64)
    r.expect_token("<M>")
    G = r.read_int()
    M = np.stack([r.read_matrix(text_dtype=np.float64).astype(np.float64) for _ in range(G)])
    r.expect_token("<SigmaInv>")
    sigma_inv = np.stack([r.read_packed().astype(np.float64) for _ in range(G)])
    r.expect_token("<IvectorOffset>")
    prior_offset = r.read_float()
    r.expect_token("</IvectorExtractor>")
    sigma_inv_m = np.einsum("gij,gjk->gik", sigma_inv, M)
    U_full = np.einsum("gji,gjk->gik", M, sigma_inv_m)
    il = np.tril_indices(M.shape[2])
    return dict(num_gauss=G, feat_dim=M.shape[1], ivector_dim=M.shape[2], w=w, w_vec=w_vec, M=M, sigma_inv=sigma_inv,
                prior_offset=float(prior_offset), sigma_inv_m=np.ascontiguousarray(sigma_inv_m),
                U=np.ascontiguousarray(U_full[:, il[0], il[1]]))


@_format_errors
def read_cmvn_stats(path: str) -> np.ndarray:
    """global_cmvn.stats: a 2 x (dim + 1) double matrix (transform/cmvn.cc)."""
    return np.asarray(read_matrix(path), np.float64)


# ----------------------------------------------------------------------------- transition model, final.mdl

def _read_topology(r: Reader) -> dict:
    """HmmTopology::Read (hmm/hmm-topology.cc:38-163).  entries[i] = list of states
    {forward_pdf_class, self_loop_pdf_class, transitions [(dst, prob)]}."""
    r.expect_token("<Topology>")
    entries, phone2idx, phones = [], {}, []
    if not r.binary:
        while True:
            tok = r.read_token()
            if tok == "</Topology>":
                break
            if tok != "<TopologyEntry>":
                raise KaldiFormatError(f"expected <TopologyEntry>, got {tok}")
            r.expect_token("<ForPhones>")
            ph = []
            while True:
                t = r.read_token()
                if t == "</ForPhones>":
                    break
                ph.append(int(t))
            states = []
            while True:
                t = r.read_token()
                if t == "</TopologyEntry>":
                    break
                if t != "<State>":
                    raise KaldiFormatError(f"expected <State>, got {t}")
                idx = r.read_int()
                if idx != len(states):
                    raise KaldiFormatError("states out of order in topology")
                st = {"forward_pdf_class": -1, "self_loop_pdf_class": -1, "transitions": []}
                while True:
                    t = r.read_token()
                    if t == "</State>":
                        break
                    if t == "<PdfClass>":
                        st["forward_pdf_class"] = st["self_loop_pdf_class"] = r.read_int()
                    elif t == "<ForwardPdfClass>":
                        st["forward_pdf_class"] = r.read_int()
                    elif t == "<SelfLoopPdfClass>":
                        st["self_loop_pdf_class"] = r.read_int()
                    elif t in ("<Transition>", "<Final>"):
                        if t == "<Transition>":
                            dst = r.read_int()
                            st["transitions"].append((dst, r.read_float()))
                        else:
                            r.read_float()
                    else:
                        raise KaldiFormatError(f"unexpected token {t} in topology state")
                states.append(st)
            for p in ph:
                phone2idx[p] = len(entries)
            phones += ph
            entries.append(states)
        return {"phones": sorted(phones), "phone2idx": phone2idx, "entries": entries}
    phones = r.read_int_vector().tolist()
    p2i = r.read_int_vector().tolist()
    n = r.read_int()
    is_hmm = True
    if n == -1:                          # the extended format with self-loop pdf classes (:213)
        is_hmm = False
        n = r.read_int()
    for _ in range(n):
        ns = r.read_int()
        states = []
        for _ in range(ns):
            fwd = r.read_int()
            sl = fwd if is_hmm else r.read_int()
            nt = r.read_int()
            tr = []
            for _ in range(nt):
                dst = r.read_int()
                tr.append((dst, r.read_float()))
            states.append({"forward_pdf_class": fwd, "self_loop_pdf_class": sl, "transitions": tr})
        entries.append(states)
    r.expect_token("</Topology>")
    return {"phones": phones, "phone2idx": {p: i for p, i in enumerate(p2i) if i >= 0}, "entries": entries}


def _read_transition_model(r: Reader) -> dict:
    """TransitionModel::Read + ComputeDerived (hmm/transition-model.cc:394-420,144-188).
    tid2pdf[t] for t = 1..num_tids (index 0 is a 0 placeholder, as CudaFst's table wants it)."""
    r.expect_token("<TransitionModel>")
    topo = _read_topology(r)
    tok = r.read_token()
    if tok not in ("<Triples>", "<Tuples>"):
        raise KaldiFormatError(f"expected <Triples>/<Tuples>, got {tok}")
    has_sl = tok == "<Tuples>"
    n = r.read_int()
    tuples = []
    for _ in range(n):
        phone, hs, fwd = r.read_int(), r.read_int(), r.read_int()
        sl = r.read_int() if has_sl else fwd
        tuples.append((phone, hs, fwd, sl))
    r.expect_token("</Triples>" if not has_sl else "</Tuples>")
    r.expect_token("<LogProbs>")
    log_probs = r.read_vector()
    r.expect_token("</LogProbs>")
    r.expect_token("</TransitionModel>")
    tid2pdf, tid2phone, self_loop = [0], [0], [False]
    for phone, hs, fwd, sl in tuples:
        state = topo["entries"][topo["phone2idx"][phone]][hs]
        for dst, _p in state["transitions"]:
            is_sl = dst == hs
            tid2pdf.append(sl if is_sl else fwd)
            tid2phone.append(phone)
            self_loop.append(is_sl)
    if len(log_probs) != len(tid2pdf):
        raise KaldiFormatError("transition model: <LogProbs> does not match the number of transition-ids")
    return dict(topology=topo, tuples=tuples, log_probs=log_probs, tid2pdf=np.array(tid2pdf, np.int32),
                tid2phone=np.array(tid2phone, np.int32), is_self_loop=np.array(self_loop, bool),
                num_pdfs=int(max(max(t[2], t[3]) for t in tuples) + 1))


@_format_errors
def read_transition_model(path: str) -> dict:
    return _read_transition_model(Reader.open(path))


@_format_errors
def read_final_mdl(path: str) -> dict:
    """final.mdl of an nnet3 acoustic model: TransitionModel + AmNnetSimple (nnet3/am-nnet-simple.cc:47-57).
    {"transition_model": ..., "nnet": parsed raw nnet3, "left_context", "right_context", "priors"}."""
    r = Reader.open(path)
    tm = _read_transition_model(r)
    nnet = _read_nnet3(r)
    r.expect_token("<LeftContext>")
    lc = r.read_int()
    r.expect_token("<RightContext>")
    rc = r.read_int()
    r.expect_token("<Priors>")
    priors = r.read_vector()
    return dict(transition_model=tm, nnet=nnet, left_context=lc, right_context=rc, priors=priors)


# ----------------------------------------------------------------------------- OpenFst binary FSTs (HCLG.fst)
#
# PARITY UNPINNED: OpenFst (1.8.4, tools/Makefile:10) is absent from this image and the reference tree holds no
# binary FST, so this reader follows the published on-disk layout of fst/fst.h (FstHeader::Write),
# fst/vector-fst.h (VectorFstImpl::Write, file version 2) and fst/const-fst.h (ConstFstImpl::Write, file version
# 2, aligned version 1) for StdArc ("standard") and has only been checked against its own writer below.

_FST_MAGIC = 2125659606
_SYMTAB_MAGIC = 2125658996
_FST_HAS_ISYMBOLS, _FST_HAS_OSYMBOLS, _FST_IS_ALIGNED = 1, 2, 4


def _fst_string(d: bytes, p: int):
    n = struct.unpack_from("<i", d, p)[0]
    return d[p + 4:p + 4 + n].decode("ascii"), p + 4 + n


def _skip_symbol_table(d: bytes, p: int) -> int:
    magic = struct.unpack_from("<i", d, p)[0]
    if magic != _SYMTAB_MAGIC:
        raise KaldiFormatError("bad symbol table magic")
    _name, p = _fst_string(d, p + 4)
    _avail, size = struct.unpack_from("<qq", d, p)
    p += 16
    for _ in range(size):
        _sym, p = _fst_string(d, p)
        p += 8
    return p


@_format_errors
def read_openfst(path: str) -> dict:
    """An OpenFst binary "vector" or "const" FST over StdArc as the CSR dictionary CudaFst takes:
    num_states, start, offsets [S+1], ilabel/olabel/nextstate [A] int32, weight [A] float32, final [S] float32
    (+inf = not final).  Arc order is file order (= ConstFst arc order, which the decoder's results depend on)."""
    with open(path, "rb") as f:
        d = f.read()
    if struct.unpack_from("<i", d, 0)[0] != _FST_MAGIC:
        raise KaldiFormatError("not an OpenFst binary file (bad magic number)")
    fsttype, p = _fst_string(d, 4)
    arctype, p = _fst_string(d, p)
    version, flags = struct.unpack_from("<ii", d, p)
    p += 8
    _props, start, nstates, narcs = struct.unpack_from("<Qqqq", d, p)
    p += 32
    if arctype != "standard":
        raise KaldiFormatError(f"arc type {arctype!r} is not supported (only StdArc)")
    if flags & _FST_HAS_ISYMBOLS:
        p = _skip_symbol_table(d, p)
    if flags & _FST_HAS_OSYMBOLS:
        p = _skip_symbol_table(d, p)
    arc_dt = np.dtype([("ilabel", "<i4"), ("olabel", "<i4"), ("weight", "<f4"), ("nextstate", "<i4")])
    if fsttype == "vector":
        finals, offs, chunks = [], [0], []
        s = 0
        while (nstates < 0 or s < nstates) and p + 12 <= len(d):
            fw = struct.unpack_from("<f", d, p)[0]
            na = struct.unpack_from("<q", d, p + 4)[0]
            p += 12
            if na < 0 or na * 16 > len(d) - p:
                raise KaldiFormatError("vector FST: implausible arc count")
            chunks.append(np.frombuffer(d, arc_dt, na, p))
            p += na * 16
            finals.append(fw)
            offs.append(offs[-1] + na)
            s += 1
        arcs = np.concatenate(chunks) if chunks else np.zeros(0, arc_dt)
        final = np.array(finals, np.float32)
        offsets = np.array(offs, np.int64)
    elif fsttype == "const":
        if (flags & _FST_IS_ALIGNED) or version == 1:
            p = (p + 15) // 16 * 16
        st_dt = np.dtype([("final", "<f4"), ("pos", "<u4"), ("narcs", "<u4"), ("nieps", "<u4"), ("noeps", "<u4")])
        if nstates < 0 or narcs < 0 or nstates * st_dt.itemsize > len(d) - p:
            raise KaldiFormatError("const FST: implausible sizes")
        st = np.frombuffer(d, st_dt, nstates, p)
        p += nstates * st_dt.itemsize
        if (flags & _FST_IS_ALIGNED) or version == 1:
            p = (p + 15) // 16 * 16
        if narcs * 16 > len(d) - p:
            raise KaldiFormatError("const FST: the arc table is cut short")
        arcs = np.frombuffer(d, arc_dt, narcs, p)
        final = st["final"].astype(np.float32)
        offsets = np.concatenate([st["pos"].astype(np.int64), [narcs]])
        if nstates and not np.array_equal(np.diff(offsets), st["narcs"].astype(np.int64)):
            raise KaldiFormatError("const FST: state table is not contiguous")
    else:
        raise KaldiFormatError(f"FST type {fsttype!r} is not supported (vector, const)")
    if offsets[-1] > np.iinfo(np.int32).max:
        raise KaldiFormatError("more than 2^31 arcs")
    return dict(num_states=int(len(final)), start=int(start), offsets=offsets.astype(np.int32),
                ilabel=np.ascontiguousarray(arcs["ilabel"]), olabel=np.ascontiguousarray(arcs["olabel"]),
                weight=np.ascontiguousarray(arcs["weight"]), nextstate=np.ascontiguousarray(arcs["nextstate"]),
                final=final, fst_type=fsttype)


def write_openfst(path: str, g: dict, fst_type: str = "const", aligned: bool = False) -> None:
    """Writer for the same layout (used by the self-consistency test and to hand graphs to OpenFst tools)."""
    S, A = int(g["num_states"]), int(g["offsets"][-1])
    arc_dt = np.dtype([("ilabel", "<i4"), ("olabel", "<i4"), ("weight", "<f4"), ("nextstate", "<i4")])
    arcs = np.zeros(A, arc_dt)
    for k in ("ilabel", "olabel", "weight", "nextstate"):
        arcs[k] = g[k][:A]
    fin = np.asarray(g["final"], np.float32)
    off = np.asarray(g["offsets"], np.int64)

    def fstr(x):
        return struct.pack("<i", len(x)) + x.encode("ascii")
    version = 2 if not (aligned and fst_type == "const") else 1
    flags = _FST_IS_ALIGNED if (aligned and fst_type == "const") else 0
    hdr = struct.pack("<i", _FST_MAGIC) + fstr(fst_type) + fstr("standard") + struct.pack("<ii", version, flags) + \
        struct.pack("<Qqqq", 0, int(g["start"]), S, A)
    with open(path, "wb") as f:
        f.write(hdr)
        if fst_type == "vector":
            for s_ in range(S):
                f.write(struct.pack("<f", float(fin[s_])) + struct.pack("<q", int(off[s_ + 1] - off[s_])))
                f.write(arcs[off[s_]:off[s_ + 1]].tobytes())
        else:
            pos = len(hdr)
            if aligned:
                f.write(b"\0" * ((-pos) % 16))
                pos += (-pos) % 16
            st_dt = np.dtype([("final", "<f4"), ("pos", "<u4"), ("narcs", "<u4"), ("nieps", "<u4"), ("noeps", "<u4")])
            st = np.zeros(S, st_dt)
            st["final"], st["pos"], st["narcs"] = fin, off[:-1], np.diff(off)
            ie = np.add.reduceat((arcs["ilabel"] == 0).astype(np.int64), off[:-1].clip(max=max(A - 1, 0))) if A else np.zeros(S, np.int64)
            oe = np.add.reduceat((arcs["olabel"] == 0).astype(np.int64), off[:-1].clip(max=max(A - 1, 0))) if A else np.zeros(S, np.int64)
            empty = np.diff(off) == 0
            st["nieps"], st["noeps"] = np.where(empty, 0, ie), np.where(empty, 0, oe)
            f.write(st.tobytes())
            pos += S * st_dt.itemsize
            if aligned:
                f.write(b"\0" * ((-pos) % 16))
            f.write(arcs.tobytes())
