"""Host-side mirror of the reference's nnet3 inference surface over the b2k
C-ABI: the pair (DecodableNnetSimpleLoopedInfo, NnetComputer) of
nnet3/decodable-simple-looped.h:102-160 / nnet-compute.h:95-200 and the batched
wrapper cuda_decoder::BatchedStaticNnet3 (cudadecoder/batched-static-nnet3.h)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from . import nnet_model as NM


class _Node(C.Structure):
    _fields_ = [("dim", C.c_int32), ("rows", C.c_int32), ("kind", C.c_int32), ("arena_off", C.c_int64)]


class _Term(C.Structure):
    _fields_ = [("src", C.c_int32), ("ratio", C.c_int32), ("shift", C.c_int32), ("lo", C.c_int32),
                ("hi", C.c_int32), ("ivec", C.c_int32), ("C", C.c_int32), ("m", C.c_int32),
                ("k0", C.c_int32), ("klen", C.c_int32), ("scale", C.c_float), ("block", C.c_int32),
                ("col_step", C.c_int32), ("col_off", C.c_int32), ("col_lim", C.c_int32)]


class _Op(C.Structure):
    _fields_ = [("type", C.c_int32), ("out", C.c_int32), ("rows", C.c_int32), ("N", C.c_int32), ("K", C.c_int32),
                ("n_terms", C.c_int32), ("terms", _Term * 12),
                ("w", C.c_int64), ("bias", C.c_int64), ("bn_scale", C.c_int64), ("bn_offset", C.c_int64),
                ("sub_vec", C.c_int64), ("relu", C.c_int32), ("has_res", C.c_int32), ("res", _Term),
                ("res_alpha", C.c_float), ("out_scale", C.c_float), ("log_softmax", C.c_int32),
                ("block_dim", C.c_int32), ("hsplit", C.c_int32)]


def _term(d: dict, k0=0, klen=0, scale=1.0, block=0) -> _Term:
    return _Term(d["src"], d["ratio"], d["shift"], d["lo"], d["hi"], d.get("ivec", 0), d.get("C", 1), d.get("m", 0),
                 d.get("k0", k0), d.get("klen", klen), d.get("scale", scale), block,
                 d.get("col_step", 0), d.get("col_off", 0), d.get("col_lim", 0))


class NnetComputer:
    """Compiled, batched forward of one TDNN-F model for utterances of a fixed
    number of feature frames (the reference compiles one NnetComputation per
    request shape as well, nnet3/nnet-compile-looped.cc:329)."""

    def __init__(self, arch: dict, W: dict, num_frames: int, max_batch: int, frames_per_chunk: int = 21,
                 acoustic_scale: float = 1.0, use_priors: bool = True, conv_mode: str | None = None):
        self.prog = prog = NM.compile_program(arch, W, num_frames, frames_per_chunk, acoustic_scale, use_priors,
                                              conv_mode=conv_mode)
        nodes = (_Node * len(prog["nodes"]))()
        for i, (name, dim, rows, t0, step) in enumerate(prog["nodes"]):
            kind = {"input": 1, "ivector": 2, "output": 3}.get(name, 0)
            nodes[i] = _Node(dim, rows, kind, prog["arena_off"].get(i, 0))
        ops = (_Op * len(prog["ops"]))()
        for i, o in enumerate(prog["ops"]):
            op = _Op()
            op.out, op.rows = o["out"], o["rows"]
            op.w = op.bias = op.sub_vec = -1
            op.bn_scale, op.bn_offset = o.get("bn_scale", -1), o.get("bn_offset", -1)
            op.out_scale, op.block_dim = 1.0, 1
            if o["type"] == "gemm":
                op.type, op.N, op.K = 0, o["N"], o["K"]
                op.n_terms = len(o["terms"])
                assert op.n_terms <= 12
                op.hsplit = o.get("hsplit", 1)
                for j, t in enumerate(o["terms"]):
                    op.terms[j] = _term(t)
                op.w, op.bias, op.sub_vec = o["w"], o["bias"], o.get("sub_vec", -1)
                op.relu, op.log_softmax, op.out_scale = o["relu"], o["log_softmax"], o["out_scale"]
                if o.get("res"):
                    op.has_res, op.res, op.res_alpha = 1, _term(o["res"]), o["res_alpha"]
            else:
                op.type, op.block_dim = 1, o["block_dim"]
                flat = [(b, t) for b, blk in enumerate(o["blocks"]) for t in blk]
                op.n_terms = len(flat)
                assert op.n_terms <= 12
                for j, (b, t) in enumerate(flat):
                    op.terms[j] = _term(t, block=b)
                op.N = op.K = 0
            ops[i] = op
        blob = np.ascontiguousarray(prog["blob"], np.float32)
        self.h = C.c_void_p()
        L = _lib.lib()
        _lib.check(L.b2k_nnet_create(C.cast(nodes, C.c_void_p), len(nodes), C.cast(ops, C.c_void_p), len(ops),
                                     blob.ctypes.data_as(C.POINTER(C.c_float)), blob.size, int(max_batch),
                                     C.byref(self.h)))
        self.max_batch = max_batch
        self.num_frames = num_frames
        self.n_out = prog["n_out"]
        self.n_chunks = prog["n_chunks"]
        self.output_dim = arch["num_pdfs"]
        self.feat_dim = arch["feat_dim"]
        self.ivector_dim = arch["ivector_dim"]
        self.flops_per_utt = float(L.b2k_nnet_flops_per_lane(self.h))
        self.launches_per_run = int(L.b2k_nnet_num_launches_per_run(self.h))

    @classmethod
    def from_model(cls, model, num_frames: int, max_batch: int, frames_per_chunk: int = 21,
                   acoustic_scale: float = 1.0, use_priors: bool = True, conv_mode: str | None = None):
        """The C++ route: model.KaldiModel (b2k_model_read) → b2k_nnet_compile → b2k_nnet_create_from_program; the
        Python compiler is not involved (it is this route's test oracle)."""
        L = _lib.lib()
        self = cls.__new__(cls)
        self.h = C.c_void_p()
        self.prog = None
        prog = model.compile(num_frames, frames_per_chunk, acoustic_scale, use_priors, conv_mode)
        try:
            info = (C.c_int64 * 8)()
            L.b2k_nnet_program_info.argtypes = [C.c_void_p, C.c_void_p]
            _lib.check(L.b2k_nnet_program_info(prog, info))
            L.b2k_nnet_create_from_program.argtypes = [C.c_void_p, C.c_int32, C.c_void_p]
            _lib.check(L.b2k_nnet_create_from_program(prog, int(max_batch), C.byref(self.h)))
        finally:
            L.b2k_nnet_program_destroy.argtypes = [C.c_void_p]
            L.b2k_nnet_program_destroy(prog)
        self.max_batch = max_batch
        self.num_frames = num_frames
        self.n_out, self.n_chunks = int(info[0]), int(info[1])
        self.output_dim = model.num_pdfs
        self.feat_dim = model.feat_dim
        self.ivector_dim = model.ivector_dim
        self.flops_per_utt = float(L.b2k_nnet_flops_per_lane(self.h))
        self.launches_per_run = int(L.b2k_nnet_num_launches_per_run(self.h))
        return self

    def __del__(self):
        try:
            if self.h:
                _lib.lib().b2k_nnet_destroy(self.h)
        except Exception:
            pass

    def Run(self, input_ptrs, in_stride, ivector_ptrs, iv_stride, output_ptrs, out_stride, stream: int = 0):
        n = len(input_ptrs)

        def arr(ptrs):
            a = (C.c_void_p * n)(*[int(p) for p in ptrs])
            return C.cast(a, C.c_void_p), a
        ip, k1 = arr(input_ptrs)
        vp, k2 = arr(ivector_ptrs) if ivector_ptrs is not None else (None, None)
        op, k3 = arr(output_ptrs)
        _lib.check(_lib.lib().b2k_nnet_run(self.h, n, ip, int(in_stride), vp, int(iv_stride), op, int(out_stride),
                                           C.c_void_p(stream)))

    # test convenience
    def forward(self, feats_list, chunk_ivectors_list):
        import torch
        d_in = [torch.from_numpy(np.ascontiguousarray(f, np.float32)).cuda() for f in feats_list]
        d_iv = [torch.from_numpy(np.ascontiguousarray(v, np.float32)).cuda() for v in chunk_ivectors_list]
        d_out = [torch.zeros(self.n_out, self.output_dim, device="cuda") for _ in feats_list]
        self.Run([x.data_ptr() for x in d_in], self.feat_dim, [x.data_ptr() for x in d_iv], self.ivector_dim,
                 [x.data_ptr() for x in d_out], self.output_dim)
        torch.cuda.synchronize()
        return [o.cpu().numpy() for o in d_out]


class BatchedStaticNnet3:
    """cuda_decoder::BatchedStaticNnet3 (cudadecoder/batched-static-nnet3.h:59-138) over b2k_nnet_stream_*: at most one
    chunk of features per channel and call, context frames carried per channel on the device, right context flushed
    on the last chunk.  `frames_per_chunk` is compute_opts.frames_per_chunk (input frames)."""

    def __init__(self, arch: dict, W: dict, max_batch: int, nchannels: int = -1, frames_per_chunk: int = 51,
                 acoustic_scale: float = 1.0, use_priors: bool = True, looped: bool = False):
        from .nnet_compile import _Cfg, abi_arrays
        L = _lib.lib()
        layers, ws, self._keep = abi_arrays(arch, W)
        cfg = _Cfg(arch["feat_dim"], arch["ivector_dim"], arch["num_pdfs"], arch["frame_subsampling_factor"], 0,
                   int(frames_per_chunk), int(use_priors), 0, float(acoustic_scale))
        self.h = C.c_void_p()
        L.b2k_nnet_stream_create.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                             C.c_void_p]
        _lib.check(L.b2k_nnet_stream_create(C.byref(cfg), layers, len(layers), ws, len(ws), int(max_batch), int(nchannels),
                                            int(bool(looped)), C.byref(self.h)))
        info = (C.c_int64 * 8)()
        L.b2k_nnet_stream_info.argtypes = [C.c_void_p, C.c_void_p]
        _lib.check(L.b2k_nnet_stream_info(self.h, info))
        (self.output_frames_per_chunk, self.left_context, self.right_context, self.window, self.input_dim,
         self.ivector_dim, self.output_dim, self.ivector_rows) = [int(x) for x in info]
        self.max_batch = max_batch
        self.frames_per_chunk = frames_per_chunk
        L.b2k_nnet_stream_run_batch.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p,
                                                C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p,
                                                C.c_void_p]

    def __del__(self):
        try:
            if self.h:
                _lib.lib().b2k_nnet_stream_destroy.argtypes = [C.c_void_p]
                _lib.lib().b2k_nnet_stream_destroy(self.h)
        except Exception:
            pass

    def GetNOutputFramesPerChunk(self) -> int:
        return self.output_frames_per_chunk

    def GetTotalNnet3RightContext(self) -> int:
        return self.right_context

    def RunBatch(self, channels, d_features, features_stride, d_ivectors, n_input_frames_valid, is_first_chunk,
                 is_last_chunk, d_all_log_posteriors: int, d_all_eos_log_posteriors: int, out_stride: int, stream: int = 0):
        """Device pointers as integers.  Returns (n_output_frames, n_eos_output_frames) per batch slot: slot i's frame k
        is row i*output_frames_per_chunk + k of d_all_log_posteriors, its flush frames the same rows of
        d_all_eos_log_posteriors (FormatOutputPtrs, batched-static-nnet3.cc:369-395)."""
        n = len(channels)
        ch = (C.c_int32 * n)(*[int(c) for c in channels])
        fp = (C.c_void_p * n)(*[int(p) if p else None for p in d_features])
        ip = (C.c_void_p * n)(*[int(p) for p in d_ivectors]) if d_ivectors is not None else None
        nv = (C.c_int32 * n)(*[int(x) for x in n_input_frames_valid])
        fi = (C.c_int32 * n)(*[int(bool(x)) for x in is_first_chunk])
        la = (C.c_int32 * n)(*[int(bool(x)) for x in is_last_chunk])
        no, ne = (C.c_int32 * n)(), (C.c_int32 * n)()
        _lib.check(_lib.lib().b2k_nnet_stream_run_batch(self.h, n, ch, fp, int(features_stride), ip, nv, fi, la,
                                                        C.c_void_p(d_all_log_posteriors),
                                                        C.c_void_p(d_all_eos_log_posteriors) if d_all_eos_log_posteriors else None,
                                                        int(out_stride), no, ne, C.c_void_p(stream)))
        return list(no), list(ne)
