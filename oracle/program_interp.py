"""TEST INFRASTRUCTURE ONLY: a numpy interpreter of the op program that
kaldi_b200.nnet_model.compile_program emits and kaldi_b200/csrc/nnet.cu executes
(same row maps, same epilogue order).  It lets the tests check a compiled program
against the reference's compiled nnet3 forward (oracle/_ref) on the CPU, so that
a new layer type (e.g. the dense-expanded TimeHeightConvolutionComponent) is
validated before any GPU time is spent; the GPU tests then only have to show
that the kernels execute the same program.
"""
from __future__ import annotations

import numpy as np


def _rows(term: dict, n_rows: int) -> np.ndarray:
    i = np.arange(n_rows, dtype=np.int64)
    j = i * term["ratio"] + term["shift"]
    if term.get("ivec"):
        j = np.floor_divide(j, term["C"]) - term["m"]          # floor division, also for negative times
    return np.clip(j, term["lo"], term["hi"])


def run_program(prog: dict, feats: np.ndarray, chunk_ivectors: np.ndarray | None) -> np.ndarray:
    blob = prog["blob"]
    nodes = prog["nodes"]                                       # (name, dim, rows, t0, step)
    val = {}
    for i, (name, dim, rows, _t0, _step) in enumerate(nodes):
        if name == "input":
            val[i] = np.ascontiguousarray(feats, np.float32)
        elif name == "ivector":
            val[i] = (np.zeros((prog["n_chunks"], dim), np.float32) if chunk_ivectors is None
                      else np.ascontiguousarray(chunk_ivectors, np.float32))
    out = None
    for op in prog["ops"]:
        name, dim, _r, _t0, _step = nodes[op["out"]]
        R = op["rows"]
        if op["type"] == "gemm":
            N, K = op["N"], op["K"]
            w = blob[op["w"]:op["w"] + N * K].reshape(N, K)
            H = op.get("hsplit", 1)
            acc = np.zeros((R, H, N), np.float32)
            for t in op["terms"]:
                rows = val[t["src"]][_rows(t, R)]
                wt = w[:, t["k0"]:t["k0"] + t["klen"]].T
                for h in range(H):
                    if t.get("col_lim", 0) > 0:                 # convolution patch window of output height h
                        cb = h * t["col_step"] + t["col_off"]
                        if not (0 <= cb < t["col_lim"]):
                            continue                            # height zero padding
                        acc[:, h] += rows[:, cb:cb + t["klen"]] @ wt
                    else:
                        acc[:, h] += rows[:, :t["klen"]] @ wt
            v = acc.reshape(R * H, N)
            if op["bias"] >= 0:
                v = v + blob[op["bias"]:op["bias"] + N]
            if op["relu"]:
                v = np.maximum(v, 0.0)
            if op["bn_scale"] >= 0:
                v = v * blob[op["bn_scale"]:op["bn_scale"] + N] + blob[op["bn_offset"]:op["bn_offset"] + N]
            if op.get("res"):
                assert H == 1
                v = np.float32(op["res_alpha"]) * val[op["res"]["src"]][_rows(op["res"], R)] + v
            if op["log_softmax"]:
                m = v.max(axis=1, keepdims=True)
                v = v - (m + np.log(np.exp(v - m).sum(axis=1, keepdims=True)))
            if op["sub_vec"] >= 0:
                v = v - blob[op["sub_vec"]:op["sub_vec"] + N]
            if op["out_scale"] != 1.0:
                v = v * np.float32(op["out_scale"])
            v = v.astype(np.float32).reshape(R, H * N)
        else:
            bd = op["block_dim"]
            v = np.zeros((R, dim), np.float32)
            for bi, blk in enumerate(op["blocks"]):
                for t in blk:
                    v[:, bi * bd:(bi + 1) * bd] += np.float32(t["scale"]) * val[t["src"]][_rows(t, R)][:, :bd]
            if op["bn_scale"] >= 0:
                v = v * blob[op["bn_scale"]:op["bn_scale"] + dim] + blob[op["bn_offset"]:op["bn_offset"] + dim]
            v = v.astype(np.float32)
        val[op["out"]] = v
        if name == "output":
            out = v
    return out


def program_from_abi(nodes, ops, blob) -> dict:
    """The same program, rebuilt from the C-ABI structs (b2k_nnet_node / b2k_nnet_op arrays as b2k_nnet_create
    takes them, include/b2k.h): lets the tests interpret exactly what the device executes, including programs
    that never existed as a Python dictionary (the C++ compiler's output)."""
    kinds = {1: "input", 2: "ivector", 3: "output"}

    def term(t):
        return dict(src=t.src, ratio=t.ratio, shift=t.shift, lo=t.lo, hi=t.hi, ivec=t.ivec, C=t.C, m=t.m, k0=t.k0,
                    klen=t.klen, scale=t.scale, col_step=t.col_step, col_off=t.col_off, col_lim=t.col_lim)
    pn = [(kinds.get(n.kind, "n%d" % i), n.dim, n.rows, 0, 0) for i, n in enumerate(nodes)]
    po = []
    for o in ops:
        d = dict(out=o.out, rows=o.rows, bn_scale=o.bn_scale, bn_offset=o.bn_offset)
        if o.type == 0:
            d.update(type="gemm", N=o.N, K=o.K, hsplit=max(o.hsplit, 1), terms=[term(o.terms[j]) for j in range(o.n_terms)],
                     w=o.w, bias=o.bias, sub_vec=o.sub_vec, relu=o.relu, log_softmax=o.log_softmax, out_scale=o.out_scale)
            if o.has_res:
                d.update(res=term(o.res), res_alpha=o.res_alpha)
        else:
            nb = 1 + max(o.terms[j].block for j in range(o.n_terms))
            d.update(type="sum", block_dim=o.block_dim,
                     blocks=[[term(o.terms[j]) for j in range(o.n_terms) if o.terms[j].block == b] for b in range(nb)])
        po.append(d)
    n_chunks = max([n.rows for n in nodes if n.kind == 2] + [0])
    return dict(nodes=pn, ops=po, blob=np.asarray(blob, np.float32), n_chunks=n_chunks)
