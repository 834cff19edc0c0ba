#!/usr/bin/env python
"""bench.py — headline benchmark of the online2 inference hot path on B200.

Metric (BASELINE.json): real-time factor RTFx = audio-seconds / wall-seconds.  A "step" is one batch of synthetic
16 kHz 10 s utterances through features -> online i-vectors -> nnet3 -> lattice decoder -> finalized raw lattices.

  python bench.py --gpus N --steps K --warmup W            # our arm; default workload = BASELINE configs[1]
  python bench.py --impl reference --gpus N ...             # the reference's own CPU path on the host cores
  python bench.py --workload librispeech_tdnn_1d/hclg50M/batch512   # BASELINE configs[2] (one GPU's share)
  python bench.py --workload decoder_sweep                  # BASELINE configs[4]: Marcs/s per (graph size, beam)
  python bench.py --workload streaming [--partials]          # serving mode: B streams chunk by chunk through b2k_stream_*

Everything on the GPU side goes through the C ABI (include/b2k.h): b2k_pipeline_* over libb2k.so.
`value`       : whole-job RTFx with the audio already resident in HBM (b2k_pipeline_run_device, CUDA events).
`e2e`         : the same through b2k_pipeline_submit_i16 / b2k_pipeline_collect with HOST int16 buffers: staging to
                pinned memory, H2D, all stages, packed lattices D2H and unpacking are inside the timed region; batch
                k+1 is submitted before batch k is collected (the calls a serving loop makes).
`roofline`    : the dominant kernel (largest share of the step, measured live with CUDA events).
`cpu_baseline`: the reference CPU path (oracle/_ref: the reference's own sources) on a bounded sample, one thread.
`parity_checked`: utterances of the last end-to-end batch whose lattices were compared, outside the timed region, with
                the compiled reference decoder run on the device's log-likelihoods (bit-identical or the bench fails).
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from kaldi_b200 import nnet_model as NM  # noqa: E402
from kaldi_b200 import synth  # noqa: E402

NUM_SAMPLES = 160000

WORKLOADS = {
    # BASELINE configs[1]: the configuration the metric is quoted on at N = 1
    "mini_librispeech_tdnn_1k/10s-utts/beam15/hclg5M": dict(
        arch="mini_librispeech_1k", num_pdfs=2336, graph_arcs=5_000_000, batch=592, cal="bench_calibration.npz",
        model="mini_librispeech tdnn_1k (4.47 M params fwd path, 2336 pdfs)"),
    # BASELINE configs[2]: one GPU's share of the 8-GPU target configuration
    "librispeech_tdnn_1d/hclg50M/batch512": dict(
        arch="librispeech_1d", num_pdfs=6024, graph_arcs=50_000_000, batch=512, cal="bench_calibration_1d.npz",
        model="librispeech tdnn_1d (1536/160 x 17 layers, 6024 pdfs)"),
}
DEFAULT_WORKLOAD = "mini_librispeech_tdnn_1k/10s-utts/beam15/hclg5M"


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm_gbs=float(d["hbm_gbs"]), bf16_tflops=float(d.get("bf16_tflops_sustained", d.get("bf16_tflops", 0))),
                    source="measured (MEASURED_PEAKS.json)")
    return dict(hbm_gbs=6650.0, bf16_tflops=1590.0, source="fallback (B200_PROFILING.md)")


def usable_cores():
    """Host threads this process can really use: the affinity mask AND the cgroup CPU quota (a 1-GPU lease of this pool
    shows 128 CPUs in the mask with cpu.max = 16 CPUs: round 1 started 128 workers on it, profiles/r02_reference_arm.md)."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    quota = None
    try:
        q, per = open("/sys/fs/cgroup/cpu.max").read().split()[:2]
        if q != "max":
            quota = max(1, int(int(q) / int(per)))
    except Exception:
        try:
            q = int(open("/sys/fs/cgroup/cpu/cpu.cfs_quota_us").read())
            per = int(open("/sys/fs/cgroup/cpu/cpu.cfs_period_us").read())
            if q > 0:
                quota = max(1, q // per)
        except Exception:
            pass
    return dict(affinity=n, cgroup_quota=quota, usable=min(n, quota) if quota else n)


def graph_for(arcs: int, num_pdfs: int):
    """Synthetic HCLG (kaldi_b200/synth.py); the 50 M-arc one takes minutes to draw and is cached for the other commands of
    the same GPU lease (under /tmp: gpurun_out/ is copied back and capped at 64 MiB)."""
    cache = os.path.join(os.environ.get("B2K_CACHE_DIR", "/tmp/b2k_cache"), f"hclg_{arcs}_{num_pdfs}.npz")
    if arcs >= 20_000_000 and os.path.exists(cache):
        z = np.load(cache)
        g = {k: z[k] for k in z.files}
        for k in ("num_states", "start", "num_pdfs"):
            g[k] = int(g[k])
        return g
    g = synth.make_hclg(arcs, num_pdfs=num_pdfs, seed=1)
    if arcs >= 20_000_000:
        try:
            os.makedirs(os.path.dirname(cache), exist_ok=True)
            np.savez(cache, **g)
        except Exception:
            pass
    return g


def arch_for(w):
    return getattr(NM, "arch_" + w["arch"])(w["num_pdfs"])


class ClockSampler:
    """nvidia-smi clocks/throttle reasons DURING the timed region (B200_PROFILING.md)."""

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        self.proc.terminate()
        try:
            self.t.join(timeout=2)
        except Exception:
            pass
        sm = [float(r[0]) for r in self.rows if len(r) >= 6 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 6 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 6 for i in range(4) if r[2 + i].lower().startswith("active")})
        return dict(sm_mhz=float(np.median(sm)) if sm else None, sm_max_mhz=max(mx) if mx else None, reasons=reasons,
                    samples=len(sm))


# ----------------------------------------------------------------------------- CPU reference path

_CPU_STATE = {}


def load_calibrated_weights(arch, seed, cal_name):
    """Synthetic weights + the prior-style output calibration (committed fixture: both arms decode the same model)."""
    W = NM.random_weights(arch, seed=seed)
    for p in (os.path.join(ROOT, "gpurun_out", cal_name), os.path.join(ROOT, "tests", "golden", cal_name)):
        if os.path.exists(p):
            c = np.load(p)
            return NM.apply_output_calibration(W, c["mean"], float(c["scale"]))
    return W


def _cpu_setup(wname):
    """Graph, model, extractor, features and decoder of the reference's CPU path (oracle/_ref), once per process."""
    if _CPU_STATE.get("workload") == wname:
        return
    os.environ.setdefault("OPENBLAS_NUM_THREADS", "1")
    from oracle import feat_oracle as F, nnet_oracle as NO, ref_decoder as RD, ivector_oracle as IV
    w = WORKLOADS[wname]
    g = _CPU_STATE.get("graph")
    if g is None:
        g = graph_for(w["graph_arcs"], w["num_pdfs"])
    arch = arch_for(w)
    W = load_calibrated_weights(arch, 0, w["cal"])
    _CPU_STATE.update(workload=wname, graph=g, nnet=NO.RefNnet(arch, W), feat=F.RefFeat(), F=F,
                      dec=RD.RefDecoder(g, synth.DEFAULT_DECODER_CFG), ivx=IV.make_cpu_extractor(0))


def cpu_reference_one(args):
    """One utterance through the reference's CPU path, all of it the reference's own sources compiled in oracle/_ref:
    features (online, 0.18 s chunks), i-vector, nnet3 looped forward and LatticeFasterDecoder."""
    idx, wname = args
    _cpu_setup(wname)
    S = _CPU_STATE
    F = S["F"]
    wave = S["waves"][idx]                                          # drawn before the clock started (fork-shared)
    t1 = time.time()
    feats = S["feat"].compute(wave, F.FeatOpts(), online_chunk=2880)
    R = S["nnet"]
    n_chunks = len(R.chunk_ivector_rows(feats.shape[0], feats.shape[0], 1))
    civ = S["ivx"].chunk_ivectors(feats, n_chunks, R.frames_per_chunk, R.right_context) if S["ivx"] is not None \
        else np.zeros((n_chunks, 100), np.float32)
    ends = [(n + 1) * R.frames_per_chunk + R.right_context for n in range(n_chunks)]
    mat = np.zeros((ends[-1] + 1, 100), np.float32)
    prev = 0
    for n, e in enumerate(ends):
        mat[prev:e + 1] = civ[n]
        prev = e + 1
    ll = R.forward(feats, mat, period=1)
    S["dec"].decode(ll, record=S.get("record", False))              # record: tools/bench_lattice_det.py
    t2 = time.time()
    ns, na, _ = S["dec"].lattice_sizes()
    return (t2 - t1, ns, na)


def run_reference_arm(a, wname):
    """`--impl reference`: the reference's CPU path on every host thread this process may use; utterances are drawn
    before the pool forks, each step is `2 x workers` utterances."""
    import multiprocessing as mp
    cores = usable_cores()
    workers = cores["usable"]
    per_step = 2 * workers
    n_steps = a.warmup + a.steps
    os.environ["OPENBLAS_NUM_THREADS"] = "1"
    os.environ["OMP_NUM_THREADS"] = "1"
    # one step's worth of distinct utterances (+ one per worker for the warm-up), reused by every step: the CPU path keeps no
    # state between utterances, and drawing 25 steps x 2 x workers utterances would cost minutes on a many-core box
    _CPU_STATE["waves"] = [synth.make_audio(NUM_SAMPLES, seed=1000 + i) for i in range(per_step + workers)]
    _cpu_setup(wname)                                               # before the fork: workers share graph and model
    pool = mp.get_context("fork").Pool(workers)
    base = per_step
    pool.map(cpu_reference_one, [(base + i, wname) for i in range(workers)], chunksize=1)   # warm every worker
    vals, t_all = [], 0.0
    for s in range(n_steps):
        t0 = time.time()
        pool.map(cpu_reference_one, [(i, wname) for i in range(per_step)], chunksize=1)
        dt = time.time() - t0
        if s >= a.warmup:
            vals.append(per_step * NUM_SAMPLES / 16000.0 / dt)
            t_all += dt
    pool.close()
    v = float(np.mean(vals))
    sample = (f"{per_step} x 10 s utterances per step on {workers} worker processes (affinity {cores['affinity']}, cgroup quota "
              f"{cores['cgroup_quota']}), audio drawn before the clock; features + i-vector + nnet3 + LatticeFasterDecoder are the "
              "reference's own sources compiled in oracle/_ref (OpenBLAS 1 thread per worker; decoder against a container-only "
              "OpenFst stand-in); lattice determinization excluded on both arms")
    w = WORKLOADS[wname]
    return dict(metric="real-time factor (audio-sec/wall-sec)", value=v, unit="RTFx", n_gpus=a.gpus, steps=a.steps,
                warmup=a.warmup, ms_per_step=1e3 * t_all / max(a.steps, 1), higher_is_better=True, scaling="weak",
                vs_baseline=None, dtype="f32", data="synthetic", impl="reference",
                config=dict(workload=wname, model=w["model"], graph_arcs=w["graph_arcs"], beam=15.0, max_active=7000,
                            lattice_beam=8.0, sample=f"{per_step} utterances of 10 s per step"),
                cpu_baseline=dict(value=v, unit="RTFx", cores=workers, kind="reference", sample=sample),
                e2e=dict(value=v, unit="RTFx", h2d_bytes_per_step=0, d2h_bytes_per_step=0))


def run_cpu_baseline(num_utts, wname):
    _CPU_STATE["waves"] = [synth.make_audio(NUM_SAMPLES, seed=900 + i) for i in range(num_utts + 1)]
    _cpu_setup(wname)
    cpu_reference_one((num_utts, wname))                            # warm (untimed)
    res = [cpu_reference_one((i, wname)) for i in range(num_utts)]
    t = sum(r[0] for r in res)
    return dict(value=num_utts * NUM_SAMPLES / 16000.0 / t, unit="RTFx", cores=1, kind="reference",
                sample=f"{num_utts} x 10 s utterances, single thread: features + i-vector + nnet3 + LatticeFasterDecoder are the "
                       "reference's own sources compiled in oracle/_ref (OpenBLAS 1 thread; decoder against a container-only "
                       "OpenFst stand-in); determinization excluded")


# ----------------------------------------------------------------------------- GPU arm

def dec_infos(L, dec_handle, n):
    L.b2k_dec_channel_info.argtypes = [C.c_void_p, C.c_int32, C.c_void_p]
    out = []
    info = (C.c_int64 * 32)()
    for c in range(n):
        L.b2k_dec_channel_info(dec_handle, c, info)
        out.append([int(x) for x in info])
    return out


def decoder_sweep(a, rank, world, local_rank):
    """BASELINE configs[4]: decoder only, fixed precomputed log-likelihoods, graph 5 M -> 200 M arcs, beam 10 -> 20."""
    import torch
    from kaldi_b200.decoder import CudaDecoder, CudaDecoderConfig, CudaFst
    torch.cuda.set_device(local_rank)
    peaks = load_peaks()
    B, T, P = a.batch or 592, 333, 2336
    sizes = [int(x) for x in (a.sweep_arcs or "5000000,20000000,50000000").split(",")]
    beams = [float(x) for x in (a.sweep_beams or "10,13,15,17,20").split(",")]
    rows = []
    for arcs in sizes:
        g = graph_for(arcs, P)
        fst = CudaFst(g)
        ll = torch.from_numpy(np.stack([synth.make_loglikes(g, T, seed=100 + i) for i in range(min(B, 16))])).cuda()
        ll = ll.repeat((B + ll.shape[0] - 1) // ll.shape[0], 1, 1)[:B].contiguous()
        for beam in beams:
            cfg = dict(synth.DEFAULT_DECODER_CFG, beam=beam)
            dc = CudaDecoderConfig.from_dict(cfg, max_frames=T + 2, max_tokens=T * 9000, max_links=T * 16000,
                                             reference_order=True, max_tokens_per_frame=a.max_tpf)
            dec = CudaDecoder(fst, dc, B)
            ch = list(range(B))
            lp = [ll.data_ptr() + 4 * T * P * i for i in range(B)]
            ts = []
            for it in range(1 + max(1, a.steps)):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                dec.InitDecoding(ch)
                e0.record()
                dec.AdvanceDecodingFrames(ch, lp, [T] * B, P)
                e1.record()
                torch.cuda.synchronize()
                if it > 0:
                    ts.append(e0.elapsed_time(e1) / 1e3)
            infos = [dec.ChannelInfo(c) for c in range(B)]
            arcs_x = sum(i["arcs_emitting"] + i["arcs_nonemitting"] for i in infos)
            ntok, nlink = sum(i["ntok"] for i in infos), sum(i["nlink"] for i in infos)
            t = float(np.mean(ts))
            alg = 16.0 * arcs_x + 32.0 * ntok + 36.0 * nlink
            rows.append(dict(graph_arcs=int(g["offsets"][-1]), beam=beam, marcs_per_s=arcs_x / t / 1e6, ms=1e3 * t,
                             tokens_per_frame=ntok / (B * T), arcs_per_frame=arcs_x / (B * T), gbs=alg / t / 1e9,
                             frac_of_hbm=alg / t / 1e9 / peaks["hbm_gbs"], errors=sum(1 for i in infos if i["status"] != 0)))
            del dec
        del fst
    if rank == 0:
        best = max(r["marcs_per_s"] for r in rows)
        print(json.dumps(dict(metric="decoder Marcs/s (arcs examined per second)", value=best, unit="Marcs/s", n_gpus=world,
                              steps=a.steps, warmup=1, ms_per_step=float(sum(r["ms"] for r in rows)), higher_is_better=True,
                              scaling="weak", vs_baseline=None, dtype="f32",
                              data="synthetic", config=dict(workload="decoder_sweep", lanes=B, frames=T, pdfs=P,
                                                            decoder_mode="reference_order"), sweep=rows)))
    return 0


def streaming_bench(a, rank, world, local_rank):
    """Serving mode: B audio streams decoded chunk by chunk through b2k_stream_* (the call structure of
    BatchedThreadedNnet3CudaOnlinePipeline::DecodeBatch): every call brings 0.51 s of 16-bit PCM per stream from HOST memory,
    computes the new feature frames, the chunked network and the decoder frames of all streams, and the last call finalizes
    them.  A step = all the calls of B utterances of 10 s.  The network sees zero i-vectors (the device i-vector stage works
    per utterance).  With --partials every call also asks for every stream's partial hypothesis (b2k_dec_best_path)."""
    import torch
    from kaldi_b200.decoder import CudaFst
    from kaldi_b200.feat import FeatureOptions
    from kaldi_b200.model import KaldiModel
    from kaldi_b200.streaming import NativeStreamingDecoder
    torch.cuda.set_device(local_rank)
    wname = DEFAULT_WORKLOAD
    w = WORKLOADS[wname]
    arch = arch_for(w)
    W = load_calibrated_weights(arch, 0, w["cal"])
    graph = graph_for(w["graph_arcs"], w["num_pdfs"])
    B, fpc = a.batch or w["batch"], 51
    nf_out = 333 + 24
    sd = NativeStreamingDecoder(KaldiModel.from_arch(arch, W), CudaFst(graph), dict(synth.DEFAULT_DECODER_CFG), nchannels=B,
                                max_seconds=NUM_SAMPLES / 16000.0 + 0.1, frames_per_chunk=fpc,
                                feature_opts=FeatureOptions(max_lanes=max(B, 64)),
                                decoder_kwargs=dict(max_tokens=nf_out * a.tok_per_frame, max_links=nf_out * a.links_per_frame,
                                                    max_tokens_per_frame=a.max_tpf, reference_order=True))
    pcm = np.stack([synth.make_audio(NUM_SAMPLES, seed=10_000 * rank + i) for i in range(B)]).astype(np.int16)
    chunk = fpc * 160
    n_calls = (NUM_SAMPLES + chunk - 1) // chunk
    ch = list(range(B))
    h2d = 0

    def one_utterance_set():
        nonlocal h2d
        res = None
        for k in range(n_calls):
            pieces = [pcm[i, k * chunk:(k + 1) * chunk] for i in range(B)]
            res = sd.DecodeBatch(ch, pieces, [k == 0] * B, [k == n_calls - 1] * B, want_partial=a.partials, lattices="batched")
            h2d += sum(len(p) for p in pieces) * 2
        return res
    for _ in range(max(1, a.warmup)):
        one_utterance_set()
    torch.cuda.synchronize()
    h2d = 0
    mon = ClockSampler(local_rank); mon.start()
    t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    frames = 0
    for _ in range(a.steps):
        res = one_utterance_set()
        frames = res[0]["frames_decoded"]
    e1.record()
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    clocks = mon.stop()
    t_dev = e0.elapsed_time(e1) / 1e3
    audio = B * NUM_SAMPLES / 16000.0 * a.steps
    infos = [sd._with_dec().ChannelInfo(c) for c in (0, B // 2, B - 1)]
    sd.dec.h = None
    if rank == 0:
        print(json.dumps(dict(metric="real-time factor (audio-sec/wall-sec)", value=audio / wall, unit="RTFx", n_gpus=world, steps=a.steps,
                              warmup=max(1, a.warmup), ms_per_step=1e3 * wall / a.steps, higher_is_better=True, scaling="weak",
                              vs_baseline=None, dtype="f32", data="synthetic",
                              config=dict(workload="streaming: " + wname, streams=B, seconds_per_stream=NUM_SAMPLES / 16000.0,
                                          chunk_seconds=chunk / 16000.0, calls_per_utterance=n_calls, frames_per_chunk=fpc,
                                          partial_hypotheses_every_call=bool(a.partials), ivectors="zeros",
                                          api="b2k_stream_decode_batch_i16 (C ABI), 16-bit PCM chunks in host memory every call",
                                          timing="wall clock around the calls (host-paced: every call copies its chunks from "
                                                 "pageable host memory); device time in e2e.device_s"),
                              e2e=dict(value=audio / wall, unit="RTFx", h2d_bytes_per_step=h2d // max(1, a.steps),
                                       d2h_bytes_per_step=int(sum(v.nbytes for v in res[0]["lattices_packed"].values())),
                                       lattice_states_per_step=int(len(res[0]["lattices_packed"]["state_frame"])),
                                       device_s=t_dev, wall_s=wall),
                              output_frames_per_stream=frames, channel_status=[i["status"] for i in infos], clocks=clocks)))
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD, choices=list(WORKLOADS) + ["decoder_sweep", "streaming"])
    ap.add_argument("--partials", action="store_true", help="--workload streaming: ask for every stream's partial hypothesis after every call")
    ap.add_argument("--batch", type=int, default=0, help="utterances per GPU per step (0 = the workload's own)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-utts", type=int, default=6)
    ap.add_argument("--parity-utts", type=int, default=2, help="utterances of the last e2e batch checked against the compiled reference decoder")
    ap.add_argument("--max-tpf", type=int, default=65536, help="decoder per-frame token capacity (hash = 2x slots)")
    ap.add_argument("--tok-per-frame", type=int, default=9000, help="decoder token arena sizing (avg tokens/frame)")
    ap.add_argument("--links-per-frame", type=int, default=16000, help="decoder link arena sizing (avg links/frame)")
    ap.add_argument("--sweep-arcs", default="")
    ap.add_argument("--sweep-beams", default="")
    a = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if a.impl == "reference":
        if rank != 0:
            return 0
        wname = a.workload if a.workload in WORKLOADS else DEFAULT_WORKLOAD
        print(json.dumps(run_reference_arm(a, wname)))
        return 0
    if a.workload == "decoder_sweep":
        return decoder_sweep(a, rank, world, local_rank)
    if a.workload == "streaming":
        return streaming_bench(a, rank, world, local_rank)

    import torch
    import torch.distributed as dist
    from kaldi_b200 import _lib
    from kaldi_b200.decoder import CudaDecoder, CudaFst, lattice_to_canonical
    from kaldi_b200.feat import FeatureOptions
    from kaldi_b200.ivector import IvectorExtractorGpu, make_synthetic_extractor
    from kaldi_b200.model import KaldiModel
    from kaldi_b200.pipeline import NativeBatchedPipeline, PipelineConfig

    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl")
    L = _lib.lib()
    peaks = load_peaks()
    wname = a.workload
    w = WORKLOADS[wname]
    P = w["num_pdfs"]
    arch = arch_for(w)
    if not any(os.path.exists(os.path.join(ROOT, d, w["cal"])) for d in ("gpurun_out", os.path.join("tests", "golden"))):
        # no calibration fixture for this synthetic model yet (nnet_model.apply_output_calibration): measure it once on 16
        # utterances and keep it (gpurun_out/ comes back from the GPU box; committed under tests/golden/ so that the
        # reference arm decodes the same model)
        from kaldi_b200.pipeline import BatchedPipeline
        W0 = NM.random_weights(arch, seed=0)
        cp = BatchedPipeline(PipelineConfig(max_batch=16, num_samples=NUM_SAMPLES, max_tokens=200_000, max_links=400_000),
                             arch, W0, synth.make_hclg(20_000, num_pdfs=P, seed=3), ivector_extractor=make_synthetic_extractor(seed=0))
        mean, scale = cp.output_calibration([synth.make_audio(NUM_SAMPLES, seed=777_000 + i) for i in range(16)], target_std=1.0)
        del cp
        torch.cuda.empty_cache()
        if rank == 0:
            os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
            np.savez(os.path.join(ROOT, "gpurun_out", w["cal"]), mean=mean, scale=np.float32(scale))
        if world > 1:
            dist.barrier()
    W = load_calibrated_weights(arch, 0, w["cal"])
    graph = graph_for(w["graph_arcs"], P)
    _CPU_STATE["graph"] = graph
    B = a.batch or w["batch"]
    nf_out = 333
    cfg = PipelineConfig(feature_opts=FeatureOptions(max_lanes=max(B, 64)), max_batch=B, num_samples=NUM_SAMPLES,
                         reference_order=True, max_tokens=nf_out * a.tok_per_frame, max_links=nf_out * a.links_per_frame,
                         max_tokens_per_frame=a.max_tpf)
    model = KaldiModel.from_arch(arch, W)
    fst = CudaFst(graph)
    T_feat = 1 + (NUM_SAMPLES - 400) // 160
    ivx = IvectorExtractorGpu(make_synthetic_extractor(seed=0), B, T_feat)
    pipe = NativeBatchedPipeline(cfg, model, fst, ivx)
    nf = pipe.plan.num_output_frames
    L.b2k_pipeline_enable_stage_timing.argtypes = [C.c_void_p, C.c_int32]
    L.b2k_pipeline_stage_times.argtypes = [C.c_void_p, C.c_void_p]
    L.b2k_pipeline_run_device.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]
    L.b2k_pipeline_nnet_flops_per_utterance.restype = C.c_double
    L.b2k_pipeline_nnet_flops_per_utterance.argtypes = [C.c_void_p]
    L.b2k_pipeline_decoder.restype = C.c_void_p
    L.b2k_pipeline_decoder.argtypes = [C.c_void_p]
    _lib.check(L.b2k_pipeline_enable_stage_timing(pipe.h, 1))

    # distinct utterances per rank and per step slot (cycled); int16 on the host, float on the device
    n_sets = 2
    host_sets = [np.stack([synth.make_audio(NUM_SAMPLES, seed=10_000 * rank + 1000 * s + i) for i in range(B)]).astype(np.int16)
                 for s in range(n_sets)]
    host_ptrs = [NativeBatchedPipeline.row_pointers(hs) for hs in host_sets]
    dev_sets = [torch.from_numpy(hs.astype(np.float32)).cuda() for hs in host_sets]
    stream = torch.cuda.current_stream().cuda_stream

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def step_device(s):
        _lib.check(L.b2k_pipeline_run_device(pipe.h, B, C.c_void_p(dev_sets[s % n_sets].data_ptr()), C.c_void_p(stream)))

    # ---- warm-up
    for s in range(a.warmup):
        step_device(s)
    sync_all()
    # ---- timed: device-resident
    clocks = ClockSampler(local_rank); clocks.start()
    launches0 = _lib.kernel_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    stage_names = ["features", "ivector", "nnet3", "decoder_advance", "decoder_finalize"]
    stage_ms = dict.fromkeys(stage_names, 0.0)
    sync_all()
    e0.record()
    for s in range(a.steps):
        step_device(s)
    e1.record()
    sync_all()
    launches = _lib.kernel_launch_count() - launches0
    t_dev = e0.elapsed_time(e1) / 1e3
    ms5 = (C.c_float * 5)()
    _lib.check(L.b2k_pipeline_stage_times(pipe.h, ms5))            # (the last step's split; the steps are alike)
    for k, v in zip(stage_names, ms5):
        stage_ms[k] = float(v)
    infos = dec_infos(L, L.b2k_pipeline_decoder(pipe.h), B)
    nerr_local = sum(1 for i in infos if i[0] != 0)
    collective_bytes = 0
    e2e_api = "b2k_pipeline_submit_i16 / b2k_pipeline_collect (C ABI), int16 host buffers, two batches in flight"
    if world == 1:
        # ---- timed: end to end through the C ABI with host buffers, batch k+1 submitted before batch k is collected
        for s in range(min(a.warmup, 2)):
            pipe.submit(None, ptrs=host_ptrs[s % n_sets])
            pipe.collect(copy=False)
        sync_all()
        t0 = time.perf_counter()
        d2h = 0
        lat_states = 0
        last = None
        pipe.submit(None, ptrs=host_ptrs[0])
        for s in range(1, a.steps + 1):
            if s < a.steps:
                pipe.submit(None, ptrs=host_ptrs[s % n_sets])
            last = pipe.collect(copy=(s == a.steps))
            d2h += sum(v.nbytes for v in last.values())
            lat_states += int(last["state_offs"][-1])
        torch.cuda.synchronize()
        t_e2e = time.perf_counter() - t0
        h2d_per_step = B * NUM_SAMPLES * 2
    else:
        # ---- timed: the whole step enters and leaves through rank 0 (kaldi_b200/ingest.py): rank 0 copies the
        #      packed int16 PCM of ALL ranks' utterances to its GPU, scatters the shards over NVLink (grouped ncclSend/Recv),
        #      every rank runs the pipeline on its shard and packs its lattices on the device, rank 0 gathers the packed
        #      buffers (size all-gather + grouped send/recv), copies them to the host and unpacks them.
        from kaldi_b200 import ingest
        e2e_api = ("rank 0: H2D of every rank's int16 PCM -> scatter (grouped ncclSend/ncclRecv) -> b2k_pipeline_run_device_i16 + "
                   "b2k_pipeline_pack_device on every rank -> size all-gather + grouped gather to rank 0 -> D2H + b2k_dec_unpack_lattices")
        dev = torch.device("cuda", local_rank)
        U = world * B
        speakers = [f"spk{u // 4}" for u in range(U)]               # four utterances per speaker: a speaker stays on one rank
        lengths = [NUM_SAMPLES] * U
        pcm_host = None
        if rank == 0:
            pcm_host = torch.from_numpy(np.concatenate([host_sets[0]] * world, axis=0)).pin_memory()
        pack_cap = max(64 << 20, B * (512 << 10))
        pack_buf = torch.empty(pack_cap, dtype=torch.uint8, device=dev)

        def compute(shard):
            n = int(shard.shape[0])
            pipe.run_device_i16(shard.data_ptr(), n, stream)
            pipe.pack_device(n, pack_buf.data_ptr(), pack_cap, stream)
            hb = int(L.b2k_dec_pack_header_bytes(n))
            hdr = pack_buf[:hb].cpu().numpy().view(np.int64)          # (waits for the stream: the sizes are needed to send)
            status, need = NativeBatchedPipeline.packed_bytes_needed(hdr, n)
            if status != 0:
                raise SystemExit(f"[bench] rank {rank}: packed lattices carry status {status} ({need} bytes needed, capacity {pack_cap})")
            return pack_buf[:need]
        L.b2k_dec_pack_header_bytes.restype = C.c_int64
        L.b2k_dec_pack_header_bytes.argtypes = [C.c_int32]

        def one_step():
            pcm_dev = pcm_host.to(dev, non_blocking=True) if rank == 0 else None
            parts, shards, nbytes = ingest.run_step(pcm_dev, speakers if rank == 0 else None, lengths if rank == 0 else None,
                                                    rank, world, compute, dev)
            got = None
            if rank == 0:
                got = [NativeBatchedPipeline.unpack_lattices(p_.cpu().numpy(), len(ids)) for p_, ids in zip(parts, shards)]
            return got, nbytes
        for s in range(min(a.warmup, 1)):
            one_step()
        sync_all()
        t0 = time.perf_counter()
        d2h = 0
        lat_states = 0
        last = None
        for s in range(a.steps):
            got, nbytes = one_step()
            if rank == 0:
                collective_bytes += nbytes
                d2h += sum(v.nbytes for g_ in got for v in g_.values())
                lat_states += sum(int(g_["state_offs"][-1]) for g_ in got)
                last = got[0]
        torch.cuda.synchronize()
        t_e2e = time.perf_counter() - t0
        h2d_per_step = U * NUM_SAMPLES * 2
        dist.barrier()
    clk = clocks.stop()
    # max over ranks
    if world > 1:
        t = torch.tensor([t_dev, t_e2e], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        t_dev, t_e2e = float(t[0]), float(t[1])
        e = torch.tensor([nerr_local], device="cuda"); dist.all_reduce(e); nerr = int(e[0])
    else:
        nerr = nerr_local
    audio_per_step = world * B * NUM_SAMPLES / 16000.0
    value = audio_per_step * a.steps / t_dev
    e2e_value = audio_per_step * a.steps / t_e2e

    # ---- parity spot check (outside the timed regions): the last e2e batch's first k lattices against the reference's own
    #      LatticeFasterDecoder (oracle/_ref) run on the log-likelihoods the device computed for them
    parity_checked = 0
    if rank == 0 and a.parity_utts > 0:
        from oracle import ref_decoder as RD
        k = min(a.parity_utts, B)
        ll = pipe.read("loglikes", k)
        rd = RD.RefDecoder(graph, synth.DEFAULT_DECODER_CFG)
        lats = CudaDecoder.SplitLattices(last)
        for i in range(k):
            rd.decode(ll[i])
            want, have = rd.lattice(), lattice_to_canonical(lats[i])
            for key in have:
                if not np.array_equal(have[key], want[key]):
                    raise SystemExit(f"[bench] parity check failed: utterance {i}, lattice field {key} differs from the compiled reference decoder")
            parity_checked += 1

    # ---- roofline of the dominant kernel (by measured share of the step)
    dom = max(stage_ms, key=stage_ms.get)
    arcs = sum(i[4] + i[5] for i in infos)
    ntok = sum(i[2] for i in infos)
    nlink = sum(i[3] for i in infos)
    traffic_file = os.path.join(ROOT, "profiles", "r02_decoder_dram.json")
    traffic = None
    traffic_source = None
    if os.path.exists(traffic_file):
        tj = json.load(open(traffic_file))
        if tj.get("workload") == wname:
            traffic = float(tj["dram_bytes_per_lane_frame"]) * B * nf
            traffic_source = tj.get("source")
    if dom.startswith("decoder"):
        # 16 B/arc examined + 16 B/source token + 36 B/link admitted + 16 B/token kept
        alg_bytes = 16.0 * arcs + 16.0 * ntok + 36.0 * nlink + 16.0 * ntok
        achieved = alg_bytes / (stage_ms["decoder_advance"] / 1e3) / 1e9
        roof = dict(kernel="dec_advance_v2_kernel" if not os.environ.get("B2K_DEC_V1") else "dec_advance_exact_kernel", bound="hbm",
                    achieved=achieved, peak=peaks["hbm_gbs"], unit="GB/s", frac=achieved / peaks["hbm_gbs"],
                    traffic=traffic, traffic_source=traffic_source, algorithmic_bytes=alg_bytes,
                    peak_source=peaks["source"], marcs_per_s=arcs / (stage_ms["decoder_advance"] / 1e3) / 1e6,
                    share_of_step=stage_ms["decoder_advance"] / sum(stage_ms.values()))
    elif dom == "nnet3":
        fl = L.b2k_pipeline_nnet_flops_per_utterance(pipe.h) * B
        achieved = fl / (stage_ms["nnet3"] / 1e3) / 1e12
        roof = dict(kernel="nnet_gemm_ts_kernel", bound="tensor", achieved=achieved, peak=peaks["bf16_tflops"], unit="TFLOP/s",
                    frac=achieved / peaks["bf16_tflops"], traffic=None, peak_source=peaks["source"],
                    share_of_step=stage_ms["nnet3"] / sum(stage_ms.values()))
    else:
        frames = B * pipe.plan.num_feature_frames
        achieved = frames * 800.0 / (stage_ms["features"] / 1e3) / 1e9
        roof = dict(kernel="feat_kernel", bound="hbm", achieved=achieved, peak=peaks["hbm_gbs"], unit="GB/s",
                    frac=achieved / peaks["hbm_gbs"], traffic=None, peak_source=peaks["source"],
                    share_of_step=stage_ms["features"] / sum(stage_ms.values()))

    cpu_base = None
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        cpu_base = run_cpu_baseline(a.cpu_utts, wname)
    if rank == 0:
        gemm = os.environ.get("B2K_NNET_GEMM", "ts")
        line = dict(metric="real-time factor (audio-sec/wall-sec)", value=value, unit="RTFx", n_gpus=world, steps=a.steps,
                    warmup=a.warmup, ms_per_step=1e3 * t_dev / a.steps, higher_is_better=True, scaling="weak",
                    vs_baseline=None, dtype="f32", data="synthetic",
                    config=dict(workload=wname, utterances_per_gpu_per_step=B, audio_s_per_step=audio_per_step,
                                model=w["model"], graph_arcs=int(graph["offsets"][-1]), beam=15.0, max_active=7000, lattice_beam=8.0,
                                decoder_mode="reference_order",
                                ivector="online 100-dim, 512-Gaussian UBM, re-estimated per nnet chunk (synthetic extractor)",
                                cache="inputs larger than L2: log-likes %.0f MB, audio %.0f MB per step" %
                                      (B * nf * P * 4 / 1e6, B * NUM_SAMPLES * 4 / 1e6),
                                parallelism=f"dp{world} (speaker-ordered utterance shards; replicas of model / extractor / HCLG; "
                                            "collectives only at ingest (PCM scatter) and egress (lattice gather) of the e2e path)"),
                    e2e=dict(value=e2e_value, unit="RTFx", h2d_bytes_per_step=h2d_per_step,
                             d2h_bytes_per_step=int(d2h / a.steps), lattice_states_per_step=int(lat_states / a.steps),
                             collective_bytes_per_step=int(collective_bytes / a.steps), api=e2e_api),
                    gpu_launches=int(launches), stage_ms={k: round(v, 3) for k, v in stage_ms.items()},
                    decoder_phase_share=None if not os.environ.get("B2K_DEC_PROF") else dict(zip(
                        # slots of dec_advance_v2_lane / finish_frame_v2 (B2K_TICK): 2 carries the walk's sizing counters, not cycles
                        ["cutoff_seed", "expand", "(walk_size_counters)", "walk_arrays_build", "closure_start", "closure_rounds", "replay_prep",
                         "walk_single_thread", "eps_keys_zero", "list_order_scan", "commit_links_remap", "clear"],
                        [round(float(sum(i[16 + k] for i in infos)) / max(1.0, float(sum(i[16 + 15] for i in infos))), 3)
                         for k in range(12)])),
                    replay_routes=None if not os.environ.get("B2K_DEC_PROF") else dict(
                        marking_in_smem=sum(i[16 + 2] & 0xfffff for i in infos) / max(1, sum(i[1] for i in infos)),
                        ids_in_smem=sum((i[16 + 2] >> 20) & 0xfffff for i in infos) / max(1, sum(i[1] for i in infos)),
                        walk_by_components=sum((i[16 + 2] >> 40) & 0xfffff for i in infos) / max(1, sum(i[1] for i in infos))),
                    eps_replay_per_frame=dict(zip(["pops", "arc_visits", "in_shared_memory"],
                        [round(float(sum(i[16 + k] for i in infos)) / max(1.0, float(sum(i[1] for i in infos))), 2)
                         for k in (12, 13, 14)])),
                    decoder=dict(marcs_per_s=arcs / (stage_ms["decoder_advance"] / 1e3) / 1e6 * world,
                                 arcs_per_frame=arcs / (B * nf),
                                 eps_arcs_per_frame=sum(i[5] for i in infos) / (B * nf),
                                 links_per_frame=nlink / (B * nf), tokens_per_frame=ntok / (B * nf), errors=nerr),
                    nnet3=dict(tflops=L.b2k_pipeline_nnet_flops_per_utterance(pipe.h) * B / (stage_ms["nnet3"] / 1e3) / 1e12 * world,
                               mma={"ts": "3xTF32 tcgen05.mma kind::tf32 (A in TMEM, W by TMA), fp32 accumulate in TMEM",
                                    "tcgen05": "3xTF32 tcgen05.mma kind::tf32 (operands staged by threads)",
                                    "mma": "3xTF32 mma.sync m16n8k8", "simt": "fp32 FFMA (SIMT)"}.get(gemm, gemm)),
                    roofline=roof, clocks=clk, parity_checked=parity_checked, host=usable_cores())
        if cpu_base:
            line["cpu_baseline"] = cpu_base
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
