#!/usr/bin/env python
"""bench.py — headline benchmark of the online2 inference hot path on B200.

Metric (BASELINE.json): real-time factor RTFx = audio-seconds / wall-seconds, on
config[1]: mini_librispeech TDNN-F chain model (run_tdnn_1k.sh shapes, 2336
pdfs, synthetic weights), synthetic 16 kHz 10 s utterances, beam 15 /
max-active 7000 / lattice-beam 8, synthetic 5 M-arc HCLG.  A "step" is one
batch of utterances through features -> i-vectors -> nnet3 -> decoder ->
finalized raw lattices.

  python bench.py --gpus N --steps K --warmup W          # our arm
  python bench.py --impl reference --gpus N ...           # the reference's CPU path on the host cores

`value`  : whole-job RTFx with the audio already resident in HBM (CUDA events).
`e2e`    : same through BatchedPipeline.decode_batch with HOST buffers: pinned
           host audio -> H2D -> ... -> finalized lattices packed -> D2H, all timed.
`roofline`: the dominant kernel (largest share of the step, measured live).
`cpu_baseline`: the reference CPU path on a bounded sample (rank 0, N=1 only).
"""
from __future__ import annotations

import argparse
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

WORKLOAD = "mini_librispeech_tdnn_1k/10s-utts/beam15/hclg5M"
NUM_SAMPLES = 160000
GRAPH_ARCS = 5_000_000
NUM_PDFS = 2336


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm_gbs=float(d["hbm_gbs"]), bf16_tflops=float(d.get("bf16_tflops_sustained", d.get("bf16_tflops", 0))),
                    source="measured (MEASURED_PEAKS.json)")
    return dict(hbm_gbs=6650.0, bf16_tflops=1590.0, source="fallback (B200_PROFILING.md)")


def build_inputs(batch: int, seed0: int):
    return [synth.make_audio(NUM_SAMPLES, seed=seed0 + i) for i in range(batch)]


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

def cpu_reference_one(args):
    """One utterance through the reference's CPU path, all of it the reference's own
    sources compiled in oracle/_ref: features, i-vector, nnet3 looped forward and
    LatticeFasterDecoder (InitDecoding / AdvanceDecoding / FinalizeDecoding)."""
    seed, arch_seed = args
    os.environ.setdefault("OPENBLAS_NUM_THREADS", "1")
    from oracle import dec_oracle as D, feat_oracle as F, nnet_oracle as NO
    g = _CPU_STATE.get("graph")
    if g is None:
        g = _CPU_STATE["graph"] = synth.make_hclg(GRAPH_ARCS, num_pdfs=NUM_PDFS, seed=1)
        arch = NM.arch_mini_librispeech_1k(NUM_PDFS)
        W = load_calibrated_weights(arch, arch_seed)
        _CPU_STATE["nnet"] = NO.RefNnet(arch, W)
        _CPU_STATE["feat"] = F.RefFeat()
        from oracle import ref_decoder as RD
        _CPU_STATE["dec"] = RD.RefDecoder(g, synth.DEFAULT_DECODER_CFG)
        from oracle import ivector_oracle as IV
        _CPU_STATE["ivx"] = IV.make_cpu_extractor(arch_seed)
    t0 = time.time()
    wave = synth.make_audio(NUM_SAMPLES, seed=seed)
    t1 = time.time()
    feats = _CPU_STATE["feat"].compute(wave, F.FeatOpts(), online_chunk=2880)
    R = _CPU_STATE["nnet"]
    n_chunks = len(R.chunk_ivector_rows(feats.shape[0], feats.shape[0], 1))
    if _CPU_STATE["ivx"] is not None:
        civ = _CPU_STATE["ivx"].chunk_ivectors(feats, n_chunks, R.frames_per_chunk, R.right_context)
    else:
        civ = np.zeros((n_chunks, 100), np.float32)
    ends = [(n + 1) * R.frames_per_chunk + R.right_context for n in range(n_chunks)]
    mat = np.zeros((ends[-1] + 1, 100), np.float32)
    prev = 0
    for n, e in enumerate(ends):
        mat[prev:e + 1] = civ[n]
        prev = e + 1
    ll = R.forward(feats, mat, period=1)
    _CPU_STATE["dec"].decode(ll, record=_CPU_STATE.get("record", False))   # record: tools/bench_lattice_det.py
    t2 = time.time()
    ns, na, _ = _CPU_STATE["dec"].lattice_sizes()
    return (t2 - t1, ns, na)


_CPU_STATE = {}
CAL_PATH = os.path.join(ROOT, "gpurun_out", "bench_calibration.npz")
CAL_FALLBACK = os.path.join(ROOT, "tests", "golden", "bench_calibration.npz")


def load_calibrated_weights(arch, seed):
    """Synthetic weights + the output calibration the GPU arm computed (committed
    fixture so that the CPU arm decodes exactly the same model)."""
    W = NM.random_weights(arch, seed=seed)
    for p in (CAL_PATH, CAL_FALLBACK):
        if os.path.exists(p):
            c = np.load(p)
            return NM.apply_output_calibration(W, c["mean"], float(c["scale"]))
    return W


def run_cpu_reference(num_utts: int, workers: int, pool=None, seed0: int = 1000):
    t0 = time.time()
    if workers <= 1:
        cpu_reference_one((900, 0))                                   # builds graph/model/extractor once (untimed)
        res = [cpu_reference_one((seed0 + i, 0)) for i in range(num_utts)]
        t_total = sum(r[0] for r in res)
    else:
        t0 = time.time()
        res = pool.map(cpu_reference_one, [(seed0 + i, 0) for i in range(num_utts)], chunksize=1)
        t_total = time.time() - t0
    audio = num_utts * NUM_SAMPLES / 16000.0
    return dict(rtfx=audio / t_total, wall_s=t_total, utts=num_utts, arcs=sum(r[2] for r in res))


def make_cpu_pool(workers: int):
    import multiprocessing as mp
    os.environ["OPENBLAS_NUM_THREADS"] = "1"
    os.environ["OMP_NUM_THREADS"] = "1"
    pool = mp.get_context("fork").Pool(workers)
    pool.map(cpu_reference_one, [(900 + i, 0) for i in range(workers)], chunksize=1)   # warm every worker
    return pool


# ----------------------------------------------------------------------------- main

# dram__bytes_read.sum + dram__bytes_write.sum of dec_advance_exact_kernel per (lane, frame), profiles/r01_final_ncu_summary.md
NCU_DRAM_BYTES_PER_LANE_FRAME = 3.66e6


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=592, help="utterances per GPU per step (592 = 148 SMs x 4 resident decoder CTAs)")
    ap.add_argument("--order-free", action="store_true", help="use the order-free decoder mode (not reference exact)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-utts", type=int, default=6)
    ap.add_argument("--max-tpf", type=int, default=32768, help="decoder per-frame token capacity (hash = 2x slots)")
    ap.add_argument("--tok-per-frame", type=int, default=9000, help="decoder token arena sizing (avg tokens/frame)")
    ap.add_argument("--links-per-frame", type=int, default=16000, help="decoder link arena sizing (avg links/frame)")
    a = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if a.impl == "reference":
        if rank != 0:
            return 0
        cores = os.cpu_count() or 1
        per_step = 2 * cores
        pool = make_cpu_pool(cores)
        vals = []
        t_all = 0.0
        for s in range(a.warmup + a.steps):
            r = run_cpu_reference(per_step, cores, pool, seed0=1000 + 100 * s)
            if s >= a.warmup:
                vals.append(r["rtfx"]); t_all += r["wall_s"]
        pool.close()
        v = float(np.mean(vals))
        line = dict(metric="real-time factor (audio-sec/wall-sec)", value=v, unit="RTFx", n_gpus=a.gpus, steps=a.steps,
                    warmup=a.warmup, ms_per_step=1e3 * t_all / max(a.steps, 1), higher_is_better=True, scaling="weak",
                    vs_baseline=None, dtype="f32", data="synthetic", impl="reference",
                    config=dict(workload=WORKLOAD, sample=f"{per_step} utterances of 10 s per step"),
                    cpu_baseline=dict(value=v, unit="RTFx", cores=cores, kind="reference",
                                      sample=f"{per_step} x 10 s utterances per step, one per worker process; features+nnet3 = "
                                             "the reference's own sources compiled in oracle/_ref (OpenBLAS, 1 thread/worker), "
                                             "decoder = the reference's lattice-faster-decoder.cc compiled against a container-only OpenFst stand-in; lattice determinization excluded"),
                    e2e=dict(value=v, unit="RTFx", h2d_bytes_per_step=0, d2h_bytes_per_step=0))
        print(json.dumps(line))
        return 0

    import torch
    import torch.distributed as dist
    from kaldi_b200 import _lib
    from kaldi_b200.pipeline import BatchedPipeline, PipelineConfig

    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl")
    peaks = load_peaks()
    arch = NM.arch_mini_librispeech_1k(NUM_PDFS)
    W = NM.random_weights(arch, seed=0)
    graph = synth.make_hclg(GRAPH_ARCS, num_pdfs=NUM_PDFS, seed=1)
    B = a.batch
    nf_out = 333
    from kaldi_b200.feat import FeatureOptions
    cfg = PipelineConfig(feature_opts=FeatureOptions(max_lanes=max(B, 64)), max_batch=B, num_samples=NUM_SAMPLES,
                         reference_order=not a.order_free,
                         max_tokens=nf_out * a.tok_per_frame, max_links=nf_out * a.links_per_frame,
                         max_tokens_per_frame=a.max_tpf)
    from kaldi_b200.ivector import make_synthetic_extractor
    ivx = make_synthetic_extractor(seed=0)
    pipe = BatchedPipeline(cfg, arch, W, graph, ivector_extractor=ivx)
    # prior-style calibration of the synthetic model (see nnet_model.apply_output_calibration)
    if os.path.exists(CAL_FALLBACK):
        c = np.load(CAL_FALLBACK)
        mean, scale = c["mean"], float(c["scale"])
    else:
        mean, scale = pipe.output_calibration(build_inputs(min(B, 16), seed0=777_000), target_std=1.0)
        if rank == 0:
            os.makedirs(os.path.dirname(CAL_PATH), exist_ok=True)
            np.savez(CAL_PATH, mean=mean, scale=np.float32(scale))
    W = NM.apply_output_calibration(W, mean, scale)
    del pipe
    torch.cuda.empty_cache()
    pipe = BatchedPipeline(cfg, arch, W, graph, ivector_extractor=ivx)
    # distinct utterances per rank and per step slot (cycled)
    n_sets = 2
    host_sets = [build_inputs(B, seed0=10_000 * rank + 1000 * s) for s in range(n_sets)]
    dev_sets = [torch.from_numpy(np.stack(hs)).cuda() for hs in host_sets]

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def step_device(s):
        pipe.d_wave[:B].copy_(dev_sets[s % n_sets])
        pipe.run_device(B)

    stage_ms = dict(features=0.0, ivector=0.0, nnet3=0.0, decoder_advance=0.0, decoder_finalize=0.0)

    def step_device_staged(s):
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(6)]
        pipe.d_wave[:B].copy_(dev_sets[s % n_sets])
        ev[0].record(); pipe.compute_features(B)
        ev[1].record(); pipe.compute_ivectors(B)
        ev[2].record(); pipe.compute_nnet(B)
        ev[3].record()
        ch = list(range(B)); P, nf = arch["num_pdfs"], pipe.nnet.n_out
        pipe.dec.InitDecoding(ch)
        lp = [pipe.d_loglikes.data_ptr() + 4 * nf * P * i for i in range(B)]
        pipe.dec.AdvanceDecodingFrames(ch, lp, [nf] * B, P)
        ev[4].record(); pipe.dec.FinalizeDecoding(ch)
        ev[5].record()
        return ev

    # ---- warm-up
    for s in range(a.warmup):
        step_device(s)
    sync_all()
    # ---- timed: device-resident
    clocks = ClockSampler(local_rank); clocks.start()
    launches0 = _lib.kernel_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync_all()
    e0.record()
    evs = [step_device_staged(s) for s in range(a.steps)]
    e1.record()
    sync_all()
    launches = _lib.kernel_launch_count() - launches0
    t_dev = e0.elapsed_time(e1) / 1e3
    for ev in evs:
        for k, (i, j) in zip(stage_ms, [(0, 1), (1, 2), (2, 3), (3, 4), (4, 5)]):
            stage_ms[k] += ev[i].elapsed_time(ev[j]) / a.steps
    infos = [pipe.dec.ChannelInfo(c) for c in range(B)]
    errs = [i["status"] for i in infos if i["status"] != 0]
    # ---- timed: end to end through the public API with host buffers
    for s in range(min(a.warmup, 2)):
        pipe.decode_batch(host_sets[s % n_sets], want_lattices=True)
    sync_all()
    t0 = time.perf_counter()
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2.record()
    d2h = 0
    lat_states = 0
    for s in range(a.steps):
        try:
            lats = pipe.decode_batch(host_sets[s % n_sets], want_lattices=True)
        except Exception as ex:          # e.g. an arena overflow on some channel: report, do not hide
            print(f"[bench] decode_batch failed: {ex}", file=sys.stderr)
            raise
        d2h += sum(v.nbytes for k, v in lats.items() if hasattr(v, "nbytes"))
        lat_states += int(lats["state_offs"][-1])
    e3.record()
    sync_all()
    t_e2e = max(e2.elapsed_time(e3) / 1e3, time.perf_counter() - t0)
    clk = clocks.stop()
    # max over ranks
    if world > 1:
        t = torch.tensor([t_dev, t_e2e], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        t_dev, t_e2e = float(t[0]), float(t[1])
        e = torch.tensor([len(errs)], device="cuda"); dist.all_reduce(e); nerr = int(e[0])
    else:
        nerr = len(errs)
    audio_per_step = world * B * NUM_SAMPLES / 16000.0
    value = audio_per_step * a.steps / t_dev
    e2e_value = audio_per_step * a.steps / t_e2e

    # ---- roofline of the dominant kernel (by measured share of the step)
    dom = max(stage_ms, key=stage_ms.get)
    arcs = sum(i["arcs_emitting"] + i["arcs_nonemitting"] for i in infos)
    ntok = sum(i["ntok"] for i in infos)
    nlink = sum(i["nlink"] for i in infos)
    if dom.startswith("decoder"):
        # 16 B/arc examined + 16 B/source token + 36 B/link admitted + 16 B/token kept
        alg_bytes = 16.0 * arcs + 16.0 * ntok + 36.0 * nlink + 16.0 * ntok
        achieved = alg_bytes / (stage_ms["decoder_advance"] / 1e3) / 1e9
        roof = dict(kernel="dec_advance_exact_kernel" if not a.order_free else "dec_advance_kernel", bound="hbm",
                    achieved=achieved, peak=peaks["hbm_gbs"], unit="GB/s", frac=achieved / peaks["hbm_gbs"],
                    traffic=(NCU_DRAM_BYTES_PER_LANE_FRAME * B * pipe.nnet.n_out) if not a.order_free else None,
                    traffic_source="ncu --set full (profiles/r01_final_ncu_summary.md): dram read+write per lane-frame at 592 lanes "
                                   "on the decoder-only synthetic load, scaled to this launch's lane-frames",
                    algorithmic_bytes=alg_bytes,
                    peak_source=peaks["source"], marcs_per_s=arcs / (stage_ms["decoder_advance"] / 1e3) / 1e6,
                    share_of_step=stage_ms["decoder_advance"] / sum(stage_ms.values()))
    elif dom == "nnet3":
        fl = pipe.nnet.flops_per_utt * B
        achieved = fl / (stage_ms["nnet3"] / 1e3) / 1e12
        roof = dict(kernel="nnet_gemm_kernel", bound="tensor", achieved=achieved, peak=peaks["bf16_tflops"], unit="TFLOP/s",
                    frac=achieved / peaks["bf16_tflops"], traffic=None, peak_source=peaks["source"],
                    share_of_step=stage_ms["nnet3"] / sum(stage_ms.values()))
    else:
        frames = B * pipe.T
        achieved = frames * 800.0 / (stage_ms["features"] / 1e3) / 1e9
        roof = dict(kernel="feat_kernel", bound="hbm", achieved=achieved, peak=peaks["hbm_gbs"], unit="GB/s",
                    frac=achieved / peaks["hbm_gbs"], traffic=None, peak_source=peaks["source"],
                    share_of_step=stage_ms["features"] / sum(stage_ms.values()))

    cpu_base = None
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        r = run_cpu_reference(a.cpu_utts, 1)
        cpu_base = dict(value=r["rtfx"], unit="RTFx", cores=1, kind="reference",
                        sample=f"{a.cpu_utts} x 10 s utterances, single thread: features + nnet3 are the reference's own "
                               "sources compiled in oracle/_ref (OpenBLAS 1 thread), decoder = the reference's lattice-faster-decoder.cc "
                               "compiled against a container-only OpenFst stand-in; determinization excluded")
    if rank == 0:
        line = dict(metric="real-time factor (audio-sec/wall-sec)", value=value, unit="RTFx", n_gpus=world, steps=a.steps,
                    warmup=a.warmup, ms_per_step=1e3 * t_dev / a.steps, higher_is_better=True, scaling="weak",
                    vs_baseline=None, dtype="f32", data="synthetic",
                    config=dict(workload=WORKLOAD, utterances_per_gpu_per_step=B, audio_s_per_step=audio_per_step,
                                model="mini_librispeech tdnn_1k (4.47 M params fwd path, 2336 pdfs)",
                                graph_arcs=int(graph["offsets"][-1]), beam=15.0, max_active=7000, lattice_beam=8.0,
                                decoder_mode="order_free" if a.order_free else "reference_order",
                                ivector="online 100-dim, 512-Gaussian UBM, re-estimated per nnet chunk (synthetic extractor)",
                                cache="inputs larger than L2: log-likes %.0f MB, audio %.0f MB per step" %
                                      (B * pipe.nnet.n_out * NUM_PDFS * 4 / 1e6, B * NUM_SAMPLES * 4 / 1e6),
                                parallelism=f"dp{world} (utterance shards, no data-path collective)"),
                    e2e=dict(value=e2e_value, unit="RTFx", h2d_bytes_per_step=B * NUM_SAMPLES * 4,
                             d2h_bytes_per_step=int(d2h / a.steps), lattice_states_per_step=int(lat_states / a.steps),
                             host_ms_last_step={k: round(v, 1) for k, v in getattr(pipe, "last_host_ms", {}).items()}),
                    gpu_launches=int(launches), stage_ms={k: round(v, 3) for k, v in stage_ms.items()},
                    decoder_phase_share=None if not os.environ.get("B2K_DEC_PROF") else dict(zip(
                        ["cutoff_seed", "expand", "rank", "bucket_scatter", "eps_init", "eps_closure", "replay_prep", "eps_replay",
                         "eps_finish", "list_order", "eps_links", "commit"],
                        [round(float(sum(i["prof_cycles"][k] for i in infos)) / max(1.0, float(sum(i["prof_cycles"][15] for i in infos))), 3)
                         for k in range(12)])),
                    eps_replay_per_frame=dict(zip(["pops", "arc_visits", "in_shared_memory"],
                        [round(float(sum(i["prof_cycles"][k] for i in infos)) / max(1.0, float(sum(i["frames_decoded"] for i in infos))), 2)
                         for k in (12, 13, 14)])),
                    decoder=dict(marcs_per_s=arcs / (stage_ms["decoder_advance"] / 1e3) / 1e6 * world,
                                 arcs_per_frame=arcs / (B * pipe.nnet.n_out),
                                 eps_arcs_per_frame=sum(i["arcs_nonemitting"] for i in infos) / (B * pipe.nnet.n_out),
                                 links_per_frame=sum(i["nlink"] for i in infos) / (B * pipe.nnet.n_out), tokens_per_frame=ntok / (B * pipe.nnet.n_out),
                                 errors=nerr),
                    nnet3=dict(tflops=pipe.nnet.flops_per_utt * B / (stage_ms["nnet3"] / 1e3) / 1e12 * world,
                               mma="3xTF32 mma.sync m16n8k8, fp32 accumulate" if os.environ.get("B2K_NNET_GEMM") != "simt" else "fp32 FFMA (SIMT)"),
                    roofline=roof, clocks=clk)
        if cpu_base:
            line["cpu_baseline"] = cpu_base
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
