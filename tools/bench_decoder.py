"""Decoder-only micro benchmark (config 5 of BASELINE.json): precomputed
log-likes, synthetic HCLG; prints Marcs/s, frames/s, RTFx-equivalent for both
decoder modes.  Not the headline bench (that is bench.py)."""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from kaldi_b200 import synth
from kaldi_b200.decoder import CudaFst, CudaDecoder, CudaDecoderConfig

ap = argparse.ArgumentParser()
ap.add_argument("--arcs", type=int, default=5_000_000)
ap.add_argument("--lanes", type=int, default=256)
ap.add_argument("--frames", type=int, default=333)
ap.add_argument("--pdfs", type=int, default=2336)
ap.add_argument("--beam", type=float, default=15.0)
ap.add_argument("--distinct", type=int, default=16, help="distinct loglike matrices (cycled over lanes)")
ap.add_argument("--modes", default="ref,free")
ap.add_argument("--reps", type=int, default=2)
a = ap.parse_args()

g = synth.make_hclg(a.arcs, num_pdfs=a.pdfs, seed=1)
fst = CudaFst(g)
lls = [torch.from_numpy(synth.make_loglikes(g, a.frames, seed=100 + i)).cuda() for i in range(a.distinct)]
cfgd = dict(synth.DEFAULT_DECODER_CFG, beam=a.beam)
for mode in a.modes.split(","):
    cfg = CudaDecoderConfig.from_dict(cfgd, max_frames=a.frames + 2, max_tokens=int(a.frames * 9000),
                                      max_links=int(a.frames * 16000), reference_order=(mode == "ref"))
    dec = CudaDecoder(fst, cfg, a.lanes)
    ch = list(range(a.lanes))
    ptrs = [lls[i % a.distinct].data_ptr() for i in ch]
    best = None
    for rep in range(a.reps + 1):
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        dec.InitDecoding(ch)
        torch.cuda.synchronize()
        e0.record()
        dec.AdvanceDecodingFrames(ch, ptrs, [a.frames] * a.lanes, lls[0].stride(0))
        e1.record()
        dec.FinalizeDecoding(ch)
        e2.record()
        torch.cuda.synchronize()
        t_adv, t_fin = e0.elapsed_time(e1) / 1e3, e1.elapsed_time(e2) / 1e3
        if (rep > 0 or a.reps == 0) and (best is None or t_adv < best[0]):
            best = (t_adv, t_fin)
    infos = [dec.ChannelInfo(c) for c in ch]
    bad = [i["status"] for i in infos if i["status"] != 0]
    arcs = sum(i["arcs_emitting"] + i["arcs_nonemitting"] for i in infos)
    toks = sum(i["ntok"] for i in infos)
    links = sum(i["nlink"] for i in infos)
    t_adv, t_fin = best
    audio_s = a.lanes * a.frames * 0.03
    print(json.dumps(dict(mode=mode, lanes=a.lanes, frames=a.frames, arcs_graph=a.arcs, beam=a.beam,
                          advance_s=round(t_adv, 4), finalize_s=round(t_fin, 4),
                          marcs_per_s=round(arcs / t_adv / 1e6, 1),
                          arcs_per_frame=round(arcs / (a.lanes * a.frames)),
                          toks_per_frame=round(toks / (a.lanes * a.frames)),
                          links_per_frame=round(links / (a.lanes * a.frames)),
                          rtfx_decoder_only=round(audio_s / (t_adv + t_fin)),
                          errors=len(bad))))
    del dec
    torch.cuda.empty_cache()
