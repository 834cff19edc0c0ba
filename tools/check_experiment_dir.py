"""Host-only check of a Kaldi online-decoding directory against the C++ readers of libb2k.so: does every file the hot path needs
parse, are the pieces consistent with each other, and what would the batched pipeline allocate for a given utterance length?
Nothing here needs a GPU.  Layout expected (what steps/online/nnet3/prepare_online_decoding.sh leaves behind):

    <dir>/final.mdl   <dir>/conf/online.conf   (+ the files online.conf points at: mfcc.conf, ivector_extractor.conf,
    splice.conf, online_cmvn.conf, final.mat, global_cmvn.stats, final.dubm, final.ie)      <graph>/HCLG.fst

usage: python tools/check_experiment_dir.py <dir> <graph>/HCLG.fst [seconds-per-utterance] [batch]
Prints one JSON object; exit status 0 if everything is usable, 1 otherwise (the message of the first failing reader is in
"error")."""
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def check(exp_dir: str, hclg: str, seconds: float = 10.0, batch: int = 64) -> dict:
    from kaldi_b200 import _lib
    from kaldi_b200.decoder import read_fst_file
    from kaldi_b200.feat import _FeatCfg
    from kaldi_b200.ivector import IvectorFiles
    from kaldi_b200.model import KaldiModel
    L = _lib.lib()
    out = {"dir": exp_dir, "graph": hclg}

    class OnlineConf(C.Structure):
        _fields_ = [("feature_type", C.c_int32), ("add_pitch", C.c_int32)] + \
                   [(k, C.c_char * 512) for k in ("mfcc_config", "fbank_config", "cmvn_config", "global_cmvn_stats", "ivector_extraction_config")] + \
                   [("rest", C.c_char * 4096), ("plp_config", C.c_char * 512)]

    class IvecCfg(C.Structure):
        _fields_ = [(k, C.c_int32) for k in ("base_dim", "splice_left", "splice_right", "feat_dim", "num_gauss", "ivector_dim", "num_gselect")] + \
                   [(k, C.c_float) for k in ("min_post", "posterior_scale", "max_count", "prior_offset")] + \
                   [(k, C.c_int32) for k in ("num_cg_iters", "cmn_window", "speaker_frames", "global_frames", "max_lanes", "max_frames")]

    class IvecPaths(C.Structure):
        _fields_ = [(k, C.c_char * 512) for k in ("lda_matrix", "global_cmvn_stats", "splice_config", "cmvn_config", "diag_ubm", "ivector_extractor")] + \
                   [("ivector_period", C.c_int32), ("use_most_recent_ivector", C.c_int32), ("greedy_ivector_extractor", C.c_int32),
                    ("online_cmvn_iextractor", C.c_int32), ("max_remembered_frames", C.c_float)]
    oc = OnlineConf()
    L.b2k_online_conf_read.argtypes = [C.c_char_p, C.c_void_p]
    _lib.check(L.b2k_online_conf_read(os.path.join(exp_dir, "conf", "online.conf").encode(), C.byref(oc)))
    out["feature_type"] = ["mfcc", "fbank", "plp"][oc.feature_type]
    out["other_options"] = oc.rest.decode().split()
    fc = _FeatCfg()
    conf = (oc.mfcc_config if oc.feature_type == 0 else oc.fbank_config if oc.feature_type == 1 else oc.plp_config).decode()
    L.b2k_feat_cfg_from_conf.argtypes = [C.c_char_p, C.c_int32, C.c_void_p]
    _lib.check(L.b2k_feat_cfg_from_conf(conf.encode(), oc.feature_type, C.byref(fc)))
    feat_dim = fc.num_ceps if oc.feature_type == 0 else fc.num_bins + (1 if fc.use_energy else 0)
    out["features"] = dict(config=conf, dim=feat_dim, samp_freq=fc.samp_freq, dither=fc.dither, num_mel_bins=fc.num_bins, snip_edges=bool(fc.snip_edges))
    warnings = []
    if fc.dither != 0.0:
        warnings.append("dither is %g: the reference's dithering is unseeded, b2k_feat_create requires --dither=0" % fc.dither)
    m = KaldiModel(os.path.join(exp_dir, "final.mdl"), is_mdl=True)
    out["model"] = dict(input_dim=m.feat_dim, ivector_dim=m.ivector_dim, num_pdfs=m.num_pdfs, frame_subsampling_factor=m.frame_subsampling_factor,
                        layers=m.n_layers, transition_ids=(len(m.tid2pdf) - 1 if m.tid2pdf is not None else 0), has_priors=m.has_priors)
    if m.feat_dim != feat_dim:
        raise RuntimeError(f"feature dimension {feat_dim} differs from the model's input dimension {m.feat_dim}")
    if m.ivector_dim > 0:
        ic, ip = IvecCfg(), IvecPaths()
        ic.base_dim, ic.max_lanes, ic.max_frames = feat_dim, batch, 1
        L.b2k_ivec_cfg_from_conf.argtypes = [C.c_char_p, C.c_void_p, C.c_void_p]
        _lib.check(L.b2k_ivec_cfg_from_conf(oc.ivector_extraction_config, C.byref(ic), C.byref(ip)))
        f = IvectorFiles(ip.ivector_extractor.decode(), ip.diag_ubm.decode(), ip.lda_matrix.decode(), ip.global_cmvn_stats.decode())
        out["ivector_extractor"] = dict(num_gauss=f.num_gauss, feat_dim=f.feat_dim, ivector_dim=f.ivector_dim, lda=[f.lda_rows, f.lda_cols],
                                        splice=[ic.splice_left, ic.splice_right], num_gselect=ic.num_gselect, max_count=ic.max_count,
                                        posterior_scale=ic.posterior_scale, min_post=ic.min_post)
        if f.ivector_dim != m.ivector_dim:
            raise RuntimeError(f"extractor i-vector dimension {f.ivector_dim} differs from the model's {m.ivector_dim}")
        if f.lda_cols != feat_dim * (ic.splice_left + 1 + ic.splice_right) + 1:
            raise RuntimeError("final.mat does not match the splicing options")
        if ic.splice_left != ic.splice_right:
            warnings.append("asymmetric splicing: the extractor kernel takes one context width on both sides")
    g = read_fst_file(hclg)
    out["graph_info"] = dict(type=g["fst_type"], states=g["num_states"], arcs=int(g["offsets"][-1]), max_ilabel=int(g["ilabel"].max()) if len(g["ilabel"]) else 0)
    if m.tid2pdf is not None and out["graph_info"]["max_ilabel"] >= len(m.tid2pdf):
        raise RuntimeError("HCLG.fst uses transition-ids the model does not have")
    # what the pipeline would allocate
    from kaldi_b200.pipeline import _native_structs
    PC, PP = _native_structs()
    pc, pl = PC(), PP()
    L.b2k_pipeline_cfg_default.argtypes = [C.c_void_p]
    L.b2k_pipeline_cfg_default.restype = None
    L.b2k_pipeline_cfg_default(C.byref(pc))
    pc.feat = fc
    pc.feat.dither = 0.0
    pc.max_batch, pc.num_samples = batch, int(round(seconds * fc.samp_freq))
    L.b2k_pipeline_plan_for.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    _lib.check(L.b2k_pipeline_plan_for(C.byref(pc), m.h, C.byref(pl)))
    out["plan"] = dict(batch=batch, seconds=seconds, feature_frames=pl.num_feature_frames, output_frames=pl.num_output_frames, nnet_chunks=pl.num_chunks,
                       decoder_max_tokens=pl.dec.max_tokens, decoder_max_links=pl.dec.max_links, device_bytes=pl.device_bytes, pinned_bytes=pl.pinned_bytes)
    out["warnings"] = warnings
    out["ok"] = True
    return out


if __name__ == "__main__":
    if len(sys.argv) < 3:
        print(__doc__)
        sys.exit(2)
    try:
        res = check(sys.argv[1], sys.argv[2], float(sys.argv[3]) if len(sys.argv) > 3 else 10.0, int(sys.argv[4]) if len(sys.argv) > 4 else 64)
    except Exception as e:                                    # noqa: BLE001 - report whatever the readers said
        res = {"ok": False, "error": str(e)}
    print(json.dumps(res, indent=1))
    sys.exit(0 if res.get("ok") else 1)
