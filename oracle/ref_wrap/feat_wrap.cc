// oracle/ref_wrap/feat_wrap.cc — TEST INFRASTRUCTURE ONLY.
// Thin extern "C" wrapper (our code) around the reference's OWN feature code,
// which is compiled from the sources where they lie under /root/reference/src
// (see oracle/ref_feat.py).  Nothing from the reference is copied here; this
// file only calls its public API:
//   Mfcc/Fbank::ComputeFeatures            feat/feature-common-inl.h
//   OnlineMfcc/OnlineFbank, OnlineCmvn     feat/online-feature.{h,cc}
//   OnlineMatrixFeature                    feat/online-feature.h
#include <cstdio>
#include <cstring>
#include <fstream>
#include <memory>
#include <vector>

#include "feat/feature-fbank.h"
#include "feat/feature-mfcc.h"
#include "feat/feature-plp.h"
#include "feat/online-feature.h"
#include "matrix/kaldi-matrix.h"
#include "feat/resample.h"
#include "feat/wave-reader.h"
#include "util/parse-options.h"

using namespace kaldi;

extern "C" {

struct ref_feat_opts {
  int feature_type;        // 0 mfcc, 1 fbank
  float samp_freq, frame_shift_ms, frame_length_ms, dither, preemph_coeff;
  int remove_dc_offset, round_to_power_of_two, snip_edges;
  int num_bins;
  float low_freq, high_freq;
  int num_ceps, use_energy;
  float energy_floor;
  int raw_energy;
  float cepstral_lifter;
  int htk_compat, use_log_fbank, use_power;
  int window_type;         // 0 povey, 1 hamming, 2 hanning, 3 rectangular
  int htk_mode;            // MelBanksOptions::htk_mode (hidden test option)
  int lpc_order;           // PlpOptions (feature_type 2)
  float compress_factor, cepstral_scale;
};

static void fill_frame(const ref_feat_opts *o, FrameExtractionOptions *f) {
  f->samp_freq = o->samp_freq; f->frame_shift_ms = o->frame_shift_ms;
  f->frame_length_ms = o->frame_length_ms; f->dither = o->dither;
  f->preemph_coeff = o->preemph_coeff; f->remove_dc_offset = o->remove_dc_offset != 0;
  static const char *kWin[] = {"povey", "hamming", "hanning", "rectangular"};
  f->window_type = kWin[o->window_type & 3]; f->round_to_power_of_two = o->round_to_power_of_two != 0;
  f->snip_edges = o->snip_edges != 0;
}
static MfccOptions mfcc_opts(const ref_feat_opts *o) {
  MfccOptions m;
  fill_frame(o, &m.frame_opts);
  m.mel_opts.num_bins = o->num_bins; m.mel_opts.low_freq = o->low_freq; m.mel_opts.high_freq = o->high_freq;
  m.mel_opts.htk_mode = o->htk_mode != 0;
  m.num_ceps = o->num_ceps; m.use_energy = o->use_energy != 0; m.energy_floor = o->energy_floor;
  m.raw_energy = o->raw_energy != 0; m.cepstral_lifter = o->cepstral_lifter; m.htk_compat = o->htk_compat != 0;
  return m;
}
static FbankOptions fbank_opts(const ref_feat_opts *o) {
  FbankOptions m;
  fill_frame(o, &m.frame_opts);
  m.mel_opts.num_bins = o->num_bins; m.mel_opts.low_freq = o->low_freq; m.mel_opts.high_freq = o->high_freq;
  m.mel_opts.htk_mode = o->htk_mode != 0;
  m.use_energy = o->use_energy != 0; m.energy_floor = o->energy_floor; m.raw_energy = o->raw_energy != 0;
  m.htk_compat = o->htk_compat != 0; m.use_log_fbank = o->use_log_fbank != 0; m.use_power = o->use_power != 0;
  return m;
}

static PlpOptions plp_opts(const ref_feat_opts *o) {
  PlpOptions m;
  fill_frame(o, &m.frame_opts);
  m.mel_opts.num_bins = o->num_bins; m.mel_opts.low_freq = o->low_freq; m.mel_opts.high_freq = o->high_freq;
  m.mel_opts.htk_mode = o->htk_mode != 0;
  m.lpc_order = o->lpc_order; m.num_ceps = o->num_ceps; m.use_energy = o->use_energy != 0; m.energy_floor = o->energy_floor;
  m.raw_energy = o->raw_energy != 0; m.compress_factor = o->compress_factor; m.cepstral_lifter = (int32)o->cepstral_lifter;
  m.cepstral_scale = o->cepstral_scale; m.htk_compat = o->htk_compat != 0;
  return m;
}

// offline: whole-utterance features.  returns number of frames (or -1), writes dim.
int ref_feat_compute(const ref_feat_opts *o, const float *wave, int n, float *out, int max_rows, int *dim) {
  try {
    SubVector<BaseFloat> w(const_cast<float *>(wave), n);
    Matrix<BaseFloat> feats;
    if (o->feature_type == 0) { Mfcc m(mfcc_opts(o)); m.ComputeFeatures(w, o->samp_freq, 1.0, &feats); }
    else if (o->feature_type == 2) { Plp m(plp_opts(o)); m.ComputeFeatures(w, o->samp_freq, 1.0, &feats); }
    else { Fbank m(fbank_opts(o)); m.ComputeFeatures(w, o->samp_freq, 1.0, &feats); }
    *dim = feats.NumCols();
    if (feats.NumRows() > max_rows) return -2;
    for (int r = 0; r < feats.NumRows(); r++) memcpy(out + (size_t)r * feats.NumCols(), feats.RowData(r), 4 * feats.NumCols());
    return feats.NumRows();
  } catch (...) { return -1; }
}

// online: feed `chunk`-sample pieces through OnlineMfcc/OnlineFbank (the path
// OnlineNnet2FeaturePipeline uses), InputFinished at the end.
int ref_feat_online(const ref_feat_opts *o, const float *wave, int n, int chunk, float *out, int max_rows, int *dim) {
  try {
    std::unique_ptr<OnlineBaseFeature> f;
    if (o->feature_type == 0) f.reset(new OnlineMfcc(mfcc_opts(o)));
    else if (o->feature_type == 2) f.reset(new OnlinePlp(plp_opts(o)));
    else f.reset(new OnlineFbank(fbank_opts(o)));
    for (int off = 0; off < n; off += chunk) {
      int len = std::min(chunk, n - off);
      SubVector<BaseFloat> w(const_cast<float *>(wave) + off, len);
      f->AcceptWaveform(o->samp_freq, w);
    }
    f->InputFinished();
    int T = f->NumFramesReady(), D = f->Dim();
    *dim = D;
    if (T > max_rows) return -2;
    for (int t = 0; t < T; t++) { SubVector<BaseFloat> row(out + (size_t)t * D, D); f->GetFrame(t, &row); }
    return T;
  } catch (...) { return -1; }
}

// OnlineCmvn over precomputed features (OnlineMatrixFeature source).
// global_stats: [2 x (D+1)] double (may be NULL -> zero stats); frames are
// requested in the order given by `order` (NULL = 0..T-1) to exercise the
// cached-stats paths (online-feature.cc:337-368).
int ref_cmvn_online(const float *feats, int T, int D, int cmn_window, int speaker_frames, int global_frames,
                    int normalize_mean, int normalize_variance, const double *global_stats,
                    const int *order, int n_order, float *out) {
  try {
    Matrix<BaseFloat> m(T, D);
    for (int t = 0; t < T; t++) memcpy(m.RowData(t), feats + (size_t)t * D, 4 * D);
    OnlineMatrixFeature src(m);
    OnlineCmvnOptions opts;
    opts.cmn_window = cmn_window; opts.speaker_frames = speaker_frames; opts.global_frames = global_frames;
    opts.normalize_mean = normalize_mean != 0; opts.normalize_variance = normalize_variance != 0;
    Matrix<double> g(2, D + 1);
    if (global_stats) for (int r = 0; r < 2; r++) for (int c = 0; c <= D; c++) g(r, c) = global_stats[r * (D + 1) + c];
    OnlineCmvnState state(g);
    OnlineCmvn cmvn(opts, state, &src);
    int n = order ? n_order : T;
    for (int i = 0; i < n; i++) {
      int t = order ? order[i] : i;
      SubVector<BaseFloat> row(out + (size_t)i * D, D);
      cmvn.GetFrame(t, &row);
    }
    return n;
  } catch (...) { return -1; }
}

// MfccOptions / FbankOptions as the reference's own ParseOptions::ReadConfigFile fills them from an option file
// (util/parse-options.cc:460-497): the oracle of b2k_feat_cfg_from_conf.  out: samp_freq, frame_shift_ms, frame_length_ms,
// dither, preemph, remove_dc, round_pow2, snip_edges, num_bins, low_freq, high_freq, num_ceps, use_energy, energy_floor,
// raw_energy, cepstral_lifter, htk_compat, use_log_fbank, use_power.  Returns 0, or 1 if the reference rejected the file.
int ref_feat_opts_from_conf(const char *path, int feature_type, float *out, char *window, int window_cap) {
  try {
    kaldi::ParseOptions po("");
    kaldi::MfccOptions m;
    kaldi::FbankOptions f;
    if (feature_type == 2) {          // PlpOptions: the same slots, then lpc_order, compress_factor, cepstral_scale
      kaldi::PlpOptions q;
      q.Register(&po);
      po.ReadConfigFile(path);
      const kaldi::FrameExtractionOptions &fr = q.frame_opts;
      int i = 0;
      out[i++] = fr.samp_freq; out[i++] = fr.frame_shift_ms; out[i++] = fr.frame_length_ms; out[i++] = fr.dither; out[i++] = fr.preemph_coeff;
      out[i++] = fr.remove_dc_offset; out[i++] = fr.round_to_power_of_two; out[i++] = fr.snip_edges;
      out[i++] = q.mel_opts.num_bins; out[i++] = q.mel_opts.low_freq; out[i++] = q.mel_opts.high_freq;
      out[i++] = q.num_ceps; out[i++] = q.use_energy; out[i++] = q.energy_floor; out[i++] = q.raw_energy; out[i++] = q.cepstral_lifter;
      out[i++] = q.htk_compat; out[i++] = 1; out[i++] = 1;
      out[i++] = q.lpc_order; out[i++] = q.compress_factor; out[i++] = q.cepstral_scale;
      snprintf(window, window_cap, "%s", fr.window_type.c_str());
      return 0;
    }
    if (feature_type == 0) m.Register(&po); else f.Register(&po);
    po.ReadConfigFile(path);
    const kaldi::FrameExtractionOptions &fr = feature_type == 0 ? m.frame_opts : f.frame_opts;
    const kaldi::MelBanksOptions &mel = feature_type == 0 ? m.mel_opts : f.mel_opts;
    int i = 0;
    out[i++] = fr.samp_freq; out[i++] = fr.frame_shift_ms; out[i++] = fr.frame_length_ms; out[i++] = fr.dither; out[i++] = fr.preemph_coeff;
    out[i++] = fr.remove_dc_offset; out[i++] = fr.round_to_power_of_two; out[i++] = fr.snip_edges;
    out[i++] = mel.num_bins; out[i++] = mel.low_freq; out[i++] = mel.high_freq;
    out[i++] = feature_type == 0 ? m.num_ceps : 0; out[i++] = feature_type == 0 ? m.use_energy : f.use_energy;
    out[i++] = feature_type == 0 ? m.energy_floor : f.energy_floor; out[i++] = feature_type == 0 ? m.raw_energy : f.raw_energy;
    out[i++] = feature_type == 0 ? m.cepstral_lifter : 0; out[i++] = feature_type == 0 ? m.htk_compat : f.htk_compat;
    out[i++] = feature_type == 0 ? 1 : f.use_log_fbank; out[i++] = feature_type == 0 ? 1 : f.use_power;
    snprintf(window, window_cap, "%s", fr.window_type.c_str());
    return 0;
  } catch (const std::exception &) {
    return 1;
  }
}

// WaveData::Read (feat/wave-reader.cc) on a file: the oracle of b2k_wave_read.  Returns 0, 1 if the reference rejected the
// file, 2 if `cap` floats are not enough.
int ref_wave_read(const char *path, float *out, long long cap, int *channels, long long *samples, float *samp_freq) {
  try {
    std::ifstream is(path, std::ios::binary);
    if (!is.good()) return 1;
    kaldi::WaveData w;
    w.Read(is);
    *channels = w.Data().NumRows(); *samples = w.Data().NumCols(); *samp_freq = w.SampFreq();
    if ((long long)w.Data().NumRows() * w.Data().NumCols() > cap) return 2;
    for (int j = 0; j < w.Data().NumRows(); j++)
      for (int i = 0; i < w.Data().NumCols(); i++) out[(long long)j * w.Data().NumCols() + i] = w.Data()(j, i);
    return 0;
  } catch (const std::exception &) {
    return 1;
  }
}

// ResampleWaveform (feat/resample.cc:368): the oracle of b2k_resample_waveform.  Returns the number of output samples (-1 on error).
long long ref_resample_waveform(float orig_freq, const float *in, long long n, float new_freq, float *out, long long cap) {
  try {
    kaldi::Vector<kaldi::BaseFloat> w(n), o;
    for (long long i = 0; i < n; i++) w(i) = in[i];
    kaldi::ResampleWaveform(orig_freq, w, new_freq, &o);
    for (long long i = 0; i < o.Dim() && i < cap; i++) out[i] = o(i);
    return o.Dim();
  } catch (const std::exception &) {
    return -1;
  }
}

}  // extern "C"
