// b2k_online2_shims.h — the reference's online2 feature pipeline with the base features on the GPU:
//
//   OnlineNnet2FeaturePipeline                        online2/online-nnet2-feature-pipeline.h:200-327
//   (built from the reference's OWN OnlineNnet2FeaturePipelineConfig / ...Info, :69-195, unchanged: Register(), the
//    config files and their checks stay the reference's code)
//
// Only the base feature object (OnlineMfcc / OnlineFbank, online-nnet2-feature-pipeline.cc:100-108) is replaced -- by
// OnlineBaseFeatureB2k over b2k_feat_* -- and everything stacked on it is the reference's own OnlineCmvn /
// OnlineIvectorFeature / OnlineAppendFeature, so adaptation state, CMVN state and frame weights keep their meaning and their
// types.  Pitch has no kernel here and is refused.
//
// online2/online-ivector-feature.h includes the OpenFst-based decoders; a build without OpenFst (this repository's check,
// oracle/check_shims.py) pre-defines their include guards and forward-declares the two decoder templates.
#ifndef B2K_ONLINE2_SHIMS_H_
#define B2K_ONLINE2_SHIMS_H_

#include <memory>
#include <string>
#include <utility>
#include <vector>

#include "b2k.h"
#include "b2k_kaldi_shims.h"
#include "feat/feature-fbank.h"
#include "feat/feature-mfcc.h"
#include "feat/feature-plp.h"
#include "feat/online-feature.h"
#include "online2/online-ivector-feature.h"
#include "online2/online-nnet2-feature-pipeline.h"

namespace kaldi {
namespace b2k_shim {

inline int32_t WindowTypeB2k(const std::string &w) {
  if (w == "povey") return 0;
  if (w == "hamming") return 1;
  if (w == "hanning") return 2;
  if (w == "rectangular") return 3;
  KALDI_ERR << "b2k has no kernel for --window-type=" << w;
  return -1;
}

inline void FrameAndMelB2k(const FrameExtractionOptions &f, const MelBanksOptions &m, b2k_feat_cfg *c) {
  c->samp_freq = f.samp_freq; c->frame_shift_ms = f.frame_shift_ms; c->frame_length_ms = f.frame_length_ms;
  c->dither = f.dither; c->preemph_coeff = f.preemph_coeff; c->remove_dc_offset = f.remove_dc_offset;
  c->round_to_power_of_two = f.round_to_power_of_two; c->snip_edges = f.snip_edges; c->window_type = WindowTypeB2k(f.window_type);
  if (f.blackman_coeff != 0.42f && f.window_type == "blackman") KALDI_ERR << "b2k has no blackman window";
  if (f.allow_downsample || f.allow_upsample) KALDI_WARN << "resampling happens before b2k (b2k_resample_waveform), not inside it";
  c->num_bins = m.num_bins; c->low_freq = m.low_freq; c->high_freq = m.high_freq; c->htk_mode = m.htk_mode;
  if (m.vtln_low != 100.0f || m.vtln_high != -500.0f) KALDI_WARN << "b2k computes features without VTLN warping (warp factor 1.0)";
}

// MfccOptions / FbankOptions -> the union b2k_feat_create takes (include/b2k.h)
inline b2k_feat_cfg ToB2kFeatCfg(const MfccOptions &o, int32 max_lanes = 1) {
  b2k_feat_cfg c;
  b2k_feat_cfg_default(&c);
  c.feature_type = 0;
  FrameAndMelB2k(o.frame_opts, o.mel_opts, &c);
  c.num_ceps = o.num_ceps; c.use_energy = o.use_energy; c.energy_floor = o.energy_floor; c.raw_energy = o.raw_energy;
  c.cepstral_lifter = o.cepstral_lifter; c.htk_compat = o.htk_compat;
  c.max_lanes = max_lanes;
  return c;
}
inline b2k_feat_cfg ToB2kFeatCfg(const FbankOptions &o, int32 max_lanes = 1) {
  b2k_feat_cfg c;
  b2k_feat_cfg_default(&c);
  c.feature_type = 1;
  FrameAndMelB2k(o.frame_opts, o.mel_opts, &c);
  c.use_energy = o.use_energy; c.energy_floor = o.energy_floor; c.raw_energy = o.raw_energy; c.htk_compat = o.htk_compat;
  c.use_log_fbank = o.use_log_fbank; c.use_power = o.use_power;
  c.max_lanes = max_lanes;
  return c;
}

inline b2k_feat_cfg ToB2kFeatCfg(const PlpOptions &o, int32 max_lanes = 1) {
  b2k_feat_cfg c;
  b2k_feat_cfg_default(&c);
  c.feature_type = 2;
  FrameAndMelB2k(o.frame_opts, o.mel_opts, &c);
  c.lpc_order = o.lpc_order; c.num_ceps = o.num_ceps; c.use_energy = o.use_energy; c.energy_floor = o.energy_floor;
  c.raw_energy = o.raw_energy; c.compress_factor = o.compress_factor; c.cepstral_lifter = static_cast<float>(o.cepstral_lifter);
  c.cepstral_scale = o.cepstral_scale; c.htk_compat = o.htk_compat;
  c.max_lanes = max_lanes;
  return c;
}

// One per OnlineNnet2FeaturePipelineInfo: the device tables (window, mel banks, DCT) every utterance's pipeline shares.
class FeatureTablesB2k {
 public:
  explicit FeatureTablesB2k(const OnlineNnet2FeaturePipelineInfo &info, int32 max_lanes = 1) {
    if (info.add_pitch) KALDI_ERR << "b2k has no pitch kernel (--add-pitch)";
    b2k_feat_cfg c;
    if (info.feature_type == "mfcc") c = ToB2kFeatCfg(info.mfcc_opts, max_lanes);
    else if (info.feature_type == "fbank") c = ToB2kFeatCfg(info.fbank_opts, max_lanes);
    else if (info.feature_type == "plp") c = ToB2kFeatCfg(info.plp_opts, max_lanes);
    else KALDI_ERR << "b2k computes mfcc, fbank and plp features, not " << info.feature_type;
    Check(b2k_feat_create(&c, &feat_), "b2k_feat_create");
  }
  ~FeatureTablesB2k() { b2k_feat_destroy(feat_); }
  b2k_feat *Handle() const { return feat_; }

 private:
  b2k_feat *feat_ = NULL;
  KALDI_DISALLOW_COPY_AND_ASSIGN(FeatureTablesB2k);
};

// OnlineNnet2FeaturePipeline (online2/online-nnet2-feature-pipeline.h:200-327), one object per utterance.
class OnlineNnet2FeaturePipelineB2k : public OnlineFeatureInterface {
 public:
  // max_seconds bounds the device buffers of the utterance (the reference grows its host buffers instead)
  OnlineNnet2FeaturePipelineB2k(const OnlineNnet2FeaturePipelineInfo &info, const FeatureTablesB2k &tables, BaseFloat max_seconds = 120.0f)
      : info_(info) {
    base_.reset(new OnlineBaseFeatureB2k(tables.Handle(), info.FrameShiftInSeconds(),
                                         static_cast<int32>(max_seconds * b2k_feat_samp_freq(tables.Handle()))));
    OnlineFeatureInterface *top = base_.get();
    if (info.use_cmvn) {                                            // online-nnet2-feature-pipeline.cc:121-131
      if (info.global_cmvn_stats.NumCols() == 0)
        KALDI_ERR << "global_cmvn_stats for OnlineCmvn must be non-empty, please assign it to OnlineNnet2FeaturePipelineInfo.";
      cmvn_.reset(new OnlineCmvn(info.cmvn_opts, OnlineCmvnState(info.global_cmvn_stats), base_.get()));
      top = cmvn_.get();
    }
    input_ = top;                                                   // what the network reads as "input"
    if (info.use_ivectors) {                                        // the extractor reads the features WITHOUT the cmvn (:137-139)
      ivector_.reset(new OnlineIvectorFeature(info.ivector_extractor_info, base_.get()));
      append_.reset(new OnlineAppendFeature(top, ivector_.get()));
      top = append_.get();
    }
    final_ = top;
  }

  int32 Dim() const override { return final_->Dim(); }
  bool IsLastFrame(int32 frame) const override { return final_->IsLastFrame(frame); }
  int32 NumFramesReady() const override { return final_->NumFramesReady(); }
  void GetFrame(int32 frame, VectorBase<BaseFloat> *feat) override { final_->GetFrame(frame, feat); }
  BaseFloat FrameShiftInSeconds() const override { return info_.FrameShiftInSeconds(); }

  void UpdateFrameWeights(const std::vector<std::pair<int32, BaseFloat> > &delta_weights) {
    IvectorFeature()->UpdateFrameWeights(delta_weights);
  }
  void SetAdaptationState(const OnlineIvectorExtractorAdaptationState &adaptation_state) {
    if (ivector_) ivector_->SetAdaptationState(adaptation_state);
  }
  void GetAdaptationState(OnlineIvectorExtractorAdaptationState *adaptation_state) const {
    if (ivector_) ivector_->GetAdaptationState(adaptation_state);
  }
  void SetCmvnState(const OnlineCmvnState &cmvn_state) { if (cmvn_) cmvn_->SetState(cmvn_state); }
  void GetCmvnState(OnlineCmvnState *cmvn_state) { if (cmvn_) cmvn_->GetState(cmvn_->NumFramesReady() - 1, cmvn_state); }

  void AcceptWaveform(BaseFloat sampling_rate, const VectorBase<BaseFloat> &waveform) { base_->AcceptWaveform(sampling_rate, waveform); }
  void InputFinished() { base_->InputFinished(); }

  OnlineIvectorFeature *IvectorFeature() { return ivector_.get(); }
  const OnlineIvectorFeature *IvectorFeature() const { return ivector_.get(); }
  OnlineFeatureInterface *InputFeature() { return input_; }

 private:
  const OnlineNnet2FeaturePipelineInfo &info_;
  // destruction runs bottom-up in reverse order of declaration: the stages on top go first
  std::unique_ptr<OnlineBaseFeatureB2k> base_;
  std::unique_ptr<OnlineCmvn> cmvn_;
  std::unique_ptr<OnlineIvectorFeature> ivector_;
  std::unique_ptr<OnlineAppendFeature> append_;
  OnlineFeatureInterface *input_ = NULL, *final_ = NULL;
  KALDI_DISALLOW_COPY_AND_ASSIGN(OnlineNnet2FeaturePipelineB2k);
};

}  // namespace b2k_shim
}  // namespace kaldi

#endif  // B2K_ONLINE2_SHIMS_H_
