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
#include "b2k_silence_weighting.h"
#include "cudamatrix/cu-matrix.h"
#include "cudamatrix/cu-vector.h"
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
    b2k_feat_cfg c;                                        // (pitch, when asked for, is appended on the host by the pipeline below)
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
    if (info.add_pitch) {                                           // :111-116: the reference's own pitch objects, on the host
      pitch_.reset(new OnlinePitchFeature(info.pitch_opts));        // (b2k has no pitch kernel; the base features come from the device)
      pitch_feature_.reset(new OnlineProcessPitch(info.pitch_process_opts, pitch_.get()));
      plus_pitch_.reset(new OnlineAppendFeature(base_.get(), pitch_feature_.get()));
      top = plus_pitch_.get();
    }
    if (info.use_cmvn) {                                            // online-nnet2-feature-pipeline.cc:121-131
      if (info.global_cmvn_stats.NumCols() == 0)
        KALDI_ERR << "global_cmvn_stats for OnlineCmvn must be non-empty, please assign it to OnlineNnet2FeaturePipelineInfo.";
      cmvn_.reset(new OnlineCmvn(info.cmvn_opts, OnlineCmvnState(info.global_cmvn_stats), top));
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

  void AcceptWaveform(BaseFloat sampling_rate, const VectorBase<BaseFloat> &waveform) {          // :219-225
    base_->AcceptWaveform(sampling_rate, waveform);
    if (pitch_) pitch_->AcceptWaveform(sampling_rate, waveform);
  }
  void InputFinished() {                                                                          // :227-231
    base_->InputFinished();
    if (pitch_) pitch_->InputFinished();
  }

  OnlineIvectorFeature *IvectorFeature() { return ivector_.get(); }
  const OnlineIvectorFeature *IvectorFeature() const { return ivector_.get(); }
  OnlineFeatureInterface *InputFeature() { return input_; }

 private:
  const OnlineNnet2FeaturePipelineInfo &info_;
  // destruction runs bottom-up in reverse order of declaration: the stages on top go first
  std::unique_ptr<OnlineBaseFeatureB2k> base_;
  std::unique_ptr<OnlinePitchFeature> pitch_;
  std::unique_ptr<OnlineProcessPitch> pitch_feature_;
  std::unique_ptr<OnlineAppendFeature> plus_pitch_;
  std::unique_ptr<OnlineCmvn> cmvn_;
  std::unique_ptr<OnlineIvectorFeature> ivector_;
  std::unique_ptr<OnlineAppendFeature> append_;
  OnlineFeatureInterface *input_ = NULL, *final_ = NULL;
  KALDI_DISALLOW_COPY_AND_ASSIGN(OnlineNnet2FeaturePipelineB2k);
};

// OnlineBatchedFeaturePipelineCuda (cudafeat/online-batched-feature-pipeline-cuda.h:44-134): one chunk of samples per lane and
// call, the frames that became computable land in rows [lane * GetMaxChunkFrames(), ...) of `input_features`, online CMVN applied
// when the configuration asks for it.  The reference stashes the samples a frame still needs between calls; here a channel's
// samples and raw features so far stay in device memory (the feature kernel reads frames that straddle two chunks from there,
// the CMVN kernel its sliding window), bounded by max_seconds per utterance.  The streaming i-vector of the reference's class
// (BatchedIvectorExtractorCuda) has no counterpart: a non-null ivector_features with i-vectors configured is refused.
class OnlineBatchedFeaturePipelineB2k {
 public:
  typedef int32 ChannelId;
  OnlineBatchedFeaturePipelineB2k(const OnlineNnet2FeaturePipelineConfig &config, int32_t max_chunk_size_samples, int32_t max_lanes,
                                  int32_t num_channels, BaseFloat max_seconds = 120.0f)
      : info_(config), tables_(info_, max_lanes), max_chunk_size_samples_(max_chunk_size_samples), max_lanes_(max_lanes),
        num_channels_(num_channels) {
    if (info_.add_pitch) KALDI_ERR << "b2k has no pitch kernel (--add-pitch); the single-utterance pipeline appends pitch on the host";
    if (info_.feature_type == "mfcc") frame_opts_ = info_.mfcc_opts.frame_opts;
    else if (info_.feature_type == "fbank") frame_opts_ = info_.fbank_opts.frame_opts;
    else frame_opts_ = info_.plp_opts.frame_opts;
    const int32_t shift = frame_opts_.WindowShift();
    max_chunk_size_frames_ = (max_chunk_size_samples_ + shift - 1) / shift;                  // :63-64
    dim_ = b2k_feat_dim(tables_.Handle());
    max_samples_ = static_cast<int32_t>(max_seconds * frame_opts_.samp_freq);
    max_frames_ = b2k_feat_num_frames(tables_.Handle(), max_samples_, 1) + 1;
    const size_t sd = 2 * (static_cast<size_t>(dim_) + 1);
    if (cudaMalloc(&d_wave_, sizeof(float) * static_cast<size_t>(num_channels) * max_samples_) != cudaSuccess ||
        cudaMalloc(&d_raw_, sizeof(float) * static_cast<size_t>(num_channels) * max_frames_ * dim_) != cudaSuccess ||
        cudaMalloc(&d_cmvn_state_, sizeof(double) * static_cast<size_t>(num_channels) * sd) != cudaSuccess ||
        cudaMalloc(&d_global_, sizeof(double) * sd) != cudaSuccess)
      KALDI_ERR << "cudaMalloc failed";
    if (info_.use_cmvn) {
      if (info_.global_cmvn_stats.NumCols() == 0) KALDI_ERR << "global_cmvn_stats for OnlineCmvn must be non-empty.";   // :78-80
      std::vector<double> g(sd);
      for (int32 r = 0; r < 2; r++)
        for (int32 c = 0; c <= dim_; c++) g[r * (dim_ + 1) + c] = info_.global_cmvn_stats(r, c);
      cudaMemcpy(d_global_, g.data(), sizeof(double) * sd, cudaMemcpyHostToDevice);
      cmvn_cfg_.cmn_window = info_.cmvn_opts.cmn_window; cmvn_cfg_.speaker_frames = info_.cmvn_opts.speaker_frames;
      cmvn_cfg_.global_frames = info_.cmvn_opts.global_frames; cmvn_cfg_.normalize_mean = info_.cmvn_opts.normalize_mean;
      cmvn_cfg_.normalize_variance = info_.cmvn_opts.normalize_variance;
    }
    samples_.assign(num_channels, 0);
  }
  ~OnlineBatchedFeaturePipelineB2k() { cudaFree(d_wave_); cudaFree(d_raw_); cudaFree(d_cmvn_state_); cudaFree(d_global_); }

  void ComputeFeaturesBatched(int32_t num_lanes, const std::vector<ChannelId> &channels, const std::vector<int32_t> &num_chunk_samples,
                              const std::vector<bool> &first, const std::vector<bool> &last, BaseFloat sample_freq,
                              const CuMatrixBase<BaseFloat> &cu_waves, CuMatrix<BaseFloat> *input_features,
                              CuVector<BaseFloat> *ivector_features, std::vector<int32_t> *num_frames_computed) {
    KALDI_ASSERT(num_lanes <= max_lanes_);                                                   // :135-136
    KALDI_ASSERT(num_lanes <= static_cast<int32_t>(num_frames_computed->size()));
    if (sample_freq != frame_opts_.samp_freq) KALDI_ERR << "Sampling frequency mismatch, expected " << frame_opts_.samp_freq << ", got " << sample_freq;
    if (info_.use_ivectors && ivector_features != NULL) KALDI_ERR << "b2k computes i-vectors per utterance (b2k_ivec_compute_batched), not per chunk";
    if (input_features->NumRows() < num_lanes * max_chunk_size_frames_ || input_features->NumCols() != dim_)
      input_features->Resize(max_lanes_ * max_chunk_size_frames_, dim_, kUndefined);
    const size_t sd = 2 * (static_cast<size_t>(dim_) + 1);
    wp_.resize(num_lanes); rawp_.resize(num_lanes); inp_.resize(num_lanes); outp_.resize(num_lanes); statep_.resize(num_lanes);
    ns_.resize(num_lanes); ff_.resize(num_lanes); nf_.resize(num_lanes);
    for (int32_t lane = 0; lane < num_lanes; lane++) {
      const ChannelId ch = channels[lane];
      KALDI_ASSERT(ch >= 0 && ch < num_channels_);                                           // :145-146
      KALDI_ASSERT(num_chunk_samples[lane] <= max_chunk_size_samples_);
      const int32_t current_sample = first[lane] ? 0 : samples_[ch];                         // :153-156
      const int32_t current_frame = b2k_feat_num_frames(tables_.Handle(), current_sample, 0);
      const int32_t num_samples = current_sample + num_chunk_samples[lane];
      if (num_samples > max_samples_) KALDI_ERR << "utterance longer than the configured capacity";
      const int32_t num_frames = b2k_feat_num_frames(tables_.Handle(), num_samples, last[lane] ? 1 : 0);
      float *w = d_wave_ + static_cast<size_t>(ch) * max_samples_;
      if (num_chunk_samples[lane] > 0 &&
          cudaMemcpyAsync(w + current_sample, cu_waves.Data() + static_cast<size_t>(lane) * cu_waves.Stride(),
                          sizeof(float) * num_chunk_samples[lane], cudaMemcpyDeviceToDevice, cudaStreamPerThread) != cudaSuccess)
        KALDI_ERR << "copying the chunk behind the channel's samples failed";
      if (first[lane]) cudaMemsetAsync(d_cmvn_state_ + static_cast<size_t>(ch) * sd, 0, sizeof(double) * sd, cudaStreamPerThread);
      samples_[ch] = num_samples;                                                            // :173-174
      (*num_frames_computed)[lane] = num_frames - current_frame;                             // :177
      float *raw = d_raw_ + static_cast<size_t>(ch) * max_frames_ * dim_;
      wp_[lane] = w; rawp_[lane] = raw; ns_[lane] = num_samples; ff_[lane] = current_frame; nf_[lane] = num_frames - current_frame;
      // frame f of the channel goes to row lane * max_chunk_frames + (f - current_frame): the kernels address frames absolutely
      inp_[lane] = raw;
      outp_[lane] = input_features->Data() + (static_cast<ptrdiff_t>(lane) * max_chunk_size_frames_ - current_frame) * input_features->Stride();
      statep_[lane] = d_cmvn_state_ + static_cast<size_t>(ch) * sd;
    }
    Check(b2k_feat_compute_batched(tables_.Handle(), num_lanes, wp_.data(), ns_.data(), ff_.data(), nf_.data(), rawp_.data(), dim_,
                                   cudaStreamPerThread), "b2k_feat_compute_batched");
    if (info_.use_cmvn) {
      Check(b2k_cmvn_apply_batched(tables_.Handle(), &cmvn_cfg_, num_lanes, inp_.data(), outp_.data(), dim_, input_features->Stride(),
                                   ff_.data(), nf_.data(), statep_.data(), d_global_, NULL, cudaStreamPerThread), "b2k_cmvn_apply_batched");
    } else {
      for (int32_t lane = 0; lane < num_lanes; lane++)
        if (nf_[lane] > 0 &&
            cudaMemcpy2DAsync(input_features->Data() + static_cast<size_t>(lane) * max_chunk_size_frames_ * input_features->Stride(),
                              sizeof(float) * input_features->Stride(), rawp_[lane] + static_cast<size_t>(ff_[lane]) * dim_, sizeof(float) * dim_,
                              sizeof(float) * dim_, nf_[lane], cudaMemcpyDeviceToDevice, cudaStreamPerThread) != cudaSuccess)
          KALDI_ERR << "copying the new feature frames failed";
    }
  }

  int32_t GetMaxChunkFrames() { return max_chunk_size_frames_; }
  int32_t FeatureDim() { return dim_; }
  int32_t IvectorDim() { return 0; }
  const FrameExtractionOptions &GetFrameOptions() { return frame_opts_; }

 private:
  OnlineNnet2FeaturePipelineInfo info_;
  FeatureTablesB2k tables_;
  FrameExtractionOptions frame_opts_;
  int32_t max_chunk_size_samples_, max_chunk_size_frames_ = 0, max_lanes_, num_channels_, dim_ = 0, max_samples_ = 0, max_frames_ = 0;
  float *d_wave_ = NULL, *d_raw_ = NULL;
  double *d_cmvn_state_ = NULL, *d_global_ = NULL;
  b2k_cmvn_cfg cmvn_cfg_;
  std::vector<int32_t> samples_, ns_, ff_, nf_;
  std::vector<const float *> wp_, inp_;
  std::vector<float *> rawp_, outp_;
  std::vector<double *> statep_;
  KALDI_DISALLOW_COPY_AND_ASSIGN(OnlineBatchedFeaturePipelineB2k);
};

// OnlineSilenceWeighting (online2/online-ivector-feature.h:460-571) over a b2k decoder.  The loop of the online2 tools stays
// as it is (online2bin/online2-wav-nnet3-latgen-faster.cc:254-262):
//     if (silence_weighting.Active() && feature_pipeline.IvectorFeature() != NULL) {
//       silence_weighting.ComputeCurrentTraceback(decoder);
//       silence_weighting.GetDeltaWeights(feature_pipeline.NumFramesReady(), frame_offset * subsampling, &delta_weights);
//       feature_pipeline.UpdateFrameWeights(delta_weights);
//     }
// The reference's ComputeCurrentTraceback is a template over its own decoders (instantiated in the .cc for three FST types);
// this one takes any decoder with NumFramesDecoded() and BestPath(use_final_probs, &ilabels, ..., &info, &arc_states):
// SingleUtteranceNnet3DecoderB2k (b2k_nnet3_shims.h).  The bookkeeping is b2k_host::SilenceWeighting, pinned to the
// reference's class call after call (tests/test_silence_weighting.py).
class OnlineSilenceWeightingB2k {
 public:
  OnlineSilenceWeightingB2k(const TransitionModel &trans_model, const OnlineSilenceWeightingConfig &config,
                            int32 frame_subsampling_factor = 1)
      : config_(config), core_(TidToPhone(trans_model), config.silence_phones_str, config.silence_weight, config.max_state_duration,
                               frame_subsampling_factor) {
    KALDI_ASSERT(frame_subsampling_factor >= 1);
    if (!core_.SilencePhonesParsed())        // the reference goes on with an empty list without a word
      KALDI_WARN << "--silence-phones=" << config.silence_phones_str << " is not a list of integers: no phone counts as silence";
  }

  bool Active() const { return config_.Active(); }

  template <class Decoder>
  void ComputeCurrentTraceback(Decoder &decoder, bool use_final_probs = false) {
    const int32 num_frames_decoded = decoder.NumFramesDecoded();
    ilabels_.clear(); arc_states_.clear();
    if (num_frames_decoded > 0) {
      b2k_best_path_info info;
      decoder.BestPath(use_final_probs, &ilabels_, NULL, NULL, NULL, &info, &arc_states_);
    }
    try {
      if (!core_.SetTracebackFromPath(ilabels_.data(), arc_states_.data(), static_cast<int32>(ilabels_.size()), num_frames_decoded))
        KALDI_ERR << "the best path does not hold one transition-id per decoded frame (" << num_frames_decoded << " frames)";
    } catch (const std::runtime_error &e) { KALDI_ERR << e.what(); }
  }

  void GetDeltaWeights(int32 num_frames_ready, int32 first_decoder_frame, std::vector<std::pair<int32, BaseFloat> > *delta_weights) {
    KALDI_ASSERT(num_frames_ready > first_decoder_frame || num_frames_ready == 0);
    core_.GetDeltaWeights(num_frames_ready, first_decoder_frame, delta_weights);
  }
  void GetDeltaWeights(int32 num_frames_ready, std::vector<std::pair<int32, BaseFloat> > *delta_weights) {
    GetDeltaWeights(num_frames_ready, 0, delta_weights);
  }
  void GetNonsilenceFrames(int32 num_frames_ready, int32 first_decoder_frame, std::vector<int32> *frames) {
    KALDI_ASSERT(num_frames_ready > first_decoder_frame || num_frames_ready == 0);
    core_.GetNonsilenceFrames(num_frames_ready, first_decoder_frame, frames);
  }

 private:
  static std::vector<int32_t> TidToPhone(const TransitionModel &trans_model) {
    std::vector<int32_t> t(trans_model.NumTransitionIds() + 1, 0);
    for (int32 tid = 1; tid <= trans_model.NumTransitionIds(); tid++) t[tid] = trans_model.TransitionIdToPhone(tid);
    return t;
  }
  const OnlineSilenceWeightingConfig &config_;
  b2k_host::SilenceWeighting core_;
  std::vector<int32> ilabels_, arc_states_;
};

}  // namespace b2k_shim
}  // namespace kaldi

#endif  // B2K_ONLINE2_SHIMS_H_
