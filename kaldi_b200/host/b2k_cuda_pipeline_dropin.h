// b2k_cuda_pipeline_dropin.h -- build cudadecoderbin/batched-wav-nnet3-cuda-online.cc against b2k WITHOUT editing it.
//
//   g++ ... -include b2k_cuda_pipeline_dropin.h cudadecoderbin/batched-wav-nnet3-cuda-online.cc ... -lb2k
//
// The tool drives two classes: cuda_decoder::BatchedThreadedNnet3CudaOnlinePipeline (cudadecoder/batched-threaded-nnet3-cuda-
// online-pipeline.h:119-330) and cuda_decoder::CudaOnlinePipelineDynamicBatcher (cuda-online-pipeline-dynamic-batcher.h:38-60).
// This header includes the reference's own headers first (so that the tool's later #includes are no-ops and the option
// structs, CudaPipelineResult, SegmentedLatticeCallbackParams, LatticePostprocessor stay the reference's types), then lets the
// two names resolve to adapters over the b2k streaming pipeline (b2k_stream_*: features, chunked nnet3 with carried context,
// decoder, all on the device) with the constructors and members the tool calls:
//
//   BatchedThreadedNnet3CudaOnlinePipeline cuda_pipeline(opts.batched_decoder_config, *decode_fst, am_nnet, trans_model);  :88-89
//   cuda_pipeline.SetSymbolTable / SetLatticePostprocessor / GetNSampsPerChunk / GetSecondsPerChunk
//   cuda_pipeline.SetBestPathCallback(corr_id, (text, partial, endpoint) -> void)                                         :191
//   cuda_pipeline.SetLatticeCallback(corr_id, SegmentedResultsCallback, result_type)                                      :246
//   CudaOnlinePipelineDynamicBatcher dynamic_batcher(dynamic_batcher_config, cuda_pipeline);  Push / WaitForCompletion    :142
//
// What the adapter does with the reference's configuration struct: feature_opts -> OnlineNnet2FeaturePipelineInfo -> the b2k
// feature tables (mfcc / fbank / plp; i-vectors are refused: the streaming pipeline has no per-chunk i-vector stage);
// decoder_opts (beam, lattice beam, max-active, queue capacities) -> b2k_dec_cfg; compute_opts (frames-per-chunk, acoustic
// scale, subsampling factor) -> the chunked executor; det_opts -> b2k_lat_determinize_phone_pruned at the decoder's lattice
// beam; num_channels / max_batch_size -> channels of the stream pipeline and the batcher's batch size.  A stream's lattice is
// determinized on the calling thread when its last chunk has been decoded (the reference hands that to a thread pool).
//
// Checked by oracle/check_shims.py: the tool's translation unit as it lies in the reference tree compiles with this header
// force-included (HAVE_CUDA=1, the container's OpenFst stand-in).  It cannot be RUN here (no CUDA build of Kaldi, no OpenFst).
#ifndef B2K_CUDA_PIPELINE_DROPIN_H_
#define B2K_CUDA_PIPELINE_DROPIN_H_

#include <functional>
#include <map>
#include <memory>
#include <random>                                    // the tool uses std::mt19937 and gets <random> through OpenFst
#include <string>
#include <utility>
#include <vector>

#include "cudadecoder/batched-threaded-nnet3-cuda-online-pipeline.h"
#include "cudadecoder/batched-threaded-nnet3-cuda-pipeline2.h"
#include "cudadecoder/cuda-online-pipeline-dynamic-batcher.h"
#include "cudadecoder/cuda-pipeline-common.h"
#include "cudadecoder/lattice-postprocessor.h"

#ifndef B2K_HAVE_OPENFST
#define B2K_HAVE_OPENFST
#endif
#include "b2k_kaldi_shims.h"
#include "b2k_nnet3_shims.h"
#include "b2k_online2_shims.h"

namespace kaldi {
namespace cuda_decoder {
namespace b2k_cuda_dropin {

class BatchedThreadedNnet3CudaOnlinePipeline {
 public:
  using CorrelationID = uint64_t;
  typedef std::function<void(const std::string &, bool, bool)> BestPathCallback;
  typedef std::function<void(CompactLattice &)> LatticeCallback;
  typedef kaldi::cuda_decoder::BatchedThreadedNnet3CudaOnlinePipelineConfig Config;

  BatchedThreadedNnet3CudaOnlinePipeline(const Config &config, const fst::Fst<fst::StdArc> &decode_fst,
                                         const nnet3::AmNnetSimple &am_nnet, const TransitionModel &trans_model)
      : config_(config), trans_model_(&trans_model), feature_info_(config.feature_opts) {
    config_.compute_opts.CheckAndFixConfigs(am_nnet.GetNnet().Modulus());           // …online-pipeline.h:151-152
    config_.CheckAndFixConfigs();
    if (feature_info_.use_ivectors)
      KALDI_ERR << "the b2k streaming pipeline has no per-chunk i-vector stage (--ivector-extraction-config)";
    if (feature_info_.add_pitch) KALDI_ERR << "b2k has no pitch kernel (--add-pitch)";
    if (!config_.determinize_lattice) KALDI_ERR << "--determinize-lattice=false is not supported";
    model_.reset(new b2k_shim::ModelB2k(am_nnet, config_.compute_opts.frame_subsampling_factor));
    graph_.reset(new b2k_shim::CudaFstB2k(decode_fst, &trans_model));

    b2k_stream_cfg c;
    b2k_stream_cfg_default(&c);
    if (feature_info_.feature_type == "mfcc") c.feat = b2k_shim::ToB2kFeatCfg(feature_info_.mfcc_opts, config_.max_batch_size);
    else if (feature_info_.feature_type == "fbank") c.feat = b2k_shim::ToB2kFeatCfg(feature_info_.fbank_opts, config_.max_batch_size);
    else if (feature_info_.feature_type == "plp") c.feat = b2k_shim::ToB2kFeatCfg(feature_info_.plp_opts, config_.max_batch_size);
    else KALDI_ERR << "b2k computes mfcc, fbank and plp features, not " << feature_info_.feature_type;
    model_frequency_ = c.feat.samp_freq;
    b2k_shim::CudaDecoderConfigB2k d;                  // the same option values, mapped as CudaDecoderB2k maps them
    d.default_beam = config_.decoder_opts.default_beam; d.lattice_beam = config_.decoder_opts.lattice_beam;
    d.max_active = config_.decoder_opts.max_active; d.ntokens_pre_allocated = config_.decoder_opts.ntokens_pre_allocated;
    d.main_q_capacity = config_.decoder_opts.main_q_capacity; d.aux_q_capacity = config_.decoder_opts.aux_q_capacity;
    c.dec = d.ToB2k(c.dec.max_frames);
    c.nchannels = config_.num_channels;
    c.frames_per_chunk = config_.compute_opts.frames_per_chunk;
    c.acoustic_scale = config_.compute_opts.acoustic_scale;
    c.use_priors = 1;
    pipeline_.reset(new b2k_shim::StreamingOnlinePipelineB2k(c, model_->Handle(), graph_->Handle()));
    seconds_per_chunk_ = pipeline_->GetNSampsPerChunk() / model_frequency_;

    const int32 nt = trans_model.NumTransitionIds() + 1;        // the transition model as the determinizer takes it
    phone_of_.assign(nt, 0); self_loop_.assign(nt, 0); phone_start_.assign(nt, 0);
    for (int32 t = 1; t < nt; t++) {
      phone_of_[t] = trans_model.TransitionIdToPhone(t);
      self_loop_[t] = trans_model.IsSelfLoop(t) ? 1 : 0;
      phone_start_[t] = trans_model.TransitionIdIsStartOfPhone(t) ? 1 : 0;
    }
  }

  const Config &GetConfig() { return config_; }
  bool TryInitCorrID(CorrelationID corr_id, int wait_for = 0) { return pipeline_->TryInitCorrID(corr_id, wait_for); }

  void SetBestPathCallback(CorrelationID corr_id, const BestPathCallback &callback) { pipeline_->SetBestPathCallback(corr_id, callback); }
  void SetBestPathCallback(CorrelationID corr_id, BestPathCallback &&callback) { pipeline_->SetBestPathCallback(corr_id, callback); }

  // (CompactLattice&) callbacks are the one-segment case of the segmented form (…online-pipeline.cc:170-188)
  void SetLatticeCallback(CorrelationID corr_id, const LatticeCallback &callback) {
    LatticeCallback cb = callback;
    SetLatticeCallback(corr_id, [cb](SegmentedLatticeCallbackParams &params) {
      if (!params.results.empty()) cb(*params.results[0].GetLatticeResult());
    }, CudaPipelineResult::RESULT_TYPE_LATTICE);
  }
  void SetLatticeCallback(CorrelationID corr_id, LatticeCallback &&callback) {
    const LatticeCallback &cb = callback;
    SetLatticeCallback(corr_id, cb);
  }
  void SetLatticeCallback(CorrelationID corr_id, const SegmentedResultsCallback &callback,
                          const int result_type = CudaPipelineResult::RESULT_TYPE_LATTICE) {
    SegmentedResultsCallback cb = callback;
    pipeline_->SetRawLatticeCallback(corr_id, [this, cb, result_type](CorrelationID, const b2k_raw_lattice &raw) {
      this->FinishStream(raw, cb, result_type);
    });
  }
  void SetLatticeCallback(CorrelationID corr_id, SegmentedResultsCallback &&callback,
                          const int result_type = CudaPipelineResult::RESULT_TYPE_LATTICE) {
    const SegmentedResultsCallback &cb = callback;
    SetLatticeCallback(corr_id, cb, result_type);
  }
  void SetLatticePostprocessor(const std::shared_ptr<LatticePostprocessor> &lattice_postprocessor) {
    lattice_postprocessor_ = lattice_postprocessor;
    lattice_postprocessor_->SetDecoderFrameShift(GetDecoderFrameShiftSeconds());     // …online-pipeline.cc:730-735
    lattice_postprocessor_->SetTransitionInformation(trans_model_);
  }

  void DecodeBatch(const std::vector<CorrelationID> &corr_ids, const std::vector<SubVector<BaseFloat> > &wave_samples,
                   const std::vector<bool> &is_first_chunk, const std::vector<bool> &is_last_chunk,
                   std::vector<const std::string *> *partial_hypotheses = nullptr, std::vector<bool> *end_point = nullptr) {
    pipeline_->DecodeBatch(corr_ids, wave_samples, is_first_chunk, is_last_chunk, partial_hypotheses, end_point);
  }

  int32 GetNSampsPerChunk() const { return pipeline_->GetNSampsPerChunk(); }
  int32 GetNInputFramesPerChunk() const { return pipeline_->GetNInputFramesPerChunk(); }
  BaseFloat GetDecoderFrameShiftSeconds() const { return pipeline_->GetDecoderFrameShiftSeconds(); }
  BaseFloat GetModelFrequency() const { return model_frequency_; }
  TransitionModel const &GetTransitionModel() const { return *trans_model_; }
  BaseFloat GetSecondsPerChunk() const { return seconds_per_chunk_; }
  // the table must outlive the pipeline's last DecodeBatch, as with the reference (cuda-decoder.h keeps the pointer)
  void SetSymbolTable(const fst::SymbolTable &word_syms) {
    const fst::SymbolTable *syms = &word_syms;
    pipeline_->SetWordMapper([syms](int32 w) { return syms->Find(w); });
  }
  void WaitForLatticeCallbacks() noexcept {}          // lattice callbacks have run when DecodeBatch returns

  b2k_shim::StreamingOnlinePipelineB2k &Streaming() { return *pipeline_; }

 private:
  // raw lattice of a finished stream -> determinized CompactLattice (+ CTM) -> the caller's callback
  void FinishStream(const b2k_raw_lattice &raw, const SegmentedResultsCallback &callback, int result_type) {
    b2k_clat *c = NULL;
    b2k_shim::CheckNnet3(b2k_lat_determinize_phone_pruned(&raw, config_.decoder_opts.lattice_beam, 0, phone_of_.data(), self_loop_.data(),
                                                          phone_start_.data(), static_cast<int32>(phone_of_.size()),
                                                          config_.det_opts.phone_determinize ? 1 : 0,
                                                          config_.det_opts.word_determinize ? 1 : 0, &c),
                         "b2k_lat_determinize_phone_pruned");
    CompactLattice clat;
    b2k_shim::BatchedOnlinePipelineB2k::FillCompactLattice(c, &clat);
    b2k_clat_destroy(c);
    SegmentedLatticeCallbackParams params;
    params.results.emplace_back();
    CudaPipelineResult &result = params.results[0];
    result.SetSegmentID(0);
    result.SetAsLastSegment();
    if (result_type & CudaPipelineResult::RESULT_TYPE_CTM) {
      if (!lattice_postprocessor_) KALDI_ERR << "CTM output needs a lattice postprocessor (SetLatticePostprocessor)";
      CTMResult ctm;
      lattice_postprocessor_->GetCTM(clat, &ctm);
      result.SetCTMResult(std::move(ctm));
    }
    if (result_type & CudaPipelineResult::RESULT_TYPE_LATTICE) {
      if (lattice_postprocessor_) {
        CompactLattice post;
        lattice_postprocessor_->GetPostprocessedLattice(clat, &post);
        result.SetLatticeResult(std::move(post));
      } else {
        result.SetLatticeResult(std::move(clat));
      }
    }
    callback(params);
  }

  Config config_;
  const TransitionModel *trans_model_;
  OnlineNnet2FeaturePipelineInfo feature_info_;
  std::unique_ptr<b2k_shim::ModelB2k> model_;
  std::unique_ptr<b2k_shim::CudaFstB2k> graph_;
  std::unique_ptr<b2k_shim::StreamingOnlinePipelineB2k> pipeline_;
  std::shared_ptr<LatticePostprocessor> lattice_postprocessor_;
  BaseFloat model_frequency_ = 16000.0f, seconds_per_chunk_ = 0.0f;
  std::vector<int32_t> phone_of_;
  std::vector<uint8_t> self_loop_, phone_start_;
  KALDI_DISALLOW_COPY_AND_ASSIGN(BatchedThreadedNnet3CudaOnlinePipeline);
};

class CudaOnlinePipelineDynamicBatcher {
 public:
  typedef BatchedThreadedNnet3CudaOnlinePipeline::CorrelationID CorrelationID;
  CudaOnlinePipelineDynamicBatcher(CudaOnlinePipelineDynamicBatcherConfig config, BatchedThreadedNnet3CudaOnlinePipeline &cuda_pipeline)
      : impl_(config.dynamic_batcher_timeout, cuda_pipeline.Streaming(), cuda_pipeline.GetConfig().max_batch_size) {}
  void Push(CorrelationID corr_id, bool is_first_chunk, bool is_last_chunk, const SubVector<BaseFloat> &wave_samples) {
    impl_.Push(corr_id, is_first_chunk, is_last_chunk, wave_samples);
  }
  void WaitForCompletion() { impl_.WaitForCompletion(); }
  int GetNumPendingChunks(CorrelationID corr_id) { return impl_.GetNumPendingChunks(corr_id); }

 private:
  b2k_shim::CudaOnlinePipelineDynamicBatcherB2k impl_;
};

}  // namespace b2k_cuda_dropin
}  // namespace cuda_decoder
}  // namespace kaldi

// From here on the two names mean the adapters (whole tokens only: ...PipelineConfig and ...DynamicBatcherConfig are other
// tokens and stay the reference's structs).
#define BatchedThreadedNnet3CudaOnlinePipeline b2k_cuda_dropin::BatchedThreadedNnet3CudaOnlinePipeline
#define CudaOnlinePipelineDynamicBatcher b2k_cuda_dropin::CudaOnlinePipelineDynamicBatcher

#endif  // B2K_CUDA_PIPELINE_DROPIN_H_
