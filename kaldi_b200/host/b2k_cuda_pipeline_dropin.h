// b2k_cuda_pipeline_dropin.h -- build cudadecoderbin/batched-wav-nnet3-cuda-online.cc and batched-wav-nnet3-cuda2.cc against
// b2k WITHOUT editing them.
//
//   g++ ... -include b2k_cuda_pipeline_dropin.h cudadecoderbin/batched-wav-nnet3-cuda-online.cc ... -lb2k
//   g++ ... -include b2k_cuda_pipeline_dropin.h cudadecoderbin/batched-wav-nnet3-cuda2.cc ... -lb2k
//
// The online tool drives two classes: cuda_decoder::BatchedThreadedNnet3CudaOnlinePipeline (cudadecoder/batched-threaded-nnet3-
// cuda-online-pipeline.h:119-330) and cuda_decoder::CudaOnlinePipelineDynamicBatcher (cuda-online-pipeline-dynamic-batcher.h:
// 38-60); the offline tool (the reference's throughput benchmark) drives cuda_decoder::BatchedThreadedNnet3CudaPipeline2
// (batched-threaded-nnet3-cuda-pipeline2.h:57-245), which is the online pipeline fed whole utterances chunk by chunk.
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
#include <cmath>
#include <cstdlib>
#include <random>                                    // the tool uses std::mt19937 and gets <random> through OpenFst
#include <set>
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
#include "b2k_utterance_pump.h"

namespace kaldi {
namespace cuda_decoder {
namespace b2k_cuda_dropin {

class BatchedThreadedNnet3CudaOnlinePipeline {
 public:
  using CorrelationID = uint64_t;
  typedef std::function<void(const std::string &, bool, bool)> BestPathCallback;
  typedef std::function<void(CompactLattice &)> LatticeCallback;
  typedef kaldi::cuda_decoder::BatchedThreadedNnet3CudaOnlinePipelineConfig Config;

  // max_seconds_per_stream: the device buffers of a channel (samples, features, decoder arenas) are sized for it -- the reference
  // grows its buffers instead; B2K_STREAM_MAX_SECONDS in the environment overrides it for a tool that cannot pass it
  BatchedThreadedNnet3CudaOnlinePipeline(const Config &config, const fst::Fst<fst::StdArc> &decode_fst,
                                         const nnet3::AmNnetSimple &am_nnet, const TransitionModel &trans_model,
                                         BaseFloat max_seconds_per_stream = 60.0f)
      : config_(config), trans_model_(&trans_model), feature_info_(config.feature_opts) {
    config_.compute_opts.CheckAndFixConfigs(am_nnet.GetNnet().Modulus());           // …online-pipeline.h:151-152
    config_.CheckAndFixConfigs();
    if (feature_info_.use_ivectors)
      KALDI_ERR << "the b2k streaming pipeline has no per-chunk i-vector stage (--ivector-extraction-config)";
    if (feature_info_.add_pitch) KALDI_ERR << "b2k has no pitch kernel (--add-pitch)";
    if (config_.reset_on_endpoint) KALDI_ERR << "--reset-on-endpoint is not supported (a stream is one segment)";
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
    c.nchannels = config_.num_channels;
    c.max_seconds = max_seconds_per_stream;
    if (const char *env = std::getenv("B2K_STREAM_MAX_SECONDS")) {
      const double v = std::atof(env);
      if (!(v > 0.0)) KALDI_ERR << "B2K_STREAM_MAX_SECONDS=" << env << " is not a positive number";
      c.max_seconds = static_cast<float>(v);
    }
    // The b2k decoder keeps a stream's tokens and links on the device until its lattice has been read (the reference moves them
    // to the host as it goes, cuda-decoder.cc:1100-1260), in arenas of a fixed size per channel: 12 bytes a token, 16 a link,
    // two links a token on average.  With num_channels in the hundreds that is THE memory of the pipeline, so it is sized from
    // what the device has: half of the free memory over the channels, at most 9000 tokens per decoder frame (what the batched
    // bench gives a lane), and refused when that leaves fewer than 400 (a stream would overflow its arena).
    {
      const int32 sub = std::max(1, config_.compute_opts.frame_subsampling_factor);
      const int64_t frames = static_cast<int64_t>(c.max_seconds * 1000.0f / c.feat.frame_shift_ms) / sub + 16;
      size_t free_bytes = 0, total_bytes = 0;
      if (cudaMemGetInfo(&free_bytes, &total_bytes) != cudaSuccess) KALDI_ERR << "cudaMemGetInfo failed: is a CUDA device selected?";
      const double per_channel = 0.5 * static_cast<double>(free_bytes) / std::max(1, config_.num_channels);
      int64_t tokens = static_cast<int64_t>(per_channel / (12.0 + 2.0 * 16.0));
      tokens = std::min<int64_t>(tokens, frames * 9000);
      if (tokens < frames * 400)
        KALDI_ERR << "not enough device memory for " << config_.num_channels << " channels of " << c.max_seconds
                  << " s each (" << (free_bytes >> 20) << " MB free): lower --num-channels / --max-batch-size, or the stream "
                  << "length (B2K_STREAM_MAX_SECONDS, --segment-length)";
      c.dec = d.ToB2k(static_cast<int32>(frames));
      c.dec.max_tokens = tokens;
      c.dec.max_links = 2 * tokens;
    }
    c.frames_per_chunk = config_.compute_opts.frames_per_chunk;
    c.acoustic_scale = config_.compute_opts.acoustic_scale;
    c.use_priors = 1;
    pipeline_.reset(new b2k_shim::StreamingOnlinePipelineB2k(c, model_->Handle(), graph_->Handle()));
    // the end-point rules are CudaDecoderConfig's (cuda-decoder.h:64,147: --endpoint.* registered beside the decoder options)
    pipeline_->SetEndpointConfig(b2k_shim::ToB2kEndpointConfig(config_.decoder_opts.endpointing_config));
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
    CompactLattice clat;
    if (config_.determinize_lattice) {                                                   // …online-pipeline.cc:755-760
      b2k_clat *c = NULL;
      b2k_shim::CheckNnet3(b2k_lat_determinize_phone_pruned(&raw, config_.decoder_opts.lattice_beam, 0, phone_of_.data(), self_loop_.data(),
                                                            phone_start_.data(), static_cast<int32>(phone_of_.size()),
                                                            config_.det_opts.phone_determinize ? 1 : 0,
                                                            config_.det_opts.word_determinize ? 1 : 0, &c),
                           "b2k_lat_determinize_phone_pruned");
      if (config_.det_opts.minimize) b2k_shim::CheckNnet3(b2k_clat_minimize(c, 1.0f / 1024.0f /* fst::kDelta, as the reference calls it */), "b2k_clat_minimize");
      b2k_shim::BatchedOnlinePipelineB2k::FillCompactLattice(c, &clat);
      b2k_clat_destroy(c);
    } else {
      // ConvertLattice(lat, &clat) (fstext/lattice-utils.h:91): the raw lattice state for state, the word as the label, the
      // transition-id (if any) as a one-element string
      for (int64_t s = 0; s < raw.num_states; s++) clat.AddState();
      if (raw.num_states > 0) clat.SetStart(0);
      for (int64_t a = 0; a < raw.num_arcs; a++) {
        std::vector<int32> str;
        if (raw.arc_ilabel[a] != 0) str.push_back(raw.arc_ilabel[a]);
        clat.AddArc(raw.arc_src[a], CompactLatticeArc(raw.arc_olabel[a], raw.arc_olabel[a],
                                                       CompactLatticeWeight(LatticeWeight(raw.arc_graph_cost[a], raw.arc_acoustic_cost[a]), str),
                                                       raw.arc_dst[a]));
      }
      for (int64_t f = 0; f < raw.num_finals; f++)
        clat.SetFinal(raw.final_state[f], CompactLatticeWeight(LatticeWeight(raw.final_cost[f], 0.0f), std::vector<int32>()));
    }
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

// cuda_decoder::BatchedThreadedNnet3CudaPipeline2 (cudadecoder/batched-threaded-nnet3-cuda-pipeline2.h:57-245): whole
// utterances in, one callback per utterance out.  Like the reference's class it owns an online pipeline and feeds it the
// utterances chunk by chunk (b2k_host::UtterancePump: the reference's control thread, here on the caller's thread -- a full
// batch of utterances is decoded as soon as there is one, WaitForAllTasks / WaitForGroup decode what is left, callbacks run on
// the calling thread).  Segmentation (seg_opts) is the reference's: NumberOfSegments, the per-segment offsets, segments shorter
// than --min-segment-length dropped (:265-334).
class BatchedThreadedNnet3CudaPipeline2 {
 public:
  typedef kaldi::cuda_decoder::BatchedThreadedNnet3CudaPipeline2Config Config;
  BatchedThreadedNnet3CudaPipeline2(const Config &config, const fst::Fst<fst::StdArc> &decode_fst,
                                    const nnet3::AmNnetSimple &am_nnet, const TransitionModel &trans_model)
      : config_(config),
        // a stream is one segment: with the default segmentation (20 s) the channels are sized for a segment, not for 60 s
        online_(config.cuda_online_pipeline_opts, decode_fst, am_nnet, trans_model,
                static_cast<BaseFloat>(std::min(60.0, config.seg_opts.segment_length_s) + 1.0)),
        backend_{&online_},
        pump_(&backend_, online_.GetConfig().max_batch_size, online_.GetNSampsPerChunk()) {
    config_.Check();
    model_freq_ = online_.GetModelFrequency();
    segment_length_nsamples_ = config_.seg_opts.segment_length_s * model_freq_;                                  // :77-83
    segment_shift_nsamples_ = (config_.seg_opts.segment_length_s - config_.seg_opts.segment_overlap_s) * model_freq_;
    min_segment_length_nsamples_ = config_.seg_opts.min_segment_length_s * model_freq_;
  }
  virtual ~BatchedThreadedNnet3CudaPipeline2() {}

  void SegmentedDecodeWithCallback(const std::shared_ptr<WaveData> &wave_data, const SegmentedResultsCallback &segmented_callback,
                                   const int result_type = CudaPipelineResult::RESULT_TYPE_LATTICE) {
    KALDI_ASSERT(result_type && "You must define at least one result type");
    KALDI_ASSERT("Mismatch in model and utt frequency" && (wave_data->SampFreq() == model_freq_));
    SubVector<BaseFloat> h_wave(wave_data->Data(), 0);
    const int total_nsamples = h_wave.Dim();
    if (total_nsamples == 0) {                       // one empty lattice (:336-345)
      SegmentedLatticeCallbackParams params;
      params.results.emplace_back();
      CompactLattice clat;
      params.results.back().SetLatticeResult(std::move(clat));
      segmented_callback(params);
      return;
    }
    const int nsegments = NumberOfSegments(total_nsamples, segment_length_nsamples_, segment_shift_nsamples_);
    // what the segments of one utterance share: the result slots, how many are still out, the audio
    struct Shared {
      std::vector<CudaPipelineResult> results;
      int not_done;
      std::shared_ptr<WaveData> wave;
      SegmentedResultsCallback callback;
    };
    std::shared_ptr<Shared> shared = std::make_shared<Shared>();
    shared->wave = wave_data;
    shared->callback = segmented_callback;
    std::vector<std::pair<int, int> > segments;      // (offset, nsamples) of the segments that are decoded
    int dropped = 0;
    for (int offset = 0;; offset += segment_shift_nsamples_) {
      const int nsamples = std::min(total_nsamples - offset, segment_length_nsamples_);
      if (nsamples >= min_segment_length_nsamples_) segments.push_back(std::make_pair(offset, nsamples));
      else dropped++;
      if (offset + nsamples >= total_nsamples) break;
    }
    KALDI_ASSERT(nsegments - dropped == static_cast<int>(segments.size()));                  // :333
    shared->results.resize(segments.size());
    shared->not_done = static_cast<int>(segments.size());
    if (segments.empty()) { SegmentedLatticeCallbackParams params; segmented_callback(params); return; }
    std::shared_ptr<LatticePostprocessor> *postprocessor = &lattice_postprocessor_;
    for (size_t i = 0; i < segments.size(); i++) {
      shared->results[i].SetTimeOffsetSeconds(std::floor(static_cast<BaseFloat>(segments[i].first) / model_freq_));
      shared->results[i].SetSegmentID(static_cast<int>(i));
      if (i + 1 == segments.size()) shared->results[i].SetAsLastSegment();
      const uint64_t corr_id = next_corr_id_++;
      online_.SetLatticeCallback(corr_id, [shared, i, result_type, postprocessor](CompactLattice &clat) {
        SetResultUsingLattice(clat, result_type, *postprocessor, &shared->results[i]);      // lattice-postprocessor.cc:115
        if (--shared->not_done == 0) {
          SegmentedLatticeCallbackParams params;
          params.results = std::move(shared->results);
          shared->callback(params);
        }
      });
      pump_.Add(corr_id, h_wave.Data() + segments[i].first, segments[i].second, [shared]() {});   // keeps the audio alive
    }
    pump_.Run(false);
  }

  void DecodeWithCallback(const std::shared_ptr<WaveData> &wave_data, const std::function<void(CompactLattice &)> &callback,
                          const std::string &group = std::string()) {
    KALDI_ASSERT("Mismatch in model and utt frequency" && (wave_data->SampFreq() == model_freq_));
    SubVector<BaseFloat> h_wave(wave_data->Data(), 0);
    std::shared_ptr<WaveData> keep = wave_data;
    Submit(h_wave.Data(), h_wave.Dim(), callback, group, [keep]() {});
  }
  // the samples must stay valid until the callback has run (the reference keeps a SubVector onto them too, :190-196)
  void DecodeWithCallback(const VectorBase<BaseFloat> &wave_data, float sample_rate,
                          const std::function<void(CompactLattice &)> &callback, const std::string &group = std::string()) {
    KALDI_ASSERT(sample_rate == model_freq_);
    Submit(wave_data.Data(), wave_data.Dim(), callback, group, std::function<void()>());
  }

  void SetLatticePostprocessor(const std::shared_ptr<LatticePostprocessor> &lattice_postprocessor) {
    lattice_postprocessor_ = lattice_postprocessor;
    lattice_postprocessor_->SetDecoderFrameShift(online_.GetDecoderFrameShiftSeconds());     // …pipeline2.cc:148-154
    lattice_postprocessor_->SetTransitionInformation(&online_.GetTransitionModel());
  }
  // groups exist so that a caller can wait for some of its tasks; on one thread every wait decodes what is pending
  void CreateTaskGroup(const std::string &group) { KALDI_ASSERT("Group is already in use" && groups_.insert(group).second); }
  void DestroyTaskGroup(const std::string &group) { KALDI_ASSERT("Group does not exist" && groups_.erase(group) == 1); }
  void WaitForGroup(const std::string &group) {
    KALDI_ASSERT("Group does not exist. Call CreateTaskGroup() first" && groups_.count(group));
    pump_.Run(true);
  }
  void WaitForAllTasks() { pump_.Run(true); }
  void SetSymbolTable(const fst::SymbolTable &word_syms) { online_.SetSymbolTable(word_syms); }

 private:
  void Submit(const BaseFloat *samples, int32 num_samples, const std::function<void(CompactLattice &)> &callback,
              const std::string &group, std::function<void()> release) {
    if (!group.empty()) KALDI_ASSERT("Group does not exist. Call CreateTaskGroup() first" && groups_.count(group));
    if (num_samples == 0) return;                    // nothing to do (:218)
    const uint64_t corr_id = next_corr_id_++;
    online_.SetLatticeCallback(corr_id, callback);
    pump_.Add(corr_id, samples, num_samples, std::move(release));
    pump_.Run(false);
  }
  struct Backend {                                   // UtterancePump's view of the online pipeline
    BatchedThreadedNnet3CudaOnlinePipeline *online;
    void DecodeBatch(const std::vector<uint64_t> &ids, const std::vector<std::pair<const float *, int64_t> > &chunks,
                     const std::vector<bool> &first, const std::vector<bool> &last) {
      std::vector<SubVector<BaseFloat> > waves;
      for (size_t i = 0; i < chunks.size(); i++)
        waves.push_back(SubVector<BaseFloat>(const_cast<BaseFloat *>(chunks[i].first), static_cast<MatrixIndexT>(chunks[i].second)));
      online->DecodeBatch(ids, waves, first, last);
    }
  };

  const Config &config_;
  BatchedThreadedNnet3CudaOnlinePipeline online_;
  Backend backend_;
  b2k_host::UtterancePump<Backend> pump_;
  std::shared_ptr<LatticePostprocessor> lattice_postprocessor_;
  std::set<std::string> groups_;
  uint64_t next_corr_id_ = 0;
  BaseFloat model_freq_ = 16000.0f;
  int segment_length_nsamples_ = 0, segment_shift_nsamples_ = 0, min_segment_length_nsamples_ = 0;
  KALDI_DISALLOW_COPY_AND_ASSIGN(BatchedThreadedNnet3CudaPipeline2);
};

}  // namespace b2k_cuda_dropin
}  // namespace cuda_decoder
}  // namespace kaldi

// From here on the three names mean the adapters (whole tokens only: ...PipelineConfig, ...Pipeline2Config and
// ...DynamicBatcherConfig are other tokens and stay the reference's structs).
#include "b2k_dropin_common.h"                      // leaves nnet3::CollapseModel out: b2k takes the model as trained
#define BatchedThreadedNnet3CudaOnlinePipeline b2k_cuda_dropin::BatchedThreadedNnet3CudaOnlinePipeline
#define CudaOnlinePipelineDynamicBatcher b2k_cuda_dropin::CudaOnlinePipelineDynamicBatcher
#define BatchedThreadedNnet3CudaPipeline2 b2k_cuda_dropin::BatchedThreadedNnet3CudaPipeline2

#endif  // B2K_CUDA_PIPELINE_DROPIN_H_
