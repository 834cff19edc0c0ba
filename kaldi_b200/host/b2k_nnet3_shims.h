// b2k_nnet3_shims.h — header-only C++ shims with the reference's nnet3 inference surfaces over the b2k C ABI:
//
//   nnet3::NnetComputer {AcceptInput, Run, GetOutput, GetOutputDestructive}            nnet3/nnet-compute.h:88-200
//   nnet3::DecodableNnetLoopedOnline / DecodableAmNnetLoopedOnline                     nnet3/decodable-online-looped.h:49-196
//   cuda_decoder::BatchedStaticNnet3 {RunBatch, FormatOutputPtrs, Get*}                cudadecoder/batched-static-nnet3.h:59-138
//   nnet3::DecodableNnetSimple / DecodableAmNnetSimple (the offline tools' decodables)  nnet3/nnet-am-decodable-simple.h:185-349
//   SingleUtteranceNnet3Decoder {AdvanceDecoding, FinalizeDecoding, EndpointDetected, GetBestPath, GetLattice}
//                                                                                       online2/online-nnet3-decoding.h:50-121
//
// The model goes from the reference's objects to libb2k.so through the reference's own serialisation
// (Nnet::Write / AmNnetSimple::Write into memory -> b2k_model_read_memory), so no component walker has to track
// nnet3's class list.  Everything here is type-checked against the reference's headers by oracle/check_shims.py
// (HAVE_CUDA=1: CuMatrix holds device memory, which is what b2k reads and writes).
#ifndef B2K_NNET3_SHIMS_H_
#define B2K_NNET3_SHIMS_H_

#include <cuda_runtime.h>

#include <algorithm>
#include <memory>
#include <sstream>
#include <string>
#include <utility>
#include <vector>

#include "b2k.h"
#include "b2k_kaldi_shims.h"
#include "base/kaldi-error.h"
#include "cudamatrix/cu-device.h"
#include "cudamatrix/cu-matrix.h"
#include "hmm/transition-model.h"
#include "itf/decodable-itf.h"
#include "itf/online-feature-itf.h"
#include "nnet3/am-nnet-simple.h"
#include "nnet3/decodable-simple-looped.h"
#include "nnet3/nnet-am-decodable-simple.h"
#include "nnet3/nnet-computation.h"
#include "nnet3/nnet-compute.h"
#include "nnet3/nnet-utils.h"

namespace kaldi {
namespace b2k_shim {

inline void CheckNnet3(int rc, const char *what) {
  if (rc != B2K_OK) KALDI_ERR << what << ": " << b2k_last_error();      // throws KaldiFatalError like the reference
}

inline void RequireCuDevice() {
#if HAVE_CUDA == 1
  if (!CuDevice::Instantiate().Enabled())
    KALDI_ERR << "b2k needs the CUDA device to be selected (CuDevice::SelectGpuId): it has no CPU path";
#else
  KALDI_ERR << "b2k needs a Kaldi built with CUDA: CuMatrix must hold device memory";
#endif
}

// The b2k view of a model: layer list + named weights (+ priors).  Owns the handle.
class ModelB2k {
 public:
  // kind 0: raw Nnet, 2: AmNnetSimple (priors included)
  ModelB2k(const nnet3::Nnet &nnet, int32 frame_subsampling_factor) { Init(nnet, NULL, frame_subsampling_factor); }
  ModelB2k(const nnet3::AmNnetSimple &am_nnet, int32 frame_subsampling_factor) {
    Init(am_nnet.GetNnet(), &am_nnet, frame_subsampling_factor);
  }
  ~ModelB2k() { b2k_model_destroy(m_); }
  const b2k_model *Handle() const { return m_; }
  int32 FeatDim() const { return info_[0]; }
  int32 IvectorDim() const { return info_[1]; }
  int32 NumPdfs() const { return info_[2]; }
  int32 NumLayers() const { return info_[4]; }
  int32 NumWeights() const { return info_[5]; }
  b2k_nnet_compile_cfg CompileConfig(int32 frames_per_chunk, BaseFloat acoustic_scale, bool use_priors) const {
    b2k_nnet_compile_cfg c;
    c.feat_dim = info_[0]; c.ivector_dim = info_[1]; c.num_pdfs = info_[2]; c.frame_subsampling_factor = info_[3];
    c.num_frames = 0; c.frames_per_chunk = frames_per_chunk; c.use_priors = use_priors ? 1 : 0; c.conv_dense = 0;
    c.acoustic_scale = acoustic_scale;
    return c;
  }

 private:
  void Init(const nnet3::Nnet &nnet, const nnet3::AmNnetSimple *am, int32 sf) {
    std::ostringstream os(std::ios::binary);
    os.put('\0'); os.put('B');                         // InitKaldiOutputStream's binary marker (base/io-funcs-inl.h)
    if (am) am->Write(os, true); else nnet.Write(os, true);
    const std::string s = os.str();
    CheckNnet3(b2k_model_read_memory(s.data(), static_cast<int64_t>(s.size()), am ? 2 : 0, &m_), "b2k_model_read_memory");
    CheckNnet3(b2k_model_set_frame_subsampling_factor(m_, sf), "b2k_model_set_frame_subsampling_factor");
    CheckNnet3(b2k_model_info(m_, info_), "b2k_model_info");
  }
  b2k_model *m_ = NULL;
  int32_t info_[8];
  KALDI_DISALLOW_COPY_AND_ASSIGN(ModelB2k);
};

// nnet3::NnetComputer for the forward computations the inference tools request (nnet3/nnet-compute.h:88-200).  The reference
// takes a compiled NnetComputation; b2k compiles its own program from the same ComputationRequest, so the constructor takes the
// request.  Supported requests are the ones DecodableNnetSimple (nnet-am-decodable-simple.cc:231-300) and BatchedStaticNnet3
// (batched-static-nnet3.cc:123-152) build: for every sequence n the same contiguous range of input times, outputs at
// t0 + k*stride, at most one "ivector" index per sequence; anything else is refused (KALDI_ERR).  No derivatives.
class NnetComputerB2k {
 public:
  NnetComputerB2k(const nnet3::NnetComputeOptions &options, const nnet3::ComputationRequest &request, const nnet3::Nnet &nnet)
      : debug_(options.debug) {
    RequireCuDevice();
    if (request.need_model_derivative || request.store_component_stats) KALDI_ERR << "b2k runs forward computations only";
    const nnet3::IoSpecification *in = NULL, *iv = NULL;
    for (size_t i = 0; i < request.inputs.size(); i++) {
      if (request.inputs[i].has_deriv) KALDI_ERR << "b2k runs forward computations only";
      if (request.inputs[i].name == "input") in = &request.inputs[i];
      else if (request.inputs[i].name == "ivector") iv = &request.inputs[i];
      else KALDI_ERR << "unexpected input " << request.inputs[i].name;
    }
    if (!in || request.outputs.size() != 1 || request.outputs[0].name != "output") KALDI_ERR << "request must have input 'input' and output 'output'";
    const std::vector<nnet3::Index> &ii = in->indexes, &oi = request.outputs[0].indexes;
    if (ii.empty() || oi.empty()) KALDI_ERR << "empty request";
    // n-major, the same times for every n
    num_seq_ = 0;
    for (size_t i = 0; i < ii.size(); i++) num_seq_ = std::max<int32>(num_seq_, ii[i].n + 1);
    if (ii.size() % num_seq_ != 0 || oi.size() % num_seq_ != 0) KALDI_ERR << "request is not rectangular over n";
    in_frames_ = ii.size() / num_seq_; out_frames_ = oi.size() / num_seq_;
    const int32 t_in0 = ii[0].t, t_out0 = oi[0].t;
    stride_ = out_frames_ > 1 ? oi[1].t - oi[0].t : 1;
    for (int32 n = 0; n < num_seq_; n++) {
      for (int32 k = 0; k < in_frames_; k++) {
        const nnet3::Index &x = ii[n * in_frames_ + k];
        if (x.n != n || x.t != t_in0 + k || x.x != 0) KALDI_ERR << "unsupported input indexes (need n-major, contiguous t)";
      }
      for (int32 k = 0; k < out_frames_; k++) {
        const nnet3::Index &x = oi[n * out_frames_ + k];
        if (x.n != n || x.t != t_out0 + k * stride_ || x.x != 0) KALDI_ERR << "unsupported output indexes (need n-major, constant stride)";
      }
    }
    if (iv && static_cast<int32>(iv->indexes.size()) != num_seq_) KALDI_ERR << "b2k takes one i-vector per sequence";
    if (stride_ <= 0) KALDI_ERR << "output stride must be positive";
    model_.reset(new ModelB2k(nnet, stride_));
    if ((model_->IvectorDim() > 0) != (iv != NULL)) KALDI_ERR << "the request and the model disagree about the i-vector input";
    b2k_nnet_compile_cfg c = model_->CompileConfig(stride_, 1.0f, false);       // raw output: no priors, no scale (the callers apply them)
    c.num_frames = in_frames_;
    b2k_nnet_program *prog = NULL;
    CheckNnet3(b2k_nnet_compile_window(&c, t_out0 - t_in0, out_frames_, 1, b2k_model_layers(model_->Handle()), model_->NumLayers(),
                                       b2k_model_weights(model_->Handle()), model_->NumWeights(), &prog),
               "b2k_nnet_compile_window");
    const int rc = b2k_nnet_create_from_program(prog, num_seq_, &nn_);
    b2k_nnet_program_destroy(prog);
    CheckNnet3(rc, "b2k_nnet_create_from_program");
  }
  ~NnetComputerB2k() { if (nn_) b2k_nnet_destroy(nn_); }

  // "the input is destroyed": swapped into the computer, as the reference does (nnet-compute.cc:538-560)
  void AcceptInput(const std::string &node_name, CuMatrix<BaseFloat> *input) {
    if (node_name == "input") {
      if (input->NumRows() != num_seq_ * in_frames_ || input->NumCols() != model_->FeatDim())
        KALDI_ERR << "Num-rows or cols mismatch for input 'input': " << input->NumRows() << " x " << input->NumCols();
      input_.Swap(input); input->Resize(0, 0);
    } else if (node_name == "ivector") {
      if (input->NumRows() != num_seq_ || input->NumCols() != model_->IvectorDim())
        KALDI_ERR << "Num-rows or cols mismatch for input 'ivector'";
      ivector_.Swap(input); input->Resize(0, 0);
    } else {
      KALDI_ERR << "No node named '" << node_name << "' in network.";
    }
  }
  void Run() {
    if (input_.NumRows() == 0 || (model_->IvectorDim() > 0 && ivector_.NumRows() == 0)) KALDI_ERR << "Run() before all inputs were provided";
    output_.Resize(num_seq_ * out_frames_, model_->NumPdfs(), kUndefined);
    std::vector<const float *> in(num_seq_), iv(num_seq_);
    std::vector<float *> out(num_seq_);
    for (int32 n = 0; n < num_seq_; n++) {
      in[n] = input_.Data() + static_cast<size_t>(n) * in_frames_ * input_.Stride();
      iv[n] = model_->IvectorDim() > 0 ? ivector_.Data() + static_cast<size_t>(n) * ivector_.Stride() : NULL;
      out[n] = output_.Data() + static_cast<size_t>(n) * out_frames_ * output_.Stride();
    }
    CheckNnet3(b2k_nnet_run(nn_, num_seq_, in.data(), input_.Stride(), model_->IvectorDim() > 0 ? iv.data() : NULL,
                            model_->IvectorDim() > 0 ? ivector_.Stride() : 0, out.data(), output_.Stride(), cudaStreamPerThread),
               "b2k_nnet_run");
    if (debug_) CheckCuda(cudaStreamSynchronize(cudaStreamPerThread));
    input_.Resize(0, 0); ivector_.Resize(0, 0);
    ran_ = true;
  }
  const CuMatrixBase<BaseFloat> &GetOutput(const std::string &node_name) {
    if (node_name != "output" || !ran_) KALDI_ERR << "GetOutput: no output named " << node_name << " has been computed";
    return output_;
  }
  void GetOutputDestructive(const std::string &output_name, CuMatrix<BaseFloat> *output) {
    if (output_name != "output" || !ran_) KALDI_ERR << "GetOutputDestructive: no output named " << output_name << " has been computed";
    output->Resize(0, 0);
    output->Swap(&output_);
    ran_ = false;
  }

 private:
  static void CheckCuda(cudaError_t e) { if (e != cudaSuccess) KALDI_ERR << cudaGetErrorString(e); }
  bool debug_;
  int32 num_seq_ = 0, in_frames_ = 0, out_frames_ = 0, stride_ = 1;
  std::unique_ptr<ModelB2k> model_;
  b2k_nnet *nn_ = NULL;
  CuMatrix<BaseFloat> input_, ivector_, output_;
  bool ran_ = false;
  KALDI_DISALLOW_COPY_AND_ASSIGN(NnetComputerB2k);
};

// cuda_decoder::BatchedStaticNnet3 (cudadecoder/batched-static-nnet3.h:59-138): same constructor arguments, same RunBatch.
// The reference sub-batches by 64 sequences (MAX_COMPUTE_BATCH_SIZE) because its computation is compiled for that many; b2k's
// executor takes the whole batch.
template <class Config /* cuda_decoder::BatchedStaticNnet3Config */>
class BatchedStaticNnet3B2k {
 public:
  BatchedStaticNnet3B2k(const Config &config, const nnet3::AmNnetSimple &am_nnet)
      : max_batch_size_(config.max_batch_size), has_ivector_(config.has_ivector) {
    RequireCuDevice();
    nchannels_ = (config.nchannels != -1) ? config.nchannels : max_batch_size_;
    KALDI_ASSERT(max_batch_size_ > 0);
    KALDI_ASSERT(nchannels_ >= max_batch_size_);
    const nnet3::NnetSimpleComputationOptions &o = config.compute_opts;
    if (o.extra_left_context != 0 || o.extra_right_context != 0 || o.extra_left_context_initial > 0 || o.extra_right_context_final > 0)
      KALDI_ERR << "b2k: extra left / right context is not supported";
    model_.reset(new ModelB2k(am_nnet, o.frame_subsampling_factor));
    if (has_ivector_ != (model_->IvectorDim() > 0)) KALDI_ERR << "has_ivector disagrees with the model";
    const b2k_nnet_compile_cfg c = model_->CompileConfig(o.frames_per_chunk, o.acoustic_scale, am_nnet.Priors().Dim() != 0);
    CheckNnet3(b2k_nnet_stream_create(&c, b2k_model_layers(model_->Handle()), model_->NumLayers(), b2k_model_weights(model_->Handle()),
                                      model_->NumWeights(), max_batch_size_, nchannels_, /*looped=*/0, &s_), "b2k_nnet_stream_create");
    int64_t info[8];
    CheckNnet3(b2k_nnet_stream_info(s_, info), "b2k_nnet_stream_info");
    output_frames_per_chunk_ = info[0]; total_nnet_right_context_ = info[2]; num_pdfs_ = info[6];
  }
  virtual ~BatchedStaticNnet3B2k() { if (s_) b2k_nnet_stream_destroy(s_); }

  void RunBatch(const std::vector<int> &channels, const std::vector<BaseFloat *> &d_features, const int features_stride,
                const std::vector<BaseFloat *> &d_ivectors, const std::vector<int> &n_input_frames_valid,
                const std::vector<bool> &is_first_chunk, const std::vector<bool> &is_last_chunk,
                CuMatrix<BaseFloat> *d_all_log_posteriors,
                std::vector<std::vector<std::pair<int, const BaseFloat *>>> *all_frames_log_posteriors_ptrs) {
    const size_t n = channels.size();
    KALDI_ASSERT(d_features.size() >= n);                      // batched-static-nnet3.cc:306-312
    KALDI_ASSERT(is_last_chunk.size() >= n);
    KALDI_ASSERT(is_first_chunk.size() >= n);
    if (has_ivector_) KALDI_ASSERT(d_ivectors.size() >= n);
    const int32 rows = max_batch_size_ * output_frames_per_chunk_;
    if (d_all_log_posteriors->NumRows() < rows || d_all_log_posteriors->NumCols() != num_pdfs_)
      d_all_log_posteriors->Resize(rows, num_pdfs_, kUndefined);
    bool any_last = false;
    first_.resize(n); last_.resize(n); n_out_.resize(n); n_eos_.resize(n); feats_.resize(n); ivecs_.resize(n);
    for (size_t i = 0; i < n; i++) {
      first_[i] = is_first_chunk[i] ? 1 : 0; last_[i] = is_last_chunk[i] ? 1 : 0; any_last |= is_last_chunk[i];
      feats_[i] = d_features[i]; ivecs_[i] = has_ivector_ ? d_ivectors[i] : NULL;
    }
    if (any_last && (d_all_eos_log_posteriors_.NumRows() != d_all_log_posteriors->NumRows() ||
                     d_all_eos_log_posteriors_.Stride() != d_all_log_posteriors->Stride()))
      d_all_eos_log_posteriors_.Resize(d_all_log_posteriors->NumRows(), d_all_log_posteriors->NumCols(), kUndefined);
    if (any_last && d_all_eos_log_posteriors_.Stride() != d_all_log_posteriors->Stride())
      KALDI_ERR << "b2k: the two posterior matrices must have the same stride";
    CheckNnet3(b2k_nnet_stream_run_batch(s_, static_cast<int32_t>(n), channels.data(), feats_.data(), features_stride,
                                         has_ivector_ ? ivecs_.data() : NULL, n_input_frames_valid.data(), first_.data(), last_.data(),
                                         d_all_log_posteriors->Data(), any_last ? d_all_eos_log_posteriors_.Data() : NULL,
                                         d_all_log_posteriors->Stride(), n_out_.data(), n_eos_.data(), cudaStreamPerThread),
               "b2k_nnet_stream_run_batch");
    all_frames_log_posteriors_ptrs->clear();
    FormatOutputPtrs(channels, d_all_log_posteriors, all_frames_log_posteriors_ptrs, n_out_);
    if (any_last) FormatOutputPtrs(channels, &d_all_eos_log_posteriors_, all_frames_log_posteriors_ptrs, n_eos_, &n_out_);
  }

  // batched-static-nnet3.cc:369-395; slot i's frames are rows i*output_frames_per_chunk + k of the matrix it is given (the
  // flush of slot i lies in the same rows of the second matrix, where the reference compacts the ended channels)
  void FormatOutputPtrs(const std::vector<int> &channels, CuMatrix<BaseFloat> *d_all_log_posteriors,
                        std::vector<std::vector<std::pair<int, const BaseFloat *>>> *all_frames_log_posteriors_ptrs,
                        const std::vector<int> &n_output_frames_valid, const std::vector<int> *n_output_frames_valid_offset = NULL) {
    KALDI_ASSERT(channels.size() == n_output_frames_valid.size());
    for (size_t i = 0; i < channels.size(); ++i) {
      const int offset = n_output_frames_valid_offset ? (*n_output_frames_valid_offset)[i] : 0;
      const int total = offset + n_output_frames_valid[i];
      if (static_cast<int>(all_frames_log_posteriors_ptrs->size()) < total) all_frames_log_posteriors_ptrs->resize(total);
      for (int iframe = offset; iframe < total; ++iframe) {
        const BaseFloat *frame = d_all_log_posteriors->Data() +
            static_cast<size_t>(i * output_frames_per_chunk_ + (iframe - offset)) * d_all_log_posteriors->Stride();
        (*all_frames_log_posteriors_ptrs)[iframe].push_back({channels[i], frame});
      }
    }
  }

  int GetNOutputFramesPerChunk() { return output_frames_per_chunk_; }
  int GetTotalNnet3RightContext() { return total_nnet_right_context_; }

 private:
  int max_batch_size_, nchannels_;
  bool has_ivector_;
  int output_frames_per_chunk_ = 0, total_nnet_right_context_ = 0, num_pdfs_ = 0;
  std::unique_ptr<ModelB2k> model_;
  b2k_nnet_stream *s_ = NULL;
  CuMatrix<BaseFloat> d_all_eos_log_posteriors_;
  std::vector<int32_t> first_, last_, n_out_, n_eos_;
  std::vector<const float *> feats_, ivecs_;
};

// nnet3::DecodableNnetLoopedOnlineBase (nnet3/decodable-online-looped.h:49-133) over b2k_nnet_stream in looped mode: the
// chunk schedule, the clamping of the input at both ends of the utterance and the i-vector every chunk receives are
// AdvanceChunk's (decodable-online-looped.cc:118-236); the network runs on the device, one chunk per call.
class DecodableNnetLoopedOnlineBaseB2k : public DecodableInterface {
 public:
  // `info` supplies the options and the model; the b2k program is compiled here, once per object -- share one object per
  // thread of decoding, or reuse it across utterances with Reset().
  DecodableNnetLoopedOnlineBaseB2k(const nnet3::DecodableNnetSimpleLoopedInfo &info, OnlineFeatureInterface *input_features,
                                   OnlineFeatureInterface *ivector_features)
      : info_(info), input_features_(input_features), ivector_features_(ivector_features) {
    RequireCuDevice();
    KALDI_ASSERT(input_features_ != NULL);
    if (info_.opts.extra_left_context_initial > 0) KALDI_ERR << "b2k: extra-left-context-initial is not supported";
    if (info_.has_ivectors && ivector_features_ == NULL) KALDI_ERR << "the model needs i-vectors but no i-vector feature was given";
    model_.reset(new ModelB2k(info_.nnet, info_.opts.frame_subsampling_factor));
    // priors and the acoustic scale exactly where the reference applies them: after the network, on this side
    // (decodable-online-looped.cc:218-223), so that NnetSimpleLoopedComputationOptions keeps its meaning
    const b2k_nnet_compile_cfg c = model_->CompileConfig(info_.frames_per_chunk, 1.0f, false);
    CheckNnet3(b2k_nnet_stream_create(&c, b2k_model_layers(model_->Handle()), model_->NumLayers(), b2k_model_weights(model_->Handle()),
                                      model_->NumWeights(), 1, 1, /*looped=*/1, &s_), "b2k_nnet_stream_create");
    int64_t si[8];
    CheckNnet3(b2k_nnet_stream_info(s_, si), "b2k_nnet_stream_info");
    if (si[1] != info_.frames_left_context || si[2] != info_.frames_right_context)
      KALDI_ERR << "b2k computed context " << si[1] << "/" << si[2] << ", the reference " << info_.frames_left_context << "/" << info_.frames_right_context;
    ivector_rows_ = si[7];
    d_out_.Resize(si[0], info_.output_dim, kUndefined);
    Reset();
  }
  ~DecodableNnetLoopedOnlineBaseB2k() override { if (s_) b2k_nnet_stream_destroy(s_); }

  void Reset() {        // a new utterance on the same object
    num_chunks_computed_ = 0; current_log_post_subsampled_offset_ = -1; frame_offset_ = 0;
    current_log_post_.Resize(0, 0);
    chunk_ivectors_.clear();
  }

  bool IsLastFrame(int32 subsampled_frame) const override {                           // decodable-online-looped.cc:88-110
    const int32 features_ready = input_features_->NumFramesReady();
    if (features_ready == 0) return subsampled_frame == -1 && input_features_->IsLastFrame(-1);
    if (!input_features_->IsLastFrame(features_ready - 1)) return false;
    const int32 sf = info_.opts.frame_subsampling_factor, num_subsampled_frames_ready = (features_ready + sf - 1) / sf;
    return subsampled_frame + frame_offset_ == num_subsampled_frames_ready - 1;
  }
  int32 NumFramesReady() const override {                                             // :56-84
    const int32 features_ready = input_features_->NumFramesReady();
    if (features_ready == 0) return 0;
    const bool input_finished = input_features_->IsLastFrame(features_ready - 1);
    const int32 sf = info_.opts.frame_subsampling_factor;
    if (input_finished) return (features_ready + sf - 1) / sf - frame_offset_;
    const int32 non_subsampled_output_frames_ready = std::max<int32>(0, features_ready - info_.frames_right_context);
    const int32 num_chunks_ready = non_subsampled_output_frames_ready / info_.frames_per_chunk;
    return num_chunks_ready * info_.frames_per_chunk / sf - frame_offset_;
  }
  int32 FrameSubsamplingFactor() const { return info_.opts.frame_subsampling_factor; }
  void SetFrameOffset(int32 frame_offset) {
    KALDI_ASSERT(0 <= frame_offset && frame_offset <= frame_offset_ + NumFramesReady());
    frame_offset_ = frame_offset;
  }
  int32 GetFrameOffset() const { return frame_offset_; }

  // For a decoder that reads the log-likelihoods where they are: the device rows of `subsampled_frame` and of the frames after
  // it in the same chunk (priors and acoustic scale applied), `*num_frames` of them, row stride `*stride` floats.
  const BaseFloat *DeviceFrames(int32 subsampled_frame, int32 *num_frames, int32 *stride) {
    subsampled_frame += frame_offset_;
    EnsureFrameIsComputed(subsampled_frame);
    const int32 row = subsampled_frame - current_log_post_subsampled_offset_;
    *num_frames = d_log_post_.NumRows() - row;
    *stride = d_log_post_.Stride();
    return d_log_post_.Data() + static_cast<size_t>(row) * d_log_post_.Stride();
  }

 protected:
  inline void EnsureFrameIsComputed(int32 subsampled_frame) {
    KALDI_ASSERT(subsampled_frame >= current_log_post_subsampled_offset_ && "Frames must be accessed in order.");
    while (subsampled_frame >= current_log_post_subsampled_offset_ + current_log_post_.NumRows()) AdvanceChunk();
  }

  void AdvanceChunk() {
    const int32 C = info_.frames_per_chunk, R = info_.frames_right_context, dim = input_features_->Dim();
    const int32 num_feature_frames_ready = input_features_->NumFramesReady();
    const bool is_finished = input_features_->IsLastFrame(num_feature_frames_ready - 1);
    const int32 end_input_frame = (num_chunks_computed_ + 1) * C + R;                 // :121-134
    if (end_input_frame > num_feature_frames_ready && !is_finished)
      KALDI_ERR << "Attempt to access frame past the end of the available input";
    auto gather = [&](int32 begin, int32 end, Matrix<BaseFloat> *m) {                  // :148-161 (t < 0: the stream repeats frame 0)
      m->Resize(end - begin, dim, kUndefined);
      for (int32 i = begin; i < end; i++) {
        SubVector<BaseFloat> row(*m, i - begin);
        input_features_->GetFrame(std::min(std::max(i, 0), num_feature_frames_ready - 1), &row);
      }
    };
    Vector<BaseFloat> ivector;
    if (info_.has_ivectors) {                                                         // :166-197
      ivector.Resize(ivector_features_->Dim());
      const int32 most_recent_input_frame = num_feature_frames_ready - 1, num_ivector_frames_ready = ivector_features_->NumFramesReady();
      if (num_ivector_frames_ready > 0)
        ivector_features_->GetFrame(std::min<int32>(most_recent_input_frame, num_ivector_frames_ready - 1), &ivector);
      chunk_ivectors_.push_back(ivector);
      if (static_cast<int32>(chunk_ivectors_.size()) > ivector_rows_) chunk_ivectors_.erase(chunk_ivectors_.begin());
      Matrix<BaseFloat> rows(ivector_rows_, ivector.Dim());
      for (int32 r = 0; r < ivector_rows_; r++) {                                      // chunks n-(rows-1) .. n, chunk 0's in front
        const int32 back = ivector_rows_ - 1 - r, have = chunk_ivectors_.size();
        rows.Row(r).CopyFromVec(chunk_ivectors_[std::max(0, have - 1 - back)]);
      }
      d_iv_.Resize(ivector_rows_, ivector.Dim(), kUndefined, kStrideEqualNumCols);   // b2k reads the rows back to back
      d_iv_.CopyFromMat(rows);
    }
    int32_t ch = 0, n_out = 0, n_eos = 0, zero = 0;
    const float *ivp = info_.has_ivectors ? d_iv_.Data() : NULL;
    if (num_chunks_computed_ == 0) {                               // the right context first, a chunk's worth at a time: no output yet
      for (int32 b = 0; b < R; b += C) {
        Matrix<BaseFloat> f;
        gather(b, std::min(b + C, R), &f);
        d_feats_.Resize(0, 0); d_feats_ = f;
        const float *fp = d_feats_.Data(); int32_t nv = f.NumRows();
        const int32_t is_first = b == 0 ? 1 : 0;
        CheckNnet3(b2k_nnet_stream_run_batch(s_, 1, &ch, &fp, d_feats_.Stride(), info_.has_ivectors ? &ivp : NULL, &nv, &is_first, &zero,
                                             d_out_.Data(), NULL, d_out_.Stride(), &n_out, &n_eos, cudaStreamPerThread), "b2k_nnet_stream_run_batch");
        KALDI_ASSERT(n_out == 0);
      }
    }
    Matrix<BaseFloat> f;
    gather(num_chunks_computed_ * C + R, end_input_frame, &f);
    d_feats_.Resize(0, 0); d_feats_ = f;
    const float *fp = d_feats_.Data(); int32_t nv = C;
    const int32_t first = (num_chunks_computed_ == 0 && R == 0) ? 1 : 0;
    CheckNnet3(b2k_nnet_stream_run_batch(s_, 1, &ch, &fp, d_feats_.Stride(), info_.has_ivectors ? &ivp : NULL, &nv, &first, &zero,
                                         d_out_.Data(), NULL, d_out_.Stride(), &n_out, &n_eos, cudaStreamPerThread), "b2k_nnet_stream_run_batch");
    if (cudaStreamSynchronize(cudaStreamPerThread) != cudaSuccess) KALDI_ERR << "cudaStreamSynchronize failed";
    KALDI_ASSERT(n_out == d_out_.NumRows());
    {                                                                                  // :206-230; the chunk stays on the device too
      d_log_post_ = d_out_;
      if (info_.log_priors.Dim() != 0) d_log_post_.AddVecToRows(-1.0, info_.log_priors);
      d_log_post_.Scale(info_.opts.acoustic_scale);
      current_log_post_.Resize(d_log_post_.NumRows(), d_log_post_.NumCols(), kUndefined);
      d_log_post_.CopyToMat(&current_log_post_);
    }
    KALDI_ASSERT(current_log_post_.NumRows() == info_.frames_per_chunk / info_.opts.frame_subsampling_factor &&
                 current_log_post_.NumCols() == info_.output_dim);
    num_chunks_computed_++;
    current_log_post_subsampled_offset_ = (num_chunks_computed_ - 1) * (info_.frames_per_chunk / info_.opts.frame_subsampling_factor);
  }

  Matrix<BaseFloat> current_log_post_;
  int32 num_chunks_computed_ = 0;
  int32 current_log_post_subsampled_offset_ = -1;
  const nnet3::DecodableNnetSimpleLoopedInfo &info_;
  int32 frame_offset_ = 0;

 private:
  OnlineFeatureInterface *input_features_;
  OnlineFeatureInterface *ivector_features_;
  std::unique_ptr<ModelB2k> model_;
  b2k_nnet_stream *s_ = NULL;
  int32 ivector_rows_ = 1;
  std::vector<Vector<BaseFloat> > chunk_ivectors_;
  CuMatrix<BaseFloat> d_feats_, d_iv_, d_out_, d_log_post_;
  KALDI_DISALLOW_COPY_AND_ASSIGN(DecodableNnetLoopedOnlineBaseB2k);
};

// nnet3::DecodableNnetLoopedOnline (decodable-online-looped.h:141-160): indexes are pdf-ids + 1
class DecodableNnetLoopedOnlineB2k : public DecodableNnetLoopedOnlineBaseB2k {
 public:
  DecodableNnetLoopedOnlineB2k(const nnet3::DecodableNnetSimpleLoopedInfo &info, OnlineFeatureInterface *input_features,
                               OnlineFeatureInterface *ivector_features)
      : DecodableNnetLoopedOnlineBaseB2k(info, input_features, ivector_features) {}
  int32 NumIndices() const override { return info_.output_dim; }
  BaseFloat LogLikelihood(int32 subsampled_frame, int32 index) override {
    subsampled_frame += frame_offset_;
    EnsureFrameIsComputed(subsampled_frame);
    return current_log_post_(subsampled_frame - current_log_post_subsampled_offset_, index - 1);
  }
};

// nnet3::DecodableAmNnetLoopedOnline (decodable-online-looped.h:168-196): indexes are transition-ids
class DecodableAmNnetLoopedOnlineB2k : public DecodableNnetLoopedOnlineBaseB2k {
 public:
  DecodableAmNnetLoopedOnlineB2k(const TransitionModel &trans_model, const nnet3::DecodableNnetSimpleLoopedInfo &info,
                                 OnlineFeatureInterface *input_features, OnlineFeatureInterface *ivector_features)
      : DecodableNnetLoopedOnlineBaseB2k(info, input_features, ivector_features), trans_model_(trans_model) {}
  int32 NumIndices() const override { return trans_model_.NumTransitionIds(); }
  BaseFloat LogLikelihood(int32 subsampled_frame, int32 index) override {
    subsampled_frame += frame_offset_;
    EnsureFrameIsComputed(subsampled_frame);
    return current_log_post_(subsampled_frame - current_log_post_subsampled_offset_, trans_model_.TransitionIdToPdfFast(index));
  }

 private:
  const TransitionModel &trans_model_;
};

// nnet3::DecodableNnetSimple / DecodableAmNnetSimple (nnet3/nnet-am-decodable-simple.h:185-349), the decodables of the OFFLINE
// tools (nnet3bin/nnet3-latgen-faster.cc, nnet3-compute.cc): the utterance is cut into chunks of --frames-per-chunk outputs, each
// computed from its own window of input frames [first output - left, last output + right] (indices clamped to the utterance,
// nnet-am-decodable-simple.cc:136-166) with ONE i-vector per chunk (the one nearest the chunk's middle frame, :181-211).  The
// reference computes a chunk when the decoder first asks for one of its frames; here ALL chunks of the utterance are the lanes
// of one batched run of a window program (b2k_nnet_compile_window) at construction -- same windows, same i-vectors, same outputs.
//
// NnetSimpleComputerB2k is the part that is shared between utterances (the role CachingOptimizingCompiler plays in the
// reference's constructor arguments): the model on the device and the compiled window program.
class NnetSimpleComputerB2k {
 public:
  NnetSimpleComputerB2k(const nnet3::NnetSimpleComputationOptions &opts, const nnet3::Nnet &nnet, int32 max_chunks_per_run = 64)
      : max_lanes_(max_chunks_per_run) {
    RequireCuDevice();
    if (opts.extra_left_context != 0 || opts.extra_right_context != 0 || opts.extra_left_context_initial > 0 || opts.extra_right_context_final > 0)
      KALDI_ERR << "b2k: extra left / right context is not supported";
    if (opts.frame_subsampling_factor < 1 || opts.frames_per_chunk < 1)
      KALDI_ERR << "--frame-subsampling-factor and --frames-per-chunk must be > 0";                      // CheckAndFixConfigs :287-290
    sf_ = opts.frame_subsampling_factor;
    frames_per_chunk_ = sf_ * ((opts.frames_per_chunk + sf_ - 1) / sf_);                                  // (:294-296; the supported nets have modulus 1)
    opc_ = frames_per_chunk_ / sf_;
    model_.reset(new ModelB2k(nnet, sf_));
    b2k_nnet_compile_cfg c = model_->CompileConfig(sf_, 1.0f, false);
    CheckNnet3(b2k_nnet_model_context(&c, b2k_model_layers(model_->Handle()), model_->NumLayers(), &left_, &right_), "b2k_nnet_model_context");
    window_ = (opc_ - 1) * sf_ + 1 + left_ + right_;                                                     // :137-139 for a full chunk
    c.num_frames = window_;
    b2k_nnet_program *prog = NULL;
    CheckNnet3(b2k_nnet_compile_window(&c, left_, opc_, 1, b2k_model_layers(model_->Handle()), model_->NumLayers(),
                                       b2k_model_weights(model_->Handle()), model_->NumWeights(), &prog), "b2k_nnet_compile_window");
    const int rc = b2k_nnet_create_from_program(prog, max_lanes_, &nn_);
    b2k_nnet_program_destroy(prog);
    CheckNnet3(rc, "b2k_nnet_create_from_program");
  }
  ~NnetSimpleComputerB2k() { if (nn_) b2k_nnet_destroy(nn_); }

  int32 OutputDim() const { return model_->NumPdfs(); }
  int32 InputDim() const { return model_->FeatDim(); }
  int32 IvectorDim() const { return model_->IvectorDim(); }
  int32 FrameSubsamplingFactor() const { return sf_; }
  void GetSimpleNnetContext(int32 *left, int32 *right) const { *left = left_; *right = right_; }

  // Raw network output for every subsampled frame of the utterance ([ceil(T / sf) x OutputDim()]; priors and acoustic scale are
  // the caller's, as in DoNnetComputation :258-262)
  void Compute(const MatrixBase<BaseFloat> &feats, const VectorBase<BaseFloat> *ivector, const MatrixBase<BaseFloat> *online_ivectors,
               int32 online_ivector_period, CuMatrix<BaseFloat> *output) {
    const int32 T = feats.NumRows(), dim = feats.NumCols(), ivd = model_->IvectorDim();
    if (dim != model_->FeatDim())
      KALDI_ERR << "Neural net expects 'input' features with dimension " << model_->FeatDim() << " but you provided " << dim;       // :103-106
    const int32 given_ivd = ivector ? ivector->Dim() : (online_ivectors ? online_ivectors->NumCols() : 0);
    if (given_ivd != ivd) KALDI_ERR << "Neural net expects 'ivector' features with dimension " << ivd << " but you provided " << given_ivd;
    KALDI_ASSERT(!(ivector != NULL && online_ivectors != NULL));                                          // :52
    KALDI_ASSERT(!(online_ivectors != NULL && online_ivector_period <= 0 && "You need to set the --online-ivector-period option!"));
    const int32 num_subsampled = (T + sf_ - 1) / sf_, num_chunks = (num_subsampled + opc_ - 1) / opc_;
    output->Resize(num_subsampled, model_->NumPdfs(), kUndefined);
    for (int32 c0 = 0; c0 < num_chunks; c0 += max_lanes_) {
      const int32 n = std::min(max_lanes_, num_chunks - c0);
      Matrix<BaseFloat> win(n * window_, dim, kUndefined), ivs(std::max(1, n), std::max(1, ivd));
      for (int32 k = 0; k < n; k++) {
        const int32 first_out = (c0 + k) * opc_ * sf_;
        const int32 outs = std::min(num_subsampled - (c0 + k) * opc_, opc_), last_out = first_out + (outs - 1) * sf_;
        for (int32 i = 0; i < window_; i++) {                                                            // :150-161 (rows past the chunk's own
          const int32 t = std::min(std::max(first_out - left_ + i, 0), T - 1);                            //  window feed outputs nobody reads)
          win.Row(k * window_ + i).CopyFromVec(feats.Row(t));
        }
        if (ivector) ivs.Row(k).CopyFromVec(*ivector);                                                    // GetCurrentIvector :181-211
        else if (online_ivectors) {
          const int32 frame_to_search = first_out + (last_out - first_out) / 2;
          int32 ivector_frame = frame_to_search / online_ivector_period;
          if (ivector_frame >= online_ivectors->NumRows()) {
            const int32 margin = ivector_frame - (online_ivectors->NumRows() - 1);
            if (margin * online_ivector_period > 50)
              KALDI_ERR << "Could not get iVector for frame " << frame_to_search << ", only available till frame " << online_ivectors->NumRows()
                        << " * ivector-period=" << online_ivector_period << " (mismatched --online-ivector-period?)";
            ivector_frame = online_ivectors->NumRows() - 1;
          }
          ivs.Row(k).CopyFromVec(online_ivectors->Row(ivector_frame));
        }
      }
      d_win_ = win;
      if (ivd > 0) { d_iv_.Resize(n, ivd, kUndefined, kStrideEqualNumCols); d_iv_.CopyFromMat(ivs.RowRange(0, n)); }
      d_out_.Resize(n * opc_, model_->NumPdfs(), kUndefined);
      std::vector<const float *> in(n), iv(n);
      std::vector<float *> out(n);
      for (int32 k = 0; k < n; k++) {
        in[k] = d_win_.Data() + static_cast<size_t>(k) * window_ * d_win_.Stride();
        iv[k] = ivd > 0 ? d_iv_.Data() + static_cast<size_t>(k) * ivd : NULL;
        out[k] = d_out_.Data() + static_cast<size_t>(k) * opc_ * d_out_.Stride();
      }
      CheckNnet3(b2k_nnet_run(nn_, n, in.data(), d_win_.Stride(), ivd > 0 ? iv.data() : NULL, ivd, out.data(), d_out_.Stride(), cudaStreamPerThread),
                 "b2k_nnet_run");
      const int32 rows = std::min(n * opc_, num_subsampled - c0 * opc_);
      output->RowRange(c0 * opc_, rows).CopyFromMat(d_out_.RowRange(0, rows));
    }
  }

 private:
  int32 max_lanes_, sf_ = 1, frames_per_chunk_ = 0, opc_ = 0, left_ = 0, right_ = 0, window_ = 0;
  std::unique_ptr<ModelB2k> model_;
  b2k_nnet *nn_ = NULL;
  CuMatrix<BaseFloat> d_win_, d_iv_, d_out_;
  KALDI_DISALLOW_COPY_AND_ASSIGN(NnetSimpleComputerB2k);
};

class DecodableNnetSimpleB2k {                        // nnet-am-decodable-simple.h:185-270
 public:
  DecodableNnetSimpleB2k(const nnet3::NnetSimpleComputationOptions &opts, const VectorBase<BaseFloat> &priors,
                         const MatrixBase<BaseFloat> &feats, NnetSimpleComputerB2k *computer, const VectorBase<BaseFloat> *ivector = NULL,
                         const MatrixBase<BaseFloat> *online_ivectors = NULL, int32 online_ivector_period = 1) {
    CuMatrix<BaseFloat> out;
    computer->Compute(feats, ivector, online_ivectors, online_ivector_period, &out);
    if (priors.Dim() != 0) {                                                                            // :47,258-260
      CuVector<BaseFloat> log_priors(priors);
      log_priors.ApplyLog();
      out.AddVecToRows(-1.0, log_priors);
    }
    out.Scale(opts.acoustic_scale);                                                                     // :262
    log_post_.Resize(out.NumRows(), out.NumCols(), kUndefined);
    out.CopyToMat(&log_post_);
  }
  inline int32 NumFrames() const { return log_post_.NumRows(); }
  inline int32 OutputDim() const { return log_post_.NumCols(); }
  void GetOutputForFrame(int32 frame, VectorBase<BaseFloat> *output) { output->CopyFromVec(log_post_.Row(frame)); }
  inline BaseFloat GetOutput(int32 subsampled_frame, int32 pdf_id) { return log_post_(subsampled_frame, pdf_id); }

 private:
  Matrix<BaseFloat> log_post_;
  KALDI_DISALLOW_COPY_AND_ASSIGN(DecodableNnetSimpleB2k);
};

class DecodableAmNnetSimpleB2k : public DecodableInterface {             // nnet-am-decodable-simple.h:272-349
 public:
  DecodableAmNnetSimpleB2k(const nnet3::NnetSimpleComputationOptions &opts, const TransitionModel &trans_model,
                           const nnet3::AmNnetSimple &am_nnet, const MatrixBase<BaseFloat> &feats, NnetSimpleComputerB2k *computer,
                           const VectorBase<BaseFloat> *ivector = NULL, const MatrixBase<BaseFloat> *online_ivectors = NULL,
                           int32 online_ivector_period = 1)
      : decodable_nnet_(opts, am_nnet.Priors(), feats, computer, ivector, online_ivectors, online_ivector_period), trans_model_(trans_model) {}
  BaseFloat LogLikelihood(int32 frame, int32 transition_id) override {
    return decodable_nnet_.GetOutput(frame, trans_model_.TransitionIdToPdfFast(transition_id));
  }
  int32 NumFramesReady() const override { return decodable_nnet_.NumFrames(); }
  int32 NumIndices() const override { return trans_model_.NumTransitionIds(); }
  bool IsLastFrame(int32 frame) const override {
    KALDI_ASSERT(frame < NumFramesReady());
    return frame == NumFramesReady() - 1;
  }

 private:
  DecodableNnetSimpleB2k decodable_nnet_;
  const TransitionModel &trans_model_;
  KALDI_DISALLOW_COPY_AND_ASSIGN(DecodableAmNnetSimpleB2k);
};

// SingleUtteranceNnet3DecoderTpl (online2/online-nnet3-decoding.h:50-121), the object online2-wav-nnet3-latgen-faster drives:
//   feature_pipeline.AcceptWaveform(...); decoder.AdvanceDecoding(); if (decoder.EndpointDetected(endpoint_opts)) break; ...
//   decoder.FinalizeDecoding(); decoder.GetLattice(true, &clat);
// The decodable is DecodableAmNnetLoopedOnlineB2k (its chunks stay on the device), the decoder one lane of b2k_dec.
//   DecoderConfig  = LatticeFasterDecoderConfig (decoder/lattice-faster-decoder.h:39-99; a template parameter because that
//                    header needs OpenFst)
//   Features       = OnlineNnet2FeaturePipeline or OnlineNnet2FeaturePipelineB2k: InputFeature(), IvectorFeature(),
//                    FrameShiftInSeconds()
//   graph          = CudaFstB2k(decode_fst, &trans_model).Handle()  (b2k_kaldi_shims.h)
template <class DecoderConfig, class Features>
class SingleUtteranceNnet3DecoderB2k {
 public:
  SingleUtteranceNnet3DecoderB2k(const DecoderConfig &decoder_opts, const TransitionModel &trans_model,
                                 const nnet3::DecodableNnetSimpleLoopedInfo &info, const b2k_fst *fst, Features *features,
                                 int32 max_frames = 4096)
      : decoder_opts_(decoder_opts),
        input_feature_frame_shift_in_seconds_(features->FrameShiftInSeconds()),
        trans_model_(trans_model),
        decodable_(trans_model, info, features->InputFeature(), features->IvectorFeature()) {
    b2k_dec_cfg c;
    b2k_dec_cfg_default(&c);
    c.beam = decoder_opts.beam; c.lattice_beam = decoder_opts.lattice_beam; c.max_active = decoder_opts.max_active;
    c.min_active = decoder_opts.min_active; c.beam_delta = decoder_opts.beam_delta; c.prune_interval = decoder_opts.prune_interval;
    c.prune_scale = decoder_opts.prune_scale; c.hash_ratio = decoder_opts.hash_ratio; c.max_frames = max_frames;
    // one channel: its arenas hold the whole utterance (no interim pruning): 2000 tokens and 4000 links a frame, ~350 MB at 4096 frames
    c.max_tokens = std::max<int64_t>(c.max_tokens, static_cast<int64_t>(max_frames) * 2000);
    c.max_links = std::max<int64_t>(c.max_links, 2 * c.max_tokens);
    CheckNnet3(b2k_dec_create(fst, &c, 1, 1, &dec_), "b2k_dec_create");
    InitDecoding();
  }
  ~SingleUtteranceNnet3DecoderB2k() { b2k_dec_destroy(dec_); }

  void InitDecoding(int32 frame_offset = 0) {                                          // online-nnet3-decoding.cc:39-43
    const int32_t ch = 0;
    CheckNnet3(b2k_dec_init_decoding(dec_, &ch, 1, cudaStreamPerThread), "InitDecoding");
    frames_decoded_ = 0;
    decodable_.SetFrameOffset(frame_offset);
  }
  void AdvanceDecoding() {                                                             // :46-48: every frame that is ready
    const int32_t ch = 0;
    while (frames_decoded_ < decodable_.NumFramesReady()) {
      int32 in_chunk = 0, stride = 0;
      const float *rows = decodable_.DeviceFrames(frames_decoded_, &in_chunk, &stride);
      int32_t n = std::min(in_chunk, decodable_.NumFramesReady() - frames_decoded_);
      CheckNnet3(b2k_dec_advance_decoding_frames(dec_, &ch, &rows, &n, stride, 1, cudaStreamPerThread), "AdvanceDecoding");
      frames_decoded_ += n;
    }
  }
  void FinalizeDecoding() {                                                            // :51-53
    const int32_t ch = 0;
    CheckNnet3(b2k_dec_finalize_decoding(dec_, &ch, 1, cudaStreamPerThread), "FinalizeDecoding");
  }
  int32 NumFramesDecoded() const { return frames_decoded_; }

  // EndpointDetected(config, trans_model, frame_shift, decoder) (online-endpoint.cc:116-135) on the LIVE best path
  template <class EndpointConfig>
  bool EndpointDetected(const EndpointConfig &config) {                                // :85-92
    if (frames_decoded_ == 0) return false;
    std::vector<int32> ilabels;
    b2k_best_path_info bp;
    BestPath(false, &ilabels, NULL, NULL, NULL, &bp);
    const BaseFloat frame_shift = input_feature_frame_shift_in_seconds_ * decodable_.FrameSubsamplingFactor();
    return EndpointDetectedB2k(config, trans_model_, frame_shift, ilabels, bp.num_frames, bp.final_relative_cost);
  }
  BaseFloat FinalRelativeCost() {
    b2k_best_path_info bp;
    BestPath(false, NULL, NULL, NULL, NULL, &bp);
    return bp.final_relative_cost;
  }
  // the best path as arrays (what GetBestPath builds its linear lattice from); arc_states: the HCLG state each arc enters
  // (a token of the decoder is a (frame, state) pair: OnlineSilenceWeightingB2k tells tokens apart by it)
  void BestPath(bool use_final_probs, std::vector<int32> *ilabels, std::vector<int32> *olabels, std::vector<BaseFloat> *graph_costs,
                std::vector<BaseFloat> *acoustic_costs, b2k_best_path_info *info, std::vector<int32> *arc_states = NULL) {
    const int32_t ch = 0, cap = 3 * std::max(1, frames_decoded_) + 64;
    std::vector<int32_t> il(cap), ol(cap), st(arc_states ? cap : 0);
    std::vector<float> g(cap), a(cap);
    CheckNnet3(b2k_dec_best_path(dec_, &ch, 1, use_final_probs ? 1 : 0, cap, il.data(), ol.data(), g.data(), a.data(), NULL,
                                 arc_states ? st.data() : NULL, info, cudaStreamPerThread), "b2k_dec_best_path");
    const int32 n = info->n_arcs;
    if (ilabels) ilabels->assign(il.begin(), il.begin() + n);
    if (olabels) olabels->assign(ol.begin(), ol.begin() + n);
    if (graph_costs) graph_costs->assign(g.begin(), g.begin() + n);
    if (acoustic_costs) acoustic_costs->assign(a.begin(), a.begin() + n);
    if (arc_states) arc_states->assign(st.begin(), st.begin() + n);
  }

#ifdef B2K_HAVE_OPENFST
  // GetBestPath(end_of_utterance, Lattice*) (:78-82): LatticeFasterOnlineDecoderTpl::GetBestPath's linear lattice
  // (lattice-faster-online-decoder.cc:54-75)
  void GetBestPath(bool end_of_utterance, Lattice *best_path) {
    std::vector<int32> il, ol;
    std::vector<BaseFloat> g, a;
    b2k_best_path_info bp;
    BestPath(end_of_utterance, &il, &ol, &g, &a, &bp);
    best_path->DeleteStates();
    if (bp.end_state < 0) return;
    for (size_t k = 0; k <= il.size(); k++) best_path->AddState();
    best_path->SetStart(0);
    for (size_t k = 0; k < il.size(); k++)
      best_path->AddArc(static_cast<int32>(k), LatticeArc(il[k], ol[k], LatticeWeight(g[k], a[k]), static_cast<int32>(k + 1)));
    best_path->SetFinal(static_cast<int32>(il.size()), LatticeWeight(bp.final_cost, 0.0));
  }
#if !defined(B2K_OPENFST_IS_STANDIN)
  // GetLattice(end_of_utterance, CompactLattice*) (:56-75): raw lattice + pruned determinization at lattice_beam.  The raw
  // lattice of this decoder exists after FinalizeDecoding (final probabilities applied): end_of_utterance must be true.
  void GetLattice(bool end_of_utterance, CompactLattice *clat) {
    if (NumFramesDecoded() == 0) KALDI_ERR << "You cannot get a lattice if you decoded no frames.";
    if (!end_of_utterance) KALDI_ERR << "b2k: the lattice is available at the end of the utterance (use GetBestPath for partial results)";
    if (!decoder_opts_.determinize_lattice) KALDI_ERR << "--determinize-lattice=false option is not supported at the moment";
    FinalizeDecoding();
    b2k_raw_lattice r = {};
    CheckNnet3(b2k_dec_get_raw_lattice(dec_, 0, &r, cudaStreamPerThread), "GetRawLattice(size)");
    std::vector<int32> sf(r.num_states), sh(r.num_states), as(r.num_arcs), ad(r.num_arcs), ai(r.num_arcs), ao(r.num_arcs), fs(r.num_finals);
    std::vector<float> st(r.num_states), se(r.num_states), ag(r.num_arcs), aa(r.num_arcs), fc(r.num_finals);
    r.state_frame = sf.data(); r.state_hclg = sh.data(); r.state_tot_cost = st.data(); r.state_extra_cost = se.data();
    r.arc_src = as.data(); r.arc_dst = ad.data(); r.arc_ilabel = ai.data(); r.arc_olabel = ao.data();
    r.arc_graph_cost = ag.data(); r.arc_acoustic_cost = aa.data(); r.final_state = fs.data(); r.final_cost = fc.data();
    CheckNnet3(b2k_dec_get_raw_lattice(dec_, 0, &r, cudaStreamPerThread), "GetRawLattice");
    // DeterminizeLatticePhonePrunedWrapper(trans_model_, &raw_lat, lat_beam, clat, decoder_opts_.det_opts) (:70-73): the
    // transition model as the three per-transition-id arrays b2k_lat_determinize_phone_pruned takes
    const int32 nt = trans_model_.NumTransitionIds() + 1;
    std::vector<int32_t> phone_of(nt, 0);
    std::vector<uint8_t> self_loop(nt, 0), phone_start(nt, 0);
    for (int32 t = 1; t < nt; t++) {
      phone_of[t] = trans_model_.TransitionIdToPhone(t);
      self_loop[t] = trans_model_.IsSelfLoop(t) ? 1 : 0;
      phone_start[t] = trans_model_.TransitionIdIsStartOfPhone(t) ? 1 : 0;
    }
    b2k_clat *c = NULL;
    CheckNnet3(b2k_lat_determinize_phone_pruned(&r, decoder_opts_.lattice_beam, 0, phone_of.data(), self_loop.data(), phone_start.data(), nt,
                                                decoder_opts_.det_opts.phone_determinize ? 1 : 0, decoder_opts_.det_opts.word_determinize ? 1 : 0, &c),
               "b2k_lat_determinize_phone_pruned");
    if (decoder_opts_.det_opts.minimize) CheckNnet3(b2k_clat_minimize(c, 1.0f / 1024.0f /* MinimizeCompactLattice's default delta, fst::kDelta */), "b2k_clat_minimize");   // :1459-1465
    BatchedOnlinePipelineB2k::FillCompactLattice(c, clat);
    b2k_clat_destroy(c);
  }
#endif
#endif

 private:
  const DecoderConfig &decoder_opts_;
  BaseFloat input_feature_frame_shift_in_seconds_;
  const TransitionModel &trans_model_;
  DecodableAmNnetLoopedOnlineB2k decodable_;
  b2k_dec *dec_ = NULL;
  int32 frames_decoded_ = 0;
  KALDI_DISALLOW_COPY_AND_ASSIGN(SingleUtteranceNnet3DecoderB2k);
};

}  // namespace b2k_shim
}  // namespace kaldi

#endif  // B2K_NNET3_SHIMS_H_
