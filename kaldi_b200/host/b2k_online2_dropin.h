// b2k_online2_dropin.h -- build online2bin/online2-wav-nnet3-latgen-faster.cc against b2k WITHOUT editing it.
//
//   g++ ... -include b2k_online2_dropin.h online2bin/online2-wav-nnet3-latgen-faster.cc ... -lb2k
//
// The tool names three classes whose work b2k does on the device.  This header includes the reference's own headers first
// (their include guards make the tool's later #includes no-ops), then lets those three names resolve to adapters with the
// constructors the tool calls:
//
//   OnlineNnet2FeaturePipeline feature_pipeline(feature_info);                      online2-wav-nnet3-latgen-faster.cc:219
//   OnlineSilenceWeighting silence_weighting(trans_model, config, subsampling);     :223-226
//   SingleUtteranceNnet3Decoder decoder(decoder_opts, trans_model, decodable_info,
//                                       *decode_fst, &feature_pipeline);            :228-230
//
// Everything else in the tool -- option registration, OnlineNnet2FeaturePipelineInfo, the model, DecodableNnetSimpleLoopedInfo,
// the FST, the table readers and writers, the lattice post-processing -- stays the reference's own code.  What differs from the
// shims underneath (b2k_online2_shims.h, b2k_nnet3_shims.h) is only what a constructor signature cannot carry: the device
// tables of a feature configuration and the device copy of a decoding graph are made once per Info / per Fst object and shared
// by the utterances that follow (the tool builds both objects once, before its loops).
//
// Checked by oracle/check_shims.py: the tool's translation unit, as it lies in the reference tree, compiles with this header
// force-included (against the container's OpenFst stand-in; declarations only for the lattice library the tool calls after
// decoding).  It cannot be RUN here: there is no OpenFst and no CUDA build of Kaldi in this image.
#ifndef B2K_ONLINE2_DROPIN_H_
#define B2K_ONLINE2_DROPIN_H_

#include <map>
#include <memory>
#include <mutex>
#include <queue>                                     // online-ivector-feature.h uses std::priority_queue
#include <utility>

#include "decoder/lattice-faster-decoder.h"          // LatticeFasterDecoderConfig
#include "fstext/fstext-lib.h"
#include "lat/lattice-functions.h"
#include "online2/online-endpoint.h"
#include "online2/online-nnet2-feature-pipeline.h"
#include "online2/online-nnet3-decoding.h"

#ifndef B2K_HAVE_OPENFST
#define B2K_HAVE_OPENFST
#endif
#include "b2k_nnet3_shims.h"
#include "b2k_online2_shims.h"

namespace kaldi {
namespace b2k_dropin {

// one set of device tables per OnlineNnet2FeaturePipelineInfo object, made at the first utterance
inline const b2k_shim::FeatureTablesB2k &TablesOf(const OnlineNnet2FeaturePipelineInfo &info) {
  static std::mutex mu;
  static std::map<const OnlineNnet2FeaturePipelineInfo *, std::unique_ptr<b2k_shim::FeatureTablesB2k> > tables;
  std::lock_guard<std::mutex> lock(mu);
  std::unique_ptr<b2k_shim::FeatureTablesB2k> &t = tables[&info];
  if (!t) t.reset(new b2k_shim::FeatureTablesB2k(info));
  return *t;
}

// one device copy per decoding graph object (and transition model: the pdf of every transition-id is baked into the arcs)
inline const b2k_fst *GraphOf(const fst::Fst<fst::StdArc> &fst, const TransitionModel &trans_model) {
  static std::mutex mu;
  static std::map<std::pair<const void *, const void *>, std::unique_ptr<b2k_shim::CudaFstB2k> > graphs;
  std::lock_guard<std::mutex> lock(mu);
  std::unique_ptr<b2k_shim::CudaFstB2k> &g = graphs[std::make_pair(static_cast<const void *>(&fst), static_cast<const void *>(&trans_model))];
  if (!g) g.reset(new b2k_shim::CudaFstB2k(fst, &trans_model));
  return g->Handle();
}

class OnlineNnet2FeaturePipeline : public b2k_shim::OnlineNnet2FeaturePipelineB2k {
 public:
  explicit OnlineNnet2FeaturePipeline(const OnlineNnet2FeaturePipelineInfo &info)
      : b2k_shim::OnlineNnet2FeaturePipelineB2k(info, TablesOf(info)) {}
};

class SingleUtteranceNnet3Decoder
    : public b2k_shim::SingleUtteranceNnet3DecoderB2k<LatticeFasterDecoderConfig, OnlineNnet2FeaturePipeline> {
  typedef b2k_shim::SingleUtteranceNnet3DecoderB2k<LatticeFasterDecoderConfig, OnlineNnet2FeaturePipeline> Base;

 public:
  SingleUtteranceNnet3Decoder(const LatticeFasterDecoderConfig &decoder_opts, const TransitionModel &trans_model,
                              const nnet3::DecodableNnetSimpleLoopedInfo &info, const fst::Fst<fst::StdArc> &fst,
                              OnlineNnet2FeaturePipeline *features)
      : Base(decoder_opts, trans_model, info, GraphOf(fst, trans_model), features) {}
  // const LatticeFasterOnlineDecoderTpl<FST> &Decoder() (online-nnet3-decoding.h:112): what the silence weighting and the
  // endpointing free functions walk; here the object itself answers NumFramesDecoded() and BestPath()
  Base &Decoder() { return *this; }
};

typedef b2k_shim::OnlineSilenceWeightingB2k OnlineSilenceWeighting;

}  // namespace b2k_dropin
}  // namespace kaldi

// From here on the three names mean the adapters (whole tokens only: OnlineNnet2FeaturePipelineInfo / ...Config,
// OnlineSilenceWeightingConfig and SingleUtteranceNnet3DecoderTpl are other tokens and stay the reference's).
#include "b2k_dropin_common.h"                      // leaves nnet3::CollapseModel out: b2k takes the model as trained
#define OnlineNnet2FeaturePipeline b2k_dropin::OnlineNnet2FeaturePipeline
#define OnlineSilenceWeighting b2k_dropin::OnlineSilenceWeighting
#define SingleUtteranceNnet3Decoder b2k_dropin::SingleUtteranceNnet3Decoder

#endif  // B2K_ONLINE2_DROPIN_H_
