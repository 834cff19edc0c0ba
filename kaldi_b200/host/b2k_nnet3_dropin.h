// b2k_nnet3_dropin.h -- build the offline nnet3 tools against b2k WITHOUT editing them.
//
//   g++ ... -include b2k_nnet3_dropin.h nnet3bin/nnet3-compute.cc ... -lb2k
//   g++ ... -include b2k_nnet3_dropin.h nnet3bin/nnet3-latgen-faster.cc ... -lb2k
//
// Those tools run the network through nnet3::DecodableNnetSimple / DecodableAmNnetSimple (nnet3/nnet-am-decodable-simple.h:
// 185-349), built per utterance with a CachingOptimizingCompiler that the tool keeps across utterances:
//
//   DecodableNnetSimple nnet_computer(opts, nnet, priors, features, &compiler, ivector, online_ivectors, period);   nnet3-compute.cc:163
//   DecodableAmNnetSimple nnet_decodable(decodable_opts, trans_model, am_nnet, features, ivector, online_ivectors,
//                                        online_ivector_period, &compiler);                                         nnet3-latgen-faster.cc:174
//
// This header includes the reference's header first and then lets the two names resolve to adapters with those constructors
// over b2k_shim::DecodableNnetSimpleB2k / DecodableAmNnetSimpleB2k (b2k_nnet3_shims.h: all chunks of the utterance as the lanes of
// one batched run on the device).  What the reference's compiler cache does for it -- compile once, run many utterances -- the
// adapters get from one NnetSimpleComputerB2k per Nnet object (and option values), made at the first utterance; the tool's
// CachingOptimizingCompiler is constructed as written and not used.  The decoder of nnet3-latgen-faster stays the reference's
// CPU LatticeFasterDecoder reading log-likelihoods through DecodableInterface (the batched GPU decoder is behind the CUDA tools).
//
// Checked by oracle/check_shims.py: both tools' translation units as they lie in the reference tree compile with this header
// force-included (HAVE_CUDA=1, the container's OpenFst stand-in with declarations of the lattice library).  No run here.
#ifndef B2K_NNET3_DROPIN_H_
#define B2K_NNET3_DROPIN_H_

#include <map>
#include <memory>
#include <mutex>
#include <tuple>

#include "nnet3/nnet-am-decodable-simple.h"
#include "nnet3/nnet-optimize.h"                     // CachingOptimizingCompiler

#include "b2k_nnet3_shims.h"

namespace kaldi {
namespace nnet3 {
namespace b2k_nnet3_dropin {

// one compiled window program + device weights per (Nnet object, the options that shape the program)
inline b2k_shim::NnetSimpleComputerB2k *ComputerOf(const NnetSimpleComputationOptions &opts, const Nnet &nnet) {
  typedef std::tuple<const Nnet *, int32, int32, int32, int32, int32, int32> Key;
  static std::mutex mu;
  static std::map<Key, std::unique_ptr<b2k_shim::NnetSimpleComputerB2k> > computers;
  std::lock_guard<std::mutex> lock(mu);
  std::unique_ptr<b2k_shim::NnetSimpleComputerB2k> &c =
      computers[Key(&nnet, opts.frames_per_chunk, opts.frame_subsampling_factor, opts.extra_left_context, opts.extra_right_context,
                    opts.extra_left_context_initial, opts.extra_right_context_final)];
  if (!c) c.reset(new b2k_shim::NnetSimpleComputerB2k(opts, nnet));
  return c.get();
}

class DecodableNnetSimple : public b2k_shim::DecodableNnetSimpleB2k {
 public:
  DecodableNnetSimple(const NnetSimpleComputationOptions &opts, const Nnet &nnet, const VectorBase<BaseFloat> &priors,
                      const MatrixBase<BaseFloat> &feats, CachingOptimizingCompiler * /*compiler*/,
                      const VectorBase<BaseFloat> *ivector = NULL, const MatrixBase<BaseFloat> *online_ivectors = NULL,
                      int32 online_ivector_period = 1)
      : b2k_shim::DecodableNnetSimpleB2k(opts, priors, feats, ComputerOf(opts, nnet), ivector, online_ivectors, online_ivector_period) {}
};

class DecodableAmNnetSimple : public b2k_shim::DecodableAmNnetSimpleB2k {
 public:
  DecodableAmNnetSimple(const NnetSimpleComputationOptions &opts, const TransitionModel &trans_model, const AmNnetSimple &am_nnet,
                        const MatrixBase<BaseFloat> &feats, const VectorBase<BaseFloat> *ivector = NULL,
                        const MatrixBase<BaseFloat> *online_ivectors = NULL, int32 online_ivector_period = 1,
                        CachingOptimizingCompiler * /*compiler*/ = NULL)
      : b2k_shim::DecodableAmNnetSimpleB2k(opts, trans_model, am_nnet, feats, ComputerOf(opts, am_nnet.GetNnet()), ivector, online_ivectors,
                                           online_ivector_period) {}
};

}  // namespace b2k_nnet3_dropin
}  // namespace nnet3
}  // namespace kaldi

// From here on the two names mean the adapters (DecodableNnetSimpleLooped... and DecodableAmNnetSimpleParallel are other tokens).
#include "b2k_dropin_common.h"                      // leaves nnet3::CollapseModel out: b2k takes the model as trained
#define DecodableNnetSimple b2k_nnet3_dropin::DecodableNnetSimple
#define DecodableAmNnetSimple b2k_nnet3_dropin::DecodableAmNnetSimple

#endif  // B2K_NNET3_DROPIN_H_
