// oracle/ref_wrap/nnet_stubs.cc — TEST INFRASTRUCTURE ONLY.
// nnet3/nnet-utils.cc (RecomputeStats, a training utility) references
// NnetComputeProb, whose implementation (nnet-diagnostics.cc) needs NnetExample
// I/O that pulls in OpenFst-dependent code.  The forward path never calls it;
// these stubs only satisfy the dynamic linker.
#include "nnet3/nnet-diagnostics.h"
namespace kaldi { namespace nnet3 {
NnetComputeProb::NnetComputeProb(const NnetComputeProbOptions &config, Nnet *nnet)
    : config_(config), nnet_(*nnet), deriv_nnet_owned_(false), deriv_nnet_(NULL), compiler_(*nnet),
      num_minibatches_processed_(0) { KALDI_ERR << "NnetComputeProb is not available in the oracle build"; }
NnetComputeProb::~NnetComputeProb() {}
void NnetComputeProb::Compute(const NnetExample &) { KALDI_ERR << "not available"; }
bool NnetComputeProb::PrintTotalStats() const { return false; }
}}
