// TEST INFRASTRUCTURE ONLY (oracle side): what ivector_wrap.cc and silence_wrap.cc both hold of an extractor and a speaker.
#ifndef B2K_ORACLE_IVECTOR_REF_TYPES_H_
#define B2K_ORACLE_IVECTOR_REF_TYPES_H_
#include <memory>
#include "feat/online-feature.h"
#include "gmm/diag-gmm.h"
#include "ivector/ivector-extractor.h"

namespace b2k_oracle {
using namespace kaldi;

struct RefIvec {
  IvectorExtractor extractor;
  DiagGmm ubm;
  Matrix<BaseFloat> lda;
};

// One speaker across utterances: what OnlineIvectorExtractorAdaptationState holds (online2/online-ivector-feature.h:218-263).
struct RefSpeaker {
  bool has = false;
  OnlineCmvnState cmvn;
  std::unique_ptr<OnlineIvectorEstimationStats> stats;
};

}  // namespace b2k_oracle
#endif
