// b2k_dropin_common.h -- shared by the drop-in headers (b2k_online2_dropin.h, b2k_nnet3_dropin.h, b2k_cuda_pipeline_dropin.h).
//
// Every inference tool of the reference reads the model and then calls
//     SetBatchnormTestMode(true, &nnet); SetDropoutTestMode(true, &nnet); nnet3::CollapseModel(nnet3::CollapseModelConfig(), &nnet);
// (online2-wav-nnet3-latgen-faster.cc:163-165, cuda-bin-tools.h ReadModels, nnet3-compute.cc, ...).  CollapseModel merges
// components (batch-norm into the affine next to it, two affines into one: nnet3/nnet-utils.cc:1700-2110) so that the CPU /
// cudamatrix executor runs fewer commands.  b2k does that folding itself, when it compiles its program from the model AS
// TRAINED -- with the arithmetic of the collapsed network: the parity tests compare against the reference's collapsed forward --
// and its reader takes the xconfig structure, not merged components (a merged network is refused with a message, never
// misread).  So in a tool built with a drop-in header the CollapseModel call leaves the network alone; the test-mode calls stay.
// The reference's own objects that the tool still builds from the network (DecodableNnetSimpleLoopedInfo: options, priors,
// the Nnet pointer) work on the uncollapsed network as they do on the collapsed one.
#ifndef B2K_DROPIN_COMMON_H_
#define B2K_DROPIN_COMMON_H_

#include "nnet3/nnet-utils.h"

namespace kaldi {
namespace nnet3 {
inline void B2kLeaveModelAsTrained(const CollapseModelConfig & /*config*/, Nnet * /*nnet*/) {}
}  // namespace nnet3
}  // namespace kaldi

#define CollapseModel B2kLeaveModelAsTrained       // CollapseModelConfig is another token and stays the reference's struct

#endif  // B2K_DROPIN_COMMON_H_
