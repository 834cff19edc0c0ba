// TEST INFRASTRUCTURE ONLY (oracle/_ref determinizer build).  The determinizer only names PruneLattice, and only on its
// retry path (determinize-lattice-pruned.cc:1230,1285: after max_mem / max_states was hit); the oracle runs with those
// limits out of reach and aborts if the path is ever taken.
#ifndef B2K_ORACLE_FST_STUB_DET_LATTICE_FUNCTIONS_H_
#define B2K_ORACLE_FST_STUB_DET_LATTICE_FUNCTIONS_H_
#include "lat/kaldi-lattice.h"
namespace kaldi {
template <class LatType> bool PruneLattice(BaseFloat, LatType *) {
  std::fprintf(stderr, "oracle/_ref determinizer: the retry path (PruneLattice) is not available in this build\n");
  std::abort();
}
}  // namespace kaldi
#endif
