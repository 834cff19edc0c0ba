// TEST INFRASTRUCTURE ONLY: option structs named by LatticeFasterDecoderConfig
// (lattice-faster-decoder.h:61) and the entry point of the deprecated GetLattice().
#ifndef B2K_ORACLE_FST_STUB_DET_LAT_PRUNED_H_
#define B2K_ORACLE_FST_STUB_DET_LAT_PRUNED_H_
#include "itf/options-itf.h"
#include "lat/kaldi-lattice.h"
namespace fst {
struct DeterminizeLatticePrunedOptions { int max_mem = 50000000; };
struct DeterminizeLatticePhonePrunedOptions {
  int max_mem = 50000000;
  void Register(kaldi::OptionsItf *) {}
};
template <class L, class C>
bool DeterminizeLatticePruned(const L &, double, C *, DeterminizeLatticePrunedOptions) {
  StubUnavailable("DeterminizeLatticePruned");
}
}  // namespace fst
#endif
