// TEST INFRASTRUCTURE ONLY (oracle/check_shims.py, -fsyntax-only): the option structs of lat/determinize-lattice-pruned.h:
// 123-150,250-295 by field name (LatticeFasterDecoderConfig and LatticeIncrementalDecoderConfig embed them; the b2k decoder shim
// reads phone_determinize / word_determinize), and the entry point the deprecated GetLattice() names.
#ifndef B2K_ORACLE_FST_STUB_TOOL_DET_LAT_PRUNED_H_
#define B2K_ORACLE_FST_STUB_TOOL_DET_LAT_PRUNED_H_
#include "itf/options-itf.h"
#include "lat/kaldi-lattice.h"
namespace fst {
struct DeterminizeLatticePrunedOptions {
  float delta = 0.0009765625f;
  int max_mem = 50000000, max_loop = 0, max_states = -1, max_arcs = -1;
  float retry_cutoff = 0.5f;
};
struct DeterminizeLatticePhonePrunedOptions {
  float delta = 0.0009765625f;
  int max_mem = 50000000;
  bool phone_determinize = true, word_determinize = true, minimize = false;
  void Register(kaldi::OptionsItf *) {}
};
template <class L, class C>
bool DeterminizeLatticePruned(const L &, double, C *, DeterminizeLatticePrunedOptions) {
  StubUnavailable("DeterminizeLatticePruned");
}
}  // namespace fst
#endif
