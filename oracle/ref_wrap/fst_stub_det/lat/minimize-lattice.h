// TEST INFRASTRUCTURE ONLY (oracle/_ref determinizer build): named by determinize-lattice-pruned.cc:1462-1464 under
// opts.minimize (false by default, DeterminizeLatticePhonePrunedOptions); the oracle never sets it.
#ifndef B2K_ORACLE_FST_STUB_DET_MINPUSH_H_
#define B2K_ORACLE_FST_STUB_DET_MINPUSH_H_
#include "lat/kaldi-lattice.h"
namespace fst {
template <class Weight, class IntType> bool MinimizeCompactLattice(MutableFst<ArcTpl<CompactLatticeWeightTpl<Weight, IntType> > > *, float = kDelta) { std::abort(); }
template <class Weight, class IntType> bool PushCompactLatticeStrings(MutableFst<ArcTpl<CompactLatticeWeightTpl<Weight, IntType> > > *) { std::abort(); }
template <class Weight, class IntType> bool PushCompactLatticeWeights(MutableFst<ArcTpl<CompactLatticeWeightTpl<Weight, IntType> > > *) { std::abort(); }
}  // namespace fst
#endif
