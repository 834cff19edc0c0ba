// TEST INFRASTRUCTURE ONLY (oracle/_ref determinizer build): the typedefs of lat/kaldi-lattice.h:33-50 over the
// reference's own fstext/lattice-weight.h; nothing of the table / I-O code of the real header.
#ifndef B2K_ORACLE_FST_STUB_DET_KALDI_LATTICE_H_
#define B2K_ORACLE_FST_STUB_DET_KALDI_LATTICE_H_
#include "base/kaldi-common.h"
#include "fst/fstlib.h"
#include "fstext/lattice-weight.h"
namespace kaldi {
typedef fst::LatticeWeightTpl<BaseFloat> LatticeWeight;
typedef fst::CompactLatticeWeightTpl<LatticeWeight, int32> CompactLatticeWeight;
typedef fst::CompactLatticeWeightCommonDivisorTpl<LatticeWeight, int32> CompactLatticeWeightCommonDivisor;
typedef fst::ArcTpl<LatticeWeight> LatticeArc;
typedef fst::ArcTpl<CompactLatticeWeight> CompactLatticeArc;
typedef fst::VectorFst<LatticeArc> Lattice;
typedef fst::VectorFst<CompactLatticeArc> CompactLattice;
}  // namespace kaldi
#endif
