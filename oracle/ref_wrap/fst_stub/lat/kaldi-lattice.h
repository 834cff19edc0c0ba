// TEST INFRASTRUCTURE ONLY (oracle/_ref build of the reference decoder): the lattice
// types of lat/kaldi-lattice.h + fstext/lattice-weight.h, reduced to what
// LatticeFasterDecoderTpl::GetRawLattice writes (lattice-faster-decoder.cc:114-197).
#ifndef B2K_ORACLE_FST_STUB_KALDI_LATTICE_H_
#define B2K_ORACLE_FST_STUB_KALDI_LATTICE_H_
#include "base/kaldi-common.h"
#include "fst/fstlib.h"

namespace fst {
template <class T>
class LatticeWeightTpl {
 public:
  LatticeWeightTpl() : a_(std::numeric_limits<T>::infinity()), b_(std::numeric_limits<T>::infinity()) {}
  LatticeWeightTpl(T a, T b) : a_(a), b_(b) {}
  T Value1() const { return a_; }
  T Value2() const { return b_; }
  static LatticeWeightTpl Zero() { return LatticeWeightTpl(); }
  static LatticeWeightTpl One() { return LatticeWeightTpl(0, 0); }
 private:
  T a_, b_;
};
// placeholder for CompactLattice (only named by the deprecated GetLattice(); lives in fst:: so that the
// reference's unqualified Connect(ofst) resolves by argument-dependent lookup as it does with OpenFst)
struct CompactLatticeStub { int NumStates() const { return 0; } };
}  // namespace fst

namespace kaldi {
typedef fst::LatticeWeightTpl<BaseFloat> LatticeWeight;
typedef fst::ArcTpl<LatticeWeight> LatticeArc;
typedef fst::VectorFst<LatticeArc> Lattice;
typedef fst::CompactLatticeStub CompactLattice;
}  // namespace kaldi
#endif
