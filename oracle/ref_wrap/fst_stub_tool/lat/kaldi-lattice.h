// TEST INFRASTRUCTURE ONLY (oracle/check_shims.py, -fsyntax-only): the lattice types of lat/kaldi-lattice.h +
// fstext/lattice-weight.h as far as online2bin/online2-wav-nnet3-latgen-faster.cc and the b2k shims name them, over the
// container stand-in of OpenFst (../fst_stub/fst/fstlib.h).  Containers and declarations only; nothing here is ever linked.
#ifndef B2K_ORACLE_FST_STUB_TOOL_KALDI_LATTICE_H_
#define B2K_ORACLE_FST_STUB_TOOL_KALDI_LATTICE_H_
#include <iostream>
#include <string>
#include <vector>
#include "base/kaldi-common.h"
#include "util/kaldi-table.h"
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
template <class W, class I>
class CompactLatticeWeightTpl {
 public:
  CompactLatticeWeightTpl() {}
  CompactLatticeWeightTpl(const W &w, const std::vector<I> &s) : w_(w), s_(s) {}
  const W &Weight() const { return w_; }
  const std::vector<I> &String() const { return s_; }
  static CompactLatticeWeightTpl Zero() { return CompactLatticeWeightTpl(W::Zero(), std::vector<I>()); }
  static CompactLatticeWeightTpl One() { return CompactLatticeWeightTpl(W::One(), std::vector<I>()); }
 private:
  W w_;
  std::vector<I> s_;
};
struct CompactLatticeStub { int NumStates() const { return 0; } };      // named by the decoder stand-in's headers
}  // namespace fst

namespace kaldi {
typedef fst::LatticeWeightTpl<BaseFloat> LatticeWeight;
typedef fst::CompactLatticeWeightTpl<LatticeWeight, int32> CompactLatticeWeight;
typedef fst::ArcTpl<LatticeWeight> LatticeArc;
typedef fst::ArcTpl<CompactLatticeWeight> CompactLatticeArc;
typedef fst::VectorFst<LatticeArc> Lattice;
typedef fst::VectorFst<CompactLatticeArc> CompactLattice;

class CompactLatticeHolder {                                             // lat/kaldi-lattice.h:90-130, declarations
 public:
  typedef CompactLattice T;
  CompactLatticeHolder();
  static bool Write(std::ostream &os, bool binary, const T &t);
  bool Read(std::istream &is);
  static bool IsReadInBinary() { return true; }
  T &Value();
  void Clear();
  void Swap(CompactLatticeHolder *other);
  bool ExtractRange(const CompactLatticeHolder &other, const std::string &range);
  ~CompactLatticeHolder();
};
typedef TableWriter<CompactLatticeHolder> CompactLatticeWriter;
class LatticeHolder {                                                    // lat/kaldi-lattice.h:132-170, declarations
 public:
  typedef Lattice T;
  LatticeHolder();
  static bool Write(std::ostream &os, bool binary, const T &t);
  bool Read(std::istream &is);
  static bool IsReadInBinary() { return true; }
  T &Value();
  void Clear();
  void Swap(LatticeHolder *other);
  bool ExtractRange(const LatticeHolder &other, const std::string &range);
  ~LatticeHolder();
};
typedef TableWriter<LatticeHolder> LatticeWriter;
}  // namespace kaldi
#endif
