// TEST INFRASTRUCTURE ONLY (oracle/_ref determinizer build): the two helpers of fstext/fstext-utils.h and
// fstext/lattice-utils.h that determinize-lattice-pruned.cc calls, written from their documented contracts
// (fstext-utils.h:45 "Returns the highest numbered input symbol id of the FST"; lattice-utils.h:43-60 ConvertLattice:
// "takes a lattice ... and produces the compact form" — one compact arc per arc, the output label moved into a
// one-element string, invert=false keeps the input label as the compact arc's label).
#ifndef B2K_ORACLE_FST_STUB_DET_FSTEXT_UTILS_H_
#define B2K_ORACLE_FST_STUB_DET_FSTEXT_UTILS_H_
#include "fst/fstlib.h"
#include "fstext/lattice-weight.h"
namespace fst {
template <class Arc> typename Arc::Label HighestNumberedInputSymbol(const Fst<Arc> &fst) {
  typename Arc::Label ans = 0;
  for (typename Arc::StateId s = 0; s < fst.NumStates(); s++) for (auto &a : fst.ArcsOf(s)) ans = std::max(ans, a.ilabel);
  return ans;
}
template <class Weight, class Int>
void ConvertLattice(const ExpandedFst<ArcTpl<Weight> > &ifst, MutableFst<ArcTpl<CompactLatticeWeightTpl<Weight, Int> > > *ofst, bool invert = true) {
  typedef ArcTpl<CompactLatticeWeightTpl<Weight, Int> > CArc;
  typedef CompactLatticeWeightTpl<Weight, Int> CW;
  ofst->DeleteStates();
  for (int s = 0; s < ifst.NumStates(); s++) ofst->AddState();
  ofst->SetStart(ifst.Start());
  for (int s = 0; s < ifst.NumStates(); s++) {
    if (ifst.Final(s) != Weight::Zero()) ofst->SetFinal(s, CW(ifst.Final(s), std::vector<Int>()));
    for (auto &a : ifst.ArcsOf(s)) {
      const int label = invert ? a.olabel : a.ilabel, other = invert ? a.ilabel : a.olabel;
      std::vector<Int> str;
      if (other != 0) str.push_back(other);
      ofst->AddArc(s, CArc(label, label, CW(a.weight, str), a.nextstate));
    }
  }
}
}  // namespace fst
#endif
