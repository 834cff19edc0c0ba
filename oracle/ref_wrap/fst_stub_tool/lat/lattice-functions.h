// TEST INFRASTRUCTURE ONLY (oracle/check_shims.py, -fsyntax-only): DECLARATIONS of the lattice library calls that
// online2bin/online2-wav-nnet3-latgen-faster.cc and online2-tcp-nnet3-decode-faster.cc make after decoding
// (lat/lattice-functions.h:73,243, fstext/lattice-utils.h:91,125,142, fstext/fstext-utils.h:131, OpenFst's TopSort); the real
// ones need OpenFst.
#ifndef B2K_ORACLE_FST_STUB_TOOL_LATTICE_FUNCTIONS_H_
#define B2K_ORACLE_FST_STUB_TOOL_LATTICE_FUNCTIONS_H_
#include <vector>
#include "lat/kaldi-lattice.h"
namespace fst {
template <class WIn, class WOut, class Int>
void ConvertLattice(const VectorFst<ArcTpl<CompactLatticeWeightTpl<WIn, Int> > > &ifst, VectorFst<ArcTpl<WOut> > *ofst);
template <class Arc, class I>
bool GetLinearSymbolSequence(const Fst<Arc> &fst, std::vector<I> *isymbols_out, std::vector<I> *osymbols_out,
                             typename Arc::Weight *tot_weight_out);
std::vector<std::vector<double> > AcousticLatticeScale(double acwt);
template <class W, class Int>
void ScaleLattice(const std::vector<std::vector<double> > &scale, VectorFst<ArcTpl<CompactLatticeWeightTpl<W, Int> > > *fst);
template <class A> bool TopSort(VectorFst<A> *fst);
}  // namespace fst
namespace kaldi {
using fst::ConvertLattice;
using fst::GetLinearSymbolSequence;
using fst::AcousticLatticeScale;
using fst::ScaleLattice;
int32 LatticeStateTimes(const Lattice &lat, std::vector<int32> *times);
void CompactLatticeShortestPath(const CompactLattice &clat, CompactLattice *shortest_path);
}  // namespace kaldi
#endif
