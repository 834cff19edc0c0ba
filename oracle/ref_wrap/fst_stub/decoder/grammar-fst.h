// TEST INFRASTRUCTURE ONLY: lattice-faster-decoder.cc:1010-1017 explicitly instantiates the
// decoder for the grammar-FST wrappers; give it two distinct FST types to instantiate on.
#ifndef B2K_ORACLE_FST_STUB_GRAMMAR_FST_H_
#define B2K_ORACLE_FST_STUB_GRAMMAR_FST_H_
#include "fst/fstlib.h"
namespace fst {
class ConstGrammarFst : public Fst<StdArc> {};
class VectorGrammarFst : public Fst<StdArc> {};
}  // namespace fst
#endif
