// TEST INFRASTRUCTURE ONLY (oracle/check_shims.py, -fsyntax-only): what online2bin/online2-wav-nnet3-latgen-faster.cc uses of
// fstext-lib.h and of OpenFst's symbol table: declarations (fstext/kaldi-fst-io.h:49, fst/symbol-table.h).
#ifndef B2K_ORACLE_FST_STUB_TOOL_FSTEXT_LIB_H_
#define B2K_ORACLE_FST_STUB_TOOL_FSTEXT_LIB_H_
#include <iostream>
#include <string>
#include "fst/fstlib.h"
#include "fst/symbol-table.h"
namespace fst {
Fst<StdArc> *ReadFstKaldiGeneric(std::string rxfilename, bool throw_on_err = true);
class VectorFstHolder {                              // fstext/kaldi-fst-io.h:107-150 (tables of decoding graphs), declarations
 public:
  typedef VectorFst<StdArc> T;
  VectorFstHolder();
  static bool Write(std::ostream &os, bool binary, const T &t);
  bool Read(std::istream &is);
  static bool IsReadInBinary() { return true; }
  T &Value();
  void Clear();
  void Swap(VectorFstHolder *other);
  bool ExtractRange(const VectorFstHolder &other, const std::string &range);
  ~VectorFstHolder();
};
}  // namespace fst
#endif
