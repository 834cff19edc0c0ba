// TEST INFRASTRUCTURE ONLY (oracle/check_shims.py, -fsyntax-only): OpenFst's symbol table by declaration.
#ifndef B2K_ORACLE_FST_STUB_TOOL_SYMBOL_TABLE_H_
#define B2K_ORACLE_FST_STUB_TOOL_SYMBOL_TABLE_H_
#include <cstdint>
#include <string>
namespace fst {
class SymbolTable {
 public:
  static SymbolTable *ReadText(const std::string &filename);
  std::string Find(int64_t key) const;
  int64_t Find(const std::string &symbol) const;
  int64_t AvailableKey() const;
  SymbolTable *Copy() const;
};
}  // namespace fst
#endif
