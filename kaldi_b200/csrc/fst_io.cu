// fst_io.cu — host-only reader of OpenFst binary files ("vector" and "const" FSTs over StdArc) into the CSR view that
// b2k_fst_create takes: what CudaFst::CudaFst does with StateIterator/ArcIterator over HCLG.fst
// (cudadecoder/cuda-fst.cc:57-198), without OpenFst.
//
// PARITY UNPINNED: OpenFst is neither vendored under the reference tree nor present in this image.  The layout below
// is the published one (fst/fst.h FstHeader::Read: magic 2125659606, type string, arc type string, version, flags,
// properties, start, numstates, numarcs; optional symbol tables; vector-fst.h: per state {float final, int64 narcs,
// arcs {int32 ilabel, int32 olabel, float weight, int32 nextstate}}; const-fst.h: state table {float final,
// uint32 pos, narcs, niepsilons, noepsilons}, then the arc table, both 16-byte aligned when the aligned flag is set
// or the file version is 1).  It is checked against kaldi_io.read_openfst / write_openfst only
// (tests/test_fst_io_cpp.py).
#include <cstdio>
#include <cstring>
#include <limits>
#include <stdexcept>
#include <string>
#include <vector>

#include "common.cuh"

struct b2k_fst_file {
  int32_t num_states = 0, start = -1;
  std::string fst_type;
  std::vector<int32_t> offsets, ilabel, olabel, nextstate;
  std::vector<float> weight, final_cost;
};

namespace {

struct Bytes {
  std::vector<unsigned char> d;
  size_t p = 0;
  void need(size_t n) const { if (p + n > d.size()) throw std::runtime_error("unexpected end of file"); }
  template <typename T> T get() { need(sizeof(T)); T v; memcpy(&v, &d[p], sizeof(T)); p += sizeof(T); return v; }
  std::string str() {
    const int32_t n = get<int32_t>();
    if (n < 0 || n > 4096) throw std::runtime_error("bad string length");
    need((size_t)n);
    std::string s((const char *)&d[p], (size_t)n);
    p += (size_t)n;
    return s;
  }
  void align16() { p = (p + 15) / 16 * 16; }
  void skip_symbol_table() {                       // symbol-table.cc SymbolTableImpl::Read
    if (get<int32_t>() != 2125658996) throw std::runtime_error("bad symbol table magic");
    str();
    get<int64_t>();                                // available key
    const int64_t size = get<int64_t>();
    if (size < 0) throw std::runtime_error("bad symbol table size");
    for (int64_t i = 0; i < size; i++) { str(); get<int64_t>(); }
  }
};

const int kHasISymbols = 1, kHasOSymbols = 2, kIsAligned = 4;

}  // namespace

extern "C" {

int b2k_fst_file_read(const char *path, b2k_fst_file **out) {
  if (!path || !out) return b2k::set_error(B2K_ERR_INVALID, "b2k_fst_file_read: bad args");
  b2k_fst_file *F = new b2k_fst_file();
  try {
    Bytes b;
    FILE *f = fopen(path, "rb");
    if (!f) throw std::runtime_error(std::string("cannot open ") + path);
    fseek(f, 0, SEEK_END);
    const long n = ftell(f);
    fseek(f, 0, SEEK_SET);
    b.d.resize((size_t)n);
    const bool ok = n <= 0 || fread(b.d.data(), 1, (size_t)n, f) == (size_t)n;
    fclose(f);
    if (!ok) throw std::runtime_error("short read");
    if (b.get<int32_t>() != 2125659606) throw std::runtime_error("not an OpenFst binary file (bad magic number)");
    F->fst_type = b.str();
    const std::string arc_type = b.str();
    const int32_t version = b.get<int32_t>(), flags = b.get<int32_t>();
    b.get<uint64_t>();                             // properties
    const int64_t start = b.get<int64_t>(), nstates = b.get<int64_t>(), narcs = b.get<int64_t>();
    if (arc_type != "standard") throw std::runtime_error("arc type " + arc_type + " is not supported (only StdArc)");
    if (flags & kHasISymbols) b.skip_symbol_table();
    if (flags & kHasOSymbols) b.skip_symbol_table();
    const int64_t lim = std::numeric_limits<int32_t>::max();
    if (F->fst_type == "vector") {
      F->offsets.push_back(0);
      int64_t s = 0, total = 0;
      while ((nstates < 0 || s < nstates) && b.p + 12 <= b.d.size()) {
        const float fw = b.get<float>();
        const int64_t na = b.get<int64_t>();
        if (na < 0 || total + na > lim) throw std::runtime_error("bad arc count");
        b.need((size_t)na * 16);
        for (int64_t a = 0; a < na; a++) {
          F->ilabel.push_back(b.get<int32_t>()); F->olabel.push_back(b.get<int32_t>());
          F->weight.push_back(b.get<float>()); F->nextstate.push_back(b.get<int32_t>());
        }
        total += na;
        F->final_cost.push_back(fw);
        F->offsets.push_back((int32_t)total);
        s++;
      }
      if (nstates >= 0 && s != nstates) throw std::runtime_error("vector FST: fewer states than the header announces");
    } else if (F->fst_type == "const") {
      if (nstates < 0 || narcs < 0 || nstates > lim || narcs > lim) throw std::runtime_error("const FST: bad sizes");
      const bool aligned = (flags & kIsAligned) || version == 1;
      if (aligned) b.align16();
      b.need((size_t)nstates * 20);
      F->final_cost.resize((size_t)nstates);
      F->offsets.resize((size_t)nstates + 1);
      int64_t expect = 0;
      for (int64_t s = 0; s < nstates; s++) {
        F->final_cost[s] = b.get<float>();
        const uint32_t pos = b.get<uint32_t>(), na = b.get<uint32_t>();
        b.get<uint32_t>(); b.get<uint32_t>();
        if ((int64_t)pos != expect) throw std::runtime_error("const FST: state table is not contiguous");
        F->offsets[s] = (int32_t)pos;
        expect += na;
      }
      if (expect != narcs) throw std::runtime_error("const FST: state table does not cover the arc table");
      F->offsets[(size_t)nstates] = (int32_t)narcs;
      if (aligned) b.align16();
      b.need((size_t)narcs * 16);
      F->ilabel.resize((size_t)narcs); F->olabel.resize((size_t)narcs); F->weight.resize((size_t)narcs); F->nextstate.resize((size_t)narcs);
      for (int64_t a = 0; a < narcs; a++) {
        F->ilabel[a] = b.get<int32_t>(); F->olabel[a] = b.get<int32_t>(); F->weight[a] = b.get<float>(); F->nextstate[a] = b.get<int32_t>();
      }
    } else {
      throw std::runtime_error("FST type " + F->fst_type + " is not supported (vector, const)");
    }
    F->num_states = (int32_t)F->final_cost.size();
    F->start = (int32_t)start;
    if (F->num_states == 0 || F->start < 0 || F->start >= F->num_states) throw std::runtime_error("FST without a start state");
    for (int32_t ns : F->nextstate) if (ns < 0 || ns >= F->num_states) throw std::runtime_error("arc to a state outside the FST");
  } catch (const std::exception &e) {
    delete F;
    return b2k::set_error(B2K_ERR_INVALID, "b2k_fst_file_read", e.what());
  }
  *out = F;
  return B2K_OK;
}

int b2k_fst_file_destroy(b2k_fst_file *f) { delete f; return B2K_OK; }

int b2k_fst_file_csr(const b2k_fst_file *f, b2k_fst_csr *csr, int32_t *is_const) {
  if (!f || !csr) return b2k::set_error(B2K_ERR_INVALID, "b2k_fst_file_csr: bad args");
  memset(csr, 0, sizeof(*csr));
  csr->num_states = f->num_states; csr->start = f->start;
  csr->offsets = f->offsets.data(); csr->ilabel = f->ilabel.data(); csr->olabel = f->olabel.data();
  csr->weight = f->weight.data(); csr->nextstate = f->nextstate.data(); csr->final_cost = f->final_cost.data();
  csr->tid2pdf = nullptr; csr->num_tids = 0;
  if (is_const) *is_const = f->fst_type == "const";
  return B2K_OK;
}

int b2k_fst_create_from_file(const b2k_fst_file *f, const int32_t *tid2pdf, int32_t num_tids, b2k_fst **out) {
  if (!f || !out || (tid2pdf && num_tids <= 0)) return b2k::set_error(B2K_ERR_INVALID, "b2k_fst_create_from_file: bad args");
  b2k_fst_csr csr;
  b2k_fst_file_csr(f, &csr, nullptr);
  csr.tid2pdf = tid2pdf; csr.num_tids = tid2pdf ? num_tids : 0;
  return b2k_fst_create(&csr, out);
}

}  // extern "C"
