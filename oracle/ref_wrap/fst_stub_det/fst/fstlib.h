// TEST INFRASTRUCTURE ONLY (oracle/_ref build of the reference's lattice determinizer).
//
// Stand-in for the parts of OpenFst (1.8.4 pinned at tools/Makefile:10; absent from this image) that
// lat/determinize-lattice-pruned.{h,cc}, fstext/lattice-weight.h and fstext/determinize-lattice{,-inl}.h touch, so that
// the reference's OWN determinizer can be compiled where it lies and used as the oracle of
// kaldi_b200/csrc/lattice_det.cu.  What lives here: one adjacency-list FST class (Fst = ExpandedFst = MutableFst =
// VectorFst), its iterators, property bits, the float/stream helpers the weight classes call, and four graph utilities
// the determinization wrapper calls around the algorithm (TopSort, ArcSort, Invert, Connect) written from their
// documented contracts.  The determinization algorithm itself, the semiring (lattice-weight.h) and the string
// repository are the reference's own sources.  A different directory from fst_stub/ (the decoder build) on purpose:
// that one stays minimal.
#ifndef B2K_ORACLE_FST_STUB_DET_FSTLIB_H_
#define B2K_ORACLE_FST_STUB_DET_FSTLIB_H_

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <functional>
#include <iostream>
#include <limits>
#include <list>
#include <map>
#include <memory>
#include <queue>
#include <set>
#include <sstream>
#include <string>
#include <unordered_map>
#include <unordered_set>
#include <utility>
#include <vector>

#ifndef CHECK
#define CHECK(x) do { if (!(x)) { std::fprintf(stderr, "CHECK failed: %s\n", #x); std::abort(); } } while (0)
#endif

static const std::string FST_FLAGS_fst_weight_separator = ",";
static const std::string FST_FLAGS_fst_field_separator = "\t ";
static const int FST_FLAGS_v = 0;

namespace fst {

// OpenFst's headers make these visible inside namespace fst; the reference relies on it
using std::unordered_map; using std::unordered_set; using std::vector; using std::pair; using std::string; using std::list;
using std::map; using std::set; using std::ostream; using std::istream; using std::unique_ptr;

constexpr int kNoStateId = -1;
constexpr int kNoLabel = -1;
constexpr float kDelta = 1.0F / 1024.0F;
constexpr char kStringSeparator = '_';   // fst/string-weight.h

// weight properties (fst/weight.h)
constexpr uint64_t kLeftSemiring = 0x1, kRightSemiring = 0x2, kSemiring = 0x3, kCommutative = 0x4, kIdempotent = 0x8, kPath = 0x10;
// fst properties (fst/properties.h), only the bits the reference tests
constexpr uint64_t kExpanded = 0x1, kMutable = 0x2, kError = 0x4, kILabelSorted = 0x10000000ULL, kNotILabelSorted = 0x20000000ULL,
                   kTopSorted = 0x4000000000ULL, kNotTopSorted = 0x8000000000ULL, kAcyclic = 0x800000000ULL, kCyclic = 0x400000000ULL;

enum DivideType { DIVIDE_LEFT, DIVIDE_RIGHT, DIVIDE_ANY };

template <class T> struct FloatLimits {
  static constexpr T PosInfinity() { return std::numeric_limits<T>::infinity(); }
  static constexpr T NegInfinity() { return -std::numeric_limits<T>::infinity(); }
  static constexpr T NumberBad() { return std::numeric_limits<T>::quiet_NaN(); }
};

template <class T> std::istream &ReadType(std::istream &strm, T *t) { return strm.read(reinterpret_cast<char *>(t), sizeof(T)); }
template <class T> std::ostream &WriteType(std::ostream &strm, const T &t) { return strm.write(reinterpret_cast<const char *>(&t), sizeof(T)); }

template <class T>
class TropicalWeightTpl {
 public:
  typedef T ValueType;
  TropicalWeightTpl() : v_(std::numeric_limits<T>::infinity()) {}
  TropicalWeightTpl(T v) : v_(v) {}   // NOLINT
  T Value() const { return v_; }
  static TropicalWeightTpl Zero() { return TropicalWeightTpl(std::numeric_limits<T>::infinity()); }
  static TropicalWeightTpl One() { return TropicalWeightTpl(0); }
  static const std::string &Type() { static const std::string t = "tropical"; return t; }
  bool operator==(const TropicalWeightTpl &o) const { return v_ == o.v_; }
  bool operator!=(const TropicalWeightTpl &o) const { return v_ != o.v_; }
 private:
  T v_;
};
typedef TropicalWeightTpl<float> TropicalWeight;
template <class T> TropicalWeightTpl<T> Plus(const TropicalWeightTpl<T> &a, const TropicalWeightTpl<T> &b) { return a.Value() < b.Value() ? a : b; }
template <class T> TropicalWeightTpl<T> Times(const TropicalWeightTpl<T> &a, const TropicalWeightTpl<T> &b) { return TropicalWeightTpl<T>(a.Value() + b.Value()); }
template <class T> TropicalWeightTpl<T> Divide(const TropicalWeightTpl<T> &a, const TropicalWeightTpl<T> &b, DivideType = DIVIDE_ANY) { return TropicalWeightTpl<T>(a.Value() - b.Value()); }
template <class T> bool ApproxEqual(const TropicalWeightTpl<T> &a, const TropicalWeightTpl<T> &b, float delta = kDelta) { return a.Value() <= b.Value() + delta && b.Value() <= a.Value() + delta; }
template <class T> std::ostream &operator<<(std::ostream &s, const TropicalWeightTpl<T> &w) { return s << w.Value(); }

template <class W1, class W2>
class PairWeight {
 public:
  PairWeight() {}
  PairWeight(W1 a, W2 b) : a_(a), b_(b) {}
  const W1 &Value1() const { return a_; }
  const W2 &Value2() const { return b_; }
 private:
  W1 a_; W2 b_;
};

template <class W> struct NaturalLess {
  bool operator()(const W &a, const W &b) const { return Plus(a, b) == a && a != b; }
};
template <class W1, class W2> struct WeightConvert { W2 operator()(const W1 &) const { std::abort(); } };

template <class W>
struct ArcTpl {
  typedef W Weight;
  typedef int Label;
  typedef int StateId;
  Label ilabel, olabel;
  Weight weight;
  StateId nextstate;
  ArcTpl() : ilabel(0), olabel(0), nextstate(kNoStateId) {}
  ArcTpl(Label i, Label o, Weight w, StateId n) : ilabel(i), olabel(o), weight(std::move(w)), nextstate(n) {}
  static const std::string &Type() { static const std::string t = "stub"; return t; }
};
typedef ArcTpl<TropicalWeight> StdArc;

class SymbolTable {};

template <class A>
class VectorFst {
 public:
  typedef A Arc;
  typedef typename A::StateId StateId;
  typedef typename A::Weight Weight;
  VectorFst() {}
  virtual ~VectorFst() {}
  virtual const std::string &Type() const { static const std::string t = "vector"; return t; }
  VectorFst *Copy(bool /*safe*/ = false) const { return new VectorFst(*this); }
  StateId Start() const { return start_; }
  Weight Final(StateId s) const { return final_[s]; }
  StateId NumStates() const { return (StateId)arcs_.size(); }
  size_t NumArcs(StateId s) const { return arcs_[s].size(); }
  size_t NumInputEpsilons(StateId s) const { size_t n = 0; for (auto &a : arcs_[s]) n += a.ilabel == 0; return n; }
  StateId AddState() { arcs_.emplace_back(); final_.push_back(Weight::Zero()); return (StateId)arcs_.size() - 1; }
  void AddArc(StateId s, const A &arc) { arcs_[s].push_back(arc); }
  void SetStart(StateId s) { start_ = s; }
  void SetFinal(StateId s, Weight w) { final_[s] = std::move(w); }
  void DeleteStates() { arcs_.clear(); final_.clear(); start_ = kNoStateId; }
  void DeleteArcs(StateId s) { arcs_[s].clear(); }
  void ReserveStates(size_t n) { arcs_.reserve(n); final_.reserve(n); }
  void ReserveArcs(StateId s, size_t n) { arcs_[s].reserve(n); }
  const SymbolTable *InputSymbols() const { return nullptr; }
  const SymbolTable *OutputSymbols() const { return nullptr; }
  void SetInputSymbols(const SymbolTable *) {}
  void SetOutputSymbols(const SymbolTable *) {}
  const std::vector<A> &ArcsOf(StateId s) const { return arcs_[s]; }
  std::vector<A> &MutableArcsOf(StateId s) { return arcs_[s]; }
  std::vector<std::vector<A> > &AllArcs() { return arcs_; }
  std::vector<Weight> &AllFinals() { return final_; }
  // Properties(mask, test): computed on demand for the bits the reference asks about
  uint64_t Properties(uint64_t mask, bool /*test*/) const {
    uint64_t p = kExpanded | kMutable;
    if (mask & (kILabelSorted | kNotILabelSorted)) {
      bool sorted = true;
      for (auto &v : arcs_) for (size_t i = 1; i < v.size(); i++) if (v[i].ilabel < v[i - 1].ilabel) sorted = false;
      p |= sorted ? kILabelSorted : kNotILabelSorted;
    }
    if (mask & (kTopSorted | kNotTopSorted)) {
      bool top = true;
      for (StateId s = 0; s < NumStates(); s++) for (auto &a : arcs_[s]) if (a.nextstate <= s) top = false;
      p |= top ? kTopSorted : kNotTopSorted;
    }
    return p & mask;
  }
 protected:
  StateId start_ = kNoStateId;
  std::vector<std::vector<A> > arcs_;
  std::vector<Weight> final_;
};
template <class A> using Fst = VectorFst<A>;
template <class A> using ExpandedFst = VectorFst<A>;
template <class A> using MutableFst = VectorFst<A>;
template <class A> using ConstFst = VectorFst<A>;
typedef VectorFst<StdArc> StdFst;
typedef VectorFst<StdArc> StdVectorFst;

template <class To, class From> To down_cast(From *f) { return static_cast<To>(f); }

template <class F>
class ArcIterator {
 public:
  typedef typename F::Arc Arc;
  typedef typename Arc::StateId StateId;
  ArcIterator(const F &fst, StateId s) : arcs_(&fst.ArcsOf(s)), i_(0) {}
  bool Done() const { return i_ >= arcs_->size(); }
  void Next() { ++i_; }
  const Arc &Value() const { return (*arcs_)[i_]; }
  void Reset() { i_ = 0; }
  void Seek(size_t i) { i_ = i; }
  size_t Position() const { return i_; }
 private:
  const std::vector<Arc> *arcs_;
  size_t i_;
};
template <class F>
class MutableArcIterator {
 public:
  typedef typename F::Arc Arc;
  typedef typename Arc::StateId StateId;
  MutableArcIterator(F *fst, StateId s) : fst_(fst), s_(s), i_(0) {}
  bool Done() const { return i_ >= fst_->ArcsOf(s_).size(); }
  void Next() { ++i_; }
  const Arc &Value() const { return fst_->ArcsOf(s_)[i_]; }
  void SetValue(const Arc &a) { fst_->MutableArcsOf(s_)[i_] = a; }
 private:
  F *fst_; StateId s_; size_t i_;
};
template <class F>
class StateIterator {
 public:
  typedef typename F::Arc::StateId StateId;
  explicit StateIterator(const F &fst) : n_(fst.NumStates()), s_(0) {}
  bool Done() const { return s_ >= n_; }
  void Next() { ++s_; }
  StateId Value() const { return s_; }
 private:
  StateId n_, s_;
};
template <class A> typename A::StateId CountStates(const VectorFst<A> &f) { return f.NumStates(); }

template <class A> struct ILabelCompare { bool operator()(const A &a, const A &b) const { return a.ilabel < b.ilabel; } };
template <class A> struct OLabelCompare { bool operator()(const A &a, const A &b) const { return a.olabel < b.olabel; } };

// ArcSort: stable per-state sort with the given comparator (fst/arcsort.h sorts with std::sort; a stable sort is one of
// the orders std::sort may produce for distinct keys and keeps equal-label arcs in input order)
template <class A, class C> void ArcSort(VectorFst<A> *f, C comp) { for (auto &v : f->AllArcs()) std::stable_sort(v.begin(), v.end(), comp); }
// Invert: swap input and output labels
template <class A> void Invert(VectorFst<A> *f) { for (auto &v : f->AllArcs()) for (auto &a : v) std::swap(a.ilabel, a.olabel); }
// TopSort: renumber states in a topological order (depth-first finishing order reversed, from the start state first,
// as fst/topsort.h does); false and no change if there is a cycle
template <class A> bool TopSort(VectorFst<A> *f) {
  typedef typename A::StateId S;
  const S n = f->NumStates();
  if (n == 0) return true;
  std::vector<char> color(n, 0);
  std::vector<S> finish;
  bool acyclic = true;
  auto dfs = [&](S root) {
    std::vector<std::pair<S, size_t> > st;
    st.push_back({root, 0}); color[root] = 1;
    while (!st.empty()) {
      S s = st.back().first;
      if (st.back().second < f->ArcsOf(s).size()) {
        S d = f->ArcsOf(s)[st.back().second++].nextstate;
        if (color[d] == 1) acyclic = false;
        else if (color[d] == 0) { color[d] = 1; st.push_back({d, 0}); }
      } else { color[s] = 2; finish.push_back(s); st.pop_back(); }
    }
  };
  if (f->Start() != kNoStateId) dfs(f->Start());
  for (S s = 0; s < n; s++) if (!color[s]) dfs(s);
  if (!acyclic) return false;
  std::vector<S> order(n);
  for (S i = 0; i < n; i++) order[finish[n - 1 - i]] = i;
  std::vector<std::vector<A> > arcs(n);
  std::vector<typename A::Weight> fin(n, A::Weight::Zero());
  for (S s = 0; s < n; s++) { arcs[order[s]] = std::move(f->AllArcs()[s]); fin[order[s]] = f->AllFinals()[s]; for (auto &a : arcs[order[s]]) a.nextstate = order[a.nextstate]; }
  f->AllArcs() = std::move(arcs); f->AllFinals() = std::move(fin);
  if (f->Start() != kNoStateId) f->SetStart(order[f->Start()]);
  return true;
}
// Connect: remove states not both accessible from the start and co-accessible to a final state; survivors keep
// their relative order
template <class A> void Connect(VectorFst<A> *f) {
  typedef typename A::StateId S;
  const S n = f->NumStates();
  if (n == 0) return;
  std::vector<char> acc(n, 0), coacc(n, 0);
  std::vector<S> st;
  if (f->Start() != kNoStateId) { acc[f->Start()] = 1; st.push_back(f->Start()); }
  while (!st.empty()) { S s = st.back(); st.pop_back(); for (auto &a : f->ArcsOf(s)) if (!acc[a.nextstate]) { acc[a.nextstate] = 1; st.push_back(a.nextstate); } }
  std::vector<std::vector<S> > rev(n);
  for (S s = 0; s < n; s++) for (auto &a : f->ArcsOf(s)) rev[a.nextstate].push_back(s);
  for (S s = 0; s < n; s++) if (f->Final(s) != A::Weight::Zero()) { coacc[s] = 1; st.push_back(s); }
  while (!st.empty()) { S s = st.back(); st.pop_back(); for (S p : rev[s]) if (!coacc[p]) { coacc[p] = 1; st.push_back(p); } }
  std::vector<S> map(n, kNoStateId);
  S k = 0;
  for (S s = 0; s < n; s++) if (acc[s] && coacc[s]) map[s] = k++;
  std::vector<std::vector<A> > arcs(k);
  std::vector<typename A::Weight> fin(k, A::Weight::Zero());
  for (S s = 0; s < n; s++) if (map[s] != kNoStateId) {
    fin[map[s]] = f->Final(s);
    for (auto &a : f->ArcsOf(s)) if (map[a.nextstate] != kNoStateId) { A b = a; b.nextstate = map[a.nextstate]; arcs[map[s]].push_back(b); }
  }
  const S start = f->Start() != kNoStateId ? map[f->Start()] : kNoStateId;
  f->AllArcs() = std::move(arcs); f->AllFinals() = std::move(fin);
  f->SetStart(start);
}

// names fstext/openfst_compat.h mentions
struct ArcMapFstOptions {};
template <class A, class B, class C> class ArcMapFst {};

}  // namespace fst
#endif  // B2K_ORACLE_FST_STUB_DET_FSTLIB_H_
