// TEST INFRASTRUCTURE ONLY (oracle/_ref build of the reference decoder).
//
// Minimal stand-in for the parts of OpenFst (1.8.4, tools/Makefile:10; absent from
// this image) that decoder/lattice-faster-decoder.{h,cc} touch, so that the
// reference's OWN decoder sources and util/hash-list-inl.h can be compiled where
// they lie.  Only containers and iterators live here -- no search logic.  The
// reference calls: Fst::Start/Final/Type/NumInputEpsilons, ArcIterator::Done/Next/
// Value, MemoryPool::Allocate/Free, and (in GetRawLattice) MutableFst::AddState/
// AddArc/SetStart/SetFinal/DeleteStates/NumStates.  Functions needed only by
// GetBestPath/GetLattice (ShortestPath, Invert, ArcSort, Connect) abort.
#ifndef B2K_ORACLE_FST_STUB_FSTLIB_H_
#define B2K_ORACLE_FST_STUB_FSTLIB_H_

#include <cstdio>
#include <cstdlib>
#include <limits>
#include <string>
#include <vector>

namespace fst {

constexpr int kNoStateId = -1;
constexpr int kNoLabel = -1;

class TropicalWeight {
 public:
  TropicalWeight() : v_(std::numeric_limits<float>::infinity()) {}
  TropicalWeight(float v) : v_(v) {}   // NOLINT (implicit, as in OpenFst)
  float Value() const { return v_; }
  static TropicalWeight Zero() { return TropicalWeight(std::numeric_limits<float>::infinity()); }
  static TropicalWeight One() { return TropicalWeight(0.0f); }
  bool operator==(const TropicalWeight &o) const { return v_ == o.v_; }
  bool operator!=(const TropicalWeight &o) const { return v_ != o.v_; }
 private:
  float v_;
};

template <class W>
struct ArcTpl {
  typedef W Weight;
  typedef int Label;
  typedef int StateId;
  Label ilabel, olabel;
  Weight weight;
  StateId nextstate;
  ArcTpl() : ilabel(0), olabel(0), nextstate(kNoStateId) {}
  ArcTpl(Label i, Label o, Weight w, StateId n) : ilabel(i), olabel(o), weight(w), nextstate(n) {}
};
typedef ArcTpl<TropicalWeight> StdArc;

// Adjacency-list FST shared by every flavour below.
template <class A>
class Fst {
 public:
  typedef A Arc;
  typedef typename A::StateId StateId;
  typedef typename A::Weight Weight;
  virtual ~Fst() {}
  virtual const std::string &Type() const { static const std::string t = "stub"; return t; }
  StateId Start() const { return start_; }
  Weight Final(StateId s) const { return final_[s]; }
  size_t NumInputEpsilons(StateId s) const { return niepsilons_[s]; }
  StateId NumStates() const { return (StateId)arcs_.size(); }
  // mutable part (the reference only uses it on the output Lattice)
  StateId AddState() {
    arcs_.emplace_back(); final_.push_back(Weight::Zero()); niepsilons_.push_back(0);
    return (StateId)arcs_.size() - 1;
  }
  void AddArc(StateId s, const A &arc) { arcs_[s].push_back(arc); if (arc.ilabel == 0) niepsilons_[s]++; }
  void SetStart(StateId s) { start_ = s; }
  void SetFinal(StateId s, Weight w) { final_[s] = w; }
  void DeleteStates() { arcs_.clear(); final_.clear(); niepsilons_.clear(); start_ = kNoStateId; }
  const std::vector<A> &ArcsOf(StateId s) const { return arcs_[s]; }
 protected:
  StateId start_ = kNoStateId;
  std::vector<std::vector<A> > arcs_;
  std::vector<Weight> final_;
  std::vector<size_t> niepsilons_;
};

template <class A>
class ConstFst : public Fst<A> {
 public:
  const std::string &Type() const override { static const std::string t = "const"; return t; }
};
template <class A>
class VectorFst : public Fst<A> {
 public:
  const std::string &Type() const override { static const std::string t = "vector"; return t; }
};
template <class A> using MutableFst = VectorFst<A>;
typedef Fst<StdArc> StdFst;
typedef VectorFst<StdArc> StdVectorFst;

template <class F>
class ArcIterator {
 public:
  typedef typename F::Arc Arc;
  typedef typename Arc::StateId StateId;
  ArcIterator(const F &fst, StateId s) : arcs_(fst.ArcsOf(s)), i_(0) {}
  bool Done() const { return i_ >= arcs_.size(); }
  void Next() { ++i_; }
  const Arc &Value() const { return arcs_[i_]; }
 private:
  const std::vector<Arc> &arcs_;
  size_t i_;
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

// fst/memory.h: fixed-size object pool (block size is only a hint in the reference too)
template <class T>
class MemoryPool {
 public:
  explicit MemoryPool(size_t block_size = 256) : block_(block_size ? block_size : 256), free_(nullptr) {}
  ~MemoryPool() { for (void *b : blocks_) std::free(b); }
  void *Allocate() {
    if (!free_) {
      const size_t sz = sizeof(T) < sizeof(void *) ? sizeof(void *) : sizeof(T);
      char *b = static_cast<char *>(std::malloc(sz * block_));
      blocks_.push_back(b);
      for (size_t i = 0; i < block_; i++) { void **p = reinterpret_cast<void **>(b + i * sz); *p = free_; free_ = p; }
    }
    void **p = static_cast<void **>(free_);
    free_ = *p;
    return p;
  }
  void Free(void *ptr) { void **p = static_cast<void **>(ptr); *p = free_; free_ = p; }
 private:
  size_t block_;
  void *free_;
  std::vector<void *> blocks_;
};

template <class A>
struct ILabelCompare {
  bool operator()(const A &a, const A &b) const { return a.ilabel < b.ilabel; }
};

[[noreturn]] inline void StubUnavailable(const char *what) {
  std::fprintf(stderr, "oracle/_ref decoder build: %s needs OpenFst, which is absent from this image\n", what);
  std::abort();
}
template <class F1, class F2> void ShortestPath(const F1 &, F2 *) { StubUnavailable("ShortestPath"); }
template <class F> void Invert(F *) { StubUnavailable("Invert"); }
template <class F, class C> void ArcSort(F *, C) { StubUnavailable("ArcSort"); }
template <class F> void Connect(F *) { StubUnavailable("Connect"); }

}  // namespace fst

#endif  // B2K_ORACLE_FST_STUB_FSTLIB_H_
