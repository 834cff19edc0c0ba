// TEST INFRASTRUCTURE ONLY.  C entry points around the reference's OWN lattice determinizer
// (lat/determinize-lattice-pruned.cc, compiled where it lies against the stand-in of fst_stub_det/): a raw lattice in
// the flat form of b2k_raw_lattice goes through DeterminizeLatticePhonePrunedWrapper
// (determinize-lattice-pruned.cc:1486-1506: Invert, TopSort, ArcSort, [phone-level pass,] word-level pass, Connect)
// and comes back as flat compact-lattice arrays.  Oracle of kaldi_b200/csrc/lattice_det.cu (tests/test_lattice_det.py).
#include <cstdint>
#include <chrono>
#include <cstring>
#include <sstream>
#include <string>
#include <vector>

#include "lat/determinize-lattice-pruned.h"

namespace {

// A transition model for synthetic graphs: transition-id t belongs to "phone" phone_of[t]; self-loop flags as given.
class ArrayTransitionInformation : public kaldi::TransitionInformation {
 public:
  ArrayTransitionInformation(const int32_t *phone_of, const uint8_t *self_loop, const uint8_t *phone_start, int32_t n)
      : phone_(phone_of, phone_of + n), loop_(self_loop, self_loop + n), start_(phone_start, phone_start + n), pdf_(n, 0) {}
  bool TransitionIdsEquivalent(int32_t a, int32_t b) const override { return phone_[a] == phone_[b]; }
  bool TransitionIdIsStartOfPhone(int32_t t) const override { return start_[t] != 0; }
  int32_t TransitionIdToPhone(int32_t t) const override { return phone_[t]; }
  bool IsFinal(int32_t) const override { return false; }
  bool IsSelfLoop(int32_t t) const override { return loop_[t] != 0; }
  const std::vector<int32_t> &TransitionIdToPdfArray() const override { return pdf_; }
  int32_t NumPdfs() const override { return 1; }
 private:
  std::vector<int32_t> phone_;
  std::vector<uint8_t> loop_, start_;
  std::vector<int32_t> pdf_;
};

double g_last_ms = 0.0;   // wall time of the last DeterminizeLatticePhonePrunedWrapper call alone

struct Out {
  std::vector<int32_t> arc_src, arc_dst, arc_word, final_state, tids;
  std::vector<float> arc_g, arc_a, final_g, final_a;
  std::vector<int64_t> arc_off, final_off;
  int64_t num_states = 0;
  int ok = 1;
};

}  // namespace

extern "C" {

void *ref_det_run(int32_t num_states, int64_t num_arcs, const int32_t *src, const int32_t *dst, const int32_t *ilabel,
                  const int32_t *olabel, const float *graph, const float *acoustic, int64_t num_finals,
                  const int32_t *final_state, const float *final_cost, double beam, int32_t phone_determinize,
                  const int32_t *phone_of, const uint8_t *self_loop, const uint8_t *phone_start, int32_t num_tids,
                  int32_t max_mem, int32_t minimize) {
  using namespace kaldi;
  Lattice lat;
  for (int32_t s = 0; s < num_states; s++) lat.AddState();
  if (num_states > 0) lat.SetStart(0);
  for (int64_t a = 0; a < num_arcs; a++)
    lat.AddArc(src[a], LatticeArc(ilabel[a], olabel[a], LatticeWeight(graph[a], acoustic[a]), dst[a]));   // GetRawLattice form
  for (int64_t f = 0; f < num_finals; f++) lat.SetFinal(final_state[f], LatticeWeight(final_cost[f], 0.0f));
  std::vector<int32_t> zero(1, 0);
  std::vector<uint8_t> zb(1, 0);
  ArrayTransitionInformation tm(phone_of ? phone_of : zero.data(), self_loop ? self_loop : zb.data(),
                                phone_start ? phone_start : zb.data(), phone_of ? num_tids : 1);
  fst::DeterminizeLatticePhonePrunedOptions opts;           // defaults: delta kDelta, word_determinize, no minimize
  opts.phone_determinize = phone_determinize != 0;
  if (max_mem > 0) opts.max_mem = max_mem;
  opts.minimize = minimize != 0;                            // push strings, push weights, minimize (:1459-1465; the reference's own
                                                            // lat/push-lattice.cc and lat/minimize-lattice.cc are compiled in)
  CompactLattice clat;
  Out *o = new Out();
  const auto t0 = std::chrono::steady_clock::now();
  o->ok = fst::DeterminizeLatticePhonePrunedWrapper(tm, &lat, beam, &clat, opts) ? 1 : 0;
  g_last_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
  o->num_states = clat.NumStates();
  o->arc_off.push_back(0);
  // the readers assume state 0 = start: TopSort put it first; after minimization it may have been merged into a later state,
  // in which case the two numbers are exchanged on the way out
  const int32_t start = clat.NumStates() > 0 ? clat.Start() : 0;
  auto id = [start](int32_t s) { return s == start ? 0 : (s == 0 ? start : s); };
  for (int32_t n = 0; n < clat.NumStates(); n++) {
    const int32_t s = id(n);                                // the state that is written as number n
    for (fst::ArcIterator<CompactLattice> it(clat, s); !it.Done(); it.Next()) {
      const CompactLatticeArc &a = it.Value();
      o->arc_src.push_back(n); o->arc_dst.push_back(id(a.nextstate)); o->arc_word.push_back(a.ilabel);
      o->arc_g.push_back(a.weight.Weight().Value1()); o->arc_a.push_back(a.weight.Weight().Value2());
      o->tids.insert(o->tids.end(), a.weight.String().begin(), a.weight.String().end());
      o->arc_off.push_back((int64_t)o->tids.size());
    }
  }
  o->final_off.push_back((int64_t)o->tids.size());
  for (int32_t n = 0; n < clat.NumStates(); n++) {
    const CompactLatticeWeight w = clat.Final(id(n));
    if (w == CompactLatticeWeight::Zero()) continue;
    o->final_state.push_back(n); o->final_g.push_back(w.Weight().Value1()); o->final_a.push_back(w.Weight().Value2());
    o->tids.insert(o->tids.end(), w.String().begin(), w.String().end());
    o->final_off.push_back((int64_t)o->tids.size());
  }
  return o;
}

void ref_det_sizes(void *h, int64_t sz[5]) {
  Out *o = (Out *)h;
  sz[0] = o->num_states; sz[1] = (int64_t)o->arc_src.size(); sz[2] = (int64_t)o->final_state.size(); sz[3] = (int64_t)o->tids.size(); sz[4] = o->ok;
}

void ref_det_copy(void *h, int32_t *arc_src, int32_t *arc_dst, int32_t *arc_word, float *arc_g, float *arc_a, int64_t *arc_off,
                  int32_t *final_state, float *final_g, float *final_a, int64_t *final_off, int32_t *tids) {
  Out *o = (Out *)h;
  auto cp = [](void *d, const void *s, size_t n) { if (n) memcpy(d, s, n); };
  const size_t na = o->arc_src.size(), nf = o->final_state.size();
  cp(arc_src, o->arc_src.data(), 4 * na); cp(arc_dst, o->arc_dst.data(), 4 * na); cp(arc_word, o->arc_word.data(), 4 * na);
  cp(arc_g, o->arc_g.data(), 4 * na); cp(arc_a, o->arc_a.data(), 4 * na); cp(arc_off, o->arc_off.data(), 8 * (na + 1));
  cp(final_state, o->final_state.data(), 4 * nf); cp(final_g, o->final_g.data(), 4 * nf); cp(final_a, o->final_a.data(), 4 * nf);
  cp(final_off, o->final_off.data(), 8 * (nf + 1)); cp(tids, o->tids.data(), 4 * o->tids.size());
}

void ref_det_free(void *h) { delete (Out *)h; }
double ref_det_last_ms() { return g_last_ms; }

// The reference's own binary encodings: CompactLatticeWeightTpl::Write (fstext/lattice-weight.h:531-540) when n >= 0,
// LatticeWeightTpl::Write (:141-146) when n < 0; and the arc type strings of the two lattice types.  Returns the byte count.
int32_t ref_lattice_weight_bytes(float g, float a, const int32_t *tids, int32_t n, uint8_t *out, int32_t cap) {
  std::ostringstream os(std::ios::binary);
  if (n >= 0) kaldi::CompactLatticeWeight(kaldi::LatticeWeight(g, a), std::vector<int32_t>(tids, tids + n)).Write(os);
  else kaldi::LatticeWeight(g, a).Write(os);
  const std::string s = os.str();
  if ((int32_t)s.size() <= cap) memcpy(out, s.data(), s.size());
  return (int32_t)s.size();
}
// The text form of the two weight types: the reference's own operator<< (fstext/lattice-weight.h:396-403, 726-741), i.e. what
// FstPrinter puts into the weight column of a text-mode lattice.  n < 0: LatticeWeight.
int32_t ref_lattice_weight_text(float g, float a, const int32_t *tids, int32_t n, char *out, int32_t cap) {
  std::ostringstream os;
  if (n >= 0) os << kaldi::CompactLatticeWeight(kaldi::LatticeWeight(g, a), std::vector<int32_t>(tids, tids + n));
  else os << kaldi::LatticeWeight(g, a);
  const std::string s = os.str();
  if ((int32_t)s.size() + 1 <= cap) memcpy(out, s.c_str(), s.size() + 1);
  return (int32_t)s.size();
}
int32_t ref_lattice_type_strings(char *out, int32_t cap) {
  const std::string s = kaldi::LatticeWeight::Type() + " " + kaldi::CompactLatticeWeight::Type();
  if ((int32_t)s.size() + 1 <= cap) memcpy(out, s.c_str(), s.size() + 1);
  return (int32_t)s.size();
}

}  // extern "C"
