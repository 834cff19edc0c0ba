// OnlineSilenceWeightingB2k (kaldi_b200/host/b2k_online2_shims.h), the Kaldi-typed wrapper, beside the reference's OWN
// OnlineSilenceWeighting (online-ivector-feature.cc as compiled into oracle/_ref/libkaldi_ref_nnet3.so over the replay decoder of
// oracle/ref_wrap/replay_decoder.h), both constructed from the same TransitionModel and OnlineSilenceWeightingConfig and driven
// by the same decoder: a random trellis of tokens whose best path after every chunk goes to the reference through BestPathEnd /
// TraceBackBestPath and to the wrapper through the BestPath(...) arrays of SingleUtteranceNnet3DecoderB2k.  Test harness only.
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>

#include "replay_decoder.h"                        // declares the decoder templates online-ivector-feature.h names
#include "b2k_online2_shims.h"
#include "hmm/transition-model.h"
#include "util/kaldi-io.h"

using namespace kaldi;

#define REQUIRE(c) do { if (!(c)) { std::fprintf(stderr, "FAILED %s:%d: %s\n", __FILE__, __LINE__, #c); std::exit(1); } } while (0)

namespace {

struct Token { int state, pred_t, pred_i, label; };          // pred_t < 0: the start token

struct Trellis {
  std::mt19937 &rng;
  int num_tids;
  std::vector<std::vector<Token> > tokens;
  Trellis(std::mt19937 &r, int n) : rng(r), num_tids(n) { tokens.push_back(std::vector<Token>(1, Token{0, -1, -1, 0})); }
  int TidInto(int t, int i) const {
    while (true) {
      const Token &k = tokens[t][i];
      if (k.label != 0 || k.pred_t < 0) return k.label;
      t = k.pred_t; i = k.pred_i;
    }
  }
  void Grow(int n) {
    for (int f = 0; f < n; f++) {
      const int t = (int)tokens.size();
      const std::vector<Token> &prev = tokens.back();
      std::vector<Token> cur;
      for (int lane = 0; lane < 3; lane++) {
        const int src = (rng() % 100 < 93 && lane < (int)prev.size()) ? lane : (int)(rng() % prev.size());
        const int last = TidInto(t - 1, src);
        const int tid = (last > 0 && rng() % 100 < 80) ? last : 1 + (int)(rng() % num_tids);
        cur.push_back(Token{1 + lane + 10 * (int)(rng() % 4), t - 1, src, tid});
      }
      if (rng() % 100 < 15) cur.push_back(Token{5 + 10 * (int)(rng() % 4), t, (int)(rng() % cur.size()), 0});   // an epsilon token
      tokens.push_back(cur);
    }
  }
  // arcs start first: ilabels, the state each arc enters, the state each arc leaves (a large id for the start token)
  void BestPath(std::vector<int> *il, std::vector<int> *dst, std::vector<int> *src) {
    il->clear(); dst->clear(); src->clear();
    int t = (int)tokens.size() - 1, i = (int)(rng() % tokens[t].size());
    while (tokens[t][i].pred_t >= 0) {
      const Token k = tokens[t][i];
      il->push_back(k.label); dst->push_back(k.state);
      t = k.pred_t; i = k.pred_i;
      src->push_back(tokens[t][i].pred_t < 0 ? (1 << 20) : tokens[t][i].state);
    }
    std::reverse(il->begin(), il->end()); std::reverse(dst->begin(), dst->end()); std::reverse(src->begin(), src->end());
  }
};

// what OnlineSilenceWeightingB2k::ComputeCurrentTraceback asks of a decoder (SingleUtteranceNnet3DecoderB2k's members)
struct ArrayDecoder {
  std::vector<int32> ilabels, arc_states;
  int32 frames = 0;
  int32 NumFramesDecoded() const { return frames; }
  void BestPath(bool, std::vector<int32> *il, std::vector<int32> *, std::vector<BaseFloat> *, std::vector<BaseFloat> *,
                b2k_best_path_info *info, std::vector<int32> *states) {
    *il = ilabels; *states = arc_states;
    info->n_arcs = (int32)ilabels.size(); info->num_frames = frames; info->status = 0;
  }
};

}  // namespace

int main(int argc, char **argv) {
  REQUIRE(argc == 2);
  TransitionModel tm;
  { bool binary; Input ki(argv[1], &binary); tm.Read(ki.Stream(), binary); }
  std::mt19937 rng(11);
  typedef LatticeFasterOnlineDecoderTpl<fst::Fst<fst::StdArc> > RefDecoder;
  int compared = 0, taken_back = 0;
  for (int trial = 0; trial < 12; trial++) {
    OnlineSilenceWeightingConfig cfg;
    cfg.silence_phones_str = trial % 3 == 2 ? "1" : "1:2";
    cfg.silence_weight = trial % 2 ? 0.001f : 0.0f;
    cfg.max_state_duration = trial % 4 == 3 ? 5.0f : -1.0f;
    const int32 fs = trial % 2 ? 3 : 1, first = trial % 3 == 1 ? 2 * fs : 0;
    OnlineSilenceWeighting ref(tm, cfg, fs);
    b2k_shim::OnlineSilenceWeightingB2k mine(tm, cfg, fs);
    REQUIRE(ref.Active() && mine.Active());
    Trellis tr(rng, tm.NumTransitionIds());
    int32 frames = 0, ready = 0;
    std::vector<int> il, dst, src;
    while (frames < 300) {
      const int n = 1 + (int)(rng() % 25);
      tr.Grow(n);
      frames += n;
      tr.BestPath(&il, &dst, &src);
      RefDecoder rd;
      rd.path = il; rd.source = src; rd.frames = frames;
      ref.ComputeCurrentTraceback(rd, false);
      ArrayDecoder ad;
      ad.ilabels.assign(il.begin(), il.end()); ad.arc_states.assign(dst.begin(), dst.end()); ad.frames = frames;
      mine.ComputeCurrentTraceback(ad);
      ready = std::max(ready, first + fs * frames + (int)(rng() % (4 * fs)));
      std::vector<std::pair<int32, BaseFloat> > a, b;
      ref.GetDeltaWeights(ready, first, &a);
      mine.GetDeltaWeights(ready, first, &b);
      REQUIRE(a == b && !a.empty());
      for (size_t k = 0; k < a.size(); k++) taken_back += a[k].second < 0;
      std::vector<int32> na, nb;
      ref.GetNonsilenceFrames(ready, first, &na);
      mine.GetNonsilenceFrames(ready, first, &nb);
      REQUIRE(na == nb);
      compared++;
    }
  }
  REQUIRE(compared > 100 && taken_back > 0);
  // the error the reference raises when a decoder goes backwards is the wrapper's too (KaldiFatalError)
  {
    OnlineSilenceWeightingConfig cfg;
    cfg.silence_phones_str = "1"; cfg.silence_weight = 0.5f;
    b2k_shim::OnlineSilenceWeightingB2k mine(tm, cfg, 1);
    ArrayDecoder ad;
    ad.ilabels.assign(10, 1); ad.arc_states.assign(10, 3); ad.frames = 10;
    mine.ComputeCurrentTraceback(ad);
    ad.ilabels.resize(4); ad.arc_states.resize(4); ad.frames = 4;
    bool threw = false;
    try { mine.ComputeCurrentTraceback(ad); } catch (const KaldiFatalError &) { threw = true; }
    REQUIRE(threw);
  }
  std::printf("silence weighting wrapper ok (%d calls compared, %d weights taken back)\n", compared, taken_back);
  return 0;
}
