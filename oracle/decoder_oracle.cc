// oracle/decoder_oracle.cc
//
// TEST INFRASTRUCTURE ONLY.  CPU restatement of Kaldi's lattice-generating
// token-passing decoder over a plain CSR graph.  Only tests/,
// __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
// may load this library; the product (kaldi_b200/) never does.
//
// PARITY STATUS: pinned to the reference itself.  The reference ships no unit
// test and no golden lattice for src/decoder, and OpenFst is
// absent from this image, but the decoder's search code needs only FST
// containers: oracle/ref_decoder.py compiles decoder/lattice-faster-decoder.cc
// and util/hash-list-inl.h from where they lie against a container-only stand-in
// (oracle/ref_wrap/fst_stub/), and tests/test_decoder_oracle.py checks this
// restatement against that library bit for bit (token set of every frame, finalized
// raw lattice) on seeded graphs incl. a 333-frame BASELINE-size utterance.  The
// restatement stays because it adds what the reference cannot give: mode 1, link
// sets per frame, cutoffs/cost offsets per frame, visit counters.  Line map:
//
//   decoder/lattice-faster-decoder.cc   (LatticeFasterDecoderTpl<ConstFst,...>)
//     InitDecoding :63-81            -> Oracle::InitDecoding
//     FindOrAddToken :261-302        -> Oracle::FindOrAddToken
//     PruneForwardLinks :308-379     -> Oracle::PruneForwardLinks
//     PruneForwardLinksFinal :385-467-> Oracle::PruneForwardLinksFinal
//     PruneTokensForFrame :488-507   -> Oracle::PruneTokensForFrame
//     PruneActiveTokens :515-542     -> Oracle::PruneActiveTokens
//     ComputeFinalCosts :545-586     -> Oracle::ComputeFinalCosts
//     AdvanceDecoding :589-628       -> Oracle::Decode (loop body)
//     FinalizeDecoding :634-649      -> Oracle::FinalizeDecoding
//     GetCutoff :653-720             -> Oracle::GetCutoff
//     ProcessEmitting :723-814       -> Oracle::ProcessEmitting
//     ProcessNonemitting :830-897    -> Oracle::ProcessNonemitting
//     GetRawLattice :114-197         -> Oracle::ExportLattice (canonical set,
//                                       no OpenFst state numbering)
//   util/hash-list-inl.h :30-175      -> HashList (bucket/list order identical)
//   base/kaldi-math.h :265-273        -> ApproxEqual
//
// Two modes:
//   mode 0 "reference order": literal restatement, including the running
//          next_cutoff that tightens in HashList iteration order (:794-796).
//   mode 1 "order free": the data-parallel semantics the CUDA decoder
//          implements: a (token, arc) pair is admitted iff tot_cost < FINAL
//          next_cutoff of the frame (the value mode 0 returns), and ties for
//          the best token are broken by smallest state id.  Everything else
//          is the same code.  Mode 0 additionally creates order-dependent
//          "extra" links/tokens with tot_cost in [final cutoff, running
//          cutoff); the tests measure whether any of them survives into the
//          finalized lattice.
//
// All arithmetic is float32 evaluated left-to-right exactly as the reference
// writes it.  Build: g++ -O2 -ffp-contract=off (no FMA contraction).

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <vector>

namespace {

typedef int32_t int32;
typedef float BaseFloat;
static const BaseFloat kInf = std::numeric_limits<BaseFloat>::infinity();

struct Cfg {          // LatticeFasterDecoderConfig, lattice-faster-decoder.h:38-106
  float beam;
  int32 max_active;
  int32 min_active;
  float lattice_beam;
  int32 prune_interval;
  float beam_delta;
  float hash_ratio;
  float prune_scale;
};

struct Graph {        // fst::ConstFst<StdArc> semantics: arcs in file order
  int32 num_states, start;
  std::vector<int32> offsets;               // [num_states+1]
  std::vector<int32> ilabel, olabel, nextstate;
  std::vector<float> weight;
  std::vector<float> final_cost;            // +inf = not final
  std::vector<int32> num_ieps;              // NumInputEpsilons(state)
  std::vector<int32> tid2pdf;               // TransitionIdToPdfFast
};

struct Token;
struct ForwardLink {
  Token *next_tok;
  int32 ilabel, olabel;
  BaseFloat graph_cost, acoustic_cost;
  ForwardLink *next;
};
struct Token {
  BaseFloat tot_cost, extra_cost;
  ForwardLink *links;
  Token *next;
  int32 state;        // not in the reference struct; used only for export
};

template <class T> struct Pool {
  std::vector<T *> blocks;
  T *free_head = nullptr;
  size_t block = 4096;
  T *Allocate() {
    if (!free_head) {
      T *b = (T *)malloc(sizeof(T) * block);
      blocks.push_back(b);
      for (size_t i = 0; i < block; i++) {
        *(T **)(&b[i]) = (i + 1 < block) ? &b[i + 1] : nullptr;
      }
      free_head = b;
    }
    T *ans = free_head;
    free_head = *(T **)ans;
    return ans;
  }
  void Free(T *p) { *(T **)p = free_head; free_head = p; }
  ~Pool() { for (T *b : blocks) free(b); }
};

// util/hash-list-inl.h
struct Elem { int32 key; Token *val; Elem *tail; };
struct HashList {
  struct HashBucket { size_t prev_bucket; Elem *last_elem; };
  Elem *list_head_ = nullptr;
  size_t bucket_list_tail_ = (size_t)-1;
  size_t hash_size_ = 0;
  std::vector<HashBucket> buckets_;
  Pool<Elem> pool_;
  void SetSize(size_t size) {                       // :38-44
    hash_size_ = size;
    if (size > buckets_.size()) buckets_.resize(size, HashBucket{0, nullptr});
  }
  size_t Size() const { return hash_size_; }
  Elem *Clear() {                                   // :46-59
    for (size_t b = bucket_list_tail_; b != (size_t)-1; b = buckets_[b].prev_bucket)
      buckets_[b].last_elem = nullptr;
    bucket_list_tail_ = (size_t)-1;
    Elem *ans = list_head_;
    list_head_ = nullptr;
    return ans;
  }
  const Elem *GetList() const { return list_head_; }
  void Delete(Elem *e) { pool_.Free(e); }
  Elem *Insert(int32 key, Token *val) {             // :126-175
    size_t index = (size_t)key % hash_size_;
    HashBucket &bucket = buckets_[index];
    if (bucket.last_elem != nullptr) {
      Elem *head = (bucket.prev_bucket == (size_t)-1
                        ? list_head_
                        : buckets_[bucket.prev_bucket].last_elem->tail),
           *tail = bucket.last_elem->tail;
      for (Elem *e = head; e != tail; e = e->tail)
        if (e->key == key) return e;
    }
    Elem *elem = pool_.Allocate();
    elem->key = key;
    elem->val = val;
    if (bucket.last_elem == nullptr) {
      if (bucket_list_tail_ == (size_t)-1) {
        list_head_ = elem;
      } else {
        buckets_[bucket_list_tail_].last_elem->tail = elem;
      }
      elem->tail = nullptr;
      bucket.last_elem = elem;
      bucket.prev_bucket = bucket_list_tail_;
      bucket_list_tail_ = index;
    } else {
      elem->tail = bucket.last_elem->tail;
      bucket.last_elem->tail = elem;
      bucket.last_elem = elem;
    }
    return elem;
  }
};

static inline bool ApproxEqual(float a, float b, float tol) {  // kaldi-math.h:265
  if (a == b) return true;
  float diff = std::abs(a - b);
  if (diff == kInf || diff != diff) return false;
  return (diff <= tol * (std::abs(a) + std::abs(b)));
}

struct TokenList {
  Token *toks = nullptr;
  bool must_prune_forward_links = true;
  bool must_prune_tokens = true;
};

// one recorded raw frame (before any pruning touches it)
struct RawFrame {
  std::vector<int32> tok_state;
  std::vector<float> tok_cost;
  // links created by this frame step: emitting (prev list -> this list) and
  // epsilon (this list -> this list).  7 int32-sized fields per link:
  // src_state, dst_state, ilabel, olabel, graph(bits), acoustic(bits), is_eps
  std::vector<int32> links;
};

struct Oracle {
  Graph g;
  Cfg cfg;
  int mode = 0;
  bool record = false;

  HashList toks_;
  std::vector<TokenList> active_toks_;
  std::vector<const Elem *> queue_;
  std::vector<BaseFloat> tmp_array_;
  std::vector<BaseFloat> cost_offsets_;
  Pool<Token> token_pool_;
  Pool<ForwardLink> link_pool_;
  int32 num_toks_ = 0;
  bool decoding_finalized_ = false;
  // final costs: per token of the last frame (map Token* -> cost); we key by
  // state since one token per (frame,state)
  std::vector<std::pair<Token *, BaseFloat>> final_costs_;
  BaseFloat final_relative_cost_ = 0, final_best_cost_ = 0;

  const float *loglikes_ = nullptr;
  int32 num_pdfs_ = 0;

  // diagnostics
  int64_t arcs_emitting_ = 0, arcs_nonemitting_ = 0, tokens_expanded_ = 0;
  int64_t n_extra_links_ = 0, n_best_ties_ = 0, n_min_active_branch_ = 0,
          n_max_active_branch_ = 0, n_extras_expanded_ = 0, n_links_admitted_ = 0,
          n_toks_created_ = 0;
  std::vector<float> frame_cutoff_;     // next_cutoff returned per frame
  std::vector<int32> frame_ntoks_;      // tokens in list after each frame step
  std::vector<RawFrame> raw_;

  // exported lattice
  std::vector<int32> lat_state_frame, lat_state_state;
  std::vector<float> lat_state_tot, lat_state_extra;
  std::vector<int32> lat_arc;          // 8 fields: src_frame, src_state, dst_frame, dst_state, ilabel, olabel, graph bits, acoustic bits
  std::vector<int32> lat_final_state;
  std::vector<float> lat_final_cost;

  inline BaseFloat LogLikelihood(int32 frame, int32 tid) const {
    return loglikes_[(size_t)frame * num_pdfs_ + g.tid2pdf[tid]];
  }
  int32 NumFramesDecoded() const { return (int32)active_toks_.size() - 1; }

  void DeleteForwardLinks(Token *tok) {
    ForwardLink *l = tok->links, *m;
    while (l != nullptr) { m = l->next; link_pool_.Free(l); l = m; }
    tok->links = nullptr;
  }
  void DeleteElems(Elem *list) {
    for (Elem *e = list, *e_tail; e != nullptr; e = e_tail) { e_tail = e->tail; toks_.Delete(e); }
  }
  void ClearActiveTokens() {
    for (size_t i = 0; i < active_toks_.size(); i++) {
      for (Token *tok = active_toks_[i].toks; tok != nullptr;) {
        DeleteForwardLinks(tok);
        Token *next_tok = tok->next;
        token_pool_.Free(tok);
        num_toks_--;
        tok = next_tok;
      }
    }
    active_toks_.clear();
  }

  void PossiblyResizeHash(size_t num_toks) {                      // :227-233
    size_t new_sz = (size_t)((BaseFloat)num_toks * cfg.hash_ratio);
    if (new_sz > toks_.Size()) toks_.SetSize(new_sz);
  }

  Elem *FindOrAddToken(int32 state, int32 frame_plus_one, BaseFloat tot_cost,
                       bool *changed) {                           // :261-302
    Token *&toks = active_toks_[frame_plus_one].toks;
    Elem *e_found = toks_.Insert(state, nullptr);
    if (e_found->val == nullptr) {
      Token *new_tok = token_pool_.Allocate();
      new_tok->tot_cost = tot_cost;
      new_tok->extra_cost = 0.0f;
      new_tok->links = nullptr;
      new_tok->next = toks;
      new_tok->state = state;
      toks = new_tok;
      num_toks_++;
      n_toks_created_++;
      e_found->val = new_tok;
      if (changed) *changed = true;
      return e_found;
    } else {
      Token *tok = e_found->val;
      if (tok->tot_cost > tot_cost) {
        tok->tot_cost = tot_cost;
        if (changed) *changed = true;
      } else {
        if (changed) *changed = false;
      }
      return e_found;
    }
  }

  void InitDecoding() {                                           // :63-81
    DeleteElems(toks_.Clear());
    cost_offsets_.clear();
    ClearActiveTokens();
    num_toks_ = 0;
    decoding_finalized_ = false;
    final_costs_.clear();
    active_toks_.resize(1);
    Token *start_tok = token_pool_.Allocate();
    start_tok->tot_cost = 0.0f; start_tok->extra_cost = 0.0f;
    start_tok->links = nullptr; start_tok->next = nullptr; start_tok->state = g.start;
    active_toks_[0].toks = start_tok;
    toks_.Insert(g.start, start_tok);
    num_toks_++;
    ProcessNonemitting(cfg.beam);
    if (record) RecordFrame(0);
  }

  void PruneForwardLinks(int32 frame_plus_one, bool *extra_costs_changed,
                         bool *links_pruned, BaseFloat delta) {   // :308-379
    *extra_costs_changed = false;
    *links_pruned = false;
    bool changed = true;
    while (changed) {
      changed = false;
      for (Token *tok = active_toks_[frame_plus_one].toks; tok != nullptr; tok = tok->next) {
        ForwardLink *link, *prev_link = nullptr;
        BaseFloat tok_extra_cost = kInf;
        for (link = tok->links; link != nullptr;) {
          Token *next_tok = link->next_tok;
          BaseFloat link_extra_cost = next_tok->extra_cost +
              ((tok->tot_cost + link->acoustic_cost + link->graph_cost) - next_tok->tot_cost);
          if (link_extra_cost > cfg.lattice_beam) {
            ForwardLink *next_link = link->next;
            if (prev_link != nullptr) prev_link->next = next_link;
            else tok->links = next_link;
            link_pool_.Free(link);
            link = next_link;
            *links_pruned = true;
          } else {
            if (link_extra_cost < 0.0f) link_extra_cost = 0.0f;
            if (link_extra_cost < tok_extra_cost) tok_extra_cost = link_extra_cost;
            prev_link = link;
            link = link->next;
          }
        }
        if (std::fabs(tok_extra_cost - tok->extra_cost) > delta) changed = true;
        tok->extra_cost = tok_extra_cost;
      }
      if (changed) *extra_costs_changed = true;
    }
  }

  void ComputeFinalCosts(std::vector<std::pair<Token *, BaseFloat>> *final_costs,
                         BaseFloat *final_relative_cost, BaseFloat *final_best_cost) { // :545-586
    if (final_costs) final_costs->clear();
    const Elem *final_toks = toks_.GetList();
    BaseFloat best_cost = kInf, best_cost_with_final = kInf;
    while (final_toks != nullptr) {
      int32 state = final_toks->key;
      Token *tok = final_toks->val;
      const Elem *next = final_toks->tail;
      BaseFloat final_cost = g.final_cost[state];
      BaseFloat cost = tok->tot_cost, cost_with_final = cost + final_cost;
      best_cost = std::min(cost, best_cost);
      best_cost_with_final = std::min(cost_with_final, best_cost_with_final);
      if (final_costs && final_cost != kInf) final_costs->push_back({tok, final_cost});
      final_toks = next;
    }
    if (final_relative_cost) {
      if (best_cost == kInf && best_cost_with_final == kInf) *final_relative_cost = kInf;
      else *final_relative_cost = best_cost_with_final - best_cost;
    }
    if (final_best_cost) {
      if (best_cost_with_final != kInf) *final_best_cost = best_cost_with_final;
      else *final_best_cost = best_cost;
    }
  }

  // final_costs_ is sorted by pointer in PruneForwardLinksFinal
  BaseFloat FinalCostOf(Token *tok, bool *found) const {
    auto it = std::lower_bound(final_costs_.begin(), final_costs_.end(), tok,
                               [](const std::pair<Token *, BaseFloat> &a, Token *t) { return a.first < t; });
    if (it != final_costs_.end() && it->first == tok) { *found = true; return it->second; }
    *found = false; return kInf;
  }

  void PruneForwardLinksFinal() {                                 // :385-467
    int32 frame_plus_one = (int32)active_toks_.size() - 1;
    ComputeFinalCosts(&final_costs_, &final_relative_cost_, &final_best_cost_);
    decoding_finalized_ = true;
    DeleteElems(toks_.Clear());
    // index final costs by token for O(1) lookup: temporarily stash in a map
    // keyed by pointer order (sorted vector + binary search)
    std::sort(final_costs_.begin(), final_costs_.end(),
              [](const std::pair<Token *, BaseFloat> &a, const std::pair<Token *, BaseFloat> &b) {
                return a.first < b.first; });
    auto lookup = [&](Token *tok, BaseFloat *out) -> bool {
      auto it = std::lower_bound(final_costs_.begin(), final_costs_.end(), tok,
                                 [](const std::pair<Token *, BaseFloat> &a, Token *t) { return a.first < t; });
      if (it != final_costs_.end() && it->first == tok) { *out = it->second; return true; }
      return false;
    };
    bool changed = true;
    BaseFloat delta = 1.0e-05f;
    while (changed) {
      changed = false;
      for (Token *tok = active_toks_[frame_plus_one].toks; tok != nullptr; tok = tok->next) {
        ForwardLink *link, *prev_link = nullptr;
        BaseFloat final_cost;
        if (final_costs_.empty()) {
          final_cost = 0.0f;
        } else {
          if (!lookup(tok, &final_cost)) final_cost = kInf;
        }
        BaseFloat tok_extra_cost = tok->tot_cost + final_cost - final_best_cost_;
        for (link = tok->links; link != nullptr;) {
          Token *next_tok = link->next_tok;
          BaseFloat link_extra_cost = next_tok->extra_cost +
              ((tok->tot_cost + link->acoustic_cost + link->graph_cost) - next_tok->tot_cost);
          if (link_extra_cost > cfg.lattice_beam) {
            ForwardLink *next_link = link->next;
            if (prev_link != nullptr) prev_link->next = next_link;
            else tok->links = next_link;
            link_pool_.Free(link);
            link = next_link;
          } else {
            if (link_extra_cost < 0.0f) link_extra_cost = 0.0f;
            if (link_extra_cost < tok_extra_cost) tok_extra_cost = link_extra_cost;
            prev_link = link;
            link = link->next;
          }
        }
        if (tok_extra_cost > cfg.lattice_beam) tok_extra_cost = kInf;
        if (!ApproxEqual(tok->extra_cost, tok_extra_cost, delta)) changed = true;
        tok->extra_cost = tok_extra_cost;
      }
    }
  }

  void PruneTokensForFrame(int32 frame_plus_one) {                // :488-507
    Token *&toks = active_toks_[frame_plus_one].toks;
    Token *tok, *next_tok, *prev_tok = nullptr;
    for (tok = toks; tok != nullptr; tok = next_tok) {
      next_tok = tok->next;
      if (tok->extra_cost == kInf) {
        if (prev_tok != nullptr) prev_tok->next = tok->next;
        else toks = tok->next;
        token_pool_.Free(tok);
        num_toks_--;
      } else {
        prev_tok = tok;
      }
    }
  }

  void PruneActiveTokens(BaseFloat delta) {                       // :515-542
    int32 cur_frame_plus_one = NumFramesDecoded();
    for (int32 f = cur_frame_plus_one - 1; f >= 0; f--) {
      if (active_toks_[f].must_prune_forward_links) {
        bool extra_costs_changed = false, links_pruned = false;
        PruneForwardLinks(f, &extra_costs_changed, &links_pruned, delta);
        if (extra_costs_changed && f > 0) active_toks_[f - 1].must_prune_forward_links = true;
        if (links_pruned) active_toks_[f].must_prune_tokens = true;
        active_toks_[f].must_prune_forward_links = false;
      }
      if (f + 1 < cur_frame_plus_one && active_toks_[f + 1].must_prune_tokens) {
        PruneTokensForFrame(f + 1);
        active_toks_[f + 1].must_prune_tokens = false;
      }
    }
  }

  void FinalizeDecoding() {                                       // :634-649
    int32 final_frame_plus_one = NumFramesDecoded();
    PruneForwardLinksFinal();
    for (int32 f = final_frame_plus_one - 1; f >= 0; f--) {
      bool b1, b2;
      PruneForwardLinks(f, &b1, &b2, 0.0f);
      PruneTokensForFrame(f + 1);
    }
    PruneTokensForFrame(0);
  }

  BaseFloat GetCutoff(Elem *list_head, size_t *tok_count, BaseFloat *adaptive_beam,
                      Elem **best_elem) {                          // :653-720
    BaseFloat best_weight = kInf;
    size_t count = 0;
    int ties = 0;
    if (cfg.max_active == std::numeric_limits<int32>::max() && cfg.min_active == 0) {
      for (Elem *e = list_head; e != nullptr; e = e->tail, count++) {
        BaseFloat w = e->val->tot_cost;
        if (w < best_weight) { best_weight = w; *best_elem = e; ties = 0; }
        else if (w == best_weight) {
          ties++;
          if (mode == 1 && *best_elem && e->key < (*best_elem)->key) *best_elem = e;
        }
      }
      if (ties) n_best_ties_++;
      *tok_count = count;
      *adaptive_beam = cfg.beam;
      return best_weight + cfg.beam;
    } else {
      tmp_array_.clear();
      for (Elem *e = list_head; e != nullptr; e = e->tail, count++) {
        BaseFloat w = e->val->tot_cost;
        tmp_array_.push_back(w);
        if (w < best_weight) { best_weight = w; *best_elem = e; ties = 0; }
        else if (w == best_weight) {
          ties++;
          if (mode == 1 && *best_elem && e->key < (*best_elem)->key) *best_elem = e;
        }
      }
      if (ties) n_best_ties_++;
      *tok_count = count;
      BaseFloat beam_cutoff = best_weight + cfg.beam, min_active_cutoff = kInf,
                max_active_cutoff = kInf;
      if (tmp_array_.size() > (size_t)cfg.max_active) {
        std::nth_element(tmp_array_.begin(), tmp_array_.begin() + cfg.max_active, tmp_array_.end());
        max_active_cutoff = tmp_array_[cfg.max_active];
      }
      if (max_active_cutoff < beam_cutoff) {
        *adaptive_beam = max_active_cutoff - best_weight + cfg.beam_delta;
        n_max_active_branch_++;
        return max_active_cutoff;
      }
      if (tmp_array_.size() > (size_t)cfg.min_active) {
        if (cfg.min_active == 0) min_active_cutoff = best_weight;
        else {
          std::nth_element(tmp_array_.begin(), tmp_array_.begin() + cfg.min_active,
                           tmp_array_.size() > (size_t)cfg.max_active
                               ? tmp_array_.begin() + cfg.max_active
                               : tmp_array_.end());
          min_active_cutoff = tmp_array_[cfg.min_active];
        }
      }
      if (min_active_cutoff > beam_cutoff) {
        *adaptive_beam = min_active_cutoff - best_weight + cfg.beam_delta;
        n_min_active_branch_++;
        return min_active_cutoff;
      } else {
        *adaptive_beam = cfg.beam;
        return beam_cutoff;
      }
    }
  }

  BaseFloat ProcessEmitting() {                                   // :723-814
    int32 frame = (int32)active_toks_.size() - 1;
    active_toks_.resize(active_toks_.size() + 1);
    Elem *final_toks = toks_.Clear();
    Elem *best_elem = nullptr;
    BaseFloat adaptive_beam;
    size_t tok_cnt;
    BaseFloat cur_cutoff = GetCutoff(final_toks, &tok_cnt, &adaptive_beam, &best_elem);
    PossiblyResizeHash(tok_cnt);
    BaseFloat next_cutoff = kInf;
    BaseFloat cost_offset = 0.0f;
    if (best_elem) {
      int32 state = best_elem->key;
      Token *tok = best_elem->val;
      cost_offset = -tok->tot_cost;
      for (int32 a = g.offsets[state]; a < g.offsets[state + 1]; a++) {
        if (g.ilabel[a] != 0) {
          BaseFloat new_weight = g.weight[a] + cost_offset -
              LogLikelihood(frame, g.ilabel[a]) + tok->tot_cost;
          if (new_weight + adaptive_beam < next_cutoff) next_cutoff = new_weight + adaptive_beam;
        }
      }
    }
    cost_offsets_.resize(frame + 1, 0.0f);
    cost_offsets_[frame] = cost_offset;

    // The FINAL value of next_cutoff: min(seed, min over all (tok,arc) of
    // tot+adaptive_beam) -- see DESIGN.md "running cutoff is a prefix-min".
    BaseFloat final_cutoff = next_cutoff;
    if (mode == 1) {
      for (Elem *e = final_toks; e != nullptr; e = e->tail) {
        Token *tok = e->val;
        if (tok->tot_cost <= cur_cutoff) {
          int32 state = e->key;
          for (int32 a = g.offsets[state]; a < g.offsets[state + 1]; a++) {
            if (g.ilabel[a] != 0) {
              BaseFloat ac_cost = cost_offset - LogLikelihood(frame, g.ilabel[a]),
                        graph_cost = g.weight[a], cur_cost = tok->tot_cost,
                        tot_cost = cur_cost + ac_cost + graph_cost;
              if (tot_cost + adaptive_beam < final_cutoff) final_cutoff = tot_cost + adaptive_beam;
            }
          }
        }
      }
      next_cutoff = final_cutoff;  // fixed for the whole main pass
    }

    size_t first_link_admitted = n_links_admitted_;
    (void)first_link_admitted;
    std::vector<float> *admitted_tot = nullptr;
    std::vector<float> admitted_store;
    if (mode == 0) admitted_tot = &admitted_store;

    for (Elem *e = final_toks, *e_tail; e != nullptr; e = e_tail) {
      int32 state = e->key;
      Token *tok = e->val;
      if (tok->tot_cost <= cur_cutoff) {
        tokens_expanded_++;
        for (int32 a = g.offsets[state]; a < g.offsets[state + 1]; a++) {
          if (g.ilabel[a] != 0) {
            arcs_emitting_++;
            BaseFloat ac_cost = cost_offset - LogLikelihood(frame, g.ilabel[a]),
                      graph_cost = g.weight[a], cur_cost = tok->tot_cost,
                      tot_cost = cur_cost + ac_cost + graph_cost;
            if (tot_cost >= next_cutoff) continue;
            else if (mode == 0 && tot_cost + adaptive_beam < next_cutoff)
              next_cutoff = tot_cost + adaptive_beam;
            Elem *e_next = FindOrAddToken(g.nextstate[a], frame + 1, tot_cost, nullptr);
            ForwardLink *l = link_pool_.Allocate();
            l->next_tok = e_next->val; l->ilabel = g.ilabel[a]; l->olabel = g.olabel[a];
            l->graph_cost = graph_cost; l->acoustic_cost = ac_cost; l->next = tok->links;
            tok->links = l;
            n_links_admitted_++;
            if (admitted_tot) admitted_tot->push_back(tot_cost);
          }
        }
      }
      e_tail = e->tail;
      toks_.Delete(e);
    }
    if (admitted_tot) {
      for (float t : *admitted_tot) if (t >= next_cutoff) n_extra_links_++;
    }
    return next_cutoff;
  }

  void ProcessNonemitting(BaseFloat cutoff) {                     // :830-897
    int32 frame = (int32)active_toks_.size() - 2;
    for (const Elem *e = toks_.GetList(); e != nullptr; e = e->tail) {
      int32 state = e->key;
      if (g.num_ieps[state] != 0) queue_.push_back(e);
    }
    while (!queue_.empty()) {
      const Elem *e = queue_.back();
      queue_.pop_back();
      int32 state = e->key;
      Token *tok = e->val;
      BaseFloat cur_cost = tok->tot_cost;
      if (cur_cost >= cutoff) continue;
      DeleteForwardLinks(tok);
      tok->links = nullptr;
      for (int32 a = g.offsets[state]; a < g.offsets[state + 1]; a++) {
        if (g.ilabel[a] == 0) {
          arcs_nonemitting_++;
          BaseFloat graph_cost = g.weight[a], tot_cost = cur_cost + graph_cost;
          if (tot_cost < cutoff) {
            bool changed;
            Elem *e_new = FindOrAddToken(g.nextstate[a], frame + 1, tot_cost, &changed);
            ForwardLink *l = link_pool_.Allocate();
            l->next_tok = e_new->val; l->ilabel = 0; l->olabel = g.olabel[a];
            l->graph_cost = graph_cost; l->acoustic_cost = 0.0f; l->next = tok->links;
            tok->links = l;
            if (changed && g.num_ieps[g.nextstate[a]] != 0) queue_.push_back(e_new);
          }
        }
      }
    }
  }

  static inline int32 Bits(float f) { int32 i; memcpy(&i, &f, 4); return i; }

  // Snapshot the frame that was just produced: tokens of the last list, the
  // emitting links out of the previous list and epsilon links inside the
  // last list.  Nothing has pruned these yet (see AdvanceDecoding order).
  void RecordFrame(int32 frame_plus_one) {
    raw_.resize(frame_plus_one + 1);
    RawFrame &r = raw_[frame_plus_one];
    for (Token *tok = active_toks_[frame_plus_one].toks; tok; tok = tok->next) {
      r.tok_state.push_back(tok->state);
      r.tok_cost.push_back(tok->tot_cost);
      for (ForwardLink *l = tok->links; l; l = l->next) {
        int32 rec[7] = {tok->state, l->next_tok->state, l->ilabel, l->olabel,
                        Bits(l->graph_cost), Bits(l->acoustic_cost), 1};
        r.links.insert(r.links.end(), rec, rec + 7);
      }
    }
    if (frame_plus_one > 0) {
      for (Token *tok = active_toks_[frame_plus_one - 1].toks; tok; tok = tok->next) {
        for (ForwardLink *l = tok->links; l; l = l->next) {
          if (l->ilabel == 0) continue;
          int32 rec[7] = {tok->state, l->next_tok->state, l->ilabel, l->olabel,
                          Bits(l->graph_cost), Bits(l->acoustic_cost), 0};
          r.links.insert(r.links.end(), rec, rec + 7);
        }
      }
    }
  }

  void ExportLattice() {                                          // GetRawLattice :114-197
    lat_state_frame.clear(); lat_state_state.clear(); lat_state_tot.clear();
    lat_state_extra.clear(); lat_arc.clear(); lat_final_state.clear(); lat_final_cost.clear();
    int32 num_frames = (int32)active_toks_.size() - 1;
    for (int32 f = 0; f <= num_frames; f++) {
      for (Token *tok = active_toks_[f].toks; tok; tok = tok->next) {
        lat_state_frame.push_back(f);
        lat_state_state.push_back(tok->state);
        lat_state_tot.push_back(tok->tot_cost);
        lat_state_extra.push_back(tok->extra_cost);
        for (ForwardLink *l = tok->links; l; l = l->next) {
          BaseFloat cost_offset = 0.0f;
          if (l->ilabel != 0) cost_offset = cost_offsets_[f];
          BaseFloat ac = l->acoustic_cost - cost_offset;
          int32 rec[8] = {f, tok->state, f + (l->ilabel != 0 ? 1 : 0), l->next_tok->state,
                          l->ilabel, l->olabel, Bits(l->graph_cost), Bits(ac)};
          lat_arc.insert(lat_arc.end(), rec, rec + 8);
        }
        if (f == num_frames) {
          if (!final_costs_.empty()) {
            bool found; BaseFloat fc = FinalCostOf(tok, &found);
            if (found) { lat_final_state.push_back(tok->state); lat_final_cost.push_back(fc); }
          } else {
            lat_final_state.push_back(tok->state); lat_final_cost.push_back(0.0f);
          }
        }
      }
    }
  }

  int Decode(const float *loglikes, int32 T, int32 num_pdfs, int mode_, bool record_, bool finalize) {
    mode = mode_; record = record_;
    loglikes_ = loglikes; num_pdfs_ = num_pdfs;
    arcs_emitting_ = arcs_nonemitting_ = tokens_expanded_ = 0;
    n_extra_links_ = n_best_ties_ = n_min_active_branch_ = n_max_active_branch_ = 0;
    n_extras_expanded_ = n_links_admitted_ = n_toks_created_ = 0;
    frame_cutoff_.clear(); frame_ntoks_.clear(); raw_.clear();
    // the reference tools construct a fresh decoder per utterance
    // (online2-wav-nnet3-latgen-faster.cc:228): hash size restarts at 1000 (:39)
    DeleteElems(toks_.Clear());
    toks_.SetSize(1000);
    InitDecoding();
    while (NumFramesDecoded() < T) {                               // :621-627
      if (NumFramesDecoded() % cfg.prune_interval == 0)
        PruneActiveTokens(cfg.lattice_beam * cfg.prune_scale);
      BaseFloat cost_cutoff = ProcessEmitting();
      ProcessNonemitting(cost_cutoff);
      frame_cutoff_.push_back(cost_cutoff);
      int32 n = 0;
      for (const Elem *e = toks_.GetList(); e; e = e->tail) n++;
      frame_ntoks_.push_back(n);
      if (record) RecordFrame(NumFramesDecoded());
    }
    if (finalize) {
      FinalizeDecoding();
      ExportLattice();
    }
    return 0;
  }

  Oracle() { toks_.SetSize(1000); }                               // :39
};

}  // namespace

extern "C" {

struct b2k_oracle_dec_cfg {
  float beam; int32_t max_active; int32_t min_active; float lattice_beam;
  int32_t prune_interval; float beam_delta; float hash_ratio; float prune_scale;
};

void *b2k_oracle_dec_create(int32_t num_states, int32_t start, const int32_t *offsets,
                            const int32_t *ilabel, const int32_t *olabel,
                            const float *weight, const int32_t *nextstate,
                            const float *final_cost, const int32_t *tid2pdf,
                            int32_t num_tids, const b2k_oracle_dec_cfg *cfg) {
  Oracle *o = new Oracle();
  Graph &g = o->g;
  g.num_states = num_states; g.start = start;
  g.offsets.assign(offsets, offsets + num_states + 1);
  int32_t na = offsets[num_states];
  g.ilabel.assign(ilabel, ilabel + na);
  g.olabel.assign(olabel, olabel + na);
  g.weight.assign(weight, weight + na);
  g.nextstate.assign(nextstate, nextstate + na);
  g.final_cost.assign(final_cost, final_cost + num_states);
  g.tid2pdf.assign(tid2pdf, tid2pdf + num_tids);
  g.num_ieps.assign(num_states, 0);
  for (int32_t s = 0; s < num_states; s++) {
    int32_t c = 0;
    for (int32_t a = offsets[s]; a < offsets[s + 1]; a++) c += (ilabel[a] == 0);
    g.num_ieps[s] = c;
  }
  o->cfg.beam = cfg->beam; o->cfg.max_active = cfg->max_active; o->cfg.min_active = cfg->min_active;
  o->cfg.lattice_beam = cfg->lattice_beam; o->cfg.prune_interval = cfg->prune_interval;
  o->cfg.beam_delta = cfg->beam_delta; o->cfg.hash_ratio = cfg->hash_ratio;
  o->cfg.prune_scale = cfg->prune_scale;
  return o;
}

void b2k_oracle_dec_destroy(void *h) { delete (Oracle *)h; }

int b2k_oracle_dec_decode(void *h, const float *loglikes, int32_t T, int32_t num_pdfs,
                          int32_t mode, int32_t record_frames, int32_t finalize) {
  return ((Oracle *)h)->Decode(loglikes, T, num_pdfs, mode, record_frames != 0, finalize != 0);
}

// stats: [arcs_emitting, arcs_nonemitting, tokens_expanded, extra_links,
//         best_ties, min_active_branch, max_active_branch, links_admitted,
//         toks_created, num_lat_states, num_lat_arcs, num_lat_finals]
void b2k_oracle_dec_stats(void *h, int64_t *out) {
  Oracle *o = (Oracle *)h;
  out[0] = o->arcs_emitting_; out[1] = o->arcs_nonemitting_; out[2] = o->tokens_expanded_;
  out[3] = o->n_extra_links_; out[4] = o->n_best_ties_; out[5] = o->n_min_active_branch_;
  out[6] = o->n_max_active_branch_; out[7] = o->n_links_admitted_; out[8] = o->n_toks_created_;
  out[9] = (int64_t)o->lat_state_frame.size(); out[10] = (int64_t)o->lat_arc.size() / 8;
  out[11] = (int64_t)o->lat_final_state.size();
}

void b2k_oracle_dec_frame_info(void *h, float *cutoffs, int32_t *ntoks, float *cost_offsets) {
  Oracle *o = (Oracle *)h;
  for (size_t i = 0; i < o->frame_cutoff_.size(); i++) {
    if (cutoffs) cutoffs[i] = o->frame_cutoff_[i];
    if (ntoks) ntoks[i] = o->frame_ntoks_[i];
    if (cost_offsets) cost_offsets[i] = o->cost_offsets_[i];
  }
}

void b2k_oracle_dec_lattice(void *h, int32_t *state_frame, int32_t *state_state,
                            float *state_tot, float *state_extra, int32_t *arcs8,
                            int32_t *final_state, float *final_cost) {
  Oracle *o = (Oracle *)h;
  size_t ns = o->lat_state_frame.size();
  memcpy(state_frame, o->lat_state_frame.data(), ns * 4);
  memcpy(state_state, o->lat_state_state.data(), ns * 4);
  memcpy(state_tot, o->lat_state_tot.data(), ns * 4);
  memcpy(state_extra, o->lat_state_extra.data(), ns * 4);
  memcpy(arcs8, o->lat_arc.data(), o->lat_arc.size() * 4);
  memcpy(final_state, o->lat_final_state.data(), o->lat_final_state.size() * 4);
  memcpy(final_cost, o->lat_final_cost.data(), o->lat_final_cost.size() * 4);
}

// raw (pre-pruning) frame record access
void b2k_oracle_dec_raw_sizes(void *h, int32_t frame_plus_one, int64_t *ntok, int64_t *nlink) {
  Oracle *o = (Oracle *)h;
  *ntok = (int64_t)o->raw_[frame_plus_one].tok_state.size();
  *nlink = (int64_t)o->raw_[frame_plus_one].links.size() / 7;
}
void b2k_oracle_dec_raw_copy(void *h, int32_t frame_plus_one, int32_t *tok_state, float *tok_cost,
                             int32_t *links7) {
  Oracle *o = (Oracle *)h;
  RawFrame &r = o->raw_[frame_plus_one];
  memcpy(tok_state, r.tok_state.data(), r.tok_state.size() * 4);
  memcpy(tok_cost, r.tok_cost.data(), r.tok_cost.size() * 4);
  memcpy(links7, r.links.data(), r.links.size() * 4);
}

}  // extern "C"
