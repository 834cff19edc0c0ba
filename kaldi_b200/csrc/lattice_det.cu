// lattice_det.cu — host-only: finalized raw lattice -> word-deterministic compact lattice.
//
// Stands where SingleUtteranceNnet3DecoderTpl::GetLattice calls DeterminizeLatticePhonePrunedWrapper
// (online2/online-nnet3-decoding.cc:60-78, lat/determinize-lattice-pruned.h:284) and where
// CompactLatticeShortestPath reads the 1-best off the result (online2bin/online2-wav-nnet3-latgen-faster.cc:43-54).
// Semantics taken from the reference, algorithm our own:
//  * weights are LatticeWeight (graph, acoustic) ordered by their sum, ties by the graph part
//    (fstext/lattice-weight.h:295-308); a compact-lattice weight adds the transition-id string, ties broken by
//    the shorter then the lexicographically smaller string (:590-604);
//  * the result accepts word sequences; for each sequence it carries the best path's weight and transition-id
//    string; word epsilons (olabel 0) are absorbed; every state has at most one arc per word;
//  * pruning: every sequence whose best cost <= best cost of the lattice + beam is kept, with that cost and its string
//    (determinize-lattice-pruned.h:126-140: "--beam" relative to the best path).  A sequence outside the beam can survive
//    where its states are shared with sequences inside it; it then carries the weight of its best SURVIVING derivation,
//    which is never below its true best cost (pruned determinization does not promise more, in the reference either).
// The wrapper's phone-level first pass (an efficiency device: it does not change the accepted language) is
// b2k_lat_determinize_phone_pruned below.  Not reproduced: max_mem / max_loop early stopping, minimization (off by default,
// DeterminizeLatticePhonePrunedOptions), and therefore the STATE NUMBERING of the reference's output.
// PARITY: pinned by equivalence.  The reference's own lat/determinize-lattice-pruned.cc is compiled in oracle/_ref
// against a container-only OpenFst stand-in (oracle/ref_det.py, oracle/ref_wrap/fst_stub_det/) and run on the same raw
// lattices, with and without its phone-level pass: same word sequences within the beam, same weights, same
// transition-id strings (tests/test_lattice_det.py; also exhaustively against path enumeration on small lattices, the
// property the reference's determinize-lattice-pruned-test.cc checks).  The state numbering is not compared.
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <limits>
#include <map>
#include <queue>
#include <stdexcept>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

#include "common.cuh"

namespace {

struct W { float g = 0.f, a = 0.f; };                      // LatticeWeight: Value1 = graph, Value2 = acoustic
inline int cmp_w(const W &x, const W &y) {                  // 1: x better (lattice-weight.h:295-308)
  const float fx = x.g + x.a, fy = y.g + y.a;
  if (fx < fy) return 1;
  if (fx > fy) return -1;
  if (x.g < y.g) return 1;
  if (x.g > y.g) return -1;
  return 0;
}
inline W times(const W &x, const W &y) { return W{x.g + y.g, x.a + y.a}; }
inline W divide(const W &x, const W &y) { return W{x.g - y.g, x.a - y.a}; }

// an element's transition-id string lives in one append-only pool per call: (offset, length) — taking a prefix off is a
// change of offset, extending copies into the pool's tail, nothing is allocated per element
struct Elem { int32_t state; W w; int64_t s_off = 0; int32_t s_len = 0; };
inline int cmp_span(const int32_t *x, int32_t nx, const int32_t *y, int32_t ny) {     // 1: x better (:594-603)
  if (nx > ny) return -1;
  if (nx < ny) return 1;
  for (int32_t i = 0; i < nx; i++) { if (x[i] < y[i]) return -1; if (x[i] > y[i]) return 1; }
  return 0;
}

struct Key {                                               // identity of a determinized state
  std::vector<int32_t> ints;                                // per element: state, g bits, a bits, len, string...
  bool operator==(const Key &o) const { return ints == o.ints; }
};
struct KeyHash {
  size_t operator()(const Key &k) const { size_t h = 1469598103934665603ull; for (int32_t v : k.ints) { h ^= (uint32_t)v; h *= 1099511628211ull; } return h; }
};

}  // namespace

struct b2k_clat {
  std::vector<int32_t> arc_src, arc_dst, arc_word, final_state, tids;
  std::vector<float> arc_g, arc_a, final_g, final_a;
  std::vector<int64_t> arc_str_off, final_str_off;          // [n+1] offsets into tids (arcs first, then finals)
  int64_t num_states = 0;
  int64_t subsets_expanded = 0, elements_total = 0;
  float effective_beam = 0.f;
};

// returns B2K_ERR_OVERFLOW (and no object) when more than max_states determinized states were created
static int determinize_once(const b2k_raw_lattice *in, float beam, int64_t max_states, b2k_clat **out) {
  if (!in || !out || in->num_states < 0 || in->num_arcs < 0 || in->num_finals < 0 || !(beam > 0.f))
    return b2k::set_error(B2K_ERR_INVALID, "b2k_lat_determinize_pruned: bad args (beam must be positive)");
  if (in->num_states > 0 && (!in->arc_src || !in->arc_dst || !in->arc_ilabel || !in->arc_olabel || !in->arc_graph_cost ||
                             !in->arc_acoustic_cost || (in->num_finals > 0 && (!in->final_state || !in->final_cost))))
    return b2k::set_error(B2K_ERR_INVALID, "b2k_lat_determinize_pruned: the raw lattice arrays are missing");
  b2k_clat *C = new b2k_clat();
  C->arc_str_off.push_back(0);
  const int64_t N = in->num_states, A = in->num_arcs;
  if (N == 0 || in->num_finals == 0) { C->final_str_off.push_back(0); *out = C; return B2K_OK; }   // empty lattice -> empty result
  for (int64_t a = 0; a < A; a++)
    if (in->arc_src[a] < 0 || in->arc_src[a] >= N || in->arc_dst[a] < 0 || in->arc_dst[a] >= N) { delete C; return b2k::set_error(B2K_ERR_INVALID, "b2k_lat_determinize_pruned: arc endpoint out of range"); }
  // CSR by source state, arcs kept in input order
  std::vector<int64_t> off(N + 1, 0);
  for (int64_t a = 0; a < A; a++) off[in->arc_src[a] + 1]++;
  for (int64_t s = 0; s < N; s++) off[s + 1] += off[s];
  std::vector<int64_t> arcs(A), fill(off.begin(), off.end() - 1);
  for (int64_t a = 0; a < A; a++) arcs[fill[in->arc_src[a]]++] = a;
  std::vector<float> final_cost(N, std::numeric_limits<float>::infinity());
  for (int64_t f = 0; f < in->num_finals; f++) {
    if (in->final_state[f] < 0 || in->final_state[f] >= N) { delete C; return b2k::set_error(B2K_ERR_INVALID, "b2k_lat_determinize_pruned: final state out of range"); }
    final_cost[in->final_state[f]] = std::min(final_cost[in->final_state[f]], in->final_cost[f]);
  }
  // topological order (the raw lattice is acyclic: lattice-faster-decoder.cc:995 asserts no epsilon cycles)
  std::vector<int32_t> indeg(N, 0), topo;
  for (int64_t a = 0; a < A; a++) indeg[in->arc_dst[a]]++;
  topo.reserve(N);
  for (int64_t s = 0; s < N; s++) if (!indeg[s]) topo.push_back((int32_t)s);
  for (size_t i = 0; i < topo.size(); i++)
    for (int64_t k = off[topo[i]]; k < off[topo[i] + 1]; k++) if (--indeg[in->arc_dst[arcs[k]]] == 0) topo.push_back(in->arc_dst[arcs[k]]);
  if ((int64_t)topo.size() != N) { delete C; return b2k::set_error(B2K_ERR_INVALID, "b2k_lat_determinize_pruned: the raw lattice has a cycle"); }
  std::vector<int32_t> rank(N);
  for (int64_t i = 0; i < N; i++) rank[topo[i]] = (int32_t)i;
  // backward best cost to a final state (double), for the pruning criterion
  const double INF = std::numeric_limits<double>::infinity();
  std::vector<double> beta(N, INF);
  for (int64_t i = N - 1; i >= 0; i--) {
    const int32_t s = topo[i];
    double b = std::isfinite(final_cost[s]) ? (double)final_cost[s] : INF;
    for (int64_t k = off[s]; k < off[s + 1]; k++) {
      const int64_t a = arcs[k];
      b = std::min(b, (double)in->arc_graph_cost[a] + (double)in->arc_acoustic_cost[a] + beta[in->arc_dst[a]]);
    }
    beta[s] = b;
  }
  if (!std::isfinite(beta[0])) { C->final_str_off.push_back(0); *out = C; return B2K_OK; }   // no final state reachable
  const double cutoff = beta[0] + (double)beam;
  std::vector<int32_t> pool;
  pool.reserve((size_t)A * 4 + 1024);
  auto better = [&](const Elem &x, const Elem &y) {
    const int c = cmp_w(x.w, y.w);
    return c ? c > 0 : cmp_span(pool.data() + x.s_off, x.s_len, pool.data() + y.s_off, y.s_len) > 0;
  };
  // string of `src` plus one transition-id (0 = nothing to add) for a new element
  auto extend = [&](const Elem &src, int32_t tid, Elem *dst) {
    if (tid == 0) { dst->s_off = src.s_off; dst->s_len = src.s_len; return; }
    const size_t off = pool.size(), need = off + (size_t)src.s_len + 1;
    if (pool.capacity() < need) pool.reserve(std::max(need, 2 * pool.capacity()));
    pool.resize(need);                                       // no reallocation after the reserve: source and target do not overlap
    if (src.s_len) memcpy(&pool[off], &pool[(size_t)src.s_off], 4 * (size_t)src.s_len);
    pool[off + (size_t)src.s_len] = tid;
    dst->s_off = (int64_t)off; dst->s_len = src.s_len + 1;
  };

  // closure over word-epsilon arcs in topological order, best element per input state; elements that cannot lie on a
  // path within the beam are dropped; then the common weight and the common string prefix are split off
  struct Subset { std::vector<Elem> elems; double alpha; int32_t min_rank; };
  std::vector<Subset> subsets;
  std::unordered_map<Key, int32_t, KeyHash> index;
  std::vector<int32_t> tmp_slot(N, -1);
  auto closure_and_normalise = [&](std::vector<Elem> &seed, double alpha, W *common_w, std::vector<int32_t> *common_s) -> bool {
    // seed: arbitrary elements; returns false if nothing survives the beam
    std::vector<Elem> best;                                  // one per state
    std::vector<int32_t> touched;
    auto offer = [&](Elem &&e) {
      // beam: the bound alpha + w + beta never decreases along an arc, so an element outside the beam has no
      // descendant inside it and the closure need not walk through it
      if (!(alpha + (double)e.w.g + (double)e.w.a + beta[e.state] <= cutoff)) return;
      int32_t &slot = tmp_slot[e.state];
      if (slot < 0) { slot = (int32_t)best.size(); touched.push_back(e.state); best.push_back(std::move(e)); }
      else if (better(e, best[slot])) best[slot] = std::move(e);
    };
    for (auto &e : seed) offer(std::move(e));
    // states in topological rank order (each enters the heap once, when its slot is created): when a state is popped
    // every closure predecessor has been expanded, so its best element is final
    std::priority_queue<std::pair<int32_t, int32_t>, std::vector<std::pair<int32_t, int32_t>>, std::greater<>> pq;
    for (int32_t s : touched) pq.push({rank[s], s});
    while (!pq.empty()) {
      const int32_t s = pq.top().second;
      pq.pop();
      const Elem cur = best[tmp_slot[s]];                    // final for s: all predecessors have lower rank
      for (int64_t k = off[s]; k < off[s + 1]; k++) {
        const int64_t a = arcs[k];
        if (in->arc_olabel[a] != 0) continue;
        Elem e;
        e.state = in->arc_dst[a];
        e.w = times(cur.w, W{in->arc_graph_cost[a], in->arc_acoustic_cost[a]});
        extend(cur, in->arc_ilabel[a], &e);
        const bool wasnew = tmp_slot[e.state] < 0;
        offer(std::move(e));
        if (wasnew && tmp_slot[in->arc_dst[a]] >= 0) pq.push({rank[in->arc_dst[a]], in->arc_dst[a]});
      }
    }
    std::vector<Elem> kept(std::move(best));
    for (int32_t s : touched) tmp_slot[s] = -1;
    if (kept.empty()) return false;
    std::sort(kept.begin(), kept.end(), [](const Elem &x, const Elem &y) { return x.state < y.state; });
    W cw = kept[0].w;
    for (auto &e : kept) if (cmp_w(e.w, cw) > 0) cw = e.w;
    int32_t pre = kept[0].s_len;
    for (auto &e : kept) {
      int32_t k = 0;
      while (k < pre && k < e.s_len && pool[e.s_off + k] == pool[kept[0].s_off + k]) k++;
      pre = k;
    }
    common_s->assign(pool.begin() + kept[0].s_off, pool.begin() + kept[0].s_off + pre);
    for (auto &e : kept) { e.w = divide(e.w, cw); e.s_off += pre; e.s_len -= pre; }
    *common_w = cw;
    seed.swap(kept);
    return true;
  };
  // Expansion order: a subset's elements all descend from elements of its parent over at least one arc, so the
  // smallest topological rank among its elements is strictly larger than its parent's.  Popping subsets by that
  // rank therefore expands every possible parent of a subset before the subset itself: its alpha (the cheapest
  // determinized prefix reaching it) is final when it is expanded, and pruning its children against that alpha
  // never drops a word sequence that is within the beam.
  std::priority_queue<std::pair<int32_t, int32_t>, std::vector<std::pair<int32_t, int32_t>>, std::greater<>> agenda;
  auto intern = [&](std::vector<Elem> &elems, double alpha) -> int32_t {
    Key k;
    int32_t mr = std::numeric_limits<int32_t>::max();
    for (auto &e : elems) {
      int32_t gb, ab;
      memcpy(&gb, &e.w.g, 4); memcpy(&ab, &e.w.a, 4);
      k.ints.push_back(e.state); k.ints.push_back(gb); k.ints.push_back(ab); k.ints.push_back(e.s_len);
      k.ints.insert(k.ints.end(), pool.begin() + e.s_off, pool.begin() + e.s_off + e.s_len);
      mr = std::min(mr, rank[e.state]);
    }
    auto it = index.find(k);
    if (it != index.end()) { subsets[it->second].alpha = std::min(subsets[it->second].alpha, alpha); return it->second; }
    const int32_t id = (int32_t)subsets.size();
    index.emplace(std::move(k), id);
    subsets.push_back(Subset{std::move(elems), alpha, mr});
    agenda.push({mr, id});
    return id;
  };

  // start: closure of input state 0; whatever it has in common precedes every path and goes to a leading arc-less
  // position: it is folded into the outgoing arcs / final weight of the start state
  std::vector<Elem> seed(1);
  seed[0].state = 0;
  W start_w; std::vector<int32_t> start_s;
  if (!closure_and_normalise(seed, 0.0, &start_w, &start_s)) { C->final_str_off.push_back(0); *out = C; return B2K_OK; }
  intern(seed, (double)start_w.g + (double)start_w.a);

  std::vector<float> fin_g, fin_a; std::vector<int32_t> fin_state; std::vector<std::vector<int32_t>> fin_str, arc_str;
  while (!agenda.empty()) {
    const size_t cur = (size_t)agenda.top().second;
    agenda.pop();
    C->subsets_expanded++;
    if (max_states > 0 && (int64_t)subsets.size() > max_states) { delete C; return B2K_ERR_OVERFLOW; }
    const double alpha = subsets[cur].alpha;
    const std::vector<Elem> elems = subsets[cur].elems;       // copy: `subsets` may reallocate
    C->elements_total += (int64_t)elems.size();
    // final weight: best over elements of elem (x) Final(state)
    {
      bool have = false; Elem bestf;
      for (auto &e : elems) if (std::isfinite(final_cost[e.state])) {
        Elem f = e; f.w = times(e.w, W{final_cost[e.state], 0.f});
        if (!have || better(f, bestf)) { bestf = std::move(f); have = true; }
      }
      if (have && alpha + (double)bestf.w.g + (double)bestf.w.a <= cutoff) {
        W w = bestf.w; std::vector<int32_t> s(pool.begin() + bestf.s_off, pool.begin() + bestf.s_off + bestf.s_len);
        if (cur == 0) { w = times(start_w, w); s.insert(s.begin(), start_s.begin(), start_s.end()); }
        fin_state.push_back((int32_t)cur); fin_g.push_back(w.g); fin_a.push_back(w.a); fin_str.push_back(std::move(s));
      }
    }
    // transitions grouped by word, words in increasing order (deterministic output)
    std::map<int32_t, std::vector<Elem>> by_word;
    for (auto &e : elems)
      for (int64_t k = off[e.state]; k < off[e.state + 1]; k++) {
        const int64_t a = arcs[k];
        const int32_t word = in->arc_olabel[a];
        if (word == 0) continue;
        Elem n; n.state = in->arc_dst[a];
        n.w = times(e.w, W{in->arc_graph_cost[a], in->arc_acoustic_cost[a]});
        extend(e, in->arc_ilabel[a], &n);
        by_word[word].push_back(std::move(n));
      }
    for (auto &kv : by_word) {
      W cw; std::vector<int32_t> cs;
      if (!closure_and_normalise(kv.second, alpha, &cw, &cs)) continue;
      const double nalpha = alpha + (double)cw.g + (double)cw.a;
      const int32_t dst = intern(kv.second, nalpha);
      if (cur == 0) { cw = times(start_w, cw); cs.insert(cs.begin(), start_s.begin(), start_s.end()); }
      C->arc_src.push_back((int32_t)cur); C->arc_dst.push_back(dst); C->arc_word.push_back(kv.first);
      C->arc_g.push_back(cw.g); C->arc_a.push_back(cw.a); arc_str.push_back(std::move(cs));
    }
  }
  // (the start subset cannot be re-entered: ranks strictly increase along determinized arcs)
  C->num_states = (int64_t)subsets.size();
  for (auto &s : arc_str) { C->tids.insert(C->tids.end(), s.begin(), s.end()); C->arc_str_off.push_back((int64_t)C->tids.size()); }
  C->final_str_off.push_back((int64_t)C->tids.size());
  for (size_t i = 0; i < fin_state.size(); i++) {
    C->final_state.push_back(fin_state[i]); C->final_g.push_back(fin_g[i]); C->final_a.push_back(fin_a[i]);
    C->tids.insert(C->tids.end(), fin_str[i].begin(), fin_str[i].end());
    C->final_str_off.push_back((int64_t)C->tids.size());
  }
  *out = C;
  return B2K_OK;
}

extern "C" {

int b2k_lat_determinize_pruned(const b2k_raw_lattice *in, float beam, int64_t max_states, b2k_clat **out) {
  if (!out) return b2k::set_error(B2K_ERR_INVALID, "b2k_lat_determinize_pruned: bad args");
  *out = nullptr;
  // Size guard in the spirit of the reference's max_mem handling (determinize-lattice-pruned.cc: when the memory
  // limit is hit the beam is reduced and the work continues; the tool then reports the "effective beam"): here the
  // determinization is simply redone with 3/4 of the beam until the state budget holds.
  float b = beam;
  for (int attempt = 0; attempt < 24; attempt++) {
    int rc = determinize_once(in, b, max_states, out);
    if (rc == B2K_OK) { (*out)->effective_beam = b; return B2K_OK; }
    if (rc != B2K_ERR_OVERFLOW) return rc;
    b *= 0.75f;
  }
  return b2k::set_error(B2K_ERR_OVERFLOW, "b2k_lat_determinize_pruned: the state budget cannot be met even with a tiny beam");
}

// n independent lattices on up to num_threads host threads (the reference determinizes on a thread pool as well,
// cudadecoder/batched-threaded-nnet3-cuda-online-pipeline.cc:727-790).  out[i] / status[i] per lattice; returns the
// first non-zero status.  Results are those of the single-lattice call: nothing is shared between lattices.
int b2k_lat_determinize_pruned_batch(const b2k_raw_lattice *in, int32_t n, float beam, int64_t max_states, int32_t num_threads,
                                     b2k_clat **out, int32_t *status) {
  if (!in || !out || n < 0) return b2k::set_error(B2K_ERR_INVALID, "b2k_lat_determinize_pruned_batch: bad args");
  for (int32_t i = 0; i < n; i++) out[i] = nullptr;
  std::vector<int> rc((size_t)n, 0);
  std::vector<std::string> msg((size_t)n);
  std::atomic<int32_t> next{0};
  auto work = [&]() {
    for (int32_t i = next.fetch_add(1); i < n; i = next.fetch_add(1)) {
      rc[i] = b2k_lat_determinize_pruned(&in[i], beam, max_states, &out[i]);
      if (rc[i]) msg[i] = b2k::g_last_error;                   // thread-local in the worker: carry it to the caller
    }
  };
  const int32_t nt = std::max<int32_t>(1, std::min<int32_t>(num_threads > 0 ? num_threads : (int32_t)std::thread::hardware_concurrency(), n));
  if (nt <= 1) work();
  else {
    std::vector<std::thread> th;
    for (int32_t t = 0; t < nt; t++) th.emplace_back(work);
    for (auto &t : th) t.join();
  }
  int first = 0;
  for (int32_t i = 0; i < n; i++) {
    if (status) status[i] = rc[i];
    if (rc[i] && !first) { first = rc[i]; b2k::g_last_error = msg[i]; }
  }
  return first;
}

// The two-pass form of DeterminizeLatticePhonePruned (lat/determinize-lattice-pruned.cc:1291-1470): first determinize over words
// AND phone labels -- a phone label is put on every arc that starts a phone (not a self-loop, not leaving the start state), on the
// arc itself when it carries no word, behind it on an extra arc otherwise (DeterminizeLatticeInsertPhones :1291-1345) -- which
// merges derivations early, phone by phone, instead of carrying whole-word sets of alternatives; then delete the phone labels
// (:1347-1370), expand the compact arcs back into transition-id arcs (ConvertLattice) and determinize over words.  Every word
// sequence within the beam keeps its best derivation through both passes (its phone sequence is within the beam too), so the
// result accepts the same sequences with the same weights and alignments as the one-pass form.
int b2k_lat_determinize_phone_pruned(const b2k_raw_lattice *in, float beam, int64_t max_states, const int32_t *phone_of,
                                     const uint8_t *self_loop, const uint8_t *phone_start, int32_t num_tids, int32_t phone_determinize,
                                     int32_t word_determinize, b2k_clat **out) {
  if (!out) return b2k::set_error(B2K_ERR_INVALID, "b2k_lat_determinize_phone_pruned: bad args");
  *out = nullptr;
  if (!phone_determinize && !word_determinize)
    return b2k::set_error(B2K_ERR_INVALID, "b2k_lat_determinize_phone_pruned: both passes are off (the reference copies the lattice then; ask for at least one)");
  if (!phone_determinize) return b2k_lat_determinize_pruned(in, beam, max_states, out);
  if (!in || !phone_of || !self_loop || !phone_start || num_tids <= 0 || in->num_states < 0 || in->num_arcs < 0)
    return b2k::set_error(B2K_ERR_INVALID, "b2k_lat_determinize_phone_pruned: bad args");
  if (in->num_states > 0 && (!in->arc_src || !in->arc_dst || !in->arc_ilabel || !in->arc_olabel || !in->arc_graph_cost || !in->arc_acoustic_cost))
    return b2k::set_error(B2K_ERR_INVALID, "b2k_lat_determinize_phone_pruned: the raw lattice arrays are missing");
  // ---- pass 1 input: phones inserted
  int32_t first_phone_label = 1;                             // HighestNumberedInputSymbol + 1 (words are the labels here)
  for (int64_t a = 0; a < in->num_arcs; a++) first_phone_label = std::max(first_phone_label, in->arc_olabel[a] + 1);
  std::vector<int32_t> src(in->arc_src, in->arc_src + in->num_arcs), dst(in->arc_dst, in->arc_dst + in->num_arcs),
      il(in->arc_ilabel, in->arc_ilabel + in->num_arcs), ol(in->arc_olabel, in->arc_olabel + in->num_arcs);
  std::vector<float> g(in->arc_graph_cost, in->arc_graph_cost + in->num_arcs), ac(in->arc_acoustic_cost, in->arc_acoustic_cost + in->num_arcs);
  int64_t nstates = in->num_states;
  for (int64_t a = 0; a < in->num_arcs; a++) {
    const int32_t tid = il[a];
    if (tid < 0 || tid >= num_tids) return b2k::set_error(B2K_ERR_INVALID, "b2k_lat_determinize_phone_pruned: transition-id outside the table");
    if (src[a] == 0 || tid == 0 || !phone_start[tid] || self_loop[tid]) continue;     // (:1307-1308 skips the start state's arcs)
    const int32_t phone = phone_of[tid];
    if (phone <= 0) return b2k::set_error(B2K_ERR_INVALID, "b2k_lat_determinize_phone_pruned: phone 0 for a transition-id");      // :1321
    if (ol[a] == 0) { ol[a] = first_phone_label + phone; continue; }
    const int32_t x = (int32_t)nstates++;                   // word arc first, the phone on an extra arc behind it
    src.push_back(x); dst.push_back(dst[a]); il.push_back(0); ol.push_back(first_phone_label + phone); g.push_back(0.f); ac.push_back(0.f);
    dst[a] = x;
  }
  b2k_raw_lattice l1 = *in;
  l1.num_states = nstates; l1.num_arcs = (int64_t)src.size();
  l1.arc_src = src.data(); l1.arc_dst = dst.data(); l1.arc_ilabel = il.data(); l1.arc_olabel = ol.data();
  l1.arc_graph_cost = g.data(); l1.arc_acoustic_cost = ac.data();
  b2k_clat *c1 = nullptr;
  int rc = b2k_lat_determinize_pruned(&l1, beam, max_states, &c1);
  if (rc) return rc;
  // ---- phones deleted, compact arcs expanded into transition-id arcs (word and weight on the first arc of a chain)
  std::vector<int32_t> s2, d2, i2, o2, fs2;
  std::vector<float> g2, a2, fc2;
  int64_t n2 = c1->num_states;
  auto chain = [&](int32_t from, int32_t to, int32_t label, float wg, float wa, const int32_t *str, int64_t len) {
    // to < 0: a fresh end state (returned)
    const int64_t n = std::max<int64_t>(len, 1);
    int32_t cur = from;
    for (int64_t j = 0; j < n; j++) {
      const int32_t nxt = (j + 1 < n || to < 0) ? (int32_t)n2++ : to;
      s2.push_back(cur); d2.push_back(nxt); i2.push_back(j < len ? str[j] : 0); o2.push_back(j == 0 ? label : 0);
      g2.push_back(j == 0 ? wg : 0.f); a2.push_back(j == 0 ? wa : 0.f);
      cur = nxt;
    }
    return cur;
  };
  for (size_t i = 0; i < c1->arc_src.size(); i++) {
    const int32_t lab = c1->arc_word[i] >= first_phone_label ? 0 : c1->arc_word[i];
    chain(c1->arc_src[i], c1->arc_dst[i], lab, c1->arc_g[i], c1->arc_a[i], c1->tids.data() + c1->arc_str_off[i], c1->arc_str_off[i + 1] - c1->arc_str_off[i]);
  }
  for (size_t f = 0; f < c1->final_state.size(); f++) {      // a compact final weight has an acoustic part and a string: an arc to a fresh final state
    const int32_t e = chain(c1->final_state[f], -1, 0, c1->final_g[f], c1->final_a[f], c1->tids.data() + c1->final_str_off[f],
                            c1->final_str_off[f + 1] - c1->final_str_off[f]);
    fs2.push_back(e); fc2.push_back(0.f);
  }
  const int64_t expanded1 = c1->subsets_expanded, elems1 = c1->elements_total;
  const float eff1 = c1->effective_beam;
  if (!word_determinize) {
    // "ConvertLattice(*ifst, ofst, false)" (:1448-1451): the phone-level result as a compact lattice, one transition-id per arc
    b2k_clat *C = new b2k_clat();
    C->num_states = n2; C->arc_src = s2; C->arc_dst = d2; C->arc_word = o2; C->arc_g = g2; C->arc_a = a2;
    C->arc_str_off.push_back(0);
    for (size_t i = 0; i < s2.size(); i++) { if (i2[i] != 0) C->tids.push_back(i2[i]); C->arc_str_off.push_back((int64_t)C->tids.size()); }
    C->final_str_off.push_back((int64_t)C->tids.size());
    for (size_t f = 0; f < fs2.size(); f++) { C->final_state.push_back(fs2[f]); C->final_g.push_back(0.f); C->final_a.push_back(0.f); C->final_str_off.push_back((int64_t)C->tids.size()); }
    C->subsets_expanded = expanded1; C->elements_total = elems1; C->effective_beam = eff1;
    delete c1;
    *out = C;
    return B2K_OK;
  }
  delete c1;
  b2k_raw_lattice l2;
  memset(&l2, 0, sizeof(l2));
  l2.num_states = n2; l2.num_arcs = (int64_t)s2.size(); l2.num_finals = (int64_t)fs2.size();
  l2.arc_src = s2.data(); l2.arc_dst = d2.data(); l2.arc_ilabel = i2.data(); l2.arc_olabel = o2.data();
  l2.arc_graph_cost = g2.data(); l2.arc_acoustic_cost = a2.data(); l2.final_state = fs2.data(); l2.final_cost = fc2.data();
  rc = b2k_lat_determinize_pruned(&l2, beam, max_states, out);
  if (rc) return rc;
  (*out)->subsets_expanded += expanded1; (*out)->elements_total += elems1;
  (*out)->effective_beam = std::min((*out)->effective_beam, eff1);
  return B2K_OK;
}

int b2k_clat_destroy(b2k_clat *c) { delete c; return B2K_OK; }

}  // extern "C"

// ---------------------------------------------------------------------------------------------------------------------
// Push + minimize: what DeterminizeLatticePhonePruned does after determinization under --minimize
// (lat/determinize-lattice-pruned.cc:1459-1465): PushCompactLatticeStrings, PushCompactLatticeWeights (lat/push-lattice.cc),
// MinimizeCompactLattice (lat/minimize-lattice.cc).  The lattice stays an acyclic deterministic acceptor over words with the
// same language, path weights (up to float rounding of the re-distribution) and alignments; transition-ids move as early and
// costs as early as they can, after which states with the same future are one state.
namespace {

struct LW {                                   // LatticeWeight: (graph, acoustic), compared by the sum, then by the first
  float g, a;
  static LW Zero() { const float inf = std::numeric_limits<float>::infinity(); return LW{inf, inf}; }
  bool IsZero() const { const float inf = std::numeric_limits<float>::infinity(); return g == inf && a == inf; }
};
inline LW lw_times(const LW &x, const LW &y) { return LW{x.g + y.g, x.a + y.a}; }
inline LW lw_plus(const LW &x, const LW &y) {            // the better of the two (fstext/lattice-weight.h:295-315)
  const float fx = x.g + x.a, fy = y.g + y.a;
  if (fx < fy) return x;
  if (fx > fy) return y;
  return x.g <= y.g ? x : y;
}
inline LW lw_divide(const LW &x, const LW &y) {          // :371-386: anything that is not a number is Zero
  const float inf = std::numeric_limits<float>::infinity();
  const float g = x.g - y.g, a = x.a - y.a;
  if (g != g || a != a || g == -inf || a == -inf || g == inf || a == inf) return LW::Zero();
  return LW{g, a};
}
inline bool lw_approx_equal(const LW &x, const LW &y, float delta) {       // :390-395
  if (x.g == y.g && x.a == y.a) return true;
  return std::fabs((x.g + x.a) - (y.g + y.a)) <= delta;
}

struct MArc { int32_t dst, word; LW w; std::vector<int32_t> str; };
struct MState { std::vector<MArc> arcs; bool is_final = false; LW fw = LW::Zero(); std::vector<int32_t> fstr; };

// the first n transition-ids met on a path out of state s (any path: the callers only ask where every path agrees)
void first_tids(const std::vector<MState> &st, int32_t s, size_t n, std::vector<int32_t> *out) {
  while (n > 0) {
    const MState &q = st[s];
    const std::vector<int32_t> *str;
    int32_t next = -1;
    if (q.is_final) str = &q.fstr;
    else if (!q.arcs.empty()) { str = &q.arcs[0].str; next = q.arcs[0].dst; }
    else return;                                          // a dead end: nothing more to give
    const size_t take = std::min(n, str->size());
    out->insert(out->end(), str->begin(), str->begin() + take);
    n -= take;
    if (next < 0) return;
    s = next;
  }
}

}  // namespace

extern "C" {

int b2k_clat_minimize(b2k_clat *c, float delta) {
  if (!c || !(delta >= 0.f)) return b2k::set_error(B2K_ERR_INVALID, "b2k_clat_minimize: bad args");
  const int64_t N = c->num_states, A = (int64_t)c->arc_src.size(), F = (int64_t)c->final_state.size();
  if (N == 0) return B2K_OK;                               // an empty lattice is pushed and minimal
  std::vector<MState> st(N);
  for (int64_t a = 0; a < A; a++) {
    MArc m;
    m.dst = c->arc_dst[a]; m.word = c->arc_word[a]; m.w = LW{c->arc_g[a], c->arc_a[a]};
    m.str.assign(c->tids.begin() + c->arc_str_off[a], c->tids.begin() + c->arc_str_off[a + 1]);
    st[c->arc_src[a]].arcs.push_back(std::move(m));
  }
  for (int64_t f = 0; f < F; f++) {
    MState &q = st[c->final_state[f]];
    q.is_final = true; q.fw = LW{c->final_g[f], c->final_a[f]};
    q.fstr.assign(c->tids.begin() + c->final_str_off[f], c->tids.begin() + c->final_str_off[f + 1]);
  }
  // topological order from the start state (0): depth first, a state after everything it reaches, then reversed
  std::vector<int32_t> order;
  {
    std::vector<uint8_t> mark(N, 0);                       // 1 = on the stack, 2 = done
    std::vector<std::pair<int32_t, size_t> > stack;
    stack.push_back(std::make_pair(0, (size_t)0));
    mark[0] = 1;
    while (!stack.empty()) {
      const int32_t s = stack.back().first;
      if (stack.back().second < st[s].arcs.size()) {
        const int32_t d = st[s].arcs[stack.back().second++].dst;
        if (mark[d] == 1) return b2k::set_error(B2K_ERR_STATE, "b2k_clat_minimize: the lattice has a cycle");
        if (mark[d] == 0) { mark[d] = 1; stack.push_back(std::make_pair(d, (size_t)0)); }
      } else {
        mark[s] = 2; order.push_back(s); stack.pop_back();
      }
    }
    std::reverse(order.begin(), order.end());              // states the start does not reach are dropped at the end
  }
  const size_t R = order.size();

  // 1. strings: shift[s] = how many transition-ids every path out of s starts with, so that they can sit on the arcs INTO s
  {
    std::vector<int64_t> shift(N, 0);
    std::vector<int32_t> ref, cmp;
    for (size_t k = R; k-- > 1;) {                         // not the start state: nothing comes before it
      const int32_t s = order[k];
      const MState &q = st[s];
      if (q.arcs.empty()) { shift[s] = (int64_t)q.fstr.size(); continue; }
      int64_t sh = q.is_final ? (int64_t)q.fstr.size() : std::numeric_limits<int64_t>::max();
      for (const MArc &m : q.arcs) sh = std::min(sh, shift[m.dst] + (int64_t)m.str.size());
      if (q.arcs.size() + (q.is_final ? 1 : 0) > 1 && sh > 0) {      // several ways on: keep what they have in common
        ref.clear();
        size_t from = 0;
        if (q.is_final) ref.assign(q.fstr.begin(), q.fstr.begin() + sh);
        else {
          ref.assign(q.arcs[0].str.begin(), q.arcs[0].str.begin() + std::min<size_t>(sh, q.arcs[0].str.size()));
          first_tids(st, q.arcs[0].dst, sh - ref.size(), &ref);
          from = 1;
        }
        for (size_t i = from; i < q.arcs.size() && sh > 0; i++) {
          const MArc &m = q.arcs[i];
          cmp.assign(m.str.begin(), m.str.begin() + std::min<size_t>(sh, m.str.size()));
          first_tids(st, m.dst, sh - cmp.size(), &cmp);
          int64_t same = 0;
          while (same < sh && same < (int64_t)ref.size() && same < (int64_t)cmp.size() && ref[same] == cmp[same]) same++;
          sh = same;
        }
      }
      shift[s] = sh;
    }
    std::vector<std::vector<std::vector<int32_t> > > new_str(N);     // computed from the old strings, then swapped in
    for (size_t k = 0; k < R; k++) {
      const int32_t s = order[k];
      new_str[s].resize(st[s].arcs.size());
      for (size_t i = 0; i < st[s].arcs.size(); i++) {
        const MArc &m = st[s].arcs[i];
        std::vector<int32_t> full(m.str);
        first_tids(st, m.dst, (size_t)shift[m.dst], &full);
        new_str[s][i].assign(full.begin() + shift[s], full.end());
      }
    }
    for (size_t k = 0; k < R; k++) {
      const int32_t s = order[k];
      for (size_t i = 0; i < st[s].arcs.size(); i++) st[s].arcs[i].str.swap(new_str[s][i]);
      if (st[s].is_final) st[s].fstr.erase(st[s].fstr.begin(), st[s].fstr.begin() + shift[s]);
    }
  }

  // 2. weights: to_end[s] = the best weight from s to the end; every arc then carries its share, the start keeps the rest
  {
    std::vector<LW> to_end(N, LW::Zero());
    for (size_t k = R; k-- > 0;) {
      const int32_t s = order[k];
      LW w = st[s].fw;
      for (const MArc &m : st[s].arcs) w = lw_plus(w, lw_times(m.w, to_end[m.dst]));
      to_end[s] = w;
    }
    to_end[0] = LW{0.f, 0.f};
    for (size_t k = 0; k < R; k++) {
      const int32_t s = order[k];
      if (to_end[s].IsZero()) continue;
      for (MArc &m : st[s].arcs)
        if (!to_end[m.dst].IsZero()) m.w = lw_times(m.w, lw_divide(to_end[m.dst], to_end[s]));
      if (st[s].is_final) st[s].fw = lw_divide(st[s].fw, to_end[s]);
    }
  }

  // 3. states with the same future become one: last states first, so that what an arc leads to is already a class
  std::vector<int32_t> cls(N);
  for (int64_t s = 0; s < N; s++) cls[s] = (int32_t)s;
  {
    struct Key { int32_t word, dst; const MArc *arc; };
    auto sorted_arcs = [&](int32_t s, std::vector<Key> *out) {
      out->clear();
      for (const MArc &m : st[s].arcs) out->push_back(Key{m.word, cls[m.dst], &m});
      std::sort(out->begin(), out->end(), [](const Key &x, const Key &y) { return x.word != y.word ? x.word < y.word : x.dst < y.dst; });
    };
    // an exact signature (words, classes reached, strings; not the costs) decides who is compared with whom
    auto signature = [&](int32_t s, const std::vector<Key> &keys) {
      uint64_t h = st[s].is_final ? 0x9E3779B97F4A7C15ull : 0x632BE59BD9B4E019ull;
      auto mix = [&h](uint64_t v) { h ^= v + 0x9E3779B97F4A7C15ull + (h << 6) + (h >> 2); };
      if (st[s].is_final) { mix(st[s].fstr.size()); for (int32_t t : st[s].fstr) mix((uint64_t)(uint32_t)t); }
      for (const Key &k : keys) {
        mix(0xABCDull); mix((uint64_t)(uint32_t)k.word); mix((uint64_t)(uint32_t)k.dst); mix(k.arc->str.size());
        for (int32_t t : k.arc->str) mix((uint64_t)(uint32_t)t);
      }
      return h;
    };
    std::unordered_map<uint64_t, std::vector<int32_t> > groups;      // representatives seen so far, latest last
    std::vector<Key> ks, kt;
    for (size_t k = R; k-- > 0;) {
      const int32_t s = order[k];
      sorted_arcs(s, &ks);
      std::vector<int32_t> &g = groups[signature(s, ks)];
      bool merged = false;
      for (size_t i = g.size(); i-- > 0 && !merged;) {     // candidates in the order the reference meets them: nearest the start first
        const int32_t t = g[i];
        if (st[s].is_final != st[t].is_final || st[s].arcs.size() != st[t].arcs.size()) continue;
        if (st[s].is_final && (!lw_approx_equal(st[s].fw, st[t].fw, delta) || st[s].fstr != st[t].fstr)) continue;
        sorted_arcs(t, &kt);
        bool same = true;
        for (size_t j = 0; j < ks.size() && same; j++)
          same = ks[j].word == kt[j].word && ks[j].dst == kt[j].dst && lw_approx_equal(ks[j].arc->w, kt[j].arc->w, 1.0f / 1024.0f) &&
                 ks[j].arc->str == kt[j].arc->str;
        if (same) { cls[s] = t; merged = true; }
      }
      if (!merged) g.push_back(s);
    }
  }

  // write back: the states the (possibly merged) start reaches, renumbered from it, arcs in their old order
  const int32_t start = cls[0];
  std::vector<int32_t> newid(N, -1), kept;
  {
    std::vector<int32_t> todo(1, start);
    newid[start] = 0; kept.push_back(start);
    while (!todo.empty()) {
      const int32_t s = todo.back(); todo.pop_back();
      for (const MArc &m : st[s].arcs) {
        const int32_t d = cls[m.dst];
        if (newid[d] < 0) { newid[d] = (int32_t)kept.size(); kept.push_back(d); todo.push_back(d); }
      }
    }
  }
  b2k_clat out;
  out.subsets_expanded = c->subsets_expanded; out.elements_total = c->elements_total; out.effective_beam = c->effective_beam;
  out.num_states = (int64_t)kept.size();
  out.arc_str_off.push_back(0);
  for (int32_t s : kept)
    for (const MArc &m : st[s].arcs) {
      out.arc_src.push_back(newid[s]); out.arc_dst.push_back(newid[cls[m.dst]]); out.arc_word.push_back(m.word);
      out.arc_g.push_back(m.w.g); out.arc_a.push_back(m.w.a);
      out.tids.insert(out.tids.end(), m.str.begin(), m.str.end());
      out.arc_str_off.push_back((int64_t)out.tids.size());
    }
  out.final_str_off.push_back((int64_t)out.tids.size());
  for (int32_t s : kept)
    if (st[s].is_final) {
      out.final_state.push_back(newid[s]); out.final_g.push_back(st[s].fw.g); out.final_a.push_back(st[s].fw.a);
      out.tids.insert(out.tids.end(), st[s].fstr.begin(), st[s].fstr.end());
      out.final_str_off.push_back((int64_t)out.tids.size());
    }
  *c = std::move(out);
  return B2K_OK;
}

// LatticeFasterDecoderTpl::GetBestPath (decoder/lattice-faster-decoder.cc:102-108 = GetRawLattice + ShortestPath over
// graph + acoustic) / CudaDecoder::GetBestPath (cudadecoder/cuda-decoder.h): the cheapest path from state 0 to a final
// state of a finalized raw lattice.  Ties: the first minimum in topological order of states and input order of arcs.
}  // extern "C"

// shared by the two entry points below: arcs of the cheapest start -> final path (in path order) and the final's index
static int best_path_core(const b2k_raw_lattice *in, std::vector<int64_t> *path_out, int64_t *best_final) {
  const int64_t N = in->num_states, A = in->num_arcs;
  path_out->clear();
  *best_final = -1;
  if (N == 0 || in->num_finals == 0) return B2K_OK;
  std::vector<int64_t> off(N + 1, 0);
  for (int64_t a = 0; a < A; a++) {
    if (in->arc_src[a] < 0 || in->arc_src[a] >= N || in->arc_dst[a] < 0 || in->arc_dst[a] >= N) return b2k::set_error(B2K_ERR_INVALID, "best path: arc endpoint out of range");
    off[in->arc_src[a] + 1]++;
  }
  for (int64_t s = 0; s < N; s++) off[s + 1] += off[s];
  std::vector<int64_t> arcs(A), fill(off.begin(), off.end() - 1);
  for (int64_t a = 0; a < A; a++) arcs[fill[in->arc_src[a]]++] = a;
  std::vector<int32_t> indeg(N, 0), topo;
  for (int64_t a = 0; a < A; a++) indeg[in->arc_dst[a]]++;
  for (int64_t s = 0; s < N; s++) if (!indeg[s]) topo.push_back((int32_t)s);
  for (size_t i = 0; i < topo.size(); i++)
    for (int64_t k = off[topo[i]]; k < off[topo[i] + 1]; k++) if (--indeg[in->arc_dst[arcs[k]]] == 0) topo.push_back(in->arc_dst[arcs[k]]);
  if ((int64_t)topo.size() != N) return b2k::set_error(B2K_ERR_INVALID, "best path: the raw lattice has a cycle");
  const double INF = std::numeric_limits<double>::infinity();
  std::vector<double> dist(N, INF);
  std::vector<int64_t> back(N, -1);
  dist[0] = 0.0;
  for (int32_t s : topo) {
    if (!(dist[s] < INF)) continue;
    for (int64_t k = off[s]; k < off[s + 1]; k++) {
      const int64_t a = arcs[k];
      const double nd = dist[s] + (double)in->arc_graph_cost[a] + (double)in->arc_acoustic_cost[a];
      if (nd < dist[in->arc_dst[a]]) { dist[in->arc_dst[a]] = nd; back[in->arc_dst[a]] = a; }
    }
  }
  int64_t best = -1;
  double best_cost = INF;
  for (int64_t f = 0; f < in->num_finals; f++) {
    const int32_t s = in->final_state[f];
    if (s < 0 || s >= N) return b2k::set_error(B2K_ERR_INVALID, "best path: final state out of range");
    const double c = dist[s] + (double)in->final_cost[f];
    if (c < best_cost) { best_cost = c; best = f; }
  }
  if (best < 0) return b2k::set_error(B2K_ERR_STATE, "best path: no final state is reachable from the start state");
  for (int32_t s = in->final_state[best]; s != 0;) { path_out->push_back(back[s]); s = in->arc_src[back[s]]; }
  std::reverse(path_out->begin(), path_out->end());
  *best_final = best;
  return B2K_OK;
}

extern "C" {

int b2k_lat_best_path(const b2k_raw_lattice *in, int32_t *words, int32_t *n_words, int32_t *tids, int32_t *n_tids, int32_t cap,
                      float *graph_cost, float *acoustic_cost) {
  if (!in || !n_words || !n_tids || cap < 0 || (cap > 0 && (!words || !tids)))
    return b2k::set_error(B2K_ERR_INVALID, "b2k_lat_best_path: bad args");
  *n_words = 0; *n_tids = 0;
  if (graph_cost) *graph_cost = std::numeric_limits<float>::infinity();
  if (acoustic_cost) *acoustic_cost = std::numeric_limits<float>::infinity();
  std::vector<int64_t> path;
  int64_t best = -1;
  const int rc = best_path_core(in, &path, &best);
  if (rc) return rc;
  if (best < 0) return B2K_OK;
  double g = (double)in->final_cost[best], ac = 0.0;
  int32_t nw = 0, nt = 0;
  for (const int64_t a : path) {
    g += (double)in->arc_graph_cost[a]; ac += (double)in->arc_acoustic_cost[a];
    if (in->arc_olabel[a] != 0) { if (nw < cap) words[nw] = in->arc_olabel[a]; nw++; }
    if (in->arc_ilabel[a] != 0) { if (nt < cap) tids[nt] = in->arc_ilabel[a]; nt++; }
  }
  *n_words = nw; *n_tids = nt;
  if (graph_cost) *graph_cost = (float)g;
  if (acoustic_cost) *acoustic_cost = (float)ac;
  if (nw > cap || nt > cap) return b2k::set_error(B2K_ERR_OVERFLOW, "b2k_lat_best_path: output buffers too small (sizes returned)");
  return B2K_OK;
}

int b2k_lat_best_path_arcs(const b2k_raw_lattice *in, int64_t *arcs, int64_t *n_arcs, int64_t cap, int64_t *final_index) {
  if (!in || !n_arcs || cap < 0 || (cap > 0 && !arcs)) return b2k::set_error(B2K_ERR_INVALID, "b2k_lat_best_path_arcs: bad args");
  std::vector<int64_t> path;
  int64_t best = -1;
  const int rc = best_path_core(in, &path, &best);
  if (rc) return rc;
  *n_arcs = (int64_t)path.size();
  if (final_index) *final_index = best;
  if ((int64_t)path.size() > cap) return b2k::set_error(B2K_ERR_OVERFLOW, "b2k_lat_best_path_arcs: output buffer too small (size returned)");
  if (!path.empty()) memcpy(arcs, path.data(), 8 * path.size());
  return B2K_OK;
}

// ---- table entries (lat/kaldi-lattice.cc WriteLattice / WriteCompactLattice): `key` + ' ' + "\0B" + VectorFst binary, or the
// text form an FstPrinter gives (weights printed by the reference's own operator<<: "%g"-style floats, "Infinity", strings
// joined by '_'; unit weights left out).  kaldi_b200/lattice.py writes the same bytes and is the test oracle of these two.
}  // extern "C"

namespace {
struct Out {
  FILE *f;
  void bytes(const void *p, size_t n) { if (n && fwrite(p, 1, n, f) != n) throw std::runtime_error("write failed"); }
  template <typename T> void put(T v) { bytes(&v, sizeof(T)); }
  void str(const std::string &s) { put<int32_t>((int32_t)s.size()); bytes(s.data(), s.size()); }
  void text(const std::string &s) { bytes(s.data(), s.size()); }
};
std::string num_text(float x) {
  if (x == std::numeric_limits<float>::infinity()) return "Infinity";
  if (x == -std::numeric_limits<float>::infinity()) return "-Infinity";
  if (x != x) return "BadNumber";
  char b[64];
  snprintf(b, sizeof(b), "%g", (double)x);
  return b;
}
void fst_header(Out &o, const char *arctype, int64_t start, int64_t ns) {
  o.put<int32_t>(2125659606); o.str("vector"); o.str(arctype); o.put<int32_t>(2); o.put<int32_t>(0);
  o.put<uint64_t>(0x3); o.put<int64_t>(start); o.put<int64_t>(ns); o.put<int64_t>(0);
}
// arcs grouped by source state, input order kept inside a state
void group_by_src(const int32_t *src, int64_t na, int64_t ns, std::vector<int64_t> *off, std::vector<int64_t> *order) {
  off->assign((size_t)ns + 1, 0);
  for (int64_t a = 0; a < na; a++) { if (src[a] < 0 || src[a] >= ns) throw std::runtime_error("arc source out of range"); (*off)[src[a] + 1]++; }
  for (int64_t s = 0; s < ns; s++) (*off)[s + 1] += (*off)[s];
  order->resize((size_t)na);
  std::vector<int64_t> fill(off->begin(), off->end() - 1);
  for (int64_t a = 0; a < na; a++) (*order)[fill[src[a]]++] = a;
}
FILE *open_out(const char *path, int32_t append) {
  FILE *f = fopen(path, append ? "ab" : "wb");
  if (!f) throw std::runtime_error(std::string("cannot open ") + path);
  return f;
}
}  // namespace

extern "C" {

int b2k_lat_write(const b2k_raw_lattice *in, const char *key, const char *path, int32_t binary, int32_t append) {
  if (!in || !key || !path || in->num_states < 0 || in->num_arcs < 0 || in->num_finals < 0) return b2k::set_error(B2K_ERR_INVALID, "b2k_lat_write: bad args");
  FILE *f = nullptr;
  try {
    std::vector<int64_t> off, order;
    group_by_src(in->arc_src, in->num_arcs, in->num_states, &off, &order);
    const float inf = std::numeric_limits<float>::infinity();
    std::vector<float> fin((size_t)in->num_states, inf);
    for (int64_t k = 0; k < in->num_finals; k++) { if (in->final_state[k] < 0 || in->final_state[k] >= in->num_states) throw std::runtime_error("final state out of range"); fin[in->final_state[k]] = in->final_cost[k]; }
    f = open_out(path, append);
    Out o{f};
    if (binary) {
      o.text(std::string(key) + " "); o.put<char>(0); o.put<char>('B');
      fst_header(o, "lattice4", in->num_states ? 0 : -1, in->num_states);
      for (int64_t s = 0; s < in->num_states; s++) {
        if (fin[s] != inf) { o.put<float>(fin[s]); o.put<float>(0.f); } else { o.put<float>(inf); o.put<float>(inf); }
        o.put<int64_t>(off[s + 1] - off[s]);
        for (int64_t k = off[s]; k < off[s + 1]; k++) {
          const int64_t a = order[k];
          o.put<int32_t>(in->arc_ilabel[a]); o.put<int32_t>(in->arc_olabel[a]); o.put<float>(in->arc_graph_cost[a]); o.put<float>(in->arc_acoustic_cost[a]); o.put<int32_t>(in->arc_dst[a]);
        }
      }
    } else {
      o.text(std::string(key) + "\n");
      for (int64_t s = 0; s < in->num_states; s++) {
        for (int64_t k = off[s]; k < off[s + 1]; k++) {
          const int64_t a = order[k];
          const float g = in->arc_graph_cost[a], ac = in->arc_acoustic_cost[a];
          std::string l = std::to_string(s) + "\t" + std::to_string(in->arc_dst[a]) + "\t" + std::to_string(in->arc_ilabel[a]) + "\t" + std::to_string(in->arc_olabel[a]);
          if (!(g == 0.f && ac == 0.f)) l += "\t" + num_text(g) + "," + num_text(ac);
          o.text(l + "\n");
        }
        if (fin[s] != inf) o.text(fin[s] == 0.f ? std::to_string(s) + "\n" : std::to_string(s) + "\t" + num_text(fin[s]) + ",0\n");
      }
      o.text("\n");
    }
    fclose(f);
  } catch (const std::exception &e) {
    if (f) fclose(f);
    return b2k::set_error(B2K_ERR_INVALID, "b2k_lat_write", e.what());
  }
  return B2K_OK;
}

int b2k_clat_write(const b2k_clat *c, const char *key, const char *path, int32_t binary, int32_t append) {
  if (!c || !key || !path) return b2k::set_error(B2K_ERR_INVALID, "b2k_clat_write: bad args");
  FILE *f = nullptr;
  try {
    const int64_t ns = c->num_states, na = (int64_t)c->arc_src.size(), nf = (int64_t)c->final_state.size();
    std::vector<int64_t> off, order;
    group_by_src(c->arc_src.data(), na, ns, &off, &order);
    std::vector<int64_t> fin((size_t)ns, -1);
    for (int64_t k = 0; k < nf; k++) fin[c->final_state[k]] = k;
    auto tids_of_arc = [&](int64_t a, const int32_t **p, int64_t *n) { *p = c->tids.data() + c->arc_str_off[a]; *n = c->arc_str_off[a + 1] - c->arc_str_off[a]; };
    auto tids_of_final = [&](int64_t k, const int32_t **p, int64_t *n) { *p = c->tids.data() + c->final_str_off[k]; *n = c->final_str_off[k + 1] - c->final_str_off[k]; };
    auto joined = [](const int32_t *p, int64_t n) { std::string s; for (int64_t i = 0; i < n; i++) { if (i) s += "_"; s += std::to_string(p[i]); } return s; };
    f = open_out(path, append);
    Out o{f};
    const float inf = std::numeric_limits<float>::infinity();
    if (binary) {
      o.text(std::string(key) + " "); o.put<char>(0); o.put<char>('B');
      fst_header(o, "compactlattice44", ns ? 0 : -1, ns);
      auto weight = [&](float g, float a, const int32_t *p, int64_t n) { o.put<float>(g); o.put<float>(a); o.put<int32_t>((int32_t)n); o.bytes(p, 4 * (size_t)n); };
      for (int64_t s = 0; s < ns; s++) {
        const int32_t *p; int64_t n;
        if (fin[s] >= 0) { tids_of_final(fin[s], &p, &n); weight(c->final_g[fin[s]], c->final_a[fin[s]], p, n); } else weight(inf, inf, nullptr, 0);
        o.put<int64_t>(off[s + 1] - off[s]);
        for (int64_t k = off[s]; k < off[s + 1]; k++) {
          const int64_t a = order[k];
          tids_of_arc(a, &p, &n);
          o.put<int32_t>(c->arc_word[a]); o.put<int32_t>(c->arc_word[a]); weight(c->arc_g[a], c->arc_a[a], p, n); o.put<int32_t>(c->arc_dst[a]);
        }
      }
    } else {
      o.text(std::string(key) + "\n");
      for (int64_t s = 0; s < ns; s++) {
        const int32_t *p; int64_t n;
        for (int64_t k = off[s]; k < off[s + 1]; k++) {
          const int64_t a = order[k];
          tids_of_arc(a, &p, &n);
          std::string l = std::to_string(s) + "\t" + std::to_string(c->arc_dst[a]) + "\t" + std::to_string(c->arc_word[a]);
          if (!(c->arc_g[a] == 0.f && c->arc_a[a] == 0.f && n == 0)) l += "\t" + num_text(c->arc_g[a]) + "," + num_text(c->arc_a[a]) + "," + joined(p, n);
          o.text(l + "\n");
        }
        if (fin[s] >= 0) {
          tids_of_final(fin[s], &p, &n);
          const float g = c->final_g[fin[s]], a = c->final_a[fin[s]];
          o.text((g == 0.f && a == 0.f && n == 0) ? std::to_string(s) + "\n" : std::to_string(s) + "\t" + num_text(g) + "," + num_text(a) + "," + joined(p, n) + "\n");
        }
      }
      o.text("\n");
    }
    fclose(f);
  } catch (const std::exception &e) {
    if (f) fclose(f);
    return b2k::set_error(B2K_ERR_INVALID, "b2k_clat_write", e.what());
  }
  return B2K_OK;
}

// CompactLatticeShortestPath + the read-out of online2-wav-nnet3-latgen-faster.cc:43-76: the cheapest path of the compact lattice
// (word ids, the concatenated transition-id strings, graph and acoustic cost with the final weight included).
int b2k_clat_best_path(const b2k_clat *c, int32_t *words, int32_t *n_words, int32_t *tids, int32_t *n_tids, int32_t cap_words, int32_t cap_tids,
                       float *graph_cost, float *acoustic_cost) {
  if (!c || !n_words || !n_tids || cap_words < 0 || cap_tids < 0 || (cap_words > 0 && !words) || (cap_tids > 0 && !tids))
    return b2k::set_error(B2K_ERR_INVALID, "b2k_clat_best_path: bad args");
  *n_words = 0; *n_tids = 0;
  if (graph_cost) *graph_cost = std::numeric_limits<float>::infinity();
  if (acoustic_cost) *acoustic_cost = std::numeric_limits<float>::infinity();
  const int64_t N = c->num_states, A = (int64_t)c->arc_src.size(), F = (int64_t)c->final_state.size();
  if (N == 0 || F == 0) return B2K_OK;
  std::vector<int64_t> off(N + 1, 0);
  for (int64_t a = 0; a < A; a++) off[c->arc_src[a] + 1]++;
  for (int64_t s = 0; s < N; s++) off[s + 1] += off[s];
  std::vector<int64_t> arcs(A), fill(off.begin(), off.end() - 1);
  for (int64_t a = 0; a < A; a++) arcs[fill[c->arc_src[a]]++] = a;
  std::vector<int32_t> indeg(N, 0), topo;
  for (int64_t a = 0; a < A; a++) indeg[c->arc_dst[a]]++;
  for (int64_t s = 0; s < N; s++) if (!indeg[s]) topo.push_back((int32_t)s);
  for (size_t i = 0; i < topo.size(); i++)
    for (int64_t k = off[topo[i]]; k < off[topo[i] + 1]; k++) if (--indeg[c->arc_dst[arcs[k]]] == 0) topo.push_back(c->arc_dst[arcs[k]]);
  const double INF = std::numeric_limits<double>::infinity();
  std::vector<double> dist(N, INF);
  std::vector<int64_t> back(N, -1);
  dist[0] = 0.0;
  for (int32_t s : topo) {
    if (!(dist[s] < INF)) continue;
    for (int64_t k = off[s]; k < off[s + 1]; k++) {
      const int64_t a = arcs[k];
      const double nd = dist[s] + (double)c->arc_g[a] + (double)c->arc_a[a];
      if (nd < dist[c->arc_dst[a]]) { dist[c->arc_dst[a]] = nd; back[c->arc_dst[a]] = a; }
    }
  }
  int64_t best = -1;
  double best_cost = INF;
  for (int64_t f = 0; f < F; f++) { const double v = dist[c->final_state[f]] + (double)c->final_g[f] + (double)c->final_a[f]; if (v < best_cost) { best_cost = v; best = f; } }
  if (best < 0) return b2k::set_error(B2K_ERR_STATE, "b2k_clat_best_path: no final state is reachable");
  std::vector<int64_t> path;
  for (int32_t s = c->final_state[best]; s != 0;) { path.push_back(back[s]); s = c->arc_src[back[s]]; }
  std::reverse(path.begin(), path.end());
  double g = (double)c->final_g[best], ac = (double)c->final_a[best];
  int64_t nw = 0, nt = 0;
  auto emit = [&](int64_t b, int64_t e) { for (int64_t i = b; i < e; i++) { if (nt < cap_tids) tids[nt] = c->tids[i]; nt++; } };
  for (int64_t a : path) {
    g += (double)c->arc_g[a]; ac += (double)c->arc_a[a];
    if (nw < cap_words) words[nw] = c->arc_word[a];
    nw++;
    emit(c->arc_str_off[a], c->arc_str_off[a + 1]);
  }
  emit(c->final_str_off[best], c->final_str_off[best + 1]);
  *n_words = (int32_t)nw; *n_tids = (int32_t)nt;
  if (graph_cost) *graph_cost = (float)g;
  if (acoustic_cost) *acoustic_cost = (float)ac;
  if (nw > cap_words || nt > cap_tids) return b2k::set_error(B2K_ERR_OVERFLOW, "b2k_clat_best_path: output buffers too small (sizes returned)");
  return B2K_OK;
}

float b2k_clat_effective_beam(const b2k_clat *c) { return c ? c->effective_beam : 0.f; }

int b2k_clat_sizes(const b2k_clat *c, int64_t sizes[6]) {
  if (!c || !sizes) return b2k::set_error(B2K_ERR_INVALID, "b2k_clat_sizes: bad args");
  sizes[0] = c->num_states; sizes[1] = (int64_t)c->arc_src.size(); sizes[2] = (int64_t)c->final_state.size();
  sizes[3] = (int64_t)c->tids.size(); sizes[4] = c->subsets_expanded; sizes[5] = c->elements_total;
  return B2K_OK;
}

int b2k_clat_copy(const b2k_clat *c, b2k_compact_lattice *out) {
  if (!c || !out) return b2k::set_error(B2K_ERR_INVALID, "b2k_clat_copy: bad args");
  const size_t na = c->arc_src.size(), nf = c->final_state.size();
  out->num_states = c->num_states; out->num_arcs = (int64_t)na; out->num_finals = (int64_t)nf; out->num_tids = (int64_t)c->tids.size();
  auto cp = [](void *dst, const void *src, size_t bytes) { if (dst && bytes) memcpy(dst, src, bytes); };
  cp(out->arc_src, c->arc_src.data(), 4 * na); cp(out->arc_dst, c->arc_dst.data(), 4 * na); cp(out->arc_word, c->arc_word.data(), 4 * na);
  cp(out->arc_graph_cost, c->arc_g.data(), 4 * na); cp(out->arc_acoustic_cost, c->arc_a.data(), 4 * na);
  cp(out->arc_tids_off, c->arc_str_off.data(), 8 * (na + 1));
  cp(out->final_state, c->final_state.data(), 4 * nf); cp(out->final_graph_cost, c->final_g.data(), 4 * nf);
  cp(out->final_acoustic_cost, c->final_a.data(), 4 * nf); cp(out->final_tids_off, c->final_str_off.data(), 8 * (nf + 1));
  cp(out->tids, c->tids.data(), 4 * c->tids.size());
  return B2K_OK;
}

}  // extern "C"
