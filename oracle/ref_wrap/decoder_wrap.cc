// oracle/_ref decoder wrapper -- TEST INFRASTRUCTURE ONLY.
//
// Drives the reference's OWN decoder (decoder/lattice-faster-decoder.{h,cc} and
// util/hash-list-inl.h, compiled from where they lie under /root/reference/src
// against the container-only OpenFst stand-in in oracle/ref_wrap/fst_stub/) on a
// CSR graph and a log-likelihood matrix, and exports what the parity tests
// compare: the token list of every frame in HashList order right after the frame
// was processed, and the finalized raw lattice in the row format of
// oracle/decoder_oracle.cc (ExportLattice).  This pins the decoder restatement
// (oracle/decoder_oracle.cc) and, through it, the CUDA decoder to the reference's
// search code; only the FST container is ours.
//
// A token's HCLG state is not stored by the reference (it is the HashList key of
// the frame being built), so the wrapper advances one frame at a time
// (AdvanceDecoding(decodable, 1), lattice-faster-decoder.cc:589) and records the
// (Token*, state) pairs of toks_ after every frame.
#include <cstring>
#include <unordered_map>
#include <vector>

#include "decoder/lattice-faster-decoder.h"

namespace {

using kaldi::BaseFloat;
using kaldi::int32;

class MatrixDecodable : public kaldi::DecodableInterface {
 public:
  MatrixDecodable(const float *ll, int32 T, int32 P, const int32 *tid2pdf, int32 ntid)
      : ll_(ll), T_(T), P_(P), tid2pdf_(tid2pdf), ntid_(ntid) {}
  BaseFloat LogLikelihood(int32 frame, int32 index) override { return ll_[(size_t)frame * P_ + tid2pdf_[index]]; }
  int32 NumFramesReady() const override { return T_; }
  bool IsLastFrame(int32 frame) const override { return frame == T_ - 1; }
  int32 NumIndices() const override { return ntid_; }
 private:
  const float *ll_; int32 T_, P_; const int32 *tid2pdf_; int32 ntid_;
};

typedef fst::ConstFst<fst::StdArc> Graph;
typedef kaldi::decoder::StdToken Token;
typedef kaldi::LatticeFasterDecoderTpl<Graph, Token> Base;

inline int32 Bits(float f) { int32 i; std::memcpy(&i, &f, 4); return i; }

class Probe : public Base {
 public:
  Probe(const Graph &g, const kaldi::LatticeFasterDecoderConfig &c) : Base(g, c) {}

  struct Frame { std::vector<int32> state; std::vector<float> cost; std::unordered_map<const void *, int32> state_of; };
  std::vector<Frame> frames;          // [T + 1], list (HashList) order
  std::vector<int32> lat_state_frame, lat_state_state, lat_arc, lat_final_state;
  std::vector<float> lat_state_tot, lat_state_extra, lat_final_cost;

  void Record() {
    frames.emplace_back();
    Frame &f = frames.back();
    for (const Elem *e = toks_.GetList(); e != NULL; e = e->tail) {
      f.state.push_back(e->key);
      f.cost.push_back(e->val->tot_cost);
      f.state_of[e->val] = e->key;
    }
  }

  // sizes of the raw lattice only (what GetRawLattice would emit), for the timing leg
  int64_t n_lat_states = 0, n_lat_arcs = 0;
  void CountLattice() {
    for (size_t f = 0; f < active_toks_.size(); f++)
      for (Token *tok = active_toks_[f].toks; tok != NULL; tok = tok->next) {
        n_lat_states++;
        for (ForwardLinkT *l = tok->links; l != NULL; l = l->next) n_lat_arcs++;
      }
  }

  int32 StateOf(int32 frame, const void *tok) const {
    auto it = frames[frame].state_of.find(tok);
    return it == frames[frame].state_of.end() ? -1 : it->second;
  }

  // the content of GetRawLattice (:114-197), keyed by (frame, HCLG state) instead of lattice state ids
  void Export() {
    const int32 num_frames = (int32)active_toks_.size() - 1;
    for (int32 f = 0; f <= num_frames; f++) {
      for (Token *tok = active_toks_[f].toks; tok != NULL; tok = tok->next) {
        const int32 s = StateOf(f, tok);
        lat_state_frame.push_back(f); lat_state_state.push_back(s);
        lat_state_tot.push_back(tok->tot_cost); lat_state_extra.push_back(tok->extra_cost);
        for (ForwardLinkT *l = tok->links; l != NULL; l = l->next) {
          BaseFloat cost_offset = 0.0;
          if (l->ilabel != 0) cost_offset = cost_offsets_[f];
          const int32 nf = f + (l->ilabel != 0 ? 1 : 0);
          const BaseFloat ac = l->acoustic_cost - cost_offset;
          const int32 rec[8] = {f, s, nf, StateOf(nf, l->next_tok), l->ilabel, l->olabel, Bits(l->graph_cost), Bits(ac)};
          lat_arc.insert(lat_arc.end(), rec, rec + 8);
        }
        if (f == num_frames) {
          if (!final_costs_.empty()) {
            auto it = final_costs_.find(tok);
            if (it != final_costs_.end()) { lat_final_state.push_back(s); lat_final_cost.push_back(it->second); }
          } else {
            lat_final_state.push_back(s); lat_final_cost.push_back(0.0f);
          }
        }
      }
    }
  }
};

struct Handle {
  Graph graph;
  std::vector<int32> tid2pdf;
  kaldi::LatticeFasterDecoderConfig cfg;
  Probe *dec = nullptr;
  ~Handle() { delete dec; }
};

}  // namespace

extern "C" {

struct b2k_refdec_cfg {
  float beam; int32_t max_active; int32_t min_active; float lattice_beam;
  int32_t prune_interval; float beam_delta; float hash_ratio; float prune_scale;
};

void *b2k_refdec_create(int32_t num_states, int32_t start, const int32_t *offsets, const int32_t *ilabel,
                        const int32_t *olabel, const float *weight, const int32_t *nextstate,
                        const float *final_cost, const int32_t *tid2pdf, int32_t num_tids,
                        const b2k_refdec_cfg *cfg) {
  Handle *h = new Handle();
  for (int32 s = 0; s < num_states; s++) {
    h->graph.AddState();
    h->graph.SetFinal(s, fst::TropicalWeight(final_cost[s]));
    for (int32 a = offsets[s]; a < offsets[s + 1]; a++)
      h->graph.AddArc(s, fst::StdArc(ilabel[a], olabel[a], fst::TropicalWeight(weight[a]), nextstate[a]));
  }
  h->graph.SetStart(start);
  h->tid2pdf.assign(tid2pdf, tid2pdf + num_tids);
  h->cfg.beam = cfg->beam; h->cfg.max_active = cfg->max_active; h->cfg.min_active = cfg->min_active;
  h->cfg.lattice_beam = cfg->lattice_beam; h->cfg.prune_interval = cfg->prune_interval;
  h->cfg.beam_delta = cfg->beam_delta; h->cfg.hash_ratio = cfg->hash_ratio; h->cfg.prune_scale = cfg->prune_scale;
  return h;
}

void b2k_refdec_destroy(void *hp) { delete (Handle *)hp; }

// InitDecoding + AdvanceDecoding + FinalizeDecoding (lattice-faster-decoder.cc:62,589,634).
// record != 0: advance frame by frame and keep every frame's token list (parity tests);
// record == 0: one AdvanceDecoding call, as online2-wav-nnet3-latgen-faster drives it (timing).
int b2k_refdec_decode(void *hp, const float *loglikes, int32_t T, int32_t num_pdfs, int32_t record) {
  Handle *h = (Handle *)hp;
  delete h->dec;
  h->dec = new Probe(h->graph, h->cfg);
  MatrixDecodable decodable(loglikes, T, num_pdfs, h->tid2pdf.data(), (int32)h->tid2pdf.size());
  try {
    h->dec->InitDecoding();
    if (record) {
      h->dec->Record();
      for (int32 t = 0; t < T; t++) {
        h->dec->AdvanceDecoding(&decodable, 1);
        h->dec->Record();
      }
    } else {
      h->dec->AdvanceDecoding(&decodable);
    }
    h->dec->FinalizeDecoding();
    if (record) h->dec->Export();
    else h->dec->CountLattice();
  } catch (const std::exception &e) {
    return 1;
  }
  return 0;
}

void b2k_refdec_frame_size(void *hp, int32_t frame_plus_one, int64_t *ntok) {
  *ntok = (int64_t)((Handle *)hp)->dec->frames[frame_plus_one].state.size();
}
void b2k_refdec_frame_copy(void *hp, int32_t frame_plus_one, int32_t *state, float *cost) {
  const Probe::Frame &f = ((Handle *)hp)->dec->frames[frame_plus_one];
  std::memcpy(state, f.state.data(), f.state.size() * 4);
  std::memcpy(cost, f.cost.data(), f.cost.size() * 4);
}
void b2k_refdec_lattice_sizes(void *hp, int64_t *ns, int64_t *na, int64_t *nf) {
  Probe *d = ((Handle *)hp)->dec;
  *ns = (int64_t)d->lat_state_frame.size(); *na = (int64_t)d->lat_arc.size() / 8; *nf = (int64_t)d->lat_final_state.size();
  if (d->frames.empty()) { *ns = d->n_lat_states; *na = d->n_lat_arcs; }   // decoded with record == 0
}
void b2k_refdec_lattice(void *hp, int32_t *state_frame, int32_t *state_state, float *state_tot, float *state_extra,
                        int32_t *arcs8, int32_t *final_state, float *final_cost) {
  Probe *d = ((Handle *)hp)->dec;
  std::memcpy(state_frame, d->lat_state_frame.data(), d->lat_state_frame.size() * 4);
  std::memcpy(state_state, d->lat_state_state.data(), d->lat_state_state.size() * 4);
  std::memcpy(state_tot, d->lat_state_tot.data(), d->lat_state_tot.size() * 4);
  std::memcpy(state_extra, d->lat_state_extra.data(), d->lat_state_extra.size() * 4);
  std::memcpy(arcs8, d->lat_arc.data(), d->lat_arc.size() * 4);
  std::memcpy(final_state, d->lat_final_state.data(), d->lat_final_state.size() * 4);
  std::memcpy(final_cost, d->lat_final_cost.data(), d->lat_final_cost.size() * 4);
}

}  // extern "C"
