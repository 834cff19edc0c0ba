// b2k_pipeline_shim.h — the device side of the batcher (b2k_batcher.h): a Backend over the C ABI that decodes batches of
// equal-length utterances with b2k_pipeline_* and turns every finalized raw lattice into a compact lattice with
// b2k_lat_determinize_pruned — what BatchedThreadedNnet3CudaOnlinePipeline does between DecodeBatch and the lattice
// callback (cudadecoder/batched-threaded-nnet3-cuda-online-pipeline.cc:316-377,727-790).  Plain C++17 over include/b2k.h;
// the Kaldi-typed wrapper is BatchedOnlinePipelineB2k in b2k_kaldi_shims.h.
//
// One b2k_pipeline per distinct utterance length (created on first use and kept: the nnet3 program, the decoder arenas
// and, with an extractor, the i-vector workspace are sized for a length).  `max_pipelines` bounds how many are kept
// (least recently used is destroyed first).  STATE: type-checked and link-checked on the CPU (tests/test_batcher_cpp.py);
// like pipeline.cu it has not run on a device yet.
#ifndef B2K_PIPELINE_SHIM_H_
#define B2K_PIPELINE_SHIM_H_

#include <list>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "b2k.h"
#include "b2k_batcher.h"

namespace b2k_host {

inline void Check(int rc, const char *what) {
  if (rc != B2K_OK) throw std::runtime_error(std::string(what) + ": " + b2k_last_error());
}

struct ClatDeleter { void operator()(b2k_clat *c) const { b2k_clat_destroy(c); } };

class B2kPipelineBackend {
 public:
  struct Result {
    std::unique_ptr<b2k_clat, ClatDeleter> clat;   // b2k_clat_sizes / b2k_clat_copy read it out
    float effective_beam = 0.f;
    int64_t raw_states = 0, raw_arcs = 0;
  };

  // ivec_files / ivec_opts may be null (the network then sees zero i-vectors); nothing passed in is owned
  B2kPipelineBackend(const b2k_pipeline_cfg &cfg, const b2k_model *model, const b2k_fst *fst, const b2k_ivec_files *ivec_files,
                     const b2k_ivec_cfg *ivec_opts, float det_beam, int64_t det_max_states = 0, size_t max_pipelines = 4)
      : cfg_(cfg), model_(model), fst_(fst), ivec_files_(ivec_files), det_beam_(det_beam), det_max_states_(det_max_states),
        max_pipelines_(max_pipelines ? max_pipelines : 1) {
    if (!model || !fst || (ivec_files && !ivec_opts)) throw std::invalid_argument("B2kPipelineBackend: bad arguments");
    if (ivec_opts) ivec_opts_ = *ivec_opts;
  }
  ~B2kPipelineBackend() { for (auto &e : pipes_) Destroy(e); }
  B2kPipelineBackend(const B2kPipelineBackend &) = delete;
  B2kPipelineBackend &operator=(const B2kPipelineBackend &) = delete;

  void Decode(int64_t num_samples, const std::vector<const float *> &waves, std::vector<Result> *out) {
    const int32_t n = (int32_t)waves.size();
    Entry &e = Get(num_samples);
    Check(b2k_pipeline_decode_batch(e.pipe, n, waves.data(), nullptr), "b2k_pipeline_decode_batch");
    // sizes, then the packed lattices of the batch
    std::vector<int64_t> so(n + 1), ao(n + 1), fo(n + 1);
    b2k_raw_lattice all = {};
    Check(b2k_pipeline_get_raw_lattices(e.pipe, n, &all, so.data(), ao.data(), fo.data(), nullptr), "b2k_pipeline_get_raw_lattices(sizes)");
    std::vector<int32_t> sf(all.num_states), sh(all.num_states), as(all.num_arcs), ad(all.num_arcs), ai(all.num_arcs), ao_(all.num_arcs), fs(all.num_finals);
    std::vector<float> st(all.num_states), se(all.num_states), ag(all.num_arcs), aa(all.num_arcs), fc(all.num_finals);
    all.state_frame = sf.data(); all.state_hclg = sh.data(); all.state_tot_cost = st.data(); all.state_extra_cost = se.data();
    all.arc_src = as.data(); all.arc_dst = ad.data(); all.arc_ilabel = ai.data(); all.arc_olabel = ao_.data();
    all.arc_graph_cost = ag.data(); all.arc_acoustic_cost = aa.data(); all.final_state = fs.data(); all.final_cost = fc.data();
    Check(b2k_pipeline_get_raw_lattices(e.pipe, n, &all, so.data(), ao.data(), fo.data(), nullptr), "b2k_pipeline_get_raw_lattices");
    std::vector<b2k_raw_lattice> ones((size_t)n);
    for (int32_t i = 0; i < n; i++) {
      b2k_raw_lattice &one = ones[i];                         // views into the packed arrays; ids are lattice-relative
      one = b2k_raw_lattice{};
      one.num_states = so[i + 1] - so[i]; one.num_arcs = ao[i + 1] - ao[i]; one.num_finals = fo[i + 1] - fo[i];
      one.state_frame = sf.data() + so[i]; one.state_hclg = sh.data() + so[i];
      one.state_tot_cost = st.data() + so[i]; one.state_extra_cost = se.data() + so[i];
      one.arc_src = as.data() + ao[i]; one.arc_dst = ad.data() + ao[i]; one.arc_ilabel = ai.data() + ao[i]; one.arc_olabel = ao_.data() + ao[i];
      one.arc_graph_cost = ag.data() + ao[i]; one.arc_acoustic_cost = aa.data() + ao[i];
      one.final_state = fs.data() + fo[i]; one.final_cost = fc.data() + fo[i];
    }
    std::vector<b2k_clat *> clats((size_t)n, nullptr);       // determinized on the host cores, one lattice per task
    const int rc = b2k_lat_determinize_pruned_batch(ones.data(), n, det_beam_, det_max_states_, det_threads_, clats.data(), nullptr);
    if (rc != B2K_OK) {
      const std::string msg = b2k_last_error();
      for (b2k_clat *c : clats) if (c) b2k_clat_destroy(c);
      throw std::runtime_error("b2k_lat_determinize_pruned_batch: " + msg);
    }
    for (int32_t i = 0; i < n; i++) {
      Result r;
      r.clat.reset(clats[i]);
      r.effective_beam = b2k_clat_effective_beam(clats[i]);
      r.raw_states = ones[i].num_states; r.raw_arcs = ones[i].num_arcs;
      out->push_back(std::move(r));
    }
  }

  void SetDeterminizeThreads(int32_t n) { det_threads_ = n; }   // 0 = all host cores

  size_t NumPipelines() const { return pipes_.size(); }

 private:
  struct Entry { int64_t num_samples; b2k_pipeline *pipe; b2k_ivec *ivec; };

  static void Destroy(Entry &e) {
    if (e.pipe) b2k_pipeline_destroy(e.pipe);
    if (e.ivec) b2k_ivec_destroy(e.ivec);
    e.pipe = nullptr; e.ivec = nullptr;
  }

  Entry &Get(int64_t num_samples) {
    for (auto it = pipes_.begin(); it != pipes_.end(); ++it)
      if (it->num_samples == num_samples) { pipes_.splice(pipes_.begin(), pipes_, it); return pipes_.front(); }
    b2k_pipeline_cfg c = cfg_;
    c.num_samples = num_samples;
    b2k_pipeline_plan plan;
    Check(b2k_pipeline_plan_for(&c, model_, &plan), "b2k_pipeline_plan_for");
    Entry e = {num_samples, nullptr, nullptr};
    if (ivec_files_) {
      b2k_ivec_cfg ic = ivec_opts_;
      ic.max_lanes = c.max_batch; ic.max_frames = plan.num_feature_frames;
      Check(b2k_ivec_create_from_files(&ic, ivec_files_, &e.ivec), "b2k_ivec_create_from_files");
    }
    const int rc = b2k_pipeline_create(&c, model_, fst_, e.ivec, &e.pipe);
    if (rc != B2K_OK) { const std::string msg = b2k_last_error(); Destroy(e); throw std::runtime_error("b2k_pipeline_create: " + msg); }
    while (pipes_.size() >= max_pipelines_) { Destroy(pipes_.back()); pipes_.pop_back(); }
    pipes_.push_front(e);
    return pipes_.front();
  }

  b2k_pipeline_cfg cfg_;
  const b2k_model *model_;
  const b2k_fst *fst_;
  const b2k_ivec_files *ivec_files_;
  b2k_ivec_cfg ivec_opts_ = {};
  float det_beam_;
  int64_t det_max_states_;
  size_t max_pipelines_;
  int32_t det_threads_ = 0;
  std::list<Entry> pipes_;
};

typedef UtteranceBatcher<B2kPipelineBackend> B2kBatcher;

}  // namespace b2k_host
#endif  // B2K_PIPELINE_SHIM_H_
