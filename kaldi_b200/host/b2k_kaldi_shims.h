// b2k_kaldi_shims.h — header-only C++ shims that keep the reference's class
// surfaces and forward to the b2k C ABI (include/b2k.h).  This is the binding a
// Kaldi maintainer adds: the tools keep calling
//   OnlineFeatureInterface::{Dim,NumFramesReady,IsLastFrame,GetFrame}   itf/online-feature-itf.h:49-110
//   OnlineBaseFeature::{AcceptWaveform,InputFinished}                   itf/online-feature-itf.h:112-125
//   cuda_decoder::CudaFst / CudaDecoder                                 cudadecoder/cuda-fst.h:75, cuda-decoder.h:224-346
//   cuda_decoder::BatchedThreadedNnet3CudaOnlinePipeline                cudadecoder/batched-threaded-nnet3-cuda-online-pipeline.h:127-275
// and the work happens in libb2k.so.  Compiled against the reference headers
// by `oracle/check_shims.py` (syntax + type check; OpenFst-typed members are
// guarded by B2K_HAVE_OPENFST because OpenFst is not in this image).
#ifndef B2K_KALDI_SHIMS_H_
#define B2K_KALDI_SHIMS_H_

#include <cuda_runtime.h>

#include <algorithm>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include <functional>
#include <map>
#include <mutex>

#include "b2k.h"
#include "b2k_dynamic_batcher.h"
#include "b2k_pipeline_shim.h"
#include "base/kaldi-error.h"
#include "itf/online-feature-itf.h"
#include "itf/transition-information.h"
#ifdef B2K_HAVE_OPENFST
#include "fst/fstlib.h"
#include "lat/kaldi-lattice.h"
#endif
#include "matrix/kaldi-matrix.h"
#include "matrix/kaldi-vector.h"

namespace kaldi {
namespace b2k_shim {

inline void Check(int rc, const char *what) {
  if (rc != B2K_OK) KALDI_ERR << what << ": " << b2k_last_error();      // throws KaldiFatalError like the reference
}

// OnlineGenericBaseFeature<MfccComputer|FbankComputer> (feat/online-feature.h:72-160) on the GPU.
// One object per utterance, not thread-safe (same contract as the reference).
class OnlineBaseFeatureB2k : public OnlineBaseFeature {
 public:
  OnlineBaseFeatureB2k(b2k_feat *feat, BaseFloat frame_shift_s, int32 max_samples)
      : feat_(feat), shift_s_(frame_shift_s), cap_(max_samples) {
    dim_ = b2k_feat_dim(feat_);
    if (cudaMalloc(&d_wave_, sizeof(float) * cap_) != cudaSuccess) KALDI_ERR << "cudaMalloc failed";
    int32 max_frames = b2k_feat_num_frames(feat_, cap_, 1) + 1;
    if (cudaMalloc(&d_feats_, sizeof(float) * max_frames * dim_) != cudaSuccess) KALDI_ERR << "cudaMalloc failed";
    host_.Resize(max_frames, dim_);
  }
  ~OnlineBaseFeatureB2k() override { cudaFree(d_wave_); cudaFree(d_feats_); }

  int32 Dim() const override { return dim_; }
  int32 NumFramesReady() const override { return frames_ready_; }
  bool IsLastFrame(int32 frame) const override { return finished_ && frame == frames_ready_ - 1; }
  BaseFloat FrameShiftInSeconds() const override { return shift_s_; }

  void GetFrame(int32 frame, VectorBase<BaseFloat> *feat) override {
    KALDI_ASSERT(frame >= 0 && frame < frames_ready_ && feat->Dim() == dim_);
    feat->CopyFromVec(host_.Row(frame));
  }

  // feat/online-feature.cc:131-159
  void AcceptWaveform(BaseFloat sampling_rate, const VectorBase<BaseFloat> &wave) override {
    if (wave.Dim() == 0) return;
    if (finished_) KALDI_ERR << "AcceptWaveform called after InputFinished() was called.";
    if (sampling_rate != b2k_feat_samp_freq(feat_))          // no resampler here (allow-downsample / allow-upsample, :138-150)
      KALDI_ERR << "Sampling frequency mismatch, expected " << b2k_feat_samp_freq(feat_) << ", got " << sampling_rate;
    if (num_samples_ + wave.Dim() > cap_) KALDI_ERR << "utterance longer than the configured capacity";
    cudaMemcpy(d_wave_ + num_samples_, wave.Data(), sizeof(float) * wave.Dim(), cudaMemcpyHostToDevice);
    num_samples_ += wave.Dim();
    Compute(false);
  }
  void InputFinished() override { finished_ = true; Compute(true); }

 private:
  void Compute(bool flush) {   // ComputeFeatures, feat/online-feature.cc:162-204
    int32 n_new = b2k_feat_num_frames(feat_, num_samples_, flush ? 1 : 0);
    if (n_new <= frames_ready_) return;
    const float *w = d_wave_;
    float *o = d_feats_;
    int32 ns = num_samples_, first = frames_ready_, cnt = n_new - frames_ready_;
    Check(b2k_feat_compute_batched(feat_, 1, &w, &ns, &first, &cnt, &o, dim_, nullptr), "b2k_feat_compute_batched");
    // Matrix rows are padded (kDefaultStride: Stride() >= NumCols(), e.g. 16 floats for a 13-dim MFCC): a pitched copy
    if (cudaMemcpy2D(host_.RowData(first), sizeof(float) * host_.Stride(), d_feats_ + (size_t)first * dim_, sizeof(float) * dim_,
                     sizeof(float) * dim_, cnt, cudaMemcpyDeviceToHost) != cudaSuccess)
      KALDI_ERR << "copying the new feature frames to the host failed";
    frames_ready_ = n_new;
  }
  b2k_feat *feat_;
  BaseFloat shift_s_;
  int32 cap_, dim_ = 0, num_samples_ = 0, frames_ready_ = 0;
  bool finished_ = false;
  float *d_wave_ = nullptr, *d_feats_ = nullptr;
  Matrix<BaseFloat> host_;
};

#ifdef B2K_HAVE_OPENFST
// cuda_decoder::CudaFst (cudadecoder/cuda-fst.h:62-82): CSR of the decoding graph in arc (file) order, with the
// transition-id -> pdf table pre-applied on the device as CudaFst::ApplyTransitionModelOnIlabels does (cuda-fst.cc:34-198).
class CudaFstB2k {
 public:
  CudaFstB2k(const fst::StdFst &fst, const TransitionInformation *trans_model = nullptr) {
    std::vector<int32> offsets(1, 0), ilabel, olabel, nextstate;
    std::vector<float> weight, final_cost;
    for (fst::StateIterator<fst::StdFst> siter(fst); !siter.Done(); siter.Next()) {
      const fst::StdArc::StateId s = siter.Value();
      for (fst::ArcIterator<fst::StdFst> aiter(fst, s); !aiter.Done(); aiter.Next()) {
        const fst::StdArc &arc = aiter.Value();
        ilabel.push_back(arc.ilabel); olabel.push_back(arc.olabel);
        weight.push_back(arc.weight.Value()); nextstate.push_back(arc.nextstate);
      }
      offsets.push_back(static_cast<int32>(ilabel.size()));
      final_cost.push_back(fst.Final(s).Value());
    }
    b2k_fst_csr csr = {};
    csr.num_states = static_cast<int32>(final_cost.size());
    csr.start = fst.Start();
    csr.offsets = offsets.data(); csr.ilabel = ilabel.data(); csr.olabel = olabel.data();
    csr.weight = weight.data(); csr.nextstate = nextstate.data(); csr.final_cost = final_cost.data();
    if (trans_model) {   // TransitionIdToPdfArray(): index = transition-id (itf/transition-information.h:94)
      const std::vector<int32_t> &t2p = trans_model->TransitionIdToPdfArray();
      csr.tid2pdf = t2p.data();
      csr.num_tids = static_cast<int32>(t2p.size());
      Check(b2k_fst_create(&csr, &fst_), "b2k_fst_create");
    } else {
      Check(b2k_fst_create(&csr, &fst_), "b2k_fst_create");
    }
  }
  ~CudaFstB2k() { b2k_fst_destroy(fst_); }
  CudaFstB2k(const CudaFstB2k &) = delete;
  CudaFstB2k &operator=(const CudaFstB2k &) = delete;
  uint32_t NumStates() const { return static_cast<uint32_t>(b2k_fst_num_states(fst_)); }
  fst::StdArc::StateId Start() const { return b2k_fst_start(fst_); }
  const b2k_fst *Handle() const { return fst_; }
 private:
  b2k_fst *fst_ = nullptr;
};
#endif

// cuda_decoder::CudaDecoderConfig (cudadecoder/cuda-decoder.h:58-163): the same fields, option names, Check() and
// ComputeConfig(), so that a pipeline config that embeds it keeps registering "--beam --lattice-beam --max-active
// --ntokens-pre-allocated --main-q-capacity --aux-q-capacity" (+ the endpointing group when the caller registers its
// own OnlineEndpointConfig beside it).  ToB2k() maps it onto the b2k decoder: this decoder keeps the CPU decoder's
// one-token-per-state semantics, so the queue capacities size its per-frame hash instead of arc-instantiated queues.
struct CudaDecoderConfigB2k {
  BaseFloat default_beam = 15.0f, lattice_beam = 10.0f;
  int32 ntokens_pre_allocated = 1000000, main_q_capacity = -1, aux_q_capacity = -1, max_active = 10000;
  template <class Opts>                                      // OptionsItf (itf/options-itf.h) or anything with its Register()
  void Register(Opts *opts) {
    opts->Register("beam", &default_beam, "Decoding beam. Larger->slower, more accurate.");
    opts->Register("lattice-beam", &lattice_beam, "The width of the lattice beam");
    opts->Register("max-active", &max_active, "At the end of each frame computation, we keep only its best max-active tokens.");
    opts->Register("ntokens-pre-allocated", &ntokens_pre_allocated, "Advanced - Number of tokens pre-allocated per channel (token arena).");
    opts->Register("main-q-capacity", &main_q_capacity, "Advanced - tokens that can be stored after pruning for each frame (-1 = 4*max-active).");
    opts->Register("aux-q-capacity", &aux_q_capacity, "Advanced - raw tokens that can be stored before pruning for each frame (-1 = 3*main-q-capacity).");
  }
  void Check() const { KALDI_ASSERT(default_beam > 0.0 && ntokens_pre_allocated >= 0 && lattice_beam >= 0.0f && max_active > 0); }
  void ComputeConfig() {                                     // KALDI_CUDA_DECODER_MAX_ACTIVE_MAIN_Q_CAPACITY_FACTOR 4, ..._AUX_Q_MAIN_Q_CAPACITIES_FACTOR 3
    if (main_q_capacity == -1) main_q_capacity = max_active * 4;
    if (aux_q_capacity == -1) aux_q_capacity = main_q_capacity * 3;
  }
  b2k_dec_cfg ToB2k(int32 max_frames) const {
    b2k_dec_cfg c;
    b2k_dec_cfg_default(&c);
    c.beam = default_beam; c.lattice_beam = lattice_beam; c.max_active = max_active;
    const int32 mq = main_q_capacity > 0 ? main_q_capacity : max_active * 4;
    c.max_tokens_per_frame = std::min<int32>(131072, std::max<int32>(1024, mq));
    c.max_frames = max_frames;
    c.max_tokens = std::max<int64_t>(ntokens_pre_allocated, (int64_t)max_frames * 9000);
    c.max_links = 2 * c.max_tokens;
    return c;
  }
};

// cuda_decoder::CudaDecoder surface (cudadecoder/cuda-decoder.h:171-346) over b2k_dec_*.
typedef int32 ChannelId;
class CudaDecoderB2k {
 public:
  CudaDecoderB2k(const b2k_fst *fst, const b2k_dec_cfg &config, int32 nlanes, int32 nchannels) {
    Check(b2k_dec_create(fst, &config, nlanes, nchannels, &dec_), "b2k_dec_create");
  }
  // CudaDecoder(const CudaFst &fst, const CudaDecoderConfig &config, int32 nlanes, int32 nchannels)  cuda-decoder.h:224
  CudaDecoderB2k(const b2k_fst *fst, const CudaDecoderConfigB2k &config, int32 nlanes, int32 nchannels, int32 max_frames = 4096) {
    config.Check();
    const b2k_dec_cfg c = config.ToB2k(max_frames);
    Check(b2k_dec_create(fst, &c, nlanes, nchannels, &dec_), "b2k_dec_create");
  }
  ~CudaDecoderB2k() { b2k_dec_destroy(dec_); }
  void InitDecoding(const std::vector<ChannelId> &channels) {
    Check(b2k_dec_init_decoding(dec_, channels.data(), (int32)channels.size(), nullptr), "InitDecoding");
  }
  // AdvanceDecoding(const std::vector<std::pair<ChannelId, const BaseFloat*>>&)  cuda-decoder.h:264-265
  void AdvanceDecoding(const std::vector<std::pair<ChannelId, const BaseFloat *>> &lanes_assignements) {
    std::vector<ChannelId> ch;
    std::vector<const float *> ll;
    for (const auto &p : lanes_assignements) { ch.push_back(p.first); ll.push_back(p.second); }
    Check(b2k_dec_advance_decoding(dec_, ch.data(), ll.data(), (int32)ch.size(), nullptr), "AdvanceDecoding");
  }
  int32 NumFramesDecoded(ChannelId ichannel) const {
    int32 n = 0;
    Check(b2k_dec_num_frames_decoded(dec_, ichannel, &n), "NumFramesDecoded");
    return n;
  }
#ifdef B2K_HAVE_OPENFST
  // GetRawLattice(const std::vector<ChannelId>&, std::vector<Lattice*>&, bool)  cuda-decoder.h
  void GetRawLattice(const std::vector<ChannelId> &channels, std::vector<Lattice *> &fst_out_vec, bool use_final_probs) {
    KALDI_ASSERT(use_final_probs);
    Check(b2k_dec_finalize_decoding(dec_, channels.data(), (int32)channels.size(), nullptr), "FinalizeDecoding");
    for (size_t i = 0; i < channels.size(); i++) {
      b2k_raw_lattice r = {};
      Check(b2k_dec_get_raw_lattice(dec_, channels[i], &r, nullptr), "GetRawLattice(size)");
      std::vector<int32> sf(r.num_states), sh(r.num_states), as(r.num_arcs), ad(r.num_arcs), ai(r.num_arcs), ao(r.num_arcs), fs(r.num_finals);
      std::vector<float> st(r.num_states), se(r.num_states), ag(r.num_arcs), aa(r.num_arcs), fc(r.num_finals);
      r.state_frame = sf.data(); r.state_hclg = sh.data(); r.state_tot_cost = st.data(); r.state_extra_cost = se.data();
      r.arc_src = as.data(); r.arc_dst = ad.data(); r.arc_ilabel = ai.data(); r.arc_olabel = ao.data();
      r.arc_graph_cost = ag.data(); r.arc_acoustic_cost = aa.data(); r.final_state = fs.data(); r.final_cost = fc.data();
      Check(b2k_dec_get_raw_lattice(dec_, channels[i], &r, nullptr), "GetRawLattice");
      Lattice *ofst = fst_out_vec[i];
      ofst->DeleteStates();
      for (int64 s = 0; s < r.num_states; s++) ofst->AddState();
      ofst->SetStart(0);
      for (int64 a = 0; a < r.num_arcs; a++)
        ofst->AddArc(as[a], LatticeArc(ai[a], ao[a], LatticeWeight(ag[a], aa[a]), ad[a]));     // lattice-faster-decoder.cc:179-182
      for (int64 f = 0; f < r.num_finals; f++) ofst->SetFinal(fs[f], LatticeWeight(fc[f], 0));  // :189
    }
  }
  // GetBestPath(const std::vector<ChannelId>&, std::vector<Lattice*>&, bool)  cuda-decoder.h; CPU semantics
  // LatticeFasterDecoderTpl::GetBestPath = GetRawLattice + ShortestPath (lattice-faster-decoder.cc:102-108): a linear
  // lattice, one state per path position, the final weight on the last state
  void GetBestPath(const std::vector<ChannelId> &channels, std::vector<Lattice *> &fst_out_vec, bool use_final_probs) {
    KALDI_ASSERT(use_final_probs);
    Check(b2k_dec_finalize_decoding(dec_, channels.data(), (int32)channels.size(), nullptr), "FinalizeDecoding");
    for (size_t i = 0; i < channels.size(); i++) {
      b2k_raw_lattice r = {};
      Check(b2k_dec_get_raw_lattice(dec_, channels[i], &r, nullptr), "GetRawLattice(size)");
      std::vector<int32> sf(r.num_states), sh(r.num_states), as(r.num_arcs), ad(r.num_arcs), ai(r.num_arcs), ao(r.num_arcs), fs(r.num_finals);
      std::vector<float> st(r.num_states), se(r.num_states), ag(r.num_arcs), aa(r.num_arcs), fc(r.num_finals);
      r.state_frame = sf.data(); r.state_hclg = sh.data(); r.state_tot_cost = st.data(); r.state_extra_cost = se.data();
      r.arc_src = as.data(); r.arc_dst = ad.data(); r.arc_ilabel = ai.data(); r.arc_olabel = ao.data();
      r.arc_graph_cost = ag.data(); r.arc_acoustic_cost = aa.data(); r.final_state = fs.data(); r.final_cost = fc.data();
      Check(b2k_dec_get_raw_lattice(dec_, channels[i], &r, nullptr), "GetRawLattice");
      std::vector<int64_t> path((size_t)r.num_arcs + 1);
      int64_t n = 0, fin = -1;
      Check(b2k_lat_best_path_arcs(&r, path.data(), &n, (int64_t)path.size(), &fin), "b2k_lat_best_path_arcs");
      Lattice *ofst = fst_out_vec[i];
      ofst->DeleteStates();
      if (fin < 0) continue;                                   // empty lattice: empty FST, as ShortestPath gives
      for (int64_t k = 0; k <= n; k++) ofst->AddState();
      ofst->SetStart(0);
      for (int64_t k = 0; k < n; k++) {
        const int64_t a = path[k];
        ofst->AddArc((int32)k, LatticeArc(ai[a], ao[a], LatticeWeight(ag[a], aa[a]), (int32)(k + 1)));
      }
      ofst->SetFinal((int32)n, LatticeWeight(fc[fin], 0));
    }
  }
#endif
 private:
  b2k_dec *dec_ = nullptr;
};

// cuda_decoder::BatchedThreadedNnet3CudaOnlinePipeline surface (cudadecoder/batched-threaded-nnet3-cuda-online-
// pipeline.h:127-275) over b2k_host::UtteranceBatcher + B2kPipelineBackend (b2k_batcher.h, b2k_pipeline_shim.h):
// TryInitCorrID / SetLatticeCallback / DecodeBatch(corr_ids, wave_samples, is_first_chunk, is_last_chunk) /
// WaitForLatticeCallbacks keep their names and argument meaning.  Results are those of the CPU tool
// online2-wav-nnet3-latgen-faster for each utterance; what differs from the reference GPU pipeline (no partial
// hypotheses, callbacks from the calling thread, utterances decoded when complete) is listed in b2k_batcher.h.
class BatchedOnlinePipelineB2k {
 public:
  using CorrelationID = uint64_t;
  // the compact lattice in ABI form; with OpenFst present use SetLatticeCallback(corr_id, LatticeCallback) below
  typedef std::function<void(CorrelationID, const b2k_clat *)> ClatCallback;

  BatchedOnlinePipelineB2k(const b2k_pipeline_cfg &config, const b2k_model *model, const b2k_fst *decode_fst,
                           const b2k_ivec_files *ivector_files, const b2k_ivec_cfg *ivector_opts, BaseFloat lattice_beam,
                           int32 num_channels)
      : backend_(config, model, decode_fst, ivector_files, ivector_opts, lattice_beam),
        batcher_(&backend_, config.max_batch, num_channels) {}

  bool TryInitCorrID(CorrelationID corr_id, int /*wait_for*/ = 0) { return batcher_.TryInitCorrId(corr_id); }

  void SetClatCallback(CorrelationID corr_id, ClatCallback cb) {
    batcher_.SetCallback(corr_id, [cb](CorrelationID id, b2k_host::B2kPipelineBackend::Result &r) { cb(id, r.clat.get()); });
  }

#if defined(B2K_HAVE_OPENFST) && !defined(B2K_OPENFST_IS_STANDIN)
  typedef std::function<void(CompactLattice &)> LatticeCallback;      // …online-pipeline.h:131
  void SetLatticeCallback(CorrelationID corr_id, const LatticeCallback &callback) {
    batcher_.SetCallback(corr_id, [callback](CorrelationID, b2k_host::B2kPipelineBackend::Result &r) {
      CompactLattice clat;
      FillCompactLattice(r.clat.get(), &clat);
      callback(clat);
    });
  }
  // b2k_clat -> CompactLattice: arcs CompactLatticeArc(word, word, CompactLatticeWeight(LatticeWeight(g, a), tids), dst)
  static void FillCompactLattice(const b2k_clat *c, CompactLattice *out) {
    int64_t sz[6];
    Check(b2k_clat_sizes(c, sz), "b2k_clat_sizes");
    std::vector<int32> as(sz[1]), ad(sz[1]), aw(sz[1]), fs(sz[2]), tids(sz[3]);
    std::vector<float> ag(sz[1]), aa(sz[1]), fg(sz[2]), fa(sz[2]);
    std::vector<int64_t> ao(sz[1] + 1), fo(sz[2] + 1);
    b2k_compact_lattice v = {};
    v.arc_src = as.data(); v.arc_dst = ad.data(); v.arc_word = aw.data(); v.arc_graph_cost = ag.data(); v.arc_acoustic_cost = aa.data();
    v.arc_tids_off = ao.data(); v.final_state = fs.data(); v.final_graph_cost = fg.data(); v.final_acoustic_cost = fa.data();
    v.final_tids_off = fo.data(); v.tids = tids.data();
    Check(b2k_clat_copy(c, &v), "b2k_clat_copy");
    out->DeleteStates();
    for (int64_t s = 0; s < sz[0]; s++) out->AddState();
    if (sz[0] > 0) out->SetStart(0);
    for (int64_t a = 0; a < sz[1]; a++) {
      std::vector<int32> str(tids.begin() + ao[a], tids.begin() + ao[a + 1]);
      out->AddArc(as[a], CompactLatticeArc(aw[a], aw[a], CompactLatticeWeight(LatticeWeight(ag[a], aa[a]), str), ad[a]));
    }
    for (int64_t f = 0; f < sz[2]; f++) {
      std::vector<int32> str(tids.begin() + fo[f], tids.begin() + fo[f + 1]);
      out->SetFinal(fs[f], CompactLatticeWeight(LatticeWeight(fg[f], fa[f]), str));
    }
  }
#endif

  // …online-pipeline.h:209-215.  partial_hypotheses / end_point are not offered (see b2k_batcher.h)
  void DecodeBatch(const std::vector<CorrelationID> &corr_ids, const std::vector<SubVector<BaseFloat>> &wave_samples,
                   const std::vector<bool> &is_first_chunk, const std::vector<bool> &is_last_chunk) {
    KALDI_ASSERT(corr_ids.size() == wave_samples.size());
    std::vector<std::pair<const float *, int64_t>> chunks;
    for (const SubVector<BaseFloat> &w : wave_samples) chunks.push_back({w.Data(), (int64_t)w.Dim()});
    try {
      batcher_.AcceptChunks(corr_ids, chunks, is_first_chunk, is_last_chunk);
    } catch (const std::exception &e) {
      KALDI_ERR << "DecodeBatch: " << e.what();
    }
  }

  void WaitForLatticeCallbacks() {
    try {
      batcher_.Flush();
    } catch (const std::exception &e) {
      KALDI_ERR << "WaitForLatticeCallbacks: " << e.what();
    }
  }

 private:
  b2k_host::B2kPipelineBackend backend_;
  b2k_host::B2kBatcher batcher_;
};

// cuda_decoder::BatchedThreadedNnet3CudaOnlinePipeline, STREAMING: every DecodeBatch advances every listed utterance by its
// chunk (features of the new frames, chunked network, decoder frames) over b2k_stream_*, reports partial hypotheses and
// end points per call and hands the raw lattice to the utterance's callback after its last chunk -- the reference's call
// structure (batched-threaded-nnet3-cuda-online-pipeline.h:160-215, .cc:316-377), where BatchedOnlinePipelineB2k above decodes
// an utterance when its last chunk has arrived.  Utterances hold a channel from TryInitCorrID until their last chunk.
class StreamingOnlinePipelineB2k {
 public:
  using CorrelationID = uint64_t;
  typedef std::function<void(const std::string &, bool, bool)> BestPathCallback;    // …online-pipeline.h:129: (text, partial, endpoint)
  typedef std::function<void(CorrelationID, const b2k_raw_lattice &)> RawLatticeCallback;

  // model: for the transition-id -> phone table the end point rules need (may be NULL when no end points are asked for)
  StreamingOnlinePipelineB2k(const b2k_stream_cfg &config, const b2k_model *model, const b2k_fst *decode_fst) : model_(model) {
    Check(b2k_stream_create(&config, model, decode_fst, &s_), "b2k_stream_create");
    int64_t info[8];
    Check(b2k_stream_info(s_, info), "b2k_stream_info");
    for (int32 c = (int32)info[0] - 1; c >= 0; c--) free_.push_back(c);
    frames_per_chunk_ = (int32)info[7];
    path_cap_ = (int32)info[2];          // arcs of a best path <= output frames + epsilon arcs: well under the feature frames of a stream
    samples_per_chunk_ = (int32)(frames_per_chunk_ * config.feat.samp_freq * 0.001f * config.feat.frame_shift_ms);
    int32_t mi[8];
    Check(b2k_model_info(model, mi), "b2k_model_info");
    decoder_frame_shift_seconds_ = 0.001f * config.feat.frame_shift_ms * mi[3];
    num_tids_ = mi[6];
    Check(b2k_endpoint_cfg_default(&endpoint_), "b2k_endpoint_cfg_default");
  }
  ~StreamingOnlinePipelineB2k() { b2k_stream_destroy(s_); }

  int32 GetNSampsPerChunk() const { return samples_per_chunk_; }                     // :248-250
  int32 GetNInputFramesPerChunk() const { return frames_per_chunk_; }
  BaseFloat GetDecoderFrameShiftSeconds() const { return decoder_frame_shift_seconds_; }

  // Like the reference's (…online-pipeline.cc:107-168: available_channels_m_, map_callbacks_m_), TryInitCorrID and the callback
  // setters may be called from other threads while ONE thread runs DecodeBatch (the dynamic batcher's worker).
  bool TryInitCorrID(CorrelationID corr_id, int /*wait_for*/ = 0) {                  // :165: false = every channel is taken
    std::lock_guard<std::mutex> lock(chan_mu_);
    if (chan_.count(corr_id)) return true;
    if (free_.empty()) return false;
    chan_[corr_id] = free_.back();
    free_.pop_back();
    return true;
  }
  void SetBestPathCallback(CorrelationID corr_id, const BestPathCallback &callback) {
    std::lock_guard<std::mutex> lock(cb_mu_);
    best_cb_[corr_id] = callback;
  }
  void SetRawLatticeCallback(CorrelationID corr_id, const RawLatticeCallback &callback) {
    std::lock_guard<std::mutex> lock(cb_mu_);
    lat_cb_[corr_id] = callback;
  }
  // how a word id is spelled in hypotheses (the reference reads its word symbol table); default: the id itself
  void SetWordMapper(const std::function<std::string(int32)> &f) { word_of_ = f; }
  void SetEndpointConfig(const b2k_endpoint_cfg &c) { endpoint_ = c; }

  // :209-215.  wave_samples hold 16-bit PCM values (what WaveData::Read leaves in its floats); anything else is refused.
  void DecodeBatch(const std::vector<CorrelationID> &corr_ids, const std::vector<SubVector<BaseFloat>> &wave_samples,
                   const std::vector<bool> &is_first_chunk, const std::vector<bool> &is_last_chunk,
                   std::vector<const std::string *> *partial_hypotheses = nullptr, std::vector<bool> *end_point = nullptr) {
    const size_t n = corr_ids.size();
    KALDI_ASSERT(wave_samples.size() == n && is_first_chunk.size() >= n && is_last_chunk.size() >= n);
    for (CorrelationID id : finished_) text_.erase(id);       // their hypothesis strings were valid until this call
    finished_.clear();
    pcm_.resize(n); chans_.resize(n); ptrs_.resize(n); ns_.resize(n); first_.resize(n); last_.resize(n);
    for (size_t i = 0; i < n; i++) {
      if (is_first_chunk[i] && !TryInitCorrID(corr_ids[i])) KALDI_ERR << "DecodeBatch: no free channel for a new utterance";
      {
        std::lock_guard<std::mutex> lock(chan_mu_);
        auto it = chan_.find(corr_ids[i]);
        if (it == chan_.end()) KALDI_ERR << "DecodeBatch: unknown correlation id (its first chunk never came)";
        chans_[i] = it->second;
      }
      const SubVector<BaseFloat> &w = wave_samples[i];
      pcm_[i].resize(w.Dim());
      for (int32 k = 0; k < w.Dim(); k++) {
        const float v = w(k);
        const int16_t q = (int16_t)v;
        if ((float)q != v) KALDI_ERR << "DecodeBatch: sample " << v << " is not a 16-bit PCM value";
        pcm_[i][k] = q;
      }
      ptrs_[i] = pcm_[i].data(); ns_[i] = w.Dim(); first_[i] = is_first_chunk[i] ? 1 : 0; last_[i] = is_last_chunk[i] ? 1 : 0;
    }
    Check(b2k_stream_decode_batch_i16(s_, (int32_t)n, chans_.data(), ptrs_.data(), ns_.data(), first_.data(), last_.data(), nullptr,
                                      nullptr, nullptr, nullptr, nullptr, cudaStreamPerThread), "b2k_stream_decode_batch_i16");
    b2k_dec *dec = b2k_stream_decoder(s_);
    // this batch's callbacks, taken under the lock and called without it (a callback may register callbacks)
    batch_best_.assign(n, BestPathCallback()); batch_lat_.assign(n, RawLatticeCallback());
    bool any_best = false;
    {
      std::lock_guard<std::mutex> lock(cb_mu_);
      for (size_t i = 0; i < n; i++) {
        auto b = best_cb_.find(corr_ids[i]);
        if (b != best_cb_.end()) { batch_best_[i] = b->second; any_best = true; }
        if (is_last_chunk[i]) {
          auto l = lat_cb_.find(corr_ids[i]);
          if (l != lat_cb_.end()) { batch_lat_[i] = l->second; lat_cb_.erase(l); }
          if (b != best_cb_.end()) best_cb_.erase(b);
        }
      }
    }
    const bool want_text = partial_hypotheses != nullptr || end_point != nullptr || any_best;
    if (partial_hypotheses) partial_hypotheses->assign(n, nullptr);
    if (end_point) end_point->assign(n, false);
    if (want_text) {
      const int32_t cap = path_cap_;
      il_.resize(n * (size_t)cap); ol_.resize(n * (size_t)cap); info_.resize(n);
      Check(b2k_dec_best_path(dec, chans_.data(), (int32_t)n, /*use_final_probs=*/0, cap, il_.data(), ol_.data(), nullptr, nullptr,
                              nullptr, nullptr, info_.data(), cudaStreamPerThread), "b2k_dec_best_path");
      for (size_t i = 0; i < n; i++) {
        std::string &text = text_[corr_ids[i]];
        text.clear();
        for (int32_t k = 0; k < info_[i].n_arcs; k++) {
          const int32 w = ol_[i * (size_t)cap + k];
          if (w == 0) continue;
          if (!text.empty()) text += ' ';
          text += word_of_ ? word_of_(w) : std::to_string(w);
        }
        bool ep = false;
        if ((end_point || any_best) && model_ && num_tids_ > 0) {
          int32_t hit = 0;
          Check(b2k_endpoint_detected_on_path(&endpoint_, b2k_model_tid2phone(model_), num_tids_, il_.data() + i * (size_t)cap, info_[i].n_arcs,
                                              info_[i].num_frames, decoder_frame_shift_seconds_, info_[i].final_relative_cost, &hit, nullptr),
                "b2k_endpoint_detected_on_path");
          ep = hit != 0;
        }
        if (partial_hypotheses) (*partial_hypotheses)[i] = &text;
        if (end_point) (*end_point)[i] = ep;
        if (batch_best_[i]) batch_best_[i](text, /*partial=*/!is_last_chunk[i], ep);
      }
    }
    for (size_t i = 0; i < n; i++) {
      if (!is_last_chunk[i]) continue;
      if (batch_lat_[i]) {
        b2k_raw_lattice q = {};
        Check(b2k_dec_get_raw_lattice(dec, chans_[i], &q, cudaStreamPerThread), "b2k_dec_get_raw_lattice");     // sizes
        st_f_.resize(q.num_states); st_h_.resize(q.num_states); st_t_.resize(q.num_states); st_e_.resize(q.num_states);
        a_s_.resize(q.num_arcs); a_d_.resize(q.num_arcs); a_i_.resize(q.num_arcs); a_o_.resize(q.num_arcs); a_g_.resize(q.num_arcs); a_a_.resize(q.num_arcs);
        f_s_.resize(q.num_finals); f_c_.resize(q.num_finals);
        q.state_frame = st_f_.data(); q.state_hclg = st_h_.data(); q.state_tot_cost = st_t_.data(); q.state_extra_cost = st_e_.data();
        q.arc_src = a_s_.data(); q.arc_dst = a_d_.data(); q.arc_ilabel = a_i_.data(); q.arc_olabel = a_o_.data();
        q.arc_graph_cost = a_g_.data(); q.arc_acoustic_cost = a_a_.data(); q.final_state = f_s_.data(); q.final_cost = f_c_.data();
        Check(b2k_dec_get_raw_lattice(dec, chans_[i], &q, cudaStreamPerThread), "b2k_dec_get_raw_lattice");
        batch_lat_[i](corr_ids[i], q);
      }
      {
        std::lock_guard<std::mutex> lock(chan_mu_);
        free_.push_back(chans_[i]);
        chan_.erase(corr_ids[i]);
      }
      finished_.push_back(corr_ids[i]);
    }
  }

 private:
  b2k_stream *s_ = NULL;
  const b2k_model *model_;
  int32 frames_per_chunk_ = 0, samples_per_chunk_ = 0, num_tids_ = 0, path_cap_ = 4096;
  BaseFloat decoder_frame_shift_seconds_ = 0.03f;
  b2k_endpoint_cfg endpoint_;
  std::vector<int32> free_;
  std::map<CorrelationID, int32> chan_;
  std::mutex chan_mu_, cb_mu_;
  std::map<CorrelationID, BestPathCallback> best_cb_;
  std::map<CorrelationID, RawLatticeCallback> lat_cb_;
  std::vector<BestPathCallback> batch_best_;
  std::vector<RawLatticeCallback> batch_lat_;
  std::map<CorrelationID, std::string> text_;
  std::vector<CorrelationID> finished_;
  std::function<std::string(int32)> word_of_;
  std::vector<std::vector<int16_t>> pcm_;
  std::vector<int32_t> chans_, ns_, first_, last_, il_, ol_;
  std::vector<const int16_t *> ptrs_;
  std::vector<b2k_best_path_info> info_;
  std::vector<int32_t> st_f_, st_h_, a_s_, a_d_, a_i_, a_o_, f_s_;
  std::vector<float> st_t_, st_e_, a_g_, a_a_, f_c_;
  KALDI_DISALLOW_COPY_AND_ASSIGN(StreamingOnlinePipelineB2k);
};

// cuda_decoder::CudaOnlinePipelineDynamicBatcher (cudadecoder/cuda-online-pipeline-dynamic-batcher.h:38-60) in front of
// StreamingOnlinePipelineB2k: producers Push chunks from any thread, a worker thread forms the batches (at most one chunk per
// stream and batch, a stream's chunks in order, a batch as soon as it is full or the timeout has passed) and calls DecodeBatch.
// The scheduling is b2k_host::DynamicBatcher (b2k_dynamic_batcher.h), unit-tested with a mock pipeline.
class CudaOnlinePipelineDynamicBatcherB2k {
 public:
  typedef StreamingOnlinePipelineB2k::CorrelationID CorrelationID;
  // config.dynamic_batcher_timeout as in CudaOnlinePipelineDynamicBatcherConfig (:34-36)
  CudaOnlinePipelineDynamicBatcherB2k(double dynamic_batcher_timeout, StreamingOnlinePipelineB2k &pipeline, int32 max_batch_size)
      : adapter_{&pipeline, max_batch_size}, batcher_(&adapter_, dynamic_batcher_timeout) {}
  void Push(CorrelationID corr_id, bool is_first_chunk, bool is_last_chunk, const SubVector<BaseFloat> &wave_samples) {
    batcher_.Push(corr_id, is_first_chunk, is_last_chunk, wave_samples.Data(), wave_samples.Dim());
  }
  void WaitForCompletion() {
    try {
      batcher_.WaitForCompletion();
    } catch (const std::exception &e) {
      KALDI_ERR << "dynamic batcher: " << e.what();
    }
  }
  int GetNumPendingChunks(CorrelationID corr_id) { return batcher_.GetNumPendingChunks(corr_id); }

 private:
  struct Adapter {
    StreamingOnlinePipelineB2k *p;
    int32 max_batch;
    int MaxBatchSize() const { return max_batch; }
    bool TryInitCorrID(uint64_t id) { return p->TryInitCorrID(id); }
    void DecodeBatch(const std::vector<uint64_t> &ids, const std::vector<std::pair<const float *, int64_t>> &chunks,
                     const std::vector<bool> &first, const std::vector<bool> &last) {
      std::vector<SubVector<BaseFloat>> waves;
      static BaseFloat none = 0.0f;
      for (const auto &c : chunks) waves.push_back(SubVector<BaseFloat>(c.second > 0 ? const_cast<BaseFloat *>(c.first) : &none, (MatrixIndexT)c.second));
      p->DecodeBatch(ids, waves, first, last);
    }
  };
  Adapter adapter_;
  b2k_host::DynamicBatcher<Adapter> batcher_;
};

// ------------------------------------------------------------------------------------------------ endpointing
// online2/online-endpoint.h:128-190.  Templates on the configuration type so that this header does not have to include
// online-endpoint.h (which pulls both online decoders, hence OpenFst, in): instantiate them with kaldi::OnlineEndpointConfig.
template <class EndpointConfig>
inline b2k_endpoint_cfg ToB2kEndpointConfig(const EndpointConfig &config) {
  b2k_endpoint_cfg c;
  Check(b2k_endpoint_cfg_default(&c), "b2k_endpoint_cfg_default");
  const decltype(config.rule1) *rules[5] = {&config.rule1, &config.rule2, &config.rule3, &config.rule4, &config.rule5};
  for (int r = 0; r < 5; r++) {
    c.rule[r].must_contain_nonsilence = rules[r]->must_contain_nonsilence ? 1 : 0;
    c.rule[r].min_trailing_silence = rules[r]->min_trailing_silence;
    c.rule[r].max_relative_cost = rules[r]->max_relative_cost;
    c.rule[r].min_utterance_length = rules[r]->min_utterance_length;
  }
  if (config.silence_phones.size() >= sizeof(c.silence_phones)) KALDI_ERR << "--endpoint.silence-phones is too long";
  config.silence_phones.copy(c.silence_phones, config.silence_phones.size());
  c.silence_phones[config.silence_phones.size()] = 0;
  return c;
}

// EndpointDetected(config, num_frames_decoded, trailing_silence_frames, frame_shift_in_seconds, final_relative_cost)  online-endpoint.h:171
template <class EndpointConfig>
inline bool EndpointDetectedB2k(const EndpointConfig &config, int32 num_frames_decoded, int32 trailing_silence_frames,
                                BaseFloat frame_shift_in_seconds, BaseFloat final_relative_cost) {
  const b2k_endpoint_cfg c = ToB2kEndpointConfig(config);
  int32_t hit = 0;
  Check(b2k_endpoint_detected(&c, num_frames_decoded, trailing_silence_frames, frame_shift_in_seconds, final_relative_cost, &hit), "EndpointDetected");
  return hit != 0;
}

// EndpointDetected(config, tmodel, frame_shift_in_seconds, decoder)  online-endpoint.h:185: the decoder's part is passed in as the
// best path's input labels in time order (CudaDecoderB2k::GetBestPath / b2k_lat_best_path_arcs), the frame count and the final
// relative cost; TransitionIdToPhone comes from the reference's own transition model.
template <class EndpointConfig>
inline bool EndpointDetectedB2k(const EndpointConfig &config, const TransitionInformation &tmodel, BaseFloat frame_shift_in_seconds,
                                const std::vector<int32> &best_path_ilabels, int32 num_frames_decoded, BaseFloat final_relative_cost) {
  const b2k_endpoint_cfg c = ToB2kEndpointConfig(config);
  int32 max_tid = 0;
  for (int32 t : best_path_ilabels) max_tid = std::max(max_tid, t);
  std::vector<int32_t> tid2phone((size_t)max_tid + 1, 0);
  for (int32 t : best_path_ilabels) if (t > 0) tid2phone[t] = tmodel.TransitionIdToPhone(t);
  int32_t hit = 0;
  Check(b2k_endpoint_detected_on_path(&c, tid2phone.data(), (int32_t)tid2phone.size(), best_path_ilabels.data(), (int64_t)best_path_ilabels.size(),
                                      num_frames_decoded, frame_shift_in_seconds, final_relative_cost, &hit, nullptr), "EndpointDetected");
  return hit != 0;
}

}  // namespace b2k_shim
}  // namespace kaldi
#endif  // B2K_KALDI_SHIMS_H_
