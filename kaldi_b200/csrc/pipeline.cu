// pipeline.cu — the batched online2 pipeline in C++ above the stage entry points of this library: host waveforms in,
// finalized raw lattices out.  It stands behind cuda_decoder::BatchedThreadedNnet3CudaOnlinePipeline::DecodeBatch
// (cudadecoder/batched-threaded-nnet3-cuda-online-pipeline.cc:316-377: ComputeGPUFeatureExtraction -> RunNnet3 ->
// RunDecoder -> finalize) with the numerical semantics of online2-wav-nnet3-latgen-faster
// (online2bin/online2-wav-nnet3-latgen-faster.cc:199-299): per-chunk i-vectors as the looped decodable would have
// received them, CPU decoder search semantics, one finalized raw lattice per utterance.  No device code of its own:
// CUDA runtime calls for the buffers and copies, the b2k stage calls for the work.  kaldi_b200/pipeline.py
// (BatchedPipeline) is the Python spelling of the same sequence; the sizing rules are shared through
// b2k_pipeline_plan_for, which needs no device (tests/test_pipeline_plan.py).
#include <algorithm>
#include <thread>
#include <vector>

#include "common.cuh"

using b2k::set_error;

struct b2k_pipeline {
  b2k_pipeline_cfg cfg;
  b2k_pipeline_plan plan;
  b2k_feat *feat = nullptr;
  b2k_nnet *nnet = nullptr;
  b2k_dec *dec = nullptr;
  b2k_ivec *ivec = nullptr;          // not owned
  float *h_wave = nullptr;           // pinned [max_batch x num_samples]
  float *d_wave = nullptr, *d_feats = nullptr, *d_ivec = nullptr, *d_loglikes = nullptr;
  float *d_feats_cmvn = nullptr;     // use_cmvn: what the network reads (the i-vector stage keeps reading d_feats)
  double *d_cmvn_state = nullptr, *d_cmvn_global = nullptr;
  std::vector<const float *> p_nnet_in;
  std::vector<float *> p_cmvn_out;
  std::vector<double *> p_cmvn_state;
  std::vector<int32_t> sched, channels, ns, zeros, nframes, nout;
  std::vector<const float *> p_wave, p_feats, p_ivec, p_ll;
  std::vector<const int16_t *> p_wave16;
  std::vector<float *> p_feats_out, p_ivec_out, p_ll_out;
  int32_t last_n = 0;
  // speaker adaptation of the NEXT batch (b2k_pipeline_set_speaker_states; cleared when the batch has been launched)
  std::vector<const double *> spk_in;
  std::vector<double *> spk_out;
  float spk_max_remembered = -1.0f;
  // pipelined operation (submit / collect): two batches in flight
  struct Slot {
    int16_t *h_wave16 = nullptr;      // pinned [max_batch x num_samples]
    int16_t *d_wave16 = nullptr;
    char *d_pack = nullptr, *h_pack = nullptr; size_t pack_cap = 0, h_pack_cap = 0;
    cudaEvent_t h2d_done = nullptr, wave_consumed = nullptr, packed = nullptr, pack_read = nullptr;
    int32_t n = 0;
    bool busy = false;
  } slot[2];
  cudaStream_t st_copy = nullptr, st_compute = nullptr;
  int32_t *d_channels = nullptr;
  int64_t n_submitted = 0, n_collected = 0;
  size_t pack_floor = 0;
  // optional stage timing (bench.py): events around the stages of the last run_device
  bool timing = false;
  cudaEvent_t tev[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};             // lower bound for the packed-lattice buffers, raised when a batch comes close to the capacity
  // host result of the last collect
  std::vector<int32_t> r_i32[7]; std::vector<float> r_f32[5];
  std::vector<int64_t> r_offs[3];
};

namespace {

int frame_count(const b2k_feat_cfg &f, int64_t num_samples) {   // NumFrames, flush = true (feat/feature-window.cc:42-87)
  const int64_t length = (int)(f.samp_freq * 0.001f * f.frame_length_ms), shift = (int)(f.samp_freq * 0.001f * f.frame_shift_ms);
  if (shift <= 0 || length <= 0) return -1;
  if (f.snip_edges) return num_samples < length ? 0 : (int)(1 + (num_samples - length) / shift);
  return (int)((num_samples + shift / 2) / shift);
}

template <typename S>
void stage_rows(float *dst, const S *const *src, int32_t n, int64_t len) {
  // float input: plain copies; int16 input: widened on the way (Kaldi keeps int16-range values in floats)
  auto work = [&](int32_t a, int32_t b) {
    for (int32_t i = a; i < b; i++) {
      float *d = dst + (size_t)i * len;
      const S *s = src[i];
      for (int64_t k = 0; k < len; k++) d[k] = (float)s[k];
    }
  };
  const int64_t total = (int64_t)n * len;
  unsigned nt = total < (1 << 21) ? 1u : std::min<unsigned>({8u, std::max(1u, std::thread::hardware_concurrency()), (unsigned)n});
  if (nt <= 1) { work(0, n); return; }
  std::vector<std::thread> th;
  for (unsigned t = 0; t < nt; t++) th.emplace_back(work, (int32_t)((int64_t)n * t / nt), (int32_t)((int64_t)n * (t + 1) / nt));
  for (auto &t : th) t.join();
}

}  // namespace

extern "C" {

void b2k_pipeline_cfg_default(b2k_pipeline_cfg *c) {
  if (!c) return;
  memset(c, 0, sizeof(*c));
  b2k_feat_cfg_default(&c->feat);
  b2k_dec_cfg_default(&c->dec);
  c->dec.max_frames = 0; c->dec.max_tokens = 0; c->dec.max_links = 0;   // sized from the utterance length
  c->frames_per_chunk = 21;          // --frames-per-chunk=20 rounded up to a multiple of 3 (GetChunkSize, nnet3/nnet-compile-looped.cc:81; decodable-simple-looped.cc:66)
  c->acoustic_scale = 1.0f;          // chain models decode with --acwt 1.0
  c->max_batch = 64;
  c->num_samples = 160000;
  c->chunk_length_secs = 0.18f;      // online2-wav-nnet3-latgen-faster --chunk-length
  c->ivector_splice_right = 3;
  c->use_priors = 1;
  c->use_cmvn = 0;                   // OnlineNnet2FeaturePipelineInfo::use_cmvn: off unless --cmvn-config is given
  c->cmvn.cmn_window = 600; c->cmvn.speaker_frames = 600; c->cmvn.global_frames = 200; c->cmvn.normalize_mean = 1; c->cmvn.normalize_variance = 0;
}

// --frame-subsampling-factor: the caller's value if given, else the model's when its layers decide it
// (b2k_model_frame_subsampling_ambiguous); never a silent guess.
static int resolve_subsampling(const b2k_pipeline_cfg *cfg, const b2k_model *model, int model_factor, int32_t *out) {
  const int amb = b2k_model_frame_subsampling_ambiguous(model);
  if (cfg->frame_subsampling_factor < 0) return set_error(B2K_ERR_INVALID, "frame_subsampling_factor must be positive (0 = take the model's)");
  if (cfg->frame_subsampling_factor == 0) {
    if (amb) return set_error(B2K_ERR_INVALID, "the model's layers do not decide the frame subsampling factor (splices at +-3 without a stride-3 TDNN-F layer): "
                                              "pass --frame-subsampling-factor (3 for chain models, 1 otherwise)");
    *out = model_factor;
    return B2K_OK;
  }
  if (!amb && model_factor == 3 && cfg->frame_subsampling_factor != 3)
    return set_error(B2K_ERR_INVALID, "--frame-subsampling-factor disagrees with the model (TDNN-F layers with time-stride 3: a chain model, factor 3)");
  *out = cfg->frame_subsampling_factor;
  return B2K_OK;
}

// samples per AcceptWaveform call of online2-wav-nnet3-latgen-faster (:234-240): int32(samp_freq * chunk_length_secs)
// in float arithmetic (truncated, not rounded), 0 becomes 1, and chunk_length_secs <= 0 (--online=false) = the whole file
static int32_t chunk_samples_of(float samp_freq, float chunk_length_secs, int64_t num_samples) {
  if (chunk_length_secs <= 0.0f) return (int32_t)std::min<int64_t>(num_samples, 0x7fffffff);
  int32_t c = (int32_t)(samp_freq * chunk_length_secs);
  return c == 0 ? 1 : c;
}

int b2k_pipeline_plan_for(const b2k_pipeline_cfg *cfg, const b2k_model *model, b2k_pipeline_plan *plan) {
  if (!cfg || !model || !plan) return set_error(B2K_ERR_INVALID, "b2k_pipeline_plan_for: bad args");
  if (cfg->max_batch <= 0 || cfg->num_samples <= 0 || cfg->frames_per_chunk <= 0)
    return set_error(B2K_ERR_INVALID, "b2k_pipeline_plan_for: max_batch, num_samples and frames_per_chunk must be positive");
  int32_t mi[8];
  int rc = b2k_model_info(model, mi);
  if (rc) return rc;
  memset(plan, 0, sizeof(*plan));
  const int T = frame_count(cfg->feat, cfg->num_samples);
  if (T <= 0) return set_error(B2K_ERR_INVALID, "b2k_pipeline_plan_for: the utterance length gives no feature frame");
  const int D = cfg->feat.feature_type == 0 ? cfg->feat.num_ceps : cfg->feat.num_bins + (cfg->feat.use_energy ? 1 : 0);
  if (D != mi[0]) return set_error(B2K_ERR_INVALID, "b2k_pipeline_plan_for: feature dimension differs from the model's input dimension");
  if (cfg->use_cmvn && !cfg->global_cmvn_stats) return set_error(B2K_ERR_INVALID, "b2k_pipeline_plan_for: use_cmvn needs global_cmvn_stats (online-feature.cc:417)");
  int sub = 0;
  if ((rc = resolve_subsampling(cfg, model, mi[3], &sub))) return rc;
  if (cfg->frames_per_chunk % sub) return set_error(B2K_ERR_INVALID, "b2k_pipeline_plan_for: frames_per_chunk must be a multiple of the frame subsampling factor");
  plan->num_feature_frames = T;
  plan->feat_dim = D;
  plan->num_output_frames = (T + sub - 1) / sub;
  plan->num_chunks = (plan->num_output_frames * sub + cfg->frames_per_chunk - 1) / cfg->frames_per_chunk;
  plan->num_pdfs = mi[2];
  plan->ivector_dim = mi[1];
  plan->chunk_samples = chunk_samples_of(cfg->feat.samp_freq, cfg->chunk_length_secs, cfg->num_samples);
  plan->dec = cfg->dec;
  const int64_t nf = plan->num_output_frames;
  if (plan->dec.max_frames <= 0) plan->dec.max_frames = (int32_t)(nf + 2);
  if (plan->dec.max_tokens <= 0) plan->dec.max_tokens = nf * 9000;
  if (plan->dec.max_links <= 0) plan->dec.max_links = nf * 16000;
  const int64_t B = cfg->max_batch;
  plan->device_bytes = 4 * B * (cfg->num_samples + (int64_t)T * D + (int64_t)plan->num_chunks * std::max(1, plan->ivector_dim) +
                                nf * plan->num_pdfs);
  if (cfg->use_cmvn) plan->device_bytes += 4 * B * (int64_t)T * D + 8 * (B + 1) * 2 * (D + 1);
  plan->pinned_bytes = 4 * B * cfg->num_samples;
  return B2K_OK;
}

int b2k_pipeline_destroy(b2k_pipeline *p) {
  if (!p) return B2K_OK;
  if (p->dec) b2k_dec_destroy(p->dec);
  if (p->nnet) b2k_nnet_destroy(p->nnet);
  if (p->feat) b2k_feat_destroy(p->feat);
  if (p->h_wave) cudaFreeHost(p->h_wave);
  for (auto &sl : p->slot) {
    if (sl.h_wave16) cudaFreeHost(sl.h_wave16);
    if (sl.d_wave16) cudaFree(sl.d_wave16);
    if (sl.d_pack) cudaFree(sl.d_pack);
    if (sl.h_pack) cudaFreeHost(sl.h_pack);
    for (cudaEvent_t e : {sl.h2d_done, sl.wave_consumed, sl.packed, sl.pack_read}) if (e) cudaEventDestroy(e);
  }
  if (p->d_channels) cudaFree(p->d_channels);
  for (cudaEvent_t e : p->tev) if (e) cudaEventDestroy(e);
  if (p->st_copy) cudaStreamDestroy(p->st_copy);
  if (p->st_compute) cudaStreamDestroy(p->st_compute);
  for (float *d : {p->d_wave, p->d_feats, p->d_ivec, p->d_loglikes, p->d_feats_cmvn}) if (d) cudaFree(d);
  for (double *d : {p->d_cmvn_state, p->d_cmvn_global}) if (d) cudaFree(d);
  delete p;
  return B2K_OK;
}

static int pipeline_create_impl(b2k_pipeline *p, const b2k_model *model, const b2k_fst *fst) {
  const b2k_pipeline_cfg &cfg = p->cfg;
  const b2k_pipeline_plan &pl = p->plan;
  int32_t mi[8];
  int rc = b2k_model_info(model, mi);
  if (rc) return rc;
  // nnet3: compile for this utterance length, upload
  b2k_nnet_compile_cfg cc;
  memset(&cc, 0, sizeof(cc));
  cc.feat_dim = mi[0]; cc.ivector_dim = mi[1]; cc.num_pdfs = mi[2];
  if ((rc = resolve_subsampling(&cfg, model, mi[3], &cc.frame_subsampling_factor))) return rc;
  cc.num_frames = pl.num_feature_frames; cc.frames_per_chunk = cfg.frames_per_chunk; cc.use_priors = cfg.use_priors;
  cc.conv_dense = cfg.conv_dense; cc.acoustic_scale = cfg.acoustic_scale;
  b2k_nnet_program *prog = nullptr;
  rc = b2k_nnet_compile(&cc, b2k_model_layers(model), mi[4], b2k_model_weights(model), mi[5], &prog);
  if (rc) return rc;
  int64_t pi[8];
  b2k_nnet_program_info(prog, pi);
  rc = b2k_nnet_create_from_program(prog, cfg.max_batch, &p->nnet);
  b2k_nnet_program_destroy(prog);
  if (rc) return rc;
  if (pi[0] != pl.num_output_frames || pi[1] != pl.num_chunks)
    return set_error(B2K_ERR_STATE, "b2k_pipeline_create: the compiled program disagrees with the plan");
  p->plan.model_right_context = (int32_t)pi[5];
  // features
  b2k_feat_cfg fc = cfg.feat;
  fc.max_lanes = cfg.max_batch;
  rc = b2k_feat_create(&fc, &p->feat);
  if (rc) return rc;
  if (b2k_feat_num_frames(p->feat, cfg.num_samples, 1) != pl.num_feature_frames || b2k_feat_dim(p->feat) != pl.feat_dim)
    return set_error(B2K_ERR_STATE, "b2k_pipeline_create: the feature stage disagrees with the plan");
  // decoder: one lane per batch slot, channel = lane
  rc = b2k_dec_create(fst, &pl.dec, cfg.max_batch, cfg.max_batch, &p->dec);
  if (rc) return rc;
  // buffers
  const size_t B = (size_t)cfg.max_batch, S = (size_t)cfg.num_samples, T = (size_t)pl.num_feature_frames, D = (size_t)pl.feat_dim;
  const size_t NC = (size_t)pl.num_chunks, IV = (size_t)std::max(1, pl.ivector_dim), NF = (size_t)pl.num_output_frames, P = (size_t)pl.num_pdfs;
  B2K_CUDA_CHECK(cudaMallocHost((void **)&p->h_wave, 4 * B * S));
  B2K_CUDA_CHECK(cudaMalloc((void **)&p->d_wave, 4 * B * S));
  B2K_CUDA_CHECK(cudaMalloc((void **)&p->d_feats, 4 * B * T * D));
  B2K_CUDA_CHECK(cudaMalloc((void **)&p->d_ivec, 4 * B * NC * IV));
  B2K_CUDA_CHECK(cudaMemset(p->d_ivec, 0, 4 * B * NC * IV));      // no extractor: the network sees zero i-vectors
  B2K_CUDA_CHECK(cudaMalloc((void **)&p->d_loglikes, 4 * B * NF * P));
  p->p_nnet_in.resize(B);
  if (cfg.use_cmvn) {      // OnlineNnet2FeaturePipeline: base -> OnlineCmvn -> network input (online-nnet2-feature-pipeline.cc:108-123)
    const size_t SD = 2 * (D + 1);
    B2K_CUDA_CHECK(cudaMalloc((void **)&p->d_feats_cmvn, 4 * B * T * D));
    B2K_CUDA_CHECK(cudaMalloc((void **)&p->d_cmvn_state, 8 * B * SD));
    B2K_CUDA_CHECK(cudaMalloc((void **)&p->d_cmvn_global, 8 * SD));
    B2K_CUDA_CHECK(cudaMemcpy(p->d_cmvn_global, cfg.global_cmvn_stats, 8 * SD, cudaMemcpyHostToDevice));
    p->p_cmvn_out.resize(B); p->p_cmvn_state.resize(B);
    for (size_t i = 0; i < B; i++) { p->p_cmvn_out[i] = p->d_feats_cmvn + i * T * D; p->p_cmvn_state[i] = p->d_cmvn_state + i * SD; }
  }
  p->channels.resize(B); p->ns.assign(B, (int32_t)S); p->zeros.assign(B, 0); p->nframes.assign(B, (int32_t)T); p->nout.assign(B, (int32_t)NF);
  p->p_wave.resize(B); p->p_feats.resize(B); p->p_ivec.resize(B); p->p_ll.resize(B);
  p->p_feats_out.resize(B); p->p_ivec_out.resize(B); p->p_ll_out.resize(B);
  for (size_t i = 0; i < B; i++) {
    p->channels[i] = (int32_t)i;
    p->p_wave[i] = p->d_wave + i * S;
    p->p_feats[i] = p->p_feats_out[i] = p->d_feats + i * T * D;
    p->p_ivec[i] = p->p_ivec_out[i] = p->d_ivec + i * NC * IV;
    p->p_ll[i] = p->p_ll_out[i] = p->d_loglikes + i * NF * P;
    p->p_nnet_in[i] = cfg.use_cmvn ? p->d_feats_cmvn + i * T * D : p->d_feats + i * T * D;
  }
  if (p->ivec) {
    p->sched.resize(NC);
    int32_t nc = 0;
    const b2k_feat_cfg &f = cfg.feat;
    rc = b2k_ivec_online_schedule(cfg.num_samples, pl.chunk_samples, (int)(f.samp_freq * 0.001f * f.frame_length_ms),
                                  (int)(f.samp_freq * 0.001f * f.frame_shift_ms), pl.num_feature_frames, p->plan.model_right_context,
                                  cfg.frames_per_chunk, mi[3], cfg.ivector_splice_right, p->sched.data(), (int32_t)NC, &nc);
    if (rc) return rc;
    if (nc != (int32_t)NC) return set_error(B2K_ERR_STATE, "b2k_pipeline_create: i-vector schedule length differs from the chunk count");
  }
  return B2K_OK;
}

int b2k_pipeline_create(const b2k_pipeline_cfg *cfg, const b2k_model *model, const b2k_fst *fst, b2k_ivec *ivec,
                        b2k_pipeline **out) {
  if (!cfg || !model || !fst || !out) return set_error(B2K_ERR_INVALID, "b2k_pipeline_create: bad args");
  *out = nullptr;
  b2k_pipeline_plan plan;
  int rc = b2k_pipeline_plan_for(cfg, model, &plan);
  if (rc) return rc;
  if (ivec && plan.ivector_dim <= 0) return set_error(B2K_ERR_INVALID, "b2k_pipeline_create: an extractor was given but the model takes no i-vector");
  rc = b2k::require_device();
  if (rc) return rc;
  b2k_pipeline *p = new b2k_pipeline();
  p->cfg = *cfg; p->plan = plan; p->ivec = ivec;
  rc = pipeline_create_impl(p, model, fst);
  if (rc) { const std::string keep = b2k::g_last_error; b2k_pipeline_destroy(p); b2k::g_last_error = keep; return rc; }
  *out = p;
  return B2K_OK;
}

int b2k_pipeline_get_plan(const b2k_pipeline *p, b2k_pipeline_plan *plan) {
  if (!p || !plan) return set_error(B2K_ERR_INVALID, "b2k_pipeline_get_plan: bad args");
  *plan = p->plan;
  return B2K_OK;
}

int b2k_pipeline_set_speaker_states(b2k_pipeline *p, int32_t n, const double *const *d_state_in, double *const *d_state_out,
                                    float max_remembered_frames) {
  if (!p || n < 0 || n > p->cfg.max_batch || (n > 0 && (!d_state_in || !d_state_out)))
    return set_error(B2K_ERR_INVALID, "b2k_pipeline_set_speaker_states: bad args");
  if (n > 0 && !p->ivec) return set_error(B2K_ERR_STATE, "b2k_pipeline_set_speaker_states: the pipeline has no i-vector extractor");
  for (int32_t i = 0; i < n; i++)
    for (int32_t j = 0; j < i; j++)
      if (d_state_out[i] && d_state_out[i] == d_state_out[j])
        return set_error(B2K_ERR_INVALID, "b2k_pipeline_set_speaker_states: two utterances of one batch write the same speaker state (a speaker's utterances go into successive batches)");
  p->spk_in.assign(d_state_in, d_state_in + n);
  p->spk_out.assign(d_state_out, d_state_out + n);
  p->spk_max_remembered = max_remembered_frames;
  return B2K_OK;
}

b2k_dec *b2k_pipeline_decoder(b2k_pipeline *p) { return p ? p->dec : nullptr; }
const float *b2k_pipeline_loglikes(const b2k_pipeline *p) { return p ? p->d_loglikes : nullptr; }
const float *b2k_pipeline_features(const b2k_pipeline *p) { return p ? p->d_feats : nullptr; }
const float *b2k_pipeline_ivectors(const b2k_pipeline *p) { return p ? p->d_ivec : nullptr; }

}  // extern "C"

// All stages for the first n batch slots, inputs already in d_wave (or, d_wave16 given, 16-bit PCM [n x num_samples] that the
// feature kernel reads directly; wave_consumed is recorded once it has); asynchronous on `stream`.
static int run_device(b2k_pipeline *p, int32_t n, void *stream, const int16_t *d_wave16 = nullptr, cudaEvent_t wave_consumed = nullptr) {
  const b2k_pipeline_plan &pl = p->plan;
  auto tick = [&](int i) { if (p->timing) cudaEventRecord(p->tev[i], (cudaStream_t)stream); };
  tick(0);
  int rc;
  if (d_wave16) {
    p->p_wave16.resize(n);
    for (int32_t i = 0; i < n; i++) p->p_wave16[i] = d_wave16 + (size_t)i * (size_t)p->cfg.num_samples;
    rc = b2k_feat_compute_batched_i16(p->feat, n, p->p_wave16.data(), p->ns.data(), p->zeros.data(), p->nframes.data(),
                                      p->p_feats_out.data(), pl.feat_dim, stream);
  } else {
    rc = b2k_feat_compute_batched(p->feat, n, p->p_wave.data(), p->ns.data(), p->zeros.data(), p->nframes.data(),
                                  p->p_feats_out.data(), pl.feat_dim, stream);
  }
  if (rc) return rc;
  if (wave_consumed) B2K_CUDA_CHECK(cudaEventRecord(wave_consumed, (cudaStream_t)stream));
  tick(1);
  if (p->ivec) {
    const bool adapt = (int32_t)p->spk_in.size() == n;
    rc = b2k_ivec_compute_batched_adapt(p->ivec, n, p->p_feats.data(), pl.feat_dim, pl.num_feature_frames, p->sched.data(),
                                        pl.num_chunks, p->p_ivec_out.data(), pl.ivector_dim, adapt ? p->spk_in.data() : nullptr,
                                        adapt ? p->spk_out.data() : nullptr, p->spk_max_remembered, stream);
    p->spk_in.clear(); p->spk_out.clear();
    if (rc) return rc;
  } else if (!p->spk_in.empty()) {
    p->spk_in.clear(); p->spk_out.clear();
    return set_error(B2K_ERR_STATE, "speaker states were set on a pipeline without an i-vector extractor");
  }
  if (p->cfg.use_cmvn) {                     // every utterance starts from empty sliding-window stats and no speaker stats
    const size_t SD = 2 * ((size_t)pl.feat_dim + 1);
    B2K_CUDA_CHECK(cudaMemsetAsync(p->d_cmvn_state, 0, 8 * (size_t)n * SD, (cudaStream_t)stream));
    rc = b2k_cmvn_apply_batched(p->feat, &p->cfg.cmvn, n, p->p_feats.data(), p->p_cmvn_out.data(), pl.feat_dim, pl.feat_dim,
                                p->zeros.data(), p->nframes.data(), p->p_cmvn_state.data(), p->d_cmvn_global, nullptr, stream);
    if (rc) return rc;
  }
  tick(2);
  rc = b2k_nnet_run(p->nnet, n, p->p_nnet_in.data(), pl.feat_dim, pl.ivector_dim > 0 ? p->p_ivec.data() : nullptr, pl.ivector_dim,
                    p->p_ll_out.data(), pl.num_pdfs, stream);
  if (rc) return rc;
  tick(3);
  rc = b2k_dec_init_decoding(p->dec, p->channels.data(), n, stream);
  if (rc) return rc;
  rc = b2k_dec_advance_decoding_frames(p->dec, p->channels.data(), p->p_ll.data(), p->nout.data(), pl.num_pdfs, n, stream);
  if (rc) return rc;
  tick(4);
  rc = b2k_dec_finalize_decoding(p->dec, p->channels.data(), n, stream);
  tick(5);
  return rc;
}

extern "C" int b2k_pipeline_enable_stage_timing(b2k_pipeline *p, int32_t on) {
  if (!p) return set_error(B2K_ERR_INVALID, "b2k_pipeline_enable_stage_timing: bad args");
  if (on && !p->tev[0]) for (auto &e : p->tev) B2K_CUDA_CHECK(cudaEventCreate(&e));
  p->timing = on != 0;
  return B2K_OK;
}
// ms of {features, i-vectors (+ CMVN), nnet3, decoder init + advance, decoder finalize} of the last run (waits for it)
extern "C" int b2k_pipeline_stage_times(b2k_pipeline *p, float ms[5]) {
  if (!p || !ms || !p->timing) return set_error(B2K_ERR_INVALID, "b2k_pipeline_stage_times: timing is not enabled");
  B2K_CUDA_CHECK(cudaEventSynchronize(p->tev[5]));
  for (int i = 0; i < 5; i++) B2K_CUDA_CHECK(cudaEventElapsedTime(&ms[i], p->tev[i], p->tev[i + 1]));
  return B2K_OK;
}
extern "C" double b2k_pipeline_nnet_flops_per_utterance(const b2k_pipeline *p) { return p && p->nnet ? b2k_nnet_flops_per_lane(p->nnet) : 0.0; }

template <typename S>
static int decode_batch(b2k_pipeline *p, int32_t n, const S *const *h_waves, void *stream) {
  if (!p || n <= 0 || n > p->cfg.max_batch || !h_waves) return set_error(B2K_ERR_INVALID, "b2k_pipeline_decode_batch: bad args");
  for (int32_t i = 0; i < n; i++) if (!h_waves[i]) return set_error(B2K_ERR_INVALID, "b2k_pipeline_decode_batch: null waveform");
  cudaStream_t st = (cudaStream_t)stream;
  // the previous batch's H2D copy reads the pinned buffer: it must have completed (it has if the caller read lattices)
  B2K_CUDA_CHECK(cudaStreamSynchronize(st));
  stage_rows<S>(p->h_wave, h_waves, n, p->cfg.num_samples);
  B2K_CUDA_CHECK(cudaMemcpyAsync(p->d_wave, p->h_wave, 4 * (size_t)n * (size_t)p->cfg.num_samples, cudaMemcpyHostToDevice, st));
  p->last_n = n;
  return run_device(p, n, stream);
}

// ---------------------------------------------------------------- pipelined operation
static int ensure_pipelined(b2k_pipeline *p) {
  if (p->st_copy) return B2K_OK;
  B2K_CUDA_CHECK(cudaStreamCreateWithFlags(&p->st_copy, cudaStreamNonBlocking));
  B2K_CUDA_CHECK(cudaStreamCreateWithFlags(&p->st_compute, cudaStreamNonBlocking));
  const size_t B = (size_t)p->cfg.max_batch, S = (size_t)p->cfg.num_samples;
  for (auto &sl : p->slot) {
    B2K_CUDA_CHECK(cudaMallocHost((void **)&sl.h_wave16, 2 * B * S));
    B2K_CUDA_CHECK(cudaMalloc((void **)&sl.d_wave16, 2 * B * S + 16));
    for (cudaEvent_t *e : {&sl.h2d_done, &sl.wave_consumed, &sl.packed, &sl.pack_read})
      B2K_CUDA_CHECK(cudaEventCreateWithFlags(e, cudaEventDisableTiming));
  }
  if (!p->d_channels) {
    B2K_CUDA_CHECK(cudaMalloc((void **)&p->d_channels, 4 * B));
    B2K_CUDA_CHECK(cudaMemcpy(p->d_channels, p->channels.data(), 4 * B, cudaMemcpyHostToDevice));
  }
  return B2K_OK;
}

extern "C" int b2k_pipeline_submit_i16(b2k_pipeline *p, int32_t n, const int16_t *const *h_waves) {
  if (!p || n <= 0 || n > p->cfg.max_batch || !h_waves) return set_error(B2K_ERR_INVALID, "b2k_pipeline_submit_i16: bad args");
  for (int32_t i = 0; i < n; i++) if (!h_waves[i]) return set_error(B2K_ERR_INVALID, "b2k_pipeline_submit_i16: null waveform");
  if (p->n_submitted - p->n_collected >= 2) return set_error(B2K_ERR_STATE, "b2k_pipeline_submit_i16: two batches are outstanding, collect one first");
  int rc = ensure_pipelined(p);
  if (rc) return rc;
  b2k_pipeline::Slot &sl = p->slot[p->n_submitted & 1];
  const size_t S = (size_t)p->cfg.num_samples, total = (size_t)n * S;
  // the pinned buffer of this slot was last read by the copy of batch k-2
  if (p->n_submitted >= 2) B2K_CUDA_CHECK(cudaEventSynchronize(sl.h2d_done));
  {
    auto work = [&](int32_t a, int32_t b) { for (int32_t i = a; i < b; i++) memcpy(sl.h_wave16 + (size_t)i * S, h_waves[i], 2 * S); };
    const unsigned nt = total < (1u << 21) ? 1u : std::min<unsigned>({8u, std::max(1u, std::thread::hardware_concurrency()), (unsigned)n});
    if (nt <= 1) work(0, n);
    else {
      std::vector<std::thread> th;
      for (unsigned t = 0; t < nt; t++) th.emplace_back(work, (int32_t)((int64_t)n * t / nt), (int32_t)((int64_t)n * (t + 1) / nt));
      for (auto &t : th) t.join();
    }
  }
  // copy stream: the device buffer of this slot was last read by the conversion of batch k-2
  if (p->n_submitted >= 2) B2K_CUDA_CHECK(cudaStreamWaitEvent(p->st_copy, sl.wave_consumed, 0));
  B2K_CUDA_CHECK(cudaMemcpyAsync(sl.d_wave16, sl.h_wave16, 2 * total, cudaMemcpyHostToDevice, p->st_copy));
  B2K_CUDA_CHECK(cudaEventRecord(sl.h2d_done, p->st_copy));
  // compute stream: the four stages (the feature kernel reads the 16-bit PCM itself), finalize, pack
  B2K_CUDA_CHECK(cudaStreamWaitEvent(p->st_compute, sl.h2d_done, 0));
  p->last_n = n;
  rc = run_device(p, n, (void *)p->st_compute, sl.d_wave16, sl.wave_consumed);
  if (rc) return rc;
  // packed lattices: capacity from the decoder's own per-channel limits would be far too large; start from 64 MB or
  // 1.5x the last batch's need and grow when the header reports an overflow (collect then falls back to the synchronous copy)
  const size_t want = std::max<size_t>({sl.pack_cap, p->pack_floor, (size_t)64 << 20, (size_t)n * ((size_t)512 << 10)});
  if (want > sl.pack_cap) {
    if (sl.d_pack) { B2K_CUDA_CHECK(cudaStreamSynchronize(p->st_copy)); cudaFree(sl.d_pack); sl.d_pack = nullptr; }
    B2K_CUDA_CHECK(cudaMalloc((void **)&sl.d_pack, want));
    sl.pack_cap = want;
  }
  // the buffer of this slot was last read by the D2H of batch k-2
  if (p->n_submitted >= 2) B2K_CUDA_CHECK(cudaStreamWaitEvent(p->st_compute, sl.pack_read, 0));
  rc = b2k_dec_pack_lattices_async(p->dec, p->d_channels, n, sl.d_pack, (int64_t)sl.pack_cap, (void *)p->st_compute);
  if (rc) return rc;
  B2K_CUDA_CHECK(cudaEventRecord(sl.packed, p->st_compute));
  sl.n = n; sl.busy = true;
  p->n_submitted++;
  return B2K_OK;
}

// int16 PCM already on the device (a shard received over NVLink, kaldi_b200/ingest.py): widen, run all stages.
extern "C" int b2k_pipeline_run_device_i16(b2k_pipeline *p, int32_t n, const int16_t *d_waves, void *stream) {
  if (!p || n <= 0 || n > p->cfg.max_batch || !d_waves) return set_error(B2K_ERR_INVALID, "b2k_pipeline_run_device_i16: bad args");
  if (reinterpret_cast<uintptr_t>(d_waves) & 15) return set_error(B2K_ERR_INVALID, "b2k_pipeline_run_device_i16: the waveform block must be 16-byte aligned");
  p->last_n = n;
  return run_device(p, n, stream, d_waves);
}

// The finalized lattices of batch slots 0..n-1 packed into a caller-owned device buffer (b2k_dec_pack_lattices_async):
// what a rank sends to the ingest rank.  *bytes_out (host) = header + body once the stream has run; a too small buffer
// is reported by the header's status (b2k_dec_unpack_lattices on a host copy returns B2K_ERR_OVERFLOW).
extern "C" int b2k_pipeline_pack_device(b2k_pipeline *p, int32_t n, void *d_buf, int64_t cap_bytes, void *stream) {
  if (!p || n <= 0 || n > p->cfg.max_batch || !d_buf) return set_error(B2K_ERR_INVALID, "b2k_pipeline_pack_device: bad args");
  if (!p->d_channels) {
    B2K_CUDA_CHECK(cudaMalloc((void **)&p->d_channels, 4 * (size_t)p->cfg.max_batch));
    B2K_CUDA_CHECK(cudaMemcpy(p->d_channels, p->channels.data(), 4 * (size_t)p->cfg.max_batch, cudaMemcpyHostToDevice));
  }
  return b2k_dec_pack_lattices_async(p->dec, p->d_channels, n, d_buf, cap_bytes, stream);
}

extern "C" int b2k_pipeline_collect(b2k_pipeline *p, int32_t *n_out, b2k_raw_lattice *view, const int64_t **state_offs,
                                    const int64_t **arc_offs, const int64_t **final_offs) {
  if (!p || !view) return set_error(B2K_ERR_INVALID, "b2k_pipeline_collect: bad args");
  if (p->n_collected >= p->n_submitted) return set_error(B2K_ERR_STATE, "b2k_pipeline_collect: nothing was submitted");
  b2k_pipeline::Slot &sl = p->slot[p->n_collected & 1];
  const int32_t n = sl.n;
  const size_t hb = (size_t)b2k_dec_pack_header_bytes(n);
  auto need_host = [&](size_t bytes) -> int {
    if (bytes <= sl.h_pack_cap) return B2K_OK;
    if (sl.h_pack) cudaFreeHost(sl.h_pack);
    sl.h_pack = nullptr; sl.h_pack_cap = 0;
    const size_t cap = bytes + bytes / 4;
    B2K_CUDA_CHECK(cudaMallocHost((void **)&sl.h_pack, cap));
    sl.h_pack_cap = cap;
    return B2K_OK;
  };
  int rc = need_host(std::max<size_t>(hb, (size_t)1 << 20));
  if (rc) return rc;
  // header first (sizes), then exactly the body; both on the copy stream, so the next batch's kernels keep running
  B2K_CUDA_CHECK(cudaStreamWaitEvent(p->st_copy, sl.packed, 0));
  B2K_CUDA_CHECK(cudaMemcpyAsync(sl.h_pack, sl.d_pack, hb, cudaMemcpyDeviceToHost, p->st_copy));
  B2K_CUDA_CHECK(cudaStreamSynchronize(p->st_copy));
  const int64_t *tail = (const int64_t *)sl.h_pack + 3 * ((size_t)n + 1);
  const int64_t status = tail[0], need = tail[1];
  if (status == B2K_ERR_OVERFLOW && need > (int64_t)sl.pack_cap) {
    // the packed form did not fit: grow the buffer for the next batches and repack this one (the channels still hold
    // it as long as no later batch was submitted after it; with one already running the lattices are gone)
    if (p->n_submitted - p->n_collected > 1) { sl.busy = false; p->n_collected++; sl.pack_cap = 0; cudaFree(sl.d_pack); sl.d_pack = nullptr;
      return set_error(B2K_ERR_OVERFLOW, "b2k_pipeline_collect: the packed lattices outgrew the buffer while the next batch was already running; the buffer grows for the following batches"); }
    B2K_CUDA_CHECK(cudaStreamSynchronize(p->st_compute));
    cudaFree(sl.d_pack); sl.d_pack = nullptr;
    const size_t want = (size_t)need + (size_t)need / 2;
    B2K_CUDA_CHECK(cudaMalloc((void **)&sl.d_pack, want));
    sl.pack_cap = want;
    rc = b2k_dec_pack_lattices_async(p->dec, p->d_channels, n, sl.d_pack, (int64_t)sl.pack_cap, (void *)p->st_compute);
    if (rc) return rc;
    B2K_CUDA_CHECK(cudaStreamSynchronize(p->st_compute));
    B2K_CUDA_CHECK(cudaMemcpy(sl.h_pack, sl.d_pack, hb, cudaMemcpyDeviceToHost));
  }
  {
    const int64_t *t2 = (const int64_t *)sl.h_pack + 3 * ((size_t)n + 1);
    if (t2[0] == B2K_OK) {
      const size_t bytes = (size_t)t2[1];
      const size_t cap_before = sl.h_pack_cap;
      if (bytes > cap_before) {                               // keep the header while the host buffer grows
        std::vector<char> keep(sl.h_pack, sl.h_pack + hb);
        if ((rc = need_host(bytes))) return rc;
        memcpy(sl.h_pack, keep.data(), hb);
      }
      B2K_CUDA_CHECK(cudaMemcpyAsync(sl.h_pack + hb, sl.d_pack + hb, bytes - hb, cudaMemcpyDeviceToHost, p->st_copy));
    }
    B2K_CUDA_CHECK(cudaEventRecord(sl.pack_read, p->st_copy));
    B2K_CUDA_CHECK(cudaStreamSynchronize(p->st_copy));
  }
  sl.busy = false;
  p->n_collected++;
  // (the device buffer grows on the next submit when a batch came within 2/3 of its capacity)
  {
    const int64_t need_now = ((const int64_t *)sl.h_pack)[3 * ((size_t)n + 1) + 1];
    if (need_now > 0 && (size_t)need_now * 3 / 2 > sl.pack_cap) p->pack_floor = std::max(p->pack_floor, (size_t)need_now * 2);
  }
  b2k_raw_lattice q;
  memset(&q, 0, sizeof(q));
  for (int k = 0; k < 3; k++) p->r_offs[k].resize((size_t)n + 1);
  rc = b2k_dec_unpack_lattices(sl.h_pack, n, &q, p->r_offs[0].data(), p->r_offs[1].data(), p->r_offs[2].data());
  if (rc) return rc;
  p->r_i32[0].resize(q.num_states); p->r_i32[1].resize(q.num_states); p->r_f32[0].resize(q.num_states); p->r_f32[1].resize(q.num_states);
  for (int k = 2; k < 6; k++) p->r_i32[k].resize(q.num_arcs);
  p->r_f32[2].resize(q.num_arcs); p->r_f32[3].resize(q.num_arcs);
  p->r_i32[6].resize(q.num_finals); p->r_f32[4].resize(q.num_finals);
  q.state_frame = p->r_i32[0].data(); q.state_hclg = p->r_i32[1].data(); q.state_tot_cost = p->r_f32[0].data(); q.state_extra_cost = p->r_f32[1].data();
  q.arc_src = p->r_i32[2].data(); q.arc_dst = p->r_i32[3].data(); q.arc_ilabel = p->r_i32[4].data(); q.arc_olabel = p->r_i32[5].data();
  q.arc_graph_cost = p->r_f32[2].data(); q.arc_acoustic_cost = p->r_f32[3].data();
  q.final_state = p->r_i32[6].data(); q.final_cost = p->r_f32[4].data();
  rc = b2k_dec_unpack_lattices(sl.h_pack, n, &q, nullptr, nullptr, nullptr);
  if (rc) return rc;
  *view = q;
  if (n_out) *n_out = n;
  if (state_offs) *state_offs = p->r_offs[0].data();
  if (arc_offs) *arc_offs = p->r_offs[1].data();
  if (final_offs) *final_offs = p->r_offs[2].data();
  return B2K_OK;
}

extern "C" {

int b2k_pipeline_decode_batch(b2k_pipeline *p, int32_t n, const float *const *h_waves, void *stream) {
  return decode_batch<float>(p, n, h_waves, stream);
}
int b2k_pipeline_decode_batch_i16(b2k_pipeline *p, int32_t n, const int16_t *const *h_waves, void *stream) {
  return decode_batch<int16_t>(p, n, h_waves, stream);
}

int b2k_pipeline_run_device(b2k_pipeline *p, int32_t n, const float *d_waves, void *stream) {
  if (!p || n <= 0 || n > p->cfg.max_batch) return set_error(B2K_ERR_INVALID, "b2k_pipeline_run_device: bad args");
  if (d_waves && d_waves != p->d_wave)
    B2K_CUDA_CHECK(cudaMemcpyAsync(p->d_wave, d_waves, 4 * (size_t)n * (size_t)p->cfg.num_samples, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
  p->last_n = n;
  return run_device(p, n, stream);
}

int b2k_pipeline_read(b2k_pipeline *p, int32_t what, int32_t n, float *h_out, void *stream) {
  if (!p || n <= 0 || n > p->cfg.max_batch || !h_out || what < 0 || what > 2) return set_error(B2K_ERR_INVALID, "b2k_pipeline_read: bad args");
  const b2k_pipeline_plan &pl = p->plan;
  const float *src = what == 0 ? p->d_feats : what == 1 ? p->d_ivec : p->d_loglikes;
  const size_t per = what == 0 ? (size_t)pl.num_feature_frames * pl.feat_dim
                   : what == 1 ? (size_t)pl.num_chunks * std::max(1, pl.ivector_dim) : (size_t)pl.num_output_frames * pl.num_pdfs;
  B2K_CUDA_CHECK(cudaMemcpyAsync(h_out, src, 4 * per * n, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
  B2K_CUDA_CHECK(cudaStreamSynchronize((cudaStream_t)stream));
  return B2K_OK;
}

int b2k_pipeline_get_raw_lattices(b2k_pipeline *p, int32_t n, b2k_raw_lattice *out, int64_t *state_offs, int64_t *arc_offs,
                                  int64_t *final_offs, void *stream) {
  if (!p || n <= 0 || n > p->cfg.max_batch) return set_error(B2K_ERR_INVALID, "b2k_pipeline_get_raw_lattices: bad args");
  return b2k_dec_get_raw_lattices(p->dec, p->channels.data(), n, out, state_offs, arc_offs, final_offs, stream);
}

}  // extern "C"
