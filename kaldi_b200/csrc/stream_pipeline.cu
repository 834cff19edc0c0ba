// stream_pipeline.cu — chunk-by-chunk decoding of many audio streams at once behind the C ABI: the call structure of
// cuda_decoder::BatchedThreadedNnet3CudaOnlinePipeline::DecodeBatch(corr_ids, wave_samples, is_first_chunk, is_last_chunk)
// (cudadecoder/batched-threaded-nnet3-cuda-online-pipeline.cc:316-377) over the stages of this library:
//
//   16-bit PCM of this call  -> H2D behind the channel's samples so far
//                            -> feat_kernel on the frames that became computable (reads the PCM itself; online == offline)
//                            -> b2k_nnet_stream_run_batch (BatchedStaticNnet3::RunBatch: context per channel, flush on the last chunk)
//                            -> b2k_dec_advance_decoding_frames on this call's output frames (and on the flushed ones)
//   last chunk               -> b2k_dec_finalize_decoding; the raw lattice / best path come from the decoder handle.
//
// Nothing of an utterance is recomputed when its next chunk arrives.  Host orchestration only; every kernel is an existing one.
#include <algorithm>
#include <vector>

#include "common.cuh"

using namespace b2k;

struct b2k_stream {
  b2k_stream_cfg cfg;
  b2k_feat *feat = nullptr;
  b2k_nnet_stream *nnet = nullptr;
  b2k_dec *dec = nullptr;
  int nch = 0, max_samples = 0, max_frames = 0, D = 0, P = 0, ivd = 0, opc = 0, fpc = 0;
  int16_t *d_wave = nullptr;     // [nch x max_samples]
  float *d_feats = nullptr;      // [nch x max_frames x D]
  float *d_zero_iv = nullptr;    // [max(1, ivd)]
  float *d_out = nullptr, *d_eos = nullptr;   // [nch x opc x P] each
  std::vector<int> samples, frames, out_frames;
  std::vector<char> started;
  // per call scratch
  std::vector<const int16_t *> wp; std::vector<float *> fp; std::vector<const float *> newp, ivp, llp;
  std::vector<int32_t> ns, ff, nf, n_new, n_out, n_eos, chan, nfr;
};

extern "C" {

void b2k_stream_cfg_default(b2k_stream_cfg *c) {
  if (!c) return;
  b2k_feat_cfg_default(&c->feat);
  b2k_dec_cfg_default(&c->dec);
  c->nchannels = 64; c->max_seconds = 60.0f; c->frames_per_chunk = 51; c->acoustic_scale = 1.0f; c->use_priors = 1;
}

int b2k_stream_destroy(b2k_stream *s) {
  if (!s) return B2K_OK;
  cudaDeviceSynchronize();
  if (s->dec) b2k_dec_destroy(s->dec);
  if (s->nnet) b2k_nnet_stream_destroy(s->nnet);
  if (s->feat) b2k_feat_destroy(s->feat);
  cudaFree(s->d_wave); cudaFree(s->d_feats); cudaFree(s->d_zero_iv); cudaFree(s->d_out); cudaFree(s->d_eos);
  delete s;
  return B2K_OK;
}

int b2k_stream_create(const b2k_stream_cfg *cfg, const b2k_model *model, const b2k_fst *fst, b2k_stream **out) {
  if (!cfg || !model || !fst || !out || cfg->nchannels <= 0 || cfg->max_seconds <= 0.0f || cfg->frames_per_chunk <= 0)
    return set_error(B2K_ERR_INVALID, "b2k_stream_create: bad args");
  *out = nullptr;
  int32_t mi[8];
  int rc = b2k_model_info(model, mi);
  if (rc) return rc;
  if (b2k_model_frame_subsampling_ambiguous(model))
    return set_error(B2K_ERR_INVALID, "b2k_stream_create: the model's layers do not decide --frame-subsampling-factor: b2k_model_set_frame_subsampling_factor first");
  if ((rc = require_device())) return rc;
  b2k_stream *s = new b2k_stream();
  s->cfg = *cfg;
  auto fail = [&](int code) { const std::string keep = g_last_error; b2k_stream_destroy(s); g_last_error = keep; return code; };
  b2k_feat_cfg fc = cfg->feat;
  fc.max_lanes = std::max(fc.max_lanes, cfg->nchannels);
  if ((rc = b2k_feat_create(&fc, &s->feat))) return fail(rc);
  s->nch = cfg->nchannels; s->D = b2k_feat_dim(s->feat); s->P = mi[2]; s->ivd = mi[1]; s->fpc = cfg->frames_per_chunk;
  if (s->D != mi[0]) { set_error(B2K_ERR_INVALID, "b2k_stream_create: the feature dimension differs from the model's input dimension"); return fail(B2K_ERR_INVALID); }
  s->max_samples = (int)(cfg->max_seconds * fc.samp_freq);
  s->max_frames = b2k_feat_num_frames(s->feat, s->max_samples, 1) + 1;
  b2k_nnet_compile_cfg cc;
  cc.feat_dim = mi[0]; cc.ivector_dim = mi[1]; cc.num_pdfs = mi[2]; cc.frame_subsampling_factor = mi[3]; cc.num_frames = 0;
  cc.frames_per_chunk = cfg->frames_per_chunk; cc.use_priors = cfg->use_priors; cc.conv_dense = 0; cc.acoustic_scale = cfg->acoustic_scale;
  if ((rc = b2k_nnet_stream_create(&cc, b2k_model_layers(model), mi[4], b2k_model_weights(model), mi[5], s->nch, s->nch, 0, &s->nnet))) return fail(rc);
  int64_t si[8];
  b2k_nnet_stream_info(s->nnet, si);
  s->opc = (int)si[0];
  b2k_dec_cfg dc = cfg->dec;
  const int sub = mi[3];
  // every call rounds its own output count up (batched-static-nnet3.cc:181-186): one spare frame per possible call
  const int need_frames = (s->max_frames + sub - 1) / sub + s->max_frames / std::max(1, cfg->frames_per_chunk) + 8;
  if (dc.max_frames < need_frames) dc.max_frames = need_frames;
  if ((rc = b2k_dec_create(fst, &dc, s->nch, s->nch, &s->dec))) return fail(rc);
  cudaError_t e = cudaMalloc(&s->d_wave, sizeof(int16_t) * (size_t)s->nch * s->max_samples + 16);
  if (e == cudaSuccess) e = cudaMalloc(&s->d_feats, sizeof(float) * (size_t)s->nch * s->max_frames * s->D);
  if (e == cudaSuccess) e = cudaMalloc(&s->d_zero_iv, sizeof(float) * std::max(1, s->ivd));
  if (e == cudaSuccess) e = cudaMemset(s->d_zero_iv, 0, sizeof(float) * std::max(1, s->ivd));
  if (e == cudaSuccess) e = cudaMalloc(&s->d_out, sizeof(float) * (size_t)s->nch * s->opc * s->P);
  if (e == cudaSuccess) e = cudaMalloc(&s->d_eos, sizeof(float) * (size_t)s->nch * s->opc * s->P);
  if (e != cudaSuccess) { set_error(B2K_ERR_CUDA, "b2k_stream_create", cudaGetErrorString(e)); return fail(B2K_ERR_CUDA); }
  s->samples.assign(s->nch, 0); s->frames.assign(s->nch, 0); s->out_frames.assign(s->nch, 0); s->started.assign(s->nch, 0);
  *out = s;
  return B2K_OK;
}

b2k_dec *b2k_stream_decoder(b2k_stream *s) { return s ? s->dec : nullptr; }
const float *b2k_stream_features(const b2k_stream *s, int32_t channel) {
  return (s && channel >= 0 && channel < s->nch) ? s->d_feats + (size_t)channel * s->max_frames * s->D : nullptr;
}
int b2k_stream_info(const b2k_stream *s, int64_t info[8]) {
  if (!s || !info) return set_error(B2K_ERR_INVALID, "b2k_stream_info: bad args");
  info[0] = s->nch; info[1] = s->max_samples; info[2] = s->max_frames; info[3] = s->D; info[4] = s->P; info[5] = s->ivd; info[6] = s->opc;
  info[7] = s->fpc;
  return B2K_OK;
}

int b2k_stream_decode_batch_i16(b2k_stream *s, int32_t n, const int32_t *channels, const int16_t *const *h_chunks,
                                const int32_t *num_samples, const int32_t *is_first_chunk, const int32_t *is_last_chunk,
                                const float *const *d_ivectors, int32_t *new_output_frames, int32_t *output_frames_so_far,
                                const float **d_new_frames, const float **d_flushed_frames, void *stream) {
  if (!s || n <= 0 || n > s->nch || !channels || !h_chunks || !num_samples || !is_first_chunk || !is_last_chunk)
    return set_error(B2K_ERR_INVALID, "b2k_stream_decode_batch_i16: bad args");
  cudaStream_t st = (cudaStream_t)stream;
  // validate everything before any state changes
  for (int i = 0; i < n; i++) {
    const int ch = channels[i];
    if (ch < 0 || ch >= s->nch) return set_error(B2K_ERR_INVALID, "b2k_stream_decode_batch_i16: channel out of range");
    for (int j = 0; j < i; j++) if (channels[j] == ch) return set_error(B2K_ERR_INVALID, "b2k_stream_decode_batch_i16: a channel appears twice in the batch");
    if (num_samples[i] < 0 || (num_samples[i] > 0 && !h_chunks[i])) return set_error(B2K_ERR_INVALID, "b2k_stream_decode_batch_i16: bad chunk");
    if (!is_first_chunk[i] && !s->started[ch]) return set_error(B2K_ERR_STATE, "b2k_stream_decode_batch_i16: a channel's first call must have is_first_chunk set");
    const long long have = (is_first_chunk[i] ? 0 : s->samples[ch]) + (long long)num_samples[i];
    if (have > s->max_samples) return set_error(B2K_ERR_OVERFLOW, "b2k_stream_decode_batch_i16: stream longer than max_seconds");
    const int ready = have > 0 ? b2k_feat_num_frames(s->feat, have, is_last_chunk[i] ? 1 : 0) : 0;
    const int done = is_first_chunk[i] ? 0 : s->frames[ch];
    if (ready - done > s->fpc) return set_error(B2K_ERR_INVALID, "b2k_stream_decode_batch_i16: a call may bring at most frames_per_chunk new frames");
  }
  s->wp.clear(); s->fp.clear(); s->ns.clear(); s->ff.clear(); s->nf.clear();
  s->newp.resize(n); s->ivp.resize(n); s->n_new.resize(n); s->n_out.resize(n); s->n_eos.resize(n);
  std::vector<int32_t> fresh;
  for (int i = 0; i < n; i++) {
    const int ch = channels[i];
    if (is_first_chunk[i]) { s->samples[ch] = s->frames[ch] = s->out_frames[ch] = 0; s->started[ch] = 1; fresh.push_back(ch); }
    int16_t *w = s->d_wave + (size_t)ch * s->max_samples;
    if (num_samples[i] > 0)
      B2K_CUDA_CHECK(cudaMemcpyAsync(w + s->samples[ch], h_chunks[i], sizeof(int16_t) * (size_t)num_samples[i], cudaMemcpyHostToDevice, st));
    s->samples[ch] += num_samples[i];
    const int ready = s->samples[ch] > 0 ? b2k_feat_num_frames(s->feat, s->samples[ch], is_last_chunk[i] ? 1 : 0) : 0;
    const int k = std::max(0, ready - s->frames[ch]);
    float *f = s->d_feats + (size_t)ch * s->max_frames * s->D;
    s->n_new[i] = k;
    s->newp[i] = f + (size_t)s->frames[ch] * s->D;
    s->ivp[i] = (d_ivectors && d_ivectors[i]) ? d_ivectors[i] : s->d_zero_iv;
    if (k > 0) { s->wp.push_back(w); s->fp.push_back(f); s->ns.push_back(s->samples[ch]); s->ff.push_back(s->frames[ch]); s->nf.push_back(k); }
  }
  int rc;
  if (!s->wp.empty()) {
    rc = b2k_feat_compute_batched_i16(s->feat, (int32_t)s->wp.size(), s->wp.data(), s->ns.data(), s->ff.data(), s->nf.data(), s->fp.data(), s->D, stream);
    if (rc) return rc;
  }
  rc = b2k_nnet_stream_run_batch(s->nnet, n, channels, s->newp.data(), s->D, s->ivd > 0 ? s->ivp.data() : nullptr, s->n_new.data(),
                                 is_first_chunk, is_last_chunk, s->d_out, s->d_eos, s->P, s->n_out.data(), s->n_eos.data(), stream);
  if (rc) return rc;
  for (int i = 0; i < n; i++) s->frames[channels[i]] += s->n_new[i];
  if (!fresh.empty() && (rc = b2k_dec_init_decoding(s->dec, fresh.data(), (int32_t)fresh.size(), stream))) return rc;
  for (int pass = 0; pass < 2; pass++) {                      // the frames of the chunks, then the flushed frames of the streams that ended
    const std::vector<int32_t> &cnt = pass == 0 ? s->n_out : s->n_eos;
    const float *base = pass == 0 ? s->d_out : s->d_eos;
    s->chan.clear(); s->llp.clear(); s->nfr.clear();
    for (int i = 0; i < n; i++)
      if (cnt[i] > 0) { s->chan.push_back(channels[i]); s->llp.push_back(base + (size_t)i * s->opc * s->P); s->nfr.push_back(cnt[i]); }
    if (!s->chan.empty() &&
        (rc = b2k_dec_advance_decoding_frames(s->dec, s->chan.data(), s->llp.data(), s->nfr.data(), s->P, (int32_t)s->chan.size(), stream))) return rc;
  }
  s->chan.clear();
  for (int i = 0; i < n; i++) {
    const int ch = channels[i];
    s->out_frames[ch] += s->n_out[i] + s->n_eos[i];
    if (new_output_frames) new_output_frames[i] = s->n_out[i] + s->n_eos[i];
    if (output_frames_so_far) output_frames_so_far[i] = s->out_frames[ch];
    if (d_new_frames) d_new_frames[i] = s->n_out[i] > 0 ? s->d_out + (size_t)i * s->opc * s->P : nullptr;
    if (d_flushed_frames) d_flushed_frames[i] = s->n_eos[i] > 0 ? s->d_eos + (size_t)i * s->opc * s->P : nullptr;
    if (is_last_chunk[i]) { s->chan.push_back(ch); s->started[ch] = 0; }
  }
  if (!s->chan.empty() && (rc = b2k_dec_finalize_decoding(s->dec, s->chan.data(), (int32_t)s->chan.size(), stream))) return rc;
  return B2K_OK;
}

}  // extern "C"
