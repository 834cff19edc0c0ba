// nnet_stream.cu — chunked nnet3 forward with carried context: the B200 counterpart of cuda_decoder::BatchedStaticNnet3
// (cudadecoder/batched-static-nnet3.{h,cc}, batched-static-nnet3-kernels.cu).  A window program (b2k_nnet_compile_window) is
// compiled once for frames_per_chunk + left + right input frames; every RunBatch gathers [context | new frames] into one window
// per batch slot, runs the same tcgen05 executor as the whole-utterance path on the batch of windows, and keeps the last
// left+right frames of every channel in HBM for its next chunk.
#include <algorithm>
#include <vector>

#include "common.cuh"

namespace b2k {

struct SlotAsg {
  const float *feat;       // new frames of this chunk (null for the flush)
  int channel;
  int in_ctx;              // frames the channel's context holds
  int n_new;
};

// One CTA per (window row, batch slot).  Rows past the frames the slot has are zeroed so that the rows of the output that the
// caller is told to ignore are at least deterministic.  (batched-static-nnet3-kernels.cu:29-94, :103-152)
__global__ void __launch_bounds__(128) nnet_stream_build_kernel(const SlotAsg *asg, const float *ctx, long long ctx_ch_stride, float *win,
                                                                int W, int dim, int L, int R, int feat_stride, int flush) {
  const SlotAsg a = asg[blockIdx.y];
  const int row = blockIdx.x;
  float *dst = win + ((long long)blockIdx.y * W + row) * dim;
  const float *c = ctx + (long long)a.channel * ctx_ch_stride;
  const float *src = nullptr;
  if (flush) {
    if (a.in_ctx > 0) {
      if (row < a.in_ctx) src = c + (long long)row * dim;
      else if (row < a.in_ctx + R) src = c + (long long)(a.in_ctx - 1) * dim;      // right context = the last frame
    }
  } else {
    const int n_left0 = a.in_ctx == 0 ? L : 0;                                      // first chunk: left context = frame 0
    if (row < n_left0) { if (a.n_new > 0) src = a.feat; }
    else if (row < n_left0 + a.in_ctx) src = c + (long long)(row - n_left0) * dim;
    else if (row < n_left0 + a.in_ctx + a.n_new) src = a.feat + (long long)(row - n_left0 - a.in_ctx) * feat_stride;
  }
  for (int i = threadIdx.x; i < dim; i += blockDim.x) dst[i] = src ? src[i] : 0.0f;
}

// The last min(total, L+R) rows of the slot's window become the channel's context (batched-static-nnet3-kernels.cu:161-205).
__global__ void __launch_bounds__(128) nnet_stream_save_kernel(const SlotAsg *asg, float *ctx, long long ctx_ch_stride, const float *win,
                                                               int W, int dim, int L, int R) {
  const SlotAsg a = asg[blockIdx.y];
  const int total = a.in_ctx + a.n_new + (a.in_ctx == 0 ? L : 0);
  const int n_copy = min(total, L + R);
  const int row = blockIdx.x;
  if (row >= n_copy) return;
  const float *src = win + ((long long)blockIdx.y * W + (total - n_copy + row)) * dim;
  float *dst = ctx + (long long)a.channel * ctx_ch_stride + (long long)row * dim;
  for (int i = threadIdx.x; i < dim; i += blockDim.x) dst[i] = src[i];
}

}  // namespace b2k

using namespace b2k;

struct b2k_nnet_stream {
  b2k_nnet *nn = nullptr;
  int L = 0, R = 0, sub = 1, fpc = 0, W = 0, opc = 0, in_dim = 0, iv_dim = 0, out_dim = 0;
  int max_batch = 0, nchannels = 0, iv_rows = 1;
  float *d_ctx = nullptr, *d_win = nullptr;
  SlotAsg *d_asg = nullptr, *h_asg = nullptr;      // 2 * max_batch entries: the main pass, then the flush
  cudaEvent_t staging_free = nullptr;
  std::vector<int> in_ctx;                         // per channel, -1 = never initialised (batched-static-nnet3.cc:93)
  std::vector<const float *> p_in, p_iv;
  std::vector<float *> p_out;
  std::vector<int> eos_slot;
};

extern "C" {

int b2k_nnet_stream_account(int32_t L, int32_t R, int32_t sub, int32_t in_ctx, int32_t n_new, int32_t flush, int32_t *after, int32_t *n_out) {
  if (L < 0 || R < 0 || sub <= 0 || in_ctx < 0 || n_new < 0 || !after || !n_out || (flush && n_new != 0))
    return set_error(B2K_ERR_INVALID, "b2k_nnet_stream_account: bad args");
  int in_batch = n_new + (in_ctx == 0 ? L : in_ctx);       // :166-171
  if (flush) in_batch += R;
  *after = std::min(in_batch, L + R);                       // :177-178
  const int minus_ctx = std::max(0, in_batch - (L + R));    // :181-186
  *n_out = (minus_ctx + sub - 1) / sub;
  return B2K_OK;
}

int b2k_nnet_stream_create(const b2k_nnet_compile_cfg *cfgp, const b2k_nnet_layer *layers, int32_t n_layers, const b2k_nnet_weight *weights,
                           int32_t n_weights, int32_t max_batch, int32_t nchannels, int32_t looped, b2k_nnet_stream **out) {
  if (!cfgp || !layers || !out || max_batch <= 0) return set_error(B2K_ERR_INVALID, "b2k_nnet_stream_create: bad args");
  if (nchannels < 0) nchannels = max_batch;
  if (nchannels < max_batch) return set_error(B2K_ERR_INVALID, "b2k_nnet_stream_create: nchannels < max_batch");   // batched-static-nnet3.h:62
  b2k_nnet_compile_cfg cfg = *cfgp;
  const int sub = cfg.frame_subsampling_factor, fpc = cfg.frames_per_chunk;
  if (sub <= 0 || fpc <= 0) return set_error(B2K_ERR_INVALID, "b2k_nnet_stream_create: frames_per_chunk and the subsampling factor must be positive");
  int L = 0, R = 0;
  cfg.num_frames = 1; cfg.frames_per_chunk = sub;
  int rc = b2k_nnet_model_context(&cfg, layers, n_layers, &L, &R);
  if (rc) return rc;
  // (the looped schedule has no such limit: its first right-context frames may arrive over several calls, none of which has output yet)
  if (!looped && fpc < R) return set_error(B2K_ERR_INVALID, "Please set --frames-per-chunk at least as large as the neural net right context");   // :172-175
  const int opc = (sub - 1 + fpc) / sub, W = fpc + L + R;
  cfg.num_frames = W;
  int iv_rows = 1;
  if (looped) {                                              // the looped computation's i-vector arithmetic: C = the chunk
    if (fpc % sub != 0) return set_error(B2K_ERR_INVALID, "b2k_nnet_stream_create: looped chunks must be a multiple of the subsampling factor");
    cfg.frames_per_chunk = fpc;
    if (cfg.ivector_dim > 0 && (rc = b2k_nnet_looped_ivector_rows(&cfg, layers, n_layers, &iv_rows))) return rc;
  }
  b2k_nnet_program *prog = nullptr;
  rc = b2k_nnet_compile_window(&cfg, L, opc, iv_rows, layers, n_layers, weights, n_weights, &prog);
  if (rc) return rc;
  if ((rc = require_device())) { b2k_nnet_program_destroy(prog); return rc; }
  b2k_nnet_stream *s = new b2k_nnet_stream();
  rc = b2k_nnet_create_from_program(prog, max_batch, &s->nn);
  b2k_nnet_program_destroy(prog);
  if (rc) { delete s; return rc; }
  s->L = L; s->R = R; s->sub = sub; s->fpc = fpc; s->W = W; s->opc = opc;
  s->in_dim = cfg.feat_dim; s->iv_dim = cfg.ivector_dim; s->out_dim = b2k_nnet_output_dim(s->nn);
  s->max_batch = max_batch; s->nchannels = nchannels; s->iv_rows = iv_rows;
  s->in_ctx.assign(nchannels, -1);
  const size_t ctx_floats = (size_t)nchannels * std::max(1, L + R) * s->in_dim, win_floats = (size_t)max_batch * W * s->in_dim;
  cudaError_t e = cudaMalloc(&s->d_ctx, ctx_floats * sizeof(float));
  if (e == cudaSuccess) e = cudaMalloc(&s->d_win, win_floats * sizeof(float));
  if (e == cudaSuccess) e = cudaMalloc(&s->d_asg, 2 * (size_t)max_batch * sizeof(SlotAsg));
  if (e == cudaSuccess) e = cudaMallocHost(&s->h_asg, 2 * (size_t)max_batch * sizeof(SlotAsg));
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&s->staging_free, cudaEventDisableTiming);
  if (e == cudaSuccess) e = cudaMemset(s->d_ctx, 0, ctx_floats * sizeof(float));
  if (e != cudaSuccess) { b2k_nnet_stream_destroy(s); return set_error(B2K_ERR_CUDA, "b2k_nnet_stream_create", cudaGetErrorString(e)); }
  *out = s;
  return B2K_OK;
}

int b2k_nnet_stream_destroy(b2k_nnet_stream *s) {
  if (!s) return B2K_OK;
  cudaDeviceSynchronize();
  if (s->nn) b2k_nnet_destroy(s->nn);
  cudaFree(s->d_ctx); cudaFree(s->d_win); cudaFree(s->d_asg);
  if (s->h_asg) cudaFreeHost(s->h_asg);
  if (s->staging_free) cudaEventDestroy(s->staging_free);
  delete s;
  return B2K_OK;
}

int b2k_nnet_stream_info(const b2k_nnet_stream *s, int64_t info[8]) {
  if (!s || !info) return set_error(B2K_ERR_INVALID, "b2k_nnet_stream_info: bad args");
  info[0] = s->opc; info[1] = s->L; info[2] = s->R; info[3] = s->W; info[4] = s->in_dim; info[5] = s->iv_dim; info[6] = s->out_dim;
  info[7] = s->iv_rows;
  return B2K_OK;
}

int b2k_nnet_stream_run_batch(b2k_nnet_stream *s, int32_t n, const int32_t *channels, const float *const *d_features, int32_t features_stride,
                              const float *const *d_ivectors, const int32_t *n_valid, const int32_t *is_first, const int32_t *is_last,
                              float *d_log_post, float *d_eos_log_post, int32_t out_stride, int32_t *n_out, int32_t *n_eos_out, void *stream) {
  if (!s || n <= 0 || n > s->max_batch || !channels || !d_features || !n_valid || !is_first || !is_last || !d_log_post || !n_out ||
      !n_eos_out || out_stride < s->out_dim || features_stride < s->in_dim)
    return set_error(B2K_ERR_INVALID, "b2k_nnet_stream_run_batch: bad args");
  if (s->iv_dim > 0 && !d_ivectors) return set_error(B2K_ERR_INVALID, "Neural net expects iVectors but none provided");
  // validate before any state changes
  bool any_last = false;
  for (int i = 0; i < n; i++) {
    const int ch = channels[i];
    if (ch < 0 || ch >= s->nchannels) return set_error(B2K_ERR_INVALID, "b2k_nnet_stream_run_batch: channel out of range");
    for (int j = 0; j < i; j++) if (channels[j] == ch) return set_error(B2K_ERR_INVALID, "b2k_nnet_stream_run_batch: a channel appears twice in the batch");
    if (n_valid[i] < 0 || n_valid[i] > s->fpc) return set_error(B2K_ERR_INVALID, "b2k_nnet_stream_run_batch: more input frames than frames_per_chunk");   // :158
    if (n_valid[i] > 0 && !d_features[i]) return set_error(B2K_ERR_INVALID, "b2k_nnet_stream_run_batch: null feature pointer");
    if (s->iv_dim > 0 && !d_ivectors[i]) return set_error(B2K_ERR_INVALID, "b2k_nnet_stream_run_batch: null i-vector pointer");
    if (!is_first[i] && s->in_ctx[ch] < 0) return set_error(B2K_ERR_STATE, "b2k_nnet_stream_run_batch: a channel's first call must have is_first_chunk set");
    if (is_last[i]) any_last = true;
  }
  if (any_last && !d_eos_log_post) return set_error(B2K_ERR_INVALID, "b2k_nnet_stream_run_batch: a last chunk needs d_eos_log_post");
  cudaStream_t st = (cudaStream_t)stream;
  B2K_CUDA_CHECK(cudaEventSynchronize(s->staging_free));            // the pinned assignments of the previous call have been read
  for (int i = 0; i < n; i++) if (is_first[i]) s->in_ctx[channels[i]] = 0;                 // InitChannel (:311-313)
  s->p_in.resize(n); s->p_iv.resize(n); s->p_out.resize(n);
  s->eos_slot.clear();
  // step 1: the chunks (:323-333)
  for (int i = 0; i < n; i++) {
    const int ch = channels[i];
    SlotAsg &a = s->h_asg[i];
    a.feat = d_features[i]; a.channel = ch; a.in_ctx = s->in_ctx[ch]; a.n_new = n_valid[i];
    int after = 0, no = 0;
    b2k_nnet_stream_account(s->L, s->R, s->sub, a.in_ctx, a.n_new, 0, &after, &no);
    s->in_ctx[ch] = after; n_out[i] = no; n_eos_out[i] = 0;
    if (is_last[i]) s->eos_slot.push_back(i);
    s->p_in[i] = s->d_win + (size_t)i * s->W * s->in_dim;
    s->p_iv[i] = s->iv_dim > 0 ? d_ivectors[i] : nullptr;
    s->p_out[i] = d_log_post + (size_t)i * s->opc * out_stride;
  }
  // step 2: the flush of the channels that ended (:339-366); their context is the one step 1 saves
  const int n_eos = (int)s->eos_slot.size();
  for (int j = 0; j < n_eos; j++) {
    const int i = s->eos_slot[j], ch = channels[i];
    SlotAsg &a = s->h_asg[s->max_batch + j];
    a.feat = nullptr; a.channel = ch; a.in_ctx = s->in_ctx[ch]; a.n_new = 0;
    int after = 0, no = 0;
    b2k_nnet_stream_account(s->L, s->R, s->sub, a.in_ctx, 0, 1, &after, &no);
    s->in_ctx[ch] = after; n_eos_out[i] = no;
  }
  B2K_CUDA_CHECK(cudaMemcpyAsync(s->d_asg, s->h_asg, sizeof(SlotAsg) * n, cudaMemcpyHostToDevice, st));
  if (n_eos) B2K_CUDA_CHECK(cudaMemcpyAsync(s->d_asg + s->max_batch, s->h_asg + s->max_batch, sizeof(SlotAsg) * n_eos, cudaMemcpyHostToDevice, st));
  B2K_CUDA_CHECK(cudaEventRecord(s->staging_free, st));
  const long long ctx_ch_stride = (long long)std::max(1, s->L + s->R) * s->in_dim;
  nnet_stream_build_kernel<<<dim3(s->W, n), 128, 0, st>>>(s->d_asg, s->d_ctx, ctx_ch_stride, s->d_win, s->W, s->in_dim, s->L, s->R, features_stride, 0);
  B2K_LAUNCH_CHECK();
  if (s->L + s->R > 0) {
    nnet_stream_save_kernel<<<dim3(s->L + s->R, n), 128, 0, st>>>(s->d_asg, s->d_ctx, ctx_ch_stride, s->d_win, s->W, s->in_dim, s->L, s->R);
    B2K_LAUNCH_CHECK();
  }
  int rc = b2k_nnet_run(s->nn, n, s->p_in.data(), s->in_dim, s->iv_dim > 0 ? s->p_iv.data() : nullptr, s->iv_dim, s->p_out.data(), out_stride, stream);
  if (rc) return rc;
  if (n_eos) {
    for (int j = 0; j < n_eos; j++) {
      const int i = s->eos_slot[j];
      s->p_in[j] = s->d_win + (size_t)j * s->W * s->in_dim;
      s->p_iv[j] = s->iv_dim > 0 ? d_ivectors[i] : nullptr;
      s->p_out[j] = d_eos_log_post + (size_t)i * s->opc * out_stride;
    }
    nnet_stream_build_kernel<<<dim3(s->W, n_eos), 128, 0, st>>>(s->d_asg + s->max_batch, s->d_ctx, ctx_ch_stride, s->d_win, s->W, s->in_dim, s->L, s->R, 0, 1);
    B2K_LAUNCH_CHECK();
    rc = b2k_nnet_run(s->nn, n_eos, s->p_in.data(), s->in_dim, s->iv_dim > 0 ? s->p_iv.data() : nullptr, s->iv_dim, s->p_out.data(), out_stride, stream);
    if (rc) return rc;
  }
  return B2K_OK;
}

}  // extern "C"
