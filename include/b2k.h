/* b2k.h — C ABI of the B200-native online2 hot path (features -> nnet3 -> decoder).
 *
 * This is the drop-in boundary: the reference has no FFI, its
 * boundary is a set of C++ class surfaces; each entry point below names the
 * reference member function it replaces (file:line under /root/reference/src).
 * The header-compatible C++ shims in kaldi_b200/host/ forward to these calls;
 * INTEGRATION.md shows the binding a Kaldi maintainer would add.
 *
 * Conventions: every function returns a b2k_status (0 = OK); no exceptions
 * cross the ABI; handles are opaque; the caller owns every buffer it passes;
 * pointers named d_* are CUDA device pointers, h_* / unprefixed are host
 * pointers; `stream` is a cudaStream_t passed as void* (NULL = default stream).
 * There is no CPU fallback: every compute entry point fails with
 * B2K_ERR_NO_DEVICE when no sm_100 device is present.
 */
#ifndef B2K_H_
#define B2K_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
  B2K_OK = 0,
  B2K_ERR_INVALID = 1,      /* bad argument                                   */
  B2K_ERR_NO_DEVICE = 2,    /* no CUDA device / wrong arch                    */
  B2K_ERR_CUDA = 3,         /* a CUDA runtime call failed (see b2k_last_error)*/
  B2K_ERR_OVERFLOW = 4,     /* a decoder arena/queue capacity was exceeded    */
  B2K_ERR_STATE = 5         /* call sequence error (e.g. advance before init) */
} b2k_status;

const char *b2k_last_error(void);
int b2k_version(void);
/* number of kernels this library has launched since load (bench "gpu_launches") */
int64_t b2k_kernel_launch_count(void);

/* ------------------------------------------------------------------ decoder */

/* Host-side CSR view of a decoding graph in fst::ConstFst<StdArc> order.
 * Replaces the argument of CudaFst::CudaFst(const fst::StdFst&, const
 * TransitionInformation*) (cudadecoder/cuda-fst.h:75-82, cuda-fst.cc:34-198). */
typedef struct {
  int32_t num_states;
  int32_t start;
  const int32_t *offsets;    /* [num_states+1] arc offsets                    */
  const int32_t *ilabel;     /* transition-ids, 0 = epsilon                   */
  const int32_t *olabel;
  const float *weight;       /* tropical weights (costs)                      */
  const int32_t *nextstate;
  const float *final_cost;   /* [num_states], +inf = not final                */
  const int32_t *tid2pdf;    /* [num_tids] TransitionIdToPdfFast; may be NULL
                                (then pdf = ilabel - 1)                       */
  int32_t num_tids;
} b2k_fst_csr;

typedef struct b2k_fst b2k_fst;

/* CudaFst::CudaFst / Initialize (cuda-fst.cc:34,57-198): builds the device CSR
 * (emitting arcs and epsilon arcs in separate arrays, tid->pdf pre-applied). */
int b2k_fst_create(const b2k_fst_csr *csr, b2k_fst **out);
int b2k_fst_destroy(b2k_fst *fst);
int32_t b2k_fst_num_states(const b2k_fst *fst);   /* CudaFst::NumStates  cuda-fst.h:81 */
int32_t b2k_fst_start(const b2k_fst *fst);        /* CudaFst::Start      cuda-fst.h:82 */

/* Decoder options: the union of LatticeFasterDecoderConfig
 * (decoder/lattice-faster-decoder.h:38-106; the parity semantics) and the
 * capacity knobs of CudaDecoderConfig (cudadecoder/cuda-decoder.h:58-163). */
typedef struct {
  float beam;                 /* --beam            (15)                       */
  float lattice_beam;         /* --lattice-beam    (8)                        */
  int32_t max_active;         /* --max-active      (7000)                     */
  int32_t min_active;         /* --min-active      (200)                      */
  float beam_delta;           /* --beam-delta      (0.5)                      */
  int32_t prune_interval;     /* --prune-interval  (25); reserved             */
  float prune_scale;          /* --prune-scale     (0.1); reserved            */
  int32_t max_tokens_per_frame; /* per-lane per-frame token capacity (pow2/2) */
  int32_t max_frames;         /* per-channel frame capacity                   */
  int64_t max_tokens;         /* per-channel token arena                      */
  int64_t max_links;          /* per-channel forward-link arena               */
  int32_t reference_order;    /* 1 (default): bit-exact emulation of the CPU
                                 decoder's HashList iteration order (running
                                 next_cutoff, lattice-faster-decoder.cc:794-796);
                                 0: order-free fast mode (admission against the
                                 final cutoff; NOT bit-compatible with the
                                 reference when extras matter)  */
  float hash_ratio;           /* --hash-ratio (2.0); reference_order only      */
  int32_t max_arcs_per_frame; /* reference_order: arc-position capacity        */
  int32_t max_lattice_states; /* per-channel capacity of the finalized lattice */
  int32_t max_lattice_arcs;
} b2k_dec_cfg;

void b2k_dec_cfg_default(b2k_dec_cfg *cfg);

typedef struct b2k_dec b2k_dec;

/* CudaDecoder::CudaDecoder(const CudaFst&, const CudaDecoderConfig&, int32
 * nlanes, int32 nchannels) (cuda-decoder.h:224-225).  The fst is not owned. */
int b2k_dec_create(const b2k_fst *fst, const b2k_dec_cfg *cfg, int32_t nlanes,
                   int32_t nchannels, b2k_dec **out);
int b2k_dec_destroy(b2k_dec *dec);

/* CudaDecoder::InitDecoding(const std::vector<ChannelId>&) (cuda-decoder.h:241;
 * CPU semantics: LatticeFasterDecoderTpl::InitDecoding lattice-faster-decoder.cc:63-81). */
int b2k_dec_init_decoding(b2k_dec *dec, const int32_t *channels, int32_t n, void *stream);

/* CudaDecoder::AdvanceDecoding(const std::vector<std::pair<ChannelId,const
 * BaseFloat*>>&) (cuda-decoder.h:264-265): one frame per listed channel;
 * d_loglikes[i] points at that frame's [num_pdfs] row (already -log prior,
 * x acoustic_scale).  CPU semantics: one iteration of AdvanceDecoding's loop,
 * lattice-faster-decoder.cc:621-627. */
int b2k_dec_advance_decoding(b2k_dec *dec, const int32_t *channels,
                             const float *const *d_loglikes, int32_t n, void *stream);

/* Batched multi-frame form (B200-native: one persistent CTA per lane loops
 * over the frames, no host round trip per frame).  Lane i decodes
 * num_frames[i] frames whose rows are d_loglikes[i] + f*row_stride. */
int b2k_dec_advance_decoding_frames(b2k_dec *dec, const int32_t *channels,
                                    const float *const *d_loglikes,
                                    const int32_t *num_frames, int32_t row_stride,
                                    int32_t n, void *stream);

/* CudaDecoder::NumFramesDecoded(ChannelId) (cuda-decoder.h). Synchronizes. */
int b2k_dec_num_frames_decoded(b2k_dec *dec, int32_t channel, int32_t *out);

/* LatticeFasterDecoderTpl::FinalizeDecoding (lattice-faster-decoder.cc:634-649):
 * final-cost aware backward pruning of tokens and forward links, on the GPU. */
int b2k_dec_finalize_decoding(b2k_dec *dec, const int32_t *channels, int32_t n, void *stream);

/* Per-channel counters after a synchronize: [0] status (b2k_status), [1] frames
 * decoded, [2] tokens in arena, [3] links in arena, [4] emitting arcs examined,
 * [5] epsilon arcs examined by the parallel closure, [6] surviving lattice
 * states, [7] surviving lattice arcs, [8] number of final states, [9] finalized,
 * [10] any final state reached, [11] source line that raised the channel's
 * first error (diagnostics for capacity overflows); [16..27] reference-order
 * kernel cycle counters per phase, [28..30] eps-replay counters (pops, arc
 * visits, frames walked in shared memory), [31] total kernel cycles. */
int b2k_dec_channel_info(b2k_dec *dec, int32_t channel, int64_t info[32]);

/* Raw lattice of a finalized channel — the content of
 * LatticeFasterDecoderTpl::GetRawLattice (lattice-faster-decoder.cc:114-197) /
 * CudaDecoder::GetRawLattice (cuda-decoder.h) as flat host arrays.  Lattice
 * states are surviving tokens; arcs are surviving forward links with
 * LatticeWeight(graph_cost, acoustic_cost - cost_offset[frame]).
 * Call once with all pointers NULL to obtain sizes via b2k_dec_channel_info. */
typedef struct {
  int64_t num_states, num_arcs, num_finals;
  int32_t *state_frame;    /* [num_states] frame index (0 = before 1st frame) */
  int32_t *state_hclg;     /* [num_states] HCLG state of the token            */
  float *state_tot_cost;   /* [num_states]                                    */
  float *state_extra_cost; /* [num_states]                                    */
  int32_t *arc_src;        /* [num_arcs] lattice-state index                  */
  int32_t *arc_dst;
  int32_t *arc_ilabel;     /* transition-id (0 for epsilon)                   */
  int32_t *arc_olabel;
  float *arc_graph_cost;
  float *arc_acoustic_cost;
  int32_t *final_state;    /* [num_finals] lattice-state index                */
  float *final_cost;       /* [num_finals]                                    */
} b2k_raw_lattice;

int b2k_dec_get_raw_lattice(b2k_dec *dec, int32_t channel, b2k_raw_lattice *out, void *stream);

/* Batched form (CudaDecoder::GetRawLattice(channels, vector<Lattice*>&, ...),
 * cuda-decoder.h): the lattices of n finalized channels concatenated; lattice i
 * owns states [state_offs[i], state_offs[i+1]) etc.; arc/final state ids are
 * relative to the lattice's own first state.  Sizes query: out->state_frame NULL. */
int b2k_dec_get_raw_lattices(b2k_dec *dec, const int32_t *channels, int32_t n, b2k_raw_lattice *out,
                             int64_t *state_offs, int64_t *arc_offs, int64_t *final_offs, void *stream);

/* The batched read-back without a host round trip before the copy (pipelined pipelines: b2k_pipeline_submit / collect).
 * b2k_dec_pack_lattices_async packs the lattices of n finalized channels (d_channels: device array) into d_buf on `stream`:
 * a header of b2k_dec_pack_header_bytes(n) bytes (offsets, status, bytes needed) followed by the body; a channel in error /
 * not finalized / a buffer that is too small is reported in the header (and by b2k_dec_unpack_lattices), nothing is written
 * past the header then.  b2k_dec_unpack_lattices reads a HOST copy of that buffer; sizes query: out->state_frame NULL. */
int64_t b2k_dec_pack_header_bytes(int32_t n);
int b2k_dec_pack_lattices_async(b2k_dec *dec, const int32_t *d_channels, int32_t n, void *d_buf, int64_t cap_bytes, void *stream);
int b2k_dec_unpack_lattices(const void *h_buf, int32_t n, b2k_raw_lattice *out, int64_t *state_offs, int64_t *arc_offs,
                            int64_t *final_offs);

/* The best path of channels that are still decoding (or finalized), without touching their state:
 * LatticeFasterOnlineDecoderTpl::GetBestPath = BestPathEnd + TraceBackBestPath (decoder/lattice-faster-online-decoder.cc:
 * 54-167), CudaDecoder::GetBestPath (cudadecoder/cuda-decoder.h:279) -- partial hypotheses, and the best path with
 * use_final_probs = 0 that endpointing reads (online2/online-endpoint.cc:78-110).  The end token is the cheapest token of
 * the last frame (with its final cost when use_final_probs and some token of that frame is final); the predecessor of a
 * token is the link whose source cost + link cost is smallest (ties: lowest arena index).  Arcs come start first, at most
 * `cap` per channel (arrays [n x cap], any of them may be NULL): ilabel (transition-id, 0 = epsilon), olabel, graph cost,
 * acoustic cost minus the frame's cost offset (:150-154), the frame index of the arc's destination token list
 * (0 = before the first frame) and its HCLG state.  final_relative_cost is ComputeFinalCosts' (lattice-faster-decoder.cc:
 * 545-586) = FinalRelativeCost(); best_cost the end token's cost including final_cost.  Synchronizes `stream`. */
typedef struct {
  int32_t status;          /* 0, B2K_ERR_STATE (not decoding / in error / walk did not reach the start), B2K_ERR_OVERFLOW */
  int32_t n_arcs;          /* 0 with end_state -1: "No final token found" (:111)                                         */
  int32_t end_state;       /* HCLG state of the end token                                                                */
  int32_t num_frames;      /* NumFramesDecoded()                                                                         */
  float final_cost;        /* LatticeWeight(final_cost, 0) of the path's last state (:63)                                */
  float best_cost;
  float final_relative_cost;
} b2k_best_path_info;
int b2k_dec_best_path(b2k_dec *dec, const int32_t *channels, int32_t n, int32_t use_final_probs, int32_t cap,
                      int32_t *ilabels, int32_t *olabels, float *graph_costs, float *acoustic_costs,
                      int32_t *arc_frame, int32_t *arc_state, b2k_best_path_info *info, void *stream);

/* Debug/parity hook: copies the un-pruned token list of frame `frame_plus_one`
 * and the links created by that frame step (tests compare these bit-for-bit
 * against the oracle before any pruning).  links7 rows:
 * src_state, dst_state, ilabel, olabel, graph bits, acoustic bits, is_eps. */
int b2k_dec_debug_frame(b2k_dec *dec, int32_t channel, int32_t frame_plus_one,
                        int32_t *tok_state, float *tok_cost, int64_t *ntok,
                        int32_t *links7, int64_t *nlink, int64_t cap_tok, int64_t cap_link);

/* Per-frame diagnostics: next_cutoff and cost_offset of frames [0, T). */
int b2k_dec_frame_info(b2k_dec *dec, int32_t channel, float *cutoff, float *cost_offset,
                       int32_t *ntoks, int32_t cap);

/* ----------------------------------------------------------------- features */

/* Union of MfccOptions / FbankOptions / FrameExtractionOptions / MelBanksOptions
 * (feat/feature-mfcc.h:38-60, feature-fbank.h:41-60, feature-window.h:38-66,
 * mel-computations.h:43-58) — what OnlineNnet2FeaturePipelineInfo reads from
 * --mfcc-config / --fbank-config (online2/online-nnet2-feature-pipeline.cc:36-60). */
typedef struct {
  int32_t feature_type;        /* 0 = mfcc, 1 = fbank, 2 = plp (feat/feature-plp.h:38-66: lpc_order, compress_factor, cepstral_scale below) */
  float samp_freq, frame_shift_ms, frame_length_ms;
  float dither;                /* must be 0 (reference dither is unseeded)      */
  float preemph_coeff;
  int32_t remove_dc_offset, round_to_power_of_two, snip_edges;
  int32_t window_type;         /* 0 povey, 1 hamming, 2 hanning, 3 rectangular  */
  int32_t num_bins;
  float low_freq, high_freq;
  int32_t num_ceps, use_energy;
  float energy_floor;
  int32_t raw_energy;
  float cepstral_lifter;
  int32_t htk_compat, use_log_fbank, use_power, htk_mode;
  int32_t max_lanes;
  /* PlpOptions only (defaults 12, 0.33333, 1.0; num_ceps <= lpc_order + 1) */
  int32_t lpc_order;
  float compress_factor, cepstral_scale;
} b2k_feat_cfg;

void b2k_feat_cfg_default(b2k_feat_cfg *cfg);   /* mfcc_hires.conf with --dither=0 */

typedef struct b2k_feat b2k_feat;

/* MfccComputer / FbankComputer constructors (feature-mfcc.cc:82-115): builds the
 * window, mel-bank, DCT and lifter tables on the host and uploads them. */
int b2k_feat_create(const b2k_feat_cfg *cfg, b2k_feat **out);
int b2k_feat_destroy(b2k_feat *f);
int32_t b2k_feat_dim(const b2k_feat *f);                                    /* Dim()       */
float b2k_feat_samp_freq(const b2k_feat *f);                                /* GetFrameOptions().samp_freq: AcceptWaveform checks it (online-feature.cc:135-150) */
int32_t b2k_feat_num_frames(const b2k_feat *f, int64_t num_samples, int32_t flush);   /* NumFrames feature-window.cc:42 */

/* OnlineBatchedFeaturePipelineCuda::ComputeFeaturesBatched
 * (cudafeat/online-batched-feature-pipeline-cuda.h:44-134) /
 * OnlineGenericBaseFeature::ComputeFeatures (feat/online-feature.cc:162-204):
 * lane i computes frames [first_frame[i], first_frame[i]+num_frames[i]) of the
 * utterance whose samples so far are d_wave[i][0..num_samples[i]) (Kaldi's
 * int16-range float convention) into d_out[i] + frame*row_stride. */
int b2k_feat_compute_batched(b2k_feat *f, int32_t num_lanes, const float *const *d_wave,
                             const int32_t *num_samples, const int32_t *first_frame,
                             const int32_t *num_frames, float *const *d_out, int32_t row_stride,
                             void *stream);

/* The same on 16-bit PCM as it arrives (WaveData's samples before their conversion to float, feat/wave-reader.cc: Kaldi keeps
 * the int16 range in its floats, so (float)sample is the value the float form holds): half the bytes per frame. */
int b2k_feat_compute_batched_i16(b2k_feat *f, int32_t num_lanes, const int16_t *const *d_wave16,
                                 const int32_t *num_samples, const int32_t *first_frame,
                                 const int32_t *num_frames, float *const *d_out, int32_t row_stride,
                                 void *stream);

/* OnlineCmvnOptions (feat/online-feature.h:203-227) */
typedef struct {
  int32_t cmn_window, speaker_frames, global_frames, normalize_mean, normalize_variance;
} b2k_cmvn_cfg;

/* OnlineCmvn::GetFrame for a run of consecutive frames per lane
 * (feat/online-feature.cc:421-452).  d_state[i]: [2*(dim+1)] doubles, the raw
 * sliding-window stats after frame first_frame[i]-1 (zero before frame 0),
 * updated in place; d_global_stats / d_speaker_stats: OnlineCmvnState
 * global_cmvn_stats / speaker_cmvn_stats as [2*(dim+1)] doubles. */
int b2k_cmvn_apply_batched(b2k_feat *f, const b2k_cmvn_cfg *cfg, int32_t num_lanes,
                           const float *const *d_in, float *const *d_out, int32_t in_stride,
                           int32_t out_stride, const int32_t *first_frame, const int32_t *num_frames,
                           double *const *d_state, const double *d_global_stats,
                           const double *d_speaker_stats, void *stream);

/* ------------------------------------------------------------------ i-vector */

/* OnlineIvectorExtractionInfo (online2/online-ivector-feature.h:165-215): LDA
 * (splice + transform), diagonal UBM, IvectorExtractor derived quantities
 * (Sigma_inv_M_, U_; ivector/ivector-extractor.cc:182-218), global CMVN stats. */
typedef struct {
  int32_t base_dim, splice_left, splice_right, feat_dim, num_gauss, ivector_dim;
  int32_t num_gselect;
  float min_post, posterior_scale, max_count, prior_offset;
  int32_t num_cg_iters;
  int32_t cmn_window, speaker_frames, global_frames;   /* OnlineCmvnOptions of the extractor */
  int32_t max_lanes, max_frames;
} b2k_ivec_cfg;

typedef struct b2k_ivec b2k_ivec;

int b2k_ivec_create(const b2k_ivec_cfg *cfg, const float *lda /* [feat_dim x (spliced+1)] */,
                    const float *gconsts, const float *means_invvars, const float *inv_vars,
                    const double *sigma_inv_m /* [G x feat_dim x ivector_dim] */,
                    const double *U /* [G x D(D+1)/2] packed */, const double *global_cmvn_stats,
                    b2k_ivec **out);
int b2k_ivec_destroy(b2k_ivec *iv);

/* OnlineIvectorFeature::GetFrame as called once per nnet3 chunk
 * (decodable-online-looped.cc:170-205): for chunk n the statistics of all frames
 * <= sched[n] are accumulated (UpdateStatsUntilFrame, online-ivector-feature.cc
 * :248-281) and the i-vector re-estimated by <= num_cg_iters CG iterations from
 * the previous solution.  d_feats[i]: raw base features [num_frames x base_dim];
 * d_out[i]: [n_chunks x ivector_dim] (prior offset removed from dim 0).  sched[n] = -1
 * (only before the first frame index): no i-vector frame was ready when chunk n ran,
 * its i-vector is all zeros (decodable-online-looped.cc:188-197). */
int b2k_ivec_compute_batched(b2k_ivec *iv, int32_t num_lanes, const float *const *d_feats,
                             int32_t feat_stride, int32_t num_frames, const int32_t *sched,
                             int32_t n_chunks, float *const *d_out, int32_t out_stride, void *stream);

/* The same with speaker adaptation carried from utterance to utterance: OnlineIvectorFeature::SetAdaptationState before
 * the utterance and GetAdaptationState after it (online2/online-ivector-feature.cc:386-396,445-453; the tool's
 * per-speaker loop, online2bin/online2-wav-nnet3-latgen-faster.cc:199-221,287).  A state is
 * b2k_ivec_adaptation_state_doubles() doubles in device memory: OnlineCmvnState::speaker_cmvn_stats [2 x (base_dim+1)],
 * then OnlineIvectorEstimationStats {num_frames, linear term [D], quadratic term [D(D+1)/2, packed lower triangle]}.
 * d_state_in[i] NULL (or the array NULL): lane i starts as a new speaker (no speaker CMVN stats, prior-only i-vector
 * stats); d_state_out[i] non-NULL: the state after this utterance -- GetState(last frame) of the CMVN (every frame added
 * to the speaker stats) and the i-vector stats, both limited to max_remembered_frames as LimitFrames does (:109-127;
 * < 0 = no limit) -- which may be the array d_state_in[i] points to. */
int64_t b2k_ivec_adaptation_state_doubles(const b2k_ivec *iv);
int b2k_ivec_compute_batched_adapt(b2k_ivec *iv, int32_t num_lanes, const float *const *d_feats, int32_t feat_stride,
                                   int32_t num_frames, const int32_t *sched, int32_t n_chunks, float *const *d_out,
                                   int32_t out_stride, const double *const *d_state_in, double *const *d_state_out,
                                   float max_remembered_frames, void *stream);

/* -------------------------------------------------------------------- nnet3 */

/* A compiled forward program for one utterance length: the analogue of the
 * reference's NnetComputation (nnet3/nnet-computation.h) for the TDNN-F and CNN-TDNN-F
 * families, produced by b2k_nnet_compile below (kaldi_b200/nnet_model.py holds the same
 * compiler in Python as its test oracle).  The executor stands
 * behind NnetComputer::{AcceptInput,Run,GetOutputDestructive}
 * (nnet3/nnet-compute.h:95-200) as driven by DecodableNnetLoopedOnlineBase::
 * AdvanceChunk (nnet3/decodable-online-looped.cc:118-236) and
 * BatchedStaticNnet3::RunBatch (cudadecoder/batched-static-nnet3.h:59-138). */
typedef struct {
  int32_t dim, rows;
  int32_t kind;          /* 0 internal, 1 "input" features, 2 "ivector" (one row per chunk), 3 "output" */
  int64_t arena_off;     /* internal nodes: float offset in the per-utterance arena   */
} b2k_nnet_node;

typedef struct {
  int32_t src;           /* source node                                              */
  int32_t ratio, shift;  /* src_row = clamp(out_row*ratio + shift, lo, hi)            */
  int32_t lo, hi;
  int32_t ivec;          /* 1: src rows are chunks: row = clamp(floor(t/C) - m, lo, hi) with t = out_row*ratio+shift */
  int32_t C, m;
  int32_t k0, klen;      /* GEMM: columns [k0, k0+klen) of W multiply this term       */
  float scale;           /* EW: coefficient                                          */
  int32_t block;         /* EW: output column block this term adds into              */
  /* GEMM ops with hsplit > 1 (time-height convolution): the term reads klen columns starting at
   * column h*col_step + col_off of the source row, or zeros when that lies outside [0, col_lim)
   * (height zero padding); col_lim = 0 means a plain term                                     */
  int32_t col_step, col_off, col_lim;
} b2k_nnet_term;

typedef struct {
  int32_t type;          /* 0 GEMM (+fused epilogue), 1 elementwise                   */
  int32_t out, rows, N, K;
  int32_t n_terms;
  b2k_nnet_term terms[12];
  int64_t w, bias, bn_scale, bn_offset, sub_vec;   /* float offsets into the blob, -1 = absent */
  int32_t relu, has_res;
  b2k_nnet_term res;     /* bypass input: out = res_alpha*res + bn(...)               */
  float res_alpha, out_scale;
  int32_t log_softmax, block_dim;
  /* > 1: TimeHeightConvolutionComponent (nnet3/nnet-convolutional-component.cc:282): every output row is
   * hsplit GEMM rows (one per output height h), written to columns [h*N, (h+1)*N) of the node's row     */
  int32_t hsplit;
} b2k_nnet_op;

typedef struct b2k_nnet b2k_nnet;

int b2k_nnet_create(const b2k_nnet_node *nodes, int32_t n_nodes, const b2k_nnet_op *ops, int32_t n_ops,
                    const float *blob, int64_t blob_len, int32_t max_batch, b2k_nnet **out);
int b2k_nnet_destroy(b2k_nnet *nn);
int32_t b2k_nnet_num_output_frames(const b2k_nnet *nn);
int32_t b2k_nnet_output_dim(const b2k_nnet *nn);
double b2k_nnet_flops_per_lane(const b2k_nnet *nn);
int32_t b2k_nnet_num_launches_per_run(const b2k_nnet *nn);

/* One forward pass for `batch` utterances of the compiled length.  d_input[i]:
 * [T x feat_dim] features (row stride in_stride); d_ivectors[i]: [n_chunks x
 * ivector_dim], the i-vector the looped computation received for each chunk;
 * d_output[i]: [ceil(T/subsampling) x num_pdfs] pseudo log-likelihoods with
 * -log prior and acoustic scale applied (decodable-online-looped.cc:218-223). */
int b2k_nnet_run(b2k_nnet *nn, int32_t batch, const float *const *d_input, int32_t in_stride,
                 const float *const *d_ivectors, int32_t iv_stride, float *const *d_output,
                 int32_t out_stride, void *stream);

/* Which feature frame OnlineIvectorFeature::GetFrame is asked for by each nnet chunk when the CPU tool feeds
 * `chunk_samples` of audio at a time (online2-wav-nnet3-latgen-faster.cc:245-268, decodable-online-looped.cc:
 * 56-84,185-193): the schedule b2k_ivec_compute_batched takes; -1 for a chunk that runs before any i-vector frame is
 * ready (the reference leaves its i-vector zero, :188-197).  Host only. */
int b2k_ivec_online_schedule(int64_t num_samples, int32_t chunk_samples, int32_t frame_length, int32_t frame_shift,
                             int32_t num_feature_frames, int32_t nnet_right_context, int32_t frames_per_chunk,
                             int32_t subsampling, int32_t splice_right, int32_t *sched, int32_t max_chunks,
                             int32_t *n_chunks);

/* ------------------------------------------------------------------ nnet3 program compiler (host only)
 *
 * From a chain-model layer list (the xconfig layers of the TDNN, TDNN-F and CNN-TDNN-F recipes) and its parameters
 * to the op program b2k_nnet_create executes: the role of nnet3's compiler for this family
 * (nnet3/nnet-compile.cc, nnet-compile-looped.cc:329, nnet-optimize.cc) for utterances of `num_frames`
 * feature frames.  Needs no device.  kaldi_b200/nnet_model.py holds the same algorithm in Python as its
 * test oracle (identical nodes / ops / blob). */
typedef struct {
  char type[24];   /* idct | batchnorm | delta | lda | relu-batchnorm | tdnnf | linear | prefinal | output |
                      ivector-linear-bn | combine | conv  (steps/libs/nnet3/xconfig layer kinds)          */
  char name[48];
  char side[48];   /* combine: the node appended to the current one before the interleave                 */
  int32_t dim, bottleneck, stride, big, small, log_softmax;
  float bypass, append_ivector, target_rms;          /* append_ivector: 0 = none, else the Scale() factor */
  int32_t height, filters1, filters2;                /* combine-feature-maps-layer                        */
  int32_t height_in, height_out, height_subsample_out, filters_in, filters_out;   /* conv-relu-batchnorm-layer */
  int32_t n_time_offsets, time_offsets[8], n_height_offsets, height_offsets[8];
  /* time_offsets, besides conv: the splice of the input of an lda / relu-batchnorm layer (input=Append(-1,0,1) of the chain TDNN
   * recipes); n_time_offsets = 0 means the TDNN-F recipes' form: -1,0,1 in front of lda, none for relu-batchnorm */
} b2k_nnet_layer;

typedef struct {
  const char *name;      /* "<component>.w" [rows x cols], ".b" [rows], ".mean" / ".var" (BatchNorm stats), "priors" */
  const float *data;
  int64_t size;
  int32_t rows, cols;
} b2k_nnet_weight;

typedef struct {
  int32_t feat_dim, ivector_dim, num_pdfs, frame_subsampling_factor;
  int32_t num_frames;        /* feature frames per utterance the program is compiled for                  */
  int32_t frames_per_chunk;  /* --frames-per-chunk of the looped computation, a multiple of the subsampling factor */
  int32_t use_priors;        /* subtract log priors at the output (decodable-online-looped.cc:218-223)     */
  int32_t conv_dense;        /* 1: every convolution as a dense map per time offset (A/B checks)           */
  float acoustic_scale;
} b2k_nnet_compile_cfg;

typedef struct b2k_nnet_program b2k_nnet_program;

int b2k_nnet_compile(const b2k_nnet_compile_cfg *cfg, const b2k_nnet_layer *layers, int32_t n_layers,
                     const b2k_nnet_weight *weights, int32_t n_weights, b2k_nnet_program **out);
int b2k_nnet_program_destroy(b2k_nnet_program *prog);
int b2k_nnet_program_sizes(const b2k_nnet_program *prog, int32_t *n_nodes, int32_t *n_ops, int64_t *blob_len);
const b2k_nnet_node *b2k_nnet_program_nodes(const b2k_nnet_program *prog);
const b2k_nnet_op *b2k_nnet_program_ops(const b2k_nnet_program *prog);
const float *b2k_nnet_program_blob(const b2k_nnet_program *prog);
/* info: [0] output rows, [1] nnet chunks, [2]/[3] input frames needed left/right of the utterance,
 * [4]/[5] model left/right context (ComputeSimpleNnetContext), [6] chunk lag m of the i-vector, [7] arena floats per utterance */
int b2k_nnet_program_info(const b2k_nnet_program *prog, int64_t info[8]);
/* b2k_nnet_create on a compiled program (needs the device) */
int b2k_nnet_create_from_program(const b2k_nnet_program *prog, int32_t max_batch, b2k_nnet **out);

/* A window program: outputs at t = first_output_t + k*subsampling, k < num_outputs, of a window of cfg->num_frames input
 * frames that holds every frame those outputs read (nothing is padded; an output outside that range is an error) and ONE
 * i-vector for the window -- the computation request BatchedStaticNnet3::SetComputationRequest builds for a chunk with its
 * context (cudadecoder/batched-static-nnet3.cc:123-152).  Only the rows each layer needs for those outputs are computed.
 * ivector_rows = 1: that.  ivector_rows > 1: the window is chunk n of a LOOPED run (DecodableNnetLoopedOnlineBase::
 * AdvanceChunk, nnet3/decodable-online-looped.cc:118-236): first_output_t is the chunk's first frame, cfg->frames_per_chunk
 * = C its length, and the "ivector" input holds the i-vectors the chunks n-(rows-1) .. n received (chunk 0's repeated in
 * front of the utterance); time t reads row floor((t - first_output_t + (rows-1)*C) / C) - m, the Round() / lag arithmetic
 * of the looped compilation (nnet-compile-looped.cc:179-205), so the window's outputs are the looped computation's.
 * b2k_nnet_looped_ivector_rows: the rows that needs = ceil(left/C) + m + 1.  Host only. */
int b2k_nnet_compile_window(const b2k_nnet_compile_cfg *cfg, int32_t first_output_t, int32_t num_outputs,
                            int32_t ivector_rows, const b2k_nnet_layer *layers, int32_t n_layers,
                            const b2k_nnet_weight *weights, int32_t n_weights, b2k_nnet_program **out);
int b2k_nnet_looped_ivector_rows(const b2k_nnet_compile_cfg *cfg, const b2k_nnet_layer *layers, int32_t n_layers,
                                 int32_t *rows);
/* ComputeSimpleNnetContext (nnet3/nnet-utils.cc) of the layer list: left / right context in input frames.  Host only. */
int b2k_nnet_model_context(const b2k_nnet_compile_cfg *cfg, const b2k_nnet_layer *layers, int32_t n_layers,
                           int32_t *left, int32_t *right);

/* ------------------------------------------------------------------ chunked nnet3 with carried context
 *
 * cuda_decoder::BatchedStaticNnet3 (cudadecoder/batched-static-nnet3.h:59-138, .cc): every call takes at most one chunk of
 * features per channel, restores the channel's context frames in front of it, runs the network on the batch of windows,
 * saves the new context, and flushes the right context (last frame repeated) of the channels whose chunk was the last.
 * cfg->frames_per_chunk is compute_opts.frames_per_chunk (INPUT frames per chunk, >= the right context as the reference
 * asserts, :172-175); cfg->num_frames is ignored.  Output frame k of a call is the network's output at input time
 * (first new frame - right context) + k*subsampling, exactly the reference's arithmetic (so frames_per_chunk should be a
 * multiple of the subsampling factor there as here). */
typedef struct b2k_nnet_stream b2k_nnet_stream;

/* The frame bookkeeping of BatchContextSwitch for one channel (batched-static-nnet3.cc:151-190): given the frames its
 * context holds (0 = first chunk), the new frames and whether this is the flush, the frames the context holds afterwards
 * and the output frames the call produces.  Host only. */
int b2k_nnet_stream_account(int32_t left_context, int32_t right_context, int32_t subsampling, int32_t frames_in_context,
                            int32_t n_new_frames, int32_t flush, int32_t *frames_in_context_after,
                            int32_t *n_output_frames);

/* looped = 0: one i-vector per window, BatchedStaticNnet3's computation.  looped = 1: the windows are the chunks of
 * DecodableNnetLoopedOnlineBase::AdvanceChunk (nnet3/decodable-online-looped.cc:118-236) and d_ivectors[i] holds
 * info[7] rows, the i-vectors chunks n-(rows-1) .. n received (b2k_nnet_compile_window); the caller feeds the frames the
 * looped schedule reads -- first the right_context frames (no output yet; in pieces of at most frames_per_chunk when the right
 * context is longer than a chunk, the first piece with is_first_chunk), then frames_per_chunk per call,
 * indices clamped to the frames that exist (:150-160) and no flush -- and every call's outputs are the looped
 * computation's for that chunk (kaldi_b200/host/b2k_nnet3_shims.h: DecodableNnetLoopedOnlineB2k). */
int b2k_nnet_stream_create(const b2k_nnet_compile_cfg *cfg, const b2k_nnet_layer *layers, int32_t n_layers,
                           const b2k_nnet_weight *weights, int32_t n_weights, int32_t max_batch, int32_t nchannels,
                           int32_t looped, b2k_nnet_stream **out);
int b2k_nnet_stream_destroy(b2k_nnet_stream *s);
/* info: {output frames per chunk (GetNOutputFramesPerChunk), total left context, total right context
 * (GetTotalNnet3RightContext), window frames, input dim, i-vector dim, output dim, i-vector rows per slot} */
int b2k_nnet_stream_info(const b2k_nnet_stream *s, int64_t info[8]);
/* RunBatch (batched-static-nnet3.cc:293-367).  Batch slot i: channel channels[i]; d_features[i] = n_input_frames_valid[i]
 * (<= frames_per_chunk) new feature rows, row stride features_stride floats; d_ivectors[i] the i-vector of this run (NULL
 * array for a model without); is_first_chunk[i] resets the channel's context; is_last_chunk[i] flushes it.  Output frame k
 * of slot i goes to d_log_post + (i*opc + k)*out_stride, k < n_output_frames[i]; the frames of the flush of slot i to
 * d_eos_log_post + (i*opc + k)*out_stride, k < n_eos_output_frames[i] (0 for a slot that is not a last chunk), with
 * opc = output frames per chunk: in FormatOutputPtrs' words (:369-395) frame n_output_frames[i] + k of the channel.
 * Priors and acoustic scale are applied (:281-286).  Asynchronous on `stream`; the frame counts are final on return. */
int b2k_nnet_stream_run_batch(b2k_nnet_stream *s, int32_t n, const int32_t *channels, const float *const *d_features,
                              int32_t features_stride, const float *const *d_ivectors,
                              const int32_t *n_input_frames_valid, const int32_t *is_first_chunk,
                              const int32_t *is_last_chunk, float *d_log_post, float *d_eos_log_post,
                              int32_t out_stride, int32_t *n_output_frames, int32_t *n_eos_output_frames, void *stream);

/* ------------------------------------------------------------------ Kaldi model files (host only)
 *
 * Raw nnet3 models as Nnet::Write emits them (nnet3/nnet-nnet.cc:630-656) and final.mdl = TransitionModel
 * (hmm/transition-model.cc:394-453) + AmNnetSimple (nnet3/am-nnet-simple.cc:34-57), binary or text, without
 * Kaldi/OpenFst: parsed and mapped onto the layer list / named weights of b2k_nnet_compile (TDNN-F and
 * CNN-TDNN-F recipe layers; anything else is B2K_ERR_INVALID).  The arrays live as long as the handle.
 * kaldi_b200/kaldi_io.py is the Python twin; both are pinned to files written by the reference's own Write(). */
typedef struct b2k_model b2k_model;
int b2k_model_read(const char *path, int32_t is_mdl /* 1: final.mdl, 0: raw nnet3 */, b2k_model **out);
/* The same parse on bytes in memory -- what a Kaldi-side shim gets from Nnet::Write / AmNnetSimple::Write into a string
 * stream, with or without the "\0B" binary marker in front (kaldi_b200/host/b2k_nnet3_shims.h).  kind 0: raw nnet3
 * (Nnet::Write); 1: TransitionModel + AmNnetSimple (final.mdl); 2: AmNnetSimple::Write alone (b2k_model_read takes the
 * same values).  Host only. */
int b2k_model_read_memory(const void *data, int64_t len, int32_t kind, b2k_model **out);
int b2k_model_destroy(b2k_model *model);
/* The same object from arrays the caller already holds (synthetic models; everything is copied; "priors" optional, tid2pdf
 * may be NULL): the input of b2k_pipeline_create without a model file. */
int b2k_model_from_arrays(int32_t feat_dim, int32_t ivector_dim, int32_t num_pdfs, int32_t frame_subsampling_factor,
                          const b2k_nnet_layer *layers, int32_t n_layers, const b2k_nnet_weight *weights, int32_t n_weights,
                          const int32_t *tid2pdf, int32_t n_tids, b2k_model **out);
/* info: [0] feat_dim, [1] ivector_dim, [2] num_pdfs, [3] frame subsampling factor, [4] layers, [5] weights,
 * [6] size of the transition-id -> pdf table (0 for raw models), [7] 1 if the file carried priors */
int b2k_model_info(const b2k_model *model, int32_t info[8]);
/* The frame subsampling factor is not stored in a model file.  1 = the layer shapes do not decide it (splices at +-3
 * without a stride-3 TDNN-F layer: chain and non-chain recipes both have them): the compile routes refuse the model
 * until b2k_model_set_frame_subsampling_factor states it (the tool's --frame-subsampling-factor). */
int32_t b2k_model_frame_subsampling_ambiguous(const b2k_model *m);
int b2k_model_set_frame_subsampling_factor(b2k_model *m, int32_t factor);
const b2k_nnet_layer *b2k_model_layers(const b2k_model *model);
const b2k_nnet_weight *b2k_model_weights(const b2k_model *model);    /* includes "priors" (ones when absent) */
const int32_t *b2k_model_tid2pdf(const b2k_model *model);           /* [tid], index 0 unused; b2k_fst_csr.tid2pdf */
const int32_t *b2k_model_tid2phone(const b2k_model *model);         /* [tid], index 0 unused; TransitionIdToPhone (hmm/transition-model.cc:798) */

/* Kaldi option files (ParseOptions::ReadConfigFile, util/parse-options.cc:460-497) for the two configurations the tool is
 * pointed at: --mfcc-config / --fbank-config (MfccOptions / FbankOptions with their frame and mel options) and
 * --ivector-extraction-config (OnlineIvectorExtractionConfig, online2/online-ivector-feature.h:55-160, with the splice and
 * CMVN option files it names).  Values not in the file keep the REFERENCE'S defaults (e.g. dither 1.0, 23 mel bins, 13
 * cepstra — not mfcc_hires.conf); an unknown option is an error, as in the reference.  Host only; pinned to the reference's
 * own ParseOptions on the same files (tests/test_conf_cpp.py). */
int b2k_feat_cfg_from_conf(const char *conf_path, int32_t feature_type /* 0 mfcc, 1 fbank */, b2k_feat_cfg *cfg);
typedef struct {
  char lda_matrix[512], global_cmvn_stats[512], splice_config[512], cmvn_config[512], diag_ubm[512], ivector_extractor[512];
  int32_t ivector_period, use_most_recent_ivector, greedy_ivector_extractor, online_cmvn_iextractor;
  float max_remembered_frames;
} b2k_ivec_paths;
/* cfg: in = the caller's base_dim / capacities, out = + every option of the three files; paths: the files to hand to
 * b2k_ivec_files_read and the options that only matter to a streaming front end */
int b2k_ivec_cfg_from_conf(const char *conf_path, b2k_ivec_cfg *cfg, b2k_ivec_paths *paths);

/* conf/online.conf, the --config file of online2-wav-nnet3-latgen-faster: the feature group
 * (OnlineNnet2FeaturePipelineConfig::Register, online2/online-nnet2-feature-pipeline.h:101-126); options of the other
 * groups the tool registers (--endpoint.*, --ivector-silence-weighting.*, decoder and decodable options) are passed
 * through in `rest`, one per line. */
typedef struct {
  int32_t feature_type;               /* 0 mfcc, 1 fbank, 2 plp              */
  int32_t add_pitch;                  /* must be false                       */
  char mfcc_config[512], fbank_config[512], cmvn_config[512], global_cmvn_stats[512], ivector_extraction_config[512];
  char rest[4096];
  char plp_config[512];
} b2k_online_conf;
int b2k_online_conf_read(const char *conf_path, b2k_online_conf *out);

/* Endpointing (online2/online-endpoint.h:54-186, online-endpoint.cc:28-135): the five-rule disjunction, its option group
 * (--endpoint.silence-phones, --endpoint.ruleN.{must-contain-nonsilence, min-trailing-silence, max-relative-cost,
 * min-utterance-length}) and the trailing-silence count the rules are evaluated on.  Host only; pinned to the reference's own
 * online-endpoint.cc (tests/test_endpoint_cpp.py). */
typedef struct {
  int32_t must_contain_nonsilence;
  float min_trailing_silence, max_relative_cost, min_utterance_length;
} b2k_endpoint_rule;
typedef struct {
  b2k_endpoint_rule rule[5];          /* rule1 .. rule5 */
  char silence_phones[512];           /* colon-separated, e.g. "1:2:3:4:5" */
} b2k_endpoint_cfg;
int b2k_endpoint_cfg_default(b2k_endpoint_cfg *cfg);                          /* OnlineEndpointConfig() (online-endpoint.h:146-151) */
/* the endpoint.* options of an option file (the other groups a tool keeps in the same file are not looked at; an endpoint.*
 * name the group does not register is an error), on top of the defaults */
int b2k_endpoint_cfg_from_conf(const char *conf_path, b2k_endpoint_cfg *cfg);
/* the same for options given as text ("--endpoint.rule2.min-trailing-silence=0.8 ...", e.g. b2k_online_conf.rest), on top of *cfg */
int b2k_endpoint_cfg_apply_options(const char *text, b2k_endpoint_cfg *cfg);
/* EndpointDetected(config, num_frames_decoded, trailing_silence_frames, frame_shift, final_relative_cost) (online-endpoint.cc:47-76);
 * *detected = 0 / 1.  num_frames_decoded < trailing_silence_frames is an error (the reference asserts). */
int b2k_endpoint_detected(const b2k_endpoint_cfg *cfg, int32_t num_frames_decoded, int32_t trailing_silence_frames,
                          float frame_shift_in_seconds, float final_relative_cost, int32_t *detected);
/* TrailingSilenceLength (online-endpoint.cc:78-114) on a best path given as its input labels in time order (epsilons allowed, as
 * b2k_lat_best_path_arcs returns them): the number of silence frames at the end, counting stops at the first non-silence phone.
 * tid2phone = b2k_model_tid2phone.  A malformed, empty or duplicated silence list and an out-of-range transition-id are errors. */
int b2k_trailing_silence_frames(const int32_t *tid2phone, int32_t num_tids, const char *silence_phones, const int32_t *ilabels,
                                int64_t n, int32_t *frames);
/* EndpointDetected(config, tmodel, frame_shift, decoder) (online-endpoint.cc:116-135) with the decoder's best path, frame count and
 * final relative cost passed in; 0 frames decoded -> not detected.  trailing_silence_frames may be NULL. */
int b2k_endpoint_detected_on_path(const b2k_endpoint_cfg *cfg, const int32_t *tid2phone, int32_t num_tids, const int32_t *ilabels, int64_t n,
                                  int32_t num_frames_decoded, float frame_shift_in_seconds, float final_relative_cost,
                                  int32_t *detected, int32_t *trailing_silence_frames);

/* RIFF/WAVE input as WaveData::Read takes it (feat/wave-reader.cc:107-321): 16-bit PCM (plain or WAVE_FORMAT_EXTENSIBLE), RIFF or
 * RIFX, extra chunks skipped, "stream mode" sizes, truncated files; samples as floats in the int16 range, one row per channel
 * ([channels x samples]).  Host only; pinned to the reference's own reader (tests/test_wave_cpp.py). */
typedef struct b2k_wave b2k_wave;
int b2k_wave_read(const char *path, b2k_wave **out);
int b2k_wave_destroy(b2k_wave *wave);
int b2k_wave_info(const b2k_wave *wave, float *samp_freq, int32_t *channels, int64_t *samples);
const float *b2k_wave_data(const b2k_wave *wave);
/* ResampleWaveform (feat/resample.cc:368-376; LinearResample, low-pass at 0.99 x Nyquist of the lower rate, 6 zero crossings,
 * flush): what AcceptWaveform applies under --allow-downsample / --allow-upsample (feat/online-feature.cc:138-150).  n_out is
 * always set; with cap too small the call returns B2K_ERR_OVERFLOW.  Host only; within 1e-5 of the reference's output scale
 * (its dot products run in BLAS order), tests/test_wave_cpp.py. */
int b2k_resample_waveform(float orig_freq, const float *in, int64_t n_in, float new_freq, float *out, int64_t cap, int64_t *n_out);

/* HCLG.fst: an OpenFst binary "vector" or "const" FST over StdArc -> the CSR view of b2k_fst_create (arc order = file
 * order = the order ConstFst iterates in, which the decoder's results depend on).  Host only, no OpenFst.
 * PARITY UNPINNED: follows the published layout (fst/fst.h, vector-fst.h, const-fst.h); OpenFst is absent from this
 * image, so the reader is checked against kaldi_io.read_openfst / write_openfst only. */
typedef struct b2k_fst_file b2k_fst_file;
int b2k_fst_file_read(const char *path, b2k_fst_file **out);
int b2k_fst_file_destroy(b2k_fst_file *file);
/* csr: views into the file object (tid2pdf NULL); is_const (may be NULL): 1 for a "const" file */
int b2k_fst_file_csr(const b2k_fst_file *file, b2k_fst_csr *csr, int32_t *is_const);
/* b2k_fst_create on the file with the transition-id -> pdf table of the model (b2k_model_tid2pdf); needs the device */
int b2k_fst_create_from_file(const b2k_fst_file *file, const int32_t *tid2pdf, int32_t num_tids, b2k_fst **out);

/* The files of an ivector_extractor directory (steps/online/nnet2/train_ivector_extractor.sh): final.ie
 * (IvectorExtractor::Read, ivector/ivector-extractor.cc:828-848, + ComputeDerivedVars :182-230), final.dubm
 * (DiagGmm::Read, gmm/diag-gmm.cc:758-800), final.mat (LDA over spliced frames, with offset column) and
 * global_cmvn.stats.  Host only; kaldi_io.read_ivector_extractor / read_diag_gmm are the Python twins. */
typedef struct b2k_ivec_files b2k_ivec_files;
int b2k_ivec_files_read(const char *ie_path, const char *dubm_path, const char *lda_mat_path,
                        const char *global_cmvn_path, b2k_ivec_files **out);
int b2k_ivec_files_destroy(b2k_ivec_files *files);
/* info: [0] Gaussians, [1] feature dim, [2] i-vector dim, [3]/[4] rows/cols of final.mat, [5] dim of the CMVN stats,
 * [6] number of UBM weights */
int b2k_ivec_files_info(const b2k_ivec_files *files, int32_t info[8], float *prior_offset);
/* which: 0 lda [rows x cols], 1 gconsts [G], 2 means_invvars [G x F], 3 inv_vars [G x F], 4 UBM weights [G] */
const float *b2k_ivec_files_f32(const b2k_ivec_files *files, int32_t which);
/* which: 0 Sigma_inv_M [G x F x D], 1 U [G x D(D+1)/2], 2 global CMVN stats [2 x (dim+1)] */
const double *b2k_ivec_files_f64(const b2k_ivec_files *files, int32_t which);
/* b2k_ivec_create on the parsed files; cfg carries what ivector_extractor.conf / splice.conf say (base_dim, splice,
 * num_gselect, min_post, posterior_scale, max_count, num_cg_iters, CMVN options, capacities); the dimensions and the
 * prior offset come from the files.  Needs the device. */
int b2k_ivec_create_from_files(const b2k_ivec_cfg *cfg, const b2k_ivec_files *files, b2k_ivec **out);

/* ------------------------------------------------------------------ raw lattice -> compact lattice (host only)
 *
 * Where SingleUtteranceNnet3DecoderTpl::GetLattice calls DeterminizeLatticePhonePrunedWrapper
 * (online2/online-nnet3-decoding.cc:60-78, lat/determinize-lattice-pruned.h:284): the finalized raw lattice
 * (b2k_dec_get_raw_lattice; state 0 = start) becomes an acceptor over words, deterministic (one arc per word per
 * state, word epsilons absorbed), each arc / final weight carrying LatticeWeight (graph, acoustic) and the
 * transition-id string of the best path (CompactLatticeWeight, fstext/lattice-weight.h:387-604).  Every word
 * sequence whose best cost is within `beam` of the best path is kept, with exactly the weight and alignment of its
 * best path in the raw lattice; sequences outside the beam may or may not survive (as in the reference).
 * PARITY: pinned by equivalence against the reference's own determinizer compiled in oracle/_ref (same word sequences
 * within the beam, same weights, same alignments; tests/test_lattice_det.py); the reference's state numbering is not
 * reproduced — see kaldi_b200/csrc/lattice_det.cu. */
typedef struct {
  int64_t num_states, num_arcs, num_finals, num_tids;   /* state 0 = start */
  int32_t *arc_src, *arc_dst, *arc_word;                /* [num_arcs]                                   */
  float *arc_graph_cost, *arc_acoustic_cost;
  int64_t *arc_tids_off;                                /* [num_arcs + 1] offsets into tids             */
  int32_t *final_state;                                 /* [num_finals]                                 */
  float *final_graph_cost, *final_acoustic_cost;
  int64_t *final_tids_off;                              /* [num_finals + 1] offsets into tids           */
  int32_t *tids;                                        /* [num_tids] transition-ids, arcs then finals  */
} b2k_compact_lattice;

/* LatticeFasterDecoderTpl::GetBestPath (lattice-faster-decoder.cc:102-108) / CudaDecoder::GetBestPath on a finalized raw
 * lattice: word ids and transition-ids (epsilons removed) of the cheapest path, its graph cost (final cost included, as
 * LatticeWeight(final, 0) does) and acoustic cost.  cap = capacity of words[] and tids[]; with too small a cap the sizes
 * are returned with B2K_ERR_OVERFLOW.  Host only. */
int b2k_lat_best_path(const b2k_raw_lattice *raw, int32_t *words, int32_t *n_words, int32_t *tids, int32_t *n_tids, int32_t cap,
                      float *graph_cost, float *acoustic_cost);
/* The same path as indices into the raw lattice's arc arrays (in path order) and the index of its final entry (-1: empty
 * lattice): what a shim needs to build the linear Lattice GetBestPath returns, weights included. */
int b2k_lat_best_path_arcs(const b2k_raw_lattice *raw, int64_t *arcs, int64_t *n_arcs, int64_t cap, int64_t *final_index);

typedef struct b2k_clat b2k_clat;
/* max_states > 0: budget of determinized states; when it is exceeded the work is redone with 3/4 of the beam (the
 * reference reduces its beam when max_mem is hit, determinize-lattice-pruned.h:126-160) — see b2k_clat_effective_beam */
int b2k_lat_determinize_pruned(const b2k_raw_lattice *raw, float beam, int64_t max_states, b2k_clat **out);
/* DeterminizeLatticePhonePruned's two passes (lat/determinize-lattice-pruned.h:238-283, .cc:1291-1470): phone labels inserted on
 * the arcs that start a phone, determinization over words and phones, phone labels deleted, determinization over words
 * (--phone-determinize / --word-determinize; at least one must be set).  The transition model as three arrays over transition-ids
 * (index 0 unused): TransitionIdToPhone, IsSelfLoop, TransitionIdIsStartOfPhone.  Same accepted sequences, weights and alignments
 * as b2k_lat_determinize_pruned within the beam; b2k_clat_sizes' counters are the sum of both passes. */
/* In place: what DeterminizeLatticePhonePruned adds under --minimize (lat/determinize-lattice-pruned.cc:1459-1465, off by
 * default): PushCompactLatticeStrings + PushCompactLatticeWeights (lat/push-lattice.cc: transition-ids and costs move as
 * early as every path allows; the start state keeps what is left) and MinimizeCompactLattice (lat/minimize-lattice.cc:
 * states with the same words, strings and -- within delta, the reference's default is 1/1024 -- weights ahead of them
 * become one).  Same language, alignments and path weights (up to float rounding).  State 0 stays the start. */
int b2k_clat_minimize(b2k_clat *clat, float delta);
int b2k_lat_determinize_phone_pruned(const b2k_raw_lattice *raw, float beam, int64_t max_states, const int32_t *phone_of,
                                     const uint8_t *self_loop, const uint8_t *phone_start, int32_t num_tids,
                                     int32_t phone_determinize, int32_t word_determinize, b2k_clat **out);
/* CompactLatticeShortestPath + read-out (online2-wav-nnet3-latgen-faster.cc:43-76): words and transition-ids of the best path
 * of the compact lattice, graph / acoustic cost with the final weight included; too small capacities: sizes + B2K_ERR_OVERFLOW */
int b2k_clat_best_path(const b2k_clat *clat, int32_t *words, int32_t *n_words, int32_t *tids, int32_t *n_tids, int32_t cap_words,
                       int32_t cap_tids, float *graph_cost, float *acoustic_cost);
float b2k_clat_effective_beam(const b2k_clat *clat);
/* One table entry (lat/kaldi-lattice.cc WriteLattice / WriteCompactLattice) written or appended to `path`: binary = 1: key, ' ',
 * "\0B", OpenFst VectorFst container (published layout, unpinned) with the reference's weight encodings; binary = 0: the text form
 * (weights as the reference's printer writes them, unit weights left out).  kaldi_b200/lattice.py writes the same bytes. */
int b2k_lat_write(const b2k_raw_lattice *raw, const char *key, const char *path, int32_t binary, int32_t append);
int b2k_clat_write(const b2k_clat *clat, const char *key, const char *path, int32_t binary, int32_t append);
/* n lattices on up to num_threads host threads (0 = all cores); out[i], status[i] per lattice (status may be NULL); returns the
 * first non-zero status.  Same results as n single calls. */
int b2k_lat_determinize_pruned_batch(const b2k_raw_lattice *raws, int32_t n, float beam, int64_t max_states, int32_t num_threads,
                                     b2k_clat **out, int32_t *status);
int b2k_clat_destroy(b2k_clat *clat);
/* sizes: [0] states, [1] arcs, [2] finals, [3] transition-ids, [4] subsets expanded, [5] subset elements in total */
int b2k_clat_sizes(const b2k_clat *clat, int64_t sizes[6]);
/* fills the counts and every non-NULL array of `out` (caller-allocated from b2k_clat_sizes) */
int b2k_clat_copy(const b2k_clat *clat, b2k_compact_lattice *out);

/* ------------------------------------------------------------------ the batched pipeline (host waveforms -> lattices)
 *
 * BatchedThreadedNnet3CudaOnlinePipeline::DecodeBatch (cudadecoder/batched-threaded-nnet3-cuda-online-pipeline.cc:
 * 316-377: ComputeGPUFeatureExtraction -> RunNnet3 -> RunDecoder -> finalize) for batches of utterances bucketed to one
 * length, with the numerical semantics of online2-wav-nnet3-latgen-faster (online2bin/online2-wav-nnet3-latgen-faster.cc:
 * 199-299).  C++ above the stage calls of this header (kaldi_b200/csrc/pipeline.cu); kaldi_b200/pipeline.py spells the
 * same sequence in Python. */
typedef struct {
  b2k_feat_cfg feat;            /* max_lanes is set from max_batch                                           */
  b2k_dec_cfg dec;              /* max_frames / max_tokens / max_links <= 0: sized from the utterance length */
  int32_t frames_per_chunk;     /* 21 = --frames-per-chunk=20 rounded up to the subsampling factor           */
  float acoustic_scale;         /* 1.0 for chain models                                                      */
  int32_t max_batch;
  int64_t num_samples;          /* every utterance of a batch has exactly this many samples                  */
  float chunk_length_secs;      /* 0.18: the CPU tool's --chunk-length, decides which i-vector a chunk sees   */
  int32_t ivector_splice_right; /* --splice-config right context of the extractor (3)                        */
  int32_t use_priors, conv_dense;
  /* OnlineNnet2FeaturePipelineInfo::use_cmvn (online-nnet2-feature-pipeline.cc:108-123): the network reads OnlineCmvn of the
   * base features (the i-vector stage keeps the base features); global_cmvn_stats: host [2 x (feat_dim + 1)] doubles */
  int32_t use_cmvn;
  b2k_cmvn_cfg cmvn;
  const double *global_cmvn_stats;
  /* --frame-subsampling-factor (nnet3/decodable-simple-looped.h:75; the reference's default is 1, chain recipes pass 3).
   * 0 = not given: the model's own factor is used when its layers decide it (b2k_model_frame_subsampling_ambiguous == 0),
   * otherwise creating the pipeline is an error.  A value that contradicts the model (TDNN-F time-stride 3) is an error. */
  int32_t frame_subsampling_factor;
} b2k_pipeline_cfg;

void b2k_pipeline_cfg_default(b2k_pipeline_cfg *cfg);

typedef struct {
  int32_t num_feature_frames, feat_dim, num_output_frames, num_chunks, num_pdfs, ivector_dim;
  int32_t model_right_context;  /* filled by b2k_pipeline_create (0 in a bare plan)                          */
  int32_t chunk_samples;
  b2k_dec_cfg dec;              /* with the capacities resolved                                              */
  int64_t device_bytes, pinned_bytes;   /* the pipeline's own buffers (decoder arenas and model not included) */
} b2k_pipeline_plan;

typedef struct b2k_pipeline b2k_pipeline;

/* The sizes the pipeline will use for (cfg, model).  Host only. */
int b2k_pipeline_plan_for(const b2k_pipeline_cfg *cfg, const b2k_model *model, b2k_pipeline_plan *plan);
/* Compiles the model for the utterance length, creates the feature / nnet3 / decoder stages (max_batch lanes,
 * channel i = batch slot i) and the buffers.  fst and ivec (may be NULL: zero i-vectors) are not owned; an
 * extractor must have been created with max_lanes >= max_batch and max_frames >= plan.num_feature_frames. */
int b2k_pipeline_create(const b2k_pipeline_cfg *cfg, const b2k_model *model, const b2k_fst *fst, b2k_ivec *ivec,
                        b2k_pipeline **out);
int b2k_pipeline_destroy(b2k_pipeline *p);
int b2k_pipeline_get_plan(const b2k_pipeline *p, b2k_pipeline_plan *plan);
/* DecodeBatch: h_waves[i] = num_samples samples of utterance i (float in int16 range, or int16).  Stages them through
 * pinned memory, copies, and queues all four stages + FinalizeDecoding on `stream`; returns without waiting. */
int b2k_pipeline_decode_batch(b2k_pipeline *p, int32_t n, const float *const *h_waves, void *stream);
int b2k_pipeline_decode_batch_i16(b2k_pipeline *p, int32_t n, const int16_t *const *h_waves, void *stream);
/* Pipelined form (what bench.py's end-to-end number times): b2k_pipeline_submit_i16 stages batch k (int16 PCM) into one of
 * two pinned buffers, copies it on the pipeline's copy stream and queues all stages + lattice packing on its compute
 * stream; it returns without waiting, so batch k+1 can be submitted (its staging and copy overlap batch k's kernels)
 * before b2k_pipeline_collect(k) copies batch k's packed lattices back (overlapping batch k+1's kernels) and unpacks
 * them.  At most two batches may be outstanding; collect returns them in submission order.  The arrays collect fills
 * belong to the pipeline and stay valid until the next collect.  (DecodeBatch + the lattice callbacks of
 * cudadecoder/batched-threaded-nnet3-cuda-online-pipeline.cc:316-377,727-790 in one thread.) */
int b2k_pipeline_submit_i16(b2k_pipeline *p, int32_t n, const int16_t *const *h_waves);
int b2k_pipeline_collect(b2k_pipeline *p, int32_t *n_out, b2k_raw_lattice *view, const int64_t **state_offs,
                         const int64_t **arc_offs, const int64_t **final_offs);
/* Multi-GPU shards (kaldi_b200/ingest.py): int16 PCM that is already on the device (received from the ingest rank), and
 * the finalized lattices packed into a caller-owned device buffer (sent back to it; layout: b2k_dec_pack_lattices_async). */
int b2k_pipeline_run_device_i16(b2k_pipeline *p, int32_t n, const int16_t *d_waves, void *stream);
int b2k_pipeline_pack_device(b2k_pipeline *p, int32_t n, void *d_buf, int64_t cap_bytes, void *stream);
/* Same with the waveforms already on the device ([n x num_samples], NULL = the pipeline's own buffer as is). */
int b2k_pipeline_run_device(b2k_pipeline *p, int32_t n, const float *d_waves, void *stream);
/* The finalized raw lattices of batch slots 0..n-1 (b2k_dec_get_raw_lattices; waits for the stream). */
int b2k_pipeline_get_raw_lattices(b2k_pipeline *p, int32_t n, b2k_raw_lattice *out, int64_t *state_offs,
                                  int64_t *arc_offs, int64_t *final_offs, void *stream);
/* Stage handles / device buffers for diagnostics: the decoder (channel info, frame info), [max_batch x frames x dim]
 * features, [max_batch x chunks x ivector_dim] i-vectors, [max_batch x output frames x pdfs] log-likelihoods. */
/* Copies a stage output of batch slots 0..n-1 to the host and waits: what = 0 features, 1 i-vectors, 2 log-likelihoods. */
int b2k_pipeline_read(b2k_pipeline *p, int32_t what, int32_t n, float *h_out, void *stream);
/* Decoder / decodable / tool options in the tool's own spelling ("--beam=15 --lattice-beam=8 --acoustic-scale=1.0
 * --frames-per-chunk=20 ...", one per line or blank-separated, e.g. b2k_online_conf.rest or a command line): applied to cfg as
 * LatticeFasterDecoderConfig / NnetSimpleLoopedComputationOptions would take them (decoder/lattice-faster-decoder.h:75-98,
 * nnet3/decodable-simple-looped.h:66-88; --frames-per-chunk is rounded up to the subsampling factor as GetChunkSize does).
 * endpoint.* / ivector-silence-weighting.* / det.* and the options without effect here are accepted and ignored; an unknown
 * option is an error.  Host only. */
int b2k_pipeline_cfg_apply_options(const char *text, b2k_pipeline_cfg *cfg);
/* Stage timing for benchmarks: CUDA events between the stages of every run; b2k_pipeline_stage_times waits for the last
 * run and returns ms of {features, i-vectors (+ CMVN), nnet3, decoder init + advance, decoder finalize}. */
int b2k_pipeline_enable_stage_timing(b2k_pipeline *p, int32_t on);
int b2k_pipeline_stage_times(b2k_pipeline *p, float ms[5]);
double b2k_pipeline_nnet_flops_per_utterance(const b2k_pipeline *p);
b2k_dec *b2k_pipeline_decoder(b2k_pipeline *p);
const float *b2k_pipeline_features(const b2k_pipeline *p);
const float *b2k_pipeline_ivectors(const b2k_pipeline *p);
/* Speaker adaptation for the NEXT batch this pipeline runs (decode_batch / submit / run_device), then forgotten: slot i
 * starts from the state d_state_in[i] points to (NULL: a new speaker) and leaves the speaker's state after the utterance in
 * d_state_out[i] (NULL: not kept) -- b2k_ivec_compute_batched_adapt's arguments; device pointers to
 * b2k_ivec_adaptation_state_doubles() doubles that must stay valid until the batch has run.  The tool's per-speaker loop
 * (online2bin/online2-wav-nnet3-latgen-faster.cc:199-221,287) becomes: batch k holds the k-th utterance of every speaker
 * (kaldi_b200/ingest.py: speaker_waves), so two slots of one batch never share a state (refused). n = 0 clears it. */
int b2k_pipeline_set_speaker_states(b2k_pipeline *p, int32_t n, const double *const *d_state_in, double *const *d_state_out,
                                    float max_remembered_frames);
const float *b2k_pipeline_loglikes(const b2k_pipeline *p);

/* ------------------------------------------------------------------ many audio streams, chunk by chunk
 *
 * cuda_decoder::BatchedThreadedNnet3CudaOnlinePipeline::DecodeBatch(corr_ids, wave_samples, is_first_chunk, is_last_chunk)
 * (cudadecoder/batched-threaded-nnet3-cuda-online-pipeline.cc:316-377): every call brings at most one chunk of 16-bit PCM per
 * channel; the features of the frames that became computable, the chunked network (b2k_nnet_stream_run_batch, context per
 * channel) and the decoder's frames run for the whole batch of channels; nothing of an utterance is recomputed.  A chunk may
 * bring at most frames_per_chunk new feature frames.  Partial hypotheses: b2k_dec_best_path on b2k_stream_decoder() with
 * use_final_probs = 0 (channel ids are the decoder's channel ids); after a channel's last chunk the call has run
 * FinalizeDecoding and its raw lattice is b2k_dec_get_raw_lattice(b2k_stream_decoder(s), channel). */
typedef struct {
  b2k_feat_cfg feat;
  b2k_dec_cfg dec;              /* max_frames is raised to what max_seconds needs                                  */
  int32_t nchannels;            /* streams in flight at most (= the largest batch of one call)                      */
  float max_seconds;            /* per stream                                                                      */
  int32_t frames_per_chunk;     /* compute_opts.frames_per_chunk of BatchedStaticNnet3 (input frames per call)      */
  float acoustic_scale;
  int32_t use_priors;
} b2k_stream_cfg;
typedef struct b2k_stream b2k_stream;
void b2k_stream_cfg_default(b2k_stream_cfg *cfg);
int b2k_stream_create(const b2k_stream_cfg *cfg, const b2k_model *model, const b2k_fst *fst, b2k_stream **out);
int b2k_stream_destroy(b2k_stream *s);
/* info: {channels, max samples, max feature frames, feature dim, pdfs, i-vector dim, output frames per chunk, frames per chunk} */
int b2k_stream_info(const b2k_stream *s, int64_t info[8]);
/* Slot i: channel channels[i], num_samples[i] new samples at h_chunks[i] (host), is_first_chunk[i] starts a new utterance on
 * the channel, is_last_chunk[i] ends it (right context flushed, FinalizeDecoding run).  d_ivectors: NULL (zeros) or one device
 * i-vector per slot.  Optional outputs per slot: output frames decoded by this call / so far, and where this call's
 * log-likelihood rows lie ([frames x pdfs], valid until the next call; NULL when there are none).  Asynchronous on `stream`
 * except for the copy of pageable host chunks. */
int b2k_stream_decode_batch_i16(b2k_stream *s, int32_t n, const int32_t *channels, const int16_t *const *h_chunks,
                                const int32_t *num_samples, const int32_t *is_first_chunk, const int32_t *is_last_chunk,
                                const float *const *d_ivectors, int32_t *new_output_frames, int32_t *output_frames_so_far,
                                const float **d_new_frames, const float **d_flushed_frames, void *stream);
b2k_dec *b2k_stream_decoder(b2k_stream *s);
const float *b2k_stream_features(const b2k_stream *s, int32_t channel);   /* [max feature frames x dim] of a channel */

#ifdef __cplusplus
}
#endif
#endif /* B2K_H_ */
