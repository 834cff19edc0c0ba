// host_utils.cu — small host-only pieces of the online2 control flow that the batched pipeline needs
// (no device code).  Python twins in kaldi_b200/ivector.py are the test oracles (tests/test_host_utils.py).
#include <algorithm>

#include "common.cuh"

extern "C" {

// For each nnet chunk n, the frame index OnlineIvectorFeature::GetFrame is called with when
// online2-wav-nnet3-latgen-faster feeds `chunk_samples` at a time (online2-wav-nnet3-latgen-faster.cc:245-268):
// the chunk is computed by the first AdvanceDecoding after which DecodableNnetLoopedOnlineBase::NumFramesReady
// (decodable-online-looped.cc:56-84) covers it, and it asks for frame min(features_ready - 1,
// ivector_frames_ready - 1) (decodable-online-looped.cc:185-193), ivector_frames_ready being smaller by the
// splice right context until the input is finished.  snip-edges framing.
int b2k_ivec_online_schedule(int64_t num_samples, int32_t chunk_samples, int32_t frame_length, int32_t frame_shift,
                             int32_t num_feature_frames, int32_t nnet_right_context, int32_t frames_per_chunk,
                             int32_t subsampling, int32_t splice_right, int32_t *sched, int32_t max_chunks,
                             int32_t *n_chunks_out) {
  if (num_samples <= 0 || chunk_samples <= 0 || frame_length <= 0 || frame_shift <= 0 || num_feature_frames <= 0 ||
      frames_per_chunk <= 0 || subsampling <= 0 || !sched || !n_chunks_out)
    return b2k::set_error(B2K_ERR_INVALID, "b2k_ivec_online_schedule: bad args");
  const int32_t T = num_feature_frames;
  const int32_t n_out = (T + subsampling - 1) / subsampling;
  const int32_t n_chunks = (n_out * subsampling + frames_per_chunk - 1) / frames_per_chunk;
  *n_chunks_out = n_chunks;
  if (n_chunks > max_chunks) return b2k::set_error(B2K_ERR_OVERFLOW, "b2k_ivec_online_schedule: schedule buffer too small");
  int64_t fed = 0;
  int32_t done = 0;
  while (done < n_chunks) {
    fed = std::min<int64_t>(fed + chunk_samples, num_samples);
    const bool finished = fed >= num_samples;
    int32_t ready = fed < frame_length ? 0 : (int32_t)(1 + (fed - frame_length) / frame_shift);
    int32_t chunks_ready, iv_frame;
    if (finished) {
      ready = T; chunks_ready = n_chunks; iv_frame = T - 1;
    } else {
      chunks_ready = std::max(0, ready - nnet_right_context) / frames_per_chunk;
      const int32_t iv_ready = std::max(0, ready - splice_right);
      iv_frame = std::min(ready - 1, iv_ready - 1);
    }
    while (done < std::min(chunks_ready, n_chunks)) sched[done++] = std::max(iv_frame, 0);
  }
  return B2K_OK;
}

}  // extern "C"
