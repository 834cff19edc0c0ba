// feat_kernels.cuh — declarations shared by feat.cu and ivector.cu
#pragma once
#include <cuda_runtime.h>

namespace b2k {

struct CmvnLane {
  const float *in;       // raw features [num_frames_total x dim] (row stride in_stride)
  float *out;
  int in_stride, out_stride;
  int first_frame, num_frames;        // frames to produce in this call
  double *state;         // [2*(dim+1)] sliding-window stats after frame first_frame-1 (persisted)
  // speaker adaptation (OnlineCmvnState::speaker_cmvn_stats): this lane's own speaker stats (null: CmvnParams::speaker_stats),
  // and where OnlineCmvn::GetState(last frame) + LimitFrames leave them after the call (null: not wanted; may equal `speaker`)
  const double *speaker;
  double *speaker_out;
  float max_remembered_frames;   // < 0: no limit
};
struct CmvnParams {
  int dim, cmn_window, speaker_frames, global_frames, normalize_mean, normalize_variance;
  const double *global_stats;   // [2*(dim+1)] or NULL
  const double *speaker_stats;  // [2*(dim+1)] or NULL
};


// launches cmvn_kernel (feat.cu): OnlineCmvn::GetFrame for runs of consecutive frames
int launch_cmvn(const CmvnParams &cp, const CmvnLane *d_lanes, int num_lanes, cudaStream_t st);

}  // namespace b2k
