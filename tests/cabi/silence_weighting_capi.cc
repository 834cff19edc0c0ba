// Test harness only: b2k_host::SilenceWeighting (kaldi_b200/host/b2k_silence_weighting.h) behind the same C functions as the
// reference's class has in oracle/ref_wrap/silence_wrap.cc, so that tests/test_silence_weighting.py drives both alike.
#include <cstdio>
#include <utility>
#include <vector>

#include "b2k_silence_weighting.h"

using b2k_host::SilenceWeighting;

extern "C" {

void *silw_create(const int *tid2phone, int n, const char *silence_phones, float silence_weight, float max_state_duration, int fs) {
  try {
    return new SilenceWeighting(std::vector<int32_t>(tid2phone, tid2phone + n), silence_phones, silence_weight, max_state_duration, fs);
  } catch (const std::exception &e) { fprintf(stderr, "silw_create: %s\n", e.what()); return NULL; }
}
void silw_destroy(void *p) { delete (SilenceWeighting *)p; }
int silw_active(void *p) { return ((SilenceWeighting *)p)->Active() ? 1 : 0; }
int silw_list_ok(void *p) { return ((SilenceWeighting *)p)->SilencePhonesParsed() ? 1 : 0; }

// a best path as b2k_dec_best_path hands it out: ilabels (0 = epsilon) and the state each arc ENTERS
int silw_traceback_from_path(void *p, const int *ilabels, const int *arc_state, int n_arcs, int frames) {
  try {
    return ((SilenceWeighting *)p)->SetTracebackFromPath(ilabels, arc_state, n_arcs, frames) ? 0 : -3;
  } catch (const std::exception &e) { return -1; }
}
int silw_delta_weights(void *p, int num_frames_ready, int first_decoder_frame, int *frame_out, float *weight_out, int cap) {
  try {
    std::vector<std::pair<int32_t, float> > d;
    ((SilenceWeighting *)p)->GetDeltaWeights(num_frames_ready, first_decoder_frame, &d);
    if ((int)d.size() > cap) return -2;
    for (size_t i = 0; i < d.size(); i++) { frame_out[i] = d[i].first; weight_out[i] = d[i].second; }
    return (int)d.size();
  } catch (const std::exception &e) { return -1; }
}
int silw_nonsilence_frames(void *p, int num_frames_ready, int first_decoder_frame, int *frame_out, int cap) {
  try {
    std::vector<int32_t> f;
    ((SilenceWeighting *)p)->GetNonsilenceFrames(num_frames_ready, first_decoder_frame, &f);
    if ((int)f.size() > cap) return -2;
    for (size_t i = 0; i < f.size(); i++) frame_out[i] = f[i];
    return (int)f.size();
  } catch (const std::exception &e) { return -1; }
}

}  // extern "C"
