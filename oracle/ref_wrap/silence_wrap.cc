// TEST INFRASTRUCTURE ONLY (oracle side).  Compiles the reference's OWN online2/online-ivector-feature.cc where it lies and
// exposes (1) OnlineSilenceWeighting (online-ivector-feature.h:460-571): the constructor, ComputeCurrentTraceback<FST> (the
// LatticeFasterOnlineDecoderTpl instantiation), GetDeltaWeights and GetNonsilenceFrames run unmodified; the decoder they walk
// is the replay device of replay_decoder.h, whose tokens are the ids the test hands in; (2) OnlineIvectorFeature itself
// (ref_ivector_run_real): the class the restated glue of ivector_wrap.cc stands for, with optional frame weights.
#include <cstdio>
#include <string>
#include <utility>
#include <vector>

#include "replay_decoder.h"
#include "util/common-utils.h"
#include "hmm/transition-model.h"

#include "online2/online-ivector-feature.cc"
#include "ivector_ref_types.h"

using namespace kaldi;
using b2k_oracle::RefIvec;
using b2k_oracle::RefSpeaker;

namespace {
typedef LatticeFasterOnlineDecoderTpl<fst::Fst<fst::StdArc> > Dec;
struct RefSilence {
  TransitionModel tm;
  OnlineSilenceWeightingConfig cfg;          // the class keeps references to both
  OnlineSilenceWeighting *w = NULL;
  ~RefSilence() { delete w; }
};
}  // namespace

extern "C" {

void *ref_silw_create(const char *mdl_path, const char *silence_phones, float silence_weight, float max_state_duration,
                      int frame_subsampling_factor) {
  try {
    RefSilence *s = new RefSilence();
    { bool binary; Input ki(mdl_path, &binary); s->tm.Read(ki.Stream(), binary); }
    s->cfg.silence_phones_str = silence_phones;
    s->cfg.silence_weight = silence_weight;
    s->cfg.max_state_duration = max_state_duration;
    s->w = new OnlineSilenceWeighting(s->tm, s->cfg, frame_subsampling_factor);
    return s;
  } catch (const std::exception &e) { fprintf(stderr, "ref_silw_create: %s\n", e.what()); return NULL; }
}
void ref_silw_destroy(void *p) { delete (RefSilence *)p; }
int ref_silw_active(void *p) { return ((RefSilence *)p)->w->Active() ? 1 : 0; }

// ComputeCurrentTraceback(decoder, use_final_probs = false) on a replayed best path: ilabels in time order (0 = epsilon),
// token_ids[k] = an id of the token arc k leaves, `frames` = NumFramesDecoded().  0, -1 on a KALDI_ERR / failed assertion.
int ref_silw_traceback(void *p, const int *ilabels, const int *token_ids, int n_arcs, int frames) {
  try {
    Dec dec;
    dec.path.assign(ilabels, ilabels + n_arcs);
    dec.source.assign(token_ids, token_ids + n_arcs);
    dec.frames = frames;
    ((RefSilence *)p)->w->ComputeCurrentTraceback(dec, false);
    return 0;
  } catch (const std::exception &e) { return -1; }
}

// GetDeltaWeights(num_frames_ready, first_decoder_frame, &delta_weights): returns the number of pairs (-1 on error, -2 when
// more than cap).
int ref_silw_delta_weights(void *p, int num_frames_ready, int first_decoder_frame, int *frame_out, float *weight_out, int cap) {
  try {
    std::vector<std::pair<int32, BaseFloat> > d;
    ((RefSilence *)p)->w->GetDeltaWeights(num_frames_ready, first_decoder_frame, &d);
    if ((int)d.size() > cap) return -2;
    for (size_t i = 0; i < d.size(); i++) { frame_out[i] = d[i].first; weight_out[i] = d[i].second; }
    return (int)d.size();
  } catch (const std::exception &e) { return -1; }
}

int ref_silw_nonsilence_frames(void *p, int num_frames_ready, int first_decoder_frame, int *frame_out, int cap) {
  try {
    std::vector<int32> f;
    ((RefSilence *)p)->w->GetNonsilenceFrames(num_frames_ready, first_decoder_frame, &f);
    if ((int)f.size() > cap) return -2;
    for (size_t i = 0; i < f.size(); i++) frame_out[i] = f[i];
    return (int)f.size();
  } catch (const std::exception &e) { return -1; }
}

// The reference's OWN OnlineIvectorFeature (online-ivector-feature.cc:399-443 ctor, :327-355 GetFrame, :445-453 / :386-396
// Set / GetAdaptationState with LimitFrames) on the inputs of ref_ivector_run_speaker (ivector_wrap.cc), same argument meaning:
// sched[n] = frame passed to GetFrame for chunk n (-1: not called, zeros), use_most_recent_ivector = true, greedy = false.
// Before chunk n's GetFrame the delta weights [dw_off[n], dw_off[n + 1]) go to UpdateFrameWeights (dw_off NULL: unweighted).
int ref_ivector_run_real(void *h, const float *feats, int T, int D, const double *global_cmvn, int cmn_window,
                         int speaker_frames, int global_frames, int splice_left, int splice_right, int num_gselect,
                         float min_post, float posterior_scale, float max_count, int num_cg_iters,
                         int online_cmvn_iextractor, const int *sched, int n_chunks, float *out,
                         void *speaker, float max_remembered_frames,
                         const int *dw_off, const int *dw_frame, const float *dw_weight) {
  try {
    RefIvec *r = (RefIvec *)h;
    RefSpeaker *spk = (RefSpeaker *)speaker;
    OnlineIvectorExtractionInfo info;
    info.lda_mat = r->lda;
    info.global_cmvn_stats.Resize(2, D + 1);
    for (int i = 0; i < 2; i++) for (int j = 0; j <= D; j++) info.global_cmvn_stats(i, j) = global_cmvn[i * (D + 1) + j];
    info.cmvn_opts.cmn_window = cmn_window; info.cmvn_opts.speaker_frames = speaker_frames; info.cmvn_opts.global_frames = global_frames;
    info.online_cmvn_iextractor = online_cmvn_iextractor != 0;
    info.splice_opts.left_context = splice_left; info.splice_opts.right_context = splice_right;
    info.diag_ubm.CopyFromDiagGmm(r->ubm);
    { std::stringstream ss(std::ios::in | std::ios::out | std::ios::binary);      // IvectorExtractor has no copy: Write / Read
      r->extractor.Write(ss, true); info.extractor.Read(ss, true); }
    info.ivector_period = 10; info.num_gselect = num_gselect; info.min_post = min_post; info.posterior_scale = posterior_scale;
    info.max_count = max_count; info.num_cg_iters = num_cg_iters; info.use_most_recent_ivector = true;
    info.greedy_ivector_extractor = false;
    info.max_remembered_frames = max_remembered_frames >= 0.0f ? max_remembered_frames : 1.0e30f;   // < 0: LimitFrames never scales
    info.Check();

    Matrix<BaseFloat> m(T, D);
    for (int t = 0; t < T; t++) memcpy(m.RowData(t), feats + (size_t)t * D, 4 * D);
    OnlineMatrixFeature base(m);
    OnlineIvectorFeature feature(info, &base);
    if (spk && spk->has) {
      OnlineIvectorExtractorAdaptationState st(info);
      st.cmvn_state = spk->cmvn;
      st.cmvn_state.global_cmvn_stats = info.global_cmvn_stats;
      st.ivector_stats = *spk->stats;
      feature.SetAdaptationState(st);
    }
    const int ivdim = feature.Dim();
    Vector<BaseFloat> row(ivdim);
    for (int n = 0; n < n_chunks; n++) {
      if (dw_off && dw_off[n + 1] > dw_off[n]) {
        std::vector<std::pair<int32, BaseFloat> > d;
        for (int k = dw_off[n]; k < dw_off[n + 1]; k++) d.push_back(std::make_pair(dw_frame[k], dw_weight[k]));
        feature.UpdateFrameWeights(d);
      }
      if (sched[n] < 0) { for (int d = 0; d < ivdim; d++) out[(size_t)n * ivdim + d] = 0.0f; continue; }
      feature.GetFrame(sched[n], &row);
      for (int d = 0; d < ivdim; d++) out[(size_t)n * ivdim + d] = row(d);
    }
    if (spk) {
      OnlineIvectorExtractorAdaptationState st(info);
      feature.GetAdaptationState(&st);
      spk->cmvn = st.cmvn_state;
      spk->stats.reset(new OnlineIvectorEstimationStats(st.ivector_stats));
      spk->has = true;
    }
    return 0;
  } catch (const std::exception &e) { fprintf(stderr, "ref_ivector_run_real: %s\n", e.what()); return -1; }
}

}  // extern "C"
