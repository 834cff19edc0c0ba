// oracle/ref_wrap/ivector_wrap.cc — TEST INFRASTRUCTURE ONLY.
// Online i-vector oracle: the arithmetic is the reference's OWN code compiled
// from /root/reference/src —
//   IvectorExtractor (text Read + ComputeDerivedVars)   ivector/ivector-extractor.cc:182-218,828-849
//   OnlineIvectorEstimationStats::{AccStats,GetIvector}  ivector/ivector-extractor.cc:611-668,732-756
//   LinearCgd<double>                                   matrix/optimization.cc:453-565
//   DiagGmm::LogLikelihoods                             gmm/diag-gmm.cc:546-562
//   VectorToPosteriorEntry                              hmm/posterior.cc:440-505
//   OnlineCmvn / OnlineSpliceFrames / OnlineTransform   feat/online-feature.cc
// and this file restates only the glue of OnlineIvectorFeature
// (online2/online-ivector-feature.cc: ctor :399-443, GetMinPost :188-199,
// UpdateStatsForFrames :201-245, UpdateStatsUntilFrame :248-281, GetFrame
// :327-355).  That translation unit also instantiates OnlineSilenceWeighting over the
// OpenFst-based decoders; silence_wrap.cc compiles it over a replay decoder instead and
// runs the reference's OWN OnlineIvectorFeature on the same inputs (ref_ivector_run_real):
// tests/test_ivector_oracle_pin.py holds the glue below to it, bit for bit.
#include <sstream>
#include <memory>
#include <vector>

#include "feat/online-feature.h"
#include "gmm/diag-gmm.h"
#include "hmm/posterior.h"
#include "ivector/ivector-extractor.h"
#include "util/kaldi-io.h"

using namespace kaldi;

#include "ivector_ref_types.h"
using namespace b2k_oracle;

extern "C" {

void *ref_ivector_speaker_create() { return new RefSpeaker(); }
void ref_ivector_speaker_destroy(void *s) { delete (RefSpeaker *)s; }
// The speaker's state as flat doubles in b2k's layout: speaker_cmvn_stats [2 x (D+1)], num_frames, linear [ivdim], quadratic
// [packed lower triangle] (read back through the reference's own binary Write / Read: the members are private).
int ref_ivector_speaker_state(void *sp, int D, int ivdim, double *out) {
  try {
    RefSpeaker *s = (RefSpeaker *)sp;
    if (!s->has) return -1;
    for (int r = 0; r < 2; r++) for (int c = 0; c <= D; c++) out[r * (D + 1) + c] = s->cmvn.speaker_cmvn_stats.NumRows() ? s->cmvn.speaker_cmvn_stats(r, c) : 0.0;
    std::stringstream ss(std::ios::in | std::ios::out | std::ios::binary);
    s->stats->Write(ss, true);
    double prior_offset, max_count, num_frames;
    ExpectToken(ss, true, "<OnlineIvectorEstimationStats>");
    ExpectToken(ss, true, "<PriorOffset>"); ReadBasicType(ss, true, &prior_offset);
    ExpectToken(ss, true, "<MaxCount>"); ReadBasicType(ss, true, &max_count);
    ExpectToken(ss, true, "<NumFrames>"); ReadBasicType(ss, true, &num_frames);
    SpMatrix<double> q; Vector<double> l;
    ExpectToken(ss, true, "<QuadraticTerm>"); q.Read(ss, true);
    ExpectToken(ss, true, "<LinearTerm>"); l.Read(ss, true);
    double *o = out + 2 * (D + 1);
    o[0] = num_frames;
    for (int i = 0; i < ivdim; i++) o[1 + i] = l(i);
    for (int i = 0, k = 0; i < ivdim; i++) for (int j = 0; j <= i; j++, k++) o[1 + ivdim + k] = q(i, j);
    return 0;
  } catch (const std::exception &e) { fprintf(stderr, "ref_ivector_speaker_state: %s\n", e.what()); return -1; }
}

void *ref_ivector_create(const char *extractor_text, const char *ubm_text, const float *lda, int lda_rows, int lda_cols) {
  try {
    RefIvec *r = new RefIvec();
    { std::istringstream is(extractor_text); r->extractor.Read(is, false); }
    { std::istringstream is(ubm_text); r->ubm.Read(is, false); }
    r->lda.Resize(lda_rows, lda_cols);
    for (int i = 0; i < lda_rows; i++) for (int j = 0; j < lda_cols; j++) r->lda(i, j) = lda[i * lda_cols + j];
    return r;
  } catch (const std::exception &e) { fprintf(stderr, "ref_ivector_create: %s\n", e.what()); return nullptr; }
}
void ref_ivector_destroy(void *h) { delete (RefIvec *)h; }

// sched[n] = frame index passed to OnlineIvectorFeature::GetFrame for nnet chunk n
// (non-decreasing).  out: [n_chunks x ivector_dim], prior offset already subtracted from dim 0.
int ref_ivector_run_speaker(void *h, const float *feats, int T, int D, const double *global_cmvn, int cmn_window,
                            int speaker_frames, int global_frames, int splice_left, int splice_right, int num_gselect,
                            float min_post, float posterior_scale, float max_count, int num_cg_iters,
                            int online_cmvn_iextractor, const int *sched, int n_chunks, float *out,
                            float *debug_lda_raw, float *debug_lda_norm, void *speaker, float max_remembered_frames);

int ref_ivector_run(void *h, const float *feats, int T, int D, const double *global_cmvn, int cmn_window,
                    int speaker_frames, int global_frames, int splice_left, int splice_right, int num_gselect,
                    float min_post, float posterior_scale, float max_count, int num_cg_iters,
                    int online_cmvn_iextractor, const int *sched, int n_chunks, float *out,
                    float *debug_lda_raw, float *debug_lda_norm) {
  return ref_ivector_run_speaker(h, feats, T, D, global_cmvn, cmn_window, speaker_frames, global_frames, splice_left, splice_right,
                                 num_gselect, min_post, posterior_scale, max_count, num_cg_iters, online_cmvn_iextractor, sched,
                                 n_chunks, out, debug_lda_raw, debug_lda_norm, nullptr, -1.0f);
}

// speaker != NULL: SetAdaptationState(speaker) before the utterance when the speaker has a state (online-ivector-feature.cc:
// 445-453), GetAdaptationState after it (:386-396: OnlineCmvn::GetState of the last frame, the stats, LimitFrames :109-127)
int ref_ivector_run_speaker(void *h, const float *feats, int T, int D, const double *global_cmvn, int cmn_window,
                            int speaker_frames, int global_frames, int splice_left, int splice_right, int num_gselect,
                            float min_post, float posterior_scale, float max_count, int num_cg_iters,
                            int online_cmvn_iextractor, const int *sched, int n_chunks, float *out,
                            float *debug_lda_raw, float *debug_lda_norm, void *speaker, float max_remembered_frames) {
  try {
    RefSpeaker *spk = (RefSpeaker *)speaker;
    RefIvec *r = (RefIvec *)h;
    Matrix<BaseFloat> m(T, D);
    for (int t = 0; t < T; t++) memcpy(m.RowData(t), feats + (size_t)t * D, 4 * D);
    OnlineMatrixFeature base(m);
    OnlineCmvnOptions copts;
    copts.cmn_window = cmn_window; copts.speaker_frames = speaker_frames; copts.global_frames = global_frames;
    Matrix<double> g(2, D + 1);
    for (int i = 0; i < 2; i++) for (int j = 0; j <= D; j++) g(i, j) = global_cmvn[i * (D + 1) + j];
    OnlineCmvnState cstate(g);
    if (spk && spk->has) { cstate = spk->cmvn; cstate.global_cmvn_stats = g; }
    OnlineCmvn cmvn(copts, cstate, &base);
    OnlineSpliceOptions sopts;
    sopts.left_context = splice_left; sopts.right_context = splice_right;
    OnlineSpliceFrames splice_norm(sopts, &cmvn), splice_raw(sopts, &base);
    OnlineTransform lda_norm(r->lda, &splice_norm), lda_raw(r->lda, &splice_raw);
    const int ivdim = r->extractor.IvectorDim();
    OnlineIvectorEstimationStats stats(ivdim, r->extractor.PriorOffset(), max_count);
    if (spk && spk->has) stats = *spk->stats;
    Vector<double> current_ivector(ivdim);
    int num_frames_stats = 0;
    if (debug_lda_raw) for (int t = 0; t < T; t++) { SubVector<BaseFloat> row(debug_lda_raw + (size_t)t * lda_raw.Dim(), lda_raw.Dim()); lda_raw.GetFrame(t, &row); }
    if (debug_lda_norm) for (int t = 0; t < T; t++) { SubVector<BaseFloat> row(debug_lda_norm + (size_t)t * lda_norm.Dim(), lda_norm.Dim()); lda_norm.GetFrame(t, &row); }
    for (int n = 0; n < n_chunks; n++) {
      int frame = sched[n];
      if (frame < 0) {   // no i-vector frame ready: GetFrame is not called and the i-vector stays zero (decodable-online-looped.cc:188-197)
        for (int d = 0; d < ivdim; d++) out[(size_t)n * ivdim + d] = 0.0f;
        continue;
      }
      // UpdateStatsUntilFrame(frame) with use_most_recent_ivector = true (:248-281)
      std::vector<int32> frames;
      for (; num_frames_stats <= frame; num_frames_stats++) frames.push_back(num_frames_stats);
      if (!frames.empty()) {
        // UpdateStatsForFrames (:201-245), all frame weights 1.0
        int nf = frames.size();
        Matrix<BaseFloat> fe(nf, lda_norm.Dim()), log_likes;
        lda_norm.GetFrames(frames, &fe);
        r->ubm.LogLikelihoods(fe, &log_likes);
        std::vector<std::vector<std::pair<int32, BaseFloat> > > post(nf);
        BaseFloat mp = min_post;           // GetMinPost(weight = 1.0) (:188-199)
        if (mp > 0.99) mp = 0.99;
        for (int i = 0; i < nf; i++) {
          VectorToPosteriorEntry(log_likes.Row(i), num_gselect, mp, &post[i]);
          for (size_t j = 0; j < post[i].size(); j++) post[i][j].second *= posterior_scale * 1.0;
        }
        if (!online_cmvn_iextractor) lda_raw.GetFrames(frames, &fe); else lda_norm.GetFrames(frames, &fe);
        stats.AccStats(r->extractor, fe, post);
        stats.GetIvector(num_cg_iters, &current_ivector);
      }
      // GetFrame (:343-349)
      for (int d = 0; d < ivdim; d++) out[(size_t)n * ivdim + d] = (float)current_ivector(d);
      out[(size_t)n * ivdim] = (float)current_ivector(0);
      { Vector<BaseFloat> f(ivdim); f.CopyFromVec(current_ivector); f(0) -= r->extractor.PriorOffset();
        for (int d = 0; d < ivdim; d++) out[(size_t)n * ivdim + d] = f(d); }
    }
    if (spk) {
      OnlineCmvnState ns;
      cmvn.GetState(cmvn.NumFramesReady() - 1, &ns);
      // LimitFrames (:109-127), restated with the reference's own Scale() methods
      if (max_remembered_frames >= 0.0f) {
        if (ns.speaker_cmvn_stats.NumRows() != 0) {
          int32 feat_dim = ns.speaker_cmvn_stats.NumCols() - 1;
          BaseFloat count = ns.speaker_cmvn_stats(0, feat_dim);
          if (count > max_remembered_frames) ns.speaker_cmvn_stats.Scale(max_remembered_frames / count);
        }
        BaseFloat max_remembered_frames_scaled = max_remembered_frames * posterior_scale;
        if (stats.Count() > max_remembered_frames_scaled) stats.Scale(max_remembered_frames_scaled / stats.Count());
      }
      spk->cmvn = ns;
      spk->stats.reset(new OnlineIvectorEstimationStats(stats));
      spk->has = true;
    }
    return 0;
  } catch (const std::exception &e) { fprintf(stderr, "ref_ivector_run: %s\n", e.what()); return -1; }
}

// IvectorExtractor::Write (ivector/ivector-extractor.cc:807) and DiagGmm::Write (gmm/diag-gmm.cc:728) with the
// Kaldi binary header, as final.ie / final.dubm are stored: what kaldi_b200/kaldi_io.py's readers are pinned against.
int ref_ivector_write(void *h, const char *extractor_path, const char *ubm_path, int binary) {
  try {
    RefIvec *r = (RefIvec *)h;
    { Output ko(extractor_path, binary != 0); r->extractor.Write(ko.Stream(), binary != 0); if (!ko.Close()) return -1; }
    { Output ko(ubm_path, binary != 0); r->ubm.Write(ko.Stream(), binary != 0); if (!ko.Close()) return -1; }
    return 0;
  } catch (const std::exception &e) { fprintf(stderr, "ref_ivector_write: %s\n", e.what()); return -1; }
}

}  // extern "C"
