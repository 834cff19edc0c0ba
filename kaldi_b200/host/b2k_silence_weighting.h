// b2k_silence_weighting.h -- which feature frames count for the i-vector, decided from the decoder's live best path.
//
// Host-only C++ (no Kaldi, no CUDA types): the bookkeeping of OnlineSilenceWeighting (online2/online-ivector-feature.h:
// 460-571, online-ivector-feature.cc:465-750).  The online2 tools call, after every chunk,
//     ComputeCurrentTraceback(decoder); GetDeltaWeights(NumFramesReady(), first_decoder_frame, &delta);
//     feature_pipeline.UpdateFrameWeights(delta);                      (online2bin/online2-wav-nnet3-latgen-faster.cc:254-262)
// so that frames the traceback calls silence (or a transition-id that repeats for too long) weigh `silence_weight` in the
// i-vector statistics, and a frame whose label changed in a later traceback has the difference sent after it.
//
// The reference walks the best path backwards through the decoder's tokens and stops at the first frame whose token is
// the one it saw there last time.  A token of the lattice decoder is one (frame, HCLG state) pair for the lifetime of
// the utterance -- a frame's tokens are hashed by state and never created again once the frame is done
// (decoder/lattice-faster-decoder.cc:258-300) -- so the state is the token's identity here: b2k_dec_best_path hands out the
// states along the path, and the walk below stops where the reference's would, including the case where the arc that LEAVES
// an unchanged token did change (the reference keeps the old transition-id for that frame; so does this).
//
// Parity: tests/test_silence_weighting.py drives this class and the reference's own class (compiled in oracle/_ref, fed by
// a replay decoder) with the same sequences of tracebacks.
#ifndef B2K_SILENCE_WEIGHTING_H_
#define B2K_SILENCE_WEIGHTING_H_

#include <algorithm>
#include <cstdint>
#include <cstdlib>
#include <stdexcept>
#include <string>
#include <unordered_set>
#include <utility>
#include <vector>

namespace b2k_host {

class SilenceWeighting {
 public:
  static constexpr int32_t kStartToken = -2;   // the token the path starts from (the decoder's one start token)

  // tid2phone[t] = TransitionIdToPhone(t) for t = 1..NumTransitionIds (entry 0 unused): b2k_model_tid2phone
  // silence_phones: "1:2:3" or "1,2,3" (--ivector-silence-weighting.silence-phones); max_state_duration <= 0: no limit
  SilenceWeighting(std::vector<int32_t> tid2phone, const std::string &silence_phones, float silence_weight,
                   float max_state_duration, int32_t frame_subsampling_factor = 1)
      : tid2phone_(std::move(tid2phone)), silence_weight_(silence_weight),
        max_run_(static_cast<int32_t>(max_state_duration)), fs_(frame_subsampling_factor),
        active_(!silence_phones.empty() && silence_weight != 1.0f) {
    if (fs_ < 1) throw std::invalid_argument("SilenceWeighting: frame_subsampling_factor must be >= 1");
    // SplitStringToIntegers(str, ":,", false, &phones) whose result the reference does not look at
    // (online-ivector-feature.cc:474-476): ONE field that is empty or not an integer leaves the list empty.
    size_t at = 0;
    while (at <= silence_phones.size() && !silence_phones.empty()) {
      size_t end = silence_phones.find_first_of(":,", at);
      if (end == std::string::npos) end = silence_phones.size();
      const std::string field = silence_phones.substr(at, end - at);
      char *stop = NULL;
      const long long v = std::strtoll(field.c_str(), &stop, 10);     // blanks before the digits pass, after them do not
      if (stop == field.c_str() || *stop != 0 || v != static_cast<int32_t>(v)) {        // (util/text-utils.h:83-98)
        silence_.clear();
        list_ok_ = false;
        break;
      }
      silence_.insert(static_cast<int32_t>(v));
      at = end + 1;
    }
  }

  // false: the silence list did not parse and is EMPTY, as in the reference (which says nothing); callers may want to say so
  bool SilencePhonesParsed() const { return list_ok_; }
  bool Active() const { return active_; }

  // The current best path, one entry per decoded frame in time order: the transition-id of the frame's (emitting) arc and
  // the HCLG state that arc leaves (kStartToken for a path's very first token when it is the start token).
  // = ComputeCurrentTraceback (online-ivector-feature.cc:480-530)
  void SetTraceback(const int32_t *tids, const int32_t *source_states, int32_t num_frames_decoded) {
    const int32_t had = static_cast<int32_t>(frames_.size());
    if (had < num_frames_decoded) frames_.resize(num_frames_decoded);
    else if (had > num_frames_decoded && frames_[num_frames_decoded].tid != -1)
      throw std::runtime_error("SilenceWeighting: number of frames decoded decreased");
    for (int32_t f = num_frames_decoded - 1; f >= 0; f--) {
      Frame &fr = frames_[f];
      if (fr.seen && fr.source == source_states[f]) break;    // same token as last time: nothing behind it has changed
      fr.seen = true;
      fr.source = source_states[f];
      fr.tid = tids[f];
    }
  }

  // From a best path as b2k_dec_best_path returns it (arcs start first, epsilons included, arc_state = the state an arc
  // enters): picks the emitting arcs and the states they leave.  Returns false when the path does not hold one emitting arc
  // per decoded frame (nothing is changed then).
  bool SetTracebackFromPath(const int32_t *ilabels, const int32_t *arc_state, int32_t n_arcs, int32_t num_frames_decoded) {
    tids_.clear(); sources_.clear();
    for (int32_t k = 0; k < n_arcs; k++)
      if (ilabels[k] != 0) {
        tids_.push_back(ilabels[k]);
        sources_.push_back(k == 0 ? kStartToken : arc_state[k - 1]);
      }
    if (static_cast<int32_t>(tids_.size()) != num_frames_decoded) return false;
    SetTraceback(tids_.data(), sources_.data(), num_frames_decoded);
    return true;
  }

  // = GetDeltaWeights (online-ivector-feature.cc:594-700).  num_frames_ready and the frames of `delta_weights` are feature
  // frames; everything in between is in decoder frames (frame_subsampling_factor feature frames each).
  void GetDeltaWeights(int32_t num_frames_ready, int32_t first_decoder_frame,
                       std::vector<std::pair<int32_t, float> > *delta_weights) {
    if (!(num_frames_ready > first_decoder_frame || num_frames_ready == 0))
      throw std::invalid_argument("SilenceWeighting: num_frames_ready must exceed first_decoder_frame");
    delta_weights->clear();
    const int32_t begin = Extend(num_frames_ready, first_decoder_frame, 100);    // not further back than 100 frames
    const int32_t end = static_cast<int32_t>(frames_.size()), count = end - begin;
    if (count == 0) return;
    weights_.assign(count, 1.0f);
    if (frames_[begin].tid == -1) {
      // no traceback anywhere in the range: repeat the last weight that went out (silence, when none did)
      const float w = begin == 0 ? silence_weight_ : frames_[begin - 1].weight;
      std::fill(weights_.begin(), weights_.end(), w);
    } else {
      int32_t run_start = 0;
      for (int32_t i = 0; i < count; i++) {
        const int32_t tid = frames_[begin + i].tid;
        if (tid == -1) {                       // decoded later than the traceback: guess what the frame before it got
          weights_[i] = weights_[i - 1];
          continue;
        }
        if (IsSilence(tid)) weights_[i] = silence_weight_;
        const bool run_ends = i + 1 == count || frames_[begin + i + 1].tid != tid;
        if (max_run_ > 0 && run_ends) {
          if (i - run_start + 1 >= max_run_)   // one transition-id for that long: not speech, whatever the phone
            std::fill(weights_.begin() + run_start, weights_.begin() + i + 1, silence_weight_);
          if (i + 1 < count) run_start = i + 1;
        }
      }
    }
    for (int32_t i = 0; i < count; i++) {
      Frame &fr = frames_[begin + i];
      const float diff = weights_[i] - fr.weight;
      fr.weight = weights_[i];
      // the newest frame is always reported, changed or not (debugging code downstream looks at it)
      if (diff != 0.0f || i + 1 == count)
        for (int32_t s = 0; s < fs_; s++)
          delta_weights->push_back(std::make_pair(first_decoder_frame + (begin + i) * fs_ + s, diff));
    }
  }
  void GetDeltaWeights(int32_t num_frames_ready, std::vector<std::pair<int32_t, float> > *delta_weights) {
    GetDeltaWeights(num_frames_ready, 0, delta_weights);
  }

  // = GetNonsilenceFrames (online-ivector-feature.cc:702-750): the decoder frames of the last 500 that the traceback calls speech
  void GetNonsilenceFrames(int32_t num_frames_ready, int32_t first_decoder_frame, std::vector<int32_t> *frames) {
    if (!(num_frames_ready > first_decoder_frame || num_frames_ready == 0))
      throw std::invalid_argument("SilenceWeighting: num_frames_ready must exceed first_decoder_frame");
    frames->clear();
    const int32_t begin = Extend(num_frames_ready, first_decoder_frame, 500);
    for (int32_t f = begin; f < static_cast<int32_t>(frames_.size()); f++)
      if (frames_[f].tid != -1 && !IsSilence(frames_[f].tid)) frames->push_back(f);
  }

 private:
  struct Frame {
    bool seen = false;        // a traceback has passed here
    int32_t source = -1;      // the token (state) the frame's arc left, at the traceback that wrote `tid`
    int32_t tid = -1;         // -1: no traceback yet
    float weight = 0.0f;      // what the i-vector statistics have been told so far
  };

  // makes room for the decoder frames up to num_frames_ready (rounded up) and returns where the output range starts:
  // `lookback` frames before what was there
  int32_t Extend(int32_t num_frames_ready, int32_t first_decoder_frame, int32_t lookback) {
    const int32_t want = (num_frames_ready - first_decoder_frame + fs_ - 1) / fs_, had = static_cast<int32_t>(frames_.size());
    if (had < want) frames_.resize(want);
    return std::max<int32_t>(0, had - lookback);
  }
  bool IsSilence(int32_t tid) const {
    if (tid <= 0 || tid >= static_cast<int32_t>(tid2phone_.size())) throw std::out_of_range("SilenceWeighting: transition-id out of range");
    return silence_.count(tid2phone_[tid]) != 0;
  }

  std::vector<int32_t> tid2phone_;
  std::unordered_set<int32_t> silence_;
  float silence_weight_;
  int32_t max_run_, fs_;
  bool active_, list_ok_ = true;
  std::vector<Frame> frames_;
  std::vector<float> weights_;
  std::vector<int32_t> tids_, sources_;
};

}  // namespace b2k_host

#endif  // B2K_SILENCE_WEIGHTING_H_
