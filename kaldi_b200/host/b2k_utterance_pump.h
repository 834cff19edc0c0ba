// b2k_utterance_pump.h -- whole utterances through a chunk-at-a-time pipeline.
//
// Host-only C++ (no Kaldi, no CUDA types).  The role of BatchedThreadedNnet3CudaPipeline2's control thread
// (cudadecoder/batched-threaded-nnet3-cuda-pipeline2.cc:383-520: AcquireTasks / BuildBatchFromCurrentTasks / ComputeTasks over the
// online pipeline it owns): up to max_batch utterances are "current"; every step sends the next chunk of each of them -- at
// most one chunk per utterance and batch, first / last flags as the online pipeline wants them -- and utterances that have sent
// their last chunk make room for waiting ones.  The reference runs this on its own thread while the caller keeps submitting;
// here it runs on the caller's thread: Run(false) whenever a full batch of utterances is available (so that submitting a
// data set keeps at most max_batch - 1 + what one call adds in memory), Run(true) to finish what is left.
//
// Pipeline: void DecodeBatch(const std::vector<uint64_t> &ids, const std::vector<std::pair<const float *, int64_t>> &chunks,
//                            const std::vector<bool> &is_first_chunk, const std::vector<bool> &is_last_chunk);
// Unit-tested with a mock pipeline (tests/cabi/utterance_pump_test.cc).
#ifndef B2K_UTTERANCE_PUMP_H_
#define B2K_UTTERANCE_PUMP_H_

#include <algorithm>
#include <cstdint>
#include <deque>
#include <functional>
#include <stdexcept>
#include <utility>
#include <vector>

namespace b2k_host {

template <class Pipeline>
class UtterancePump {
 public:
  UtterancePump(Pipeline *pipeline, int max_batch, int64_t samples_per_chunk)
      : pipeline_(pipeline), max_batch_(max_batch), chunk_(samples_per_chunk) {
    if (max_batch < 1 || samples_per_chunk < 1) throw std::invalid_argument("UtterancePump: max_batch and samples_per_chunk must be positive");
  }

  // `samples` must stay valid until `done` has been called (right after the batch with the utterance's last chunk; it may be empty).
  // An utterance without samples is refused: the reference's pipeline drops it before it becomes a task (:218).
  void Add(uint64_t id, const float *samples, int64_t num_samples, std::function<void()> done = std::function<void()>()) {
    if (num_samples <= 0) throw std::invalid_argument("UtterancePump: an utterance needs at least one sample");
    waiting_.push_back(Utterance{id, samples, num_samples, 0, std::move(done)});
  }

  size_t NumUnfinished() const { return waiting_.size() + current_.size(); }

  // drain = false: steps while a FULL batch of utterances is there; drain = true: until nothing is left
  void Run(bool drain) {
    while (drain ? NumUnfinished() > 0 : NumUnfinished() >= static_cast<size_t>(max_batch_)) Step();
  }

 private:
  struct Utterance {
    uint64_t id;
    const float *samples;
    int64_t num_samples, sent;
    std::function<void()> done;
  };

  void Step() {
    while (current_.size() < static_cast<size_t>(max_batch_) && !waiting_.empty()) {
      current_.push_back(std::move(waiting_.front()));
      waiting_.pop_front();
    }
    ids_.clear(); chunks_.clear(); first_.clear(); last_.clear();
    for (Utterance &u : current_) {
      const int64_t n = std::min(chunk_, u.num_samples - u.sent);
      ids_.push_back(u.id);
      chunks_.push_back(std::make_pair(u.samples + u.sent, n));
      first_.push_back(u.sent == 0);
      u.sent += n;
      last_.push_back(u.sent == u.num_samples);
    }
    pipeline_->DecodeBatch(ids_, chunks_, first_, last_);
    // finished utterances leave; the others keep their order (a stream's place in the batch does not matter to the pipeline)
    size_t keep = 0;
    for (size_t i = 0; i < current_.size(); i++) {
      if (current_[i].sent == current_[i].num_samples) {
        if (current_[i].done) current_[i].done();
      } else {
        if (keep != i) current_[keep] = std::move(current_[i]);
        keep++;
      }
    }
    current_.resize(keep);
  }

  Pipeline *pipeline_;
  int max_batch_;
  int64_t chunk_;
  std::deque<Utterance> waiting_;
  std::vector<Utterance> current_;
  std::vector<uint64_t> ids_;
  std::vector<std::pair<const float *, int64_t> > chunks_;
  std::vector<bool> first_, last_;
};

}  // namespace b2k_host

#endif  // B2K_UTTERANCE_PUMP_H_
