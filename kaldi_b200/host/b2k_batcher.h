// b2k_batcher.h — host logic between the reference's chunked, correlation-id based entry point
// (cuda_decoder::BatchedThreadedNnet3CudaOnlinePipeline::DecodeBatch, cudadecoder/batched-threaded-nnet3-cuda-online-
// pipeline.h:209-215: per call up to max_batch_size chunks, each tagged first/last) and the whole-utterance batches the
// b2k pipeline decodes (b2k_pipeline_decode_batch: n utterances of ONE length).  Header-only C++17, no Kaldi and no
// CUDA types: the device work sits behind the Backend template parameter, so that this logic is unit-tested on the
// CPU with a mock backend (tests/test_batcher_cpp.py).
//
// Behaviour:
//  * InitCorrId / the first chunk opens an utterance; chunks are appended to a host buffer; the last chunk closes it;
//  * closed utterances are bucketed by their exact number of samples (no padding: padding changes the features of the
//    last frames and with them the lattice); a bucket is decoded as soon as it holds max_batch utterances, and
//    Flush() decodes what is left (partial batches, buckets in increasing length, utterances in arrival order);
//  * results are delivered through the per-utterance callback in batch order, from the calling thread.
// Differences from the reference pipeline, by design: nothing is decoded before an utterance is complete (the b2k
// decoder keeps the CPU tool's whole-utterance semantics), so partial hypotheses / endpointing are
// not offered here, and callbacks run synchronously instead of on a thread pool.
#ifndef B2K_BATCHER_H_
#define B2K_BATCHER_H_

#include <algorithm>
#include <cstdint>
#include <deque>
#include <functional>
#include <map>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <utility>
#include <vector>

namespace b2k_host {

// Backend concept:
//   struct Result;                                                       // one utterance's output
//   void Decode(int64_t num_samples, const std::vector<const float *> &waves, std::vector<Result> *out);
// Decode receives 1..max_batch utterances of exactly num_samples samples each and fills one Result per utterance.
template <class Backend>
class UtteranceBatcher {
 public:
  typedef uint64_t CorrelationID;
  typedef typename Backend::Result Result;
  typedef std::function<void(CorrelationID, Result &)> Callback;

  UtteranceBatcher(Backend *backend, int max_batch, int max_open_utterances)
      : backend_(backend), max_batch_(max_batch), max_open_(max_open_utterances) {
    if (!backend || max_batch <= 0 || max_open_utterances <= 0) throw std::invalid_argument("UtteranceBatcher: bad arguments");
  }

  // TryInitCorrID (…online-pipeline.h:170): false when no slot is free or the id is already open
  bool TryInitCorrId(CorrelationID id) {
    if (open_.count(id) || (int)open_.size() >= max_open_) return false;
    open_.emplace(id, Utterance());
    return true;
  }
  void SetCallback(CorrelationID id, Callback cb) {
    auto it = open_.find(id);
    if (it == open_.end()) throw std::invalid_argument("SetCallback: unknown correlation id " + std::to_string(id));
    it->second.callback = std::move(cb);
  }
  // a default for utterances that set none
  void SetDefaultCallback(Callback cb) { default_callback_ = std::move(cb); }

  // One DecodeBatch call of the reference: chunk i belongs to corr_ids[i].  Returns the number of utterances decoded
  // (and called back) during this call.
  int AcceptChunks(const std::vector<CorrelationID> &corr_ids, const std::vector<std::pair<const float *, int64_t>> &chunks,
                   const std::vector<bool> &is_first_chunk, const std::vector<bool> &is_last_chunk) {
    const size_t n = corr_ids.size();
    if (chunks.size() != n || is_first_chunk.size() != n || is_last_chunk.size() != n)
      throw std::invalid_argument("AcceptChunks: argument sizes differ");
    // validate everything before touching any state (the reference asserts; here a bad call changes nothing)
    int implicit_new = 0;
    for (size_t i = 0; i < n; i++) {
      for (size_t j = 0; j < i; j++) if (corr_ids[j] == corr_ids[i]) throw std::invalid_argument("AcceptChunks: a correlation id appears twice in one batch");
      auto it = open_.find(corr_ids[i]);
      const bool known = it != open_.end(), started = known && it->second.started;
      if (is_first_chunk[i]) {
        if (started) throw std::invalid_argument("AcceptChunks: first chunk of an utterance that already has audio");
        if (!known) {                                          // first chunk without TryInitCorrId: opened implicitly
          if ((int)open_.size() + implicit_new >= max_open_) throw std::runtime_error("AcceptChunks: no free utterance slot");
          implicit_new++;
        }
      } else if (!started) {
        throw std::invalid_argument("AcceptChunks: chunk for a correlation id that was not started");
      }
      if (chunks[i].second < 0 || (chunks[i].second > 0 && !chunks[i].first)) throw std::invalid_argument("AcceptChunks: bad chunk");
    }
    int decoded = 0;
    for (size_t i = 0; i < n; i++) {
      Utterance &u = open_[corr_ids[i]];                       // creates the entry for an implicit first chunk
      u.started = true;
      u.samples.insert(u.samples.end(), chunks[i].first, chunks[i].first + chunks[i].second);
      if (is_last_chunk[i]) {
        const int64_t len = (int64_t)u.samples.size();
        std::deque<Closed> &b = buckets_[len];
        b.push_back(Closed{corr_ids[i], std::move(u.samples), std::move(u.callback)});
        open_.erase(corr_ids[i]);
        if ((int)b.size() >= max_batch_) decoded += DecodeBucket(len, max_batch_);
      }
    }
    return decoded;
  }

  // WaitForLatticeCallbacks (…online-pipeline.h:274): decodes every closed utterance that is still waiting
  int Flush() {
    int decoded = 0;
    while (!buckets_.empty()) {
      const int64_t len = buckets_.begin()->first;
      const int k = (int)std::min<size_t>(buckets_.begin()->second.size(), (size_t)max_batch_);
      decoded += DecodeBucket(len, k);
    }
    return decoded;
  }

  int NumOpen() const { return (int)open_.size(); }
  int NumWaiting() const { int n = 0; for (auto &kv : buckets_) n += (int)kv.second.size(); return n; }

 private:
  struct Utterance { std::vector<float> samples; Callback callback; bool started = false; };
  struct Closed { CorrelationID id; std::vector<float> samples; Callback callback; };

  int DecodeBucket(int64_t len, int k) {
    std::deque<Closed> &b = buckets_[len];
    std::vector<Closed> batch;
    for (int i = 0; i < k; i++) { batch.push_back(std::move(b.front())); b.pop_front(); }
    if (b.empty()) buckets_.erase(len);
    std::vector<const float *> waves;
    for (auto &c : batch) waves.push_back(c.samples.data());
    std::vector<Result> results;
    backend_->Decode(len, waves, &results);
    if ((int)results.size() != k) throw std::runtime_error("UtteranceBatcher: the backend returned a different number of results");
    for (int i = 0; i < k; i++) {
      if (batch[i].callback) batch[i].callback(batch[i].id, results[i]);
      else if (default_callback_) default_callback_(batch[i].id, results[i]);
    }
    return k;
  }

  Backend *backend_;
  int max_batch_, max_open_;
  std::unordered_map<CorrelationID, Utterance> open_;
  std::map<int64_t, std::deque<Closed>> buckets_;
  Callback default_callback_;
};

}  // namespace b2k_host
#endif  // B2K_BATCHER_H_
