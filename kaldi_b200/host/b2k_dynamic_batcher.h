// b2k_dynamic_batcher.h — the serving front between many producer threads that push audio chunks of many streams and a
// streaming pipeline that decodes one batch of chunks per call: the role of cuda_decoder::CudaOnlinePipelineDynamicBatcher
// (cudadecoder/cuda-online-pipeline-dynamic-batcher.{h,cc}) in front of BatchedThreadedNnet3CudaOnlinePipeline::DecodeBatch.
// Header-only C++17, no Kaldi and no CUDA types: the pipeline is a template parameter, so the scheduling is unit-tested on
// the CPU with a mock pipeline (tests/cabi/dynamic_batcher_test.cc); kaldi::b2k_shim::StreamingOnlinePipelineB2k fits the
// concept through DynamicBatcherPipelineAdapter in b2k_kaldi_shims.h.
//
// Contract (the reference's, :52-60 of its header and the .cc):
//  * Push(corr_id, is_first_chunk, is_last_chunk, samples) may be called from any thread; the samples are copied;
//  * a batch holds at most max_batch_size chunks and at most ONE chunk per stream; a stream's chunks are decoded in the order
//    they were pushed; a first chunk enters a batch only when the pipeline has a free channel for it (TryInitCorrID);
//  * a worker thread runs a batch as soon as max_batch_size chunks are waiting, or when `timeout` seconds have passed since
//    the previous batch and anything is waiting;
//  * WaitForCompletion() returns when everything pushed so far has been decoded; GetNumPendingChunks(corr_id).
// Design of our own: one FIFO per stream and a ready list of streams instead of a backlog list that is rescanned, and a
// condition variable instead of a 100-microsecond polling loop.
//
// Pipeline concept:
//   int  MaxBatchSize() const;
//   bool TryInitCorrID(uint64_t corr_id);                              // false: no free channel now, try again later
//   void DecodeBatch(const std::vector<uint64_t> &corr_ids, const std::vector<std::pair<const float *, int64_t>> &chunks,
//                    const std::vector<bool> &is_first_chunk, const std::vector<bool> &is_last_chunk);
#ifndef B2K_DYNAMIC_BATCHER_H_
#define B2K_DYNAMIC_BATCHER_H_

#include <chrono>
#include <condition_variable>
#include <cstdint>
#include <deque>
#include <exception>
#include <stdexcept>
#include <mutex>
#include <string>
#include <thread>
#include <unordered_map>
#include <utility>
#include <vector>

namespace b2k_host {

template <class Pipeline>
class DynamicBatcher {
 public:
  typedef uint64_t CorrelationID;

  DynamicBatcher(Pipeline *pipeline, double timeout_seconds = 2e-3)
      : pipeline_(pipeline), max_batch_(pipeline->MaxBatchSize()), timeout_(timeout_seconds) {
    if (max_batch_ <= 0) throw std::invalid_argument("DynamicBatcher: the pipeline's max batch size must be positive");
    worker_ = std::thread(&DynamicBatcher::Loop, this);
  }
  ~DynamicBatcher() {
    {
      std::lock_guard<std::mutex> lk(m_);
      stop_ = true;
    }
    wake_.notify_all();
    worker_.join();
  }
  DynamicBatcher(const DynamicBatcher &) = delete;
  DynamicBatcher &operator=(const DynamicBatcher &) = delete;

  void Push(CorrelationID corr_id, bool is_first_chunk, bool is_last_chunk, const float *samples, int64_t num_samples) {
    Chunk c;
    c.first = is_first_chunk; c.last = is_last_chunk;
    if (num_samples > 0) c.samples.assign(samples, samples + num_samples);
    {
      std::lock_guard<std::mutex> lk(m_);
      Stream &s = streams_[corr_id];
      if (s.chunks.empty() && !s.in_flight) ready_.push_back(corr_id);       // the stream has nothing scheduled: it becomes ready
      s.chunks.push_back(std::move(c));
      pending_++;
    }
    wake_.notify_all();
  }

  // Everything pushed before this call has been decoded when it returns.  An exception thrown by the pipeline on the worker
  // thread is rethrown here (and by the next Push-free call), once.
  void WaitForCompletion() {
    std::unique_lock<std::mutex> lk(m_);
    done_.wait(lk, [this] { return pending_ == 0 || error_; });
    if (error_) { std::exception_ptr e = error_; error_ = nullptr; std::rethrow_exception(e); }
  }
  int GetNumPendingChunks(CorrelationID corr_id) {
    std::lock_guard<std::mutex> lk(m_);
    auto it = streams_.find(corr_id);
    return it == streams_.end() ? 0 : static_cast<int>(it->second.chunks.size()) + (it->second.in_flight ? 1 : 0);
  }
  int64_t NumBatchesRun() {
    std::lock_guard<std::mutex> lk(m_);
    return batches_;
  }

 private:
  struct Chunk { bool first = false, last = false; std::vector<float> samples; };
  struct Stream { std::deque<Chunk> chunks; bool in_flight = false; bool started = false; };

  // m_ held.  Streams are taken from the ready list in the order they became ready; a first chunk that finds no free channel
  // keeps its place at the front of the next attempt.
  void FillBatch(std::vector<CorrelationID> *ids, std::vector<Chunk> *chunks) {
    std::deque<CorrelationID> retry;
    while (!ready_.empty() && static_cast<int>(ids->size()) < max_batch_) {
      const CorrelationID id = ready_.front();
      ready_.pop_front();
      Stream &s = streams_[id];
      if (s.chunks.front().first && !s.started) {
        if (!pipeline_->TryInitCorrID(id)) { retry.push_back(id); continue; }
        s.started = true;
      }
      ids->push_back(id);
      chunks->push_back(std::move(s.chunks.front()));
      s.chunks.pop_front();
      s.in_flight = true;
    }
    for (auto it = retry.rbegin(); it != retry.rend(); ++it) ready_.push_front(*it);
  }

  void Loop() {
    using clock = std::chrono::steady_clock;
    const auto period = std::chrono::duration_cast<clock::duration>(std::chrono::duration<double>(timeout_));
    auto deadline = clock::now() + period;
    std::unique_lock<std::mutex> lk(m_);
    for (;;) {
      // sleep until a full batch is waiting, the timeout has come with something waiting, or the batcher is being destroyed
      while (!stop_ && static_cast<int>(ready_.size()) < max_batch_ && !(clock::now() >= deadline && !ready_.empty())) {
        if (ready_.empty()) wake_.wait(lk);                     // nothing to time out on
        else wake_.wait_until(lk, deadline);
        if (ready_.empty()) deadline = clock::now() + period;   // the timeout counts from the moment something waits
      }
      if (stop_) return;
      std::vector<CorrelationID> ids;
      std::vector<Chunk> chunks;
      FillBatch(&ids, &chunks);
      if (ids.empty()) {                                       // only first chunks without a free channel: wait for one to end
        deadline = clock::now() + period;
        wake_.wait_until(lk, deadline);
        continue;
      }
      lk.unlock();
      std::vector<std::pair<const float *, int64_t>> views;
      std::vector<bool> first, last;
      for (const Chunk &c : chunks) {
        views.push_back({c.samples.data(), static_cast<int64_t>(c.samples.size())});
        first.push_back(c.first); last.push_back(c.last);
      }
      std::exception_ptr err;
      try {
        pipeline_->DecodeBatch(ids, views, first, last);
      } catch (...) {
        err = std::current_exception();
      }
      lk.lock();
      batches_++;
      for (size_t i = 0; i < ids.size(); i++) {
        auto it = streams_.find(ids[i]);
        Stream &s = it->second;
        s.in_flight = false;
        pending_--;
        if (chunks[i].last) s.started = false;                 // the channel is free again; a later utterance may reuse the id
        if (!s.chunks.empty()) ready_.push_back(ids[i]);       // its next chunk takes its turn behind the streams already waiting
        else if (chunks[i].last) streams_.erase(it);
      }
      if (err && !error_) error_ = err;
      deadline = clock::now() + period;
      done_.notify_all();
    }
  }

  Pipeline *pipeline_;
  const int max_batch_;
  const double timeout_;
  std::mutex m_;
  std::condition_variable wake_, done_;
  std::unordered_map<CorrelationID, Stream> streams_;
  std::deque<CorrelationID> ready_;                             // streams with a chunk to schedule and none in flight
  int64_t pending_ = 0, batches_ = 0;
  bool stop_ = false;
  std::exception_ptr error_;
  std::thread worker_;
};

}  // namespace b2k_host

#endif  // B2K_DYNAMIC_BATCHER_H_
