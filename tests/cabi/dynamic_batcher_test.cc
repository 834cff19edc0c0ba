// tests/cabi/dynamic_batcher_test.cc — b2k_host::DynamicBatcher (kaldi_b200/host/b2k_dynamic_batcher.h) with a mock
// pipeline: the properties cuda_decoder::CudaOnlinePipelineDynamicBatcher promises its callers.
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <map>
#include <random>
#include <set>
#include <stdexcept>
#include <thread>

#include "b2k_dynamic_batcher.h"

#define REQUIRE(x) do { if (!(x)) { std::fprintf(stderr, "%s:%d: %s\n", __FILE__, __LINE__, #x); std::exit(1); } } while (0)

struct MockPipeline {
  int max_batch, channels;
  std::set<uint64_t> live;                       // streams holding a channel
  std::map<uint64_t, std::vector<float>> audio;   // what each stream received, in order
  std::map<uint64_t, int> chunks_seen;
  std::vector<size_t> batch_sizes;
  bool saw_duplicate = false, saw_bad_order = false, saw_uninitialised = false, saw_oversize = false;
  int throw_on_batch = -1;
  MockPipeline(int mb, int ch) : max_batch(mb), channels(ch) {}
  int MaxBatchSize() const { return max_batch; }
  bool TryInitCorrID(uint64_t id) {
    if ((int)live.size() >= channels) return false;
    live.insert(id);
    return true;
  }
  void DecodeBatch(const std::vector<uint64_t> &ids, const std::vector<std::pair<const float *, int64_t>> &chunks,
                   const std::vector<bool> &first, const std::vector<bool> &last) {
    if ((int)batch_sizes.size() == throw_on_batch) { batch_sizes.push_back(ids.size()); throw std::runtime_error("device fell over"); }
    batch_sizes.push_back(ids.size());
    if ((int)ids.size() > max_batch) saw_oversize = true;
    std::set<uint64_t> in_batch;
    for (size_t i = 0; i < ids.size(); i++) {
      if (!in_batch.insert(ids[i]).second) saw_duplicate = true;
      if (!live.count(ids[i])) saw_uninitialised = true;
      if (first[i] != (chunks_seen[ids[i]] == 0)) saw_bad_order = true;
      // a chunk carries its own sequence number in its first sample
      if (chunks[i].second > 0 && (int)chunks[i].first[0] != chunks_seen[ids[i]]) saw_bad_order = true;
      chunks_seen[ids[i]]++;
      audio[ids[i]].insert(audio[ids[i]].end(), chunks[i].first, chunks[i].first + chunks[i].second);
      if (last[i]) { live.erase(ids[i]); chunks_seen.erase(ids[i]); }
    }
    std::this_thread::sleep_for(std::chrono::microseconds(200));      // a batch takes time: chunks pile up meanwhile
  }
};

int main() {
  // 1. many producers, more streams than channels, more chunks than a batch holds
  {
    MockPipeline p(8, 12);
    b2k_host::DynamicBatcher<MockPipeline> b(&p, 1e-3);
    const int n_streams = 40, n_threads = 4;
    std::vector<int> n_chunks(n_streams);
    std::mt19937 rng(1);
    for (int s = 0; s < n_streams; s++) n_chunks[s] = 1 + (int)(rng() % 9);
    std::vector<std::thread> th;
    for (int t = 0; t < n_threads; t++)
      th.emplace_back([&, t] {
        for (int s = t; s < n_streams; s += n_threads)
          for (int k = 0; k < n_chunks[s]; k++) {
            std::vector<float> x(16 + (size_t)(s % 5), (float)k);     // the chunk's number in every sample
            b.Push((uint64_t)(1000 + s), k == 0, k == n_chunks[s] - 1, x.data(), (int64_t)x.size());
          }
      });
    for (auto &t : th) t.join();
    b.WaitForCompletion();
    REQUIRE(!p.saw_duplicate && !p.saw_bad_order && !p.saw_uninitialised && !p.saw_oversize);
    REQUIRE(p.live.empty());
    for (int s = 0; s < n_streams; s++) {
      REQUIRE(p.audio[(uint64_t)(1000 + s)].size() == (size_t)n_chunks[s] * (16 + (size_t)(s % 5)));
      REQUIRE(b.GetNumPendingChunks((uint64_t)(1000 + s)) == 0);
    }
    size_t total = 0;
    for (size_t n : p.batch_sizes) total += n;
    int sum = 0;
    for (int c : n_chunks) sum += c;
    REQUIRE(total == (size_t)sum);
  }
  // 2. the timeout: one lonely chunk is decoded without waiting for a full batch; a full batch does not wait for the timeout
  {
    MockPipeline p(4, 4);
    b2k_host::DynamicBatcher<MockPipeline> b(&p, 5e-3);
    float x[4] = {0, 0, 0, 0};
    b.Push(1, true, false, x, 4);
    b.WaitForCompletion();
    REQUIRE(p.batch_sizes.size() == 1 && p.batch_sizes[0] == 1 && b.GetNumPendingChunks(1) == 0);
    const auto t0 = std::chrono::steady_clock::now();
    {
      MockPipeline q(4, 8);
      b2k_host::DynamicBatcher<MockPipeline> slow(&q, 30.0);       // a timeout nobody waits for
      for (uint64_t id = 10; id < 14; id++) slow.Push(id, true, true, x, 4);
      slow.WaitForCompletion();
      REQUIRE(q.batch_sizes.size() == 1 && q.batch_sizes[0] == 4);
    }
    REQUIRE(std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() < 5.0);
  }
  // 3. an empty last chunk (a stream ended at an end point), an id reused for a second utterance, and pending counts
  {
    MockPipeline p(2, 2);
    b2k_host::DynamicBatcher<MockPipeline> b(&p, 1e-3);
    float a[2] = {0, 0}, c1[2] = {1, 1};
    b.Push(7, true, false, a, 2);
    b.Push(7, false, false, c1, 2);
    b.Push(7, false, true, nullptr, 0);
    b.WaitForCompletion();
    REQUIRE(p.audio[7].size() == 4 && p.live.empty());
    b.Push(7, true, true, a, 2);                                   // the same id again: a new utterance
    b.WaitForCompletion();
    REQUIRE(p.audio[7].size() == 6 && !p.saw_bad_order && !p.saw_uninitialised);
  }
  // 4. a pipeline that throws: the exception reaches the caller of WaitForCompletion, the batcher keeps working afterwards
  {
    MockPipeline p(2, 4);
    p.throw_on_batch = 0;
    b2k_host::DynamicBatcher<MockPipeline> b(&p, 1e-3);
    float a[1] = {0};
    b.Push(1, true, true, a, 1);
    bool thrown = false;
    try { b.WaitForCompletion(); } catch (const std::runtime_error &e) { thrown = std::string(e.what()) == "device fell over"; }
    REQUIRE(thrown);
    p.live.clear();
    b.Push(2, true, true, a, 1);
    b.WaitForCompletion();
    REQUIRE(p.audio[2].size() == 1);
  }
  std::printf("dynamic batcher ok\n");
  return 0;
}
