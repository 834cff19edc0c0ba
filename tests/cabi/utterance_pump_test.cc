// b2k_host::UtterancePump with a mock pipeline: every sample of every utterance is delivered exactly once and in order, first /
// last flags are right, a batch holds at most max_batch chunks and never two of one utterance, a full batch is used whenever
// enough utterances are unfinished, Run(false) leaves fewer than max_batch unfinished, done callbacks fire once, right after
// the batch with the last chunk.
#include <cstdio>
#include <cstdlib>
#include <map>
#include <random>
#include <set>
#include <vector>

#include "b2k_utterance_pump.h"

#define REQUIRE(c) do { if (!(c)) { std::fprintf(stderr, "FAILED %s:%d: %s\n", __FILE__, __LINE__, #c); std::exit(1); } } while (0)

struct Mock {
  int max_batch;
  int64_t chunk;
  std::map<uint64_t, std::vector<float> > got;
  std::map<uint64_t, int64_t> expect_len;
  std::set<uint64_t> started, finished;
  std::vector<size_t> batch_sizes;
  size_t unfinished_before = 0;          // set by the test before Run: utterances not finished yet
  void DecodeBatch(const std::vector<uint64_t> &ids, const std::vector<std::pair<const float *, int64_t> > &chunks,
                   const std::vector<bool> &first, const std::vector<bool> &last) {
    REQUIRE(ids.size() == chunks.size() && ids.size() == first.size() && ids.size() == last.size());
    REQUIRE(!ids.empty() && ids.size() <= (size_t)max_batch);
    std::set<uint64_t> in_batch(ids.begin(), ids.end());
    REQUIRE(in_batch.size() == ids.size());
    // a batch is as full as the unfinished utterances allow
    REQUIRE(ids.size() == std::min<size_t>(max_batch, expect_len.size() - finished.size()));
    for (size_t i = 0; i < ids.size(); i++) {
      REQUIRE(!finished.count(ids[i]));
      REQUIRE(first[i] == !started.count(ids[i]));
      started.insert(ids[i]);
      std::vector<float> &g = got[ids[i]];
      REQUIRE(chunks[i].second >= 1 && chunks[i].second <= chunk);
      g.insert(g.end(), chunks[i].first, chunks[i].first + chunks[i].second);
      const bool is_last = (int64_t)g.size() == expect_len[ids[i]];
      REQUIRE(last[i] == is_last);
      if (!is_last) REQUIRE(chunks[i].second == chunk);        // only the last chunk may be short
      if (is_last) finished.insert(ids[i]);
    }
    batch_sizes.push_back(ids.size());
  }
};

int main() {
  std::mt19937 rng(7);
  for (int trial = 0; trial < 40; trial++) {
    const int max_batch = 1 + (int)(rng() % 9);
    const int64_t chunk = 1 + (int64_t)(rng() % 50);
    const int n_utt = 1 + (int)(rng() % 60);
    Mock mock;
    mock.max_batch = max_batch; mock.chunk = chunk;
    b2k_host::UtterancePump<Mock> pump(&mock, max_batch, chunk);
    std::vector<std::vector<float> > audio(n_utt);
    std::map<uint64_t, int> done_count;
    // all lengths are known to the mock up front, the pump gets them one by one (with eager runs in between)
    for (int u = 0; u < n_utt; u++) {
      const int64_t len = 1 + (int64_t)(rng() % (6 * chunk));
      audio[u].resize(len);
      for (int64_t k = 0; k < len; k++) audio[u][k] = (float)(u * 100000 + k);
    }
    bool eager = trial % 2 == 0;
    if (!eager) for (int u = 0; u < n_utt; u++) mock.expect_len[1000 + u] = (int64_t)audio[u].size();
    for (int u = 0; u < n_utt; u++) {
      const uint64_t id = 1000 + u;
      if (eager) mock.expect_len[id] = (int64_t)audio[u].size();
      pump.Add(id, audio[u].data(), (int64_t)audio[u].size(), [&done_count, &mock, id]() {
        REQUIRE(mock.finished.count(id));          // after the batch with the last chunk
        done_count[id]++;
      });
      if (eager) {
        pump.Run(false);
        REQUIRE(pump.NumUnfinished() < (size_t)max_batch);
      }
    }
    pump.Run(true);
    REQUIRE(pump.NumUnfinished() == 0);
    for (int u = 0; u < n_utt; u++) {
      const uint64_t id = 1000 + u;
      REQUIRE(mock.got[id] == audio[u]);
      REQUIRE(done_count[id] == 1);
    }
  }
  // refused input
  {
    Mock mock; mock.max_batch = 2; mock.chunk = 4;
    b2k_host::UtterancePump<Mock> pump(&mock, 2, 4);
    bool threw = false;
    float x = 0;
    try { pump.Add(1, &x, 0); } catch (const std::invalid_argument &) { threw = true; }
    REQUIRE(threw);
    threw = false;
    try { b2k_host::UtterancePump<Mock> bad(&mock, 0, 4); } catch (const std::invalid_argument &) { threw = true; }
    REQUIRE(threw);
  }
  std::printf("utterance pump ok\n");
  return 0;
}
