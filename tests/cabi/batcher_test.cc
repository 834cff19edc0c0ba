// CPU unit test of kaldi_b200/host/b2k_batcher.h with a mock backend (compiled and run by tests/test_batcher_cpp.py).
#include <cstdio>
#include <cstdlib>
#include <numeric>

#include "b2k_batcher.h"

#define CHECK(c) do { if (!(c)) { std::fprintf(stderr, "CHECK failed at line %d: %s\n", __LINE__, #c); std::exit(1); } } while (0)

struct MockBackend {
  struct Result { int64_t len; double sum; int batch_index, batch_size; };
  std::vector<std::pair<int64_t, int>> calls;                 // (length, batch size)
  void Decode(int64_t n, const std::vector<const float *> &waves, std::vector<Result> *out) {
    calls.push_back({n, (int)waves.size()});
    for (size_t i = 0; i < waves.size(); i++)
      out->push_back(Result{n, std::accumulate(waves[i], waves[i] + n, 0.0), (int)i, (int)waves.size()});
  }
};

typedef b2k_host::UtteranceBatcher<MockBackend> Batcher;

template <class F> bool throws(F f) { try { f(); } catch (const std::exception &) { return true; } return false; }

int main() {
  MockBackend be;
  Batcher b(&be, /*max_batch=*/2, /*max_open=*/4);
  std::vector<std::pair<uint64_t, MockBackend::Result>> got;
  b.SetDefaultCallback([&](uint64_t id, MockBackend::Result &r) { got.push_back({id, r}); });

  std::vector<float> a(10, 1.f), c(10, 2.f), d(7, 3.f), e(10, 4.f);
  // utterance 1 arrives in three chunks (4 + 4 + 2 samples), utterance 2 in one; both have 10 samples
  CHECK(b.TryInitCorrId(1));
  CHECK(!b.TryInitCorrId(1));
  int special = 0;
  b.SetCallback(1, [&](uint64_t id, MockBackend::Result &r) { special++; CHECK(id == 1 && r.sum == 10.0 && r.len == 10); });
  CHECK(b.AcceptChunks({1}, {{a.data(), 4}}, {true}, {false}) == 0);
  CHECK(b.AcceptChunks({1, 2}, {{a.data() + 4, 4}, {c.data(), 10}}, {false, true}, {false, true}) == 0);
  CHECK(b.NumOpen() == 1 && b.NumWaiting() == 1 && be.calls.empty());
  // the last chunk of 1 completes a batch of two 10-sample utterances: decoded at once, in arrival order (2 closed first)
  CHECK(b.AcceptChunks({1}, {{a.data() + 8, 2}}, {false}, {true}) == 2);
  CHECK(be.calls.size() == 1 && be.calls[0].first == 10 && be.calls[0].second == 2);
  CHECK(special == 1 && got.size() == 1 && got[0].first == 2 && got[0].second.sum == 20.0 && got[0].second.batch_index == 0);
  CHECK(b.NumOpen() == 0 && b.NumWaiting() == 0);

  // different lengths never share a batch; Flush decodes partial batches, shortest bucket first
  CHECK(b.AcceptChunks({3, 4}, {{d.data(), 7}, {e.data(), 10}}, {true, true}, {true, true}) == 0);
  CHECK(b.NumWaiting() == 2);
  CHECK(b.Flush() == 2);
  CHECK(be.calls.size() == 3 && be.calls[1] == std::make_pair((int64_t)7, 1) && be.calls[2] == std::make_pair((int64_t)10, 1));
  CHECK(got.size() == 3 && got[1].first == 3 && got[1].second.sum == 21.0 && got[2].first == 4 && got[2].second.sum == 40.0);
  CHECK(b.Flush() == 0);

  // misuse: nothing changes state
  CHECK(throws([&] { b.AcceptChunks({9}, {{a.data(), 4}}, {false}, {false}); }));             // not started
  CHECK(throws([&] { b.AcceptChunks({9, 9}, {{a.data(), 1}, {a.data(), 1}}, {true, true}, {false, false}); }));   // twice in one batch
  CHECK(throws([&] { b.AcceptChunks({9}, {{a.data(), 4}}, {true, false}, {false}); }));       // size mismatch
  CHECK(throws([&] { b.SetCallback(77, Batcher::Callback()); }));
  CHECK(b.NumOpen() == 0);
  CHECK(b.AcceptChunks({9}, {{a.data(), 4}}, {true}, {false}) == 0);
  CHECK(throws([&] { b.AcceptChunks({9}, {{a.data(), 4}}, {true}, {false}); }));              // first chunk twice
  // slot limit: 4 open utterances
  CHECK(b.AcceptChunks({10, 11, 12}, {{a.data(), 1}, {a.data(), 1}, {a.data(), 1}}, {true, true, true}, {false, false, false}) == 0);
  CHECK(b.NumOpen() == 4 && !b.TryInitCorrId(13));
  CHECK(throws([&] { b.AcceptChunks({13}, {{a.data(), 1}}, {true}, {false}); }));
  // an empty last chunk closes the utterance with what it has
  CHECK(b.AcceptChunks({9}, {{nullptr, 0}}, {false}, {true}) == 0);
  CHECK(b.NumOpen() == 3 && b.NumWaiting() == 1);
  CHECK(b.Flush() == 1 && got.back().first == 9 && got.back().second.len == 4);
  std::printf("batcher ok: %zu backend calls, %zu callbacks\n", be.calls.size(), got.size() + special);
  return 0;
}
