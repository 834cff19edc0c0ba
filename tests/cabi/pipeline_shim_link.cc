// Type- and link-check of kaldi_b200/host/b2k_pipeline_shim.h against libb2k.so: builds the backend on the golden model
// and asks for a decode; without a device the first pipeline creation must throw with the library's "no CUDA device"
// message (there is no CPU path).  Compiled and run by tests/test_batcher_cpp.py.
#include <cstdio>
#include <cstring>

#include "b2k_pipeline_shim.h"

int main(int argc, char **argv) {
  if (argc < 2) return 2;
  b2k_model *m = nullptr;
  b2k_host::Check(b2k_model_read(argv[1], 1, &m), "b2k_model_read");
  b2k_pipeline_cfg cfg;
  b2k_pipeline_cfg_default(&cfg);
  cfg.max_batch = 2;
  const b2k_fst *fake_fst = reinterpret_cast<const b2k_fst *>(8);   // never dereferenced before the device check
  b2k_host::B2kPipelineBackend backend(cfg, m, fake_fst, nullptr, nullptr, 8.0f);
  b2k_host::B2kBatcher batcher(&backend, cfg.max_batch, 8);
  std::vector<float> wave(16000, 0.f);
  int rc = 1;
  try {
    batcher.AcceptChunks({1}, {{wave.data(), (int64_t)wave.size()}}, {true}, {true});
    batcher.Flush();
    std::printf("decoded\n");                                  // only on a box with a GPU (and then fake_fst would be a bug)
  } catch (const std::exception &e) {
    std::printf("threw: %s\n", e.what());
    rc = std::strstr(e.what(), "no CUDA device") ? 0 : 3;
  }
  b2k_model_destroy(m);
  return rc;
}
