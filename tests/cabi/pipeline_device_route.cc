// The whole route in C++ on the device, through the host shims only (no Python between the files and the lattices):
//   final.mdl -> b2k_model_read;  HCLG.fst -> b2k_fst_file_read -> b2k_fst_create_from_file;
//   B2kPipelineBackend + UtteranceBatcher: chunks in (as BatchedThreadedNnet3CudaOnlinePipeline::DecodeBatch receives
//   them), compact lattices out through the per-utterance callback.
// usage: pipeline_device_route final.mdl HCLG.fst waves.f32 num_utts num_samples out.bin tid2pdf.i32
// (tid2pdf.i32: the transition-id -> pdf table of the synthetic graph; with a real model it is b2k_model_tid2pdf)
// out.bin, per utterance in callback order: int64 {corr_id, states, arcs, finals, tids} then the b2k_compact_lattice
// arrays (arc_src, arc_dst, arc_word [int32], arc_graph, arc_acoustic [float], arc_tids_off [int64], final_state
// [int32], final_graph, final_acoustic [float], final_tids_off [int64], tids [int32]).
// Built and run by tests/test_zz_model_route.py on a GPU box.
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "b2k_pipeline_shim.h"

int main(int argc, char **argv) {
  if (argc < 8) { std::fprintf(stderr, "usage\n"); return 2; }
  const int n = std::atoi(argv[4]);
  const int64_t S = std::atoll(argv[5]);
  try {
    b2k_model *m = nullptr;
    b2k_host::Check(b2k_model_read(argv[1], 1, &m), "b2k_model_read");
    b2k_fst_file *ff = nullptr;
    b2k_host::Check(b2k_fst_file_read(argv[2], &ff), "b2k_fst_file_read");
    b2k_fst *fst = nullptr;
    std::vector<int32_t> t2p;
    {
      FILE *tf = std::fopen(argv[7], "rb");
      if (!tf) { std::fprintf(stderr, "cannot read tid2pdf\n"); return 2; }
      int32_t v;
      while (std::fread(&v, 4, 1, tf) == 1) t2p.push_back(v);
      std::fclose(tf);
    }
    b2k_host::Check(b2k_fst_create_from_file(ff, t2p.data(), (int32_t)t2p.size(), &fst), "b2k_fst_create_from_file");
    b2k_fst_file_destroy(ff);

    std::vector<float> waves((size_t)n * S);
    FILE *f = std::fopen(argv[3], "rb");
    if (!f || std::fread(waves.data(), 4, waves.size(), f) != waves.size()) { std::fprintf(stderr, "cannot read waves\n"); return 2; }
    std::fclose(f);

    b2k_pipeline_cfg cfg;
    b2k_pipeline_cfg_default(&cfg);
    cfg.max_batch = 2;                         // n = 3 utterances of one length: one full batch + one flushed
    b2k_host::B2kPipelineBackend backend(cfg, m, fst, nullptr, nullptr, /*det_beam=*/cfg.dec.lattice_beam);
    b2k_host::B2kBatcher batcher(&backend, cfg.max_batch, 16);
    FILE *out = std::fopen(argv[6], "wb");
    if (!out) return 2;
    batcher.SetDefaultCallback([&](uint64_t id, b2k_host::B2kPipelineBackend::Result &r) {
      int64_t sz[6];
      b2k_host::Check(b2k_clat_sizes(r.clat.get(), sz), "b2k_clat_sizes");
      std::vector<int32_t> as(sz[1]), ad(sz[1]), aw(sz[1]), fs(sz[2]), tids(sz[3]);
      std::vector<float> ag(sz[1]), aa(sz[1]), fg(sz[2]), fa(sz[2]);
      std::vector<int64_t> ao(sz[1] + 1), fo(sz[2] + 1);
      b2k_compact_lattice v = {};
      v.arc_src = as.data(); v.arc_dst = ad.data(); v.arc_word = aw.data(); v.arc_graph_cost = ag.data(); v.arc_acoustic_cost = aa.data();
      v.arc_tids_off = ao.data(); v.final_state = fs.data(); v.final_graph_cost = fg.data(); v.final_acoustic_cost = fa.data();
      v.final_tids_off = fo.data(); v.tids = tids.data();
      b2k_host::Check(b2k_clat_copy(r.clat.get(), &v), "b2k_clat_copy");
      const int64_t head[5] = {(int64_t)id, sz[0], sz[1], sz[2], sz[3]};
      std::fwrite(head, 8, 5, out);
      std::fwrite(as.data(), 4, as.size(), out); std::fwrite(ad.data(), 4, ad.size(), out); std::fwrite(aw.data(), 4, aw.size(), out);
      std::fwrite(ag.data(), 4, ag.size(), out); std::fwrite(aa.data(), 4, aa.size(), out); std::fwrite(ao.data(), 8, ao.size(), out);
      std::fwrite(fs.data(), 4, fs.size(), out); std::fwrite(fg.data(), 4, fg.size(), out); std::fwrite(fa.data(), 4, fa.size(), out);
      std::fwrite(fo.data(), 8, fo.size(), out); std::fwrite(tids.data(), 4, tids.size(), out);
      std::printf("utt %llu raw %lld/%lld clat %lld/%lld\n", (unsigned long long)id, (long long)r.raw_states, (long long)r.raw_arcs,
                  (long long)sz[0], (long long)sz[1]);
    });
    // feed each utterance in three chunks, interleaved across utterances like a streaming front end would
    const int64_t cuts[4] = {0, S / 3, 2 * S / 3, S};
    for (int c = 0; c < 3; c++) {
      std::vector<uint64_t> ids;
      std::vector<std::pair<const float *, int64_t>> chunks;
      std::vector<bool> first, last;
      for (int i = 0; i < n; i++) {
        ids.push_back(100 + i);
        chunks.push_back({waves.data() + (size_t)i * S + cuts[c], cuts[c + 1] - cuts[c]});
        first.push_back(c == 0); last.push_back(c == 2);
      }
      batcher.AcceptChunks(ids, chunks, first, last);
    }
    batcher.Flush();
    std::fclose(out);
    std::printf("pipelines kept: %zu\n", backend.NumPipelines());
    b2k_fst_destroy(fst);
    b2k_model_destroy(m);
  } catch (const std::exception &e) {
    std::fprintf(stderr, "error: %s\n", e.what());
    return 1;
  }
  return 0;
}
