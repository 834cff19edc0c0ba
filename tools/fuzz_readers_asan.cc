// tools/fuzz_readers_asan.cc -- NOT part of the product build.  AddressSanitizer / UBSan harness for the host-only readers and
// the nnet3 program compiler: compile it together with the host-only translation units (they are plain C++ apart from their
// .cu suffix) and feed it corrupted files; tests/test_reader_fuzz.py does the same through libb2k.so without the sanitizers.
//
//   g++ -g -O1 -fsanitize=address,undefined -fno-sanitize-recover=undefined -std=c++17 -x c++ -Iinclude -Ikaldi_b200/csrc \
//       -I/usr/local/cuda/include tools/fuzz_readers_asan.cc kaldi_b200/csrc/{model_io,common,nnet_compile,fst_io,host_utils}.cu \
//       -o /tmp/fuzz -L/usr/local/cuda/lib64 -lcudart
//   ASAN_OPTIONS=detect_leaks=0 /tmp/fuzz mdl variants/*        (kinds: mdl raw fst wav conf ie dubm)
//
// Round 1 result: 600-700 truncated / bit-flipped / extreme-count variants per file kind (binary and text model files with the
// recipe extras, const / aligned / vector graphs, WAVE, option files, final.ie / final.dubm): clean after fixing, in model_io.cu,
// counts that were trusted before being checked against the bytes left, a phone index outside the topology table and an
// end-of-file loop in text mode, and, in nnet_compile.cu, weights whose sizes were not checked against the layer dimensions.
// The input-descriptor parser (splices of the chain TDNN layers) was run the same way on 800 variants of a reference-written chain
// TDNN model whose config section was mutated (flipped bytes, deleted spans, inserted "Offset(" / "Append(" fragments, huge offsets): clean.
#include "b2k.h"
#include <cstdio>
#include <cstring>
extern "C" int b2k_ivec_create(const b2k_ivec_cfg*, const float*, const float*, const float*, const float*, const double*, const double*, const double*, b2k_ivec**) { return 1; }
extern "C" int b2k_nnet_create(const b2k_nnet_node*, int32_t, const b2k_nnet_op*, int32_t, const float*, int64_t, int32_t, b2k_nnet**) { return 1; }
extern "C" int b2k_fst_create(const b2k_fst_csr*, b2k_fst**) { return 1; }
// usage: h2 kind files...   kind: mdl raw fst wav conf ie dubm
int main(int argc, char** argv){
  const char *k = argv[1];
  int ok = 0;
  for (int i = 2; i < argc; i++) {
    int rc = 1;
    if (!strcmp(k, "mdl") || !strcmp(k, "raw")) { b2k_model* m=nullptr; rc=b2k_model_read(argv[i], !strcmp(k,"mdl"), &m);
      if(!rc) { // also push it through the compiler, as a caller would
        int32_t mi[8]; b2k_model_info(m, mi); b2k_nnet_compile_cfg cc; memset(&cc,0,sizeof(cc)); cc.feat_dim=mi[0]; cc.ivector_dim=mi[1]; cc.num_pdfs=mi[2];
        cc.frame_subsampling_factor=mi[3]; cc.num_frames=60; cc.frames_per_chunk=21*(mi[3]>0?1:1); cc.use_priors=1; cc.acoustic_scale=1.f;
        if (mi[3] > 0 && cc.frames_per_chunk % mi[3] == 0) { b2k_nnet_program *p=nullptr; if (!b2k_nnet_compile(&cc, b2k_model_layers(m), mi[4], b2k_model_weights(m), mi[5], &p)) b2k_nnet_program_destroy(p); }
        b2k_model_destroy(m); } }
    else if (!strcmp(k, "fst")) { b2k_fst_file* f=nullptr; rc=b2k_fst_file_read(argv[i], &f); if(!rc) b2k_fst_file_destroy(f); }
    else if (!strcmp(k, "wav")) { b2k_wave* w=nullptr; rc=b2k_wave_read(argv[i], &w); if(!rc) b2k_wave_destroy(w); }
    else if (!strcmp(k, "conf")) { b2k_feat_cfg c; c.max_lanes = 0; rc=b2k_feat_cfg_from_conf(argv[i], 0, &c); }
    else if (!strcmp(k, "ie")) { b2k_ivec_files* f=nullptr; rc=b2k_ivec_files_read(argv[i], argv[argc-3], argv[argc-2], argv[argc-1], &f); if(!rc) b2k_ivec_files_destroy(f); if (i >= argc-4) break; }
    else if (!strcmp(k, "dubm")) { b2k_ivec_files* f=nullptr; rc=b2k_ivec_files_read(argv[argc-3], argv[i], argv[argc-2], argv[argc-1], &f); if(!rc) b2k_ivec_files_destroy(f); if (i >= argc-4) break; }
    ok += rc == 0;
  }
  printf("done %d accepted %d\n", argc - 2, ok); return 0; }
