/* A plain C99 caller of libb2k.so (what a cgo / JNI / FFI binding would do): the device-free part of the route
 * final.mdl -> b2k_model_read -> b2k_pipeline_plan_for -> b2k_nnet_compile, then the first device call, which must
 * fail with B2K_ERR_NO_DEVICE on a box without a GPU and succeed on one.  Prints one line the test parses. */
#include <stdio.h>
#include <string.h>

#include "b2k.h"

int main(int argc, char **argv) {
  if (argc < 2) { fprintf(stderr, "usage: host_route final.mdl\n"); return 2; }
  b2k_model *m = NULL;
  if (b2k_model_read(argv[1], 1, &m) != B2K_OK) { fprintf(stderr, "%s\n", b2k_last_error()); return 1; }
  int32_t mi[8];
  b2k_model_info(m, mi);

  b2k_pipeline_cfg cfg;
  b2k_pipeline_cfg_default(&cfg);
  cfg.max_batch = 4;
  cfg.num_samples = 32000;
  b2k_pipeline_plan plan;
  if (b2k_pipeline_plan_for(&cfg, m, &plan) != B2K_OK) { fprintf(stderr, "%s\n", b2k_last_error()); return 1; }

  b2k_nnet_compile_cfg cc;
  memset(&cc, 0, sizeof(cc));
  cc.feat_dim = mi[0]; cc.ivector_dim = mi[1]; cc.num_pdfs = mi[2]; cc.frame_subsampling_factor = mi[3];
  cc.num_frames = plan.num_feature_frames; cc.frames_per_chunk = cfg.frames_per_chunk; cc.use_priors = 1;
  cc.acoustic_scale = 1.0f;
  b2k_nnet_program *prog = NULL;
  if (b2k_nnet_compile(&cc, b2k_model_layers(m), mi[4], b2k_model_weights(m), mi[5], &prog) != B2K_OK) {
    fprintf(stderr, "%s\n", b2k_last_error());
    return 1;
  }
  int32_t n_nodes, n_ops;
  int64_t blob, pi[8];
  b2k_nnet_program_sizes(prog, &n_nodes, &n_ops, &blob);
  b2k_nnet_program_info(prog, pi);

  b2k_nnet *nn = NULL;
  int rc = b2k_nnet_create_from_program(prog, cfg.max_batch, &nn);
  printf("pdfs=%d frames=%d out=%d chunks=%d nodes=%d ops=%d blob=%lld prog_out=%lld prog_chunks=%lld tids=%d device_rc=%d\n",
         mi[2], plan.num_feature_frames, plan.num_output_frames, plan.num_chunks, n_nodes, n_ops, (long long)blob,
         (long long)pi[0], (long long)pi[1], mi[6], rc);
  if (nn) b2k_nnet_destroy(nn);
  b2k_nnet_program_destroy(prog);
  b2k_model_destroy(m);
  return 0;
}
