/* From a Kaldi online-decoding directory to a lattice archive through the C ABI only (C99, no Python, no Kaldi, no OpenFst):
 * what a cgo / JNI / FFI host would do.  Everything up to the first device call runs on any machine; without a GPU the program
 * stops there with the library's own message and exit status 3.
 *
 *   experiment_route <dir>/conf/online.conf <dir>/final.mdl <graph>/HCLG.fst utt.wav out.ark
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "b2k.h"

#define CHECK(call) do { if ((call) != B2K_OK) { fprintf(stderr, "%s: %s\n", #call, b2k_last_error()); return 1; } } while (0)

int main(int argc, char **argv) {
  if (argc < 6) { fprintf(stderr, "usage: experiment_route online.conf final.mdl HCLG.fst utt.wav out.ark\n"); return 2; }
  /* option files */
  b2k_online_conf oc;
  CHECK(b2k_online_conf_read(argv[1], &oc));
  b2k_pipeline_cfg cfg;
  b2k_pipeline_cfg_default(&cfg);
  cfg.feat.max_lanes = 1;
  CHECK(b2k_feat_cfg_from_conf(oc.feature_type == 0 ? oc.mfcc_config : oc.feature_type == 1 ? oc.fbank_config : oc.plp_config, oc.feature_type, &cfg.feat));
  CHECK(b2k_pipeline_cfg_apply_options(oc.rest, &cfg));   /* --beam, --lattice-beam, --acoustic-scale ... if the file carries them */
  cfg.feat.dither = 0.0f;                         /* the reference's dither is unseeded: results are defined only without it */
  b2k_endpoint_cfg ep;                            /* the --endpoint.* group of the same file (online2-wav-nnet3-latgen-faster.cc:128) */
  CHECK(b2k_endpoint_cfg_default(&ep));
  CHECK(b2k_endpoint_cfg_apply_options(oc.rest, &ep));
  /* model, graph, waveform */
  b2k_model *model = NULL;
  CHECK(b2k_model_read(argv[2], 1, &model));
  int32_t mi[8];
  CHECK(b2k_model_info(model, mi));
  b2k_fst_file *graph = NULL;
  CHECK(b2k_fst_file_read(argv[3], &graph));
  b2k_wave *wav = NULL;
  CHECK(b2k_wave_read(argv[4], &wav));
  float rate; int32_t channels; int64_t samples;
  CHECK(b2k_wave_info(wav, &rate, &channels, &samples));
  const float *pcm = b2k_wave_data(wav);          /* channel 0 */
  float *resampled = NULL;
  if (rate != cfg.feat.samp_freq) {               /* --allow-downsample / --allow-upsample */
    int64_t n = 0;
    b2k_resample_waveform(rate, pcm, samples, cfg.feat.samp_freq, NULL, 0, &n);
    resampled = (float *)malloc(sizeof(float) * (size_t)(n > 0 ? n : 1));
    CHECK(b2k_resample_waveform(rate, pcm, samples, cfg.feat.samp_freq, resampled, n, &n));
    pcm = resampled; samples = n;
  }
  /* i-vector extractor, if the model takes one */
  b2k_ivec_files *ivf = NULL;
  b2k_ivec_cfg icfg;
  memset(&icfg, 0, sizeof(icfg));
  cfg.max_batch = 1;
  cfg.num_samples = samples;
  b2k_pipeline_plan plan;
  CHECK(b2k_pipeline_plan_for(&cfg, model, &plan));
  if (mi[1] > 0) {
    b2k_ivec_paths ip;
    icfg.base_dim = plan.feat_dim; icfg.max_lanes = 1; icfg.max_frames = plan.num_feature_frames;
    CHECK(b2k_ivec_cfg_from_conf(oc.ivector_extraction_config, &icfg, &ip));
    CHECK(b2k_ivec_files_read(ip.ivector_extractor, ip.diag_ubm, ip.lda_matrix, ip.global_cmvn_stats, &ivf));
    cfg.ivector_splice_right = icfg.splice_right;
  }
  {
    /* what the endpoint rules say about 2 s of speech followed by 0.6 s of silence, had the decoder produced that path */
    int32_t hit = 0;
    CHECK(b2k_endpoint_detected(&ep, 87, 20, 0.03f, 0.0f, &hit));
    printf("endpointing: silence phones %s, rule2 fires after 0.6 s of trailing silence: %s\n", ep.silence_phones, hit ? "yes" : "no");
  }
  printf("host side ready: %d feature frames -> %d decoder frames, %d pdfs, %lld transition-ids\n", plan.num_feature_frames,
         plan.num_output_frames, plan.num_pdfs, (long long)mi[6] - 1);
  /* device side */
  b2k_fst *fst = NULL;
  if (b2k_fst_create_from_file(graph, b2k_model_tid2pdf(model), mi[6], &fst) != B2K_OK) { fprintf(stderr, "%s\n", b2k_last_error()); return 3; }
  b2k_ivec *ivec = NULL;
  if (ivf) CHECK(b2k_ivec_create_from_files(&icfg, ivf, &ivec));
  b2k_pipeline *pipe = NULL;
  CHECK(b2k_pipeline_create(&cfg, model, fst, ivec, &pipe));
  const float *waves[1] = {pcm};
  CHECK(b2k_pipeline_decode_batch(pipe, 1, waves, NULL));
  b2k_raw_lattice raw;
  memset(&raw, 0, sizeof(raw));
  CHECK(b2k_dec_get_raw_lattice(b2k_pipeline_decoder(pipe), 0, &raw, NULL));           /* sizes */
  raw.state_frame = malloc(4 * raw.num_states); raw.state_hclg = malloc(4 * raw.num_states);
  raw.state_tot_cost = malloc(4 * raw.num_states); raw.state_extra_cost = malloc(4 * raw.num_states);
  raw.arc_src = malloc(4 * raw.num_arcs); raw.arc_dst = malloc(4 * raw.num_arcs); raw.arc_ilabel = malloc(4 * raw.num_arcs);
  raw.arc_olabel = malloc(4 * raw.num_arcs); raw.arc_graph_cost = malloc(4 * raw.num_arcs); raw.arc_acoustic_cost = malloc(4 * raw.num_arcs);
  raw.final_state = malloc(4 * raw.num_finals + 4); raw.final_cost = malloc(4 * raw.num_finals + 4);
  CHECK(b2k_dec_get_raw_lattice(b2k_pipeline_decoder(pipe), 0, &raw, NULL));
  b2k_clat *clat = NULL;
  CHECK(b2k_lat_determinize_pruned(&raw, cfg.dec.lattice_beam, 0, &clat));              /* GetLattice */
  CHECK(b2k_clat_write(clat, "utt", argv[5], 1, 0));                                    /* ark:out.ark */
  int32_t words[256], tids[4096], nw = 0, nt = 0;
  float g = 0.f, a = 0.f;
  CHECK(b2k_clat_best_path(clat, words, &nw, tids, &nt, 256, 4096, &g, &a));             /* CompactLatticeShortestPath */
  printf("best path: %d words, %d transition-ids, graph cost %g, acoustic cost %g\n", nw, nt, g, a);
  int64_t sz[6];
  b2k_clat_sizes(clat, sz);
  printf("lattice: %lld raw arcs -> %lld compact arcs, written to %s\n", (long long)raw.num_arcs, (long long)sz[1], argv[5]);
  return 0;
}
