/* One utterance decoded CHUNK BY CHUNK through the C ABI only (C99, no Python, no Kaldi, no OpenFst): the loop of
 * online2-wav-nnet3-latgen-faster.cc:246-283 / BatchedThreadedNnet3CudaOnlinePipeline::DecodeBatch as a cgo / JNI / FFI host writes
 * it -- feed a chunk, read the partial hypothesis and the end-point rule, stop at an end point or at the end of the file, then
 * the lattice.  Everything up to the first device call runs on any machine; without a GPU the program stops there with the
 * library's own message and exit status 3.
 *
 *   stream_route <dir>/conf/online.conf <dir>/final.mdl <graph>/HCLG.fst utt.wav out.ark
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "b2k.h"

#define CHECK(call) do { if ((call) != B2K_OK) { fprintf(stderr, "%s: %s\n", #call, b2k_last_error()); return 1; } } while (0)

int main(int argc, char **argv) {
  if (argc < 6) { fprintf(stderr, "usage: stream_route online.conf final.mdl HCLG.fst utt.wav out.ark\n"); return 2; }
  b2k_online_conf oc;
  CHECK(b2k_online_conf_read(argv[1], &oc));
  b2k_stream_cfg cfg;
  b2k_stream_cfg_default(&cfg);
  CHECK(b2k_feat_cfg_from_conf(oc.feature_type == 0 ? oc.mfcc_config : oc.feature_type == 1 ? oc.fbank_config : oc.plp_config, oc.feature_type, &cfg.feat));
  cfg.feat.dither = 0.0f;                         /* the reference's dither is unseeded: results are defined only without it */
  cfg.feat.max_lanes = 1;
  cfg.nchannels = 1; cfg.max_seconds = 30.0f; cfg.frames_per_chunk = 51;
  b2k_endpoint_cfg ep;
  CHECK(b2k_endpoint_cfg_default(&ep));
  CHECK(b2k_endpoint_cfg_apply_options(oc.rest, &ep));
  b2k_model *model = NULL;
  CHECK(b2k_model_read(argv[2], 1, &model));
  int32_t mi[8];
  CHECK(b2k_model_info(model, mi));
  b2k_fst_file *graph = NULL;
  CHECK(b2k_fst_file_read(argv[3], &graph));
  b2k_wave *wav = NULL;
  CHECK(b2k_wave_read(argv[4], &wav));
  float rate; int32_t wav_channels; int64_t samples;
  CHECK(b2k_wave_info(wav, &rate, &wav_channels, &samples));
  const float *pcm = b2k_wave_data(wav);
  float *resampled = NULL;
  if (rate != cfg.feat.samp_freq) {
    int64_t n = 0;
    b2k_resample_waveform(rate, pcm, samples, cfg.feat.samp_freq, NULL, 0, &n);
    resampled = (float *)malloc(sizeof(float) * (size_t)(n > 0 ? n : 1));
    CHECK(b2k_resample_waveform(rate, pcm, samples, cfg.feat.samp_freq, resampled, n, &n));
    pcm = resampled; samples = n;
  }
  /* the stream takes 16-bit PCM, which is what a WAVE file holds (resampled audio is rounded back to it) */
  int16_t *pcm16 = (int16_t *)malloc(sizeof(int16_t) * (size_t)(samples > 0 ? samples : 1));
  for (int64_t i = 0; i < samples; i++) {
    float v = pcm[i] < 0.f ? pcm[i] - 0.5f : pcm[i] + 0.5f;
    pcm16[i] = (int16_t)(v > 32767.f ? 32767.f : v < -32768.f ? -32768.f : v);
  }
  const int32_t chunk = (int32_t)(cfg.frames_per_chunk * cfg.feat.samp_freq * 0.001f * cfg.feat.frame_shift_ms);
  printf("host side ready: %lld samples in chunks of %d, %d pdfs, %lld transition-ids, silence phones %s\n", (long long)samples, chunk, mi[2],
         (long long)mi[6] - 1, ep.silence_phones);
  /* device side */
  b2k_fst *fst = NULL;
  if (b2k_fst_create_from_file(graph, b2k_model_tid2pdf(model), mi[6], &fst) != B2K_OK) { fprintf(stderr, "%s\n", b2k_last_error()); return 3; }
  b2k_stream *s = NULL;
  CHECK(b2k_stream_create(&cfg, model, fst, &s));
  b2k_dec *dec = b2k_stream_decoder(s);
  const float frame_shift = 0.001f * cfg.feat.frame_shift_ms * (float)mi[3];             /* decoder frames */
  enum { CAP = 8192 };
  static int32_t ilabels[CAP], olabels[CAP];
  int32_t channel = 0, calls = 0, ended_early = 0;
  for (int64_t pos = 0; pos < samples || pos == 0; pos += chunk) {
    const int32_t n = (int32_t)(samples - pos < chunk ? samples - pos : chunk);
    const int16_t *piece = pcm16 + pos;
    int32_t first = pos == 0, last = pos + chunk >= samples, new_frames = 0, so_far = 0;
    CHECK(b2k_stream_decode_batch_i16(s, 1, &channel, &piece, &n, &first, &last, NULL, &new_frames, &so_far, NULL, NULL, NULL));
    calls++;
    if (last) break;
    b2k_best_path_info bp;                                                               /* the partial hypothesis */
    CHECK(b2k_dec_best_path(dec, &channel, 1, 0, CAP, ilabels, olabels, NULL, NULL, NULL, NULL, &bp, NULL));
    int32_t words = 0, hit = 0, sil = 0;
    for (int32_t k = 0; k < bp.n_arcs; k++) words += olabels[k] != 0;
    CHECK(b2k_endpoint_detected_on_path(&ep, b2k_model_tid2phone(model), mi[6], ilabels, bp.n_arcs, bp.num_frames, frame_shift,
                                        bp.final_relative_cost, &hit, &sil));
    printf("after %.2f s: %d frames decoded, %d words so far, %d trailing silence frames, endpoint %s\n",
           (double)(pos + n) / cfg.feat.samp_freq, so_far, words, sil, hit ? "yes" : "no");
    if (hit) {                                                                           /* end the utterance here: an empty last chunk */
      int32_t zero = 0, yes = 1, no = 0;
      const int16_t *none = NULL;
      CHECK(b2k_stream_decode_batch_i16(s, 1, &channel, &none, &zero, &no, &yes, NULL, &new_frames, &so_far, NULL, NULL, NULL));
      ended_early = 1;
      break;
    }
  }
  b2k_raw_lattice raw;
  memset(&raw, 0, sizeof(raw));
  CHECK(b2k_dec_get_raw_lattice(dec, channel, &raw, NULL));                              /* sizes */
  raw.state_frame = malloc(4 * raw.num_states + 4); raw.state_hclg = malloc(4 * raw.num_states + 4);
  raw.state_tot_cost = malloc(4 * raw.num_states + 4); raw.state_extra_cost = malloc(4 * raw.num_states + 4);
  raw.arc_src = malloc(4 * raw.num_arcs + 4); raw.arc_dst = malloc(4 * raw.num_arcs + 4); raw.arc_ilabel = malloc(4 * raw.num_arcs + 4);
  raw.arc_olabel = malloc(4 * raw.num_arcs + 4); raw.arc_graph_cost = malloc(4 * raw.num_arcs + 4); raw.arc_acoustic_cost = malloc(4 * raw.num_arcs + 4);
  raw.final_state = malloc(4 * raw.num_finals + 4); raw.final_cost = malloc(4 * raw.num_finals + 4);
  CHECK(b2k_dec_get_raw_lattice(dec, channel, &raw, NULL));
  b2k_clat *clat = NULL;
  CHECK(b2k_lat_determinize_pruned(&raw, cfg.dec.lattice_beam, 0, &clat));
  CHECK(b2k_clat_write(clat, "utt", argv[5], 1, 0));
  int32_t words[256], tids[8192], nw = 0, nt = 0;
  float g = 0.f, a = 0.f;
  CHECK(b2k_clat_best_path(clat, words, &nw, tids, &nt, 256, 8192, &g, &a));
  printf("streamed in %d calls%s: best path %d words, %d transition-ids, graph cost %g, acoustic cost %g; lattice written to %s\n", calls,
         ended_early ? " (ended at an end point)" : "", nw, nt, g, a, argv[5]);
  b2k_stream_destroy(s);
  return 0;
}
