// TEST INFRASTRUCTURE ONLY (oracle side).  Compiles the reference's OWN online2/online-endpoint.cc where it lies:
// EndpointDetected (the five rules), TrailingSilenceLength<DEC> and OnlineEndpointConfig::Register run unmodified.
// The decoder it queries is the replay device of replay_decoder.h (the decoder headers need OpenFst, which this image
// does not have): the endpoint code itself sees the interface it was written for.
#include <cstdio>
#include <cstring>
#include <sstream>
#include <string>
#include <vector>

#include "replay_decoder.h"
#include "util/common-utils.h"
#include "hmm/transition-model.h"

#include "online2/online-endpoint.cc"

using namespace kaldi;

namespace {
typedef LatticeFasterOnlineDecoderTpl<fst::Fst<fst::StdArc> > Dec;

bool config_from_text(const char *conf_text, OnlineEndpointConfig *cfg) {
  // the registration path of the tools: ParseOptions + Register + ReadConfigFile (util/parse-options.cc:479)
  ParseOptions po("");
  cfg->Register(&po);
  char path[] = "/tmp/b2k_endpoint_conf_XXXXXX";
  int fd = mkstemp(path);
  if (fd < 0) return false;
  FILE *f = fdopen(fd, "w");
  fputs(conf_text, f);
  fclose(f);
  try {
    po.ReadConfigFile(path);
  } catch (...) { remove(path); throw; }
  remove(path);
  return true;
}
}  // namespace

extern "C" {

// out[20] = per rule {must_contain_nonsilence, min_trailing_silence, max_relative_cost, min_utterance_length};
// silence_phones copied to sil (NUL-terminated).  Returns 0, -1 on a rejected option file.
int ref_endpoint_config(const char *conf_text, float *out, char *sil, int sil_cap) {
  try {
    OnlineEndpointConfig cfg;
    if (!config_from_text(conf_text, &cfg)) return -1;
    const OnlineEndpointRule *r[5] = {&cfg.rule1, &cfg.rule2, &cfg.rule3, &cfg.rule4, &cfg.rule5};
    for (int i = 0; i < 5; i++) {
      out[4 * i + 0] = r[i]->must_contain_nonsilence ? 1.0f : 0.0f;
      out[4 * i + 1] = r[i]->min_trailing_silence;
      out[4 * i + 2] = r[i]->max_relative_cost;
      out[4 * i + 3] = r[i]->min_utterance_length;
    }
    snprintf(sil, sil_cap, "%s", cfg.silence_phones.c_str());
    return 0;
  } catch (const std::exception &e) { return -1; }
}

// EndpointDetected(config, num_frames_decoded, trailing_silence_frames, frame_shift, final_relative_cost)
// (online-endpoint.cc:47): 1 / 0, -1 on error.
int ref_endpoint_detected(const char *conf_text, int num_frames_decoded, int trailing_silence_frames,
                          float frame_shift, float final_relative_cost) {
  try {
    OnlineEndpointConfig cfg;
    if (!config_from_text(conf_text, &cfg)) return -1;
    return EndpointDetected(cfg, num_frames_decoded, trailing_silence_frames, frame_shift, final_relative_cost) ? 1 : 0;
  } catch (const std::exception &e) { return -1; }
}

// The decoder form (online-endpoint.cc:118) on a replayed best path, with the transition model read from `mdl_path`
// by the reference's own TransitionModel::Read.  trailing_out = TrailingSilenceLength (online-endpoint.cc:78).
int ref_endpoint_detected_on_path(const char *conf_text, const char *mdl_path, const int *ilabels, int n, int frames,
                                  float frame_shift, float final_relative_cost, int *trailing_out) {
  try {
    OnlineEndpointConfig cfg;
    if (!config_from_text(conf_text, &cfg)) return -1;
    TransitionModel tm;
    { bool binary; Input ki(mdl_path, &binary); tm.Read(ki.Stream(), binary); }
    Dec dec;
    dec.path.assign(ilabels, ilabels + n);
    dec.frames = frames;
    dec.final_relative_cost = final_relative_cost;
    if (trailing_out) *trailing_out = frames == 0 ? 0 : TrailingSilenceLength(tm, cfg.silence_phones, dec);
    return EndpointDetected(cfg, tm, frame_shift, dec) ? 1 : 0;
  } catch (const std::exception &e) { fprintf(stderr, "ref_endpoint_detected_on_path: %s\n", e.what()); return -1; }
}

// TransitionIdToPhone for t = 1..NumTransitionIds (out[0] = 0); returns NumTransitionIds or -1 / -2.
int ref_tid2phone(const char *mdl_path, int *out, int max_tids) {
  try {
    TransitionModel tm;
    { bool binary; Input ki(mdl_path, &binary); tm.Read(ki.Stream(), binary); }
    const int n = tm.NumTransitionIds();
    if (n + 1 > max_tids) return -2;
    out[0] = 0;
    for (int t = 1; t <= n; t++) out[t] = tm.TransitionIdToPhone(t);
    return n;
  } catch (const std::exception &e) { return -1; }
}

}  // extern "C"
