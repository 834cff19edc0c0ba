// The option mappings of the Kaldi-typed shims, RUN: a config file goes through the reference's own ParseOptions and
// MfccOptions / FbankOptions / PlpOptions / OnlineEndpointConfig (compiled in oracle/_ref) and then through the shim's mapping
// (ToB2kFeatCfg, ToB2kEndpointConfig: what a Kaldi tool built against b2k uses), and -- independently -- through b2k's own
// readers of the same file (b2k_feat_cfg_from_conf, b2k_endpoint_cfg_from_conf: what a C caller uses).  The two structs must
// be identical, field for field.  Test harness only; needs no device (both paths are host code).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

#include "replay_decoder.h"                        // declares the decoder templates the online2 headers name
#include "b2k_online2_shims.h"
#include "online2/online-endpoint.h"
#include "util/parse-options.h"

using namespace kaldi;

#define REQUIRE(c) do { if (!(c)) { std::fprintf(stderr, "FAILED %s:%d: %s\n", __FILE__, __LINE__, #c); std::exit(1); } } while (0)

static void same_feat(const b2k_feat_cfg &a, const b2k_feat_cfg &b, const char *what) {
#define F(x) if (a.x != b.x) { std::fprintf(stderr, "%s: field %s differs: %g vs %g\n", what, #x, (double)a.x, (double)b.x); std::exit(1); }
  F(feature_type) F(samp_freq) F(frame_shift_ms) F(frame_length_ms) F(dither) F(preemph_coeff) F(remove_dc_offset) F(round_to_power_of_two)
  F(snip_edges) F(window_type) F(num_bins) F(low_freq) F(high_freq) F(use_energy) F(energy_floor) F(raw_energy) F(htk_compat) F(htk_mode)
  // the fields only one feature type reads are compared for that type
  if (a.feature_type != 1) { F(num_ceps) F(cepstral_lifter) }
  if (a.feature_type == 1) { F(use_log_fbank) F(use_power) }
  if (a.feature_type == 2) { F(lpc_order) F(compress_factor) F(cepstral_scale) }
#undef F
}

template <class Options>
static b2k_feat_cfg through_kaldi(const char *path) {
  ParseOptions po("");
  Options o;
  o.Register(&po);
  po.ReadConfigFile(path);
  return b2k_shim::ToB2kFeatCfg(o);
}

int main(int argc, char **argv) {
  // argv: (type conf-path)*, type = mfcc | fbank | plp | endpoint
  REQUIRE(argc >= 3 && argc % 2 == 1);
  int checked = 0;
  for (int i = 1; i + 1 < argc; i += 2) {
    const std::string type = argv[i];
    const char *path = argv[i + 1];
    if (type == "endpoint") {
      ParseOptions po("");
      OnlineEndpointConfig e;
      e.Register(&po);
      po.ReadConfigFile(path);
      const b2k_endpoint_cfg a = b2k_shim::ToB2kEndpointConfig(e);
      b2k_endpoint_cfg b;
      REQUIRE(b2k_endpoint_cfg_from_conf(path, &b) == B2K_OK);
      for (int r = 0; r < 5; r++) {
        REQUIRE(a.rule[r].must_contain_nonsilence == b.rule[r].must_contain_nonsilence);
        REQUIRE(a.rule[r].min_trailing_silence == b.rule[r].min_trailing_silence);
        REQUIRE(a.rule[r].max_relative_cost == b.rule[r].max_relative_cost);
        REQUIRE(a.rule[r].min_utterance_length == b.rule[r].min_utterance_length);
      }
      REQUIRE(std::strcmp(a.silence_phones, b.silence_phones) == 0);
    } else {
      const int32_t t = type == "mfcc" ? 0 : type == "fbank" ? 1 : 2;
      REQUIRE(t != 2 || type == "plp");
      const b2k_feat_cfg a = t == 0 ? through_kaldi<MfccOptions>(path) : t == 1 ? through_kaldi<FbankOptions>(path) : through_kaldi<PlpOptions>(path);
      b2k_feat_cfg b;
      if (b2k_feat_cfg_from_conf(path, t, &b) != B2K_OK) { std::fprintf(stderr, "%s: %s\n", path, b2k_last_error()); return 1; }
      same_feat(a, b, path);
    }
    checked++;
  }
  std::printf("option mappings ok (%d files)\n", checked);
  return 0;
}
