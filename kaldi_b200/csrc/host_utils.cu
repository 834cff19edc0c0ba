// host_utils.cu — small host-only pieces of the online2 control flow that the batched pipeline needs
// (no device code).  Python twins in kaldi_b200/ivector.py are the test oracles (tests/test_host_utils.py).
#include <algorithm>
#include <cmath>

#include "common.cuh"

extern "C" {

// For each nnet chunk n, the frame index OnlineIvectorFeature::GetFrame is called with when
// online2-wav-nnet3-latgen-faster feeds `chunk_samples` at a time (online2-wav-nnet3-latgen-faster.cc:245-268):
// the chunk is computed by the first AdvanceDecoding after which DecodableNnetLoopedOnlineBase::NumFramesReady
// (decodable-online-looped.cc:56-84) covers it, and it asks for frame min(features_ready - 1,
// ivector_frames_ready - 1) (decodable-online-looped.cc:185-193), ivector_frames_ready being smaller by the
// splice right context until the input is finished.  snip-edges framing.
int b2k_ivec_online_schedule(int64_t num_samples, int32_t chunk_samples, int32_t frame_length, int32_t frame_shift,
                             int32_t num_feature_frames, int32_t nnet_right_context, int32_t frames_per_chunk,
                             int32_t subsampling, int32_t splice_right, int32_t *sched, int32_t max_chunks,
                             int32_t *n_chunks_out) {
  if (num_samples <= 0 || chunk_samples <= 0 || frame_length <= 0 || frame_shift <= 0 || num_feature_frames <= 0 ||
      frames_per_chunk <= 0 || subsampling <= 0 || !sched || !n_chunks_out)
    return b2k::set_error(B2K_ERR_INVALID, "b2k_ivec_online_schedule: bad args");
  const int32_t T = num_feature_frames;
  const int32_t n_out = (T + subsampling - 1) / subsampling;
  const int32_t n_chunks = (n_out * subsampling + frames_per_chunk - 1) / frames_per_chunk;
  *n_chunks_out = n_chunks;
  if (n_chunks > max_chunks) return b2k::set_error(B2K_ERR_OVERFLOW, "b2k_ivec_online_schedule: schedule buffer too small");
  int64_t fed = 0;
  int32_t done = 0;
  while (done < n_chunks) {
    fed = std::min<int64_t>(fed + chunk_samples, num_samples);
    const bool finished = fed >= num_samples;
    int32_t ready = fed < frame_length ? 0 : (int32_t)(1 + (fed - frame_length) / frame_shift);
    int32_t chunks_ready, iv_frame;
    if (finished) {
      ready = T; chunks_ready = n_chunks; iv_frame = T - 1;
    } else {
      chunks_ready = std::max(0, ready - nnet_right_context) / frames_per_chunk;
      const int32_t iv_ready = std::max(0, ready - splice_right);
      iv_frame = std::min(ready - 1, iv_ready - 1);
    }
    // no i-vector frame is ready yet (tiny chunks at the start of a file): the reference leaves the i-vector zero
    // (decodable-online-looped.cc:188-197); -1 tells b2k_ivec_compute_batched to do the same
    while (done < std::min(chunks_ready, n_chunks)) sched[done++] = iv_frame >= 0 ? iv_frame : -1;
  }
  return B2K_OK;
}

}  // extern "C"

// ------------------------------------------------------------------ Kaldi option files (conf/*.conf)
//
// ParseOptions::ReadConfigFile (util/parse-options.cc:460-497): '#' starts a comment, lines are trimmed, every non-empty
// line is --name=value or --name (a bool set to true), names are lower-cased with '_' -> '-' (NormalizeArgName, :523-536),
// bools accept true/t/1/"" and false/f/0 (ToBool, :569-586), an unknown name is an error.  The option names and defaults
// below are the ones MfccOptions / FbankOptions / FrameExtractionOptions / MelBanksOptions, OnlineIvectorExtractionConfig,
// OnlineSpliceOptions and OnlineCmvnOptions register (feat/feature-window.h:69-104, mel-computations.h:60-74,
// feature-mfcc.h:62-79, feature-fbank.h:62-80, online2/online-ivector-feature.h:112-160, feat/online-feature.h:234-251,
// :446-456); tests/test_conf_cpp.py compares with the reference's own ParseOptions on the same files.
#include <cerrno>
#include <cstdlib>
#include <limits>
#include <fstream>
#include <iterator>
#include <map>
#include <string>
#include <vector>

namespace {

struct ConfError { std::string msg; };

static std::string trim(const std::string &s) {
  const char *ws = " \t\n\r\f\v";
  const size_t a = s.find_first_not_of(ws);
  if (a == std::string::npos) return "";
  return s.substr(a, s.find_last_not_of(ws) - a + 1);
}

static std::map<std::string, std::pair<std::string, bool>> read_conf(const char *path) {     // name -> (value, has '=')
  std::ifstream is(path);
  if (!is.good()) throw ConfError{std::string("Cannot open config file: ") + path};
  std::map<std::string, std::pair<std::string, bool>> out;
  std::string line;
  int ln = 0;
  while (std::getline(is, line)) {
    ln++;
    const size_t h = line.find('#');
    if (h != std::string::npos) line.erase(h);
    line = trim(line);
    if (line.empty()) continue;
    if (line.compare(0, 2, "--") != 0) throw ConfError{std::string(path) + ": line " + std::to_string(ln) + " is not of the form --x=y"};
    const size_t eq = line.find('=');
    std::string key = eq == std::string::npos ? line.substr(2) : line.substr(2, eq - 2), val = eq == std::string::npos ? "" : trim(line.substr(eq + 1));
    for (auto &c : key) c = c == '_' ? '-' : (char)std::tolower((unsigned char)c);
    if (key.empty()) throw ConfError{std::string(path) + ": line " + std::to_string(ln) + ": empty option name"};
    out[key] = {val, eq != std::string::npos};
  }
  return out;
}

static bool to_bool(const std::string &name, std::string v) {
  for (auto &c : v) c = (char)std::tolower((unsigned char)c);
  if (v == "true" || v == "t" || v == "1" || v == "") return true;
  if (v == "false" || v == "f" || v == "0") return false;
  throw ConfError{"Invalid format for boolean argument --" + name + ": " + v};
}
// ConvertStringToInteger / ConvertStringToReal<float> (util/text-utils.h:118-136, text-utils.cc:165-246) as ParseOptions applies
// them: integers through strtoll with a range check (no hex prefix games: base 10), reals through istream >> float, i.e. only
// decimal notation, overflow is an error, and the spellings of infinity / NaN the reference lists are accepted by name.
static double to_num(const std::string &name, const std::string &v, bool integer) {
  const std::string what = "Invalid " + std::string(integer ? "integer" : "floating-point") + " option --" + name + "=" + v;
  if (v.empty()) throw ConfError{what};
  char *end = nullptr;
  if (integer) {
    errno = 0;
    const long long i = strtoll(v.c_str(), &end, 10);
    if (end == v.c_str() || *end != 0 || errno != 0 || i != (long long)(int32_t)i) throw ConfError{what};
    return (double)i;
  }
  if (v.find_first_not_of("0123456789+-.eE") == std::string::npos) {
    const float f = strtof(v.c_str(), &end);
    if (end == v.c_str() || *end != 0 || std::isinf(f)) throw ConfError{what};
    return (double)f;
  }
  std::string u = v;
  for (auto &c : u) c = (char)std::toupper((unsigned char)c);
  const double inf = std::numeric_limits<double>::infinity(), nan = std::numeric_limits<double>::quiet_NaN();
  if (u == "INF" || u == "+INF" || u == "INFINITY" || u == "+INFINITY" || u == "1.#INF") return inf;
  if (u == "-INF" || u == "-INFINITY" || u == "-1.#INF") return -inf;
  if (u == "NAN" || u == "+NAN" || u == "1.#QNAN") return nan;
  if (u == "-NAN" || u == "-1.#QNAN") return -nan;
  throw ConfError{what};
}

// one table entry per registered option: where it goes and what kind it is
struct Opt { const char *name; char kind; void *dst; };   // kind: b bool->int32, i int32, f float, s string (char[512]), x parsed but unused

static void apply(const char *path, const std::map<std::string, std::pair<std::string, bool>> &kv, const Opt *opts, size_t n) {
  for (auto &e : kv) {
    const Opt *o = nullptr;
    for (size_t i = 0; i < n; i++) if (e.first == opts[i].name) o = &opts[i];
    if (!o) throw ConfError{"Invalid option --" + e.first + " in config file " + path};
    const std::string &v = e.second.first;
    // parse-options.cc:540-547: only a bool may be given without '=' (--x means true); "--x" for any other type is an error
    if (o->kind != 'b' && !e.second.second) throw ConfError{"Invalid option --" + e.first + " (option format is --x=y)."};
    switch (o->kind) {
      case 'b':
        if (e.second.second && v.empty()) throw ConfError{"Invalid option --" + e.first + "="};      // parse-options.cc:545: --x is true, --x= is not
        *(int32_t *)o->dst = to_bool(e.first, v) ? 1 : 0;
        break;
      case 'i': *(int32_t *)o->dst = (int32_t)to_num(e.first, v, true); break;
      case 'f': *(float *)o->dst = (float)to_num(e.first, v, false); break;
      case 's': { char *d = (char *)o->dst; if (v.size() >= 512) throw ConfError{"value of --" + e.first + " is too long"}; memcpy(d, v.c_str(), v.size() + 1); break; }
      default: break;
    }
  }
}

}  // namespace

extern "C" {

int b2k_feat_cfg_from_conf(const char *conf_path, int32_t feature_type, b2k_feat_cfg *cfg) {
  if (!conf_path || !cfg || feature_type < 0 || feature_type > 2) return b2k::set_error(B2K_ERR_INVALID, "b2k_feat_cfg_from_conf: bad args");
  // the reference's own defaults (NOT mfcc_hires.conf): MfccOptions() / FbankOptions() / PlpOptions(), FrameExtractionOptions(), MelBanksOptions(23)
  b2k_feat_cfg c;
  memset(&c, 0, sizeof(c));
  c.feature_type = feature_type; c.samp_freq = 16000.f; c.frame_shift_ms = 10.f; c.frame_length_ms = 25.f; c.dither = 1.0f;
  c.preemph_coeff = 0.97f; c.remove_dc_offset = 1; c.round_to_power_of_two = 1; c.snip_edges = 1; c.window_type = 0;
  c.num_bins = 23; c.low_freq = 20.f; c.high_freq = 0.f; c.num_ceps = 13; c.use_energy = feature_type != 1 ? 1 : 0;
  c.lpc_order = 12; c.compress_factor = 0.33333f; c.cepstral_scale = 1.0f;       // PlpOptions (feat/feature-plp.h:55-66)
  int32_t plp_lifter = 22;                                                         // PlpOptions::cepstral_lifter is an int32
  c.energy_floor = 0.f; c.raw_energy = 1; c.cepstral_lifter = 22.f; c.htk_compat = 0; c.use_log_fbank = 1; c.use_power = 1; c.htk_mode = 0;
  c.max_lanes = cfg->max_lanes > 0 ? cfg->max_lanes : 1024;
  char window[512] = "povey";
  float blackman = 0.42f, vtln_low = 100.f, vtln_high = -500.f;
  int32_t allow_down = 0, allow_up = 0, max_fv = -1, debug_mel = 0;
  std::vector<Opt> o = {
      {"sample-frequency", 'f', &c.samp_freq}, {"frame-length", 'f', &c.frame_length_ms}, {"frame-shift", 'f', &c.frame_shift_ms},
      {"preemphasis-coefficient", 'f', &c.preemph_coeff}, {"remove-dc-offset", 'b', &c.remove_dc_offset}, {"dither", 'f', &c.dither},
      {"window-type", 's', window}, {"blackman-coeff", 'f', &blackman}, {"round-to-power-of-two", 'b', &c.round_to_power_of_two},
      {"snip-edges", 'b', &c.snip_edges}, {"allow-downsample", 'b', &allow_down}, {"allow-upsample", 'b', &allow_up},
      {"max-feature-vectors", 'i', &max_fv}, {"num-mel-bins", 'i', &c.num_bins}, {"low-freq", 'f', &c.low_freq}, {"high-freq", 'f', &c.high_freq},
      {"vtln-low", 'f', &vtln_low}, {"vtln-high", 'f', &vtln_high}, {"debug-mel", 'b', &debug_mel},
      {"use-energy", 'b', &c.use_energy}, {"energy-floor", 'f', &c.energy_floor}, {"raw-energy", 'b', &c.raw_energy}, {"htk-compat", 'b', &c.htk_compat}};
  if (feature_type == 0) { o.push_back({"num-ceps", 'i', &c.num_ceps}); o.push_back({"cepstral-lifter", 'f', &c.cepstral_lifter}); }
  else if (feature_type == 2) {                                                    // PlpOptions::Register (feature-plp.h:68-90)
    o.push_back({"lpc-order", 'i', &c.lpc_order}); o.push_back({"num-ceps", 'i', &c.num_ceps}); o.push_back({"compress-factor", 'f', &c.compress_factor});
    o.push_back({"cepstral-lifter", 'i', &plp_lifter}); o.push_back({"cepstral-scale", 'f', &c.cepstral_scale});
  } else { o.push_back({"use-log-fbank", 'b', &c.use_log_fbank}); o.push_back({"use-power", 'b', &c.use_power}); }
  try {
    apply(conf_path, read_conf(conf_path), o.data(), o.size());
    const std::string w = window;
    c.window_type = w == "povey" ? 0 : w == "hamming" ? 1 : w == "hanning" ? 2 : w == "rectangular" ? 3 : -1;
    if (c.window_type < 0) throw ConfError{"window type " + w + " is not supported (povey, hamming, hanning, rectangular)"};
    if (feature_type == 2) c.cepstral_lifter = (float)plp_lifter;
  } catch (const ConfError &e) {
    return b2k::set_error(B2K_ERR_INVALID, "b2k_feat_cfg_from_conf", e.msg.c_str());
  }
  *cfg = c;
  return B2K_OK;
}

int b2k_ivec_cfg_from_conf(const char *conf_path, b2k_ivec_cfg *cfg, b2k_ivec_paths *paths) {
  if (!conf_path || !cfg || !paths) return b2k::set_error(B2K_ERR_INVALID, "b2k_ivec_cfg_from_conf: bad args");
  b2k_ivec_cfg c = *cfg;                       // base_dim, max_lanes, max_frames and the dimensions stay the caller's
  b2k_ivec_paths p;
  memset(&p, 0, sizeof(p));
  // OnlineIvectorExtractionConfig(): online-ivector-feature.h:104-111
  c.num_gselect = 5; c.min_post = 0.025f; c.posterior_scale = 0.1f; c.max_count = 0.0f; c.num_cg_iters = 15;
  // OnlineSpliceOptions(): 4 / 4;  OnlineCmvnOptions(): 600 / 600 / 200
  c.splice_left = 4; c.splice_right = 4; c.cmn_window = 600; c.speaker_frames = 600; c.global_frames = 200;
  p.ivector_period = 10; p.use_most_recent_ivector = 1; p.max_remembered_frames = 1000.f;
  int32_t norm_vars = 0, norm_means = 1;
  char skip_dims[512] = "";
  const Opt top[] = {
      {"lda-matrix", 's', p.lda_matrix}, {"global-cmvn-stats", 's', p.global_cmvn_stats}, {"cmvn-config", 's', p.cmvn_config},
      {"online-cmvn-iextractor", 'b', &p.online_cmvn_iextractor}, {"splice-config", 's', p.splice_config}, {"diag-ubm", 's', p.diag_ubm},
      {"ivector-extractor", 's', p.ivector_extractor}, {"ivector-period", 'i', &p.ivector_period}, {"num-gselect", 'i', &c.num_gselect},
      {"min-post", 'f', &c.min_post}, {"posterior-scale", 'f', &c.posterior_scale}, {"max-count", 'f', &c.max_count},
      {"use-most-recent-ivector", 'b', &p.use_most_recent_ivector}, {"greedy-ivector-extractor", 'b', &p.greedy_ivector_extractor},
      {"max-remembered-frames", 'f', &p.max_remembered_frames}};
  const Opt splice[] = {{"left-context", 'i', &c.splice_left}, {"right-context", 'i', &c.splice_right}};
  const Opt cmvn[] = {{"cmn-window", 'i', &c.cmn_window}, {"global-frames", 'i', &c.global_frames}, {"speaker-frames", 'i', &c.speaker_frames},
                      {"norm-vars", 'b', &norm_vars}, {"norm-means", 'b', &norm_means}, {"skip-dims", 's', skip_dims}};
  try {
    apply(conf_path, read_conf(conf_path), top, sizeof(top) / sizeof(top[0]));
    if (p.splice_config[0]) apply(p.splice_config, read_conf(p.splice_config), splice, 2);     // OnlineIvectorExtractionInfo::Init reads both
    if (p.cmvn_config[0]) apply(p.cmvn_config, read_conf(p.cmvn_config), cmvn, sizeof(cmvn) / sizeof(cmvn[0]));
    if (norm_vars || !norm_means || skip_dims[0]) throw ConfError{"the extractor's online CMVN must be mean-only over all dimensions (norm-vars / skip-dims are not supported)"};
    if (p.online_cmvn_iextractor) throw ConfError{"--online-cmvn-iextractor=true is not supported"};
    // OnlineIvectorFeature has two more modes this library has no kernel for: an i-vector per requested frame instead of the
    // most recent one (online-ivector-feature.cc:327-349) and the greedy stats update (:392-402).  Values that would change
    // what the reference computes are errors, not silently ignored.
    if (!p.use_most_recent_ivector) throw ConfError{"--use-most-recent-ivector=false is not supported (the i-vector of a chunk is the most recent estimate)"};
    if (p.greedy_ivector_extractor) throw ConfError{"--greedy-ivector-extractor=true is not supported"};
  } catch (const ConfError &e) {
    return b2k::set_error(B2K_ERR_INVALID, "b2k_ivec_cfg_from_conf", e.msg.c_str());
  }
  *cfg = c;
  *paths = p;
  return B2K_OK;
}


// conf/online.conf: the --config file of online2-wav-nnet3-latgen-faster.  The tool registers several option groups into
// one ParseOptions; this reader takes the feature group (OnlineNnet2FeaturePipelineConfig::Register,
// online2/online-nnet2-feature-pipeline.h:101-126) and lets the other groups a recipe puts in the same file pass
// (--endpoint.*, --ivector-silence-weighting.*, the decoder / decodable options): they are returned verbatim in `rest`.
int b2k_online_conf_read(const char *conf_path, b2k_online_conf *out) {
  if (!conf_path || !out) return b2k::set_error(B2K_ERR_INVALID, "b2k_online_conf_read: bad args");
  b2k_online_conf c;
  memset(&c, 0, sizeof(c));
  char feature_type[512] = "mfcc", pitch_config[512] = "";
  const Opt opts[] = {{"feature-type", 's', feature_type}, {"mfcc-config", 's', c.mfcc_config}, {"plp-config", 's', c.plp_config},
                      {"fbank-config", 's', c.fbank_config}, {"cmvn-config", 's', c.cmvn_config}, {"global-cmvn-stats", 's', c.global_cmvn_stats},
                      {"add-pitch", 'b', &c.add_pitch}, {"online-pitch-config", 's', pitch_config},
                      {"ivector-extraction-config", 's', c.ivector_extraction_config}};
  try {
    auto kv = read_conf(conf_path);
    std::map<std::string, std::pair<std::string, bool>> mine;
    std::string rest;
    for (auto &e : kv) {
      bool known = false;
      for (auto &o : opts) known = known || e.first == o.name;
      if (known) mine[e.first] = e.second;
      else rest += "--" + e.first + (e.second.second ? "=" + e.second.first : "") + "\n";
    }
    apply(conf_path, mine, opts, sizeof(opts) / sizeof(opts[0]));
    const std::string ft = feature_type;
    c.feature_type = ft == "mfcc" ? 0 : ft == "fbank" ? 1 : ft == "plp" ? 2 : -1;
    if (c.feature_type < 0) throw ConfError{"Invalid feature type: " + ft + " (expected mfcc, plp or fbank)"};     // online-nnet2-feature-pipeline.cc:49-52
    if (c.add_pitch) throw ConfError{"--add-pitch=true is not supported"};
    if (rest.size() >= sizeof(c.rest)) throw ConfError{"too many other options in the file"};
    memcpy(c.rest, rest.c_str(), rest.size() + 1);
  } catch (const ConfError &e) {
    return b2k::set_error(B2K_ERR_INVALID, "b2k_online_conf_read", e.msg.c_str());
  }
  *out = c;
  return B2K_OK;
}

}  // extern "C"

// The decoder / decodable / tool options that may sit beside the feature group in the same option file or on the tool's command
// line: LatticeFasterDecoderConfig::Register (decoder/lattice-faster-decoder.h:75-98), NnetSimpleLoopedComputationOptions::Register
// (nnet3/decodable-simple-looped.h:66-88) and the tool's own --chunk-length (online2-wav-nnet3-latgen-faster.cc:112).  `text` is one
// option per line or blank-separated (b2k_online_conf.rest has that form).  Option groups this library does not act on are accepted
// and ignored by prefix (endpoint.*, ivector-silence-weighting.*, det.*) or by name (word-symbol-table, do-endpointing, online,
// num-threads-startup, determinize-lattice, memory-pool-*, debug-computation); values that would change what is computed and
// are not supported are errors (extra-left-context-initial != 0, a frame-subsampling-factor other than the model's is caught
// later by b2k_pipeline_plan_for through frames_per_chunk).
extern "C" int b2k_pipeline_cfg_apply_options(const char *text, b2k_pipeline_cfg *cfg) {
  if (!text || !cfg) return b2k::set_error(B2K_ERR_INVALID, "b2k_pipeline_cfg_apply_options: bad args");
  b2k_pipeline_cfg c = *cfg;
  int32_t prune_interval = c.dec.prune_interval, sub = 3, fpc = 20, extra_left = 0, ign_i = 0, online_flag = 1, do_endpointing = 0;
  float ign_f = 0.f;
  char ign_s[512] = "";
  const Opt opts[] = {{"beam", 'f', &c.dec.beam}, {"max-active", 'i', &c.dec.max_active}, {"min-active", 'i', &c.dec.min_active},
                      {"lattice-beam", 'f', &c.dec.lattice_beam}, {"prune-interval", 'i', &prune_interval}, {"beam-delta", 'f', &c.dec.beam_delta},
                      {"hash-ratio", 'f', &c.dec.hash_ratio}, {"acoustic-scale", 'f', &c.acoustic_scale}, {"frames-per-chunk", 'i', &fpc},
                      {"frame-subsampling-factor", 'i', &sub}, {"extra-left-context-initial", 'i', &extra_left}, {"chunk-length", 'f', &c.chunk_length_secs},
                      {"determinize-lattice", 'b', &ign_i}, {"memory-pool-tokens-block-size", 'i', &ign_i}, {"memory-pool-links-block-size", 'i', &ign_i},
                      {"debug-computation", 'b', &ign_i}, {"word-symbol-table", 's', ign_s}, {"do-endpointing", 'b', &do_endpointing}, {"online", 'b', &online_flag},
                      {"num-threads-startup", 'i', &ign_i}};
  (void)ign_f;
  try {
    std::map<std::string, std::pair<std::string, bool>> kv;
    bool fpc_given = false;
    std::string sil_phones;
    float sil_weight = 1.0f;
    std::string t(text), tok;
    size_t i = 0;
    while (i <= t.size()) {
      const bool end = i == t.size() || t[i] == '\n' || t[i] == ' ' || t[i] == '\t' || t[i] == '\r';
      if (!end) tok.push_back(t[i]);
      else if (!tok.empty()) {
        if (tok.compare(0, 2, "--") != 0) throw ConfError{"option " + tok + " does not start with --"};
        const size_t eq = tok.find('=');
        std::string key = eq == std::string::npos ? tok.substr(2) : tok.substr(2, eq - 2);
        for (auto &ch : key) ch = ch == '_' ? '-' : (char)std::tolower((unsigned char)ch);
        const std::string val = eq == std::string::npos ? "" : tok.substr(eq + 1);
        if (key.compare(0, 9, "endpoint.") != 0 && key.compare(0, 26, "ivector-silence-weighting.") != 0 && key.compare(0, 4, "det.") != 0) {
          kv[key] = {val, eq != std::string::npos};
          if (key == "frames-per-chunk") fpc_given = true;
        }
        // OnlineSilenceWeightingConfig::Active() (online-ivector-feature.h:414): silence-phones given and silence-weight != 1
        // re-weights the i-vector statistics from the decoder's traceback, which this pipeline does not do
        if (key == "ivector-silence-weighting.silence-phones") sil_phones = val;
        if (key == "ivector-silence-weighting.silence-weight") sil_weight = (float)to_num(key, val, false);
        tok.clear();
      }
      i++;
    }
    apply("(options)", kv, opts, sizeof(opts) / sizeof(opts[0]));
    if (extra_left != 0) throw ConfError{"--extra-left-context-initial other than 0 is not supported"};
    if (!sil_phones.empty() && sil_weight != 1.0f)
      throw ConfError{"--ivector-silence-weighting.silence-weight other than 1 with silence phones is not supported (i-vector statistics are not re-weighted from the traceback)"};
    if (do_endpointing) throw ConfError{"--do-endpointing=true is not supported by the batched pipeline (whole utterances are decoded; the endpoint rules are b2k_endpoint_*)"};
    if (!online_flag) c.chunk_length_secs = -1.0f;         // --online=false: the tool sets chunk_length_secs = -1 = the whole file in one call (:128-130)
    if (sub <= 0) throw ConfError{"--frame-subsampling-factor must be positive"};
    c.dec.prune_interval = prune_interval;
    if (kv.count("frame-subsampling-factor")) c.frame_subsampling_factor = sub;   // carried to the compile (never inferred from a default)
    else if (c.frame_subsampling_factor > 0) sub = c.frame_subsampling_factor;
    if (fpc_given) {                                          // GetChunkSize (nnet-compile-looped.cc:81): rounded up to a multiple of the subsampling factor
      if (fpc <= 0) throw ConfError{"--frames-per-chunk must be positive"};
      c.frames_per_chunk = (fpc + sub - 1) / sub * sub;
    }
  } catch (const ConfError &e) {
    return b2k::set_error(B2K_ERR_INVALID, "b2k_pipeline_cfg_apply_options", e.msg.c_str());
  }
  *cfg = c;
  return B2K_OK;
}

// ------------------------------------------------------------------ RIFF/WAVE input
//
// WaveData::Read (feat/wave-reader.cc:107-321) in our own words: RIFF or RIFX (byte-swapped), any chunks before "fmt " and
// between "fmt " and "data" skipped, format 1 (PCM) or 0xFFFE (extensible with the PCM sub-format GUID), 16 bits only,
// byte rate and block align must be consistent, "stream mode" (read to the end of the file) when the RIFF or data size is
// 0 / 0xFFFFFFFF / 0x7FFFF000 (SoX), a short file gives the samples that are there, no data at all is an error.  Samples
// stay in the int16 range as floats (Kaldi's convention, feat/wave-reader.h:60-62), one row per channel.
struct b2k_wave { float samp_freq = 0; int32_t channels = 0; int64_t samples = 0; std::vector<float> data; };

namespace {
struct WavIn {
  std::vector<unsigned char> d; size_t p = 0; bool swap = false;
  void need(size_t n) const { if (p + n > d.size()) throw ConfError{"WaveData: unexpected end of file or read error"}; }
  std::string tag() { need(4); std::string t((const char *)&d[p], 4); p += 4; return t; }
  uint32_t u32() { need(4); uint32_t v; unsigned char b[4] = {d[p], d[p + 1], d[p + 2], d[p + 3]}; if (swap) { std::swap(b[0], b[3]); std::swap(b[1], b[2]); } memcpy(&v, b, 4); p += 4; return v; }
  uint16_t u16() { need(2); uint16_t v; unsigned char b[2] = {d[p], d[p + 1]}; if (swap) std::swap(b[0], b[1]); memcpy(&v, b, 2); p += 2; return v; }
  void skip(uint32_t n) { p = std::min(d.size(), p + (size_t)n); }      // the reference's is.get() loop does not fail at EOF either
};
}  // namespace

extern "C" {

int b2k_wave_read(const char *path, b2k_wave **out) {
  if (!path || !out) return b2k::set_error(B2K_ERR_INVALID, "b2k_wave_read: bad args");
  b2k_wave *W = new b2k_wave();
  try {
    WavIn r;
    {
      std::ifstream is(path, std::ios::binary);
      if (!is.good()) throw ConfError{std::string("cannot open ") + path};
      r.d.assign(std::istreambuf_iterator<char>(is), std::istreambuf_iterator<char>());
    }
    const std::string riff = r.tag();
    if (riff == "RIFX") r.swap = true;
    else if (riff != "RIFF") throw ConfError{"WaveData: expected RIFF or RIFX, got " + riff};
    const uint32_t riff_size = r.u32();
    if (r.tag() != "WAVE") throw ConfError{"WaveData: expected WAVE"};
    std::string t = r.tag();
    while (t != "fmt ") { r.skip(r.u32()); t = r.tag(); }
    const uint32_t fmt_size = r.u32();
    const uint16_t format = r.u16();
    // WaveInfo keeps the channel count in a uint8 (wave-reader.h:101): the 16-bit field of the file is truncated, and every later check
    // (no channels, byte rate, block_align) sees the truncated value.  Mirrored, so that the same files are accepted with the same shape.
    const uint16_t channels = (uint16_t)(r.u16() & 0xFFu);
    const uint32_t rate = r.u32(), byte_rate = r.u32();
    const uint32_t block_align = r.u16(), bits = r.u16();
    uint32_t fmt_read = 16;
    if (format == 1) {
      if (fmt_size < 16) throw ConfError{"WaveData: expect PCM format data to have fmt chunk of at least size 16."};
    } else if (format == 0xFFFE) {
      const uint16_t extra = r.u16();
      if (fmt_size < 40 || extra < 22) throw ConfError{"WaveData: malformed WAVE_FORMAT_EXTENSIBLE format data."};
      r.u16(); r.u32();
      const uint32_t g1 = r.u32(), g2 = r.u32(), g3 = r.u32(), g4 = r.u32();
      fmt_read = 40;
      if (g1 != 0x00000001u || g2 != 0x00100000u || g3 != 0xAA000080u || g4 != 0x719B3800u) throw ConfError{"WaveData: unsupported WAVE_FORMAT_EXTENSIBLE format."};
    } else {
      throw ConfError{"WaveData: can read only PCM data, format id in file is: " + std::to_string(format)};
    }
    if (fmt_size > fmt_read) r.skip(fmt_size - fmt_read);
    if (channels == 0) throw ConfError{"WaveData: no channels present"};
    if (bits != 16) throw ConfError{"WaveData: unsupported bits_per_sample = " + std::to_string(bits)};
    if (byte_rate != rate * (bits / 8) * channels) throw ConfError{"WaveData: unexpected byte rate"};
    if (block_align != (uint32_t)channels * (bits / 8)) throw ConfError{"WaveData: unexpected block_align"};
    t = r.tag();
    while (t != "data") { r.skip(r.u32()); t = r.tag(); }
    const uint32_t data_size = r.u32();
    const bool stream = riff_size == 0 || riff_size == 0xFFFFFFFFu || data_size == 0 || data_size == 0xFFFFFFFFu || data_size == 0x7FFFF000u;
    size_t avail = r.d.size() - r.p;
    if (!stream) avail = std::min<size_t>(avail, (size_t)(data_size / block_align) * block_align);   // DataBytes() = samp_count * BlockAlign()
    if (avail == 0) throw ConfError{"WaveData: empty file (no data)"};
    const size_t ns = avail / block_align;
    if (ns == 0) throw ConfError{"WaveData: less than one sample block of data"};     // the reference cannot read such a file either (Matrix::Resize(channels, 0) asserts)
    W->samp_freq = (float)rate; W->channels = channels; W->samples = (int64_t)ns;
    W->data.resize((size_t)channels * ns);
    for (size_t i = 0; i < ns; i++)
      for (int j = 0; j < channels; j++) {
        unsigned char b0 = r.d[r.p], b1 = r.d[r.p + 1];
        r.p += 2;
        if (r.swap) std::swap(b0, b1);
        W->data[(size_t)j * ns + i] = (float)(int16_t)((uint16_t)b0 | ((uint16_t)b1 << 8));
      }
  } catch (const ConfError &e) {
    delete W;
    return b2k::set_error(B2K_ERR_INVALID, "b2k_wave_read", e.msg.c_str());
  }
  *out = W;
  return B2K_OK;
}

int b2k_wave_destroy(b2k_wave *w) { delete w; return B2K_OK; }
int b2k_wave_info(const b2k_wave *w, float *samp_freq, int32_t *channels, int64_t *samples) {
  if (!w) return b2k::set_error(B2K_ERR_INVALID, "b2k_wave_info: bad args");
  if (samp_freq) *samp_freq = w->samp_freq;
  if (channels) *channels = w->channels;
  if (samples) *samples = w->samples;
  return B2K_OK;
}
const float *b2k_wave_data(const b2k_wave *w) { return w ? w->data.data() : nullptr; }

}  // extern "C"

// ------------------------------------------------------------------ sample-rate conversion of a whole waveform
//
// ResampleWaveform (feat/resample.cc:368-376): LinearResample with a low-pass at 0.99 * Nyquist of the lower rate and 6
// zero crossings, flush = true.  What OnlineGenericBaseFeature::AcceptWaveform applies when --allow-downsample /
// --allow-upsample lets a file of another rate in (feat/online-feature.cc:138-150).  Output sample k sits at time k / new_rate;
// it is the dot product of the input around that time with a Hann-windowed sinc (resample.h:40-106), whose weights repeat
// with the period of gcd(orig_rate, new_rate) and are computed once per phase.  The filter function is evaluated in float
// like the reference's (FilterFunc takes and returns BaseFloat), the weights' time arguments in double.
extern "C" int b2k_resample_waveform(float orig_freq, const float *in, int64_t n_in, float new_freq, float *out, int64_t cap, int64_t *n_out) {
  if (!n_out || n_in < 0 || (n_in > 0 && !in) || cap < 0 || (cap > 0 && !out) || !(orig_freq > 0.f) || !(new_freq > 0.f) ||
      orig_freq != (float)(int32_t)orig_freq || new_freq != (float)(int32_t)new_freq)
    return b2k::set_error(B2K_ERR_INVALID, "b2k_resample_waveform: bad args (rates must be positive whole numbers of Hz)");
  const int32_t rin = (int32_t)orig_freq, rout = (int32_t)new_freq;
  const float min_freq = std::min(orig_freq, new_freq);
  const float cutoff = 0.99f * 0.5f * min_freq;              // BaseFloat arithmetic, as in ResampleWaveform
  const int32_t num_zeros = 6;
  auto gcd = [](int32_t a, int32_t b) { while (b) { const int32_t t = a % b; a = b; b = t; } return a; };
  const int32_t base = gcd(rin, rout), in_unit = rin / base, out_unit = rout / base;
  // number of output samples: the largest k with k / rout < n_in / rin (GetNumOutputSamples with flush = true)
  const int64_t tick = (int64_t)rin / base * rout;            // lcm
  const int64_t per_in = tick / rin, per_out = tick / rout, interval = n_in * per_in;
  int64_t total = 0;
  if (interval > 0) { int64_t last = interval / per_out; if (last * per_out == interval) last--; total = last + 1; }
  *n_out = total;
  if (total > cap) return b2k::set_error(B2K_ERR_OVERFLOW, "b2k_resample_waveform: output buffer too small (size returned)");
  auto filter_func = [&](float t) -> float {                  // LinearResample::FilterFunc (resample.cc:246-259)
    float window, filter;
    if (fabs(t) < num_zeros / (2.0 * cutoff)) window = (float)(0.5 * (1 + cos(6.283185307179586476925286766559005 * cutoff / num_zeros * t)));
    else window = 0.0f;
    if (t != 0) filter = (float)(sin(6.283185307179586476925286766559005 * cutoff * t) / (3.1415926535897932384626433832795 * t));
    else filter = 2 * cutoff;
    return filter * window;
  };
  const double window_width = num_zeros / (2.0 * cutoff);
  std::vector<int32_t> first((size_t)out_unit);
  std::vector<std::vector<float>> weights((size_t)out_unit);
  for (int32_t i = 0; i < out_unit; i++) {
    const double output_t = i / (double)rout, min_t = output_t - window_width, max_t = output_t + window_width;
    const int32_t lo = (int32_t)ceil(min_t * rin), hi = (int32_t)floor(max_t * rin);
    first[i] = lo;
    weights[i].resize((size_t)std::max(0, hi - lo + 1));
    for (int32_t j = 0; j <= hi - lo; j++) {
      const double input_t = (lo + j) / (double)rin, delta_t = input_t - output_t;
      weights[i][j] = filter_func((float)delta_t) / rin;
    }
  }
  for (int64_t k = 0; k < total; k++) {
    const int64_t unit = k / out_unit;
    const int32_t ph = (int32_t)(k - unit * out_unit);
    const int64_t f0 = first[ph] + unit * in_unit;
    const std::vector<float> &w = weights[ph];
    float acc = 0.0f;
    for (size_t j = 0; j < w.size(); j++) {
      const int64_t idx = f0 + (int64_t)j;
      if (idx >= 0 && idx < n_in) acc += w[j] * in[idx];      // beyond either end: zero (flush = true, no remainder)
    }
    out[k] = acc;
  }
  return B2K_OK;
}

// ---------------------------------------------------------------------------------------------------------------------
// Endpointing (online2/online-endpoint.{h,cc}): the rule set, its option group and the two quantities the rules are
// evaluated on.  Host only.  The best path comes from b2k_lat_best_path_arcs / CudaDecoderB2k::GetBestPath; the relative
// cost of the final states is the caller's (FinalRelativeCost of the decoder, lattice-faster-decoder.cc:283).
namespace {

static void endpoint_defaults(b2k_endpoint_cfg *c) {            // OnlineEndpointConfig() (online-endpoint.h:146-151)
  const float inf = std::numeric_limits<float>::infinity();
  memset(c, 0, sizeof(*c));
  c->rule[0] = {0, 5.0f, inf, 0.0f};
  c->rule[1] = {1, 0.5f, 2.0f, 0.0f};
  c->rule[2] = {1, 1.0f, 8.0f, 0.0f};
  c->rule[3] = {1, 2.0f, inf, 0.0f};
  c->rule[4] = {0, 0.0f, inf, 20.0f};
}

// the endpoint.* entries of `kv` -> cfg; an endpoint.* name the group does not register is an error, other names are not looked at
static void endpoint_apply(const char *where, const std::map<std::string, std::pair<std::string, bool>> &kv, b2k_endpoint_cfg *c) {
  std::vector<Opt> o = {{"endpoint.silence-phones", 's', c->silence_phones}};
  static const char *const names[5][4] = {
      {"endpoint.rule1.must-contain-nonsilence", "endpoint.rule1.min-trailing-silence", "endpoint.rule1.max-relative-cost", "endpoint.rule1.min-utterance-length"},
      {"endpoint.rule2.must-contain-nonsilence", "endpoint.rule2.min-trailing-silence", "endpoint.rule2.max-relative-cost", "endpoint.rule2.min-utterance-length"},
      {"endpoint.rule3.must-contain-nonsilence", "endpoint.rule3.min-trailing-silence", "endpoint.rule3.max-relative-cost", "endpoint.rule3.min-utterance-length"},
      {"endpoint.rule4.must-contain-nonsilence", "endpoint.rule4.min-trailing-silence", "endpoint.rule4.max-relative-cost", "endpoint.rule4.min-utterance-length"},
      {"endpoint.rule5.must-contain-nonsilence", "endpoint.rule5.min-trailing-silence", "endpoint.rule5.max-relative-cost", "endpoint.rule5.min-utterance-length"}};
  for (int r = 0; r < 5; r++) {
    o.push_back({names[r][0], 'b', &c->rule[r].must_contain_nonsilence});
    o.push_back({names[r][1], 'f', &c->rule[r].min_trailing_silence});
    o.push_back({names[r][2], 'f', &c->rule[r].max_relative_cost});
    o.push_back({names[r][3], 'f', &c->rule[r].min_utterance_length});
  }
  std::map<std::string, std::pair<std::string, bool>> mine;
  for (auto &e : kv) if (e.first.compare(0, 9, "endpoint.") == 0) mine[e.first] = e.second;
  apply(where, mine, o.data(), o.size());
}

// "1:2:3" -> sorted phones; SplitStringToIntegers(":", omit_empty = false) + the two asserts of TrailingSilenceLength
// (online-endpoint.cc:82-91): no empty fields, no trailing characters after a number, no duplicates, not empty
static std::vector<int32_t> silence_set(const char *str) {
  std::vector<int32_t> out;
  const std::string s(str);
  size_t a = 0;
  while (!s.empty()) {
    const size_t b = s.find(':', a);
    const std::string f = s.substr(a, b == std::string::npos ? std::string::npos : b - a);
    char *end = nullptr;
    errno = 0;
    const long long v = strtoll(f.c_str(), &end, 10);
    if (end == f.c_str() || *end != 0 || v != (long long)(int32_t)v) throw ConfError{"Bad --silence-phones option in endpointing config: " + s};
    out.push_back((int32_t)v);
    if (b == std::string::npos) break;
    a = b + 1;
  }
  if (out.empty()) throw ConfError{"Endpointing requires nonempty --endpoint.silence-phones option"};
  std::sort(out.begin(), out.end());
  if (std::adjacent_find(out.begin(), out.end()) != out.end()) throw ConfError{"Duplicates in --silence-phones option in endpointing config"};
  return out;
}

static bool rule_activated(const b2k_endpoint_rule &r, float trailing_silence, float relative_cost, float utterance_length) {
  const bool contains_nonsilence = utterance_length > trailing_silence;           // online-endpoint.cc:31-37
  return (contains_nonsilence || !r.must_contain_nonsilence) && trailing_silence >= r.min_trailing_silence &&
         relative_cost <= r.max_relative_cost && utterance_length >= r.min_utterance_length;
}

static int32_t trailing_silence(const int32_t *tid2phone, int32_t num_tids, const std::vector<int32_t> &sil, const int32_t *ilabels, int64_t n) {
  int32_t count = 0;
  for (int64_t i = n - 1; i >= 0; i--) {                                            // backwards in time from the last decoded frame
    const int32_t t = ilabels[i];
    if (t == 0) continue;
    if (t < 0 || t >= num_tids) throw ConfError{"best path holds a transition-id outside the model's range"};
    if (!std::binary_search(sil.begin(), sil.end(), tid2phone[t])) break;
    count++;
  }
  return count;
}

}  // namespace

extern "C" {

int b2k_endpoint_cfg_default(b2k_endpoint_cfg *cfg) {
  if (!cfg) return b2k::set_error(B2K_ERR_INVALID, "b2k_endpoint_cfg_default: bad args");
  endpoint_defaults(cfg);
  return B2K_OK;
}

int b2k_endpoint_cfg_from_conf(const char *conf_path, b2k_endpoint_cfg *cfg) {
  if (!conf_path || !cfg) return b2k::set_error(B2K_ERR_INVALID, "b2k_endpoint_cfg_from_conf: bad args");
  b2k_endpoint_cfg c;
  endpoint_defaults(&c);
  try {
    endpoint_apply(conf_path, read_conf(conf_path), &c);
  } catch (const ConfError &e) {
    return b2k::set_error(B2K_ERR_INVALID, "b2k_endpoint_cfg_from_conf", e.msg.c_str());
  }
  *cfg = c;
  return B2K_OK;
}

int b2k_endpoint_cfg_apply_options(const char *text, b2k_endpoint_cfg *cfg) {
  if (!text || !cfg) return b2k::set_error(B2K_ERR_INVALID, "b2k_endpoint_cfg_apply_options: bad args");
  b2k_endpoint_cfg c = *cfg;
  try {
    std::map<std::string, std::pair<std::string, bool>> kv;
    std::string t(text), tok;
    for (size_t i = 0; i <= t.size(); i++) {
      const bool end = i == t.size() || t[i] == '\n' || t[i] == ' ' || t[i] == '\t' || t[i] == '\r';
      if (!end) { tok.push_back(t[i]); continue; }
      if (tok.empty()) continue;
      if (tok.compare(0, 2, "--") != 0) throw ConfError{"option " + tok + " does not start with --"};
      const size_t eq = tok.find('=');
      std::string key = eq == std::string::npos ? tok.substr(2) : tok.substr(2, eq - 2);
      for (auto &ch : key) ch = ch == '_' ? '-' : (char)std::tolower((unsigned char)ch);
      kv[key] = {eq == std::string::npos ? "" : tok.substr(eq + 1), eq != std::string::npos};
      tok.clear();
    }
    endpoint_apply("(options)", kv, &c);
  } catch (const ConfError &e) {
    return b2k::set_error(B2K_ERR_INVALID, "b2k_endpoint_cfg_apply_options", e.msg.c_str());
  }
  *cfg = c;
  return B2K_OK;
}

int b2k_endpoint_detected(const b2k_endpoint_cfg *cfg, int32_t num_frames_decoded, int32_t trailing_silence_frames,
                          float frame_shift_in_seconds, float final_relative_cost, int32_t *detected) {
  if (!cfg || !detected) return b2k::set_error(B2K_ERR_INVALID, "b2k_endpoint_detected: bad args");
  if (num_frames_decoded < trailing_silence_frames)              // KALDI_ASSERT (online-endpoint.cc:52)
    return b2k::set_error(B2K_ERR_INVALID, "b2k_endpoint_detected: more trailing silence than decoded frames");
  const float utterance_length = num_frames_decoded * frame_shift_in_seconds, trailing = trailing_silence_frames * frame_shift_in_seconds;
  int32_t ans = 0;
  for (int r = 0; r < 5 && !ans; r++) ans = rule_activated(cfg->rule[r], trailing, final_relative_cost, utterance_length) ? 1 : 0;
  *detected = ans;
  return B2K_OK;
}

int b2k_trailing_silence_frames(const int32_t *tid2phone, int32_t num_tids, const char *silence_phones, const int32_t *ilabels,
                                int64_t n, int32_t *frames) {
  if (!tid2phone || num_tids <= 0 || !silence_phones || (!ilabels && n > 0) || n < 0 || !frames)
    return b2k::set_error(B2K_ERR_INVALID, "b2k_trailing_silence_frames: bad args");
  try {
    *frames = trailing_silence(tid2phone, num_tids, silence_set(silence_phones), ilabels, n);
  } catch (const ConfError &e) {
    return b2k::set_error(B2K_ERR_INVALID, "b2k_trailing_silence_frames", e.msg.c_str());
  }
  return B2K_OK;
}

int b2k_endpoint_detected_on_path(const b2k_endpoint_cfg *cfg, const int32_t *tid2phone, int32_t num_tids, const int32_t *ilabels, int64_t n,
                                  int32_t num_frames_decoded, float frame_shift_in_seconds, float final_relative_cost,
                                  int32_t *detected, int32_t *trailing_silence_frames) {
  if (!cfg || !tid2phone || num_tids <= 0 || (!ilabels && n > 0) || n < 0 || num_frames_decoded < 0 || !detected)
    return b2k::set_error(B2K_ERR_INVALID, "b2k_endpoint_detected_on_path: bad args");
  if (trailing_silence_frames) *trailing_silence_frames = 0;
  if (num_frames_decoded == 0) { *detected = 0; return B2K_OK; }                  // online-endpoint.cc:123
  int32_t sil = 0;
  try {
    sil = trailing_silence(tid2phone, num_tids, silence_set(cfg->silence_phones), ilabels, n);
  } catch (const ConfError &e) {
    return b2k::set_error(B2K_ERR_INVALID, "b2k_endpoint_detected_on_path", e.msg.c_str());
  }
  if (trailing_silence_frames) *trailing_silence_frames = sil;
  return b2k_endpoint_detected(cfg, num_frames_decoded, sil, frame_shift_in_seconds, final_relative_cost, detected);
}

}  // extern "C"
