// nnet_compile.cu — host-side (no device code) compiler from a chain-model layer list to the op program that
// nnet.cu executes.  C++ counterpart of kaldi_b200/nnet_model.py (build_graph + compile_program), which stays
// as its test oracle: tests/test_nnet_compile_cpp.py requires the two to emit identical nodes, ops and blobs.
//
// It plays the role of the reference's nnet3 compiler for this model family (nnet3/nnet-compile.cc,
// nnet-compile-looped.cc, nnet-optimize.cc): per node it derives the time grid (step 1 before the first
// stride-3 layer, 3 after) and the range the outputs need (the reference also computes only required Indexes),
// turns Append/Offset/Sum/Scale/ReplaceIndex descriptors into per-term row maps (no splice copies), folds
// test-mode BatchNorm / ReLU / bypass / -log prior / acoustic scale into GEMM epilogues, maps
// TimeHeightConvolutionComponent to height-split GEMM ops, and lays the activations out in a per-utterance
// arena with liveness-based reuse.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <map>
#include <set>
#include <string>
#include <vector>

#include "common.cuh"

namespace {

using b2k::set_error;

struct Term { std::string src; int off, c0, c1; int kind; };      // kind: 0 row, 1 ivec (chunk-indexed source), 2 chunk (identity over chunks)
struct EwTerm { std::string src; int off; float scale; };
struct ColWin { int step, off, lim; };

struct Node {
  std::string name;
  int dim = 0;
  int kind = 0;                 // 0 input, 1 ivector, 2 gemm, 3 ew
  std::set<int> residues;
  int tmin = 1000000000, tmax = -1000000000, step = 0, t0 = 0, rows = 0;
  // gemm
  std::vector<Term> terms;
  std::vector<ColWin> term_cols;
  int hsplit = 1;
  const float *w = nullptr; int w_rows = 0, w_cols = 0;
  std::vector<float> w_own;     // expanded convolution weights
  const float *b = nullptr; std::vector<float> b_own; bool has_b = false; long long b_size = 0;
  bool relu = false, has_bn = false, has_res = false, log_softmax = false, ivector_rows = false;
  std::vector<float> bn_scale, bn_offset;
  std::string res_src; float res_alpha = 0.f;
  // ew
  std::vector<std::vector<EwTerm>> blocks;
  int block_dim = 0;
};

struct Weights {
  std::map<std::string, const b2k_nnet_weight *> m;
  const b2k_nnet_weight *get(const std::string &k) const {
    auto it = m.find(k);
    return it == m.end() ? nullptr : it->second;
  }
};

// BatchNormComponent::ComputeDerived (nnet-normalize-component.cc:209-246): scale = (var + eps)^-0.5 * target_rms, offset = -mean * scale
static const float kBnEps = 1e-3f;
static bool bn_scale_offset(const Weights &W, const std::string &name, float target_rms, int tile, std::vector<float> *scale,
                            std::vector<float> *offset) {
  const b2k_nnet_weight *mean = W.get(name + ".mean"), *var = W.get(name + ".var");
  if (!mean || !var || mean->size != var->size) return false;
  const int n = (int)mean->size;
  scale->resize((size_t)n * tile);
  offset->resize((size_t)n * tile);
  for (int i = 0; i < n; i++) {
    float s = powf(std::max(var->data[i], 0.0f) + kBnEps, -0.5f);
    if (target_rms != 1.0f) s = s * target_rms;
    const float o = -(mean->data[i]) * s;
    for (int t = 0; t < tile; t++) { (*scale)[(size_t)t * n + i] = s; (*offset)[(size_t)t * n + i] = o; }
  }
  return true;
}

struct Graph {
  std::vector<Node> nodes;
  std::map<std::string, int> dims;
  Node *find(const std::string &n) { for (auto &x : nodes) if (x.name == n) return &x; return nullptr; }
};

static std::string S(const char *c, size_t cap) { return std::string(c, strnlen(c, cap)); }
// the time offsets a layer splices its input over: time_offsets when given, else the recipe default of the layer kind
static std::vector<int> splice_of(const b2k_nnet_layer &L, std::vector<int> dflt) {
  if (L.n_time_offsets <= 0) return dflt;
  return std::vector<int>(L.time_offsets, L.time_offsets + std::min(L.n_time_offsets, 8));
}

#define NEED(ptr, what) do { if (!(ptr)) { err = std::string("missing weight ") + (what); return false; } } while (0)

static bool set_w(Node *n, const b2k_nnet_weight *w) { n->w = w->data; n->w_rows = w->rows; n->w_cols = w->cols; return true; }

// expand_conv_weights of nnet_model.py: one dense [H_out*F_out, H_in*F_in] block per time offset, the
// combine-feature-maps interleave folded into the column order ([main columns | side columns])
static void expand_conv(const b2k_nnet_layer &L, const float *w, const float *b, const b2k_nnet_layer *combine, Node *n,
                        int *k_main, int *k_side) {
  const int Hi = L.height_in, Ho = L.height_out, sub = L.height_subsample_out, Fi = L.filters_in, Fo = L.filters_out;
  const int nt = L.n_time_offsets, nh = L.n_height_offsets, K = Hi * Fi;
  int f1 = Fi, f2 = 0;
  if (combine) { f1 = combine->filters1; f2 = combine->filters2; }
  *k_main = Hi * f1; *k_side = Hi * f2;
  n->w_own.assign((size_t)Ho * Fo * nt * K, 0.0f);
  const int wc = nt * K, src_cols = nt * nh * Fi;
  for (int ti = 0; ti < nt; ti++)
    for (int hi = 0; hi < nh; hi++) {
      const int dh = L.height_offsets[hi];
      for (int ho = 0; ho < Ho; ho++) {
        const int h_in = ho * sub + dh;
        if (h_in < 0 || h_in >= Hi) continue;
        for (int fo = 0; fo < Fo; fo++) {
          const float *srow = w + (size_t)fo * src_cols + (size_t)(ti * nh + hi) * Fi;
          float *drow = n->w_own.data() + (size_t)(ho * Fo + fo) * wc + (size_t)ti * K;
          for (int c = 0; c < f1; c++) drow[h_in * f1 + c] = srow[c];
          for (int c = 0; c < f2; c++) drow[*k_main + h_in * f2 + c] = srow[f1 + c];
        }
      }
    }
  n->w = n->w_own.data(); n->w_rows = Ho * Fo; n->w_cols = wc;
  n->b_own.resize((size_t)Ho * Fo);
  for (int ho = 0; ho < Ho; ho++) for (int fo = 0; fo < Fo; fo++) n->b_own[(size_t)ho * Fo + fo] = b[fo];
  n->b = n->b_own.data(); n->has_b = true; n->b_size = (long long)n->b_own.size();
}

static bool build_graph(const b2k_nnet_compile_cfg &cfg, const b2k_nnet_layer *layers, int n_layers, const Weights &W,
                        bool structural, Graph *g, std::string &err) {
  const int fd = cfg.feat_dim, ivd = cfg.ivector_dim;
  g->nodes.clear();
  g->nodes.reserve(2 * n_layers + 4);
  { Node a; a.name = "input"; a.dim = fd; a.kind = 0; g->nodes.push_back(a); }
  { Node a; a.name = "ivector"; a.dim = ivd; a.kind = 1; g->nodes.push_back(a); }
  g->dims["input"] = fd; g->dims["ivector"] = ivd;
  std::string cur = "input";
  const b2k_nnet_layer *pending_combine = nullptr;
  auto add = [&](Node &n) { g->dims[n.name] = n.dim; g->nodes.push_back(std::move(n)); };
  for (int li = 0; li < n_layers; li++) {
    const b2k_nnet_layer &L = layers[li];
    const std::string t = S(L.type, sizeof(L.type)), n = S(L.name, sizeof(L.name));
    auto bn = [&](Node *nd, const std::string &name, float rms, int tile) -> bool {
      nd->has_bn = true;
      if (structural) return true;
      if (!bn_scale_offset(W, name, rms, tile, &nd->bn_scale, &nd->bn_offset)) { err = "missing batchnorm statistics " + name; return false; }
      return true;
    };
    auto lin = [&](Node *nd, const std::string &wname, bool bias) -> bool {
      if (structural) return true;
      const b2k_nnet_weight *w = W.get(wname + ".w");
      NEED(w, wname + ".w");
      set_w(nd, w);
      if (bias) { const b2k_nnet_weight *b = W.get(wname + ".b"); NEED(b, wname + ".b"); nd->b = b->data; nd->has_b = true; nd->b_size = b->size; }
      return true;
    };
    if (t == "idct") {
      Node x; x.name = n; x.dim = fd; x.kind = 2; x.terms = {{cur, 0, 0, fd, 0}};
      if (!lin(&x, n, true)) return false;
      add(x); cur = n;
    } else if (t == "batchnorm") {          // folded into the producing node's epilogue
      Node *prod = g->find(cur);
      if (!prod) { err = "batchnorm without producer"; return false; }
      if (!bn(prod, n, 1.0f, 1)) return false;
    } else if (t == "delta") {
      const int d = g->dims[cur];
      Node x; x.name = n; x.dim = 3 * d; x.kind = 3; x.block_dim = d;
      x.blocks = {{{cur, 0, 1.0f}}, {{cur, -1, -1.0f}, {cur, 1, 1.0f}}, {{cur, -2, 1.0f}, {cur, 2, 1.0f}, {cur, 0, -2.0f}}};
      if (!bn(&x, n, 1.0f, 1)) return false;
      add(x); cur = n;
    } else if (t == "lda") {
      const int d = g->dims[cur];
      const std::vector<int> sp = splice_of(L, {-1, 0, 1});
      const int k = (int)sp.size() * d + ivd;
      Node x; x.name = n; x.dim = k; x.kind = 2;
      for (size_t i = 0; i < sp.size(); i++) x.terms.push_back({cur, sp[i], (int)i * d, (int)(i + 1) * d, 0});
      x.terms.push_back({"ivector", 0, (int)sp.size() * d, k, 1});
      if (!lin(&x, n, true)) return false;
      add(x); cur = n;
    } else if (t == "relu-batchnorm") {
      const int d0 = g->dims[cur];
      const std::vector<int> sp = splice_of(L, {0});
      const int d = (int)sp.size() * d0;
      Node x; x.name = n + ".batchnorm"; x.dim = L.dim; x.kind = 2; x.relu = true;
      for (size_t i = 0; i < sp.size(); i++) x.terms.push_back({cur, sp[i], (int)i * d0, (int)(i + 1) * d0, 0});
      if (L.append_ivector != 0.0f) {     // Scale(s, ReplaceIndex(ivector, t, 0)) is a node of its own (reference's association)
        Node s; s.name = n + ".ivscaled"; s.dim = ivd; s.kind = 3; s.block_dim = ivd; s.ivector_rows = true;
        s.blocks = {{{"ivector", 0, L.append_ivector}}};
        add(s);
        x.terms.push_back({n + ".ivscaled", 0, d, d + ivd, 1});
      }
      if (!lin(&x, n + ".affine", true) || !bn(&x, n + ".batchnorm", 1.0f, 1)) return false;
      add(x); cur = n + ".batchnorm";
    } else if (t == "tdnnf") {
      const int d = g->dims[cur], s = L.stride, bt = L.bottleneck;
      std::vector<int> o1 = s ? std::vector<int>{-s, 0} : std::vector<int>{0};
      std::vector<int> o2 = s ? std::vector<int>{0, s} : std::vector<int>{0};
      Node a; a.name = n + ".linear"; a.dim = bt; a.kind = 2;
      for (size_t i = 0; i < o1.size(); i++) a.terms.push_back({cur, o1[i], (int)i * d, (int)(i + 1) * d, 0});
      if (!lin(&a, n + ".linear", false)) return false;
      add(a);
      Node x; x.name = n + ".noop"; x.dim = L.dim; x.kind = 2; x.relu = true;
      for (size_t i = 0; i < o2.size(); i++) x.terms.push_back({n + ".linear", o2[i], (int)i * bt, (int)(i + 1) * bt, 0});
      if (!lin(&x, n + ".affine", true) || !bn(&x, n + ".batchnorm", 1.0f, 1)) return false;
      if (L.bypass != 0.0f) { x.has_res = true; x.res_src = cur; x.res_alpha = L.bypass; }
      add(x); cur = n + ".noop";
    } else if (t == "linear") {
      const int d = g->dims[cur];
      Node x; x.name = n; x.dim = L.dim; x.kind = 2; x.terms = {{cur, 0, 0, d, 0}};
      if (!lin(&x, n, false)) return false;
      add(x); cur = n;
    } else if (t == "prefinal") {
      const int d = g->dims[cur];
      Node a; a.name = n + ".batchnorm1"; a.dim = L.big; a.kind = 2; a.relu = true; a.terms = {{cur, 0, 0, d, 0}};
      if (!lin(&a, n + ".affine", true) || !bn(&a, n + ".batchnorm1", 1.0f, 1)) return false;
      add(a);
      Node x; x.name = n + ".batchnorm2"; x.dim = L.small; x.kind = 2; x.terms = {{n + ".batchnorm1", 0, 0, L.big, 0}};
      if (!lin(&x, n + ".linear", false) || !bn(&x, n + ".batchnorm2", 1.0f, 1)) return false;
      add(x); cur = n + ".batchnorm2";
    } else if (t == "ivector-linear-bn") {   // once per nnet chunk (the rows of the ivector input)
      Node x; x.name = n + "-batchnorm"; x.dim = L.dim; x.kind = 2; x.ivector_rows = true;
      x.terms = {{"ivector", 0, 0, ivd, 2}};
      if (!lin(&x, n + "-linear", false) || !bn(&x, n + "-batchnorm", L.target_rms, 1)) return false;
      add(x);
    } else if (t == "combine") {
      pending_combine = &L;
    } else if (t == "conv") {
      const bool patch = !cfg.conv_dense && !pending_combine && !structural;
      Node x; x.name = n + ".batchnorm"; x.dim = L.height_out * L.filters_out; x.kind = 2; x.relu = true;
      if (patch) {
        const int Fi = L.filters_in;
        for (int ti = 0; ti < L.n_time_offsets; ti++)
          for (int hi = 0; hi < L.n_height_offsets; hi++) {
            const int k0 = (ti * L.n_height_offsets + hi) * Fi;
            x.terms.push_back({cur, L.time_offsets[ti], k0, k0 + Fi, 0});
            x.term_cols.push_back({L.height_subsample_out * Fi, L.height_offsets[hi] * Fi, L.height_in * Fi});
          }
        x.hsplit = L.height_out;
        if (!lin(&x, n + ".conv", true) || !bn(&x, n + ".batchnorm", 1.0f, 1)) return false;
        add(x); cur = n + ".batchnorm";
        continue;
      }
      int k_main, k_side;
      if (structural) {
        const int f2 = pending_combine ? pending_combine->filters2 : 0;
        k_side = L.height_in * f2; k_main = L.height_in * L.filters_in - k_side;
      } else {
        const b2k_nnet_weight *w = W.get(n + ".conv.w"), *b = W.get(n + ".conv.b");
        NEED(w, n + ".conv.w"); NEED(b, n + ".conv.b");
        if (w->size != (long long)L.filters_out * L.n_time_offsets * L.n_height_offsets * L.filters_in || b->size != L.filters_out) {
          err = "convolution parameters of " + n + " do not match the layer's filter and offset counts"; return false;
        }
        if (pending_combine && (pending_combine->filters1 + pending_combine->filters2 != L.filters_in || pending_combine->height != L.height_in)) {
          err = "combine-feature-maps in front of " + n + " does not match its input maps"; return false;
        }
        expand_conv(L, w->data, b->data, pending_combine, &x, &k_main, &k_side);
        if (!bn(&x, n + ".batchnorm", 1.0f, L.height_out)) return false;
      }
      x.has_bn = true;
      const int K = k_main + k_side;
      for (int ti = 0; ti < L.n_time_offsets; ti++) {
        x.terms.push_back({cur, L.time_offsets[ti], ti * K, ti * K + k_main, 0});
        if (k_side) x.terms.push_back({S(pending_combine->side, sizeof(pending_combine->side)), L.time_offsets[ti], ti * K + k_main, (ti + 1) * K, 1});
      }
      add(x);
      pending_combine = nullptr;
      cur = n + ".batchnorm";
    } else if (t == "output") {
      const int d = g->dims[cur];
      Node x; x.name = "output"; x.dim = L.dim; x.kind = 2; x.terms = {{cur, 0, 0, d, 0}}; x.log_softmax = L.log_softmax != 0;
      if (!lin(&x, n + ".affine", true)) return false;
      add(x); cur = "output";
    } else {
      err = "unknown layer type " + t;
      return false;
    }
  }
  return true;
}

// The layer list and the weights come from the caller (or from a file): everything the later stages index with is
// checked here, so that an inconsistent model is B2K_ERR_INVALID instead of an out-of-bounds read.
static bool validate_layers(const b2k_nnet_compile_cfg &cfg, const b2k_nnet_layer *layers, int n_layers, std::string &err) {
  const int kMaxDim = 1 << 20;
  if (cfg.feat_dim <= 0 || cfg.feat_dim > kMaxDim || cfg.ivector_dim < 0 || cfg.ivector_dim > kMaxDim || cfg.num_pdfs <= 0 || cfg.num_pdfs > kMaxDim ||
      cfg.num_frames > (1 << 24) || n_layers > 4096) { err = "dimensions out of range"; return false; }
  for (int i = 0; i < n_layers; i++) {
    const b2k_nnet_layer &L = layers[i];
    const std::string t = S(L.type, sizeof(L.type)), n = S(L.name, sizeof(L.name));
    auto dim_ok = [&](int d) { return d > 0 && d <= kMaxDim; };
    bool ok = true;
    if (t == "linear" || t == "output" || t == "ivector-linear-bn") ok = dim_ok(L.dim);
    else if (t == "relu-batchnorm" || t == "lda") {          // time_offsets: the splice of the layer's input (0 = the recipe default)
      ok = (t == "lda" || dim_ok(L.dim)) && L.n_time_offsets >= 0 && L.n_time_offsets <= 8;
      for (int k = 0; ok && k < L.n_time_offsets; k++) ok = L.time_offsets[k] >= -64 && L.time_offsets[k] <= 64;
    }
    else if (t == "tdnnf") ok = dim_ok(L.dim) && dim_ok(L.bottleneck) && L.stride >= 0 && L.stride <= 64;
    else if (t == "prefinal") ok = dim_ok(L.big) && dim_ok(L.small);
    else if (t == "combine") ok = L.height > 0 && L.filters1 > 0 && L.filters2 >= 0 && (long long)L.height * (L.filters1 + (long long)L.filters2) <= kMaxDim;
    else if (t == "conv") {
      ok = L.height_in > 0 && L.height_out > 0 && L.height_subsample_out > 0 && L.filters_in > 0 && L.filters_out > 0 &&
           L.n_time_offsets > 0 && L.n_time_offsets <= 8 && L.n_height_offsets > 0 && L.n_height_offsets <= 8 &&
           (long long)L.height_in * L.filters_in <= kMaxDim && (long long)L.height_out * L.filters_out <= kMaxDim;
      for (int k = 0; ok && k < L.n_time_offsets; k++) ok = L.time_offsets[k] >= -64 && L.time_offsets[k] <= 64;
      for (int k = 0; ok && k < L.n_height_offsets; k++) ok = L.height_offsets[k] >= -64 && L.height_offsets[k] <= 64;
    }
    if (!ok) { err = "layer " + n + " (" + t + ") has dimensions out of range"; return false; }
  }
  return true;
}

static bool validate_graph(Graph &g, std::string &err) {
  std::map<std::string, const Node *> by;
  for (auto &n : g.nodes) {
    if (by.count(n.name)) { err = "two nodes are called " + n.name; return false; }
    by[n.name] = &n;
  }
  auto src_dim = [&](const std::string &s) -> int { auto it = by.find(s); return it == by.end() ? -1 : it->second->dim; };
  for (auto &n : g.nodes) {
    if (n.kind == 2) {
      const int H = std::max(1, n.hsplit);
      if (!n.w || n.dim <= 0 || n.dim % H != 0 || n.w_rows != n.dim / H || n.w_cols <= 0) { err = "parameters of " + n.name + " do not match its output dimension"; return false; }
      int K = 0;
      for (size_t i = 0; i < n.terms.size(); i++) {
        const Term &t = n.terms[i];
        const int sd = src_dim(t.src);
        if (sd < 0) { err = n.name + " reads the unknown node " + t.src; return false; }
        const int len = t.c1 - t.c0;
        if (t.c0 < 0 || len <= 0) { err = "bad column range in " + n.name; return false; }
        if (i < n.term_cols.size()) {                      // convolution patch: a window of the source row
          const ColWin &c = n.term_cols[i];
          if (c.lim <= 0 || c.lim > sd || c.step <= 0) { err = "bad convolution window in " + n.name; return false; }
          // every in-range window start leaves len columns inside [0, lim)
          for (int h = 0; h < H; h++) { const int cb = h * c.step + c.off; if (cb >= 0 && cb < c.lim && cb + len > sd) { err = "convolution window of " + n.name + " leaves its source row"; return false; } }
        } else if (len > sd) { err = n.name + " reads more columns than " + t.src + " has"; return false; }
        K = std::max(K, t.c1);
      }
      if (K != n.w_cols) { err = "parameters of " + n.name + " do not match its input dimension"; return false; }
      if (n.has_b && n.b_size != n.w_rows) { err = "bias of " + n.name + " has the wrong length"; return false; }
      if (n.has_bn && ((int)n.bn_scale.size() != n.w_rows || (int)n.bn_offset.size() != n.w_rows)) { err = "batchnorm statistics of " + n.name + " have the wrong length"; return false; }
      if (n.has_res && src_dim(n.res_src) != n.dim) { err = "bypass input of " + n.name + " has a different dimension"; return false; }
    } else if (n.kind == 3) {
      if (n.block_dim <= 0 || (long long)n.block_dim * (long long)n.blocks.size() != n.dim) { err = "block layout of " + n.name + " does not match its dimension"; return false; }
      for (auto &blk : n.blocks) for (auto &t : blk) if (src_dim(t.src) < n.block_dim) { err = n.name + " reads more columns than " + t.src + " has"; return false; }
      if (n.has_bn && ((int)n.bn_scale.size() != n.dim || (int)n.bn_offset.size() != n.dim)) { err = "batchnorm statistics of " + n.name + " have the wrong length"; return false; }
    }
  }
  return true;
}

static void deps(const Node &n, std::vector<std::pair<std::string, int>> *d) {
  d->clear();
  if (n.kind == 2) {
    if (n.ivector_rows) return;
    for (const Term &t : n.terms) if (t.kind == 0) d->push_back({t.src, t.off});
    if (n.has_res) d->push_back({n.res_src, 0});
  } else if (n.kind == 3) {
    if (n.ivector_rows) return;
    for (const auto &blk : n.blocks) for (const EwTerm &t : blk) d->push_back({t.src, t.off});
  }
}

// ComputeSimpleNnetContext (nnet3/nnet-utils.cc) for this family
static bool model_context(const b2k_nnet_compile_cfg &cfg, const b2k_nnet_layer *layers, int n_layers, int *left, int *right, std::string &err) {
  Graph g;
  Weights none;
  if (!build_graph(cfg, layers, n_layers, none, true, &g, err)) return false;
  std::map<std::string, std::pair<bool, std::pair<int, int>>> need;
  for (auto &n : g.nodes) need[n.name] = {false, {0, 0}};
  if (!need.count("output")) { err = "no output layer"; return false; }
  need["output"] = {true, {0, 0}};
  std::vector<std::pair<std::string, int>> d;
  for (int i = (int)g.nodes.size() - 1; i >= 0; i--) {
    const Node &n = g.nodes[i];
    auto &me = need[n.name];
    if (!me.first) continue;
    deps(n, &d);
    for (auto &so : d) {
      auto &s = need[so.first];
      const int a = me.second.first + so.second, b = me.second.second + so.second;
      if (!s.first) s = {true, {a, b}};
      else { s.second.first = std::min(s.second.first, a); s.second.second = std::max(s.second.second, b); }
    }
  }
  *left = -need["input"].second.first;
  *right = need["input"].second.second;
  return true;
}

}  // namespace

struct b2k_nnet_program {
  std::vector<b2k_nnet_node> nodes;
  std::vector<b2k_nnet_op> ops;
  std::vector<float> blob;
  int32_t n_out = 0, n_chunks = 0, left = 0, right = 0, model_left = 0, model_right = 0, ivector_m = 0;
  int64_t arena_size = 0;
};

extern "C" {

// win_n_out <= 0: the whole utterance (outputs at t = 0, sub, ...; the input is padded by repeating its first / last frame).
// win_n_out > 0: a window program -- outputs at t = win_t0 + k*sub, k < win_n_out, of a window of cfg.num_frames input frames that
// must hold every frame those outputs read (no padding), one i-vector for the whole window: the computation request of
// BatchedStaticNnet3::SetComputationRequest (cudadecoder/batched-static-nnet3.cc:123-152).
// win_iv_rows > 1 (window programs only): the looped computation's i-vectors -- the window is chunk n of a looped run
// (its first output is the first frame of the chunk, cfg.frames_per_chunk = C frames per chunk) and the i-vector input holds
// the i-vectors of chunks n-(rows-1) .. n: time t of the window reads row floor((t - win_t0 + (rows-1)*C) / C) - m, the same
// Round() / lag arithmetic as the whole-utterance program (nnet-compile-looped.cc:179-205).
static int compile_impl(const b2k_nnet_compile_cfg *cfgp, int win_t0, int win_n_out, int win_iv_rows, const b2k_nnet_layer *layers, int32_t n_layers,
                        const b2k_nnet_weight *weights, int32_t n_weights, b2k_nnet_program **out) {
  if (!cfgp || !layers || n_layers <= 0 || (!weights && n_weights > 0) || !out)
    return set_error(B2K_ERR_INVALID, "b2k_nnet_compile: bad args");
  const b2k_nnet_compile_cfg &cfg = *cfgp;
  const bool window = win_n_out > 0;
  const int sub = cfg.frame_subsampling_factor, T = cfg.num_frames, C = cfg.frames_per_chunk;
  if (sub <= 0 || T <= 0 || C <= 0 || C % sub != 0) return set_error(B2K_ERR_INVALID, "b2k_nnet_compile: frames_per_chunk must be a positive multiple of the subsampling factor");
  if (window && (win_t0 < 0 || (long long)win_t0 + (long long)sub * (win_n_out - 1) >= T))
    return set_error(B2K_ERR_INVALID, "b2k_nnet_compile_window: the outputs lie outside the window");
  Weights W;
  for (int i = 0; i < n_weights; i++) {
    const b2k_nnet_weight &w = weights[i];
    if (!w.name || !w.data || w.size < 0 || w.rows < 0 || w.cols < 0 || (long long)w.rows * std::max(w.cols, 1) != w.size)
      return set_error(B2K_ERR_INVALID, "b2k_nnet_compile: a weight's size does not match its shape", w.name ? w.name : "(unnamed)");
    W.m[w.name] = &w;
  }
  std::string err;
  Graph g;
  if (!validate_layers(cfg, layers, n_layers, err)) return set_error(B2K_ERR_INVALID, "b2k_nnet_compile", err.c_str());
  if (!build_graph(cfg, layers, n_layers, W, false, &g, err)) return set_error(B2K_ERR_INVALID, "b2k_nnet_compile", err.c_str());
  if (!validate_graph(g, err)) return set_error(B2K_ERR_INVALID, "b2k_nnet_compile", err.c_str());
  std::map<std::string, Node *> by;
  for (auto &n : g.nodes) by[n.name] = &n;
  if (!by.count("output")) return set_error(B2K_ERR_INVALID, "b2k_nnet_compile: no output layer");
  const int n_out = window ? win_n_out : (T + sub - 1) / sub;
  const int out_t0 = window ? win_t0 : 0;
  Node *o = by["output"];
  o->residues = {out_t0 % sub}; o->tmin = out_t0; o->tmax = out_t0 + sub * (n_out - 1);
  std::vector<std::pair<std::string, int>> d;
  // backward pass: required time range and residues (mod sub) of every node
  for (int i = (int)g.nodes.size() - 1; i >= 0; i--) {
    Node &n = g.nodes[i];
    if (n.tmax < n.tmin) continue;
    deps(n, &d);
    for (auto &so : d) {
      Node *s = by[so.first];
      s->tmin = std::min(s->tmin, n.tmin + so.second);
      s->tmax = std::max(s->tmax, n.tmax + so.second);
      if (n.residues.size() == 1 && n.kind != 0) {
        const int r = *n.residues.begin();
        s->residues.insert(((r + so.second) % sub + sub) % sub);
      } else {
        for (int r = 0; r < sub; r++) s->residues.insert(r);
      }
    }
  }
  b2k_nnet_program *P = new b2k_nnet_program();
  P->n_out = n_out;
  P->left = out_t0 - by["input"]->tmin;
  P->right = by["input"]->tmax - (out_t0 + sub * (n_out - 1));
  if (window && (by["input"]->tmin < 0 || by["input"]->tmax > T - 1)) {
    char msg[160];
    snprintf(msg, sizeof(msg), "the outputs read input frames [%d, %d] of a window of %d", by["input"]->tmin, by["input"]->tmax, T);
    delete P;
    return set_error(B2K_ERR_INVALID, "b2k_nnet_compile_window", msg);
  }
  for (auto &n : g.nodes) {
    if (n.kind == 0) { n.step = 1; n.t0 = 0; n.rows = T; }
    else if (n.kind == 1) { n.step = 0; n.t0 = 0; n.rows = 0; }
    else if (n.tmax < n.tmin) n.rows = 0;
    else if (n.residues.size() == 1) {
      const int r = *n.residues.begin();
      n.step = sub;
      n.t0 = n.tmin + (((r - n.tmin) % sub) + sub) % sub;
      n.rows = (n.tmax - n.t0) / sub + 1;
    } else { n.step = 1; n.t0 = n.tmin; n.rows = n.tmax - n.tmin + 1; }
  }
  int Lk = 0, Rk = 0;
  if (!model_context(cfg, layers, n_layers, &Lk, &Rk, err)) { delete P; return set_error(B2K_ERR_INVALID, "b2k_nnet_compile", err.c_str()); }
  // chunk n of the looped computation supplies one i-vector; input time t uses chunk max(0, floor(t / C) - m),
  // m = floor((C + R - 1) / C)   (nnet-compile-looped.cc:179-205)
  const int m = (C + Rk - 1) / C;
  const int n_chunks = window ? std::max(1, win_iv_rows) : (n_out * sub + C - 1) / C;
  const int iv_shift = (window && win_iv_rows > 1) ? (win_iv_rows - 1) * C - out_t0 : 0;
  by["ivector"]->rows = n_chunks;
  P->model_left = Lk; P->model_right = Rk; P->ivector_m = m; P->n_chunks = n_chunks;
  auto put = [&](const float *a, size_t n) -> int64_t {
    const int64_t off = (int64_t)P->blob.size();
    P->blob.insert(P->blob.end(), a, a + n);
    return off;
  };
  std::map<std::string, int> idx;
  for (size_t i = 0; i < g.nodes.size(); i++) idx[g.nodes[i].name] = (int)i;
  struct OpSrc { std::vector<int> srcs; };
  std::vector<OpSrc> op_srcs;
  bool failed = false;
  for (auto &n : g.nodes) {
    if (n.kind == 0 || n.kind == 1) continue;
    if (n.rows == 0 && !n.ivector_rows) continue;
    b2k_nnet_op op;
    memset(&op, 0, sizeof(op));
    op.out = idx[n.name];
    op.w = op.bias = op.sub_vec = op.bn_scale = op.bn_offset = -1;
    op.out_scale = 1.0f; op.block_dim = 1; op.hsplit = 0;
    OpSrc os;
    auto rowmap = [&](const std::string &src, int off, int kind, b2k_nnet_term *t) {
      const Node *s = by[src];
      memset(t, 0, sizeof(*t));
      t->src = idx[src]; t->C = 1; t->scale = 1.0f;
      if (kind == 1) { t->ratio = n.step; t->shift = n.t0 + off + iv_shift; t->lo = 0; t->hi = s->rows - 1; t->ivec = 1; t->C = C; t->m = m; return; }
      if (s->kind == 0) { t->ratio = n.step; t->shift = n.t0 + off; t->lo = 0; t->hi = T - 1; return; }
      if (s->step == 0 || n.step % s->step != 0 || (n.t0 + off - s->t0) % s->step != 0) { failed = true; return; }
      t->ratio = n.step / s->step; t->shift = (n.t0 + off - s->t0) / s->step; t->lo = 0; t->hi = s->rows - 1;
    };
    if (n.ivector_rows) {
      n.step = 0; n.t0 = 0; n.rows = n_chunks;
      op.rows = n_chunks;
      if (n.kind == 3) {
        op.type = 1; op.block_dim = n.block_dim;
        int j = 0;
        for (size_t bi = 0; bi < n.blocks.size(); bi++)
          for (const EwTerm &e : n.blocks[bi]) {
            b2k_nnet_term &t = op.terms[j++];
            memset(&t, 0, sizeof(t));
            t.src = idx["ivector"]; t.ratio = 1; t.shift = 0; t.lo = 0; t.hi = n_chunks - 1; t.C = 1; t.scale = e.scale; t.block = (int)bi;
            os.srcs.push_back(t.src);
          }
        op.n_terms = j;
      } else {
        op.type = 0; op.N = n.w_rows; op.K = n.w_cols;
        int j = 0;
        for (const Term &tt : n.terms) {
          b2k_nnet_term &t = op.terms[j++];
          memset(&t, 0, sizeof(t));
          t.src = idx["ivector"]; t.ratio = 1; t.lo = 0; t.hi = n_chunks - 1; t.C = 1; t.scale = 1.0f; t.k0 = tt.c0; t.klen = tt.c1 - tt.c0;
          os.srcs.push_back(t.src);
        }
        op.n_terms = j;
        op.w = put(n.w, (size_t)n.w_rows * n.w_cols);
        if (n.has_b) op.bias = put(n.b, n.w_rows);
        op.relu = n.relu;
        if (n.has_bn) { op.bn_scale = put(n.bn_scale.data(), n.bn_scale.size()); op.bn_offset = put(n.bn_offset.data(), n.bn_offset.size()); }
      }
      P->ops.push_back(op); op_srcs.push_back(os);
      continue;
    }
    op.rows = n.rows;
    if (n.kind == 3) {
      op.type = 1; op.block_dim = n.block_dim;
      int j = 0;
      for (size_t bi = 0; bi < n.blocks.size(); bi++)
        for (const EwTerm &e : n.blocks[bi]) {
          if (j >= 12) { failed = true; break; }
          b2k_nnet_term &t = op.terms[j++];
          rowmap(e.src, e.off, 0, &t);
          t.scale = e.scale; t.block = (int)bi;
          os.srcs.push_back(t.src);
        }
      op.n_terms = j;
      if (n.has_bn) { op.bn_scale = put(n.bn_scale.data(), n.bn_scale.size()); op.bn_offset = put(n.bn_offset.data(), n.bn_offset.size()); }
    } else {
      op.type = 0; op.N = n.w_rows; op.K = n.w_cols;
      if (n.terms.size() > 12) failed = true;
      int j = 0;
      for (size_t ti = 0; ti < n.terms.size() && ti < 12; ti++) {
        const Term &tt = n.terms[ti];
        b2k_nnet_term &t = op.terms[j++];
        rowmap(tt.src, tt.off, tt.kind == 1 ? 1 : 0, &t);
        t.k0 = tt.c0; t.klen = tt.c1 - tt.c0;
        if (!n.term_cols.empty()) { t.col_step = n.term_cols[ti].step; t.col_off = n.term_cols[ti].off; t.col_lim = n.term_cols[ti].lim; }
        os.srcs.push_back(t.src);
      }
      op.n_terms = j;
      op.w = put(n.w, (size_t)n.w_rows * n.w_cols);
      if (n.has_b) op.bias = put(n.b, n.w_rows);
      op.relu = n.relu;
      if (n.has_bn) { op.bn_scale = put(n.bn_scale.data(), n.bn_scale.size()); op.bn_offset = put(n.bn_offset.data(), n.bn_offset.size()); }
      op.log_softmax = n.log_softmax;
      op.hsplit = n.hsplit;
      if (n.has_res) { op.has_res = 1; rowmap(n.res_src, 0, 0, &op.res); op.res_alpha = n.res_alpha; os.srcs.push_back(op.res.src); }
      if (n.name == "output") {    // AddVecToRows(-1, log_priors); Scale(acoustic_scale)  (decodable-online-looped.cc:218-223)
        if (cfg.use_priors) {
          const b2k_nnet_weight *pri = W.get("priors");
          if (!pri) { delete P; return set_error(B2K_ERR_INVALID, "b2k_nnet_compile: missing weight priors"); }
          std::vector<float> lp((size_t)pri->size);
          for (int64_t i = 0; i < pri->size; i++) lp[(size_t)i] = logf(pri->data[i]);
          op.sub_vec = put(lp.data(), lp.size());
        }
        op.out_scale = cfg.acoustic_scale;
      }
    }
    P->ops.push_back(op); op_srcs.push_back(os);
  }
  if (failed) { delete P; return set_error(B2K_ERR_INVALID, "b2k_nnet_compile: inconsistent time grids or too many terms"); }
  // per-utterance arena with liveness-based reuse (first fit)
  std::map<int, int> last_use;
  for (size_t oi = 0; oi < P->ops.size(); oi++) for (int s : op_srcs[oi].srcs) last_use[s] = (int)oi;
  std::map<int, int64_t> arena_off;
  struct Live { int64_t off, size; int node; };
  std::vector<Live> live;
  int64_t arena_size = 0;
  for (size_t oi = 0; oi < P->ops.size(); oi++) {
    const int on = P->ops[oi].out;
    const Node &nd = g.nodes[on];
    if (nd.name != "output") {
      const int64_t size = (((int64_t)nd.rows * nd.dim + 31) / 32) * 32;
      std::sort(live.begin(), live.end(), [](const Live &a, const Live &b) { return a.off != b.off ? a.off < b.off : (a.size != b.size ? a.size < b.size : a.node < b.node); });
      int64_t pos = 0;
      for (const Live &l : live) {
        if (pos + size <= l.off) break;
        pos = std::max(pos, l.off + l.size);
      }
      arena_off[on] = pos;
      live.push_back({pos, size, on});
      arena_size = std::max(arena_size, pos + size);
    }
    std::vector<Live> keep;
    for (const Live &l : live) {
      auto it = last_use.find(l.node);
      const int lu = it == last_use.end() ? -1 : it->second;
      if (lu > (int)oi || l.node == on) keep.push_back(l);
    }
    live.swap(keep);
  }
  P->arena_size = arena_size;
  P->nodes.resize(g.nodes.size());
  for (size_t i = 0; i < g.nodes.size(); i++) {
    const Node &n = g.nodes[i];
    b2k_nnet_node &x = P->nodes[i];
    x.dim = n.dim; x.rows = n.rows;
    x.kind = n.name == "input" ? 1 : n.name == "ivector" ? 2 : n.name == "output" ? 3 : 0;
    auto it = arena_off.find((int)i);
    x.arena_off = it == arena_off.end() ? 0 : it->second;
  }
  if (P->blob.empty()) P->blob.push_back(0.0f);
  *out = P;
  return B2K_OK;
}

int b2k_nnet_compile(const b2k_nnet_compile_cfg *cfg, const b2k_nnet_layer *layers, int32_t n_layers,
                     const b2k_nnet_weight *weights, int32_t n_weights, b2k_nnet_program **out) {
  return compile_impl(cfg, 0, 0, 1, layers, n_layers, weights, n_weights, out);
}

int b2k_nnet_compile_window(const b2k_nnet_compile_cfg *cfg, int32_t first_output_t, int32_t num_outputs, int32_t ivector_rows,
                            const b2k_nnet_layer *layers, int32_t n_layers, const b2k_nnet_weight *weights, int32_t n_weights,
                            b2k_nnet_program **out) {
  if (num_outputs <= 0 || ivector_rows <= 0) return set_error(B2K_ERR_INVALID, "b2k_nnet_compile_window: num_outputs and ivector_rows must be positive");
  return compile_impl(cfg, first_output_t, num_outputs, ivector_rows, layers, n_layers, weights, n_weights, out);
}

int b2k_nnet_looped_ivector_rows(const b2k_nnet_compile_cfg *cfg, const b2k_nnet_layer *layers, int32_t n_layers, int32_t *rows) {
  if (!cfg || !rows || cfg->frames_per_chunk <= 0) return set_error(B2K_ERR_INVALID, "b2k_nnet_looped_ivector_rows: bad args");
  int32_t L = 0, R = 0;
  const int rc = b2k_nnet_model_context(cfg, layers, n_layers, &L, &R);
  if (rc) return rc;
  const int C = cfg->frames_per_chunk;
  *rows = (L + C - 1) / C + (C + R - 1) / C + 1;        // chunks reached back by the left context + the lag m + the chunk itself
  return B2K_OK;
}

int b2k_nnet_model_context(const b2k_nnet_compile_cfg *cfg, const b2k_nnet_layer *layers, int32_t n_layers, int32_t *left, int32_t *right) {
  if (!cfg || !layers || n_layers <= 0 || !left || !right) return set_error(B2K_ERR_INVALID, "b2k_nnet_model_context: bad args");
  std::string err;
  if (!validate_layers(*cfg, layers, n_layers, err)) return set_error(B2K_ERR_INVALID, "b2k_nnet_model_context", err.c_str());
  int L = 0, R = 0;
  if (!model_context(*cfg, layers, n_layers, &L, &R, err)) return set_error(B2K_ERR_INVALID, "b2k_nnet_model_context", err.c_str());
  *left = L; *right = R;
  return B2K_OK;
}

int b2k_nnet_program_destroy(b2k_nnet_program *p) { delete p; return B2K_OK; }

int b2k_nnet_program_sizes(const b2k_nnet_program *p, int32_t *n_nodes, int32_t *n_ops, int64_t *blob_len) {
  if (!p || !n_nodes || !n_ops || !blob_len) return set_error(B2K_ERR_INVALID, "b2k_nnet_program_sizes: bad args");
  *n_nodes = (int32_t)p->nodes.size(); *n_ops = (int32_t)p->ops.size(); *blob_len = (int64_t)p->blob.size();
  return B2K_OK;
}
const b2k_nnet_node *b2k_nnet_program_nodes(const b2k_nnet_program *p) { return p ? p->nodes.data() : nullptr; }
const b2k_nnet_op *b2k_nnet_program_ops(const b2k_nnet_program *p) { return p ? p->ops.data() : nullptr; }
const float *b2k_nnet_program_blob(const b2k_nnet_program *p) { return p ? p->blob.data() : nullptr; }

int b2k_nnet_program_info(const b2k_nnet_program *p, int64_t info[8]) {
  if (!p || !info) return set_error(B2K_ERR_INVALID, "b2k_nnet_program_info: bad args");
  info[0] = p->n_out; info[1] = p->n_chunks; info[2] = p->left; info[3] = p->right; info[4] = p->model_left;
  info[5] = p->model_right; info[6] = p->ivector_m; info[7] = p->arena_size;
  return B2K_OK;
}

int b2k_nnet_create_from_program(const b2k_nnet_program *p, int32_t max_batch, b2k_nnet **out) {
  if (!p || !out) return set_error(B2K_ERR_INVALID, "b2k_nnet_create_from_program: bad args");
  return b2k_nnet_create(p->nodes.data(), (int32_t)p->nodes.size(), p->ops.data(), (int32_t)p->ops.size(), p->blob.data(),
                         (int64_t)p->blob.size(), max_batch, out);
}

}  // extern "C"
