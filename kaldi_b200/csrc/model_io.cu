// model_io.cu — host-only readers for the Kaldi model files on the input side of the hot path (no device code,
// no Kaldi/OpenFst): raw nnet3 models as Nnet::Write emits them (nnet3/nnet-nnet.cc:630-656) and final.mdl =
// TransitionModel (hmm/transition-model.cc:394-453, hmm-topology.cc:38-230) + AmNnetSimple
// (nnet3/am-nnet-simple.cc:34-57), binary or text, mapped onto the layer list / named weights that
// b2k_nnet_compile takes.  C++ counterpart of kaldi_b200/kaldi_io.py (read_nnet3_raw, read_final_mdl,
// nnet3_to_arch), which stays as its test oracle; both are pinned to files written by the reference's own
// Write() methods (tests/test_model_io_cpp.py, tests/test_kaldi_io.py).
#include <algorithm>
#include <cctype>
#include <cstdio>
#include <cstring>
#include <map>
#include <set>
#include <stdexcept>
#include <string>
#include <vector>

#include "common.cuh"

namespace {

struct FormatError : std::runtime_error { using std::runtime_error::runtime_error; };

struct Value {
  enum Kind { SCALAR, BOOL, MAT, VEC, IVEC, PAIRS } kind = SCALAR;
  unsigned char raw[8] = {0}; int raw_size = 0; std::string text;     // SCALAR: 4/8 raw bytes (binary) or the text number
  bool b = false;
  std::vector<float> f; int rows = 0, cols = 0;                        // MAT / VEC (doubles are narrowed)
  std::vector<long long> iv;                                           // IVEC, PAIRS (flattened)
  long long as_int() const {
    if (!text.empty()) return atoll(text.c_str());
    if (raw_size == 4) { int32_t v; memcpy(&v, raw, 4); return v; }
    long long v; memcpy(&v, raw, 8); return v;
  }
  double as_float() const {
    if (!text.empty()) return atof(text.c_str());
    if (raw_size == 4) { float v; memcpy(&v, raw, 4); return v; }
    double v; memcpy(&v, raw, 8); return v;
  }
};

struct Reader {
  std::vector<unsigned char> d;
  size_t p = 0;
  bool binary = false;
  Reader(const void *data, size_t len) : d((const unsigned char *)data, (const unsigned char *)data + len) {
    if (d.size() >= 2 && d[0] == 0 && d[1] == 'B') { binary = true; p = 2; }
  }
  explicit Reader(const char *path) {
    FILE *f = fopen(path, "rb");
    if (!f) throw FormatError(std::string("cannot open ") + path);
    fseek(f, 0, SEEK_END);
    long n = ftell(f);
    fseek(f, 0, SEEK_SET);
    d.resize((size_t)n);
    if (n > 0 && fread(d.data(), 1, (size_t)n, f) != (size_t)n) { fclose(f); throw FormatError("short read"); }
    fclose(f);
    if (d.size() >= 2 && d[0] == 0 && d[1] == 'B') { binary = true; p = 2; }
  }
  static bool sp(unsigned char c) { return c == ' ' || c == '\n' || c == '\t' || c == '\r' || c == '\f' || c == '\v'; }
  void ws() { while (p < d.size() && sp(d[p])) p++; }
  void need(size_t n) const { if (p > d.size() || n > d.size() - p) throw FormatError("unexpected end of file"); }
  // a count read from the file: non-negative and small enough for the bytes that are left (checked BEFORE anything is allocated)
  size_t count(long long n, size_t elem_bytes) const {
    if (n < 0 || p > d.size() || (unsigned long long)n > (d.size() - p) / elem_bytes) throw FormatError("implausible element count in the file");
    return (size_t)n;
  }
  std::string token() {                       // ReadToken (io-funcs.cc:154)
    ws();
    size_t e = p;
    while (e < d.size() && !sp(d[e])) e++;
    if (e == p) throw FormatError("empty token");
    std::string t((const char *)&d[p], e - p);
    p = std::min(e + 1, d.size());
    return t;
  }
  void expect(const char *t) { std::string g = token(); if (g != t) throw FormatError(std::string("expected ") + t + ", got " + g); }
  std::string line() {
    if (p >= d.size()) throw FormatError("unexpected end of file");
    size_t e = p;
    while (e < d.size() && d[e] != '\n') e++;
    std::string s((const char *)&d[p], e - p);
    p = std::min(e + 1, d.size());
    return s;
  }
  std::string text_number() {
    ws();
    size_t e = p;
    while (e < d.size() && !sp(d[e])) e++;
    if (e == p) throw FormatError("unexpected end of file (number expected)");
    std::string s((const char *)&d[p], e - p);
    p = e;
    return s;
  }
  long long read_int() {                      // ReadBasicType<integer> (io-funcs-inl.h:34-110)
    if (!binary) return atoll(text_number().c_str());
    need(1);
    int sz = (signed char)d[p++];
    int n = sz < 0 ? -sz : sz;
    need(n);
    long long v = 0;
    if (n == 1) v = sz > 0 ? (long long)(signed char)d[p] : (long long)d[p];
    else if (n == 2) { int16_t x; memcpy(&x, &d[p], 2); v = sz > 0 ? (long long)x : (long long)(uint16_t)x; }
    else if (n == 4) { int32_t x; memcpy(&x, &d[p], 4); v = sz > 0 ? (long long)x : (long long)(uint32_t)x; }
    else if (n == 8) { memcpy(&v, &d[p], 8); }
    else throw FormatError("bad integer size");
    p += n;
    return v;
  }
  double read_float() {
    if (!binary) return atof(text_number().c_str());
    need(1);
    int n = d[p++];
    need(n);
    double v;
    if (n == 4) { float x; memcpy(&x, &d[p], 4); v = x; } else if (n == 8) memcpy(&v, &d[p], 8); else throw FormatError("bad float size");
    p += n;
    return v;
  }
  void text_brackets(std::vector<std::vector<std::string>> *rows) {
    ws();
    if (p >= d.size() || d[p] != '[') throw FormatError("expected [");
    size_t e = p;
    while (e < d.size() && d[e] != ']') e++;
    if (e >= d.size()) throw FormatError("unterminated [");
    rows->clear();
    std::vector<std::string> cur;
    std::string tok;
    for (size_t i = p + 1; i <= e; i++) {
      const unsigned char c = i < e ? d[i] : '\n';
      if (sp(c)) {
        if (!tok.empty()) { cur.push_back(tok); tok.clear(); }
        if (c == '\n' && !cur.empty()) { rows->push_back(cur); cur.clear(); }
      } else tok.push_back((char)c);
    }
    p = e + 1;
  }
  void read_vector(Value *v) {
    v->kind = Value::VEC;
    if (binary) {
      need(3);
      const bool dbl = d[p] == 'D';
      if (!((d[p] == 'F' || dbl) && d[p + 1] == 'V' && d[p + 2] == ' ')) throw FormatError("expected FV/DV");
      p += 3;
      const size_t n = count(read_int(), dbl ? 8 : 4);
      v->f.resize(n);
      if (dbl) { for (size_t i = 0; i < n; i++) { double x; memcpy(&x, &d[p + 8 * i], 8); v->f[i] = (float)x; } p += 8 * n; }
      else { if (n) memcpy(v->f.data(), &d[p], 4 * n); p += 4 * n; }
      v->rows = (int)n; v->cols = 1;
      return;
    }
    std::vector<std::vector<std::string>> rows;
    text_brackets(&rows);
    v->f.clear();
    for (auto &r : rows) for (auto &x : r) v->f.push_back((float)atof(x.c_str()));
    v->rows = (int)v->f.size(); v->cols = 1;
  }
  void read_matrix(Value *v) {
    v->kind = Value::MAT;
    if (binary) {
      need(3);
      if (d[p] == 'C' && d[p + 1] == 'M') throw FormatError("compressed matrices are not supported");
      const bool dbl = d[p] == 'D';
      if (!((d[p] == 'F' || dbl) && d[p + 1] == 'M' && d[p + 2] == ' ')) throw FormatError("expected FM/DM");
      p += 3;
      const long long r = read_int(), c = read_int();
      if (r < 0 || c < 0 || r > 0x7fffffffLL || c > 0x7fffffffLL) throw FormatError("implausible matrix size in the file");
      const size_t n = count(r * c, dbl ? 8 : 4);
      v->f.resize(n);
      if (dbl) { for (size_t i = 0; i < n; i++) { double x; memcpy(&x, &d[p + 8 * i], 8); v->f[i] = (float)x; } p += 8 * n; }
      else { if (n) memcpy(v->f.data(), &d[p], 4 * n); p += 4 * n; }
      v->rows = (int)r; v->cols = (int)c;
      return;
    }
    std::vector<std::vector<std::string>> rows;
    text_brackets(&rows);
    v->rows = (int)rows.size(); v->cols = rows.empty() ? 0 : (int)rows[0].size();
    v->f.clear();
    for (auto &r : rows) { if ((int)r.size() != v->cols) throw FormatError("ragged matrix"); for (auto &x : r) v->f.push_back((float)atof(x.c_str())); }
  }
  // double-precision forms for the i-vector extractor files (values kept as stored; floats are widened)
  void read_dense_d(char kind /* 'M' | 'V' | 'P' */, std::vector<double> *out, int *rows, int *cols) {
    out->clear();
    if (binary) {
      need(3);
      const bool dbl = d[p] == 'D';
      if (!((d[p] == 'F' || dbl) && d[p + 1] == (unsigned char)kind && d[p + 2] == ' ')) {
        if (d[p] == 'C' && d[p + 1] == 'M') throw FormatError("compressed matrices are not supported");
        throw FormatError(std::string("expected a binary F") + kind + "/D" + kind + " object");
      }
      p += 3;
      long long r = read_int(), c = kind == 'M' ? read_int() : 1;
      if (r < 0 || c < 0 || r > 0x7fffffffLL || c > 0x7fffffffLL) throw FormatError("implausible matrix size in the file");
      const size_t n = count(kind == 'P' ? r * (r + 1) / 2 : r * c, dbl ? 8 : 4);
      out->resize(n);
      if (dbl) { if (n) memcpy(out->data(), &d[p], 8 * n); p += 8 * n; }
      else { for (size_t i = 0; i < n; i++) { float x; memcpy(&x, &d[p + 4 * i], 4); (*out)[i] = x; } p += 4 * n; }
      *rows = (int)r; *cols = kind == 'P' ? (int)r : (int)c;
      return;
    }
    std::vector<std::vector<std::string>> rw;
    text_brackets(&rw);
    for (auto &r : rw) for (auto &x : r) out->push_back(atof(x.c_str()));
    if (kind == 'V') { *rows = (int)out->size(); *cols = 1; return; }
    *rows = (int)rw.size();
    if (kind == 'P') {
      for (size_t i = 0; i < rw.size(); i++) if (rw[i].size() != i + 1) throw FormatError("bad packed matrix row");
      *cols = *rows;
    } else {
      *cols = rw.empty() ? 0 : (int)rw[0].size();
      for (auto &r : rw) if ((int)r.size() != *cols) throw FormatError("ragged matrix");
    }
  }
  void read_int_vector(Value *v, bool pairs) {   // ReadIntegerVector / ReadIntegerPairVector (io-funcs-inl.h:113-290)
    v->kind = pairs ? Value::PAIRS : Value::IVEC;
    v->iv.clear();
    if (binary) {
      need(5);
      const int sz = d[p++];
      int32_t n; memcpy(&n, &d[p], 4); p += 4;
      if (sz != 1 && sz != 2 && sz != 4 && sz != 8) throw FormatError("bad integer vector element size");
      const long long cnt = (long long)count((long long)n * (pairs ? 2 : 1), (size_t)sz);
      for (long long i = 0; i < cnt; i++) {
        long long x = 0;
        if (sz == 4) { int32_t y; memcpy(&y, &d[p + 4 * i], 4); x = y; } else if (sz == 8) memcpy(&x, &d[p + 8 * i], 8);
        else if (sz == 2) { int16_t y; memcpy(&y, &d[p + 2 * i], 2); x = y; } else if (sz == 1) x = (signed char)d[p + i];
        else throw FormatError("bad integer vector element size");
        v->iv.push_back(x);
      }
      p += (size_t)cnt * sz;
      return;
    }
    std::vector<std::vector<std::string>> rows;
    text_brackets(&rows);
    for (auto &r : rows) for (auto &x : r) {
      if (pairs) { long long a = 0, b = 0; if (sscanf(x.c_str(), "%lld,%lld", &a, &b) != 2) throw FormatError("bad pair"); v->iv.push_back(a); v->iv.push_back(b); }
      else v->iv.push_back(atoll(x.c_str()));
    }
  }
  // "<A> v <B> v v ... </End>" -> {token: [values]}
  void read_fields(const std::string &end, std::map<std::string, std::vector<Value>> *out) {
    static const std::set<std::string> ivec_tokens = {"<TimeOffsets>", "<RequiredTimeOffsets>", "<ColumnMap>"};
    static const std::set<std::string> pair_tokens = {"<Offsets>"};
    while (true) {
      std::string tok = token();
      if (tok == end) return;
      if (tok.empty() || tok[0] != '<') throw FormatError("expected a token, got " + tok);
      std::vector<Value> &vals = (*out)[tok];
      vals.clear();
      while (true) {
        if (!binary) ws();
        need(1);
        const unsigned char c = d[p];
        if (c == '<') break;
        Value v;
        if (pair_tokens.count(tok)) read_int_vector(&v, true);
        else if (ivec_tokens.count(tok)) read_int_vector(&v, false);
        else if (binary) {
          need(3);
          if ((c == 'F' || c == 'D') && d[p + 1] == 'M' && d[p + 2] == ' ') read_matrix(&v);
          else if ((c == 'F' || c == 'D') && d[p + 1] == 'V' && d[p + 2] == ' ') read_vector(&v);
          else if (c == 'C' && d[p + 1] == 'M') throw FormatError("compressed matrices are not supported");
          else if (c == 'T' || c == 'F') { v.kind = Value::BOOL; v.b = c == 'T'; p++; }
          else if (c == 4 || c == 8) { v.kind = Value::SCALAR; v.raw_size = c; need(1 + c); memcpy(v.raw, &d[p + 1], c); p += 1 + c; }
          else throw FormatError("cannot parse a binary value in " + tok);
        } else if (c == '[') {
          std::vector<std::vector<std::string>> rows;
          text_brackets(&rows);
          v.kind = rows.size() <= 1 ? Value::VEC : Value::MAT;
          v.rows = (int)rows.size(); v.cols = rows.empty() ? 0 : (int)rows[0].size();
          for (auto &r : rows) for (auto &x : r) v.f.push_back((float)atof(x.c_str()));
          if (v.kind == Value::VEC) { v.rows = (int)v.f.size(); v.cols = 1; }
        } else {
          std::string s = text_number();
          if (s == "T" || s == "F") { v.kind = Value::BOOL; v.b = s == "T"; } else { v.kind = Value::SCALAR; v.text = s; }
        }
        vals.push_back(std::move(v));
      }
    }
  }
};

struct Component { std::string type; std::map<std::string, std::vector<Value>> f; };

struct ParsedNnet {
  std::vector<std::string> config;
  std::vector<std::pair<std::string, Component>> comps;      // file order
  const Component *get(const std::string &n) const { for (auto &c : comps) if (c.first == n) return &c.second; return nullptr; }
};

static void read_nnet3(Reader &r, ParsedNnet *out) {
  r.expect("<Nnet3>");
  if (r.p < r.d.size() && r.d[r.p] == '\n') r.p++;
  while (true) {                                              // config-like section up to the first blank line (text in both modes)
    std::string l = r.line();
    size_t a = l.find_first_not_of(" \t\r"), b = l.find_last_not_of(" \t\r");
    if (a == std::string::npos) { if (!out->config.empty()) break; continue; }
    out->config.push_back(l.substr(a, b - a + 1));
  }
  r.expect("<NumComponents>");
  const long long n = r.read_int();
  for (long long i = 0; i < n; i++) {
    r.expect("<ComponentName>");
    std::string name = r.token(), typ = r.token();
    if (typ.size() < 3 || typ.front() != '<' || typ.back() != '>') throw FormatError("bad component type token " + typ);
    Component c;
    c.type = typ.substr(1, typ.size() - 2);
    r.read_fields("</" + typ.substr(1), &c.f);
    out->comps.push_back({name, std::move(c)});
  }
  r.expect("</Nnet3>");
}

// ---- transition model (tid -> pdf)
struct TopoState { int fwd = -1, sl = -1; std::vector<std::pair<int, float>> tr; };
static void read_topology(Reader &r, std::vector<std::vector<TopoState>> *entries, std::map<int, int> *phone2idx) {
  r.expect("<Topology>");
  if (!r.binary) {
    while (true) {
      std::string tok = r.token();
      if (tok == "</Topology>") break;
      if (tok != "<TopologyEntry>") throw FormatError("expected <TopologyEntry>, got " + tok);
      r.expect("<ForPhones>");
      std::vector<int> ph;
      while (true) { std::string t = r.token(); if (t == "</ForPhones>") break; ph.push_back(atoi(t.c_str())); }
      std::vector<TopoState> states;
      while (true) {
        std::string t = r.token();
        if (t == "</TopologyEntry>") break;
        if (t != "<State>") throw FormatError("expected <State>, got " + t);
        if (r.read_int() != (long long)states.size()) throw FormatError("states out of order in topology");
        TopoState st;
        while (true) {
          std::string u = r.token();
          if (u == "</State>") break;
          if (u == "<PdfClass>") st.fwd = st.sl = (int)r.read_int();
          else if (u == "<ForwardPdfClass>") st.fwd = (int)r.read_int();
          else if (u == "<SelfLoopPdfClass>") st.sl = (int)r.read_int();
          else if (u == "<Transition>") { int dst = (int)r.read_int(); st.tr.push_back({dst, (float)r.read_float()}); }
          else if (u == "<Final>") r.read_float();
          else throw FormatError("unexpected token " + u + " in topology state");
        }
        states.push_back(st);
      }
      for (int p : ph) (*phone2idx)[p] = (int)entries->size();
      entries->push_back(states);
    }
    return;
  }
  Value phones, p2i;
  r.read_int_vector(&phones, false);
  r.read_int_vector(&p2i, false);
  long long n = r.read_int();
  bool is_hmm = true;
  if (n == -1) { is_hmm = false; n = r.read_int(); }           // extended format with self-loop pdf classes (:213)
  for (long long e = 0; e < n; e++) {
    const long long ns = r.read_int();
    std::vector<TopoState> states;
    for (long long s = 0; s < ns; s++) {
      TopoState st;
      st.fwd = (int)r.read_int();
      st.sl = is_hmm ? st.fwd : (int)r.read_int();
      const long long nt = r.read_int();
      for (long long k = 0; k < nt; k++) { int dst = (int)r.read_int(); st.tr.push_back({dst, (float)r.read_float()}); }
      states.push_back(st);
    }
    entries->push_back(states);
  }
  r.expect("</Topology>");
  for (size_t p = 0; p < p2i.iv.size(); p++) if (p2i.iv[p] >= 0) (*phone2idx)[(int)p] = (int)p2i.iv[p];
}

// TransitionModel::Read + ComputeDerived (hmm/transition-model.cc:394-420,144-188): tid2pdf[t], t = 1..num_tids, [0] = 0
static void read_transition_model(Reader &r, std::vector<int32_t> *tid2pdf, std::vector<int32_t> *tid2phone) {
  r.expect("<TransitionModel>");
  std::vector<std::vector<TopoState>> entries;
  std::map<int, int> p2i;
  read_topology(r, &entries, &p2i);
  std::string tok = r.token();
  if (tok != "<Triples>" && tok != "<Tuples>") throw FormatError("expected <Triples>/<Tuples>, got " + tok);
  const bool has_sl = tok == "<Tuples>";
  const long long n = r.read_int();
  tid2pdf->assign(1, 0);
  tid2phone->assign(1, 0);
  for (long long i = 0; i < n; i++) {
    const int phone = (int)r.read_int(), hs = (int)r.read_int(), fwd = (int)r.read_int();
    const int sl = has_sl ? (int)r.read_int() : fwd;
    auto it = p2i.find(phone);
    if (it == p2i.end() || it->second < 0 || it->second >= (int)entries.size() || hs < 0 || hs >= (int)entries[it->second].size())
      throw FormatError("transition model: bad tuple");
    for (auto &tr : entries[it->second][hs].tr) {
      tid2pdf->push_back(tr.first == hs ? sl : fwd);
      tid2phone->push_back(phone);                           // TransitionIdToPhone (transition-model.cc:798)
    }
  }
  r.expect(has_sl ? "</Tuples>" : "</Triples>");
  r.expect("<LogProbs>");
  Value lp;
  r.read_vector(&lp);
  r.expect("</LogProbs>");
  r.expect("</TransitionModel>");
  if (lp.f.size() != tid2pdf->size()) throw FormatError("transition model: <LogProbs> does not match the number of transition-ids");
}

}  // namespace

struct b2k_ivec_files {
  int num_gauss = 0, feat_dim = 0, ivector_dim = 0, lda_rows = 0, lda_cols = 0, cmvn_dim = 0;
  double prior_offset = 0.0;
  std::vector<float> lda, gconsts, ubm_weights, means_invvars, inv_vars;
  std::vector<double> sigma_inv_m, U, cmvn;
};

struct b2k_model {
  int32_t feat_dim = 0, ivector_dim = 0, num_pdfs = 0, subsampling = 1;
  int32_t subsampling_ambiguous = 0;          // the layer shapes do not decide the factor (see b2k_model_set_frame_subsampling_factor)
  int32_t has_stride3_tdnnf = 0;
  std::vector<b2k_nnet_layer> layers;
  std::vector<std::string> wnames;
  std::vector<std::vector<float>> wdata;
  std::vector<std::pair<int, int>> wshape;
  std::vector<b2k_nnet_weight> weights;       // views into the three above
  std::vector<int32_t> tid2pdf, tid2phone;
  int32_t has_priors = 0;
};

namespace {

static std::map<std::string, std::string> parse_kv(const std::string &rest) {
  // key=value pairs where values may contain spaces (descriptors): a value runs up to the next " key="
  std::map<std::string, std::string> kv;
  size_t i = 0;
  while (i < rest.size()) {
    while (i < rest.size() && rest[i] == ' ') i++;
    size_t eq = rest.find('=', i);
    if (eq == std::string::npos) break;
    std::string key = rest.substr(i, eq - i);
    size_t j = eq + 1, end = rest.size();
    for (size_t k = j; k < rest.size(); k++) {
      if (rest[k] == ' ') {
        size_t m = k + 1;
        while (m < rest.size() && rest[m] != ' ' && rest[m] != '=' && rest[m] != '(' && rest[m] != ')' && rest[m] != ',') m++;
        if (m < rest.size() && rest[m] == '=' && m > k + 1) { end = k; break; }
      }
    }
    std::string val = rest.substr(j, end - j);
    while (!val.empty() && val.back() == ' ') val.pop_back();
    kv[key] = val;
    i = end;
  }
  return kv;
}

static void set_name(char *dst, size_t cap, const std::string &s) {
  if (s.size() >= cap) throw FormatError("name too long: " + s);
  memset(dst, 0, cap);
  memcpy(dst, s.data(), s.size());
}

static bool ends_with(const std::string &s, const std::string &suf) { return s.size() >= suf.size() && s.compare(s.size() - suf.size(), suf.size(), suf) == 0; }

// first number after "Scale(" in a descriptor, e.g. "Sum(Scale(0.66, tdnn1.batchnorm), x)"
static bool scale_in(const std::string &desc, float *v) {
  size_t p = desc.find("Scale(");
  if (p == std::string::npos) return false;
  *v = (float)atof(desc.c_str() + p + 6);
  return true;
}

struct Arch {
  b2k_model *M;
  const ParsedNnet &P;
  std::map<std::string, int> node_dim;
  void add_w(const std::string &name, const Value &v) {
    M->wnames.push_back(name);
    M->wdata.push_back(v.f);
    M->wshape.push_back(v.kind == Value::MAT ? std::make_pair(v.rows, v.cols) : std::make_pair((int)v.f.size(), 1));
  }
  const Value &field(const Component &c, const char *tok) const {
    auto it = c.f.find(tok);
    if (it == c.f.end() && !strcmp(tok, "<LinearParams>")) it = c.f.find("<Params>");    // LinearComponent
    if (it == c.f.end() || it->second.empty()) throw FormatError(std::string("component lacks ") + tok);
    return it->second[0];
  }
  void bn(const std::string &comp, const std::string &key) {
    const Component *c = P.get(comp);
    if (!c || c->type != "BatchNormComponent") throw FormatError("expected BatchNormComponent " + comp);
    add_w(key + ".mean", field(*c, "<StatsMean>"));
    add_w(key + ".var", field(*c, "<StatsVar>"));
  }
};

// nnet3_to_arch of kaldi_io.py: the node patterns steps/libs/nnet3/xconfig emits for the TDNN-F and CNN-TDNN-F recipes
static void to_arch(const ParsedNnet &P, b2k_model *M) {
  Arch A{M, P, {}};
  std::vector<std::pair<std::string, std::map<std::string, std::string>>> cn;     // component-nodes in file order
  std::map<std::string, std::string> inputs;
  // every node line, then the view a decoder has of a trained recipe model (kaldi_io._inference_view): dropout
  // components are the identity in test mode (SetDropoutTestMode, online2-wav-nnet3-latgen-faster.cc:147), so their
  // nodes are replaced by their inputs; and only what the output node "output" depends on is kept (chain recipes
  // leave the cross-entropy branch, output-xent, in final.mdl)
  struct NodeLine { std::string kind; std::map<std::string, std::string> kv; };
  std::vector<NodeLine> lines;
  for (const std::string &l : P.config) {
    size_t sp = l.find(' ');
    if (sp == std::string::npos) continue;
    const std::string kind = l.substr(0, sp);
    auto kv = parse_kv(l.substr(sp + 1));
    if (kind == "input-node") {
      if (kv["name"] == "input") M->feat_dim = atoi(kv["dim"].c_str());
      if (kv["name"] == "ivector") M->ivector_dim = atoi(kv["dim"].c_str());
    } else if (kind == "component-node" || kind == "dim-range-node" || kind == "output-node") lines.push_back({kind, kv});
  }
  auto is_word_char = [](char ch) { return isalnum((unsigned char)ch) || ch == '_' || ch == '.' || ch == '-'; };
  auto descriptor_nodes = [&](const std::string &d) {
    static const std::set<std::string> words = {"Append", "Offset", "Sum", "Scale", "ReplaceIndex", "Round", "IfDefined", "Failover", "Switch", "Const", "t", "x"};
    std::vector<std::string> out;
    for (size_t i = 0; i < d.size();) {
      if (isalpha((unsigned char)d[i]) || d[i] == '_') {
        size_t j = i;
        while (j < d.size() && is_word_char(d[j])) j++;
        const std::string w = d.substr(i, j - i);
        if (!words.count(w)) out.push_back(w);
        i = j;
      } else if (is_word_char(d[i])) { while (i < d.size() && is_word_char(d[i])) i++; }   // a number
      else i++;
    }
    return out;
  };
  // name -> descriptor text that replaces it; nodes are defined before they are used, so one pass in file order is enough
  std::map<std::string, std::string> alias;
  auto subst = [&](const std::string &d) {
    std::string out;
    for (size_t i = 0; i < d.size();) {
      if (isalpha((unsigned char)d[i]) || d[i] == '_') {
        size_t j = i;
        while (j < d.size() && is_word_char(d[j])) j++;
        const std::string w = d.substr(i, j - i);
        auto it = alias.find(w);
        out += it != alias.end() ? it->second : w;
        i = j;
      } else if (is_word_char(d[i])) { while (i < d.size() && is_word_char(d[i])) out.push_back(d[i++]); }
      else out.push_back(d[i++]);
    }
    return out;
  };
  std::vector<NodeLine> view;
  for (auto &ln : lines) {
    NodeLine v = ln;
    for (const char *key : {"input", "input-node"}) if (v.kv.count(key)) v.kv[key] = subst(v.kv[key]);
    if (v.kind == "component-node") {
      const Component *c = P.get(v.kv["component"]);
      std::string src = v.kv["input"];
      while (!src.empty() && src.front() == ' ') src.erase(0, 1);
      while (!src.empty() && src.back() == ' ') src.pop_back();
      const std::string &nm = v.kv["name"];
      if (c && (c->type == "GeneralDropoutComponent" || c->type == "DropoutComponent" || c->type == "SpecAugmentTimeMaskComponent")) {
        const auto names = descriptor_nodes(src);
        if (names.size() != 1 || names[0] != src) throw FormatError("dropout node " + nm + " has a compound input descriptor: " + src);
        alias[nm] = src;
        continue;
      }
      // no-op-component: a name for a descriptor (input2 = Append(delta, Scale(0.4, ivector)), run_tdnn_1k.sh:181); the NoOp nodes
      // that belong to a layer pattern stay: the delta-layer's "<input>_2" and the tdnnf-layer's "<name>.noop"
      if (c && c->type == "NoOpComponent" && !ends_with(nm, ".noop") && !(ends_with(nm, "_2") && src.find("_copy1") != std::string::npos)) {
        alias[nm] = src;
        continue;
      }
    }
    view.push_back(v);
  }
  std::set<std::string> keep;
  bool have_root = false;
  for (auto &ln : view) if (ln.kind == "output-node" && ln.kv["name"] == "output") have_root = true;
  if (have_root) {
    std::map<std::string, const NodeLine *> by_name;
    for (auto &ln : view) by_name[ln.kv["name"]] = &ln;
    std::vector<std::string> todo(1, "output");
    while (!todo.empty()) {
      const std::string n = todo.back();
      todo.pop_back();
      if (keep.count(n) || !by_name.count(n)) continue;
      keep.insert(n);
      const NodeLine *ln = by_name[n];
      for (const char *key : {"input", "input-node"}) {
        auto it = ln->kv.find(key);
        if (it != ln->kv.end()) for (auto &w : descriptor_nodes(it->second)) todo.push_back(w);
      }
    }
  }
  for (auto &ln : view)
    if (ln.kind == "component-node" && (!have_root || keep.count(ln.kv["name"]))) { inputs[ln.kv["name"]] = ln.kv["input"]; cn.push_back({ln.kv["name"], ln.kv}); }
  // nnet3::CollapseModel (nnet3/nnet-utils.cc:1793-1796, 1891-1894) replaces two components by one called "<first>.<second>" and
  // leaves the node its name: such a network is not the model as trained, and the patterns below would misread it
  for (auto &e : cn) {
    const std::string &node = e.first, &comp = e.second.at("component");
    if (comp != node && comp.size() > node.size() + 1 &&
        (comp.compare(comp.size() - node.size() - 1, std::string::npos, "." + node) == 0 || comp.compare(0, node.size() + 1, node + ".") == 0))
      throw FormatError("component " + comp + " of node " + node + " is a merged one: the network has been through nnet3::CollapseModel; "
                        "b2k folds batch-norm and dropout itself and takes the model as it was trained (build the tools with the "
                        "b2k drop-in headers, which leave CollapseModel out)");
  }
  A.node_dim["input"] = M->feat_dim; A.node_dim["ivector"] = M->ivector_dim;
  auto comp_of = [&](size_t i) -> const Component & {
    const Component *c = P.get(cn[i].second.at("component"));
    if (!c) throw FormatError("missing component " + cn[i].second.at("component"));
    return *c;
  };
  // The time offsets over which an input descriptor splices ONE source node: x -> {0}; Append(Offset(x, -1), x, Offset(x, 1)) ->
  // {-1, 0, 1} (what xconfig writes for input=Append(-1,0,1)); a trailing term on the ivector input is not part of the splice.
  auto trim = [](const std::string &x) -> std::string {
    const size_t a = x.find_first_not_of(" \t");
    return a == std::string::npos ? std::string() : x.substr(a, x.find_last_not_of(" \t") - a + 1);
  };
  // A source other than `expect` (the layer before) is a skip connection: not supported, and an error rather than a wrong network.
  auto splice_offsets = [&](const std::string &desc, const std::string &expect) -> std::vector<int> {
    const std::string d = trim(desc);
    auto skip = [&]() { return FormatError("input " + desc + " is not the layer before (" + expect + "): skip connections are not supported"); };
    if (d.size() < 8 || d.compare(0, 7, "Append(") != 0 || d.back() != ')') {
      if (d != expect) throw skip();
      return {0};
    }
    std::vector<std::string> terms;
    int depth = 0;
    size_t start = 7;
    for (size_t p = 7; p + 1 < d.size(); p++) {
      if (d[p] == '(') depth++;
      else if (d[p] == ')') depth--;
      else if (d[p] == ',' && depth == 0) { terms.push_back(d.substr(start, p - start)); start = p + 1; }
    }
    terms.push_back(d.substr(start, d.size() - 1 - start));
    auto is_name = [](const std::string &w) {
      if (w.empty() || !(std::isalpha((unsigned char)w[0]) || w[0] == '_')) return false;
      for (char ch : w) if (!(std::isalnum((unsigned char)ch) || ch == '_' || ch == '.' || ch == '-')) return false;
      return true;
    };
    std::vector<int> offs;
    std::string src;
    for (size_t k = 0; k < terms.size(); k++) {
      const std::string term = trim(terms[k]);
      bool on_ivector = false;
      for (auto &w : descriptor_nodes(term)) on_ivector = on_ivector || w == "ivector";
      if (on_ivector) {
        if (k + 1 != terms.size()) throw FormatError("unsupported input descriptor " + desc);
        continue;
      }
      std::string node = term;
      long o = 0;
      if (term.compare(0, 7, "Offset(") == 0 && term.back() == ')') {
        const size_t comma = term.find(',');
        if (comma == std::string::npos) throw FormatError("unsupported input descriptor " + desc);
        node = trim(term.substr(7, comma - 7));
        const std::string num = trim(term.substr(comma + 1, term.size() - 2 - comma));
        char *end = nullptr;
        o = strtol(num.c_str(), &end, 10);
        if (num.empty() || !end || *end != 0 || o < -64 || o > 64) throw FormatError("unsupported input descriptor " + desc);
      }
      if (!is_name(node) || (!src.empty() && node != src)) throw FormatError("unsupported input descriptor " + desc);
      src = node;
      offs.push_back((int)o);
    }
    if (offs.empty() || offs.size() > 8) throw FormatError("unsupported input descriptor " + desc);
    if (src != expect) throw skip();
    return offs;
  };
  auto set_splice = [&](b2k_nnet_layer &L, const std::vector<int> &sp, const std::vector<int> &dflt) {
    if (sp == dflt) return;
    L.n_time_offsets = (int32_t)sp.size();
    for (size_t k = 0; k < sp.size(); k++) L.time_offsets[k] = sp[k];
  };
  auto new_layer = [&](const char *type, const std::string &name) -> b2k_nnet_layer & {
    b2k_nnet_layer L;
    memset(&L, 0, sizeof(L));
    set_name(L.type, sizeof(L.type), type);
    set_name(L.name, sizeof(L.name), name);
    L.target_rms = 1.0f;
    M->layers.push_back(L);
    return M->layers.back();
  };
  size_t i = 0;
  while (i < cn.size()) {
    const std::string n = cn[i].first;
    const Component &c = comp_of(i);
    const std::string &t = c.type;
    auto next_type = [&](size_t k) -> std::string { return i + k < cn.size() ? comp_of(i + k).type : std::string(); };
    auto next_name = [&](size_t k) -> std::string { return i + k < cn.size() ? cn[i + k].first : std::string(); };
    if (t == "LinearComponent" && inputs[n].find("ivector") != std::string::npos && next_type(1) == "BatchNormComponent" && ends_with(n, "-linear")) {
      const std::string base = n.substr(0, n.size() - 7), bnn = next_name(1);
      const Value &w = A.field(c, "<LinearParams>");
      b2k_nnet_layer &L = new_layer("ivector-linear-bn", base);
      L.dim = w.rows;
      L.target_rms = (float)A.field(comp_of(i + 1), "<TargetRms>").as_float();
      A.add_w(n + ".w", w);
      A.bn(bnn, bnn);
      A.node_dim[bnn] = w.rows;
      i += 2;
    } else if (t == "PermuteComponent") {
      const std::string &in = inputs[n];
      size_t a = in.find("Append("), comma = in.find(','), close = in.rfind(')');
      if (a != 0 || comma == std::string::npos || close == std::string::npos) throw FormatError("unsupported PermuteComponent input " + in);
      std::string main = in.substr(7, comma - 7), side = in.substr(comma + 1, close - comma - 1);
      while (!side.empty() && side[0] == ' ') side.erase(0, 1);
      if (next_type(1) != "TimeHeightConvolutionComponent") throw FormatError("PermuteComponent is only supported as combine-feature-maps in front of a convolution");
      const int h = (int)A.field(comp_of(i + 1), "<HeightIn>").as_int();
      const int f1 = A.node_dim[main] / h, f2 = A.node_dim[side] / h;
      const Value &cm = A.field(c, "<ColumnMap>");
      std::vector<long long> want;
      for (int hh = 0; hh < h; hh++) { for (int f = 0; f < f1; f++) want.push_back(hh * f1 + f); for (int f = 0; f < f2; f++) want.push_back(h * f1 + hh * f2 + f); }
      if (cm.iv != want) throw FormatError("PermuteComponent column map is not a combine-feature-maps interleave");
      b2k_nnet_layer &L = new_layer("combine", n);
      set_name(L.side, sizeof(L.side), side);
      L.height = h; L.filters1 = f1; L.filters2 = f2;
      A.node_dim[n] = A.node_dim[main] + A.node_dim[side];
      i += 1;
    } else if (t == "TimeHeightConvolutionComponent" && ends_with(n, ".conv")) {
      const std::string base = n.substr(0, n.size() - 5);
      const Value &offs = A.field(c, "<Offsets>");
      std::set<long long> ts, hs;
      for (size_t k = 0; k + 1 < offs.iv.size(); k += 2) { ts.insert(offs.iv[k]); hs.insert(offs.iv[k + 1]); }
      if (ts.size() * hs.size() * 2 != offs.iv.size() || ts.size() > 8 || hs.size() > 8) throw FormatError("convolution offsets are not a full time x height grid");
      const Value &req = A.field(c, "<RequiredTimeOffsets>");
      if (std::set<long long>(req.iv.begin(), req.iv.end()) != ts) throw FormatError("time zero-padding (required-time-offsets) is not supported");
      if (next_name(1) != base + ".relu" || next_name(2) != base + ".batchnorm") throw FormatError("expected conv-relu-batchnorm at " + base);
      b2k_nnet_layer &L = new_layer("conv", base);
      L.height_in = (int)A.field(c, "<HeightIn>").as_int(); L.height_out = (int)A.field(c, "<HeightOut>").as_int();
      L.height_subsample_out = (int)A.field(c, "<HeightSubsampleOut>").as_int();
      L.filters_in = (int)A.field(c, "<NumFiltersIn>").as_int(); L.filters_out = (int)A.field(c, "<NumFiltersOut>").as_int();
      for (long long x : ts) L.time_offsets[L.n_time_offsets++] = (int)x;
      for (long long x : hs) L.height_offsets[L.n_height_offsets++] = (int)x;
      A.add_w(n + ".w", A.field(c, "<LinearParams>"));
      A.add_w(n + ".b", A.field(c, "<BiasParams>"));
      A.bn(base + ".batchnorm", base + ".batchnorm");
      A.node_dim[base + ".batchnorm"] = L.height_out * L.filters_out;
      i += 3;
    } else if (t == "FixedAffineComponent") {
      const bool idct = inputs[n] == "input";
      const Value &w = A.field(c, "<LinearParams>");
      b2k_nnet_layer &L = new_layer(idct ? "idct" : "lda", n);
      if (idct) L.dim = w.rows;
      else set_splice(L, splice_offsets(inputs[n], i ? cn[i - 1].first : std::string("input")), {-1, 0, 1});
      A.add_w(n + ".w", w);
      A.add_w(n + ".b", A.field(c, "<BiasParams>"));
      A.node_dim[n] = w.rows;
      i += 1;
    } else if (t == "BatchNormComponent" && next_type(1) == "NoOpComponent" && ends_with(next_name(1), "_2") &&
               inputs[next_name(1)].find("_copy1") != std::string::npos) {
      new_layer("batchnorm", n);                      // batchnorm-component followed by delta-layer (trivial_layers.py:236-256)
      A.bn(n, n);
      const std::string dn = next_name(2);
      new_layer("delta", dn);
      A.bn(dn, dn);
      i += 3;
    } else if (t == "BatchNormComponent") {
      new_layer("batchnorm", n);
      A.bn(n, n);
      A.node_dim[n] = (int)A.field(c, "<Dim>").as_int();
      i += 1;
    } else if ((t == "NaturalGradientAffineComponent" || t == "AffineComponent") && ends_with(n, ".affine") && next_name(1) == n.substr(0, n.size() - 7) + ".relu") {
      const std::string base = n.substr(0, n.size() - 7);
      bool prefinal = false;
      for (size_t k = 0; k < 8 && i + k < cn.size(); k++) if (cn[i + k].first == base + ".batchnorm1") prefinal = true;
      const Value &w = A.field(c, "<LinearParams>");
      if (prefinal) {                                 // affine relu batchnorm1 linear batchnorm2
        const Component *lc = P.get(base + ".linear");
        if (!lc) throw FormatError("prefinal layer without .linear");
        const Value &wl = A.field(*lc, "<LinearParams>");
        b2k_nnet_layer &L = new_layer("prefinal", base);
        L.big = w.rows; L.small = wl.rows;
        A.add_w(n + ".w", w); A.add_w(n + ".b", A.field(c, "<BiasParams>"));
        A.bn(base + ".batchnorm1", base + ".batchnorm1");
        A.add_w(base + ".linear.w", wl);
        A.bn(base + ".batchnorm2", base + ".batchnorm2");
        i += 5;
      } else {                                        // relu-batchnorm-layer
        b2k_nnet_layer &L = new_layer("relu-batchnorm", base);
        L.dim = w.rows;
        if (inputs[n].find("ivector") != std::string::npos) { float s = 1.0f; scale_in(inputs[n], &s); L.append_ivector = s; }
        set_splice(L, splice_offsets(inputs[n], i ? cn[i - 1].first : std::string("input")), {0});
        A.add_w(n + ".w", w); A.add_w(n + ".b", A.field(c, "<BiasParams>"));
        A.bn(base + ".batchnorm", base + ".batchnorm");
        i += 3;
      }
    } else if (t == "TdnnComponent" && ends_with(n, ".linear")) {
      const std::string base = n.substr(0, n.size() - 7);
      const Value &offs = A.field(c, "<TimeOffsets>");
      long long stride = 0;
      for (long long o : offs.iv) stride = std::max(stride, o < 0 ? -o : o);
      const bool has_noop = inputs.count(base + ".noop") > 0;
      const Component *ac = P.get(base + ".affine");
      if (!ac) throw FormatError("tdnnf layer without .affine");
      const Value &wl = A.field(c, "<LinearParams>"), &wa = A.field(*ac, "<LinearParams>");
      b2k_nnet_layer &L = new_layer("tdnnf", base);
      L.dim = wa.rows; L.bottleneck = wl.rows; L.stride = (int)stride;
      float b = 0.0f;
      if (has_noop) scale_in(inputs[base + ".noop"], &b);
      L.bypass = b;
      A.add_w(n + ".w", wl);
      A.add_w(base + ".affine.w", wa); A.add_w(base + ".affine.b", A.field(*ac, "<BiasParams>"));
      A.bn(base + ".batchnorm", base + ".batchnorm");
      i += has_noop ? 5 : 4;
    } else if (t == "LinearComponent") {
      const Value &w = A.field(c, "<LinearParams>");
      b2k_nnet_layer &L = new_layer("linear", n);
      L.dim = w.rows;
      A.add_w(n + ".w", w);
      i += 1;
    } else if ((t == "NaturalGradientAffineComponent" || t == "AffineComponent") && ends_with(n, ".affine")) {
      const std::string base = n.substr(0, n.size() - 7);      // output-layer (+ log-softmax for xent outputs)
      const Value &w = A.field(c, "<LinearParams>");
      b2k_nnet_layer &L = new_layer("output", base);
      L.dim = w.rows;
      A.add_w(n + ".w", w); A.add_w(n + ".b", A.field(c, "<BiasParams>"));
      i += 1;
      if (i < cn.size() && comp_of(i).type == "LogSoftmaxComponent") { M->layers.back().log_softmax = 1; i += 1; }
      if (M->num_pdfs == 0) M->num_pdfs = w.rows;
    } else {
      throw FormatError("unsupported node pattern at " + n + " (" + t + ")");
    }
  }
  // The factor is NOT stored in the file: the tools take --frame-subsampling-factor (default 1,
  // nnet3/decodable-simple-looped.h:75).  A TDNN-F layer with time-stride 3 only exists in chain recipes (factor 3).
  // Splices at +-3 in relu-batchnorm layers do not decide it: chain TDNNs (egs/wsj/s5/local/chain/tuning/run_tdnn_1f.sh,
  // factor 3) and plain nnet3 TDNNs (egs/aishell/s5/local/nnet3/tuning/run_tdnn_1a.sh, factor 1) both have them: such a
  // model is marked ambiguous and refuses to compile until the caller states the factor.
  M->subsampling = 1;
  for (auto &L : M->layers) {
    if (!strcmp(L.type, "tdnnf") && L.stride == 3) { M->subsampling = 3; M->has_stride3_tdnnf = 1; }
    if (!strcmp(L.type, "relu-batchnorm")) for (int k = 0; k < L.n_time_offsets; k++) if (L.time_offsets[k] == 3 || L.time_offsets[k] == -3) M->subsampling_ambiguous = 1;
  }
  if (M->has_stride3_tdnnf) M->subsampling_ambiguous = 0;
  if (M->subsampling_ambiguous) M->subsampling = 3;     // what the chain recipes use; only consulted after the caller confirms it
}

static void finish(b2k_model *M, const std::vector<float> *priors) {
  // priors: AmNnetSimple::Priors; empty (chain models) = no prior subtraction = all ones
  std::vector<float> pri(priors && !priors->empty() ? *priors : std::vector<float>((size_t)M->num_pdfs, 1.0f));
  M->has_priors = priors && !priors->empty();
  M->wnames.push_back("priors"); M->wdata.push_back(pri); M->wshape.push_back({(int)pri.size(), 1});
  M->weights.resize(M->wnames.size());
  for (size_t i = 0; i < M->wnames.size(); i++) {
    M->weights[i].name = M->wnames[i].c_str();
    M->weights[i].data = M->wdata[i].data();
    M->weights[i].size = (int64_t)M->wdata[i].size();
    M->weights[i].rows = M->wshape[i].first; M->weights[i].cols = M->wshape[i].second;
  }
}

}  // namespace

extern "C" {

// kind 0: raw nnet3 (Nnet::Write); 1: final.mdl (TransitionModel + AmNnetSimple); 2: AmNnetSimple::Write alone
static void read_model(Reader &r, int kind, b2k_model *M) {
  ParsedNnet P;
  std::vector<float> priors;
  if (kind == 1) read_transition_model(r, &M->tid2pdf, &M->tid2phone);
  read_nnet3(r, &P);
  if (kind == 1 || kind == 2) {
    r.expect("<LeftContext>"); r.read_int();
    r.expect("<RightContext>"); r.read_int();
    r.expect("<Priors>");
    Value pv;
    r.read_vector(&pv);
    priors = pv.f;
  }
  to_arch(P, M);
  finish(M, &priors);
}

int b2k_model_read(const char *path, int32_t is_mdl, b2k_model **out) {
  if (!path || !out || is_mdl < 0 || is_mdl > 2) return b2k::set_error(B2K_ERR_INVALID, "b2k_model_read: bad args");
  b2k_model *M = new b2k_model();
  try {
    Reader r(path);
    read_model(r, is_mdl, M);
  } catch (const std::exception &e) {
    delete M;
    return b2k::set_error(B2K_ERR_INVALID, "b2k_model_read", e.what());
  }
  *out = M;
  return B2K_OK;
}

int b2k_model_read_memory(const void *data, int64_t len, int32_t kind, b2k_model **out) {
  if (!data || len <= 0 || !out || kind < 0 || kind > 2) return b2k::set_error(B2K_ERR_INVALID, "b2k_model_read_memory: bad args");
  b2k_model *M = new b2k_model();
  try {
    Reader r(data, (size_t)len);
    read_model(r, kind, M);
  } catch (const std::exception &e) {
    delete M;
    return b2k::set_error(B2K_ERR_INVALID, "b2k_model_read_memory", e.what());
  }
  *out = M;
  return B2K_OK;
}

int b2k_model_destroy(b2k_model *m) { delete m; return B2K_OK; }

// A model that never was a file: the layer list and named weights a caller already holds (synthetic models of bench.py
// and the tests; what b2k_nnet_compile takes).  Everything is copied.  "priors" may be absent (ones are used);
// tid2pdf may be NULL (raw model).
int b2k_model_from_arrays(int32_t feat_dim, int32_t ivector_dim, int32_t num_pdfs, int32_t frame_subsampling_factor,
                          const b2k_nnet_layer *layers, int32_t n_layers, const b2k_nnet_weight *weights, int32_t n_weights,
                          const int32_t *tid2pdf, int32_t n_tids, b2k_model **out) {
  if (!out || !layers || !weights || n_layers <= 0 || n_weights <= 0 || feat_dim <= 0 || num_pdfs <= 0 || ivector_dim < 0 ||
      frame_subsampling_factor <= 0 || n_tids < 0)
    return b2k::set_error(B2K_ERR_INVALID, "b2k_model_from_arrays: bad args");
  b2k_model *M = new b2k_model();
  M->feat_dim = feat_dim; M->ivector_dim = ivector_dim; M->num_pdfs = num_pdfs; M->subsampling = frame_subsampling_factor;
  M->layers.assign(layers, layers + n_layers);
  for (auto &L : M->layers) if (!strcmp(L.type, "tdnnf") && L.stride == 3) M->has_stride3_tdnnf = 1;
  bool have_priors = false;
  for (int i = 0; i < n_weights; i++) {
    const b2k_nnet_weight &w = weights[i];
    if (!w.name || !w.data || w.size <= 0 || (int64_t)w.rows * std::max(1, w.cols) != w.size) { delete M; return b2k::set_error(B2K_ERR_INVALID, "b2k_model_from_arrays: bad weight entry"); }
    M->wnames.push_back(w.name);
    M->wdata.emplace_back(w.data, w.data + w.size);
    M->wshape.push_back({w.rows, w.cols});
    if (!strcmp(w.name, "priors")) have_priors = true;
  }
  if (!have_priors) { M->wnames.push_back("priors"); M->wdata.emplace_back((size_t)num_pdfs, 1.0f); M->wshape.push_back({num_pdfs, 1}); }
  M->has_priors = have_priors;
  M->weights.resize(M->wnames.size());
  for (size_t i = 0; i < M->wnames.size(); i++) {
    M->weights[i].name = M->wnames[i].c_str();
    M->weights[i].data = M->wdata[i].data();
    M->weights[i].size = (int64_t)M->wdata[i].size();
    M->weights[i].rows = M->wshape[i].first; M->weights[i].cols = M->wshape[i].second;
  }
  if (tid2pdf && n_tids > 0) M->tid2pdf.assign(tid2pdf, tid2pdf + n_tids);
  if (M->has_stride3_tdnnf && frame_subsampling_factor != 3) { delete M; return b2k::set_error(B2K_ERR_INVALID, "b2k_model_from_arrays: TDNN-F layers with time-stride 3 need frame_subsampling_factor 3"); }
  *out = M;
  return B2K_OK;
}

int b2k_model_info(const b2k_model *m, int32_t info[8]) {
  if (!m || !info) return b2k::set_error(B2K_ERR_INVALID, "b2k_model_info: bad args");
  info[0] = m->feat_dim; info[1] = m->ivector_dim; info[2] = m->num_pdfs; info[3] = m->subsampling;
  info[4] = (int32_t)m->layers.size(); info[5] = (int32_t)m->weights.size(); info[6] = (int32_t)m->tid2pdf.size(); info[7] = m->has_priors;
  return B2K_OK;
}
int32_t b2k_model_frame_subsampling_ambiguous(const b2k_model *m) { return m ? m->subsampling_ambiguous : -1; }
int b2k_model_set_frame_subsampling_factor(b2k_model *m, int32_t factor) {
  if (!m || factor <= 0) return b2k::set_error(B2K_ERR_INVALID, "b2k_model_set_frame_subsampling_factor: bad args");
  if (m->has_stride3_tdnnf && factor != 3)
    return b2k::set_error(B2K_ERR_INVALID, "--frame-subsampling-factor disagrees with the model: its TDNN-F layers have time-stride 3 (a chain model, factor 3)");
  m->subsampling = factor;
  m->subsampling_ambiguous = 0;
  return B2K_OK;
}
const b2k_nnet_layer *b2k_model_layers(const b2k_model *m) { return m ? m->layers.data() : nullptr; }
const b2k_nnet_weight *b2k_model_weights(const b2k_model *m) { return m ? m->weights.data() : nullptr; }
const int32_t *b2k_model_tid2pdf(const b2k_model *m) { return m && !m->tid2pdf.empty() ? m->tid2pdf.data() : nullptr; }
const int32_t *b2k_model_tid2phone(const b2k_model *m) { return m && !m->tid2phone.empty() ? m->tid2phone.data() : nullptr; }


// ------------------------------------------------------------------ i-vector extractor directory
// final.ie = IvectorExtractor::Read (ivector/ivector-extractor.cc:828-848), final.dubm = DiagGmm::Read
// (gmm/diag-gmm.cc:758-800), final.mat = the LDA/splice transform (a Matrix), global_cmvn.stats (a 2 x (dim+1) double
// Matrix).  Derived quantities as IvectorExtractor::ComputeDerivedVars (ivector-extractor.cc:182-230):
// Sigma_inv_M_[g] = Sigma_inv_[g] M_[g],  U_[g] = packed lower triangle of M_[g]^T Sigma_inv_[g] M_[g].

int b2k_ivec_files_read(const char *ie_path, const char *dubm_path, const char *lda_mat_path, const char *global_cmvn_path,
                        b2k_ivec_files **out) {
  if (!ie_path || !dubm_path || !lda_mat_path || !global_cmvn_path || !out)
    return b2k::set_error(B2K_ERR_INVALID, "b2k_ivec_files_read: bad args");
  b2k_ivec_files *F = new b2k_ivec_files();
  try {
    int r = 0, c = 0;
    {   // final.ie
      Reader rd(ie_path);
      rd.expect("<IvectorExtractor>");
      rd.expect("<w>");
      std::vector<double> w; rd.read_dense_d('M', &w, &r, &c);
      rd.expect("<w_vec>");
      rd.read_dense_d('V', &w, &r, &c);
      rd.expect("<M>");
      const long long G = rd.read_int();
      if (G <= 0 || (unsigned long long)G > rd.d.size()) throw FormatError("final.ie: implausible number of Gaussians");
      std::vector<std::vector<double>> M((size_t)G), S((size_t)G);
      int Fd = 0, D = 0;
      for (long long g = 0; g < G; g++) {
        rd.read_dense_d('M', &M[g], &r, &c);
        if (g == 0) { Fd = r; D = c; } else if (r != Fd || c != D) throw FormatError("final.ie: M_ sizes differ");
      }
      rd.expect("<SigmaInv>");
      for (long long g = 0; g < G; g++) { rd.read_dense_d('P', &S[g], &r, &c); if (r != Fd) throw FormatError("final.ie: SigmaInv size"); }
      rd.expect("<IvectorOffset>");
      F->prior_offset = rd.read_float();
      rd.expect("</IvectorExtractor>");
      F->num_gauss = (int)G; F->feat_dim = Fd; F->ivector_dim = D;
      F->sigma_inv_m.assign((size_t)G * Fd * D, 0.0);
      const int DP = D * (D + 1) / 2;
      F->U.assign((size_t)G * DP, 0.0);
      std::vector<double> full((size_t)Fd * Fd);
      for (long long g = 0; g < G; g++) {
        for (int i = 0, k = 0; i < Fd; i++) for (int j = 0; j <= i; j++, k++) full[(size_t)i * Fd + j] = full[(size_t)j * Fd + i] = S[g][k];
        double *sm = &F->sigma_inv_m[(size_t)g * Fd * D];
        for (int i = 0; i < Fd; i++)
          for (int j = 0; j < Fd; j++) {
            const double a = full[(size_t)i * Fd + j];
            const double *mr = &M[g][(size_t)j * D];
            for (int k = 0; k < D; k++) sm[(size_t)i * D + k] += a * mr[k];
          }
        double *u = &F->U[(size_t)g * DP];
        for (int i = 0, k = 0; i < D; i++)
          for (int j = 0; j <= i; j++, k++) {
            double acc = 0.0;
            for (int f = 0; f < Fd; f++) acc += M[g][(size_t)f * D + i] * sm[(size_t)f * D + j];
            u[k] = acc;
          }
      }
    }
    {   // final.dubm
      Reader rd(dubm_path);
      std::string tok = rd.token();
      if (tok != "<DiagGMM>" && tok != "<DiagGMMBegin>") throw FormatError("final.dubm: not a DiagGmm: " + tok);
      bool have[4] = {false, false, false, false};
      while (true) {
        tok = rd.token();
        if (tok == "</DiagGMM>" || tok == "<DiagGMMEnd>") break;
        Value v;
        if (tok == "<GCONSTS>") { rd.read_vector(&v); F->gconsts = v.f; have[0] = true; }
        else if (tok == "<WEIGHTS>") { rd.read_vector(&v); F->ubm_weights = v.f; have[1] = true; }
        else if (tok == "<MEANS_INVVARS>") { rd.read_matrix(&v); F->means_invvars = v.f; have[2] = true; if (v.rows != F->num_gauss || v.cols != F->feat_dim) throw FormatError("final.dubm and final.ie disagree on the number of Gaussians / feature dimension"); }
        else if (tok == "<INV_VARS>") { rd.read_matrix(&v); F->inv_vars = v.f; have[3] = true; if (v.rows != F->num_gauss || v.cols != F->feat_dim) throw FormatError("final.dubm and final.ie disagree on the number of Gaussians / feature dimension"); }
        else throw FormatError("unexpected token " + tok + " in DiagGmm");
      }
      if (!(have[0] && have[2] && have[3]) || (int)F->gconsts.size() != F->num_gauss) throw FormatError("final.dubm: incomplete DiagGmm");
    }
    {   // final.mat
      Reader rd(lda_mat_path);
      Value v;
      rd.read_matrix(&v);
      F->lda = v.f; F->lda_rows = v.rows; F->lda_cols = v.cols;
      if (v.rows != F->feat_dim) throw FormatError("final.mat: row count differs from the extractor's feature dimension");
    }
    {   // global_cmvn.stats
      Reader rd(global_cmvn_path);
      rd.read_dense_d('M', &F->cmvn, &r, &c);
      if (r != 2 || c < 2) throw FormatError("global_cmvn.stats: expected a 2 x (dim + 1) matrix");
      F->cmvn_dim = c - 1;
    }
  } catch (const std::exception &e) {
    delete F;
    return b2k::set_error(B2K_ERR_INVALID, "b2k_ivec_files_read", e.what());
  }
  *out = F;
  return B2K_OK;
}

int b2k_ivec_files_destroy(b2k_ivec_files *f) { delete f; return B2K_OK; }

int b2k_ivec_files_info(const b2k_ivec_files *f, int32_t info[8], float *prior_offset) {
  if (!f || !info) return b2k::set_error(B2K_ERR_INVALID, "b2k_ivec_files_info: bad args");
  info[0] = f->num_gauss; info[1] = f->feat_dim; info[2] = f->ivector_dim; info[3] = f->lda_rows; info[4] = f->lda_cols;
  info[5] = f->cmvn_dim; info[6] = (int32_t)f->ubm_weights.size(); info[7] = 0;
  if (prior_offset) *prior_offset = (float)f->prior_offset;
  return B2K_OK;
}

const float *b2k_ivec_files_f32(const b2k_ivec_files *f, int32_t which) {
  if (!f) return nullptr;
  switch (which) {
    case 0: return f->lda.data();
    case 1: return f->gconsts.data();
    case 2: return f->means_invvars.data();
    case 3: return f->inv_vars.data();
    case 4: return f->ubm_weights.empty() ? nullptr : f->ubm_weights.data();
    default: return nullptr;
  }
}

const double *b2k_ivec_files_f64(const b2k_ivec_files *f, int32_t which) {
  if (!f) return nullptr;
  switch (which) {
    case 0: return f->sigma_inv_m.data();
    case 1: return f->U.data();
    case 2: return f->cmvn.data();
    default: return nullptr;
  }
}

int b2k_ivec_create_from_files(const b2k_ivec_cfg *cfg, const b2k_ivec_files *f, b2k_ivec **out) {
  if (!cfg || !f || !out) return b2k::set_error(B2K_ERR_INVALID, "b2k_ivec_create_from_files: bad args");
  b2k_ivec_cfg c = *cfg;
  c.feat_dim = f->feat_dim; c.num_gauss = f->num_gauss; c.ivector_dim = f->ivector_dim; c.prior_offset = (float)f->prior_offset;
  if (f->lda_cols != c.base_dim * (c.splice_left + c.splice_right + 1) + 1)
    return b2k::set_error(B2K_ERR_INVALID, "b2k_ivec_create_from_files: final.mat does not match base_dim x (splice_left + 1 + splice_right) + 1 columns");
  if (f->cmvn_dim != c.base_dim) return b2k::set_error(B2K_ERR_INVALID, "b2k_ivec_create_from_files: global_cmvn.stats dimension differs from base_dim");
  return b2k_ivec_create(&c, f->lda.data(), f->gconsts.data(), f->means_invvars.data(), f->inv_vars.data(), f->sigma_inv_m.data(),
                         f->U.data(), f->cmvn.data(), out);
}

}  // extern "C"
