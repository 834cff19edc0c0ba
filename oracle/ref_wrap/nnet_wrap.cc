// oracle/ref_wrap/nnet_wrap.cc — TEST INFRASTRUCTURE ONLY.
// Thin extern "C" wrapper (our code) around the reference's OWN nnet3 CPU
// forward, compiled from /root/reference/src (oracle/ref_nnet.py).  It only
// calls public reference API:
//   Nnet::ReadConfig                       nnet3/nnet-nnet.cc:189
//   UpdatableComponent::{Vectorize,UnVectorize}, Component::ReadNew
//   SetBatchnormTestMode / SetDropoutTestMode / CollapseModel   nnet3/nnet-utils.cc
//   AmNnetSimple, DecodableNnetSimpleLoopedInfo, DecodableNnetSimpleLooped
//                                          nnet3/decodable-simple-looped.{h,cc}
#include <fstream>
#include <cstring>
#include <memory>
#include <sstream>
#include <string>
#include <vector>

#include "hmm/hmm-topology.h"
#include "hmm/transition-model.h"
#include "nnet3/am-nnet-simple.h"
#include "tree/context-dep.h"
#include "util/kaldi-io.h"
#include "nnet3/decodable-simple-looped.h"
#include "nnet3/nnet-am-decodable-simple.h"
#include "nnet3/nnet-nnet.h"
#include "nnet3/nnet-normalize-component.h"
#include "nnet3/nnet-simple-component.h"
#include "nnet3/nnet-utils.h"

using namespace kaldi;
using namespace kaldi::nnet3;

struct RefNnet {
  Nnet nnet;
  std::unique_ptr<AmNnetSimple> am;
  std::unique_ptr<AmNnetSimple> am_simple;   // the same model before DecodableNnetSimpleLoopedInfo rewrites its i-vector descriptors in place
  std::unique_ptr<DecodableNnetSimpleLoopedInfo> info;
  NnetSimpleLoopedComputationOptions opts;
  std::string tmp;
};

extern "C" {

void *ref_nnet_create(const char *config_text) {
  try {
    RefNnet *r = new RefNnet();
    std::istringstream is(config_text);
    r->nnet.ReadConfig(is);
    return r;
  } catch (const std::exception &e) { fprintf(stderr, "ref_nnet_create: %s\n", e.what()); return nullptr; }
}
void ref_nnet_destroy(void *h) { delete (RefNnet *)h; }

int ref_nnet_num_components(void *h) { return ((RefNnet *)h)->nnet.NumComponents(); }
const char *ref_nnet_component_name(void *h, int i) {
  RefNnet *r = (RefNnet *)h; r->tmp = r->nnet.GetComponentName(i); return r->tmp.c_str();
}
const char *ref_nnet_component_type(void *h, int i) {
  RefNnet *r = (RefNnet *)h; r->tmp = r->nnet.GetComponent(i)->Type(); return r->tmp.c_str();
}
// number of parameters of an updatable component (0 if not updatable)
int ref_nnet_num_params(void *h, int i) {
  Component *c = ((RefNnet *)h)->nnet.GetComponent(i);
  if (!(c->Properties() & kUpdatableComponent)) return 0;
  return dynamic_cast<UpdatableComponent *>(c)->NumParameters();
}
int ref_nnet_get_params(void *h, int i, float *out) {
  try {
    UpdatableComponent *c = dynamic_cast<UpdatableComponent *>(((RefNnet *)h)->nnet.GetComponent(i));
    if (!c) return -1;
    SubVector<BaseFloat> v(out, c->NumParameters());
    c->Vectorize(&v);
    return 0;
  } catch (...) { return -1; }
}
int ref_nnet_set_params(void *h, int i, const float *in) {
  try {
    UpdatableComponent *c = dynamic_cast<UpdatableComponent *>(((RefNnet *)h)->nnet.GetComponent(i));
    if (!c) return -1;
    SubVector<BaseFloat> v(const_cast<float *>(in), c->NumParameters());
    c->UnVectorize(v);
    return 0;
  } catch (...) { return -1; }
}
// replace a BatchNormComponent by one with the given stats (text Read path,
// nnet-normalize-component.cc:591-614)
int ref_nnet_set_batchnorm(void *h, int i, int dim, int block_dim, float epsilon, float target_rms, float count,
                           const float *mean, const float *var) {
  try {
    RefNnet *r = (RefNnet *)h;
    std::ostringstream os;
    os.precision(9);
    os << "<BatchNormComponent> <Dim> " << dim << " <BlockDim> " << block_dim << " <Epsilon> " << epsilon
       << " <TargetRms> " << target_rms << " <TestMode> F <Count> " << count << " <StatsMean> [ ";
    for (int d = 0; d < block_dim; d++) os << mean[d] << " ";      // statistics are per block element
    os << "] <StatsVar> [ ";
    for (int d = 0; d < block_dim; d++) os << var[d] << " ";
    os << "] </BatchNormComponent> ";
    std::istringstream is(os.str());
    Component *c = Component::ReadNew(is, false);
    if (!c) return -1;
    r->nnet.SetComponent(i, c);
    return 0;
  } catch (const std::exception &e) { fprintf(stderr, "ref_nnet_set_batchnorm: %s\n", e.what()); return -1; }
}

// test-mode + CollapseModel + looped compilation (what online2-wav-nnet3-latgen-faster.cc:162-177 does)
int ref_nnet_prepare(void *h, int frames_per_chunk, int frame_subsampling_factor, float acoustic_scale,
                     const float *priors, int num_priors, int collapse) {
  try {
    RefNnet *r = (RefNnet *)h;
    SetBatchnormTestMode(true, &r->nnet);
    SetDropoutTestMode(true, &r->nnet);
    if (collapse) CollapseModel(CollapseModelConfig(), &r->nnet);
    r->am.reset(new AmNnetSimple(r->nnet));
    if (priors && num_priors > 0) {
      Vector<BaseFloat> p(num_priors);
      for (int i = 0; i < num_priors; i++) p(i) = priors[i];
      r->am->SetPriors(p);
    }
    r->opts.frames_per_chunk = frames_per_chunk;
    r->opts.frame_subsampling_factor = frame_subsampling_factor;
    r->opts.acoustic_scale = acoustic_scale;
    r->am_simple.reset(new AmNnetSimple(*r->am));        // (ModifyNnetIvectorPeriod, decodable-simple-looped.cc:74, changes r->am's nnet)
    r->info.reset(new DecodableNnetSimpleLoopedInfo(r->opts, r->am.get()));
    return 0;
  } catch (const std::exception &e) { fprintf(stderr, "ref_nnet_prepare: %s\n", e.what()); return -1; }
}

// out: [left_context, right_context, frames_per_chunk, output_dim]
int ref_nnet_info(void *h, int *out) {
  RefNnet *r = (RefNnet *)h;
  if (!r->info) return -1;
  out[0] = r->info->frames_left_context; out[1] = r->info->frames_right_context;
  out[2] = r->info->frames_per_chunk; out[3] = r->info->output_dim;
  return 0;
}

// DecodableNnetSimpleLooped over a whole utterance; ivectors [M x idim] with
// the given period (decodable-simple-looped.cc:262-279), or M = 0 for none.
int ref_nnet_forward(void *h, const float *feats, int T, int D, const float *ivectors, int M, int idim,
                     int period, float *out, int max_rows) {
  try {
    RefNnet *r = (RefNnet *)h;
    Matrix<BaseFloat> f(T, D);
    for (int t = 0; t < T; t++) memcpy(f.RowData(t), feats + (size_t)t * D, 4 * D);
    Matrix<BaseFloat> iv;
    if (M > 0) { iv.Resize(M, idim); for (int m = 0; m < M; m++) memcpy(iv.RowData(m), ivectors + (size_t)m * idim, 4 * idim); }
    DecodableNnetSimpleLooped dec(*r->info, f, NULL, M > 0 ? &iv : NULL, M > 0 ? period : 1);
    int n = dec.NumFrames(), P = dec.OutputDim();
    if (n > max_rows) return -2;
    for (int t = 0; t < n; t++) { SubVector<BaseFloat> row(out + (size_t)t * P, P); dec.GetOutputForFrame(t, &row); }
    return n;
  } catch (const std::exception &e) { fprintf(stderr, "ref_nnet_forward: %s\n", e.what()); return -1; }
}

// DecodableNnetSimple over a whole utterance (nnet3/nnet-am-decodable-simple.cc: the decodable of the offline tools): chunks of
// frames_per_chunk outputs, each from its own clamped window, ONE i-vector per chunk (the row nearest the chunk's middle frame).
// online_ivectors [M x idim] with the given period, or M = 0 for none.  Uses the options of ref_nnet_prepare.
int ref_nnet_forward_simple(void *h, const float *feats, int T, int D, const float *ivectors, int M, int idim, int period,
                            int frames_per_chunk, float *out, int max_rows) {
  try {
    RefNnet *r = (RefNnet *)h;
    Matrix<BaseFloat> f(T, D);
    for (int t = 0; t < T; t++) memcpy(f.RowData(t), feats + (size_t)t * D, 4 * D);
    Matrix<BaseFloat> iv;
    if (M > 0) { iv.Resize(M, idim); for (int m = 0; m < M; m++) memcpy(iv.RowData(m), ivectors + (size_t)m * idim, 4 * idim); }
    NnetSimpleComputationOptions so;
    so.frames_per_chunk = frames_per_chunk;
    so.frame_subsampling_factor = r->opts.frame_subsampling_factor;
    so.acoustic_scale = r->opts.acoustic_scale;
    CachingOptimizingCompiler compiler(r->am_simple->GetNnet(), so.optimize_config, so.compiler_config);
    DecodableNnetSimple dec(so, r->am_simple->GetNnet(), r->am_simple->Priors(), f, &compiler, NULL, M > 0 ? &iv : NULL, M > 0 ? period : 0);
    int n = dec.NumFrames(), P = dec.OutputDim();
    if (n > max_rows) return -2;
    for (int t = 0; t < n; t++) { SubVector<BaseFloat> row(out + (size_t)t * P, P); dec.GetOutputForFrame(t, &row); }
    return n;
  } catch (const std::exception &e) { fprintf(stderr, "ref_nnet_forward_simple: %s\n", e.what()); return -1; }
}

// Nnet::Write (nnet3/nnet-nnet.cc:630) of the model as it stands (call before ref_nnet_prepare for the
// un-collapsed training-form model): the "raw" nnet3 file the readers in kaldi_b200/kaldi_io.py are
// pinned against.
int ref_nnet_write(void *h, const char *path, int binary) {
  try {
    RefNnet *r = (RefNnet *)h;
    Output ko(path, binary != 0);     // with the "\\0B" header, as nnet3-copy writes final.raw
    r->nnet.Write(ko.Stream(), binary != 0);
    return ko.Close() ? 0 : -1;
  } catch (const std::exception &e) { fprintf(stderr, "ref_nnet_write: %s\n", e.what()); return -1; }
}

// final.mdl as nnet3-am-init / nnet3-am-copy write it: TransitionModel::Write (hmm/transition-model.cc:422)
// followed by AmNnetSimple::Write (nnet3/am-nnet-simple.cc:34).  The transition model is built by the reference
// from a topology text and a monophone tree (tree/context-dep.cc:331); tid2pdf_out[t] = TransitionIdToPdf(t)
// for t = 1..NumTransitionIds (index 0 unused) is returned for the reader's parity test.
int ref_write_final_mdl(void *h, const char *path, int binary, const char *topo_text, int num_phones,
                        int pdf_classes_per_phone, const float *priors, int num_priors,
                        int *tid2pdf_out, int max_tids) {
  try {
    RefNnet *r = (RefNnet *)h;
    HmmTopology topo;
    { std::istringstream is(topo_text); topo.Read(is, false); }
    std::vector<int32> phones, num_classes(num_phones + 1, 0);
    for (int p = 1; p <= num_phones; p++) { phones.push_back(p); num_classes[p] = pdf_classes_per_phone; }
    std::unique_ptr<ContextDependency> ctx(MonophoneContextDependency(phones, num_classes));
    TransitionModel tm(*ctx, topo);
    AmNnetSimple am(r->nnet);
    if (priors && num_priors > 0) {
      Vector<BaseFloat> pv(num_priors);
      for (int i = 0; i < num_priors; i++) pv(i) = priors[i];
      am.SetPriors(pv);
    }
    Output ko(path, binary != 0);
    tm.Write(ko.Stream(), binary != 0);
    am.Write(ko.Stream(), binary != 0);
    if (!ko.Close()) return -1;
    const int n = tm.NumTransitionIds();
    if (n + 1 > max_tids) return -2;
    tid2pdf_out[0] = 0;
    for (int t = 1; t <= n; t++) tid2pdf_out[t] = tm.TransitionIdToPdf(t);
    return n;
  } catch (const std::exception &e) { fprintf(stderr, "ref_write_final_mdl: %s\n", e.what()); return -1; }
}

// AmNnetSimple::Write alone (nnet3/am-nnet-simple.cc:34-57) -- what kaldi_b200/host/b2k_nnet3_shims.h serialises into memory
// from the AmNnetSimple a Kaldi tool holds; with_header = the "\0B" marker Output writes.
int ref_write_am_nnet(void *h, const char *path, int binary, int with_header, const float *priors, int num_priors) {
  try {
    RefNnet *r = (RefNnet *)h;
    AmNnetSimple am(r->nnet);
    if (priors && num_priors > 0) {
      Vector<BaseFloat> pv(num_priors);
      for (int i = 0; i < num_priors; i++) pv(i) = priors[i];
      am.SetPriors(pv);
    }
    if (with_header) {
      Output ko(path, binary != 0);
      am.Write(ko.Stream(), binary != 0);
      return ko.Close() ? 0 : -1;
    }
    std::ofstream os(path, std::ios::binary);
    am.Write(os, binary != 0);
    return os.good() ? 0 : -1;
  } catch (const std::exception &e) { fprintf(stderr, "ref_write_am_nnet: %s\n", e.what()); return -1; }
}

// Matrix<float>::Read / Vector<float>::Read through Input (util/kaldi-io.h), binary or text: the reference-side
// check of the writers in kaldi_b200/kaldi_io.py.  Returns rows (matrix) or dim (vector), -1 on failure.
int ref_read_matrix(const char *path, float *out, int max_elems, int *cols) {
  try {
    bool binary;
    Input ki(path, &binary);
    Matrix<BaseFloat> m;
    m.Read(ki.Stream(), binary);
    if ((long long)m.NumRows() * m.NumCols() > max_elems) return -2;
    for (int r = 0; r < m.NumRows(); r++) memcpy(out + (size_t)r * m.NumCols(), m.RowData(r), sizeof(float) * m.NumCols());
    *cols = m.NumCols();
    return m.NumRows();
  } catch (const std::exception &e) { fprintf(stderr, "ref_read_matrix: %s\n", e.what()); return -1; }
}
int ref_read_vector(const char *path, float *out, int max_elems) {
  try {
    bool binary;
    Input ki(path, &binary);
    Vector<BaseFloat> v;
    v.Read(ki.Stream(), binary);
    if (v.Dim() > max_elems) return -2;
    memcpy(out, v.Data(), sizeof(float) * v.Dim());
    return v.Dim();
  } catch (const std::exception &e) { fprintf(stderr, "ref_read_vector: %s\n", e.what()); return -1; }
}

}  // extern "C"
