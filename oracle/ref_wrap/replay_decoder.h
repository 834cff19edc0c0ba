// TEST INFRASTRUCTURE ONLY (oracle side).  A stand-in for the reference's online decoders that both endpoint_wrap.cc and
// silence_wrap.cc use to drive the reference's own code (online2/online-endpoint.cc, online2/online-ivector-feature.cc).
// Those files include decoder headers that need OpenFst, which this image does not have, so the headers' include guards are
// pre-defined and the decoder class templates are declared here as a replay device: a decoder that "has decoded" a given
// best path and hands it back, last arc first, through the BestPathEnd / TraceBackBestPath interface
// (decoder/lattice-faster-online-decoder.h:62-118).  The reference code sees the interface it was written for.
// ONE definition for every translation unit of oracle/_ref/libkaldi_ref_nnet3.so (the class names are the reference's).
#ifndef B2K_ORACLE_REPLAY_DECODER_H_
#define B2K_ORACLE_REPLAY_DECODER_H_

#define KALDI_LAT_KALDI_LATTICE_H_
#define KALDI_DECODER_LATTICE_FASTER_ONLINE_DECODER_H_
#define KALDI_DECODER_LATTICE_INCREMENTAL_ONLINE_DECODER_H_
#define KALDI_DECODER_GRAMMAR_FST_H_
#include <cstdint>
#include <queue>            // online-ivector-feature.h uses std::priority_queue (it gets <queue> from the decoder headers)
#include <vector>
#include "base/kaldi-common.h"
#include <fst/fst-decl.h>     // forward declarations only (written by oracle/ref_nnet.py)

namespace fst {
struct ConstGrammarFst {};
struct VectorGrammarFst {};
}  // namespace fst

namespace kaldi {
struct LatticeArc { int ilabel = 0, olabel = 0, nextstate = 0; };

struct ReplayDecoder {
  std::vector<int> path;                    // ilabels of the best path in time order, epsilons (0) allowed
  std::vector<int> source;                  // per arc: an id of the token the arc leaves (empty: tokens are not looked at)
  int frames = 0;
  float final_relative_cost = 0.0f;
  struct BestPathIterator {
    void *tok;                              // the token the walk stands on
    int frame;                              // the reference's convention: index of the last frame behind `tok`, -1 at the start
    int pos;                                // index of the arc handed out next; -1 = done
    bool Done() const { return pos < 0; }
  };
  BestPathIterator BestPathEnd(bool, BaseFloat *) const {
    return BestPathIterator{NULL, frames - 1, (int)path.size() - 1};
  }
  BestPathIterator TraceBackBestPath(BestPathIterator it, LatticeArc *arc) const {
    arc->ilabel = path[it.pos];
    void *tok = source.empty() ? NULL : (void *)(intptr_t)(source[it.pos] + 16);
    return BestPathIterator{tok, it.frame - (arc->ilabel != 0 ? 1 : 0), it.pos - 1};
  }
  int32 NumFramesDecoded() const { return frames; }
  BaseFloat FinalRelativeCost() const { return final_relative_cost; }
};
template <class F> struct LatticeFasterOnlineDecoderTpl : public ReplayDecoder {};
template <class F> struct LatticeIncrementalOnlineDecoderTpl : public ReplayDecoder {};
}  // namespace kaldi

#endif
