"""Compile check of kaldi_b200/host/b2k_kaldi_shims.h against the reference's own
headers (only possible where /root/reference exists).  Not part of the product."""
import os, subprocess, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_feat as RF


def check() -> bool:
    if not os.path.isdir(RF.SRC):
        return False
    with tempfile.TemporaryDirectory() as td:
        src = os.path.join(td, "t.cc")
        open(src, "w").write('#include "b2k_kaldi_shims.h"\nint main() { return 0; }\n')
        cmd = ["g++", "-fsyntax-only"] + RF.cxxflags(["-I" + os.path.join(ROOT, "include"), "-I" + os.path.join(ROOT, "kaldi_b200", "host"),
                                                      "-I/usr/local/cuda/include"]) + [src]
        subprocess.check_call(cmd)
        # the OpenFst-typed members (CudaFstB2k, CudaDecoderB2k::GetRawLattice), type-checked against the
        # container-only stand-in of oracle/ref_wrap/fst_stub (OpenFst itself is not in this image)
        stub = os.path.join(ROOT, "oracle", "ref_wrap", "fst_stub")
        flags = RF.cxxflags(["-I" + os.path.join(ROOT, "include"), "-I" + os.path.join(ROOT, "kaldi_b200", "host"),
                             "-I/usr/local/cuda/include"])
        flags = [f for f in flags if not f.startswith("-I")] + ["-I" + stub] + [f for f in flags if f.startswith("-I")]
        subprocess.check_call(["g++", "-fsyntax-only", "-DB2K_HAVE_OPENFST", "-DB2K_OPENFST_IS_STANDIN"] + flags + [src])
        # the endpointing templates, instantiated with the reference's own OnlineEndpointConfig.  online-endpoint.h includes both
        # online decoders (OpenFst); their include guards are pre-defined, as in oracle/ref_wrap/endpoint_wrap.cc
        ep = os.path.join(td, "e.cc")
        open(ep, "w").write(
            "#define KALDI_LAT_KALDI_LATTICE_H_\n#define KALDI_DECODER_LATTICE_FASTER_ONLINE_DECODER_H_\n"
            "#define KALDI_DECODER_LATTICE_INCREMENTAL_ONLINE_DECODER_H_\n"
            '#include "online2/online-endpoint.h"\n#include "hmm/transition-model.h"\n#include "b2k_kaldi_shims.h"\n'
            "bool f(const kaldi::OnlineEndpointConfig &c, const kaldi::TransitionModel &tm, const std::vector<kaldi::int32> &path) {\n"
            "  b2k_endpoint_cfg x = kaldi::b2k_shim::ToB2kEndpointConfig(c); (void)x;\n"
            "  return kaldi::b2k_shim::EndpointDetectedB2k(c, 100, 10, 0.03f, 0.0f) ||\n"
            "         kaldi::b2k_shim::EndpointDetectedB2k(c, tm, 0.03f, path, 100, 0.0f);\n}\n")
        subprocess.check_call(["g++", "-fsyntax-only"] + RF.cxxflags(["-DHAVE_CUDA=0", "-I" + os.path.join(ROOT, "include"),
                              "-I" + os.path.join(ROOT, "kaldi_b200", "host"), "-I/usr/local/cuda/include",
                              "-I" + os.path.join(ROOT, "oracle", "_ref", "inc")]) + [ep])
        # CudaDecoderConfig's stand-in registers on the reference's own ParseOptions (util/parse-options.h) and its
        # defaults / ComputeConfig() equal the reference's constants (cudadecoder/cuda-decoder-common.h)
        cc = os.path.join(td, "c.cc")
        open(cc, "w").write(
            '#include "util/parse-options.h"\n#include "b2k_kaldi_shims.h"\n'
            "int f(int argc, const char *const *argv) {\n"
            '  kaldi::ParseOptions po("x");\n  kaldi::b2k_shim::CudaDecoderConfigB2k c;\n  c.Register(&po);\n  po.Read(argc, argv);\n'
            "  c.Check(); c.ComputeConfig();\n  b2k_dec_cfg d = c.ToB2k(400);\n  return d.max_active + c.main_q_capacity;\n}\n")
        subprocess.check_call(["g++", "-fsyntax-only"] + RF.cxxflags(["-I" + os.path.join(ROOT, "include"),
                              "-I" + os.path.join(ROOT, "kaldi_b200", "host"), "-I/usr/local/cuda/include"]) + [cc])
        # the streaming pipeline shim: DecodeBatch with exactly the argument types the reference's own callers pass
        # (cudadecoderbin/batched-wav-nnet3-cuda-online.cc), callbacks, end points
        sp = os.path.join(td, "s.cc")
        open(sp, "w").write(
            '#include "b2k_kaldi_shims.h"\nusing namespace kaldi;\n'
            "void f(const b2k_stream_cfg &c, const b2k_model *m, const b2k_fst *g, const std::vector<SubVector<BaseFloat>> &waves) {\n"
            "  b2k_shim::StreamingOnlinePipelineB2k p(c, m, g);\n"
            "  std::vector<b2k_shim::StreamingOnlinePipelineB2k::CorrelationID> ids(1, 7);\n"
            "  std::vector<bool> first(1, true), last(1, false), ep; std::vector<const std::string *> hyp;\n"
            "  p.TryInitCorrID(7);\n"
            "  p.SetBestPathCallback(7, [](const std::string &, bool, bool) {});\n"
            "  p.SetRawLatticeCallback(7, [](uint64_t, const b2k_raw_lattice &) {});\n"
            "  p.DecodeBatch(ids, waves, first, last, &hyp, &ep);\n"
            "  p.DecodeBatch(ids, waves, first, last);\n"
            "  int x = p.GetNSampsPerChunk() + p.GetNInputFramesPerChunk(); (void)x; p.GetDecoderFrameShiftSeconds();\n"
            "  b2k_shim::CudaOnlinePipelineDynamicBatcherB2k batcher(2e-3, p, 64);\n"
            "  batcher.Push(7, true, false, waves[0]); batcher.WaitForCompletion(); int n = batcher.GetNumPendingChunks(7); (void)n;\n"
            "}\n")
        subprocess.check_call(["g++", "-fsyntax-only"] + RF.cxxflags(["-I" + os.path.join(ROOT, "include"),
                              "-I" + os.path.join(ROOT, "kaldi_b200", "host"), "-I/usr/local/cuda/include"]) + [sp])
        # the nnet3 surfaces (b2k_nnet3_shims.h) against the reference's own nnet3 / cudadecoder headers, HAVE_CUDA=1 as in a CUDA
        # build of Kaldi: NnetComputer's members, DecodableAmNnetLoopedOnline as a DecodableInterface, and BatchedStaticNnet3's
        # RunBatch called with exactly the argument types batched-threaded-nnet3-cuda-online-pipeline.cc:662-668 passes
        nn = os.path.join(td, "n.cc")
        open(nn, "w").write(
            '#include "cudadecoder/batched-static-nnet3.h"\n#include "nnet3/decodable-online-looped.h"\n#include "b2k_nnet3_shims.h"\n'
            "using namespace kaldi;\n"
            "typedef b2k_shim::BatchedStaticNnet3B2k<cuda_decoder::BatchedStaticNnet3Config> Static;\n"
            "void f(const cuda_decoder::BatchedStaticNnet3Config &c, const nnet3::AmNnetSimple &am, const TransitionModel &tm,\n"
            "       const nnet3::DecodableNnetSimpleLoopedInfo &info, OnlineFeatureInterface *feats, OnlineFeatureInterface *ivecs,\n"
            "       const nnet3::NnetComputeOptions &co, const nnet3::ComputationRequest &req) {\n"
            "  Static s(c, am);\n"
            "  std::vector<int> channels, nvalid; std::vector<BaseFloat *> d_features, d_ivectors; std::vector<bool> first, last;\n"
            "  CuMatrix<BaseFloat> post; std::vector<std::vector<std::pair<int, const BaseFloat *>>> ptrs;\n"
            "  s.RunBatch(channels, d_features, 40, d_ivectors, nvalid, first, last, &post, &ptrs);\n"
            "  int a = s.GetNOutputFramesPerChunk() + s.GetTotalNnet3RightContext(); (void)a;\n"
            "  b2k_shim::DecodableAmNnetLoopedOnlineB2k d(tm, info, feats, ivecs);\n"
            "  DecodableInterface *itf = &d;\n"
            "  BaseFloat l = itf->LogLikelihood(0, 1); (void)l; itf->NumFramesReady(); itf->IsLastFrame(0); itf->NumIndices();\n"
            "  d.SetFrameOffset(d.GetFrameOffset()); d.FrameSubsamplingFactor();\n"
            "  b2k_shim::DecodableNnetLoopedOnlineB2k d2(info, feats, ivecs); d2.LogLikelihood(0, 1);\n"
            "  b2k_shim::NnetComputerB2k comp(co, req, am.GetNnet());\n"
            "  CuMatrix<BaseFloat> in, out; comp.AcceptInput(\"input\", &in); comp.Run();\n"
            "  const CuMatrixBase<BaseFloat> &o = comp.GetOutput(\"output\"); (void)o; comp.GetOutputDestructive(\"output\", &out);\n"
            "  // the decodable of nnet3bin/nnet3-latgen-faster.cc:212-216, arguments as the tool passes them\n"
            "  nnet3::NnetSimpleComputationOptions so; Matrix<BaseFloat> features, online_ivectors; Vector<BaseFloat> ivector;\n"
            "  b2k_shim::NnetSimpleComputerB2k shared(so, am.GetNnet());\n"
            "  b2k_shim::DecodableAmNnetSimpleB2k offline(so, tm, am, features, &shared, &ivector, &online_ivectors, 10);\n"
            "  DecodableInterface *oi = &offline; oi->LogLikelihood(0, 1); oi->NumFramesReady(); oi->IsLastFrame(0); oi->NumIndices();\n"
            "  b2k_shim::DecodableNnetSimpleB2k raw(so, am.Priors(), features, &shared); Vector<BaseFloat> row(raw.OutputDim());\n"
            "  raw.GetOutputForFrame(0, &row); raw.NumFrames(); raw.GetOutput(0, 0);\n"
            "}\n")
        subprocess.check_call(["g++", "-fsyntax-only", "-DHAVE_CUDA=1"] + RF.cxxflags(["-I" + os.path.join(ROOT, "include"),
                              "-I" + os.path.join(ROOT, "kaldi_b200", "host"), "-I/usr/local/cuda/include",
                              "-I" + os.path.join(ROOT, "oracle", "_ref", "inc")]) + [nn])
        # the online2 feature pipeline (b2k_online2_shims.h) against the reference's own online2 headers; online-ivector-feature.h
        # pulls in the OpenFst-based decoders for OnlineSilenceWeighting: their guards are pre-defined and the two templates declared
        o2 = os.path.join(td, "o.cc")
        open(o2, "w").write(
            "#define KALDI_DECODER_LATTICE_FASTER_ONLINE_DECODER_H_\n#define KALDI_DECODER_LATTICE_INCREMENTAL_ONLINE_DECODER_H_\n"
            '#include <queue>\n#include "base/kaldi-common.h"\n'
            "namespace kaldi { template <class FST> class LatticeFasterOnlineDecoderTpl; template <class FST> class LatticeIncrementalOnlineDecoderTpl; }\n"
            '#include "b2k_online2_shims.h"\n'
            "using namespace kaldi;\n"
            "void f(const OnlineNnet2FeaturePipelineConfig &cfg, const VectorBase<BaseFloat> &wave) {\n"
            "  OnlineNnet2FeaturePipelineInfo info(cfg);\n"
            "  b2k_shim::FeatureTablesB2k tables(info);\n"
            "  b2k_shim::OnlineNnet2FeaturePipelineB2k p(info, tables);\n"
            "  OnlineFeatureInterface *itf = &p;\n"
            "  OnlineIvectorExtractorAdaptationState st(info.ivector_extractor_info);\n"
            "  p.SetAdaptationState(st); p.AcceptWaveform(16000.0f, wave); p.InputFinished(); p.GetAdaptationState(&st);\n"
            "  OnlineCmvnState cs; p.GetCmvnState(&cs); p.SetCmvnState(cs);\n"
            "  std::vector<std::pair<int32, BaseFloat> > dw; p.UpdateFrameWeights(dw);\n"
            "  Vector<BaseFloat> v(itf->Dim()); itf->GetFrame(0, &v); itf->NumFramesReady(); itf->IsLastFrame(0); p.FrameShiftInSeconds();\n"
            "  OnlineFeatureInterface *in = p.InputFeature(); OnlineIvectorFeature *iv = p.IvectorFeature(); (void)in; (void)iv;\n"
            "  // OnlineBatchedFeaturePipelineCuda's call as batched-threaded-nnet3-cuda-online-pipeline.cc:272-279 makes it\n"
            "  b2k_shim::OnlineBatchedFeaturePipelineB2k bp(cfg, 8160, 4, 16);\n"
            "  std::vector<int32> channels(1, 0), nsamp(1, 8160), nframes(1, 0); std::vector<bool> first(1, true), last(1, false);\n"
            "  CuMatrix<BaseFloat> waves(4, 8160), feats; CuVector<BaseFloat> ivecs;\n"
            "  bp.ComputeFeaturesBatched(1, channels, nsamp, first, last, 16000.0f, waves, &feats, NULL, &nframes);\n"
            "  int32 x = bp.GetMaxChunkFrames() + bp.FeatureDim() + bp.IvectorDim(); (void)x; bp.GetFrameOptions();\n"
            "}\n")
        subprocess.check_call(["g++", "-fsyntax-only", "-DHAVE_CUDA=0"] + RF.cxxflags(["-I" + os.path.join(ROOT, "include"),
                              "-I" + os.path.join(ROOT, "kaldi_b200", "host"), "-I/usr/local/cuda/include",
                              "-I" + os.path.join(ROOT, "oracle", "_ref", "inc")]) + [o2])
        # SingleUtteranceNnet3DecoderB2k with the reference's OWN LatticeFasterDecoderConfig (decoder/lattice-faster-decoder.h, compiled
        # against the container-only OpenFst stand-in as the decoder oracle is) and OnlineEndpointConfig: the calls of
        # online2bin/online2-wav-nnet3-latgen-faster.cc:246-283
        su = os.path.join(td, "u.cc")
        open(su, "w").write(
            '#include "decoder/lattice-faster-decoder.h"\n'
            "#define KALDI_DECODER_LATTICE_FASTER_ONLINE_DECODER_H_\n#define KALDI_DECODER_LATTICE_INCREMENTAL_ONLINE_DECODER_H_\n"
            "namespace kaldi { template <class FST> class LatticeFasterOnlineDecoderTpl; template <class FST> class LatticeIncrementalOnlineDecoderTpl; }\n"
            '#include <queue>\n#include "online2/online-endpoint.h"\n#include "b2k_nnet3_shims.h"\n#include "b2k_online2_shims.h"\n'
            "using namespace kaldi;\n"
            "struct Feats { OnlineFeatureInterface *InputFeature(); OnlineFeatureInterface *IvectorFeature(); BaseFloat FrameShiftInSeconds() const; };\n"
            "void f(const LatticeFasterDecoderConfig &opts, const TransitionModel &tm, const nnet3::DecodableNnetSimpleLoopedInfo &info,\n"
            "       const b2k_fst *g, Feats *feats, const OnlineEndpointConfig &ep) {\n"
            "  b2k_shim::SingleUtteranceNnet3DecoderB2k<LatticeFasterDecoderConfig, Feats> d(opts, tm, info, g, feats);\n"
            "  d.InitDecoding(0); d.AdvanceDecoding(); bool e = d.EndpointDetected(ep); (void)e; d.FinalizeDecoding();\n"
            "  int32 n = d.NumFramesDecoded(); (void)n; BaseFloat c = d.FinalRelativeCost(); (void)c;\n"
            "  Lattice best; d.GetBestPath(true, &best);\n"
            "}\n"
            "// the silence weighting of the tool's loop (online2-wav-nnet3-latgen-faster.cc:222-262) over the b2k decoder and pipeline\n"
            "void g(const LatticeFasterDecoderConfig &opts, const TransitionModel &tm, const nnet3::DecodableNnetSimpleLoopedInfo &info,\n"
            "       const b2k_fst *graph, const OnlineNnet2FeaturePipelineInfo &finfo, const b2k_shim::FeatureTablesB2k &tables) {\n"
            "  b2k_shim::OnlineNnet2FeaturePipelineB2k feature_pipeline(finfo, tables);\n"
            "  b2k_shim::OnlineSilenceWeightingB2k silence_weighting(tm, finfo.silence_weighting_config, info.opts.frame_subsampling_factor);\n"
            "  b2k_shim::SingleUtteranceNnet3DecoderB2k<LatticeFasterDecoderConfig, b2k_shim::OnlineNnet2FeaturePipelineB2k> decoder(opts, tm, info, graph, &feature_pipeline);\n"
            "  std::vector<std::pair<int32, BaseFloat> > delta_weights;\n"
            "  if (silence_weighting.Active() && feature_pipeline.IvectorFeature() != NULL) {\n"
            "    silence_weighting.ComputeCurrentTraceback(decoder);\n"
            "    silence_weighting.GetDeltaWeights(feature_pipeline.NumFramesReady(), 0, &delta_weights);\n"
            "    silence_weighting.GetDeltaWeights(feature_pipeline.NumFramesReady(), &delta_weights);\n"
            "    feature_pipeline.UpdateFrameWeights(delta_weights);\n"
            "    std::vector<int32> ns; silence_weighting.GetNonsilenceFrames(feature_pipeline.NumFramesReady(), 0, &ns);\n"
            "  }\n"
            "  decoder.AdvanceDecoding();\n"
            "}\n")
        stub = os.path.join(ROOT, "oracle", "ref_wrap", "fst_stub")
        flags = RF.cxxflags(["-I" + os.path.join(ROOT, "include"), "-I" + os.path.join(ROOT, "kaldi_b200", "host"),
                             "-I/usr/local/cuda/include", "-I" + os.path.join(ROOT, "oracle", "_ref", "inc")])
        flags = [f for f in flags if not f.startswith("-I")] + ["-I" + stub] + [f for f in flags if f.startswith("-I")]
        subprocess.check_call(["g++", "-fsyntax-only", "-DHAVE_CUDA=1", "-DB2K_HAVE_OPENFST", "-DB2K_OPENFST_IS_STANDIN"] + flags + [su])
        # The reference's TOOLS, as they lie in the reference tree, against the drop-in header (kaldi_b200/host/b2k_online2_dropin.h
        # force-included: OnlineNnet2FeaturePipeline, OnlineSilenceWeighting and SingleUtteranceNnet3Decoder become the b2k adapters;
        # every other statement of the tools is compiled as written).  OpenFst is the container stand-in plus fst_stub_tool/
        # (CompactLattice as a container, DECLARATIONS of the lattice library calls the tools make after decoding).
        tool_stub = os.path.join(ROOT, "oracle", "ref_wrap", "fst_stub_tool")
        tflags = [f for f in flags if not f.startswith("-I")] + ["-I" + tool_stub] + [f for f in flags if f.startswith("-I")]
        decoding = ("b2k_dropin::OnlineNnet2FeaturePipeline feature_pipeline(", "b2k_dropin::SingleUtteranceNnet3Decoder decoder(",
                    "b2k_dropin::OnlineSilenceWeighting silence_weighting(")
        for hdr, tool, adapters in (
                ("b2k_online2_dropin.h", "online2bin/online2-wav-nnet3-latgen-faster.cc", decoding),
                ("b2k_online2_dropin.h", "online2bin/online2-tcp-nnet3-decode-faster.cc", decoding),
                # the feature-side tool: the pipeline alone (features + i-vectors as matrices)
                ("b2k_online2_dropin.h", "online2bin/online2-wav-dump-features.cc", ("b2k_dropin::OnlineNnet2FeaturePipeline feature_pipeline(",)),
                # the offline nnet3 tools: DecodableNnetSimple / DecodableAmNnetSimple (kaldi_b200/host/b2k_nnet3_dropin.h)
                ("b2k_nnet3_dropin.h", "nnet3bin/nnet3-compute.cc", ("b2k_nnet3_dropin::DecodableNnetSimple nnet_computer(",)),
                ("b2k_nnet3_dropin.h", "nnet3bin/nnet3-latgen-faster.cc", ("b2k_nnet3_dropin::DecodableAmNnetSimple nnet_decodable(",))):
            pre = subprocess.run(["g++", "-E", "-DHAVE_CUDA=1", "-include", hdr] + tflags + [os.path.join(RF.SRC, tool)],
                                 check=True, capture_output=True, text=True).stdout
            for adapter in adapters:
                assert adapter in pre, (tool, adapter)          # the tool's own objects ARE the adapters
            if "dump-features" not in tool:                     # ... and its CollapseModel call leaves the model as trained
                assert "B2kLeaveModelAsTrained(nnet3::CollapseModelConfig()" in pre or "B2kLeaveModelAsTrained(CollapseModelConfig()" in pre, tool
            subprocess.check_call(["g++", "-fsyntax-only", "-DHAVE_CUDA=1", "-include", hdr] + tflags + [os.path.join(RF.SRC, tool)])
        # ... and the CUDA online tool against kaldi_b200/host/b2k_cuda_pipeline_dropin.h: BatchedThreadedNnet3CudaOnlinePipeline and
        # CudaOnlinePipelineDynamicBatcher become adapters over the b2k streaming pipeline; option structs, result and callback
        # types, the lattice postprocessor and cuda-bin-tools.h stay the reference's
        for name, adapters in (("cudadecoderbin/batched-wav-nnet3-cuda-online.cc",
                                ("b2k_cuda_dropin::BatchedThreadedNnet3CudaOnlinePipeline cuda_pipeline(",
                                 "b2k_cuda_dropin::CudaOnlinePipelineDynamicBatcher dynamic_batcher(")),
                               # the offline tool (whole utterances, the reference's throughput benchmark): BatchedThreadedNnet3CudaPipeline2
                               ("cudadecoderbin/batched-wav-nnet3-cuda2.cc", ("b2k_cuda_dropin::BatchedThreadedNnet3CudaPipeline2 cuda_pipeline(",))):
            tool = os.path.join(RF.SRC, name)
            pre = subprocess.run(["g++", "-E", "-DHAVE_CUDA=1", "-include", "b2k_cuda_pipeline_dropin.h"] + tflags + [tool],
                                 check=True, capture_output=True, text=True).stdout
            for adapter in adapters:
                assert adapter in pre, (name, adapter)
            assert "B2kLeaveModelAsTrained(nnet3::CollapseModelConfig()" in pre, name
            subprocess.check_call(["g++", "-fsyntax-only", "-DHAVE_CUDA=1", "-include", "b2k_cuda_pipeline_dropin.h"] + tflags + [tool])
    return True


if __name__ == "__main__":
    print("shims compile against the reference headers:", check())
