"""TDNN-F chain acoustic models for the nnet3 stage: architecture descriptions
with the exact layer shapes of the named recipes, seeded synthetic weights
(there are no trained models in the reference tree, SURVEY.md §7 hard part 8),
the nnet3 config text the reference itself can read (Nnet::ReadConfig,
nnet3/nnet-nnet.cc:189), and the compiler from an architecture to the flat op
program the CUDA executor runs (kaldi_b200/csrc/nnet.cu).

The compile step plays the role of the reference's nnet3 compiler
(nnet3/nnet-compile.cc, nnet-compile-looped.cc) for this model family: it
works out, per node, the time grid (step 1 or frame_subsampling_factor) and
range that the requested outputs need, exactly as the reference only computes
required Indexes, and turns descriptors (Append/Offset/Sum/Scale/ReplaceIndex,
nnet3/nnet-descriptor.h) into row mappings of fused GEMM / elementwise ops.

Layer types (xconfig name -> components it expands to, composite_layers.py /
basic_layers.py / trivial_layers.py of egs/wsj/s5/steps/libs/nnet3/xconfig):
  idct            FixedAffineComponent
  batchnorm       BatchNormComponent (test mode)
  delta           NoOp over Append(Offset..)/Sum/Scale + BatchNorm   (trivial_layers.py:189-257)
  lda             FixedAffineComponent over Append(-1,0,1,ReplaceIndex(ivector,t,0)); `time_offsets` gives another splice
                  (Append(-2,-1,0,1,2,...) of the chain TDNN recipes, wsj run_tdnn_1f.sh:171)
  relu-batchnorm  NaturalGradientAffine + ReLU + BatchNorm; `time_offsets` = the input=Append(-1,0,1) / Append(-3,0,3) /
                  Append(-6,-3,0) splice of a TDNN layer without factorisation
  tdnnf           TdnnComponent(linear) + TdnnComponent(affine) + ReLU + BatchNorm + NoOp(Sum(Scale(bypass,in),bn))
  linear          LinearComponent
  prefinal        NaturalGradientAffine + ReLU + BatchNorm + Linear + BatchNorm
  output          NaturalGradientAffine (no log-softmax for chain) [+ LogSoftmax]
  ivector-linear-bn  LinearComponent over ReplaceIndex(ivector,t,0) + BatchNorm(target-rms)   (CNN-TDNN-F front)
  combine         PermuteComponent interleaving the feature map of `cur` with a side node (combine-feature-maps-layer)
  conv            TimeHeightConvolutionComponent + ReLU + BatchNorm(block-dim = filters) (conv-relu-batchnorm-layer)
"""
from __future__ import annotations

import os
from dataclasses import dataclass, field

import numpy as np

# ----------------------------------------------------------------------------- architectures


def arch_tdnnf(name: str, feat_dim: int, ivector_dim: int, num_pdfs: int, dim: int, bottleneck: int,
               strides, prefinal_small: int, front: str, log_softmax: bool = False) -> dict:
    layers = []
    if front == "idct-delta":          # mini_librispeech run_tdnn_1k.sh:176-186
        layers += [dict(type="idct", name="idct", dim=feat_dim),
                   dict(type="batchnorm", name="batchnorm0"),
                   dict(type="delta", name="delta"),
                   dict(type="relu-batchnorm", name="tdnn1", dim=dim, append_ivector=0.4)]
    elif front == "lda":               # librispeech run_tdnn_1d.sh:219-226
        layers += [dict(type="lda", name="lda"),
                   dict(type="relu-batchnorm", name="tdnn1", dim=dim)]
    else:
        raise ValueError(front)
    for i, s in enumerate(strides):
        layers.append(dict(type="tdnnf", name=f"tdnnf{i + 2}", dim=dim, bottleneck=bottleneck, stride=s,
                           bypass=0.66))
    layers += [dict(type="linear", name="prefinal-l", dim=prefinal_small),
               dict(type="prefinal", name="prefinal-chain", small=prefinal_small, big=dim),
               dict(type="output", name="output", dim=num_pdfs, log_softmax=log_softmax)]
    return dict(name=name, feat_dim=feat_dim, ivector_dim=ivector_dim, num_pdfs=num_pdfs,
                frame_subsampling_factor=3, layers=layers)


def arch_cnn_tdnnf(name: str, feat_dim: int, ivector_dim: int, num_pdfs: int, ivector_linear_dim: int, convs,
                   dim: int, bottleneck_first: int, bottleneck: int, n_tdnnf: int, prefinal_small: int) -> dict:
    """egs/librispeech/s5/local/chain/tuning/run_cnn_tdnn_1a.sh:127-165.  convs: list of
    (height_in, height_out, height_subsample_out, num_filters_out); 3x3 patches (time and height offsets -1,0,1)."""
    assert ivector_linear_dim % feat_dim == 0
    f2 = ivector_linear_dim // feat_dim
    layers = [dict(type="idct", name="idct", dim=feat_dim),
              dict(type="ivector-linear-bn", name="ivector", dim=ivector_linear_dim, target_rms=0.025),
              dict(type="batchnorm", name="idct-batchnorm"),
              dict(type="combine", name="combine_inputs", side="ivector-batchnorm", height=feat_dim, filters1=1, filters2=f2)]
    fin = 1 + f2
    for i, (hi, ho, sub, fo) in enumerate(convs):
        layers.append(dict(type="conv", name=f"cnn{i + 1}", height_in=hi, height_out=ho, height_subsample_out=sub,
                           filters_in=fin, filters_out=fo, time_offsets=[-1, 0, 1], height_offsets=[-1, 0, 1]))
        fin = fo
    first = len(convs) + 1
    layers.append(dict(type="tdnnf", name=f"tdnnf{first}", dim=dim, bottleneck=bottleneck_first, stride=0, bypass=0.0))
    for i in range(n_tdnnf - 1):
        layers.append(dict(type="tdnnf", name=f"tdnnf{first + 1 + i}", dim=dim, bottleneck=bottleneck, stride=3, bypass=0.75))
    layers += [dict(type="linear", name="prefinal-l", dim=prefinal_small),
               dict(type="prefinal", name="prefinal-chain", small=prefinal_small, big=dim),
               dict(type="output", name="output", dim=num_pdfs, log_softmax=False)]
    return dict(name=name, feat_dim=feat_dim, ivector_dim=ivector_dim, num_pdfs=num_pdfs,
                frame_subsampling_factor=3, layers=layers)


def arch_librispeech_cnn_tdnn_1a(num_pdfs: int = 6024) -> dict:
    """BASELINE config 3: librispeech CNN-TDNN-F (run_cnn_tdnn_1a.sh)."""
    return arch_cnn_tdnnf("librispeech_cnn_tdnn_1a", 40, 100, num_pdfs, 200,
                          [(40, 40, 1, 64), (40, 40, 1, 64), (40, 20, 2, 128), (20, 20, 1, 128), (20, 10, 2, 256), (10, 10, 1, 256)],
                          1536, 256, 160, 12, 256)


def arch_tiny_cnn(num_pdfs: int = 48) -> dict:
    return arch_cnn_tdnnf("tiny_cnn_tdnnf", 40, 100, num_pdfs, 80, [(40, 40, 1, 4), (40, 20, 2, 6), (20, 10, 2, 8)],
                          48, 16, 12, 3, 24)


def arch_mini_librispeech_1k(num_pdfs: int = 2336) -> dict:
    """egs/mini_librispeech/s5/local/chain/tuning/run_tdnn_1k.sh:170-207 (5.2 M params)."""
    return arch_tdnnf("mini_librispeech_tdnn_1k", 40, 100, num_pdfs, 768, 96,
                      [1, 1, 1, 0, 3, 3, 3, 3, 3, 3, 3, 3], 192, "idct-delta")


def arch_librispeech_1d(num_pdfs: int = 6024) -> dict:
    """egs/librispeech/s5/local/chain/tuning/run_tdnn_1d.sh:219-253 (22.6 M params)."""
    return arch_tdnnf("librispeech_tdnn_1d", 40, 100, num_pdfs, 1536, 160,
                      [1, 1, 1, 0, 3, 3, 3, 3, 3, 3, 3, 3, 3, 3, 3, 3], 256, "lda")


def splice_of(L: dict) -> list:
    """Time offsets a layer splices its input over (the Append(-1,0,1) of an xconfig `input=`): `time_offsets` when given,
    else what the TDNN-F recipes use: -1,0,1 in front of the LDA-like transform, none for a relu-batchnorm-layer."""
    v = L.get("time_offsets")
    return [int(o) for o in v] if v else ([-1, 0, 1] if L["type"] == "lda" else [0])


def arch_tdnn(name: str, feat_dim: int, ivector_dim: int, num_pdfs: int, dim: int, splices, lda_splice=(-2, -1, 0, 1, 2)) -> dict:
    """The chain TDNN without factorisation (egs/wsj/s5/local/chain/tuning/run_tdnn_1f.sh:165-186 and 89 more recipes):
    fixed-affine-layer over Append(-2..2, ivector), relu-batchnorm-layers whose inputs are spliced Append(-1,0,1) /
    Append(-3,0,3) / Append(-6,-3,0), prefinal-chain, output-layer."""
    layers = [dict(type="lda", name="lda", time_offsets=list(lda_splice))]
    for i, sp in enumerate(splices):
        L = dict(type="relu-batchnorm", name=f"tdnn{i + 1}", dim=dim)
        if list(sp) != [0]:
            L["time_offsets"] = list(sp)
        layers.append(L)
    layers += [dict(type="relu-batchnorm", name="prefinal-chain", dim=dim),
               dict(type="output", name="output", dim=num_pdfs, log_softmax=False)]
    return dict(name=name, feat_dim=feat_dim, ivector_dim=ivector_dim, num_pdfs=num_pdfs, frame_subsampling_factor=3, layers=layers)


def arch_wsj_tdnn_1f(num_pdfs: int = 2880, dim: int = 448) -> dict:
    """egs/wsj/s5/local/chain/tuning/run_tdnn_1f.sh:165-186."""
    return arch_tdnn("wsj_tdnn_1f", 40, 100, num_pdfs, dim,
                     [[0], [-1, 0, 1], [0], [-1, 0, 1], [0], [-3, 0, 3], [-3, 0, 3], [-6, -3, 0]])


def arch_tiny_tdnn(num_pdfs: int = 56) -> dict:
    return arch_tdnn("tiny_tdnn", 40, 100, num_pdfs, 48, [[0], [-1, 0, 1], [-3, 0, 3], [-6, -3, 0]])


def arch_tiny(num_pdfs: int = 64, front: str = "idct-delta") -> dict:
    return arch_tdnnf("tiny_tdnnf", 40, 100, num_pdfs, 64, 16, [1, 0, 3, 3], 32, front)


# ----------------------------------------------------------------------------- weights

def random_weights(arch: dict, seed: int = 0) -> dict:
    """Seeded synthetic parameters: weights N(0, 1/fan_in), biases N(0, 0.1),
    BatchNorm mean N(0, 0.1), var U(0.5, 1.5)."""
    rng = np.random.default_rng(seed)
    W = {}

    def lin(name, out_dim, in_dim, bias=True):
        W[name + ".w"] = (rng.standard_normal((out_dim, in_dim)) / np.sqrt(in_dim)).astype(np.float32)
        if bias:
            W[name + ".b"] = (rng.standard_normal(out_dim) * 0.1).astype(np.float32)

    def bn(name, dim):
        W[name + ".mean"] = (rng.standard_normal(dim) * 0.1).astype(np.float32)
        W[name + ".var"] = rng.uniform(0.5, 1.5, dim).astype(np.float32)

    fd, ivd = arch["feat_dim"], arch["ivector_dim"]
    cur = fd
    for L in arch["layers"]:
        t, n = L["type"], L["name"]
        if t == "idct":
            lin(n, fd, fd)
            cur = fd
        elif t == "batchnorm":
            bn(n, cur)
        elif t == "delta":
            bn(n, 3 * cur)
            cur = 3 * cur
        elif t == "lda":
            k = len(splice_of(L)) * cur + ivd
            lin(n, k, k)
            cur = k
        elif t == "relu-batchnorm":
            k = cur * len(splice_of(L)) + (ivd if L.get("append_ivector") else 0)
            lin(n + ".affine", L["dim"], k)
            bn(n + ".batchnorm", L["dim"])
            cur = L["dim"]
        elif t == "tdnnf":
            nt = 1 if L["stride"] == 0 else 2
            lin(n + ".linear", L["bottleneck"], cur * nt, bias=False)
            lin(n + ".affine", L["dim"], L["bottleneck"] * nt)
            bn(n + ".batchnorm", L["dim"])
            cur = L["dim"]
        elif t == "linear":
            lin(n, L["dim"], cur, bias=False)
            cur = L["dim"]
        elif t == "prefinal":
            lin(n + ".affine", L["big"], cur)
            bn(n + ".batchnorm1", L["big"])
            lin(n + ".linear", L["small"], L["big"], bias=False)
            bn(n + ".batchnorm2", L["small"])
            cur = L["small"]
        elif t == "output":
            lin(n + ".affine", L["dim"], cur)
            cur = L["dim"]
        elif t == "ivector-linear-bn":
            lin(n + "-linear", L["dim"], ivd, bias=False)
            bn(n + "-batchnorm", L["dim"])
        elif t == "combine":
            cur = cur + L["height"] * L["filters2"]
        elif t == "conv":
            k = len(L["time_offsets"]) * len(L["height_offsets"]) * L["filters_in"]
            lin(n + ".conv", L["filters_out"], k)
            bn(n + ".batchnorm", L["filters_out"])        # block-dim = filters_out: one statistic per filter
            cur = L["height_out"] * L["filters_out"]
        else:
            raise ValueError(t)
    p = rng.uniform(0.5, 1.5, arch["num_pdfs"])
    W["priors"] = (p / p.sum()).astype(np.float32)
    return W


def apply_output_calibration(W: dict, raw_output_mean: np.ndarray, scale: float) -> dict:
    """Prior-style calibration of a synthetic model: subtract the per-pdf mean of
    the raw output over calibration data (a random network's output is
    dominated by a static per-pdf offset, which real systems remove through the
    priors estimated as the average posterior) and scale the residual so the
    per-frame spread of the log-likelihoods is speech-like (std ~1).  Without it
    the beam collapses to < 300 tokens/frame; with it decoding runs in the
    2-8 k tokens/frame regime of SURVEY.md §8a.  Returns a new weight dict."""
    W2 = dict(W)
    W2["output.affine.w"] = (W["output.affine.w"] * np.float32(scale)).astype(np.float32)
    W2["output.affine.b"] = ((W["output.affine.b"] - raw_output_mean.astype(np.float32)) * np.float32(scale)).astype(np.float32)
    return W2


def num_parameters(arch: dict, W: dict) -> int:
    return int(sum(v.size for k, v in W.items() if k.endswith(".w") or k.endswith(".b")))


# ----------------------------------------------------------------------------- nnet3 config text (for the reference)

def _write_kaldi_matrix(path: str, m: np.ndarray) -> None:
    with open(path, "w") as f:
        f.write(" [\n")
        for r in m:
            f.write("  " + " ".join(repr(float(x)) for x in r) + "\n")
        f.write(" ]\n")


def to_nnet3_config(arch: dict, W: dict, tmpdir: str) -> str:
    """nnet3 config lines (what steps/nnet3/xconfig_to_configs.py emits for these
    layers) for Nnet::ReadConfig.  FixedAffine matrices go to files in tmpdir."""
    fd, ivd = arch["feat_dim"], arch["ivector_dim"]
    c = [f"input-node name=ivector dim={ivd}", f"input-node name=input dim={fd}"]
    cur, cur_dim = "input", fd
    extras = bool(arch.get("recipe_extras"))     # tests: dropout components and the xent branch of a trained recipe model
    xent_from = None
    for L in arch["layers"]:
        t, n = L["type"], L["name"]
        if t in ("idct", "lda"):
            path = os.path.join(tmpdir, f"{n}.mat")
            _write_kaldi_matrix(path, np.concatenate([W[n + ".w"], W[n + ".b"][:, None]], 1))
            if t == "lda":
                sp = ", ".join(cur if o == 0 else f"Offset({cur}, {o})" for o in splice_of(L))
                inp = f"Append({sp}, ReplaceIndex(ivector, t, 0))"
                cur_dim = len(splice_of(L)) * cur_dim + ivd
            else:
                inp = cur
            c.append(f"component name={n} type=FixedAffineComponent matrix={path}")
            c.append(f"component-node name={n} component={n} input={inp}")
            cur = n
        elif t == "batchnorm":
            c.append(f"component name={n} type=BatchNormComponent dim={cur_dim}")
            c.append(f"component-node name={n} component={n} input={cur}")
            cur = n
            if extras:       # spec-augment-layer (basic_layers.py:1277-1360): two masking components, identity in test mode
                sa = n + "-spec-augment"
                c.append(f"component name={sa}.freq-mask type=GeneralDropoutComponent dim={cur_dim} specaugment-max-proportion=0.5")
                c.append(f"component-node name={sa}.freq-mask component={sa}.freq-mask input={cur}")
                c.append(f"component name={sa}.time-mask type=SpecAugmentTimeMaskComponent dim={cur_dim} zeroed-proportion=0.2 time-mask-max-frames=20")
                c.append(f"component-node name={sa}.time-mask component={sa}.time-mask input={sa}.freq-mask")
                cur = sa + ".time-mask"
        elif t == "delta":   # trivial_layers.py:236-256
            c.append(f"dim-range-node name={cur}_copy1 input-node={cur} dim={cur_dim} dim-offset=0")
            c.append(f"dim-range-node name={cur}_copy2 input-node={cur} dim={cur_dim} dim-offset=0")
            c.append(f"component name={cur}_2 type=NoOpComponent dim={3 * cur_dim}")
            c.append(f"component-node name={cur}_2 component={cur}_2 input=Append(Offset({cur},0),"
                     f" Sum(Offset(Scale(-1.0,{cur}_copy1),-1), Offset({cur},1)), Sum(Offset({cur},-2), Offset({cur},2),"
                     f" Offset(Scale(-2.0,{cur}_copy2),0)))")
            c.append(f"component name={n} type=BatchNormComponent dim={3 * cur_dim}")
            c.append(f"component-node name={n} component={n} input={cur}_2")
            cur, cur_dim = n, 3 * cur_dim
        elif t == "relu-batchnorm":
            sp = splice_of(L)
            spliced = ", ".join(cur if o == 0 else f"Offset({cur}, {o})" for o in sp)
            if L.get("append_ivector"):
                inp = f"Append({spliced}, Scale({L['append_ivector']}, ReplaceIndex(ivector, t, 0)))"
                k = len(sp) * cur_dim + ivd
            else:
                inp, k = (cur if sp == [0] else f"Append({spliced})"), len(sp) * cur_dim
            c.append(f"component name={n}.affine type=NaturalGradientAffineComponent input-dim={k} output-dim={L['dim']}")
            c.append(f"component-node name={n}.affine component={n}.affine input={inp}")
            c.append(f"component name={n}.relu type=RectifiedLinearComponent dim={L['dim']}")
            c.append(f"component-node name={n}.relu component={n}.relu input={n}.affine")
            c.append(f"component name={n}.batchnorm type=BatchNormComponent dim={L['dim']}")
            c.append(f"component-node name={n}.batchnorm component={n}.batchnorm input={n}.relu")
            cur, cur_dim = n + ".batchnorm", L["dim"]
            if extras:       # relu-batchnorm-dropout-layer: GeneralDropoutComponent, identity at test time
                c.append(f"component name={n}.dropout type=GeneralDropoutComponent dim={L['dim']} dropout-proportion=0.0 continuous=true")
                c.append(f"component-node name={n}.dropout component={n}.dropout input={n}.batchnorm")
                cur = n + ".dropout"
        elif t == "tdnnf":   # composite_layers.py:140-225
            s = L["stride"]
            o1 = f"{-s},0" if s else "0"
            o2 = f"0,{s}" if s else "0"
            c.append(f"component name={n}.linear type=TdnnComponent input-dim={cur_dim} output-dim={L['bottleneck']} "
                     f"use-bias=false time-offsets={o1} orthonormal-constraint=-1.0")
            c.append(f"component-node name={n}.linear component={n}.linear input={cur}")
            c.append(f"component name={n}.affine type=TdnnComponent input-dim={L['bottleneck']} output-dim={L['dim']} time-offsets={o2}")
            c.append(f"component-node name={n}.affine component={n}.affine input={n}.linear")
            c.append(f"component name={n}.relu type=RectifiedLinearComponent dim={L['dim']}")
            c.append(f"component-node name={n}.relu component={n}.relu input={n}.affine")
            c.append(f"component name={n}.batchnorm type=BatchNormComponent dim={L['dim']}")
            c.append(f"component-node name={n}.batchnorm component={n}.batchnorm input={n}.relu")
            last_bn = n + ".batchnorm"
            if extras:       # tdnnf-layer with dropout-proportion set: the dropout sits between batchnorm and the bypass sum
                c.append(f"component name={n}.dropout type=GeneralDropoutComponent dim={L['dim']} dropout-proportion=0.0 continuous=true")
                c.append(f"component-node name={n}.dropout component={n}.dropout input={n}.batchnorm")
                last_bn = n + ".dropout"
            if L["bypass"] != 0.0:
                c.append(f"component name={n}.noop type=NoOpComponent dim={L['dim']}")
                c.append(f"component-node name={n}.noop component={n}.noop input=Sum(Scale({L['bypass']}, {cur}), {last_bn})")
                cur, cur_dim = n + ".noop", L["dim"]
            else:                # bypass-scale=0.0: tdnnf-layer emits no NoOp (composite_layers.py:213-222)
                cur, cur_dim = last_bn, L["dim"]
        elif t == "linear":
            c.append(f"component name={n} type=LinearComponent input-dim={cur_dim} output-dim={L['dim']} orthonormal-constraint=-1.0")
            c.append(f"component-node name={n} component={n} input={cur}")
            cur, cur_dim = n, L["dim"]
        elif t == "prefinal":   # composite_layers.py:280-330
            xent_from = (cur, cur_dim)                   # the xent branch forks where the chain prefinal does
            c.append(f"component name={n}.affine type=NaturalGradientAffineComponent input-dim={cur_dim} output-dim={L['big']}")
            c.append(f"component-node name={n}.affine component={n}.affine input={cur}")
            c.append(f"component name={n}.relu type=RectifiedLinearComponent dim={L['big']}")
            c.append(f"component-node name={n}.relu component={n}.relu input={n}.affine")
            c.append(f"component name={n}.batchnorm1 type=BatchNormComponent dim={L['big']}")
            c.append(f"component-node name={n}.batchnorm1 component={n}.batchnorm1 input={n}.relu")
            c.append(f"component name={n}.linear type=LinearComponent input-dim={L['big']} output-dim={L['small']} orthonormal-constraint=-1")
            c.append(f"component-node name={n}.linear component={n}.linear input={n}.batchnorm1")
            c.append(f"component name={n}.batchnorm2 type=BatchNormComponent dim={L['small']}")
            c.append(f"component-node name={n}.batchnorm2 component={n}.batchnorm2 input={n}.linear")
            cur, cur_dim = n + ".batchnorm2", L["small"]
        elif t == "ivector-linear-bn":   # linear-component + batchnorm-component on ReplaceIndex(ivector, t, 0)
            c.append(f"component name={n}-linear type=LinearComponent input-dim={ivd} output-dim={L['dim']}")
            c.append(f"component-node name={n}-linear component={n}-linear input=ReplaceIndex(ivector, t, 0)")
            c.append(f"component name={n}-batchnorm type=BatchNormComponent dim={L['dim']} target-rms={L['target_rms']}")
            c.append(f"component-node name={n}-batchnorm component={n}-batchnorm input={n}-linear")
        elif t == "combine":             # trivial_layers.py:402-430
            h, f1, f2 = L["height"], L["filters1"], L["filters2"]
            cmap = []
            for hh in range(h):
                cmap += [hh * f1 + f for f in range(f1)] + [h * f1 + hh * f2 + f for f in range(f2)]
            c.append(f"component name={n} type=PermuteComponent column-map={','.join(map(str, cmap))}")
            c.append(f"component-node name={n} component={n} input=Append({cur}, {L['side']})")
            cur, cur_dim = n, h * (f1 + f2)
        elif t == "conv":                # convolution.py:260-310 (conv-relu-batchnorm-layer)
            od = L["height_out"] * L["filters_out"]
            c.append(f"component name={n}.conv type=TimeHeightConvolutionComponent height-in={L['height_in']} "
                     f"height-out={L['height_out']} height-subsample-out={L['height_subsample_out']} "
                     f"num-filters-in={L['filters_in']} num-filters-out={L['filters_out']} "
                     f"time-offsets={','.join(map(str, L['time_offsets']))} height-offsets={','.join(map(str, L['height_offsets']))}")
            c.append(f"component-node name={n}.conv component={n}.conv input={cur}")
            c.append(f"component name={n}.relu type=RectifiedLinearComponent dim={od} block-dim={L['filters_out']}")
            c.append(f"component-node name={n}.relu component={n}.relu input={n}.conv")
            c.append(f"component name={n}.batchnorm type=BatchNormComponent dim={od} block-dim={L['filters_out']}")
            c.append(f"component-node name={n}.batchnorm component={n}.batchnorm input={n}.relu")
            cur, cur_dim = n + ".batchnorm", od
        elif t == "output":
            if extras:       # the cross-entropy branch chain recipes train with and leave in final.mdl (run_tdnn_1k.sh:205-207)
                xin, xdim = (xent_from if xent_from else (cur, cur_dim))
                c.append(f"component name=prefinal-xent.affine type=NaturalGradientAffineComponent input-dim={xdim} output-dim=32")
                c.append(f"component-node name=prefinal-xent.affine component=prefinal-xent.affine input={xin}")
                c.append("component name=prefinal-xent.relu type=RectifiedLinearComponent dim=32")
                c.append("component-node name=prefinal-xent.relu component=prefinal-xent.relu input=prefinal-xent.affine")
                c.append("component name=prefinal-xent.batchnorm1 type=BatchNormComponent dim=32")
                c.append("component-node name=prefinal-xent.batchnorm1 component=prefinal-xent.batchnorm1 input=prefinal-xent.relu")
                c.append(f"component name=output-xent.affine type=NaturalGradientAffineComponent input-dim=32 output-dim={L['dim']}")
                c.append("component-node name=output-xent.affine component=output-xent.affine input=prefinal-xent.batchnorm1")
                c.append(f"component name=output-xent.log-softmax type=LogSoftmaxComponent dim={L['dim']}")
                c.append("component-node name=output-xent.log-softmax component=output-xent.log-softmax input=output-xent.affine")
                c.append("output-node name=output-xent input=output-xent.log-softmax objective=linear")
            c.append(f"component name={n}.affine type=NaturalGradientAffineComponent input-dim={cur_dim} output-dim={L['dim']}")
            c.append(f"component-node name={n}.affine component={n}.affine input={cur}")
            last = n + ".affine"
            if L.get("log_softmax"):
                c.append(f"component name={n}.log-softmax type=LogSoftmaxComponent dim={L['dim']}")
                c.append(f"component-node name={n}.log-softmax component={n}.log-softmax input={n}.affine")
                last = n + ".log-softmax"
            c.append(f"output-node name=output input={last}")
    return "\n".join(c) + "\n"


# ----------------------------------------------------------------------------- graph -> op program

BN_EPS = 1e-3      # BatchNormComponent default epsilon (nnet-normalize-component.h)


def bn_scale_offset(mean: np.ndarray, var: np.ndarray, target_rms: float = 1.0):
    """BatchNormComponent::ComputeDerived (nnet-normalize-component.cc:209-246):
    scale = (var + eps)^-0.5 * target_rms, offset = -mean * scale."""
    scale = np.power(np.maximum(var.astype(np.float32), 0.0) + np.float32(BN_EPS), np.float32(-0.5)).astype(np.float32)
    if target_rms != 1.0:
        scale = (scale * np.float32(target_rms)).astype(np.float32)
    offset = (-(mean.astype(np.float32)) * scale).astype(np.float32)
    return scale, offset


def expand_conv_weights(L: dict, w: np.ndarray, b: np.ndarray, combine: dict | None):
    """TimeHeightConvolutionComponent (nnet3/nnet-convolutional-component.cc:282-299, convolution.h:88-130) as a
    dense affine map per time offset: returns (W_exp [H_out*F_out, n_t * K_in], b_exp, K_main, K_side) with
      out[t, h*F_out + f] = b[f] + sum_{ti, dh, c} w[f, (ti*n_h + dhi)*F_in + c] * in[t + dt_ti, (h*sub + dh)*F_in + c]
    (inputs outside 0 <= h_in < H_in are zero padding).  With `combine` (combine-feature-maps-layer in front:
    the input map interleaves f1 filters of the main node with f2 of a side node) the columns of each time
    block are reordered to [main node columns | side node columns] so the permutation costs nothing.
    This is a first, functional mapping onto the GEMM kernel (it multiplies the structural zeros too:
    ~n_h/H_in density); a patch-gather convolution kernel is the follow-up."""
    Hi, Ho, sub = L["height_in"], L["height_out"], L["height_subsample_out"]
    Fi, Fo = L["filters_in"], L["filters_out"]
    toffs, hoffs = L["time_offsets"], L["height_offsets"]
    K = Hi * Fi
    if combine:
        f1, f2 = combine["filters1"], combine["filters2"]
        assert f1 + f2 == Fi and combine["height"] == Hi
        k_main, k_side = Hi * f1, Hi * f2
    else:
        f1, f2, k_main, k_side = Fi, 0, K, 0
    We = np.zeros((Ho * Fo, len(toffs) * K), np.float32)
    for ti in range(len(toffs)):
        for hi_, dh in enumerate(hoffs):
            blk = w[:, (ti * len(hoffs) + hi_) * Fi:(ti * len(hoffs) + hi_ + 1) * Fi]      # [Fo, Fi]
            for ho in range(Ho):
                h_in = ho * sub + dh
                if not (0 <= h_in < Hi):
                    continue
                rows = slice(ho * Fo, (ho + 1) * Fo)
                base = ti * K
                We[rows, base + h_in * f1: base + h_in * f1 + f1] = blk[:, :f1]
                if f2:
                    We[rows, base + k_main + h_in * f2: base + k_main + h_in * f2 + f2] = blk[:, f1:]
    return We, np.tile(b.astype(np.float32), Ho), k_main, k_side


@dataclass
class Node:
    name: str
    dim: int
    kind: str                 # "input", "ivector", "gemm", "ew"
    # consumers fill these during the backward range analysis
    residues: set = field(default_factory=set)
    tmin: int = 10**9
    tmax: int = -10**9
    step: int = 0
    t0: int = 0
    rows: int = 0
    spec: dict = field(default_factory=dict)


CONV_MODE_DEFAULT = "patch"      # time-height convolutions as hsplit GEMM ops (true convolution FLOPs); "dense": one dense
                                 # [H_out*F_out, H_in*F_in] map per time offset (kept for A/B checks and for the first layer,
                                 # whose combine-feature-maps permutation is folded into the weights)


def build_graph(arch: dict, W: dict, structural: bool = False, conv_mode: str | None = None):
    """Nodes in topological order.  Each non-input node has spec:
       gemm: terms=[(src, time_offset, w_cols(lo,hi), kind)], w, b, relu, bn=(scale,offset)|None,
             res=(src, alpha)|None, post=(sub_vec, mul)|None, log_softmax
       ew:   blocks=[[(src, time_offset, scale), ...] per column block], bn
    """
    fd, ivd = arch["feat_dim"], arch["ivector_dim"]
    nodes = [Node("input", fd, "input"), Node("ivector", ivd, "ivector")]
    cur = "input"
    dims = {"input": fd, "ivector": ivd}

    def add(n):
        nodes.append(n)
        dims[n.name] = n.dim

    pending_combine = None
    for L in arch["layers"]:
        t, n = L["type"], L["name"]
        if t == "idct":
            add(Node(n, fd, "gemm", spec=dict(terms=[(cur, 0, (0, fd), "row")], w=W[n + ".w"], b=W[n + ".b"])))
            cur = n
        elif t == "batchnorm":
            # fold into the producing node's epilogue (same arithmetic: x*scale + offset)
            prod = next(x for x in nodes if x.name == cur)
            prod.spec["bn"] = bn_scale_offset(W[n + ".mean"], W[n + ".var"])
        elif t == "delta":
            d = dims[cur]
            blocks = [[(cur, 0, 1.0)], [(cur, -1, -1.0), (cur, 1, 1.0)], [(cur, -2, 1.0), (cur, 2, 1.0), (cur, 0, -2.0)]]
            add(Node(n, 3 * d, "ew", spec=dict(blocks=blocks, block_dim=d, bn=bn_scale_offset(W[n + ".mean"], W[n + ".var"]))))
            cur = n
        elif t == "lda":
            d, sp = dims[cur], splice_of(L)
            k = len(sp) * d + ivd
            terms = [(cur, o, (i * d, (i + 1) * d), "row") for i, o in enumerate(sp)] + [("ivector", 0, (len(sp) * d, k), "ivec")]
            add(Node(n, k, "gemm", spec=dict(terms=terms, w=W[n + ".w"], b=W[n + ".b"])))
            cur = n
        elif t == "relu-batchnorm":
            d0, sp = dims[cur], splice_of(L)
            d = len(sp) * d0
            terms = [(cur, o, (i * d0, (i + 1) * d0), "row") for i, o in enumerate(sp)]
            if L.get("append_ivector"):
                # Scale(0.4, ReplaceIndex(ivector,t,0)) is a node of its own so the product keeps the
                # reference's association (0.4*iv rounded to f32, then the affine)
                add(Node(n + ".ivscaled", ivd, "ew", spec=dict(blocks=[[("ivector", 0, float(L["append_ivector"]))]],
                                                             block_dim=ivd, bn=None, ivector_rows=True)))
                terms.append((n + ".ivscaled", 0, (d, d + ivd), "ivec"))
            add(Node(n + ".batchnorm", L["dim"], "gemm",
                     spec=dict(terms=terms, w=W[n + ".affine.w"], b=W[n + ".affine.b"], relu=True,
                               bn=bn_scale_offset(W[n + ".batchnorm.mean"], W[n + ".batchnorm.var"]))))
            cur = n + ".batchnorm"
        elif t == "tdnnf":
            d, s, bt = dims[cur], L["stride"], L["bottleneck"]
            o1 = [-s, 0] if s else [0]
            o2 = [0, s] if s else [0]
            add(Node(n + ".linear", bt, "gemm",
                     spec=dict(terms=[(cur, o, (i * d, (i + 1) * d), "row") for i, o in enumerate(o1)],
                               w=W[n + ".linear.w"], b=None)))
            add(Node(n + ".noop", L["dim"], "gemm",
                     spec=dict(terms=[(n + ".linear", o, (i * bt, (i + 1) * bt), "row") for i, o in enumerate(o2)],
                               w=W[n + ".affine.w"], b=W[n + ".affine.b"], relu=True,
                               bn=bn_scale_offset(W[n + ".batchnorm.mean"], W[n + ".batchnorm.var"]),
                               res=(cur, float(L["bypass"])) if L["bypass"] != 0.0 else None)))
            cur = n + ".noop"
        elif t == "linear":
            d = dims[cur]
            add(Node(n, L["dim"], "gemm", spec=dict(terms=[(cur, 0, (0, d), "row")], w=W[n + ".w"], b=None)))
            cur = n
        elif t == "prefinal":
            d = dims[cur]
            add(Node(n + ".batchnorm1", L["big"], "gemm",
                     spec=dict(terms=[(cur, 0, (0, d), "row")], w=W[n + ".affine.w"], b=W[n + ".affine.b"], relu=True,
                               bn=bn_scale_offset(W[n + ".batchnorm1.mean"], W[n + ".batchnorm1.var"]))))
            add(Node(n + ".batchnorm2", L["small"], "gemm",
                     spec=dict(terms=[(n + ".batchnorm1", 0, (0, L["big"]), "row")], w=W[n + ".linear.w"], b=None,
                               bn=bn_scale_offset(W[n + ".batchnorm2.mean"], W[n + ".batchnorm2.var"]))))
            cur = n + ".batchnorm2"
        elif t == "ivector-linear-bn":
            # evaluated once per nnet chunk (the rows of the ivector input), like Scale(0.4, ReplaceIndex(ivector, t, 0))
            add(Node(n + "-batchnorm", L["dim"], "gemm",
                     spec=dict(terms=[("ivector", 0, (0, ivd), "chunk")], w=W[n + "-linear.w"], b=None, ivector_rows=True,
                               bn=bn_scale_offset(W[n + "-batchnorm.mean"], W[n + "-batchnorm.var"], L["target_rms"]))))
        elif t == "combine":
            pending_combine = dict(L)
        elif t == "conv":
            mode = conv_mode or CONV_MODE_DEFAULT
            if mode == "patch" and not pending_combine and not structural:
                # one GEMM row per (time, output height): 9 terms of klen = filters_in, each reading the source
                # row's column window of height h*subsample + dh (zeros outside = height padding)
                Fi, Fo, Hi, Ho = L["filters_in"], L["filters_out"], L["height_in"], L["height_out"]
                nh = len(L["height_offsets"])
                terms, cols = [], []
                for ti, dt in enumerate(L["time_offsets"]):
                    for hi_, dh in enumerate(L["height_offsets"]):
                        k0 = (ti * nh + hi_) * Fi
                        terms.append((cur, dt, (k0, k0 + Fi), "row"))
                        cols.append((L["height_subsample_out"] * Fi, dh * Fi, Hi * Fi))
                add(Node(n + ".batchnorm", Ho * Fo, "gemm",
                         spec=dict(terms=terms, term_cols=cols, hsplit=Ho, w=W[n + ".conv.w"], b=W[n + ".conv.b"], relu=True,
                                   bn=bn_scale_offset(W[n + ".batchnorm.mean"], W[n + ".batchnorm.var"]))))
                cur = n + ".batchnorm"
                continue
            if structural:       # context / range analysis only: no weights needed
                f2 = pending_combine["filters2"] if pending_combine else 0
                k_side = L["height_in"] * f2
                k_main = L["height_in"] * L["filters_in"] - k_side
                we, be = np.zeros((1, 1), np.float32), np.zeros(1, np.float32)
            else:
                we, be, k_main, k_side = expand_conv_weights(L, W[n + ".conv.w"], W[n + ".conv.b"], pending_combine)
            K = k_main + k_side
            terms = []
            for ti, dt in enumerate(L["time_offsets"]):
                terms.append((cur, dt, (ti * K, ti * K + k_main), "row"))
                if k_side:
                    terms.append((pending_combine["side"], dt, (ti * K + k_main, (ti + 1) * K), "ivec"))
            sc, of = bn_scale_offset(W[n + ".batchnorm.mean"], W[n + ".batchnorm.var"])
            add(Node(n + ".batchnorm", L["height_out"] * L["filters_out"], "gemm",
                     spec=dict(terms=terms, w=we, b=be, relu=True,
                               bn=(np.tile(sc, L["height_out"]), np.tile(of, L["height_out"])))))
            pending_combine = None
            cur = n + ".batchnorm"
        elif t == "output":
            d = dims[cur]
            add(Node("output", L["dim"], "gemm",
                     spec=dict(terms=[(cur, 0, (0, d), "row")], w=W[n + ".affine.w"], b=W[n + ".affine.b"],
                               log_softmax=bool(L.get("log_softmax")))))
            cur = "output"
    return nodes


def model_context(arch: dict):
    """(left_context, right_context) of the network = ComputeSimpleNnetContext
    (nnet3/nnet-utils.cc) for these architectures."""
    nodes = build_graph(arch, _zero_weights(arch), structural=True)
    lo = {n.name: 0 for n in nodes}
    hi = {n.name: 0 for n in nodes}
    order = {n.name: i for i, n in enumerate(nodes)}
    lo["output"], hi["output"] = 0, 0
    need_lo = {n.name: None for n in nodes}
    need_hi = {n.name: None for n in nodes}
    need_lo["output"], need_hi["output"] = 0, 0
    for n in reversed(nodes):
        if need_lo[n.name] is None:
            continue
        for (src, off) in _deps(n):
            a, b = need_lo[n.name] + off, need_hi[n.name] + off
            need_lo[src] = a if need_lo[src] is None else min(need_lo[src], a)
            need_hi[src] = b if need_hi[src] is None else max(need_hi[src], b)
    return -need_lo["input"], need_hi["input"]


def _deps(n: Node):
    if n.kind == "gemm":
        if n.spec.get("ivector_rows"):
            return []
        d = [(src, off) for (src, off, _c, kind) in n.spec["terms"] if kind == "row"]
        if n.spec.get("res"):
            d.append((n.spec["res"][0], 0))
        return d
    if n.kind == "ew":
        if n.spec.get("ivector_rows"):
            return []
        return [(src, off) for blk in n.spec["blocks"] for (src, off, _s) in blk]
    return []


def _zero_weights(arch):
    class Z(dict):
        def __missing__(self, k):
            return np.zeros((1, 1), np.float32) if k.endswith(".w") else np.ones(1, np.float32)
    return Z()


def compile_program(arch: dict, W: dict, num_frames: int, frames_per_chunk: int = 21,
                    acoustic_scale: float = 1.0, use_priors: bool = True, conv_mode: str | None = None) -> dict:
    """Compile for utterances of `num_frames` feature frames.  Returns a dict:
       nodes: [(name, dim, rows, t0, step)], ops: [op dicts], blob: float32 array,
       left/right context, num_out rows (= ceil(T / subsampling)), ivector chunk mapping.
    Row mapping of a term: src_row = clamp(out_row * ratio + shift, lo, hi)."""
    sub = arch["frame_subsampling_factor"]
    T = int(num_frames)
    n_out = (T + sub - 1) // sub
    nodes = build_graph(arch, W, conv_mode=conv_mode)
    by = {n.name: n for n in nodes}
    out = by["output"]
    out.residues, out.tmin, out.tmax = {0}, 0, sub * (n_out - 1)
    # backward pass: required time range and residues (mod sub) of every node
    for n in reversed(nodes):
        if n.tmax < n.tmin:
            continue
        for (src, off) in _deps(n):
            s = by[src]
            s.tmin = min(s.tmin, n.tmin + off)
            s.tmax = max(s.tmax, n.tmax + off)
            if len(n.residues) == 1 and n.kind != "input":
                r = next(iter(n.residues))
                s.residues.add((r + off) % sub)
            else:
                s.residues |= set(range(sub))
    L, R = -by["input"].tmin, by["input"].tmax - (sub * (n_out - 1))
    for n in nodes:
        if n.kind == "input":
            n.step, n.t0, n.rows = 1, 0, T            # raw features, reads are clamped
        elif n.kind == "ivector":
            n.step, n.t0, n.rows = 0, 0, 0            # set below
        elif n.tmax < n.tmin:
            n.rows = 0
        elif len(n.residues) == 1:
            r = next(iter(n.residues))
            n.step = sub
            n.t0 = n.tmin + ((r - n.tmin) % sub)
            n.rows = (n.tmax - n.t0) // sub + 1
        else:
            n.step, n.t0, n.rows = 1, n.tmin, n.tmax - n.tmin + 1
    # ivector rows: chunk n of the looped computation supplies one i-vector
    # (decodable-online-looped.cc:170-205); input time t uses the i-vector of chunk
    # max(0, floor(t / C) - m), m = floor((C + R - 1) / C)   (nnet-compile-looped.cc:179-205)
    C = frames_per_chunk
    assert C % sub == 0
    Lk, Rk = model_context(arch)
    m = (C + Rk - 1) // C
    n_chunks = (n_out * sub + C - 1) // C
    by["ivector"].rows = n_chunks
    blob = []
    blob_off = [0]

    def put(a):
        a = np.ascontiguousarray(a, np.float32).reshape(-1)
        off = blob_off[0]
        blob.append(a)
        blob_off[0] += a.size
        return off

    idx = {n.name: i for i, n in enumerate(nodes)}
    ops = []
    for n in nodes:
        if n.kind in ("input", "ivector"):
            continue
        if n.rows == 0 and not n.spec.get("ivector_rows"):
            continue
        if n.kind == "gemm" and n.spec.get("ivector_rows"):
            n.step, n.t0, n.rows = 0, 0, n_chunks
            sp = n.spec
            bn = sp.get("bn")
            ops.append(dict(type="gemm", out=idx[n.name], rows=n_chunks, N=sp["w"].shape[0], K=sp["w"].shape[1],
                            terms=[dict(src=idx["ivector"], ratio=1, shift=0, lo=0, hi=n_chunks - 1, ivec=0, k0=c0, klen=c1 - c0)
                                   for (_src, _off, (c0, c1), _kind) in sp["terms"]],
                            w=put(sp["w"]), bias=put(sp["b"]) if sp.get("b") is not None else -1, relu=int(bool(sp.get("relu"))),
                            bn_scale=put(bn[0]) if bn else -1, bn_offset=put(bn[1]) if bn else -1,
                            res=None, res_alpha=0.0, sub_vec=-1, out_scale=1.0, log_softmax=0))
            continue
        if n.kind == "ew" and n.spec.get("ivector_rows"):
            n.step, n.t0, n.rows = 0, 0, n_chunks
            ops.append(dict(type="ew", out=idx[n.name], rows=n_chunks, block_dim=n.spec["block_dim"],
                            blocks=[[dict(src=idx["ivector"], ratio=1, shift=0, lo=0, hi=n_chunks - 1, scale=s, ivec=0)
                                     for (src, off, s) in blk] for blk in n.spec["blocks"]],
                            bn_scale=-1, bn_offset=-1))
            continue

        def rowmap(src, off, kind):
            s = by[src]
            if kind == "ivec":
                # src rows are chunks: row = clamp((t0 + i*step + off) // C - m, 0, n_chunks-1)
                return dict(src=idx[src], ratio=n.step, shift=n.t0 + off, lo=0, hi=s.rows - 1, ivec=1, C=C, m=m)
            if s.kind == "input":
                return dict(src=idx[src], ratio=n.step, shift=n.t0 + off, lo=0, hi=T - 1, ivec=0)
            assert n.step % s.step == 0 and (n.t0 + off - s.t0) % s.step == 0, (n.name, src)
            return dict(src=idx[src], ratio=n.step // s.step, shift=(n.t0 + off - s.t0) // s.step, lo=0,
                        hi=s.rows - 1, ivec=0)

        if n.kind == "ew":
            blocks = []
            for blk in n.spec["blocks"]:
                blocks.append([dict(rowmap(src, off, "row"), scale=float(sc)) for (src, off, sc) in blk])
            bn = n.spec.get("bn")
            ops.append(dict(type="ew", out=idx[n.name], rows=n.rows, block_dim=n.spec["block_dim"], blocks=blocks,
                            bn_scale=put(bn[0]) if bn else -1, bn_offset=put(bn[1]) if bn else -1))
        else:
            sp = n.spec
            w = sp["w"]
            terms = []
            for ti_, (src, off, (c0, c1), kind) in enumerate(sp["terms"]):
                td = dict(rowmap(src, off, kind), k0=c0, klen=c1 - c0)
                if sp.get("term_cols"):
                    td["col_step"], td["col_off"], td["col_lim"] = sp["term_cols"][ti_]
                terms.append(td)
            bn = sp.get("bn")
            res = sp.get("res")
            op = dict(type="gemm", out=idx[n.name], rows=n.rows, N=w.shape[0], K=w.shape[1], terms=terms,
                      w=put(w), bias=put(sp["b"]) if sp.get("b") is not None else -1, relu=int(bool(sp.get("relu"))),
                      bn_scale=put(bn[0]) if bn else -1, bn_offset=put(bn[1]) if bn else -1,
                      res=None, res_alpha=0.0, sub_vec=-1, out_scale=1.0, log_softmax=int(bool(sp.get("log_softmax"))),
                      hsplit=int(sp.get("hsplit", 1)))
            if res:
                op["res"] = rowmap(res[0], 0, "row")
                op["res_alpha"] = res[1]
            if n.name == "output":
                # AddVecToRows(-1, log_priors); Scale(acoustic_scale)  (decodable-online-looped.cc:218-223)
                if use_priors:
                    op["sub_vec"] = put(np.log(W["priors"]).astype(np.float32))
                op["out_scale"] = float(acoustic_scale)
            ops.append(op)
    # per-utterance arena with liveness-based reuse (first fit)
    last_use = {}
    for oi, op in enumerate(ops):
        srcs = [t["src"] for t in (op["terms"] if op["type"] == "gemm" else [x for b in op["blocks"] for x in b])]
        if op.get("res"):
            srcs.append(op["res"]["src"])
        for sidx in srcs:
            last_use[sidx] = oi
    arena_off = {}
    live = []           # (offset, size, node_idx)
    arena_size = 0
    for oi, op in enumerate(ops):
        o = op["out"]
        nd = nodes[o]
        if nd.name != "output":
            size = ((nd.rows * nd.dim + 31) // 32) * 32
            live.sort()
            pos = 0
            for (a, sz, _) in live:
                if pos + size <= a:
                    break
                pos = max(pos, a + sz)
            arena_off[o] = pos
            live.append((pos, size, o))
            arena_size = max(arena_size, pos + size)
        # free nodes whose last use is this op (after allocating the output)
        live = [(a, sz, ni) for (a, sz, ni) in live if last_use.get(ni, -1) > oi or ni == o]
    return dict(arch=arch, T=T, n_out=n_out, left_context=L, right_context=R, model_left=Lk, model_right=Rk,
                arena_off=arena_off, arena_size=arena_size,
                frames_per_chunk=C, ivector_m=m, n_chunks=n_chunks,
                nodes=[(n.name, n.dim, n.rows, n.t0, n.step) for n in nodes], ops=ops,
                blob=np.concatenate(blob) if blob else np.zeros(1, np.float32),
                node_index=idx)


def flops_per_output_frame(prog: dict) -> float:
    """2 * sum_l K_l * N_l * rows_l / n_out (algorithmic FLOPs counted once)."""
    tot = 0.0
    for op in prog["ops"]:
        if op["type"] == "gemm":
            tot += 2.0 * op["K"] * op["N"] * op["rows"] * op.get("hsplit", 1)
    return tot / prog["n_out"]


def load_kaldi_raw(path: str, priors: np.ndarray | None = None, name: str | None = None,
                   frame_subsampling_factor: int | None = None) -> tuple[dict, dict]:
    """(arch, weights) of a TDNN-F chain model stored as a raw nnet3 file (nnet3-am-copy --raw=true final.mdl
    final.raw; text or binary).  `priors` (AmNnetSimple::Priors, the <Priors> vector of final.mdl) default to
    none, which is what chain models ship."""
    from . import kaldi_io as KIO
    arch, W = KIO.nnet3_to_arch(KIO.read_nnet3_raw(path), name=name or os.path.basename(path),
                                frame_subsampling_factor=frame_subsampling_factor)
    W["priors"] = (np.ones(arch["num_pdfs"], np.float32) if priors is None else np.ascontiguousarray(priors, np.float32))
    return arch, W


def load_kaldi_mdl(path: str, name: str | None = None, frame_subsampling_factor: int | None = None) -> tuple[dict, dict, np.ndarray]:
    """(arch, weights, tid2pdf) from a final.mdl (TransitionModel + AmNnetSimple).  tid2pdf[tid] is what
    CudaFst applies to the ilabels of HCLG (cudadecoder/cuda-fst.cc: TransitionIdToPdf); an empty <Priors>
    vector (chain models) means no prior subtraction."""
    from . import kaldi_io as KIO
    m = KIO.read_final_mdl(path)
    arch, W = KIO.nnet3_to_arch(m["nnet"], name=name or os.path.basename(path), frame_subsampling_factor=frame_subsampling_factor)
    pri = m["priors"]
    W["priors"] = np.ones(arch["num_pdfs"], np.float32) if len(pri) == 0 else np.ascontiguousarray(pri, np.float32)
    return arch, W, m["transition_model"]["tid2pdf"]
