// ivector.cu — B200-native online i-vector extraction (sm_100a).
//
// Semantics: OnlineIvectorFeature (online2/online-ivector-feature.cc) with
// use_most_recent_ivector=true, as driven chunk by chunk by
// DecodableNnetLoopedOnlineBase::AdvanceChunk (decodable-online-looped.cc:170-205):
//   front kernel (one warp per frame):
//     OnlineCmvn'd and raw base features -> OnlineSpliceFrames(+-3, edge clamp)
//     (online-feature.cc:504-519) -> OnlineTransform y = b + A x (:538-554), both
//     paths; DiagGmm::LogLikelihoods (gmm/diag-gmm.cc:546-562) on the normalised
//     path; VectorToPosteriorEntry (hmm/posterior.cc:440-505): min_post cut,
//     top num_gselect, tail pruning, renormalise; x posterior_scale.
//   stats + CG kernel (one CTA per utterance, chunks in order):
//     OnlineIvectorEstimationStats::AccStats (ivector/ivector-extractor.cc:611-668)
//     in double: linear += Sigma_inv_M_g^T (w x), quadratic += w U_g, max_count
//     prior scaling; GetIvector (:732-756) = LinearCgd<double> (matrix/
//     optimization.cc:453-565) warm-started from the previous chunk's solution,
//     <= num_cg_iters iterations with the reference's residual-recompute rule.
// This replaces cudafeat's batched i-vector path (full softmax
// without gselect pruning + Cholesky in float), which is not comparable with the
// CPU reference; here the pruned posteriors and the double-precision CG follow
// the CPU code.  FP64 is confined to the 100-dim statistics (hard part 5).

#include <vector>

#include "common.cuh"
#include "feat_kernels.cuh"

namespace b2k {

struct IvecParams {
  int base_dim, splice_left, splice_right, feat_dim, num_gauss, ivector_dim, num_gselect;
  float min_post, posterior_scale, max_count, prior_offset;
  int num_cg_iters;
  const float *lda;            // [feat_dim x (spliced_dim+1)]
  const float *gconsts;        // [G]
  const float *means_invvars;  // [G x feat_dim]
  const float *inv_vars;       // [feat_dim x G] (transposed at creation, like means_invvars)
  const double *sigma_inv_m;   // [G x feat_dim x ivector_dim]
  const double *U;             // [G x ivd*(ivd+1)/2] packed lower triangle (row-major rows of growing length)
};

struct IvecRun {
  const float *const *d_feats;   // raw base features per lane [T x base_dim], stride feat_stride
  int feat_stride, T;
  float *cmvn;                   // [lanes x T x base_dim]
  float *lda_raw;                // [lanes x T x feat_dim]
  int *post_idx;                 // [lanes x T x 8]
  float *post_val;               // [lanes x T x 8]
  int *post_cnt;                 // [lanes x T]
  const int *sched; int n_chunks;
  float *const *d_out; int out_stride;
  // speaker adaptation state per lane (null arrays / null entries: a new speaker, nothing kept): the i-vector half is
  // {num_frames, linear[D], quadratic[D(D+1)/2]} after the CMVN half (2*(base_dim+1) doubles); in and out may be the same array
  const double *const *state_in; double *const *state_out;
  int state_iv_off;              // doubles in front of the i-vector half
  float max_remembered_frames;
};

#define IV_WARPS 8
#define IV_FR 4

__global__ void __launch_bounds__(IV_WARPS * 32) ivec_front_kernel(IvecParams p, IvecRun r, int frames_per_cta) {
  extern __shared__ float sm[];
  const int SD = p.base_dim * (p.splice_left + p.splice_right + 1);
  const int LW = SD + 1;                    // lda row length
  float *s_lda = sm;                        // feat_dim * LW
  float *s_w = s_lda + p.feat_dim * LW;     // per warp: 2*SD + feat_dim
  const int per_warp = 2 * SD + IV_FR * p.feat_dim;
  const int tid = threadIdx.x, lane_id = tid & 31, warp = tid >> 5;
  for (int i = tid; i < p.feat_dim * LW; i += blockDim.x) s_lda[i] = p.lda[i];
  __syncthreads();
  float *spl_raw = s_w + warp * per_warp, *spl_norm = spl_raw + SD, *xa0 = spl_norm + SD;   // xa0: IV_FR frames x feat_dim
  const int L = blockIdx.y;
  const float *feats = r.d_feats[L];
  const float *cm = r.cmvn + (size_t)L * r.T * p.base_dim;
  const int f0 = blockIdx.x * frames_per_cta, f1 = min(f0 + frames_per_cta, r.T);
  const int W = p.splice_left + p.splice_right + 1;
  (void)W;
  // IV_FR frames per warp pass: the UBM parameters (164 KB for 512 Gaussians x 40 dims, more than L1) are loaded once per
  // pass and applied to all of them, so the L2 -> SM traffic of the kernel is 1 / IV_FR of one-frame-per-pass (97 GB per
  // 592-utterance step, the kernel sat at L2 bandwidth: 15.6 ms).  Every (Gaussian, frame) sum keeps its order over d.
  for (int t0 = f0 + warp * IV_FR; t0 < f1; t0 += IV_WARPS * IV_FR) {
    const int nfr = min(IV_FR, f1 - t0);
    for (int f = 0; f < nfr; f++) {
      const int t = t0 + f;
      float *xa = xa0 + f * p.feat_dim;
      // splice with edge clamping (online-feature.cc:504-519)
      for (int k = lane_id; k < SD; k += 32) {
        int w = k / p.base_dim, d = k - w * p.base_dim;
        int t2 = t - p.splice_left + w;
        t2 = min(max(t2, 0), r.T - 1);
        spl_raw[k] = feats[(size_t)t2 * r.feat_stride + d];
        spl_norm[k] = cm[(size_t)t2 * p.base_dim + d];
      }
      __syncwarp();
      // OnlineTransform: y = offset + A x, both paths
      float *lraw = r.lda_raw + ((size_t)L * r.T + t) * p.feat_dim;
      for (int j = lane_id; j < p.feat_dim; j += 32) {
        const float *row = s_lda + j * LW;
        float a0 = 0.f, a1 = 0.f;
        for (int k = 0; k < SD; k++) { float w = row[k]; a0 = fmaf(w, spl_raw[k], a0); a1 = fmaf(w, spl_norm[k], a1); }
        lraw[j] = row[SD] + a0;
        xa[j] = row[SD] + a1;
      }
      __syncwarp();
    }
    // DiagGmm log-likelihoods of the normalised path, IV_FR frames against one load of the parameters
    float llf[IV_FR][16];
#pragma unroll
    for (int i = 0; i < 16; i++) {
      const int g = lane_id + 32 * i;
      float a1[IV_FR], a2[IV_FR];
#pragma unroll
      for (int f = 0; f < IV_FR; f++) { a1[f] = 0.f; a2[f] = 0.f; }
      if (g < p.num_gauss) {
        const float *mv = p.means_invvars + g, *iv = p.inv_vars + g;      // [feat_dim][G]
        for (int d = 0; d < p.feat_dim; d++) {
          const float m = __ldg(&mv[(size_t)d * p.num_gauss]), v = __ldg(&iv[(size_t)d * p.num_gauss]);
#pragma unroll
          for (int f = 0; f < IV_FR; f++) {
            const float x = xa0[f * p.feat_dim + d];          // (frames beyond nfr: stale values, results unused)
            a1[f] = fmaf(m, x, a1[f]);
            a2[f] = fmaf(v, x * x, a2[f]);
          }
        }
        const float gc = __ldg(&p.gconsts[g]);
#pragma unroll
        for (int f = 0; f < IV_FR; f++) { float v = gc + a1[f]; llf[f][i] = v + (-0.5f) * a2[f]; }
      } else {
#pragma unroll
        for (int f = 0; f < IV_FR; f++) llf[f][i] = -INFINITY;
      }
    }
#pragma unroll
    for (int f = 0; f < IV_FR; f++) {
      if (f >= nfr) break;
      const int t = t0 + f;
      float (&ll)[16] = llf[f];
      float best = -INFINITY;
#pragma unroll
      for (int i = 0; i < 16; i++) best = fmaxf(best, ll[i]);
      for (int o = 16; o > 0; o >>= 1) best = fmaxf(best, __shfl_xor_sync(0xffffffffu, best, o));
      // VectorToPosteriorEntry: candidates like > max + log(min_post), post = exp(like - max)
      const float cutoff = best + logf(p.min_post);
      float post[16];
      int ncand = 0;
  #pragma unroll
      for (int i = 0; i < 16; i++) {
        bool c = (p.min_post != 0.0f) && (ll[i] > cutoff);
        post[i] = c ? expf(ll[i] - best) : -1.0f;
        ncand += c;
      }
      ncand = __reduce_add_sync(0xffffffffu, ncand);
      if (ncand == 0) {     // none reached the threshold (or min_post == 0): take them all (:467-473)
  #pragma unroll
        for (int i = 0; i < 16; i++) post[i] = (lane_id + 32 * i < p.num_gauss) ? expf(ll[i] - best) : -1.0f;
      }
      // top num_gselect by posterior (descending)
      float sel_v[8]; int sel_i[8]; int nsel = 0;
      for (int s = 0; s < p.num_gselect && s < 8; s++) {
        float bv = -1.0f; int bi = -1;
  #pragma unroll
        for (int i = 0; i < 16; i++) if (post[i] > bv) { bv = post[i]; bi = lane_id + 32 * i; }
        // warp arg-max (ties -> lowest index)
        for (int o = 16; o > 0; o >>= 1) {
          float ov = __shfl_xor_sync(0xffffffffu, bv, o);
          int oi = __shfl_xor_sync(0xffffffffu, bi, o);
          if (ov > bv || (ov == bv && oi >= 0 && (bi < 0 || oi < bi))) { bv = ov; bi = oi; }
        }
        if (bi < 0 || bv < 0.f) break;
        sel_v[nsel] = bv; sel_i[nsel] = bi; nsel++;
        if ((bi & 31) == lane_id) {
          int slot = bi >> 5;
  #pragma unroll
          for (int i = 0; i < 16; i++) if (i == slot) post[i] = -1.0f;
        }
      }
      // tail pruning and renormalisation (:492-503)
      float tot = 0.f;
      for (int s = 0; s < nsel; s++) tot += sel_v[s];
      const float cut2 = p.min_post * tot;
      while (nsel > 1 && sel_v[nsel - 1] < cut2) { tot -= sel_v[nsel - 1]; nsel--; }
      const float inv_tot = 1.0f / tot;
      if (lane_id == 0) {
        size_t o = ((size_t)L * r.T + t) * 8;
        const float sc = p.posterior_scale * 1.0f;           // posterior_scale * frame weight (1.0)
        for (int s = 0; s < nsel; s++) { r.post_idx[o + s] = sel_i[s]; r.post_val[o + s] = (sel_v[s] * inv_tot) * sc; }
        r.post_cnt[(size_t)L * r.T + t] = nsel;
      }
      __syncwarp();
    }
  }
}

// ---- stats accumulation + CG, one CTA per utterance
__device__ __forceinline__ double block_sum_d(double v, double *sh, int nthreads) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  double t = 0.0;
  for (int i = 0; i < nthreads / 32; i++) t += sh[i];
  return t;
}

__device__ __forceinline__ double sp_at(const double *q, int i, int j) {   // packed lower triangle
  return (i >= j) ? q[(size_t)i * (i + 1) / 2 + j] : q[(size_t)j * (j + 1) / 2 + i];
}

// (A v)_i of a packed symmetric matrix, FOUR threads per row (row = tid / 4, each sums a contiguous quarter of the
// columns, the quarters are added as (q0 + q1) + (q2 + q3)): the dependent chain of double FMAs is D / 4 long instead of
// D.  Every thread of the warps that hold rows gets the row's value; rows >= D return 0.  4 * D <= the CTA width.
__device__ __forceinline__ double sp_matvec_row4(const double *q, const double *v, int D) {
  const int i = threadIdx.x >> 2, part = threadIdx.x & 3;
  double a = 0.0;
  if (i < D) {
    const int per = (D + 3) >> 2, j0 = part * per, j1 = min(D, j0 + per);
    for (int j = j0; j < j1; j++) a += sp_at(q, i, j) * v[j];
  }
  a += __shfl_xor_sync(0xffffffffu, a, 1);
  a += __shfl_xor_sync(0xffffffffu, a, 2);
  return a;
}

#define IVS_THREADS 512
#define IVS_STAGE 64

__global__ void __launch_bounds__(IVS_THREADS, 3) ivec_stats_cg_kernel(IvecParams p, IvecRun r) {
  extern __shared__ double sd[];
  const int D = p.ivector_dim, Q = D * (D + 1) / 2, F = p.feat_dim;
  double *quad = sd;            // Q
  double *lin = quad + Q;       // D
  double *x = lin + D, *rr = x + D, *pp = rr + D, *Ap = pp + D;   // D each
  double *xf = Ap + D;          // 8 * F: weighted_feats of up to 8 Gaussians of the chunk
  double *red = xf + 8 * F;     // 32
  double *sa = red + 32;        // 8 * D: Sigma_inv_M_g^T weighted_feats of those Gaussians
  float *occ_s = reinterpret_cast<float *>(sa + 8 * D);        // num_gauss: a Gaussian's summed frame weights (float, as GaussInfo::tot_weight)
  int *glist = reinterpret_cast<int *>(occ_s + p.num_gauss);   // num_gauss: the chunk's distinct Gaussians
  int *wcnt_s = glist + p.num_gauss;                           // IVS_THREADS / 32
  int *nd_s = wcnt_s + IVS_THREADS / 32;                       // 4
  // a chunk of up to IVS_STAGE frames is staged once: posteriors and the LDA-transformed frames
  int *pcnt_s = nd_s + 4;                                      // IVS_STAGE
  int *pidx_s = pcnt_s + IVS_STAGE;                            // 8 * IVS_STAGE
  float *pval_s = reinterpret_cast<float *>(pidx_s + 8 * IVS_STAGE);   // 8 * IVS_STAGE
  float *feat_s = pval_s + 8 * IVS_STAGE;                      // IVS_STAGE * F
  const int tid = threadIdx.x, L = blockIdx.x;
  const double *st_in = (r.state_in && r.state_in[L]) ? r.state_in[L] + r.state_iv_off : nullptr;
  double num_frames = 0.0;
  if (st_in) {
    // SetAdaptationState (online-ivector-feature.cc:445-453): ivector_stats_ = the speaker's; the i-vector itself starts
    // from its default again (a new OnlineIvectorFeature object, :438-440)
    num_frames = st_in[0];
    for (int i = tid; i < D; i += IVS_THREADS) { lin[i] = st_in[1 + i]; x[i] = 0.0; }
    for (int k = tid; k < Q; k += IVS_THREADS) quad[k] = st_in[1 + D + k];
  } else {
    // OnlineIvectorEstimationStats ctor (:786-795): linear(0) = prior_offset, quadratic = I
    for (int k = tid; k < Q; k += IVS_THREADS) quad[k] = 0.0;
    for (int i = tid; i < D; i += IVS_THREADS) { lin[i] = 0.0; x[i] = 0.0; }
    __syncthreads();
    for (int i = tid; i < D; i += IVS_THREADS) quad[(size_t)i * (i + 1) / 2 + i] = 1.0;
    if (tid == 0) lin[0] = (double)p.prior_offset;
  }
  __syncthreads();
  int next_t = 0;
  const float *lda_raw = r.lda_raw + (size_t)L * r.T * F;
  for (int n = 0; n < r.n_chunks; n++) {
    if (r.sched[n] < 0) {          // no i-vector frame was ready when this chunk ran: all zeros (decodable-online-looped.cc:188-197)
      float *z = r.d_out[L] + (size_t)n * r.out_stride;
      for (int i = tid; i < D; i += IVS_THREADS) z[i] = 0.0f;
      continue;
    }
    const int upto = min(r.sched[n], r.T - 1);
    double tot_weight = 0.0;
    const bool any = next_t <= upto;
    if (any) {
      // OnlineIvectorEstimationStats::AccStats(extractor, features, gauss_post) (ivector-extractor.cc:611-668) over the
      // frames of this request, as the reference does it: PER GAUSSIAN.  ConvertPostToGaussInfo sums a Gaussian's frame
      // weights in float, in frame order; weighted_feats = sum_t weight * x_t in double; then ONE
      // linear += Sigma_inv_M_g^T weighted_feats and ONE quadratic += tot_weight_g * U_g per distinct Gaussian, so a
      // 40 KB U_g row and a 32 KB Sigma_inv_M_g block are read once per chunk instead of once per frame that selects g
      // (round 1 read them per frame: 213 GB of L2 traffic per 592-utterance step, the kernel sat at L2 bandwidth).
      const int t0 = next_t;
      const int G = p.num_gauss;
      const int nfr = upto - t0 + 1;
      // the chunk's posteriors and frames: out of shared memory when the chunk fits (every thread scans them several
      // times), else where the front kernel left them (same code: generic pointers, indexed from the chunk's first frame)
      const int *cntp = r.post_cnt + (size_t)L * r.T + t0;
      const int *idxp = r.post_idx + ((size_t)L * r.T + t0) * 8;
      const float *valp = r.post_val + ((size_t)L * r.T + t0) * 8;
      const float *featp = lda_raw + (size_t)t0 * F;
      if (nfr <= IVS_STAGE) {
        for (int q = tid; q < nfr; q += IVS_THREADS) pcnt_s[q] = cntp[q];
        for (int q = tid; q < nfr * 8; q += IVS_THREADS) { pidx_s[q] = idxp[q]; pval_s[q] = valp[q]; }
        for (int q = tid; q < nfr * F; q += IVS_THREADS) feat_s[q] = featp[q];
        cntp = pcnt_s; idxp = pidx_s; valp = pval_s; featp = feat_s;
        __syncthreads();
      }
      for (int g = tid; g < G; g += IVS_THREADS) {
        float o = 0.f;
        for (int t = 0; t < nfr; t++) {
          const int cnt = cntp[t];
          for (int j = 0; j < cnt; j++)
            if (idxp[t * 8 + j] == g) o += valp[t * 8 + j];
        }
        occ_s[g] = o;
      }
      if (tid == 0) *nd_s = 0;
      __syncthreads();
      // distinct Gaussians of the chunk, in index order (one pass: G <= 512 = the CTA width)
      for (int gb = 0; gb < G; gb += IVS_THREADS) {
        const int g = gb + tid;
        const bool on = g < G && occ_s[g] != 0.f;
        const unsigned m = __ballot_sync(0xffffffffu, on);
        __syncthreads();
        if ((tid & 31) == 0) wcnt_s[tid >> 5] = __popc(m);
        __syncthreads();
        int base = *nd_s;
        for (int w = 0; w < (tid >> 5); w++) base += wcnt_s[w];
        if (on) glist[base + __popc(m & ((1u << (tid & 31)) - 1u))] = g;
        __syncthreads();
        if (tid == 0) { int tot = 0; for (int w = 0; w < IVS_THREADS / 32; w++) tot += wcnt_s[w]; *nd_s += tot; }
        __syncthreads();
      }
      const int nd = *nd_s;
      for (int gb = 0; gb < nd; gb += 8) {
        const int nj = min(8, nd - gb);
        // weighted_feats of up to 8 Gaussians side by side: thread (j, d)
        for (int q = tid; q < nj * F; q += IVS_THREADS) {
          const int j = q / F, d = q - j * F, g = glist[gb + j];
          double a = 0.0;
          for (int t = 0; t < nfr; t++) {
            const int cnt = cntp[t];
            for (int jj = 0; jj < cnt; jj++)
              if (idxp[t * 8 + jj] == g) a += (double)valp[t * 8 + jj] * (double)featp[(size_t)t * F + d];
          }
          xf[q] = a;
        }
        __syncthreads();
        for (int q = tid; q < nj * D; q += IVS_THREADS) {
          const int j = q / D, i = q - j * D;
          const double *SiM = p.sigma_inv_m + (size_t)glist[gb + j] * F * D;
          double a = 0.0;
          for (int d = 0; d < F; d++) a += SiM[(size_t)d * D + i] * xf[j * F + d];
          sa[q] = a;
        }
        __syncthreads();
        int gj[8];
        double wj[8];
#pragma unroll
        for (int j = 0; j < 8; j++) {
          gj[j] = (j < nj) ? glist[gb + j] : 0;
          wj[j] = (j < nj) ? (double)occ_s[gj[j]] : 0.0;
        }
        for (int i = tid; i < D; i += IVS_THREADS) {
          double l = lin[i];
#pragma unroll
          for (int j = 0; j < 8; j++)
            if (j < nj) l += sa[j * D + i];
          lin[i] = l;
        }
        for (int k = tid; k < Q; k += IVS_THREADS) {
          double qv = quad[k];
#pragma unroll
          for (int j = 0; j < 8; j++)
            if (j < nj) qv += wj[j] * __ldg(&p.U[(size_t)gj[j] * Q + k]);
          quad[k] = qv;
        }
#pragma unroll
        for (int j = 0; j < 8; j++)
          if (j < nj) tot_weight += wj[j];
        __syncthreads();
      }
      next_t = upto + 1;
    }
    if (any) {
      // max_count prior scaling (:650-664)
      if (p.max_count > 0.0f) {
        double mc = (double)p.max_count;
        double old_n = num_frames, new_n = num_frames + tot_weight;
        double old_s = fmax(old_n, mc) / mc, new_s = fmax(new_n, mc) / mc;
        double ch = new_s - old_s;
        if (ch != 0.0) {
          if (tid == 0) lin[0] += (double)p.prior_offset * ch;
          for (int i = tid; i < D; i += IVS_THREADS) quad[(size_t)i * (i + 1) / 2 + i] += ch;
        }
      }
      num_frames += tot_weight;
      __syncthreads();
      // ---- GetIvector (:732-756) -> LinearCgd<double> (optimization.cc:453-565)
      if (num_frames > 0.0) {
        if (tid == 0 && x[0] == 0.0) x[0] = (double)p.prior_offset;
        __syncthreads();
        // p_0 = b - A x_0 ; r_0 = -p_0
        {
          const double a = sp_matvec_row4(quad, x, D);
          if ((tid & 3) == 0 && (tid >> 2) < D) { pp[tid >> 2] = lin[tid >> 2] - a; rr[tid >> 2] = -pp[tid >> 2]; }
        }
        __syncthreads();
        double loc = 0.0;
        for (int i = tid; i < D; i += IVS_THREADS) loc += rr[i] * rr[i];
        double r_cur = block_sum_d(loc, red, IVS_THREADS);
        double r_recompute = r_cur;
        const double max_error_sq = 2.2250738585072014e-308;          // max(0*0, DBL_MIN)
        const double residual_factor = 0.01 * 0.01, inv_residual_factor = 1.0 / residual_factor;
        for (int k = 0; k < D + 5 && k != p.num_cg_iters; k++) {
          {
            const double a = sp_matvec_row4(quad, pp, D);
            if ((tid & 3) == 0 && (tid >> 2) < D) Ap[tid >> 2] = a;
          }
          __syncthreads();
          double l1 = 0.0, l2 = 0.0;
          for (int i = tid; i < D; i += IVS_THREADS) { l1 += pp[i] * rr[i]; l2 += pp[i] * Ap[i]; }
          double pr = block_sum_d(l1, red, IVS_THREADS);
          double pAp = block_sum_d(l2, red, IVS_THREADS);
          double alpha = -pr / pAp;
          for (int i = tid; i < D; i += IVS_THREADS) { x[i] += alpha * pp[i]; rr[i] += alpha * Ap[i]; }
          __syncthreads();
          loc = 0.0;
          for (int i = tid; i < D; i += IVS_THREADS) loc += rr[i] * rr[i];
          double r_next = block_sum_d(loc, red, IVS_THREADS);
          if (r_next < residual_factor * r_recompute || r_next > inv_residual_factor * r_recompute) {
            {
              const double a = sp_matvec_row4(quad, x, D);
              if ((tid & 3) == 0 && (tid >> 2) < D) Ap[tid >> 2] = a - lin[tid >> 2];
            }
            __syncthreads();
            for (int i = tid; i < D; i += IVS_THREADS) rr[i] = Ap[i];
            __syncthreads();
            loc = 0.0;
            for (int i = tid; i < D; i += IVS_THREADS) loc += rr[i] * rr[i];
            r_next = block_sum_d(loc, red, IVS_THREADS);
            r_recompute = r_next;
          }
          if (r_next <= max_error_sq) break;
          double beta = r_next / r_cur;
          for (int i = tid; i < D; i += IVS_THREADS) pp[i] = pp[i] * beta - rr[i];
          r_cur = r_next;
          __syncthreads();
        }
      } else {
        for (int i = tid; i < D; i += IVS_THREADS) x[i] = (i == 0) ? (double)p.prior_offset : 0.0;
      }
      __syncthreads();
    }
    // GetFrame (:343-349): float copy, then subtract the prior offset from dim 0
    float *o = r.d_out[L] + (size_t)n * r.out_stride;
    for (int i = tid; i < D; i += IVS_THREADS) {
      float v = (float)x[i];
      if (i == 0) v = v - p.prior_offset;
      o[i] = v;
    }
    __syncthreads();
  }
  if (r.state_out && r.state_out[L]) {
    // GetAdaptationState (:386-396): the stats as they stand, then LimitFrames' i-vector half (:119-126) =
    // OnlineIvectorEstimationStats::Scale (ivector-extractor.cc:671-693)
    double *so = r.state_out[L] + r.state_iv_off;
    double scale = 1.0, diag_add = 0.0, nf = num_frames;
    const float limit = r.max_remembered_frames * p.posterior_scale;            // BaseFloat product
    if (r.max_remembered_frames >= 0.0f && num_frames > (double)limit) {
      scale = (double)limit / num_frames;
      nf = num_frames * scale;
      if (p.max_count == 0.0f) diag_add = 1.0 - scale;
      else {
        const double mc = (double)p.max_count;
        diag_add = fmax(nf, mc) / mc - scale * fmax(num_frames, mc) / mc;
      }
    }
    for (int i = tid; i < D; i += IVS_THREADS) so[1 + i] = lin[i] * scale + (i == 0 ? (double)p.prior_offset * diag_add : 0.0);
    for (int k = tid; k < Q; k += IVS_THREADS) so[1 + D + k] = quad[k] * scale;
    __syncthreads();
    for (int i = tid; i < D; i += IVS_THREADS) so[1 + D + (size_t)i * (i + 1) / 2 + i] += diag_add;
    if (tid == 0) so[0] = nf;
  }
}

}  // namespace b2k

using namespace b2k;

struct b2k_ivec {
  b2k_ivec_cfg cfg;
  IvecParams p;
  std::vector<void *> allocs;
  float *d_cmvn = nullptr, *d_lda_raw = nullptr, *d_post_val = nullptr;
  int *d_post_idx = nullptr, *d_post_cnt = nullptr, *d_sched = nullptr;
  double *d_cmvn_state = nullptr, *d_global = nullptr;
  CmvnLane *d_clanes = nullptr, *h_clanes = nullptr;
  double **d_stp = nullptr, **h_stp = nullptr;     // [2 * max_lanes]: state in pointers, then state out pointers
  const float **d_featp = nullptr, **h_featp = nullptr;
  float **d_outp = nullptr, **h_outp = nullptr;
  int *h_sched = nullptr;
  cudaEvent_t staging_free = nullptr;
  size_t smem_front = 0, smem_stats = 0;
};

extern "C" {

int b2k_ivec_create(const b2k_ivec_cfg *cfg, const float *lda, const float *gconsts, const float *means_invvars,
                    const float *inv_vars, const double *sigma_inv_m, const double *U,
                    const double *global_cmvn_stats, b2k_ivec **out) {
  if (!cfg || !lda || !gconsts || !means_invvars || !inv_vars || !sigma_inv_m || !U || !global_cmvn_stats || !out)
    return set_error(B2K_ERR_INVALID, "b2k_ivec_create: bad args");
  if (cfg->num_gauss < 1 || cfg->num_gauss > 512 || cfg->ivector_dim < 1 || cfg->ivector_dim > 128 ||
      cfg->feat_dim > 64 || cfg->num_gselect < 1 || cfg->num_gselect > 8 || !(cfg->min_post >= 0.f && cfg->min_post < 1.f))
    return set_error(B2K_ERR_INVALID, "b2k_ivec_create: unsupported sizes (num_gauss<=512, ivector_dim<=128, feat_dim<=64, num_gselect<=8)");
  int rc = require_device();
  if (rc) return rc;
  b2k_ivec *iv = new b2k_ivec();
  iv->cfg = *cfg;
  IvecParams &p = iv->p;
  memset(&p, 0, sizeof(p));
  const int SD = cfg->base_dim * (cfg->splice_left + cfg->splice_right + 1), F = cfg->feat_dim, G = cfg->num_gauss,
            D = cfg->ivector_dim, Q = D * (D + 1) / 2;
  auto up = [&](const void *h, size_t bytes, const void **d) -> int {
    void *ptr = nullptr;
    B2K_CUDA_CHECK(cudaMalloc(&ptr, bytes));
    iv->allocs.push_back(ptr);
    B2K_CUDA_CHECK(cudaMemcpy(ptr, h, bytes, cudaMemcpyHostToDevice));
    *d = ptr;
    return 0;
  };
  if ((rc = up(lda, 4 * (size_t)F * (SD + 1), (const void **)&p.lda))) return rc;
  if ((rc = up(gconsts, 4 * (size_t)G, (const void **)&p.gconsts))) return rc;
  {
    // the posterior kernel reads Gaussian g = lane + 32 i: keep the UBM parameters [feat_dim][G]
    // so that a warp's 32 Gaussians are one coalesced row segment
    std::vector<float> mvT((size_t)G * F), ivT((size_t)G * F);
    for (int g = 0; g < G; g++)
      for (int d = 0; d < F; d++) {
        mvT[(size_t)d * G + g] = means_invvars[(size_t)g * F + d];
        ivT[(size_t)d * G + g] = inv_vars[(size_t)g * F + d];
      }
    if ((rc = up(mvT.data(), 4 * (size_t)G * F, (const void **)&p.means_invvars))) return rc;
    if ((rc = up(ivT.data(), 4 * (size_t)G * F, (const void **)&p.inv_vars))) return rc;
  }
  if ((rc = up(sigma_inv_m, 8 * (size_t)G * F * D, (const void **)&p.sigma_inv_m))) return rc;
  if ((rc = up(U, 8 * (size_t)G * Q, (const void **)&p.U))) return rc;
  if ((rc = up(global_cmvn_stats, 8 * 2 * (size_t)(cfg->base_dim + 1), (const void **)&iv->d_global))) return rc;
  p.base_dim = cfg->base_dim; p.splice_left = cfg->splice_left; p.splice_right = cfg->splice_right; p.feat_dim = F;
  p.num_gauss = G; p.ivector_dim = D; p.num_gselect = cfg->num_gselect; p.min_post = cfg->min_post;
  p.posterior_scale = cfg->posterior_scale; p.max_count = cfg->max_count; p.prior_offset = cfg->prior_offset;
  p.num_cg_iters = cfg->num_cg_iters;
  const size_t NL = cfg->max_lanes, T = cfg->max_frames;
  auto al = [&](void **ptr, size_t bytes) -> int { B2K_CUDA_CHECK(cudaMalloc(ptr, bytes)); iv->allocs.push_back(*ptr); return 0; };
  if ((rc = al((void **)&iv->d_cmvn, 4 * NL * T * cfg->base_dim))) return rc;
  if ((rc = al((void **)&iv->d_lda_raw, 4 * NL * T * F))) return rc;
  if ((rc = al((void **)&iv->d_post_idx, 4 * NL * T * 8))) return rc;
  if ((rc = al((void **)&iv->d_post_val, 4 * NL * T * 8))) return rc;
  if ((rc = al((void **)&iv->d_post_cnt, 4 * NL * T))) return rc;
  if ((rc = al((void **)&iv->d_cmvn_state, 8 * NL * 2 * (cfg->base_dim + 1)))) return rc;
  if ((rc = al((void **)&iv->d_sched, 4 * 4096))) return rc;
  if ((rc = al((void **)&iv->d_clanes, sizeof(CmvnLane) * NL))) return rc;
  if ((rc = al((void **)&iv->d_featp, sizeof(void *) * NL))) return rc;
  if ((rc = al((void **)&iv->d_outp, sizeof(void *) * NL))) return rc;
  B2K_CUDA_CHECK(cudaMallocHost((void **)&iv->h_clanes, sizeof(CmvnLane) * NL));
  if ((rc = al((void **)&iv->d_stp, sizeof(double *) * 2 * NL))) return rc;
  B2K_CUDA_CHECK(cudaMallocHost((void **)&iv->h_stp, sizeof(double *) * 2 * NL));
  B2K_CUDA_CHECK(cudaMallocHost((void **)&iv->h_featp, sizeof(void *) * NL));
  B2K_CUDA_CHECK(cudaMallocHost((void **)&iv->h_outp, sizeof(void *) * NL));
  B2K_CUDA_CHECK(cudaMallocHost((void **)&iv->h_sched, 4 * 4096));
  B2K_CUDA_CHECK(cudaEventCreateWithFlags(&iv->staging_free, cudaEventDisableTiming));
  iv->smem_front = sizeof(float) * ((size_t)F * (SD + 1) + IV_WARPS * (2 * (size_t)SD + IV_FR * (size_t)F));
  iv->smem_stats = sizeof(double) * ((size_t)Q + 5 * D + 8 * (size_t)F + 32 + 8 * (size_t)D) + 4 * (2 * (size_t)G + IVS_THREADS / 32 + 4) +
                   4 * ((size_t)IVS_STAGE * (1 + 8 + 8) + (size_t)IVS_STAGE * F);
  B2K_CUDA_CHECK(cudaFuncSetAttribute(ivec_front_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)iv->smem_front));
  B2K_CUDA_CHECK(cudaFuncSetAttribute(ivec_stats_cg_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)iv->smem_stats));
  *out = iv;
  return B2K_OK;
}

int b2k_ivec_destroy(b2k_ivec *iv) {
  if (!iv) return B2K_OK;
  cudaDeviceSynchronize();
  for (void *p : iv->allocs) cudaFree(p);
  cudaFreeHost(iv->h_clanes); cudaFreeHost(iv->h_featp); cudaFreeHost(iv->h_outp); cudaFreeHost(iv->h_sched); cudaFreeHost(iv->h_stp);
  if (iv->staging_free) cudaEventDestroy(iv->staging_free);
  delete iv;
  return B2K_OK;
}

int64_t b2k_ivec_adaptation_state_doubles(const b2k_ivec *iv) {
  if (!iv) return -1;
  const int64_t D = iv->cfg.ivector_dim;
  return 2 * ((int64_t)iv->cfg.base_dim + 1) + 1 + D + D * (D + 1) / 2;
}

int b2k_ivec_compute_batched(b2k_ivec *iv, int32_t num_lanes, const float *const *d_feats, int32_t feat_stride,
                             int32_t num_frames, const int32_t *sched, int32_t n_chunks, float *const *d_out,
                             int32_t out_stride, void *stream) {
  return b2k_ivec_compute_batched_adapt(iv, num_lanes, d_feats, feat_stride, num_frames, sched, n_chunks, d_out, out_stride, nullptr,
                                        nullptr, -1.0f, stream);
}

int b2k_ivec_compute_batched_adapt(b2k_ivec *iv, int32_t num_lanes, const float *const *d_feats, int32_t feat_stride,
                                   int32_t num_frames, const int32_t *sched, int32_t n_chunks, float *const *d_out,
                                   int32_t out_stride, const double *const *d_state_in, double *const *d_state_out,
                                   float max_remembered_frames, void *stream) {
  if (!iv || num_lanes <= 0 || num_lanes > iv->cfg.max_lanes || !d_feats || !sched || !d_out || n_chunks <= 0 ||
      n_chunks > 4096 || num_frames <= 0 || num_frames > iv->cfg.max_frames)
    return set_error(B2K_ERR_INVALID, "b2k_ivec_compute_batched: bad args");
  cudaStream_t st = (cudaStream_t)stream;
  B2K_CUDA_CHECK(cudaEventSynchronize(iv->staging_free));
  const int bd = iv->cfg.base_dim;
  for (int i = 0; i < num_lanes; i++) {
    CmvnLane &L = iv->h_clanes[i];
    L.in = d_feats[i]; L.out = iv->d_cmvn + (size_t)i * num_frames * bd; L.in_stride = feat_stride; L.out_stride = bd;
    L.first_frame = 0; L.num_frames = num_frames; L.state = iv->d_cmvn_state + (size_t)i * 2 * (bd + 1);
    // the CMVN half of a speaker's adaptation state lies in front of the i-vector half
    L.speaker = d_state_in ? d_state_in[i] : nullptr;
    L.speaker_out = d_state_out ? d_state_out[i] : nullptr;
    L.max_remembered_frames = max_remembered_frames;
    iv->h_stp[i] = d_state_in ? const_cast<double *>(d_state_in[i]) : nullptr;
    iv->h_stp[iv->cfg.max_lanes + i] = d_state_out ? d_state_out[i] : nullptr;
    iv->h_featp[i] = d_feats[i]; iv->h_outp[i] = d_out[i];
  }
  for (int n = 0; n < n_chunks; n++) {
    if (sched[n] < -1 || (n > 0 && sched[n] < sched[n - 1])) return set_error(B2K_ERR_INVALID, "schedule must be non-decreasing (-1 = no i-vector ready yet)");
    iv->h_sched[n] = sched[n];
  }
  B2K_CUDA_CHECK(cudaMemcpyAsync(iv->d_clanes, iv->h_clanes, sizeof(CmvnLane) * num_lanes, cudaMemcpyHostToDevice, st));
  B2K_CUDA_CHECK(cudaMemcpyAsync((void *)iv->d_featp, iv->h_featp, sizeof(void *) * num_lanes, cudaMemcpyHostToDevice, st));
  B2K_CUDA_CHECK(cudaMemcpyAsync((void *)iv->d_outp, iv->h_outp, sizeof(void *) * num_lanes, cudaMemcpyHostToDevice, st));
  B2K_CUDA_CHECK(cudaMemcpyAsync(iv->d_sched, iv->h_sched, 4 * (size_t)n_chunks, cudaMemcpyHostToDevice, st));
  B2K_CUDA_CHECK(cudaMemcpyAsync((void *)iv->d_stp, iv->h_stp, sizeof(double *) * 2 * (size_t)iv->cfg.max_lanes, cudaMemcpyHostToDevice, st));
  B2K_CUDA_CHECK(cudaEventRecord(iv->staging_free, st));
  B2K_CUDA_CHECK(cudaMemsetAsync(iv->d_cmvn_state, 0, 8 * (size_t)num_lanes * 2 * (bd + 1), st));
  CmvnParams cp;
  cp.dim = bd; cp.cmn_window = iv->cfg.cmn_window; cp.speaker_frames = iv->cfg.speaker_frames;
  cp.global_frames = iv->cfg.global_frames; cp.normalize_mean = 1; cp.normalize_variance = 0;
  cp.global_stats = iv->d_global; cp.speaker_stats = nullptr;
  int rc = launch_cmvn(cp, iv->d_clanes, num_lanes, st);
  if (rc) return rc;
  IvecRun r;
  r.d_feats = iv->d_featp; r.feat_stride = feat_stride; r.T = num_frames; r.cmvn = iv->d_cmvn; r.lda_raw = iv->d_lda_raw;
  r.post_idx = iv->d_post_idx; r.post_val = iv->d_post_val; r.post_cnt = iv->d_post_cnt; r.sched = iv->d_sched;
  r.n_chunks = n_chunks; r.d_out = iv->d_outp; r.out_stride = out_stride;
  r.state_in = d_state_in ? iv->d_stp : nullptr; r.state_out = d_state_out ? iv->d_stp + iv->cfg.max_lanes : nullptr;
  r.state_iv_off = 2 * (bd + 1); r.max_remembered_frames = max_remembered_frames;
  const int fpc = 64;
  ivec_front_kernel<<<dim3((num_frames + fpc - 1) / fpc, num_lanes), IV_WARPS * 32, iv->smem_front, st>>>(iv->p, r, fpc);
  B2K_LAUNCH_CHECK();
  ivec_stats_cg_kernel<<<num_lanes, IVS_THREADS, iv->smem_stats, st>>>(iv->p, r);
  B2K_LAUNCH_CHECK();
  return B2K_OK;
}

}  // extern "C"
