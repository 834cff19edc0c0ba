// nnet.cu — B200-native batched nnet3 forward for the TDNN-F family (sm_100a).
//
// Executes the flat op program produced by kaldi_b200/nnet_model.py
// (compile_program): every op is either
//   GEMM   out[r,:] = epilogue( sum_terms A_term[map_term(r), :] . W[:, k0:k0+klen]^T )
//          = TdnnComponent / AffineComponent / LinearComponent / FixedAffineComponent
//            ::Propagate (nnet3/nnet-tdnn-component.cc:181-211, nnet-simple-component.cc
//            :1242,3224,3392) with the time-offset splicing expressed as per-term row
//            maps (no Append/Offset copies: descriptors nnet3/nnet-descriptor.h become
//            address arithmetic), and the epilogue fusing bias, ReLU (:964-972),
//            test-mode BatchNorm y = x*scale + offset (nnet-normalize-component.cc
//            :455-466), the TDNN-F bypass Sum(Scale(0.66, in), bn), and the decodable's
//            output post-processing (-log prior, x acoustic_scale,
//            decodable-online-looped.cc:218-223);
//   EW     out[:, block] = sum_terms scale * src[map(r), :]  (+ BatchNorm): the delta layer
//          and Scale(0.4, ReplaceIndex(ivector, t, 0)).
// This replaces the reference's NnetComputer interpreter loop over ~300-400
// commands per chunk (nnet3/nnet-compute.cc:236-459) and its three-pass
// ReLU/BatchNorm kernels with ~35 launches per batch.
//
// Numerics: float32 inputs; the GEMM runs on the tensor cores as 3xTF32 (hi/lo
// operand split, fp32 accumulation: nnet_gemm_tc_kernel below), which keeps the
// log-likelihoods within ~1e-6 relative of the CPU reference (north star: 1e-4).
// The fp32 FFMA kernel (nnet_gemm_kernel) is kept for A/B checks (B2K_NNET_GEMM=simt).
// Round 2 replaces mma.sync by tcgen05 kind::tf32.

#include <cstdlib>
#include <cstring>
#include <vector>

#include "common.cuh"

namespace b2k {

struct TermDev {
  int src_kind;          // 0 internal node, 1 external input (features), 2 external ivector
  long long src_off;     // internal: float offset inside the per-lane arena
  int src_dim;           // row stride of src (floats)
  int ratio, shift, lo, hi, ivec, C, m;
  int k0, klen;
  float scale;
  int col_step, col_off, col_lim;   // hsplit ops: source column window per output height (col_lim 0 = plain)
};

struct OpDev {
  int type;              // 0 gemm, 1 ew
  long long out_off; int out_dim; int out_kind;   // out_kind 0 internal, 3 external output
  int rows, N, K;
  int n_terms; TermDev terms[12];
  const float *w, *bias, *bn_scale, *bn_offset, *sub_vec;
  int relu, has_res; TermDev res; float res_alpha, out_scale;
  int block_dim; int term_block[12];
  int log_softmax;       // output post-processing is then applied after the log-softmax kernel
  int hsplit;            // > 1: time-height convolution, hsplit GEMM rows (output heights) per node row
};

struct RunCtx {
  float *arena; long long arena_stride;          // per-lane internal storage
  const float *const *d_input; int in_stride;    // per-lane features
  const float *const *d_ivec; int iv_stride;     // per-lane chunk i-vectors
  float *const *d_out; int out_stride;           // per-lane outputs
  int batch;
};

__device__ __forceinline__ int map_row(const TermDev &t, int i) {
  int j = i * t.ratio + t.shift;
  if (t.ivec) {
    // floor division for possibly negative times, then chunk lag m (nnet-compile-looped.cc:179-205)
    int q = (j >= 0) ? (j / t.C) : -((-j + t.C - 1) / t.C);
    j = q - t.m;
  }
  return min(max(j, t.lo), t.hi);
}

// GEMM row r -> (utterance lane, node row, output height) and the term's source pointer for it
struct RowIdx { int lane, i, h; };
__device__ __forceinline__ RowIdx split_row(const OpDev &op, int r) {
  const int H = op.hsplit > 1 ? op.hsplit : 1;
  const int per_lane = op.rows * H;
  RowIdx x;
  x.lane = r / per_lane;
  const int rem = r - x.lane * per_lane;
  x.i = rem / H;
  x.h = rem - x.i * H;
  return x;
}

__device__ __forceinline__ const float *src_row_ptr(const RunCtx &c, const TermDev &t, int lane, int row) {
  if (t.src_kind == 0) return c.arena + (long long)lane * c.arena_stride + t.src_off + (long long)row * t.src_dim;
  if (t.src_kind == 1) return c.d_input[lane] + (long long)row * c.in_stride;
  return c.d_ivec[lane] + (long long)row * c.iv_stride;
}

#define GM_BM 64
#define GM_BN 64
#define GM_BK 16

// C tile 64x64, 256 threads, 4x4 micro-tile per thread
__global__ void __launch_bounds__(256) nnet_gemm_kernel(OpDev op, RunCtx c) {
  __shared__ float As[GM_BK][GM_BM + 4];
  __shared__ float Bs[GM_BK][GM_BN + 4];
  __shared__ const float *rowp[GM_BM];
  const int tid = threadIdx.x;
  const int M = c.batch * op.rows * (op.hsplit > 1 ? op.hsplit : 1);
  const int m0 = blockIdx.x * GM_BM, n0 = blockIdx.y * GM_BN;    // x = row tiles (can exceed 65535 with hsplit)
  const int tx = tid & 15, ty = tid >> 4;         // 16 x 16 thread grid
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; i++)
#pragma unroll
    for (int j = 0; j < 4; j++) acc[i][j] = 0.f;

  for (int ti = 0; ti < op.n_terms; ti++) {
    const TermDev t = op.terms[ti];
    __syncthreads();
    if (tid < GM_BM) {
      int r = m0 + tid;
      const float *p = nullptr;
      if (r < M) {
        const RowIdx x = split_row(op, r);
        p = src_row_ptr(c, t, x.lane, map_row(t, x.i));
        if (t.col_lim > 0) {                              // convolution patch: column window of this output height
          const int cb = x.h * t.col_step + t.col_off;
          p = (cb >= 0 && cb < t.col_lim) ? p + cb : nullptr;   // outside = height zero padding
        }
      }
      rowp[tid] = p;
    }
    __syncthreads();
    for (int kk = 0; kk < t.klen; kk += GM_BK) {
      // A tile: 64 rows x 16 k  -> 1024 elements, 4 per thread
#pragma unroll
      for (int e = 0; e < 4; e++) {
        int idx = tid + e * 256;
        int r = idx >> 4, k = idx & 15;
        const float *p = rowp[r];
        float v = 0.f;
        if (p && kk + k < t.klen) v = p[kk + k];
        As[k][r] = v;
      }
      // B tile: W[n0 + n][k0 + kk + k]
#pragma unroll
      for (int e = 0; e < 4; e++) {
        int idx = tid + e * 256;
        int n = idx >> 4, k = idx & 15;
        float v = 0.f;
        if (n0 + n < op.N && kk + k < t.klen) v = __ldg(&op.w[(long long)(n0 + n) * op.K + t.k0 + kk + k]);
        Bs[k][n] = v;
      }
      __syncthreads();
#pragma unroll
      for (int k = 0; k < GM_BK; k++) {
        float a[4], b[4];
#pragma unroll
        for (int i = 0; i < 4; i++) a[i] = As[k][ty * 4 + i];
#pragma unroll
        for (int j = 0; j < 4; j++) b[j] = Bs[k][tx * 4 + j];
#pragma unroll
        for (int i = 0; i < 4; i++)
#pragma unroll
          for (int j = 0; j < 4; j++) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
      }
      __syncthreads();
    }
  }
  // ---- epilogue
#pragma unroll
  for (int i = 0; i < 4; i++) {
    int r = m0 + ty * 4 + i;
    if (r >= M) continue;
    const RowIdx x = split_row(op, r);
    const int lane = x.lane, ri = x.i;
    float *orow = (op.out_kind == 0)
                      ? c.arena + (long long)lane * c.arena_stride + op.out_off + (long long)ri * op.out_dim + (long long)x.h * op.N
                      : c.d_out[lane] + (long long)ri * c.out_stride;
    const float *rrow = nullptr;
    if (op.has_res) rrow = src_row_ptr(c, op.res, lane, map_row(op.res, ri));
#pragma unroll
    for (int j = 0; j < 4; j++) {
      int n = n0 + tx * 4 + j;
      if (n >= op.N) continue;
      float v = acc[i][j];
      if (op.bias) v = __fadd_rn(v, __ldg(&op.bias[n]));
      if (op.relu) v = fmaxf(v, 0.f);
      if (op.bn_scale) v = __fadd_rn(__fmul_rn(v, __ldg(&op.bn_scale[n])), __ldg(&op.bn_offset[n]));
      if (rrow) v = __fadd_rn(__fmul_rn(op.res_alpha, rrow[n]), v);
      if (!op.log_softmax) {
        if (op.sub_vec) v = __fadd_rn(v, -__ldg(&op.sub_vec[n]));
        if (op.out_scale != 1.0f) v = __fmul_rn(v, op.out_scale);
      }
      orow[n] = v;
    }
  }
}

// ---------------------------------------------------------------- tensor-core GEMM (3xTF32)
//
// Same operator as nnet_gemm_kernel with the inner product on the tensor cores.
// Every fp32 operand x is split when it is staged in shared memory into
// hi = tf32(x) and lo = tf32(x - hi); the product is accumulated in fp32 as
// a_lo*b_hi + a_hi*b_lo + a_hi*b_hi (three m16n8k8 TF32 MMAs, small terms
// first), which keeps the log-likelihoods within ~1e-6 relative of the fp32
// CPU reference (the dropped a_lo*b_lo term is 2^-22 relative) -- plain TF32
// (one MMA) would be ~1e-3 and miss the 1e-4 north-star tolerance.
// CTA tile 64 x (32*NT) x 32 with NT = 4 or 3 (N = 96 / 192 layers), 8 warps of
// 32 x (8*NT), operands padded to a stride of 36 floats so that the fragment
// loads are bank-conflict free.
#define TC_BM 64
#define TC_BK 32
#define TC_LD 36

__device__ __forceinline__ uint32_t to_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return r;
}

__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

template <int NT>
__global__ void __launch_bounds__(256, 2) nnet_gemm_tc_kernel(OpDev op, RunCtx c) {
  constexpr int TC_BN = 32 * NT;
  extern __shared__ __align__(16) unsigned char tc_smem[];
  uint32_t *As_hi = reinterpret_cast<uint32_t *>(tc_smem);          // [TC_BM][TC_LD]
  uint32_t *As_lo = As_hi + TC_BM * TC_LD;
  uint32_t *Bs_hi = As_lo + TC_BM * TC_LD;                          // [TC_BN][TC_LD]
  uint32_t *Bs_lo = Bs_hi + TC_BN * TC_LD;
  __shared__ const float *rowp[TC_BM];
  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane_id = tid & 31;
  const int g = lane_id >> 2, t4 = lane_id & 3;
  const int wm = warp & 1, wn = warp >> 1;                          // 2 x 4 warps
  const int M = c.batch * op.rows * (op.hsplit > 1 ? op.hsplit : 1);
  const int m0 = blockIdx.x * TC_BM, n0 = blockIdx.y * TC_BN;    // x = row tiles (can exceed 65535 with hsplit)
  float acc[2][NT][4];
#pragma unroll
  for (int i = 0; i < 2; i++)
#pragma unroll
    for (int j = 0; j < NT; j++)
#pragma unroll
      for (int q = 0; q < 4; q++) acc[i][j][q] = 0.f;

  // staging: thread -> (row = idx >> 5, k = idx & 31), 8 A elements and 16 B elements per slab
  float ra[8], rb[TC_BN / 8];
  for (int ti = 0; ti < op.n_terms; ti++) {
    const TermDev t = op.terms[ti];
    __syncthreads();
    if (tid < TC_BM) {
      int r = m0 + tid;
      const float *p = nullptr;
      if (r < M) {
        const RowIdx x = split_row(op, r);
        p = src_row_ptr(c, t, x.lane, map_row(t, x.i));
        if (t.col_lim > 0) {                              // convolution patch: column window of this output height
          const int cb = x.h * t.col_step + t.col_off;
          p = (cb >= 0 && cb < t.col_lim) ? p + cb : nullptr;   // outside = height zero padding
        }
      }
      rowp[tid] = p;
    }
    __syncthreads();
    auto fetch = [&](int kk) {
      const int k = kk + lane_id;
      const bool kin = k < t.klen;
#pragma unroll
      for (int e = 0; e < 8; e++) {
        const float *p = rowp[warp + e * 8];
        ra[e] = (p && kin) ? p[k] : 0.f;
      }
#pragma unroll
      for (int e = 0; e < TC_BN / 8; e++) {
        const int n = n0 + warp + e * 8;
        rb[e] = (n < op.N && kin) ? __ldg(&op.w[(long long)n * op.K + t.k0 + k]) : 0.f;
      }
    };
    fetch(0);
    for (int kk = 0; kk < t.klen; kk += TC_BK) {
      // registers -> shared (split)
#pragma unroll
      for (int e = 0; e < 8; e++) {
        const uint32_t hi = to_tf32(ra[e]);
        As_hi[(warp + e * 8) * TC_LD + lane_id] = hi;
        As_lo[(warp + e * 8) * TC_LD + lane_id] = to_tf32(ra[e] - __uint_as_float(hi));
      }
#pragma unroll
      for (int e = 0; e < TC_BN / 8; e++) {
        const uint32_t hi = to_tf32(rb[e]);
        Bs_hi[(warp + e * 8) * TC_LD + lane_id] = hi;
        Bs_lo[(warp + e * 8) * TC_LD + lane_id] = to_tf32(rb[e] - __uint_as_float(hi));
      }
      __syncthreads();
      if (kk + TC_BK < t.klen) fetch(kk + TC_BK);                   // next slab in flight during the MMAs
#pragma unroll
      for (int ks = 0; ks < TC_BK; ks += 8) {
        uint32_t ah[2][4], al[2][4], bh[NT][2], bl[NT][2];
#pragma unroll
        for (int mi = 0; mi < 2; mi++) {
          const int r = wm * 32 + mi * 16 + g;
          const int o0 = r * TC_LD + ks + t4, o1 = (r + 8) * TC_LD + ks + t4;
          ah[mi][0] = As_hi[o0]; ah[mi][1] = As_hi[o1]; ah[mi][2] = As_hi[o0 + 4]; ah[mi][3] = As_hi[o1 + 4];
          al[mi][0] = As_lo[o0]; al[mi][1] = As_lo[o1]; al[mi][2] = As_lo[o0 + 4]; al[mi][3] = As_lo[o1 + 4];
        }
#pragma unroll
        for (int ni = 0; ni < NT; ni++) {
          const int o = (wn * 8 * NT + ni * 8 + g) * TC_LD + ks + t4;
          bh[ni][0] = Bs_hi[o]; bh[ni][1] = Bs_hi[o + 4];
          bl[ni][0] = Bs_lo[o]; bl[ni][1] = Bs_lo[o + 4];
        }
#pragma unroll
        for (int mi = 0; mi < 2; mi++)
#pragma unroll
          for (int ni = 0; ni < NT; ni++) {
            mma_tf32(acc[mi][ni], al[mi], bh[ni]);
            mma_tf32(acc[mi][ni], ah[mi], bl[ni]);
            mma_tf32(acc[mi][ni], ah[mi], bh[ni]);
          }
      }
      __syncthreads();
    }
  }
  // ---- epilogue (same fused chain as the SIMT kernel); c0,c1 -> row g, c2,c3 -> row g + 8
#pragma unroll
  for (int mi = 0; mi < 2; mi++)
#pragma unroll
    for (int half = 0; half < 2; half++) {
      const int r = m0 + wm * 32 + mi * 16 + g + half * 8;
      if (r >= M) continue;
      const RowIdx x = split_row(op, r);
      const int lane = x.lane, ri = x.i;
      float *orow = (op.out_kind == 0)
                        ? c.arena + (long long)lane * c.arena_stride + op.out_off + (long long)ri * op.out_dim + (long long)x.h * op.N
                        : c.d_out[lane] + (long long)ri * c.out_stride;
      const float *rrow = nullptr;
      if (op.has_res) rrow = src_row_ptr(c, op.res, lane, map_row(op.res, ri));
#pragma unroll
      for (int ni = 0; ni < NT; ni++)
#pragma unroll
        for (int q = 0; q < 2; q++) {
          const int n = n0 + wn * 8 * NT + ni * 8 + 2 * t4 + q;
          if (n >= op.N) continue;
          float v = acc[mi][ni][half * 2 + q];
          if (op.bias) v = __fadd_rn(v, __ldg(&op.bias[n]));
          if (op.relu) v = fmaxf(v, 0.f);
          if (op.bn_scale) v = __fadd_rn(__fmul_rn(v, __ldg(&op.bn_scale[n])), __ldg(&op.bn_offset[n]));
          if (rrow) v = __fadd_rn(__fmul_rn(op.res_alpha, rrow[n]), v);
          if (!op.log_softmax) {
            if (op.sub_vec) v = __fadd_rn(v, -__ldg(&op.sub_vec[n]));
            if (op.out_scale != 1.0f) v = __fmul_rn(v, op.out_scale);
          }
          orow[n] = v;
        }
    }
}

// ---------------------------------------------------------------- tcgen05 GEMM (3xTF32, accumulator in TMEM)
//
// EXPERIMENTAL, OFF BY DEFAULT (B2K_NNET_GEMM=tcgen05): written after round 1's GPU budget was spent and never run.
// The descriptor and shared-memory layout math is the one of tools/tcgen05_gemm_probe.cu, cross-checked on the host
// against CuTe (tools/check_tcgen05_layout.cu); the execution protocol (fences, mbarrier phases) has not met a device.
// Same operator and fused epilogue as the two kernels above.  One CTA = 256 threads = one 128 x TN output tile held in
// TN fp32 TMEM columns; K slabs of 32 are double-buffered in shared memory: all threads gather / split / store slab s
// while the tensor core runs the 12 MMAs (4 K-steps x {lo*hi, hi*lo, hi*hi}) of slab s-1, issued by thread 0 and
// tracked with one mbarrier per buffer (tcgen05.commit).  Operand tiles are K-major, no swizzle: one panel per
// 16-byte K chunk holding 16 bytes of every row, panels padded by 16 bytes so that a warp storing 32 consecutive k of
// one row hits 32 different banks (LBO = rows*16 + 16, SBO = 128).
#define T5_BM 128
#define T5_BK 32
#define T5_CH (T5_BK / 4)

__device__ __forceinline__ uint32_t t5_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t t5_desc(uint32_t saddr, uint32_t lbo_bytes) {          // SBO 128 B, version 1, no swizzle
  return (uint64_t)((saddr >> 4) & 0x3fff) | ((uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16) | ((uint64_t)(128u >> 4) << 32) | ((uint64_t)1 << 46);
}
__device__ __forceinline__ void t5_mma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      :: "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void t5_mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "T5_WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P, [%0], %1;\n\t"
      "@P bra T5_DONE_%=;\n\t"
      "bra T5_WAIT_%=;\n\t"
      "T5_DONE_%=:\n\t}"
      :: "r"(bar), "r"(parity) : "memory");
}

template <int TN>
__global__ void __launch_bounds__(256, 1) nnet_gemm_tc5_kernel(OpDev op, RunCtx c) {
  constexpr uint32_t PANEL_A = T5_BM * 16 + 16, PANEL_B = TN * 16 + 16;            // bytes; LBO of the descriptors
  constexpr uint32_t TILE_A = T5_CH * PANEL_A, TILE_B = T5_CH * PANEL_B;
  constexpr uint32_t STAGE = 2 * TILE_A + 2 * TILE_B;                               // a_hi, a_lo, b_hi, b_lo
  constexpr uint32_t TMEM_COLS = TN <= 128 ? 128 : 256;                             // power of two >= TN
  extern __shared__ __align__(16) unsigned char t5_smem[];       // no-swizzle operands need 16-byte alignment only
  __shared__ const float *rowp[T5_BM];
  __shared__ __align__(8) unsigned long long bars[2];
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane_id = tid & 31;
  const int M = c.batch * op.rows * (op.hsplit > 1 ? op.hsplit : 1);
  const int m0 = blockIdx.x * T5_BM, n0 = blockIdx.y * TN;

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(t5_smem_u32(&tmem_base_s)), "r"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(t5_smem_u32(&bars[0])) : "memory");
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(t5_smem_u32(&bars[1])) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_d = tmem_base_s;
  // F32 accumulate (bit 4), TF32 x TF32 (2 << 7, 2 << 10), both K-major, N >> 3 at bit 17, M >> 4 at bit 24
  const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(TN >> 3) << 17) | ((uint32_t)(T5_BM >> 4) << 24);

  uint32_t uses0 = 0u, uses1 = 0u;         // commits issued on each buffer so far (same value in every thread)
  uint32_t slab = 0;
  for (int ti = 0; ti < op.n_terms; ti++) {
    const TermDev t = op.terms[ti];
    __syncthreads();
    if (tid < T5_BM) {
      const int r = m0 + tid;
      const float *p = nullptr;
      if (r < M) {
        const RowIdx x = split_row(op, r);
        p = src_row_ptr(c, t, x.lane, map_row(t, x.i));
        if (t.col_lim > 0) {
          const int cb = x.h * t.col_step + t.col_off;
          p = (cb >= 0 && cb < t.col_lim) ? p + cb : nullptr;
        }
      }
      rowp[tid] = p;
    }
    __syncthreads();
    for (int kk = 0; kk < t.klen; kk += T5_BK, slab++) {
      const uint32_t b = slab & 1u;
      unsigned char *a_hi = t5_smem + b * STAGE, *a_lo = a_hi + TILE_A, *b_hi = a_lo + TILE_A, *b_lo = b_hi + TILE_B;
      // global -> registers (a warp reads 32 consecutive k of one row: 128-byte segments), before waiting for the buffer
      const int k = kk + lane_id;
      const bool kin = k < t.klen;
      float ra[T5_BM / 8], rb[TN / 8];
#pragma unroll
      for (int e = 0; e < T5_BM / 8; e++) {
        const float *p = rowp[warp + e * 8];
        ra[e] = (p && kin) ? p[k] : 0.f;
      }
#pragma unroll
      for (int e = 0; e < TN / 8; e++) {
        const int n = n0 + warp + e * 8;
        rb[e] = (n < op.N && kin) ? __ldg(&op.w[(long long)n * op.K + t.k0 + k]) : 0.f;
      }
      // the MMAs that read this buffer two slabs ago must have retired
      const uint32_t used = b ? uses1 : uses0;
      if (used > 0) t5_mbar_wait(t5_smem_u32(&bars[b]), (used - 1u) & 1u);
      const uint32_t koff = (uint32_t)(lane_id >> 2), kin4 = (uint32_t)(lane_id & 3) * 4u;
#pragma unroll
      for (int e = 0; e < T5_BM / 8; e++) {
        const uint32_t off = koff * PANEL_A + (uint32_t)(warp + e * 8) * 16u + kin4;
        const uint32_t hi = to_tf32(ra[e]);
        *reinterpret_cast<uint32_t *>(a_hi + off) = hi;
        *reinterpret_cast<uint32_t *>(a_lo + off) = to_tf32(ra[e] - __uint_as_float(hi));
      }
#pragma unroll
      for (int e = 0; e < TN / 8; e++) {
        const uint32_t off = koff * PANEL_B + (uint32_t)(warp + e * 8) * 16u + kin4;
        const uint32_t hi = to_tf32(rb[e]);
        *reinterpret_cast<uint32_t *>(b_hi + off) = hi;
        *reinterpret_cast<uint32_t *>(b_lo + off) = to_tf32(rb[e] - __uint_as_float(hi));
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");       // generic-proxy stores -> visible to the tensor core
      __syncthreads();
      if (tid == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
        for (int ks = 0; ks < T5_BK / 8; ks++) {                          // K = 8 per instruction = two 16-byte chunks = two panels
          const uint64_t dah = t5_desc(t5_smem_u32(a_hi) + (uint32_t)ks * 2u * PANEL_A, PANEL_A), dal = t5_desc(t5_smem_u32(a_lo) + (uint32_t)ks * 2u * PANEL_A, PANEL_A);
          const uint64_t dbh = t5_desc(t5_smem_u32(b_hi) + (uint32_t)ks * 2u * PANEL_B, PANEL_B), dbl = t5_desc(t5_smem_u32(b_lo) + (uint32_t)ks * 2u * PANEL_B, PANEL_B);
          t5_mma(tmem_d, dal, dbh, idesc, (slab > 0 || ks > 0) ? 1u : 0u);
          t5_mma(tmem_d, dah, dbl, idesc, 1u);
          t5_mma(tmem_d, dah, dbh, idesc, 1u);
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(t5_smem_u32(&bars[b])) : "memory");
      }
      if (b) uses1++; else uses0++;
    }
  }
  // all MMAs retired: a commit tracks every tcgen05 operation issued before it, so the most recent one is enough
  if (slab > 0) {
    const uint32_t lb = (slab - 1u) & 1u;
    t5_mbar_wait(t5_smem_u32(&bars[lb]), ((lb ? uses1 : uses0) - 1u) & 1u);
  }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

  // ---- epilogue: warp w reads TMEM lanes 32*(w%4) .. +31 (its quarter) = tile rows, columns of half w/4
  {
    const int q = warp & 3, half = warp >> 2;
    const int r = m0 + q * 32 + lane_id;
    const bool live = r < M;
    RowIdx x = {0, 0, 0};
    float *orow = nullptr;
    const float *rrow = nullptr;
    if (live) {
      x = split_row(op, r);
      orow = (op.out_kind == 0) ? c.arena + (long long)x.lane * c.arena_stride + op.out_off + (long long)x.i * op.out_dim + (long long)x.h * op.N
                                : c.d_out[x.lane] + (long long)x.i * c.out_stride;
      if (op.has_res) rrow = src_row_ptr(c, op.res, x.lane, map_row(op.res, x.i));
    }
    for (int c0 = half * (TN / 2); c0 < (half + 1) * (TN / 2); c0 += 16) {
      uint32_t v[16];
      const uint32_t taddr = tmem_d + ((uint32_t)(q * 32) << 16) + (uint32_t)c0;
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
          : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
            "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
          : "r"(taddr));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      if (live) {
#pragma unroll
        for (int j = 0; j < 16; j++) {
          const int n = n0 + c0 + j;
          if (n >= op.N) continue;
          float val = slab > 0 ? __uint_as_float(v[j]) : 0.f;
          if (op.bias) val = __fadd_rn(val, __ldg(&op.bias[n]));
          if (op.relu) val = fmaxf(val, 0.f);
          if (op.bn_scale) val = __fadd_rn(__fmul_rn(val, __ldg(&op.bn_scale[n])), __ldg(&op.bn_offset[n]));
          if (rrow) val = __fadd_rn(__fmul_rn(op.res_alpha, rrow[n]), val);
          if (!op.log_softmax) {
            if (op.sub_vec) val = __fadd_rn(val, -__ldg(&op.sub_vec[n]));
            if (op.out_scale != 1.0f) val = __fmul_rn(val, op.out_scale);
          }
          orow[n] = val;
        }
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_d), "r"(TMEM_COLS) : "memory");
}

// out[r, blk*block_dim + c] = sum_terms(scale * src[map(r), c]) (+ BatchNorm)
__global__ void nnet_ew_kernel(OpDev op, RunCtx c) {
  const long long total = (long long)c.batch * op.rows * op.out_dim;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total;
       e += (long long)gridDim.x * blockDim.x) {
    int col = (int)(e % op.out_dim);
    long long r = e / op.out_dim;
    int lane = (int)(r / op.rows), ri = (int)(r - (long long)lane * op.rows);
    int blk = col / op.block_dim, cc = col - blk * op.block_dim;
    float v = 0.f;
    bool first = true;
    for (int ti = 0; ti < op.n_terms; ti++) {
      if (op.term_block[ti] != blk) continue;
      const TermDev &t = op.terms[ti];
      float x = src_row_ptr(c, t, lane, map_row(t, ri))[cc];
      float tv = (t.scale == 1.0f) ? x : __fmul_rn(t.scale, x);
      v = first ? tv : __fadd_rn(v, tv);
      first = false;
    }
    if (op.bn_scale) v = __fadd_rn(__fmul_rn(v, __ldg(&op.bn_scale[col])), __ldg(&op.bn_offset[col]));
    float *orow = c.arena + (long long)lane * c.arena_stride + op.out_off + (long long)ri * op.out_dim;
    orow[col] = v;
  }
}

// LogSoftmaxComponent::Propagate (nnet-simple-component.cc:3618-3625), one warp per row, in place
__global__ void nnet_logsoftmax_kernel(OpDev op, RunCtx c) {
  int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane_id = threadIdx.x & 31;
  int M = c.batch * op.rows;
  if (warp >= M) return;
  int lane = warp / op.rows, ri = warp - lane * op.rows;
  float *row = (op.out_kind == 0) ? c.arena + (long long)lane * c.arena_stride + op.out_off + (long long)ri * op.out_dim
                                  : c.d_out[lane] + (long long)ri * c.out_stride;
  float mx = -INFINITY;
  for (int n = lane_id; n < op.N; n += 32) mx = fmaxf(mx, row[n]);
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  float s = 0.f;
  for (int n = lane_id; n < op.N; n += 32) s += expf(row[n] - mx);
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  float lse = mx + logf(s);
  for (int n = lane_id; n < op.N; n += 32) {
    float v = row[n] - lse;
    if (op.sub_vec) v = __fadd_rn(v, -__ldg(&op.sub_vec[n]));        // then -log prior, x acoustic_scale
    if (op.out_scale != 1.0f) v = __fmul_rn(v, op.out_scale);
    row[n] = v;
  }
}

}  // namespace b2k

using namespace b2k;

struct b2k_nnet {
  std::vector<OpDev> ops;
  std::vector<int> log_softmax;      // per op flag
  float *d_blob = nullptr;
  float *d_arena = nullptr;
  long long arena_stride = 0;
  int max_batch = 0;
  int n_out = 0, out_dim = 0, in_rows = 0, in_dim = 0, iv_rows = 0, iv_dim = 0;
  const float **d_in = nullptr, **d_iv = nullptr;
  float **d_outp = nullptr;
  const float **h_in = nullptr, **h_iv = nullptr;
  float **h_outp = nullptr;
  cudaEvent_t staging_free = nullptr;
  double flops_per_lane = 0;
};

// B2K_NNET_GEMM=simt selects the fp32 FFMA kernel (kept for A/B numerics and timing); =tcgen05 the experimental
// TMEM kernel above (never run on a device yet); anything else, and the default, is the mma.sync 3xTF32 kernel
static int gemm_mode() {
  static int v = -1;
  if (v < 0) { const char *e = getenv("B2K_NNET_GEMM"); v = (e && !strcmp(e, "simt")) ? 1 : (e && !strcmp(e, "tcgen05")) ? 2 : 0; }
  return v;
}
static bool use_simt_gemm() { return gemm_mode() == 1; }

extern "C" {

int b2k_nnet_create(const b2k_nnet_node *nodes, int32_t n_nodes, const b2k_nnet_op *ops, int32_t n_ops,
                    const float *blob, int64_t blob_len, int32_t max_batch, b2k_nnet **out) {
  if (!nodes || !ops || !blob || !out || n_nodes <= 0 || n_ops <= 0 || max_batch <= 0)
    return set_error(B2K_ERR_INVALID, "b2k_nnet_create: bad args");
  int rc = require_device();
  if (rc) return rc;
  b2k_nnet *nn = new b2k_nnet();
  nn->max_batch = max_batch;
  B2K_CUDA_CHECK(cudaMalloc((void **)&nn->d_blob, sizeof(float) * (size_t)blob_len));
  B2K_CUDA_CHECK(cudaMemcpy(nn->d_blob, blob, sizeof(float) * (size_t)blob_len, cudaMemcpyHostToDevice));
  // arena: offsets were assigned by the compiler (liveness-based reuse)
  long long arena = 0;
  for (int i = 0; i < n_nodes; i++) {
    if (nodes[i].kind == 0 && nodes[i].rows > 0)
      arena = std::max(arena, (long long)nodes[i].arena_off + (long long)nodes[i].rows * nodes[i].dim);
    if (nodes[i].kind == 1) { nn->in_rows = nodes[i].rows; nn->in_dim = nodes[i].dim; }
    if (nodes[i].kind == 2) { nn->iv_rows = nodes[i].rows; nn->iv_dim = nodes[i].dim; }
    if (nodes[i].kind == 3) { nn->n_out = nodes[i].rows; nn->out_dim = nodes[i].dim; }
  }
  nn->arena_stride = (arena + 31) / 32 * 32;
  B2K_CUDA_CHECK(cudaMalloc((void **)&nn->d_arena, sizeof(float) * (size_t)nn->arena_stride * max_batch));
  auto bp = [&](int64_t off) -> const float * { return off < 0 ? nullptr : nn->d_blob + off; };
  auto mk_term = [&](const b2k_nnet_term &t, TermDev *d) -> int {
    if (t.src < 0 || t.src >= n_nodes) return set_error(B2K_ERR_INVALID, "term source out of range");
    const b2k_nnet_node &s = nodes[t.src];
    d->src_kind = s.kind == 3 ? 0 : s.kind; d->src_off = s.arena_off; d->src_dim = s.dim;
    d->ratio = t.ratio; d->shift = t.shift; d->lo = t.lo; d->hi = t.hi; d->ivec = t.ivec; d->C = t.C > 0 ? t.C : 1;
    d->m = t.m; d->k0 = t.k0; d->klen = t.klen; d->scale = t.scale;
    d->col_step = t.col_step; d->col_off = t.col_off; d->col_lim = t.col_lim;
    if (t.col_lim < 0 || (t.col_lim > 0 && (t.klen > s.dim || t.col_lim > s.dim))) return set_error(B2K_ERR_INVALID, "bad column window");
    return 0;
  };
  for (int i = 0; i < n_ops; i++) {
    const b2k_nnet_op &o = ops[i];
    if (o.n_terms < 1 || o.n_terms > 12 || o.out < 0 || o.out >= n_nodes || o.hsplit < 0) { return set_error(B2K_ERR_INVALID, "bad op"); }
    OpDev d;
    memset(&d, 0, sizeof(d));
    d.type = o.type; d.out_off = nodes[o.out].arena_off; d.out_dim = nodes[o.out].dim;
    d.out_kind = nodes[o.out].kind == 3 ? 3 : 0;
    d.rows = o.rows; d.N = o.N; d.K = o.K; d.n_terms = o.n_terms;
    for (int t = 0; t < o.n_terms; t++) { if ((rc = mk_term(o.terms[t], &d.terms[t]))) return rc; d.term_block[t] = o.terms[t].block; }
    d.w = bp(o.w); d.bias = bp(o.bias); d.bn_scale = bp(o.bn_scale); d.bn_offset = bp(o.bn_offset); d.sub_vec = bp(o.sub_vec);
    d.hsplit = o.hsplit > 1 ? o.hsplit : 1;
    if (d.hsplit > 1 && (o.type != 0 || o.log_softmax || o.has_res || nodes[o.out].kind == 3 || (long long)d.hsplit * o.N != nodes[o.out].dim))
      return set_error(B2K_ERR_INVALID, "hsplit is for internal convolution GEMM ops whose node dim is hsplit * N");
    d.log_softmax = o.log_softmax; d.relu = o.relu; d.has_res = o.has_res; d.res_alpha = o.res_alpha; d.out_scale = o.out_scale; d.block_dim = o.block_dim > 0 ? o.block_dim : 1;
    if (o.has_res && (rc = mk_term(o.res, &d.res))) return rc;
    nn->ops.push_back(d);
    nn->log_softmax.push_back(o.log_softmax);
    if (o.type == 0) nn->flops_per_lane += 2.0 * o.K * o.N * o.rows * d.hsplit;
  }
  size_t pb = sizeof(void *) * max_batch;
  B2K_CUDA_CHECK(cudaMalloc((void **)&nn->d_in, pb)); B2K_CUDA_CHECK(cudaMalloc((void **)&nn->d_iv, pb));
  B2K_CUDA_CHECK(cudaMalloc((void **)&nn->d_outp, pb));
  B2K_CUDA_CHECK(cudaMallocHost((void **)&nn->h_in, pb)); B2K_CUDA_CHECK(cudaMallocHost((void **)&nn->h_iv, pb));
  B2K_CUDA_CHECK(cudaMallocHost((void **)&nn->h_outp, pb));
  B2K_CUDA_CHECK(cudaEventCreateWithFlags(&nn->staging_free, cudaEventDisableTiming));
  *out = nn;
  return B2K_OK;
}

int b2k_nnet_destroy(b2k_nnet *nn) {
  if (!nn) return B2K_OK;
  cudaDeviceSynchronize();
  cudaFree(nn->d_blob); cudaFree(nn->d_arena); cudaFree(nn->d_in); cudaFree(nn->d_iv); cudaFree(nn->d_outp);
  cudaFreeHost(nn->h_in); cudaFreeHost(nn->h_iv); cudaFreeHost(nn->h_outp);
  if (nn->staging_free) cudaEventDestroy(nn->staging_free);
  delete nn;
  return B2K_OK;
}

int32_t b2k_nnet_num_output_frames(const b2k_nnet *nn) { return nn ? nn->n_out : -1; }
int32_t b2k_nnet_output_dim(const b2k_nnet *nn) { return nn ? nn->out_dim : -1; }
double b2k_nnet_flops_per_lane(const b2k_nnet *nn) { return nn ? nn->flops_per_lane : 0.0; }
int32_t b2k_nnet_num_launches_per_run(const b2k_nnet *nn) {
  if (!nn) return -1;
  int n = (int)nn->ops.size();
  for (int f : nn->log_softmax) n += f ? 1 : 0;
  return n;
}

int b2k_nnet_run(b2k_nnet *nn, int32_t batch, const float *const *d_input, int32_t in_stride,
                 const float *const *d_ivectors, int32_t iv_stride, float *const *d_output,
                 int32_t out_stride, void *stream) {
  if (!nn || batch <= 0 || batch > nn->max_batch || !d_input || !d_output)
    return set_error(B2K_ERR_INVALID, "b2k_nnet_run: bad args");
  if (nn->iv_dim > 0 && !d_ivectors) return set_error(B2K_ERR_INVALID, "Neural net expects iVectors but none provided");  // decodable-simple-looped.cc:271
  cudaStream_t st = (cudaStream_t)stream;
  B2K_CUDA_CHECK(cudaEventSynchronize(nn->staging_free));
  for (int i = 0; i < batch; i++) { nn->h_in[i] = d_input[i]; nn->h_iv[i] = d_ivectors ? d_ivectors[i] : nullptr; nn->h_outp[i] = d_output[i]; }
  size_t pb = sizeof(void *) * batch;
  B2K_CUDA_CHECK(cudaMemcpyAsync((void *)nn->d_in, nn->h_in, pb, cudaMemcpyHostToDevice, st));
  B2K_CUDA_CHECK(cudaMemcpyAsync((void *)nn->d_iv, nn->h_iv, pb, cudaMemcpyHostToDevice, st));
  B2K_CUDA_CHECK(cudaMemcpyAsync((void *)nn->d_outp, nn->h_outp, pb, cudaMemcpyHostToDevice, st));
  B2K_CUDA_CHECK(cudaEventRecord(nn->staging_free, st));
  RunCtx c;
  c.arena = nn->d_arena; c.arena_stride = nn->arena_stride; c.d_input = nn->d_in; c.in_stride = in_stride;
  c.d_ivec = nn->d_iv; c.iv_stride = iv_stride; c.d_out = nn->d_outp; c.out_stride = out_stride; c.batch = batch;
  for (size_t i = 0; i < nn->ops.size(); i++) {
    const OpDev &op = nn->ops[i];
    long long M = (long long)batch * op.rows * (op.type == 0 && op.hsplit > 1 ? op.hsplit : 1);
    if (op.type == 0) {
      if (use_simt_gemm()) {
        dim3 grid((unsigned)((M + GM_BM - 1) / GM_BM), (op.N + GM_BN - 1) / GM_BN);
        nnet_gemm_kernel<<<grid, 256, 0, st>>>(op, c);
      } else if (gemm_mode() == 2) {
        static bool configured5 = false;
        const int s128 = 2 * (2 * T5_CH * (T5_BM * 16 + 16) + 2 * T5_CH * (128 * 16 + 16)), s96 = 2 * (2 * T5_CH * (T5_BM * 16 + 16) + 2 * T5_CH * (96 * 16 + 16));
        if (!configured5) {
          B2K_CUDA_CHECK(cudaFuncSetAttribute(nnet_gemm_tc5_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, s128));
          B2K_CUDA_CHECK(cudaFuncSetAttribute(nnet_gemm_tc5_kernel<96>, cudaFuncAttributeMaxDynamicSharedMemorySize, s96));
          configured5 = true;
        }
        const int pad128 = (op.N + 127) / 128 * 128, pad96 = (op.N + 95) / 96 * 96;
        if (pad96 < pad128) {
          dim3 grid((unsigned)((M + T5_BM - 1) / T5_BM), pad96 / 96);
          nnet_gemm_tc5_kernel<96><<<grid, 256, s96, st>>>(op, c);
        } else {
          dim3 grid((unsigned)((M + T5_BM - 1) / T5_BM), pad128 / 128);
          nnet_gemm_tc5_kernel<128><<<grid, 256, s128, st>>>(op, c);
        }
      } else {
        static bool configured = false;
        const int smem4 = (int)(sizeof(uint32_t) * 2 * (TC_BM + 128) * TC_LD), smem3 = (int)(sizeof(uint32_t) * 2 * (TC_BM + 96) * TC_LD);
        if (!configured) {
          B2K_CUDA_CHECK(cudaFuncSetAttribute(nnet_gemm_tc_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem4));
          B2K_CUDA_CHECK(cudaFuncSetAttribute(nnet_gemm_tc_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem3));
          configured = true;
        }
        // column tile 128 or 96, whichever pads N less (the N = 96 / 192 bottlenecks of TDNN-F)
        const int pad4 = (op.N + 127) / 128 * 128, pad3 = (op.N + 95) / 96 * 96;
        if (pad3 < pad4) {
          dim3 grid((unsigned)((M + TC_BM - 1) / TC_BM), pad3 / 96);
          nnet_gemm_tc_kernel<3><<<grid, 256, smem3, st>>>(op, c);
        } else {
          dim3 grid((unsigned)((M + TC_BM - 1) / TC_BM), pad4 / 128);
          nnet_gemm_tc_kernel<4><<<grid, 256, smem4, st>>>(op, c);
        }
      }
      B2K_LAUNCH_CHECK();
      if (nn->log_softmax[i]) {
        long long threads = M * 32;
        nnet_logsoftmax_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(op, c);
        B2K_LAUNCH_CHECK();
      }
    } else {
      long long total = M * op.out_dim;
      int blocks = (int)std::min<long long>((total + 255) / 256, 148 * 16);
      nnet_ew_kernel<<<blocks, 256, 0, st>>>(op, c);
      B2K_LAUNCH_CHECK();
    }
  }
  return B2K_OK;
}

}  // extern "C"
