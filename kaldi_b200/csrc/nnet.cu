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

#include <cuda.h>

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

// ---------------------------------------------------------------- tcgen05 GEMM, A in TMEM, W by TMA (the default)
//
// Same operator and fused epilogue as the kernels above; the structure was validated stand-alone first
// (tools/tcgen05_gemm_probe2.cu, profiles/r02_nnet_gemm.md).  One CTA = one 128 x TN output tile, accumulator in TN
// fp32 TMEM columns, 3xTF32: D += A_lo*W_hi + A_hi*W_lo + A_hi*W_hi per K step of 8.
//   * W is split ONCE, when the network is uploaded, into W_hi = tf32(W) and W_lo = tf32(W - W_hi) (both are valid
//     TF32 bit patterns, rows padded to a multiple of 4 floats); a K slab of 32 (TN rows x 128 bytes of each) is
//     fetched by TMA (cp.async.bulk.tensor.2d, SWIZZLE_128B, zero fill outside the matrix) into a 3-stage ring and read
//     by the tensor core through K-major SW128 descriptors.  A slab that runs past a term's columns reads the next
//     term's weights against zeros of A.
//   * A is gathered by two groups of four loader warps (even / odd slabs): thread = tile row (row maps, clamping, the
//     convolution's column windows and zero padding are per-thread pointer arithmetic), 128 bytes per slab, split in
//     registers, written straight into TMEM (tcgen05.st 32x32b: lane = row, column = k) and consumed by tcgen05.mma with
//     the A operand in TMEM (.ts form): A never touches shared memory.  TMEM: TN accumulator columns + 2 stages x
//     (32 hi + 32 lo) = 256 columns for TN <= 128, so two CTAs share an SM and one's epilogue overlaps the other's MMAs.
//   * warp 0 = TMA producer, warp 1 = MMA issuer (one thread; tcgen05.commit frees the stages), warps 2..9 = loaders,
//     then epilogue: TMEM -> registers -> 32 x 32 transposes in the (now idle) W ring -> the fused bias / ReLU /
//     BatchNorm / bypass / prior / scale chain with one column per thread -> 128-byte row segments to global memory.
#define TS_BM 128
#define TS_BK 32
#define TS_THREADS 320

struct TsMaps { CUtensorMap hi, lo; };

__device__ __forceinline__ void ts_mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count) : "memory"); }
__device__ __forceinline__ void ts_mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(bar) : "memory"); }
__device__ __forceinline__ void ts_mbar_expect_tx(uint32_t bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ void ts_tma_load_2d(uint32_t dst, const CUtensorMap *map, uint32_t bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               :: "r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void ts_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(bar) : "memory");
}
// K-major SWIZZLE_128B operand: 8-row groups of 1024 B (SBO), LBO field 1 (unused for swizzled K-major), version 1, layout 2
__device__ __forceinline__ uint64_t ts_desc_sw128(uint32_t saddr) {
  return (uint64_t)((saddr >> 4) & 0x3fff) | ((uint64_t)1 << 16) | ((uint64_t)(1024u >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
__device__ __forceinline__ void ts_mma(uint32_t d, uint32_t a_tmem, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
               :: "r"(d), "r"(a_tmem), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void ts_tmem_st32(uint32_t taddr, const uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,"
      "%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"
      :: "r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),
         "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]),
         "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]),
         "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31]) : "memory");
}
__device__ __forceinline__ void ts_tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,"
      "%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
        "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
        "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// slab s of the op -> (term, k offset inside the term); terms are walked in order by every role.
// TMA fetches 16-byte aligned row segments only: the box of a term whose first weight column k0 is not a multiple of 4
// starts at the aligned column below it, d = k0 & 3 columns early, and the loaders place the term's A values d positions
// later in the slab (zeros in front: those columns belong to the previous term).  kk is the slab's offset from the
// ALIGNED start, so slab position e holds term element kk + e - d.
struct TsSlab { int ti, kk; };
__device__ __forceinline__ TsSlab ts_slab(const int *term_slab0, int n_terms, int s) {
  int ti = 0;
  while (ti + 1 < n_terms && term_slab0[ti + 1] <= s) ti++;
  return {ti, (s - term_slab0[ti]) * TS_BK};
}

// FL > 0: the accumulator is emptied every FL slabs (K = 32 FL).  The tensor core adds into its fp32 accumulator with
// truncation, so the error of a long contraction grows with the number of additions into the same accumulator (full-width
// tdnn_1d, K up to 3072: 1.8e-4 of the output scale against 1.9e-6 for the FFMA kernel, tools/nnet_precision_probe.py).
// With the flush a segment's partial sum is read out of TMEM and added to a running tile in shared memory in fp32,
// round to nearest, as the software-accumulation schemes for 3xTF32 do; short contractions (<= FL slabs) never flush.
template <int TN, int SB, int FL>
__global__ void __launch_bounds__(TS_THREADS, (TN <= 128 ? 2 : 1))
nnet_gemm_ts_kernel(const __grid_constant__ OpDev op, const RunCtx c, const __grid_constant__ TsMaps maps) {
  constexpr uint32_t B_TILE = TN * 128;                       // bytes of one hi (or lo) tile: TN rows x 32 floats
  constexpr uint32_t B_STAGE = 2 * B_TILE;
  constexpr uint32_t TMEM_COLS = (TN + 128 <= 256) ? 256 : 512;
  static_assert(SB * B_STAGE >= 8 * 32 * 33 * 4, "the W ring doubles as the epilogue's transpose space");
  extern __shared__ __align__(1024) unsigned char ts_smem_raw[];
  unsigned char *smem = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(ts_smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ __align__(8) unsigned long long bars[2 * SB + 7];   // b_full[SB], b_empty[SB], a_full[2], a_empty[2], acc_full, flush_full, flush_done
  __shared__ uint32_t tmem_base_s;
  __shared__ int term_slab0[13];
  const int tid = threadIdx.x, warp = tid >> 5, lane_id = tid & 31;
  const int M = c.batch * op.rows * (op.hsplit > 1 ? op.hsplit : 1);
  const int ntn_ = (op.N + TN - 1) / TN;
  const int m0 = (int)(blockIdx.x / (unsigned)ntn_) * TS_BM, n0 = (int)(blockIdx.x % (unsigned)ntn_) * TN;
  const uint32_t bar0 = (uint32_t)__cvta_generic_to_shared(&bars[0]);
  auto B_FULL = [&](int s) { return bar0 + 8u * (uint32_t)s; };
  auto B_EMPTY = [&](int s) { return bar0 + 8u * (uint32_t)(SB + s); };
  auto A_FULL = [&](int s) { return bar0 + 8u * (uint32_t)(2 * SB + s); };
  auto A_EMPTY = [&](int s) { return bar0 + 8u * (uint32_t)(2 * SB + 2 + s); };
  const uint32_t ACC_FULL = bar0 + 8u * (uint32_t)(2 * SB + 4);
  const uint32_t FLUSH_FULL = bar0 + 8u * (uint32_t)(2 * SB + 5), FLUSH_DONE = bar0 + 8u * (uint32_t)(2 * SB + 6);
  float *run_s = reinterpret_cast<float *>(smem + (size_t)SB * B_STAGE);   // FL > 0: running sums [TN][128], column major

  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"((uint32_t)__cvta_generic_to_shared(&tmem_base_s)), "r"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 0) {
    for (int s = 0; s < SB; s++) { ts_mbar_init(B_FULL(s), 1); ts_mbar_init(B_EMPTY(s), 1); }
    for (int s = 0; s < 2; s++) { ts_mbar_init(A_FULL(s), 128); ts_mbar_init(A_EMPTY(s), 1); }
    ts_mbar_init(ACC_FULL, 1);
    ts_mbar_init(FLUSH_FULL, 1); ts_mbar_init(FLUSH_DONE, 256);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" :: "l"(&maps.hi) : "memory");
    asm volatile("prefetch.tensormap [%0];" :: "l"(&maps.lo) : "memory");
    int acc = 0;
    for (int ti = 0; ti < op.n_terms; ti++) { term_slab0[ti] = acc; acc += (op.terms[ti].klen + (op.terms[ti].k0 & 3) + TS_BK - 1) / TS_BK; }
    term_slab0[op.n_terms] = acc;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const int nslabs = term_slab0[op.n_terms];
  const uint32_t tmem_base = tmem_base_s;
  const uint32_t tmem_acc = tmem_base;                        // TN columns
  const uint32_t tmem_a = tmem_base + (uint32_t)TN;           // 2 stages x 64 columns
  // F32 accumulate (bit 4), TF32 x TF32 (2 << 7, 2 << 10), both K-major, N >> 3 at bit 17, M >> 4 at bit 24
  const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(TN >> 3) << 17) | ((uint32_t)(TS_BM >> 4) << 24);

  if (warp == 0) {
    // ------------------------------------------------ TMA producer
    if (lane_id == 0) {
      for (int s = 0; s < nslabs; s++) {
        const int st = s % SB;
        const TsSlab sl = ts_slab(term_slab0, op.n_terms, s);
        t5_mbar_wait(B_EMPTY(st), (((uint32_t)(s / SB)) & 1u) ^ 1u);
        ts_mbar_expect_tx(B_FULL(st), B_STAGE);
        const uint32_t dst = (uint32_t)__cvta_generic_to_shared(smem + (size_t)st * B_STAGE);
        const int kcol = (op.terms[sl.ti].k0 & ~3) + sl.kk;     // 16-byte aligned box start
        ts_tma_load_2d(dst, &maps.hi, B_FULL(st), kcol, n0);
        ts_tma_load_2d(dst + B_TILE, &maps.lo, B_FULL(st), kcol, n0);
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------ MMA issuer
    if (lane_id == 0) {
      for (int s = 0; s < nslabs; s++) {
        const int st = s % SB, as = s & 1;
        bool fresh = (s == 0);
        if (FL > 0 && s > 0 && s % (FL > 0 ? FL : 1) == 0) {                 // segment boundary: hand the accumulator to the flush, start a new sum
          ts_commit(FLUSH_FULL);
          t5_mbar_wait(FLUSH_DONE, ((uint32_t)(s / (FL > 0 ? FL : 1) - 1)) & 1u);
          fresh = true;
        }
        t5_mbar_wait(B_FULL(st), ((uint32_t)(s / SB)) & 1u);
        t5_mbar_wait(A_FULL(as), ((uint32_t)(s >> 1)) & 1u);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t bh = (uint32_t)__cvta_generic_to_shared(smem + (size_t)st * B_STAGE), bl = bh + B_TILE;
#pragma unroll
        for (int ks = 0; ks < TS_BK / 8; ks++) {
          const uint64_t dbh = ts_desc_sw128(bh + (uint32_t)ks * 32u), dbl = ts_desc_sw128(bl + (uint32_t)ks * 32u);
          const uint32_t ah = tmem_a + (uint32_t)as * 64u + (uint32_t)ks * 8u, al = ah + 32u;
          ts_mma(tmem_acc, al, dbh, idesc, (!fresh || ks > 0) ? 1u : 0u);
          ts_mma(tmem_acc, ah, dbl, idesc, 1u);
          ts_mma(tmem_acc, ah, dbh, idesc, 1u);
        }
        ts_commit(B_EMPTY(st));
        ts_commit(A_EMPTY(as));
      }
      ts_commit(ACC_FULL);
    }
  } else {
    // ------------------------------------------------ A loaders (group g = even / odd slabs), then the epilogue
    const int g = (warp - 2) >> 2, q = warp & 3;             // q = the TMEM lane quarter this warp may access
    const int r = q * 32 + lane_id;                           // tile row owned by this thread
    const bool live = m0 + r < M;
    RowIdx x = {0, 0, 0};
    if (live) x = split_row(op, m0 + r);
    float cur[32];
    int cur_ti = -1;
    const float *arow = nullptr;
    auto fetch = [&](int s) {
      const TsSlab sl = ts_slab(term_slab0, op.n_terms, s);
      const TermDev &t = op.terms[sl.ti];
      if (sl.ti != cur_ti) {                                  // this thread's source row for the term
        cur_ti = sl.ti;
        arow = nullptr;
        if (live) {
          arow = src_row_ptr(c, t, x.lane, map_row(t, x.i));
          if (t.col_lim > 0) {                                // convolution patch: column window of this output height
            const int cb = x.h * t.col_step + t.col_off;
            arow = (cb >= 0 && cb < t.col_lim) ? arow + cb : nullptr;   // outside = height zero padding
          }
        }
      }
      const int klen = t.klen;
      const int k0e = sl.kk - (t.k0 & 3);                     // term element held by slab position 0 (may be negative)
      if (!arow) {
#pragma unroll
        for (int e = 0; e < 32; e++) cur[e] = 0.f;
      } else if (k0e >= 0 && k0e + 32 <= klen && ((reinterpret_cast<uintptr_t>(arow + k0e) & 15) == 0)) {
#pragma unroll
        for (int cc = 0; cc < 8; cc++) {
          const float4 v = *reinterpret_cast<const float4 *>(arow + k0e + cc * 4);
          cur[cc * 4 + 0] = v.x; cur[cc * 4 + 1] = v.y; cur[cc * 4 + 2] = v.z; cur[cc * 4 + 3] = v.w;
        }
      } else {
#pragma unroll
        for (int e = 0; e < 32; e++) { const int ke = k0e + e; cur[e] = (ke >= 0 && ke < klen) ? arow[ke] : 0.f; }
      }
    };
    constexpr int HALF = TN / 2;
    const int n_flush = (FL > 0 && nslabs > 0) ? (nslabs - 1) / (FL > 0 ? FL : 1) : 0;
    int flushed = 0;
    auto flush = [&]() {                                      // this thread's row, its group's column half: TMEM -> running tile
      t5_mbar_wait(FLUSH_FULL, (uint32_t)flushed & 1u);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      for (int c0 = g * HALF; c0 < (g + 1) * HALF; c0 += 16) {   // 16 columns at a time (HALF = 48 or 80): the next slab's 32 values stay in registers
        uint32_t v[16];
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
            : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
              "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
            : "r"(tmem_acc + ((uint32_t)(q * 32) << 16) + (uint32_t)c0));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int j = 0; j < 16; j++) {
          float *rp = run_s + (size_t)(c0 + j) * TS_BM + r;    // consecutive rows across a warp: conflict free
          *rp = flushed ? __fadd_rn(*rp, __uint_as_float(v[j])) : __uint_as_float(v[j]);
        }
      }
      flushed++;
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      ts_mbar_arrive(FLUSH_DONE);
    };
    if (g < nslabs) fetch(g);
    for (int s = g; s < nslabs; s += 2) {
      uint32_t hv[32];
      t5_mbar_wait(A_EMPTY(g), (((uint32_t)(s >> 1)) & 1u) ^ 1u);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t t0 = tmem_a + ((uint32_t)(q * 32) << 16) + (uint32_t)g * 64u;
#pragma unroll
      for (int e = 0; e < 32; e++) hv[e] = to_tf32(cur[e]);
      ts_tmem_st32(t0, hv);
#pragma unroll
      for (int e = 0; e < 32; e++) hv[e] = to_tf32(cur[e] - __uint_as_float(hv[e]));
      ts_tmem_st32(t0 + 32u, hv);
      if (s + 2 < nslabs) fetch(s + 2);                       // next slab of this group in flight while the MMAs run
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      ts_mbar_arrive(A_FULL(g));
      // boundaries lie before slabs FL, 2 FL, ...: this slab's A is already in place, so the flush costs the tensor core only
      // the read-out itself
      if (FL > 0) while (flushed < s / (FL > 0 ? FL : 1)) flush();
    }
    // ---- epilogue.  Warp (q, g): rows 32q .. 32q+31, columns g*TN/2 .. +TN/2, 32 columns at a time: TMEM gives a
    //      thread its row's 32 columns; a 32 x 32 transpose in shared memory gives it one column of the 32 rows, so the
    //      per-column parameters are loaded once and every row is stored as one 128-byte segment.
    if (FL > 0) while (flushed < n_flush) flush();             // (a group whose last slab lies before the last boundary)
    if (nslabs > 0) t5_mbar_wait(ACC_FULL, 0u);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    float *tsp = reinterpret_cast<float *>(smem) + (size_t)(warp - 2) * (32 * 33);   // the W ring is idle now
    float *orow = nullptr;
    const float *rrow = nullptr;
    if (live) {
      orow = (op.out_kind == 0) ? c.arena + (long long)x.lane * c.arena_stride + op.out_off + (long long)x.i * op.out_dim + (long long)x.h * op.N
                                : c.d_out[x.lane] + (long long)x.i * c.out_stride;
      if (op.has_res) rrow = src_row_ptr(c, op.res, x.lane, map_row(op.res, x.i));
    }
    const unsigned long long orow_u = reinterpret_cast<unsigned long long>(orow), rrow_u = reinterpret_cast<unsigned long long>(rrow);
    for (int c0 = g * HALF; c0 < (g + 1) * HALF; c0 += 32) {
      uint32_t v[32];
      const int ncols = min(32, (g + 1) * HALF - c0);         // TN / 2 need not be a multiple of 32 (TN = 96, 160)
      if (nslabs > 0) {
        if (ncols == 32) ts_tmem_ld32(tmem_acc + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
        else {                                                // 16 columns
          asm volatile(
              "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
              : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
              : "r"(tmem_acc + ((uint32_t)(q * 32) << 16) + (uint32_t)c0));
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
          for (int j = 16; j < 32; j++) v[j] = 0u;
        }
      } else {
#pragma unroll
        for (int j = 0; j < 32; j++) v[j] = 0u;
      }
      if (FL > 0 && n_flush > 0) {                             // last segment + the running sums of the earlier ones
#pragma unroll
        for (int j = 0; j < 32; j++)
          if (j < ncols) v[j] = __float_as_uint(__fadd_rn(run_s[(size_t)(c0 + j) * TS_BM + r], __uint_as_float(v[j])));
      }
      __syncwarp();
#pragma unroll
      for (int j = 0; j < 32; j++) tsp[lane_id * 33 + j] = __uint_as_float(v[j]);   // (row = lane, column j): conflict free
      __syncwarp();
      const int n = n0 + c0 + lane_id;                        // this thread's column
      const bool ncol_ok = lane_id < ncols && n < op.N;
      float bias = 0.f, bsc = 1.f, bof = 0.f, sub = 0.f;
      if (ncol_ok) {
        if (op.bias) bias = __ldg(&op.bias[n]);
        if (op.bn_scale) { bsc = __ldg(&op.bn_scale[n]); bof = __ldg(&op.bn_offset[n]); }
        if (op.sub_vec) sub = __ldg(&op.sub_vec[n]);
      }
      // the bypass term's 32 loads (one per row, 128 contiguous bytes per warp) are all issued before the first is used:
      // one memory latency per chunk instead of one per group of rows (the epilogue was the longest part of a short-K tile)
      float resv[32];
      if (op.has_res) {
#pragma unroll
        for (int rr = 0; rr < 32; rr++) {
          const unsigned long long ru = __shfl_sync(0xffffffffu, rrow_u, rr);
          resv[rr] = (ru && ncol_ok) ? reinterpret_cast<const float *>(ru)[n] : 0.f;
        }
      }
#pragma unroll
      for (int rr = 0; rr < 32; rr++) {
        const unsigned long long ou = __shfl_sync(0xffffffffu, orow_u, rr);
        if (!ou || !ncol_ok) continue;
        float val = tsp[rr * 33 + lane_id];
        if (op.bias) val = __fadd_rn(val, bias);
        if (op.relu) val = fmaxf(val, 0.f);
        if (op.bn_scale) val = __fadd_rn(__fmul_rn(val, bsc), bof);
        if (op.has_res) val = __fadd_rn(__fmul_rn(op.res_alpha, resv[rr]), val);
        if (!op.log_softmax) {
          if (op.sub_vec) val = __fadd_rn(val, -sub);
          if (op.out_scale != 1.0f) val = __fmul_rn(val, op.out_scale);
        }
        reinterpret_cast<float *>(ou)[n] = val;
      }
      __syncwarp();
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "r"(TMEM_COLS) : "memory");
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
  // tcgen05 path: W split once into hi / lo TF32 matrices (rows padded to 4 floats) + their TMA descriptors, per GEMM op
  float *d_wsplit = nullptr;
  std::vector<TsMaps> ts_maps;       // per op (unused entries for non-GEMM ops)
  std::vector<int> ts_tn;            // column tile per op: 96, 128 (two CTAs per SM) or 160
};

// The default is the tcgen05 kernel with A in TMEM and W by TMA (nnet_gemm_ts_kernel).  B2K_NNET_GEMM selects the others
// for A/B numerics and timing: mma = mma.sync 3xTF32 (round 1's measured kernel), simt = fp32 FFMA, tcgen05 = the first
// TMEM kernel (both operands staged through shared memory by the CTA's threads).
static int gemm_mode() {
  static int v = -1;
  if (v < 0) {
    const char *e = getenv("B2K_NNET_GEMM");
    v = !e ? 3 : !strcmp(e, "simt") ? 1 : !strcmp(e, "tcgen05") ? 2 : !strcmp(e, "mma") ? 0 : 3;
  }
  return v;
}

static float host_tf32_rna(float x) {                        // cvt.rna.tf32.f32
  uint32_t u; memcpy(&u, &x, 4);
  if ((u & 0x7f800000u) == 0x7f800000u) return x;            // inf / nan unchanged
  u = (u + 0x1000u) & 0xffffe000u;
  float r; memcpy(&r, &u, 4);
  return r;
}

// Accumulator flush interval of the long contractions, in K slabs of 32 (B2K_NNET_FLUSH=0 switches it off for A/B runs).
#define TS_FL 16
static int ts_flush_slabs() {
  static int v = -1;
  if (v < 0) { const char *e = getenv("B2K_NNET_FLUSH"); v = (e && atoi(e) == 0) ? 0 : TS_FL; }
  return v;
}
static int ts_num_slabs(const OpDev &op) {
  int n = 0;
  for (int ti = 0; ti < op.n_terms; ti++) n += (op.terms[ti].klen + (op.terms[ti].k0 & 3) + TS_BK - 1) / TS_BK;
  return n;
}

static int ts_pick_tn(int N) {
  const int cand[3] = {128, 96, 160};                        // ties go to 128, then 96 (both keep two CTAs per SM)
  int best = 128, best_pad = (N + 127) / 128 * 128;
  for (int i = 1; i < 3; i++) { const int pad = (N + cand[i] - 1) / cand[i] * cand[i]; if (pad < best_pad) { best = cand[i]; best_pad = pad; } }
  return best;
}

static int ts_make_map(CUtensorMap *m, const float *dptr, int N, int Kp, int TN) {
  cuuint64_t dims[2] = {(cuuint64_t)Kp, (cuuint64_t)N};
  cuuint64_t strides[1] = {(cuuint64_t)Kp * 4};
  cuuint32_t box[2] = {32, (cuuint32_t)TN};
  cuuint32_t es[2] = {1, 1};
  // the driver entry point is looked up at run time: libb2k.so must load (host-only functions, CPU tests) where no
  // libcuda.so.1 exists
  typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                    const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static EncodeTiledFn encode = nullptr;
  if (!encode) {
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    B2K_CUDA_CHECK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    if (!fn || qres != cudaDriverEntryPointSuccess) return set_error(B2K_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
    encode = (EncodeTiledFn)fn;
  }
  CUresult rc = encode(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void *)dptr, dims, strides, box, es,
                       CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                       CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (rc != CUDA_SUCCESS) return set_error(B2K_ERR_CUDA, "cuTensorMapEncodeTiled failed for a weight matrix");
  return B2K_OK;
}

extern "C++" {
template <int TN, int SB, int FL>
static int ts_launch(const OpDev &op, const RunCtx &c, const TsMaps &maps, long long M, cudaStream_t st) {
  static bool configured = false;
  const int smem = SB * 2 * TN * 128 + (FL > 0 ? TN * TS_BM * 4 : 0) + 1024;
  if (!configured) {
    B2K_CUDA_CHECK(cudaFuncSetAttribute(nnet_gemm_ts_kernel<TN, SB, FL>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = true;
  }
  // one-dimensional grid, column tile fastest: the CTAs that share a block of A rows are launched together, so the
  // rows come out of L2 for all but the first of them (with the column tile in blockIdx.y they were N/TN waves apart
  // and every column tile re-read its A rows from HBM)
  const long long ntn = (op.N + TN - 1) / TN, ntm = (M + TS_BM - 1) / TS_BM;
  nnet_gemm_ts_kernel<TN, SB, FL><<<(unsigned)(ntm * ntn), TS_THREADS, smem, st>>>(op, c, maps);
  return B2K_OK;
}
}  // extern "C++"

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
  {
    // W -> W_hi, W_lo (TF32 bit patterns, rows padded to a multiple of 4 floats for TMA's 16-byte stride rule)
    size_t total = 0;
    std::vector<size_t> off(nn->ops.size(), 0);
    for (size_t i = 0; i < nn->ops.size(); i++) {
      if (nn->ops[i].type != 0) continue;
      const size_t Kp = ((size_t)nn->ops[i].K + 3) / 4 * 4;
      off[i] = total;
      total += 2 * (size_t)nn->ops[i].N * Kp;
      total = (total + 63) / 64 * 64;                        // 256-byte aligned matrices
    }
    nn->ts_maps.resize(nn->ops.size());
    nn->ts_tn.assign(nn->ops.size(), 0);
    if (total) {
      std::vector<float> split(total, 0.f);
      for (size_t i = 0; i < nn->ops.size(); i++) {
        const OpDev &d = nn->ops[i];
        if (d.type != 0) continue;
        const size_t Kp = ((size_t)d.K + 3) / 4 * 4;
        const float *w = blob + (d.w - nn->d_blob);          // the same offset in the host blob
        float *hi = split.data() + off[i], *lo = hi + (size_t)d.N * Kp;
        for (int n = 0; n < d.N; n++)
          for (int k = 0; k < d.K; k++) {
            const float v = w[(size_t)n * d.K + k], h = host_tf32_rna(v);
            hi[(size_t)n * Kp + k] = h;
            lo[(size_t)n * Kp + k] = host_tf32_rna(v - h);
          }
      }
      B2K_CUDA_CHECK(cudaMalloc((void **)&nn->d_wsplit, sizeof(float) * total));
      B2K_CUDA_CHECK(cudaMemcpy(nn->d_wsplit, split.data(), sizeof(float) * total, cudaMemcpyHostToDevice));
      for (size_t i = 0; i < nn->ops.size(); i++) {
        const OpDev &d = nn->ops[i];
        if (d.type != 0) continue;
        const int Kp = (d.K + 3) / 4 * 4;
        nn->ts_tn[i] = ts_pick_tn(d.N);
        if (nn->ts_tn[i] == 128 && ts_flush_slabs() > 0 && ts_num_slabs(d) > ts_flush_slabs()) {
          // the running tile of a flushed 128-column tile does not fit beside the weight ring at two CTAs per SM
          nn->ts_tn[i] = ((d.N + 159) / 160 * 160 < (d.N + 95) / 96 * 96) ? 160 : 96;
        }
        const float *hi = nn->d_wsplit + off[i], *lo = hi + (size_t)d.N * Kp;
        if (getenv("B2K_NNET_SYNC")) fprintf(stderr, "[b2k nnet] op %zu: N %d K %d Kp %d rows %d hsplit %d terms %d tn %d hi %p lo %p\n", i, d.N, d.K, Kp, d.rows, d.hsplit, d.n_terms, nn->ts_tn[i], (const void *)hi, (const void *)lo);
        if ((rc = ts_make_map(&nn->ts_maps[i].hi, hi, d.N, Kp, nn->ts_tn[i]))) return rc;
        if ((rc = ts_make_map(&nn->ts_maps[i].lo, lo, d.N, Kp, nn->ts_tn[i]))) return rc;
      }
    }
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
  cudaFree(nn->d_blob); cudaFree(nn->d_arena); cudaFree(nn->d_in); cudaFree(nn->d_iv); cudaFree(nn->d_outp); cudaFree(nn->d_wsplit);
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
      } else if (gemm_mode() == 3) {
        int rc2;
        const bool long_k = ts_flush_slabs() > 0 && ts_num_slabs(op) > ts_flush_slabs();   // (create() gave such ops a 96 / 160 tile)
        if (nn->ts_tn[i] == 96) rc2 = long_k ? ts_launch<96, 2, TS_FL>(op, c, nn->ts_maps[i], M, st) : ts_launch<96, 3, 0>(op, c, nn->ts_maps[i], M, st);
        else if (nn->ts_tn[i] == 160) rc2 = long_k ? ts_launch<160, 3, TS_FL>(op, c, nn->ts_maps[i], M, st) : ts_launch<160, 3, 0>(op, c, nn->ts_maps[i], M, st);
        else rc2 = ts_launch<128, 3, 0>(op, c, nn->ts_maps[i], M, st);
        if (rc2) return rc2;
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
      if (getenv("B2K_NNET_SYNC")) {                          // debugging aid: find the op whose kernel faults
        cudaError_t e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) {
          char msg[256];
          snprintf(msg, sizeof(msg), "op %zu (N %d K %d rows %d hsplit %d terms %d klen0 %d mode %d tn %d): %s", i, op.N, op.K, op.rows, op.hsplit,
                   op.n_terms, op.terms[0].klen, gemm_mode(), nn->ts_tn[i], cudaGetErrorString(e));
          return set_error(B2K_ERR_CUDA, "b2k_nnet_run", msg);
        }
      }
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
