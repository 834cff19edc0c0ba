// tools/tcgen05_gemm_probe2.cu -- NOT part of the product build.
//
// Stand-alone probe of the round-2 nnet3 GEMM pipeline before it goes into kaldi_b200/csrc/nnet.cu:
//   C[M x N] = A[M x K] * W[N x K]^T, fp32-equivalent (3xTF32), accumulator in TMEM, with
//   * W split ONCE on the host into W_hi / W_lo (both valid TF32 bit patterns) and loaded by TMA
//     (cp.async.bulk.tensor.2d, SWIZZLE_128B, box 32 k x TN rows) -> K-major SW128 UMMA descriptors;
//   * A gathered by loader warps (one row per thread, 128 B per K slab), split in registers and
//       MODE 1: written straight to TMEM (tcgen05.st 32x32b) and consumed by tcgen05.mma with the
//               A operand in TMEM (.ts form): no shared-memory traffic for A at all;
//       MODE 0: written to shared memory in the no-swizzle K-major panel layout (the layout the
//               round-1 kernel validated on the device), A from shared memory (.ss form);
//   * warp roles: warp 0 = TMA producer, warp 1 = MMA issuer (+ TMEM alloc), warps 2..9 = two loader
//     groups of four warps (even / odd K slabs), which also run the epilogue (tcgen05.ld -> global).
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/_bin/tc5b tools/tcgen05_gemm_probe2.cu -lcuda
//   tools/_bin/tc5b [mode] [M] [N] [K] [accmode 0|1|2] [dist 0|1]
//
// prints max relative error against a double-precision CPU product and the TFLOP/s of the kernel.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cmath>
#include <cuda.h>
#include <cuda_runtime.h>

#define CHECK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

constexpr int TM = 128, TK = 32;
constexpr int NLOAD_WARPS = 8;
constexpr int NTHREADS = 64 + NLOAD_WARPS * 32;

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t to_tf32(float x) { uint32_t r; asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x)); return r; }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(bar) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P, [%0], %1;\n\t"
      "@P bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}"
      :: "r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap *map, uint32_t bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               :: "r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(bar) : "memory");
}
// K-major SWIZZLE_128B operand: 8-row groups of 1024 B (SBO), LBO field 1, version 1, layout type 2
__device__ __forceinline__ uint64_t desc_sw128(uint32_t saddr) {
  return (uint64_t)((saddr >> 4) & 0x3fff) | ((uint64_t)1 << 16) | ((uint64_t)(1024u >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
// K-major no-swizzle panels (round-1 layout): LBO = panel bytes, SBO = 128
__device__ __forceinline__ uint64_t desc_nosw(uint32_t saddr, uint32_t lbo) {
  return (uint64_t)((saddr >> 4) & 0x3fff) | ((uint64_t)((lbo >> 4) & 0x3fff) << 16) | ((uint64_t)(128u >> 4) << 32) | ((uint64_t)1 << 46);
}
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
               :: "r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a_tmem, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
               :: "r"(d), "r"(a_tmem), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,"
      "%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"
      :: "r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),
         "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]),
         "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]),
         "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31]) : "memory");
}

// MODE 1: A in TMEM (two stages of 64 columns: 32 hi + 32 lo);  MODE 0: A in shared memory (two stages of hi + lo panels)
template <int MODE, int TN, int SB>
__global__ void __launch_bounds__(NTHREADS, 1)
gemm_tc5b(const float *__restrict__ A, float *__restrict__ C, int M, int N, int K,
          const __grid_constant__ CUtensorMap map_hi, const __grid_constant__ CUtensorMap map_lo, int accmode) {
  constexpr uint32_t B_TILE = TN * 128;                       // bytes of one hi (or lo) tile: TN rows x 32 floats
  constexpr uint32_t B_STAGE = 2 * B_TILE;
  constexpr uint32_t A_PANEL = TM * 16, A_TILE = (TK / 4) * A_PANEL, A_STAGE = 2 * A_TILE;   // MODE 0 only
  constexpr uint32_t TMEM_COLS = (MODE == 1) ? 512 : (TN <= 128 ? 128 : 256);   // (probe: room for the accumulation experiments)
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char *smem = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  unsigned char *b_st = smem;                                 // SB stages of {hi, lo}
  unsigned char *a_st = smem + SB * B_STAGE;                  // MODE 0: 2 stages of {hi, lo}
  __shared__ __align__(8) unsigned long long bars[2 * SB + 5];   // b_full[SB], b_empty[SB], a_full[2], a_empty[2], acc_full
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int m0 = blockIdx.y * TM, n0 = blockIdx.x * TN;
  const int nslabs = (K + TK - 1) / TK;
  const uint32_t bar0 = smem_u32(&bars[0]);
  auto B_FULL = [&](int s) { return bar0 + 8u * (uint32_t)s; };
  auto B_EMPTY = [&](int s) { return bar0 + 8u * (uint32_t)(SB + s); };
  auto A_FULL = [&](int s) { return bar0 + 8u * (uint32_t)(2 * SB + s); };
  auto A_EMPTY = [&](int s) { return bar0 + 8u * (uint32_t)(2 * SB + 2 + s); };
  const uint32_t ACC_FULL = bar0 + 8u * (uint32_t)(2 * SB + 4);

  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&tmem_base_s)), "r"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 0) {
    for (int s = 0; s < SB; s++) { mbar_init(B_FULL(s), 1); mbar_init(B_EMPTY(s), 1); }
    for (int s = 0; s < 2; s++) { mbar_init(A_FULL(s), 128); mbar_init(A_EMPTY(s), 1); }
    mbar_init(ACC_FULL, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" :: "l"(&map_hi) : "memory");
    asm volatile("prefetch.tensormap [%0];" :: "l"(&map_lo) : "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_base_s;
  const uint32_t tmem_acc = tmem_base;                        // TN columns
  const uint32_t tmem_a = tmem_base + (uint32_t)TN;           // MODE 1: 2 x 64 columns
  const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(TN >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);

  if (warp == 0) {
    // ------------------------------------------------ TMA producer: W_hi / W_lo tiles of slab s into stage s % SB
    if (lane == 0) {
      for (int s = 0; s < nslabs; s++) {
        const int st = s % SB;
        mbar_wait(B_EMPTY(st), (((uint32_t)(s / SB)) & 1u) ^ 1u);
        mbar_expect_tx(B_FULL(st), B_STAGE);
        const uint32_t dst = smem_u32(b_st + (size_t)st * B_STAGE);
        tma_load_2d(dst, &map_hi, B_FULL(st), s * TK, n0);
        tma_load_2d(dst + B_TILE, &map_lo, B_FULL(st), s * TK, n0);
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------ MMA issuer
    if (lane == 0) {
      for (int s = 0; s < nslabs; s++) {
        const int st = s % SB, as = s & 1;
        mbar_wait(B_FULL(st), ((uint32_t)(s / SB)) & 1u);
        mbar_wait(A_FULL(as), ((uint32_t)(s >> 1)) & 1u);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t bh = smem_u32(b_st + (size_t)st * B_STAGE), bl = bh + B_TILE;
#pragma unroll
        for (int ks = 0; ks < TK / 8; ks++) {
          const uint64_t dbh = desc_sw128(bh + (uint32_t)ks * 32u), dbl = desc_sw128(bl + (uint32_t)ks * 32u);
          const uint32_t acc0 = (s > 0 || ks > 0) ? 1u : 0u;
          if (MODE == 1) {
            const uint32_t ah = tmem_a + (uint32_t)as * 64u + (uint32_t)ks * 8u, al = ah + 32u;
            // accmode 0: one accumulator.  1: the two correction products in their own accumulator (TN + 128 ...).
            // 2: also the main products alternate between two accumulators by k-step parity.
            if (accmode == 0) {
              mma_ts(tmem_acc, al, dbh, idesc, acc0);
              mma_ts(tmem_acc, ah, dbl, idesc, 1u);
              mma_ts(tmem_acc, ah, dbh, idesc, 1u);
            } else {
              const uint32_t cross = tmem_base + (uint32_t)TN + 128u, main2 = cross + (uint32_t)TN;
              mma_ts(cross, al, dbh, idesc, acc0);
              mma_ts(cross, ah, dbl, idesc, 1u);
              if (accmode == 2 && (ks & 1)) mma_ts(main2, ah, dbh, idesc, (s > 0 || ks > 1) ? 1u : 0u);
              else mma_ts(tmem_acc, ah, dbh, idesc, acc0);
            }
          } else {
            const uint32_t ah = smem_u32(a_st + (size_t)as * A_STAGE) + (uint32_t)ks * 2u * A_PANEL, al = ah + A_TILE;
            mma_ss(tmem_acc, desc_nosw(al, A_PANEL), dbh, idesc, acc0);
            mma_ss(tmem_acc, desc_nosw(ah, A_PANEL), dbl, idesc, 1u);
            mma_ss(tmem_acc, desc_nosw(ah, A_PANEL), dbh, idesc, 1u);
          }
        }
        umma_commit(B_EMPTY(st));
        umma_commit(A_EMPTY(as));
      }
      umma_commit(ACC_FULL);
    }
  } else {
    // ------------------------------------------------ A loaders (group g = even / odd slabs), then epilogue
    const int lw = warp - 2, g = lw >> 2, q = warp & 3;      // q = TMEM lane quarter this warp may access
    const int r = q * 32 + lane;                              // tile row owned by this thread
    const bool live = m0 + r < M;
    const float *arow = A + (size_t)(live ? m0 + r : 0) * K;
    float cur[32];
    auto fetch = [&](int s) {
      const int k0 = s * TK;
#pragma unroll
      for (int c = 0; c < 8; c++) {
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (live && k0 + c * 4 + 3 < K) v = *reinterpret_cast<const float4 *>(arow + k0 + c * 4);
        else if (live) { float t[4] = {0, 0, 0, 0}; for (int e = 0; e < 4; e++) if (k0 + c * 4 + e < K) t[e] = arow[k0 + c * 4 + e]; v = make_float4(t[0], t[1], t[2], t[3]); }
        cur[c * 4 + 0] = v.x; cur[c * 4 + 1] = v.y; cur[c * 4 + 2] = v.z; cur[c * 4 + 3] = v.w;
      }
    };
    if (g < nslabs) fetch(g);
    for (int s = g; s < nslabs; s += 2) {
      uint32_t hv[32];
      mbar_wait(A_EMPTY(g), (((uint32_t)(s >> 1)) & 1u) ^ 1u);
      if (MODE == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t t0 = tmem_a + ((uint32_t)(q * 32) << 16) + (uint32_t)g * 64u;
#pragma unroll
        for (int e = 0; e < 32; e++) hv[e] = to_tf32(cur[e]);
        tmem_st32(t0, hv);
#pragma unroll
        for (int e = 0; e < 32; e++) hv[e] = to_tf32(cur[e] - __uint_as_float(hv[e]));
        tmem_st32(t0 + 32u, hv);
        if (s + 2 < nslabs) fetch(s + 2);                     // next slab of this group in flight while the MMAs run
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      } else {
        unsigned char *ah = a_st + (size_t)g * A_STAGE, *al = ah + A_TILE;
#pragma unroll
        for (int e = 0; e < 32; e++) hv[e] = to_tf32(cur[e]);
#pragma unroll
        for (int c = 0; c < 8; c++)
          *reinterpret_cast<uint4 *>(ah + (uint32_t)c * A_PANEL + (uint32_t)r * 16u) = make_uint4(hv[c * 4], hv[c * 4 + 1], hv[c * 4 + 2], hv[c * 4 + 3]);
#pragma unroll
        for (int e = 0; e < 32; e++) hv[e] = to_tf32(cur[e] - __uint_as_float(hv[e]));
#pragma unroll
        for (int c = 0; c < 8; c++)
          *reinterpret_cast<uint4 *>(al + (uint32_t)c * A_PANEL + (uint32_t)r * 16u) = make_uint4(hv[c * 4], hv[c * 4 + 1], hv[c * 4 + 2], hv[c * 4 + 3]);
        if (s + 2 < nslabs) fetch(s + 2);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      }
      mbar_arrive(A_FULL(g));
    }
    // epilogue: quarter q, column half g
    mbar_wait(ACC_FULL, 0u);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int row = m0 + r;
    for (int c0 = g * (TN / 2); c0 < (g + 1) * (TN / 2); c0 += 16) {
      uint32_t v[16];
      const uint32_t taddr = tmem_acc + ((uint32_t)(q * 32) << 16) + (uint32_t)c0;
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
          : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
            "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
          : "r"(taddr));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      for (int extra = 0; extra < accmode; extra++) {           // main + (main2 +) cross, added in fp32 (round to nearest)
        uint32_t w[16];
        const uint32_t ta2 = taddr + (uint32_t)TN + 128u + (uint32_t)(extra == 0 ? 0 : TN);
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
            : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7]),
              "=r"(w[8]), "=r"(w[9]), "=r"(w[10]), "=r"(w[11]), "=r"(w[12]), "=r"(w[13]), "=r"(w[14]), "=r"(w[15])
            : "r"(ta2));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (accmode == 2 && extra == 0) continue;              // (order: main + main2 first, then cross)
        for (int j = 0; j < 16; j++) v[j] = __float_as_uint(__uint_as_float(v[j]) + __uint_as_float(w[j]));
      }
      if (accmode == 2) {                                      // cross last
        uint32_t w[16];
        const uint32_t ta2 = taddr + (uint32_t)TN + 128u;
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
            : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7]),
              "=r"(w[8]), "=r"(w[9]), "=r"(w[10]), "=r"(w[11]), "=r"(w[12]), "=r"(w[13]), "=r"(w[14]), "=r"(w[15])
            : "r"(ta2));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int j = 0; j < 16; j++) v[j] = __float_as_uint(__uint_as_float(v[j]) + __uint_as_float(w[j]));
      }
      if (row < M)
        for (int j = 0; j < 16; j++)
          if (n0 + c0 + j < N) C[(size_t)row * N + n0 + c0 + j] = __uint_as_float(v[j]);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "r"(TMEM_COLS) : "memory");
}

static float tf32_rna(float x) {                              // cvt.rna.tf32.f32 on the host
  uint32_t u; memcpy(&u, &x, 4);
  u = (u + 0x1000u) & 0xffffe000u;
  float r; memcpy(&r, &u, 4);
  return r;
}

static CUtensorMap make_map(const float *dptr, int N, int Kp, int TN) {
  CUtensorMap m;
  cuuint64_t dims[2] = {(cuuint64_t)Kp, (cuuint64_t)N};
  cuuint64_t strides[1] = {(cuuint64_t)Kp * 4};
  cuuint32_t box[2] = {32, (cuuint32_t)TN};
  cuuint32_t es[2] = {1, 1};
  CUresult rc = cuTensorMapEncodeTiled(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void *)dptr, dims, strides, box, es,
                                       CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                       CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (rc != CUDA_SUCCESS) { printf("cuTensorMapEncodeTiled failed: %d\n", (int)rc); exit(1); }
  return m;
}

template <int MODE, int TN, int SB>
static void run(int M, int N, int K, int accmode, int dist) {
  const int Kp = (K + 3) / 4 * 4;
  std::vector<float> hA((size_t)M * K), hW((size_t)N * K), hWh((size_t)N * Kp, 0.f), hWl((size_t)N * Kp, 0.f), hC((size_t)M * N);
  srand(1);
  for (auto &x : hA) x = (float)rand() / RAND_MAX - (dist == 1 ? 0.0f : 0.5f);   // dist 1: non-negative activations (after ReLU)
  for (auto &x : hW) x = (float)rand() / RAND_MAX - 0.5f;
  for (int n = 0; n < N; n++)
    for (int k = 0; k < K; k++) {
      float w = hW[(size_t)n * K + k], h = tf32_rna(w);
      hWh[(size_t)n * Kp + k] = h; hWl[(size_t)n * Kp + k] = tf32_rna(w - h);
    }
  float *dA, *dWh, *dWl, *dC;
  CHECK(cudaMalloc(&dA, hA.size() * 4)); CHECK(cudaMalloc(&dWh, hWh.size() * 4)); CHECK(cudaMalloc(&dWl, hWl.size() * 4)); CHECK(cudaMalloc(&dC, hC.size() * 4));
  CHECK(cudaMemcpy(dA, hA.data(), hA.size() * 4, cudaMemcpyHostToDevice));
  CHECK(cudaMemcpy(dWh, hWh.data(), hWh.size() * 4, cudaMemcpyHostToDevice));
  CHECK(cudaMemcpy(dWl, hWl.data(), hWl.size() * 4, cudaMemcpyHostToDevice));
  CHECK(cudaMemset(dC, 0xff, hC.size() * 4));
  CUtensorMap mh = make_map(dWh, N, Kp, TN), ml = make_map(dWl, N, Kp, TN);
  const size_t smem = (size_t)SB * 2 * TN * 128 + (MODE == 0 ? 2 * 2 * (TK / 4) * TM * 16 : 0) + 1024;
  CHECK(cudaFuncSetAttribute(gemm_tc5b<MODE, TN, SB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid((N + TN - 1) / TN, (M + TM - 1) / TM);
  gemm_tc5b<MODE, TN, SB><<<grid, NTHREADS, smem>>>(dA, dC, M, N, K, mh, ml, accmode);
  CHECK(cudaDeviceSynchronize());
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  const int reps = 20;
  for (int i = 0; i < reps; i++) gemm_tc5b<MODE, TN, SB><<<grid, NTHREADS, smem>>>(dA, dC, M, N, K, mh, ml, accmode);
  cudaEventRecord(e1);
  CHECK(cudaDeviceSynchronize());
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  CHECK(cudaMemcpy(hC.data(), dC, hC.size() * 4, cudaMemcpyDeviceToHost));
  double max_err = 0, max_ref = 0, se = 0, sr = 0;
  int bad = 0;
  for (int i = 0; i < M; i += (M > 512 ? 37 : 1))
    for (int j = 0; j < N; j += (N > 256 ? 11 : 1)) {
      double ref = 0;
      for (int k = 0; k < K; k++) ref += (double)hA[(size_t)i * K + k] * hW[(size_t)j * K + k];
      double e = fabs(ref - hC[(size_t)i * N + j]);
      if (!(e == e)) { bad++; e = 1e30; }
      max_err = fmax(max_err, e);
      max_ref = fmax(max_ref, fabs(ref));
      if (e < 1e29) { se += e * e; sr += ref * ref; }
    }
  printf("mode %d acc %d dist %d TN %d SB %d  M %d N %d K %d: max |err| / max |ref| = %.3e  rms err / rms ref = %.3e  nan %d   %.1f TFLOP/s fp32-equivalent (%.3f ms)\n",
         MODE, accmode, dist, TN, SB, M, N, K, max_err / max_ref, sqrt(se / sr), bad, 2.0 * M * N * K * reps / (ms * 1e-3) / 1e12, ms / reps);
  cudaFree(dA); cudaFree(dWh); cudaFree(dWl); cudaFree(dC);
}

int main(int argc, char **argv) {
  const int mode = argc > 1 ? atoi(argv[1]) : 1;
  const int M = argc > 2 ? atoi(argv[2]) : 16384, N = argc > 3 ? atoi(argv[3]) : 1536, K = argc > 4 ? atoi(argv[4]) : 320;
  const int accmode = argc > 5 ? atoi(argv[5]) : 0, dist = argc > 6 ? atoi(argv[6]) : 0;
  CHECK(cudaFree(0));
  if (mode == 0) run<0, 128, 2>(M, N, K, 0, dist);
  else if (mode == 1) run<1, 128, 3>(M, N, K, accmode == 2 ? 1 : accmode, dist);   // (three accumulators of 128 do not fit)
  else if (mode == 2) run<1, 96, 3>(M, N, K, accmode, dist);
  else if (mode == 3) run<1, 256, 2>(M, N, K, 0, dist);
  else if (mode == 4) run<1, 160, 3>(M, N, K, accmode == 2 ? 1 : accmode, dist);
  return 0;
}
