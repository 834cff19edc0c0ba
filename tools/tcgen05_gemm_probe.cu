// tools/tcgen05_gemm_probe.cu -- NOT part of the product build (kaldi_b200/build.py does not compile it).
//
// Stand-alone probe for the round-2 nnet3 GEMM: C[M x N] = A[M x K] * B[N x K]^T in fp32-equivalent precision
// on the 5th-generation tensor cores (tcgen05.mma kind::tf32, 3xTF32 split, accumulator in TMEM), written with raw
// PTX only.  It exists so that the first GPU minutes of round 2 can go to validating descriptors and the
// canonical shared-memory layout instead of writing boilerplate:
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o /tmp/tc5 tools/tcgen05_gemm_probe.cu && /tmp/tc5
//
// prints the max relative error against a double-precision CPU product and the TFLOP/s of the kernel.
// STATUS: compiles (ptxas accepts every tcgen05 form used); NEVER RUN -- no GPU time was left in round 1.
//
// Shape of the kernel (deliberately the simplest correct structure, single-buffered):
//   one CTA = 128 threads = one 128 x 128 output tile, accumulator = 128 TMEM lanes x 128 fp32 columns;
//   per K slab of 32: all threads stage A and B (global fp32 -> hi/lo TF32 -> shared, canonical K-major
//   no-swizzle layout), fence to the async proxy, one thread issues 4 x 3 MMAs and commits to an mbarrier,
//   everybody waits on it before the slab buffers are overwritten;
//   epilogue: warp w reads TMEM lanes [32w, 32w+32) with tcgen05.ld.32x32b and stores the rows.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cmath>
#include <cuda_runtime.h>

#define CHECK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

constexpr int TM = 128, TN = 128, TK = 32;            // CTA tile; K slab
constexpr int CHUNKS = TK / 4;                          // 16-byte K chunks per slab (4 tf32 each)
constexpr uint32_t PANEL_BYTES = TM * 16;               // one K chunk of all 128 rows: LBO
constexpr uint32_t TILE_BYTES = CHUNKS * PANEL_BYTES;   // 16 KB per operand tile

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint32_t to_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return r;
}

// K-major, no swizzle: start >> 4 | LBO >> 4 << 16 | SBO >> 4 << 32 | version 1 << 46 | layout 0 << 61
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3fff);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}

// F32 accumulate, TF32 x TF32, both K-major, N = 128, M = 128
__device__ __forceinline__ uint32_t make_idesc() {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(TN >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);
}

__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      :: "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}

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

// element (row r, k) of a staged operand tile: panel k/4, 16 bytes per row, 4 bytes per element
__device__ __forceinline__ uint32_t tile_off(int r, int k) { return (uint32_t)(k >> 2) * PANEL_BYTES + (uint32_t)r * 16u + (uint32_t)(k & 3) * 4u; }

__global__ void __launch_bounds__(128) gemm_tc5(const float *A, const float *B, float *C, int M, int N, int K) {
  extern __shared__ __align__(1024) unsigned char smem[];
  unsigned char *a_hi = smem, *a_lo = smem + TILE_BYTES, *b_hi = smem + 2 * TILE_BYTES, *b_lo = smem + 3 * TILE_BYTES;
  __shared__ __align__(8) unsigned long long bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int m0 = blockIdx.y * TM, n0 = blockIdx.x * TN;

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&tmem_base_s)), "r"(TN) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_d = tmem_base_s;
  const uint32_t idesc = make_idesc();
  uint32_t parity = 0;

  for (int k0 = 0; k0 < K; k0 += TK) {
    // stage: 128 rows x 32 k of A and B; lane -> (row within 8 = lane % 8, chunk pair = lane / 8), conflict-free 16-byte stores
    for (int it = 0; it < (TM / 8) * (CHUNKS / 4) / 4; it++) {        // 16 row groups x 2 chunk groups = 32 units, 4 warps
      const int unit = it * 4 + warp;
      const int rg = unit >> 1, cg = unit & 1;
      const int r = rg * 8 + (lane & 7), ch = cg * 4 + (lane >> 3);
      float4 va = make_float4(0, 0, 0, 0), vb = make_float4(0, 0, 0, 0);
      if (m0 + r < M) va = *reinterpret_cast<const float4 *>(A + (size_t)(m0 + r) * K + k0 + ch * 4);
      if (n0 + r < N) vb = *reinterpret_cast<const float4 *>(B + (size_t)(n0 + r) * K + k0 + ch * 4);
      const float xa[4] = {va.x, va.y, va.z, va.w}, xb[4] = {vb.x, vb.y, vb.z, vb.w};
      uint4 ah, al, bh, bl;
      uint32_t *pah = &ah.x, *pal = &al.x, *pbh = &bh.x, *pbl = &bl.x;
#pragma unroll
      for (int e = 0; e < 4; e++) {
        pah[e] = to_tf32(xa[e]); pal[e] = to_tf32(xa[e] - __uint_as_float(pah[e]));
        pbh[e] = to_tf32(xb[e]); pbl[e] = to_tf32(xb[e] - __uint_as_float(pbh[e]));
      }
      const uint32_t off = tile_off(r, ch * 4);
      *reinterpret_cast<uint4 *>(a_hi + off) = ah; *reinterpret_cast<uint4 *>(a_lo + off) = al;
      *reinterpret_cast<uint4 *>(b_hi + off) = bh; *reinterpret_cast<uint4 *>(b_lo + off) = bl;
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");       // generic-proxy stores -> visible to the tensor core
    __syncthreads();
    if (tid == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
      for (int ks = 0; ks < TK / 8; ks++) {
        const uint32_t adv = (uint32_t)ks * 2u * PANEL_BYTES;            // two 16-byte chunks per K = 8 instruction
        const uint64_t dah = make_desc(smem_u32(a_hi) + adv, PANEL_BYTES, 128), dal = make_desc(smem_u32(a_lo) + adv, PANEL_BYTES, 128);
        const uint64_t dbh = make_desc(smem_u32(b_hi) + adv, PANEL_BYTES, 128), dbl = make_desc(smem_u32(b_lo) + adv, PANEL_BYTES, 128);
        mma_tf32(tmem_d, dal, dbh, idesc, (k0 > 0 || ks > 0) ? 1u : 0u);
        mma_tf32(tmem_d, dah, dbl, idesc, 1u);
        mma_tf32(tmem_d, dah, dbh, idesc, 1u);
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(&bar)) : "memory");
    }
    mbar_wait(smem_u32(&bar), parity);                                  // MMAs of this slab retired: buffers reusable, D complete
    parity ^= 1;
  }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

  // epilogue: warp w owns TMEM lanes 32w .. 32w+31 = output rows m0 + 32w + lane
  const int row = m0 + warp * 32 + lane;
  for (int c0 = 0; c0 < TN; c0 += 16) {
    uint32_t v[16];
    const uint32_t taddr = tmem_d + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0;
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    if (row < M)
      for (int j = 0; j < 16; j++)
        if (n0 + c0 + j < N) C[(size_t)row * N + n0 + c0 + j] = __uint_as_float(v[j]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_d), "r"(TN) : "memory");
}

int main() {
  const int M = 1024, N = 768, K = 1536;
  std::vector<float> hA((size_t)M * K), hB((size_t)N * K), hC((size_t)M * N);
  srand(1);
  for (auto &x : hA) x = (float)rand() / RAND_MAX - 0.5f;
  for (auto &x : hB) x = (float)rand() / RAND_MAX - 0.5f;
  float *dA, *dB, *dC;
  CHECK(cudaMalloc(&dA, hA.size() * 4)); CHECK(cudaMalloc(&dB, hB.size() * 4)); CHECK(cudaMalloc(&dC, hC.size() * 4));
  CHECK(cudaMemcpy(dA, hA.data(), hA.size() * 4, cudaMemcpyHostToDevice));
  CHECK(cudaMemcpy(dB, hB.data(), hB.size() * 4, cudaMemcpyHostToDevice));
  const size_t smem = 4 * TILE_BYTES + 1024;
  CHECK(cudaFuncSetAttribute(gemm_tc5, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid(N / TN, M / TM);
  gemm_tc5<<<grid, 128, smem>>>(dA, dB, dC, M, N, K);
  CHECK(cudaDeviceSynchronize());
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  for (int i = 0; i < 10; i++) gemm_tc5<<<grid, 128, smem>>>(dA, dB, dC, M, N, K);
  cudaEventRecord(e1);
  CHECK(cudaDeviceSynchronize());
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  CHECK(cudaMemcpy(hC.data(), dC, hC.size() * 4, cudaMemcpyDeviceToHost));
  double max_err = 0, max_ref = 0;
  for (int i = 0; i < M; i += 37)
    for (int j = 0; j < N; j += 11) {
      double ref = 0;
      for (int k = 0; k < K; k++) ref += (double)hA[(size_t)i * K + k] * hB[(size_t)j * K + k];
      max_err = fmax(max_err, fabs(ref - hC[(size_t)i * N + j]));
      max_ref = fmax(max_ref, fabs(ref));
    }
  printf("max |err| / max |ref| = %.3e   (3xTF32 target: ~1e-6)   %.1f TFLOP/s fp32-equivalent\n",
         max_err / max_ref, 2.0 * M * N * K * 10 / (ms * 1e-3) / 1e12);
  return 0;
}
