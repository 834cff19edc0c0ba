// common.cuh — shared helpers for the b2k CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <atomic>
#include <string>

#include "b2k.h"

namespace b2k {

extern thread_local std::string g_last_error;
extern std::atomic<int64_t> g_launch_count;

inline int set_error(int code, const char *what, const char *detail = nullptr) {
  g_last_error = what;
  if (detail) { g_last_error += ": "; g_last_error += detail; }
  return code;
}

#define B2K_CUDA_CHECK(expr)                                                     \
  do {                                                                           \
    cudaError_t _e = (expr);                                                     \
    if (_e != cudaSuccess) {                                                     \
      char _buf[256];                                                            \
      snprintf(_buf, sizeof(_buf), "%s at %s:%d", cudaGetErrorString(_e), __FILE__, __LINE__); \
      return ::b2k::set_error(B2K_ERR_CUDA, #expr, _buf);                        \
    }                                                                            \
  } while (0)

#define B2K_LAUNCH_CHECK()                                                       \
  do {                                                                           \
    ::b2k::g_launch_count.fetch_add(1, std::memory_order_relaxed);               \
    B2K_CUDA_CHECK(cudaGetLastError());                                          \
  } while (0)

// Fails loudly when there is no Blackwell device: the product has no CPU path.
int require_device();

// order-preserving float <-> uint32 (for unsigned atomicMin on costs)
__host__ __device__ __forceinline__ uint32_t f2ord(float f) {
#ifdef __CUDA_ARCH__
  uint32_t u = __float_as_uint(f);
#else
  uint32_t u; memcpy(&u, &f, 4);
#endif
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__host__ __device__ __forceinline__ float ord2f(uint32_t o) {
  uint32_t u = (o & 0x80000000u) ? (o & 0x7fffffffu) : ~o;
#ifdef __CUDA_ARCH__
  return __uint_as_float(u);
#else
  float f; memcpy(&f, &u, 4); return f;
#endif
}
#define B2K_INF_ORD 0xff800000u   // f2ord(+inf)

}  // namespace b2k
