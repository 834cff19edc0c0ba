// common.cu — process-wide state of the b2k library.
#include "common.cuh"

namespace b2k {
thread_local std::string g_last_error;
std::atomic<int64_t> g_launch_count{0};

int require_device() {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0)
    return set_error(B2K_ERR_NO_DEVICE, "no CUDA device: the b2k library has no CPU path",
                     e != cudaSuccess ? cudaGetErrorString(e) : nullptr);
  int dev = 0;
  cudaGetDevice(&dev);
  int major = 0;
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  if (major != 10)
    return set_error(B2K_ERR_NO_DEVICE, "b2k kernels are built for sm_100a only");
  return B2K_OK;
}
}  // namespace b2k

using namespace b2k;

extern "C" {

const char *b2k_last_error(void) { return g_last_error.c_str(); }
int b2k_version(void) { return 100; }
int64_t b2k_kernel_launch_count(void) { return g_launch_count.load(); }

}  // extern "C"
