// api.cu — library-level entry points and process-wide state of liblcr.
#include "common.cuh"

namespace lcr {
thread_local int g_last_cuda_error = 0;
std::atomic<uint64_t> g_launch_count{0};

int sm_count() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}
}  // namespace lcr

extern "C" int lcr_version(void) { return LCR_VERSION; }

extern "C" const char* lcr_error_string(int status) {
  switch (status) {
    case LCR_OK: return "ok";
    case LCR_ERR_INVALID_ARG: return "invalid argument (null pointer, non-positive size or unsupported parameter)";
    case LCR_ERR_CAPACITY: return "size exceeds a compiled-in capacity (LCR_MAX_*)";
    case LCR_ERR_WORKSPACE: return "workspace missing, misaligned or too small";
    case LCR_ERR_ALIGNMENT: return "pointer or stride alignment requirement not met";
    case LCR_ERR_CUDA: return "CUDA launch failed (see lcr_last_cuda_error)";
    case LCR_ERR_NO_DEVICE: return "no sm_100 CUDA device available";
    default: return "unknown lcr status";
  }
}

extern "C" int lcr_last_cuda_error(void) { return lcr::g_last_cuda_error; }
extern "C" uint64_t lcr_launch_count(void) { return lcr::g_launch_count.load(std::memory_order_relaxed); }
