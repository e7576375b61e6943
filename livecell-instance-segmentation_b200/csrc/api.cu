// api.cu — library-level entry points and process-wide state of liblcr.
#include <stdlib.h>
#include <string.h>

#include <mutex>
#include <string>

#include "common.cuh"

extern char** environ;

namespace lcr {
// ---- tuning switches ------------------------------------------------------------------------------------------------
// The alternative kernels kept for A/B runs and tests are selected by LCR_* switches.  The environment is read ONCE, the
// first time a switch is consulted (no getenv on the launch paths); lcr_set_tuning() changes a switch at run time (tests,
// tools/bench_kernels.py).
namespace {
struct TuneStore {
  std::mutex mu;
  bool loaded = false;
  static constexpr int kMax = 64;
  std::string name[kMax], value[kMax];
  int n = 0;
  void load_env() {
    if (loaded) return;
    loaded = true;
    for (char** e = environ; e && *e; ++e) {
      if (strncmp(*e, "LCR_", 4) != 0) continue;
      const char* eq = strchr(*e, '=');
      if (!eq || n >= kMax) continue;
      name[n] = std::string(*e, eq - *e);
      value[n] = std::string(eq + 1);
      ++n;
    }
  }
  int find(const char* key) {
    for (int i = 0; i < n; ++i)
      if (name[i] == key) return i;
    return -1;
  }
};
TuneStore g_tune;
}  // namespace

const char* tune_get(const char* key) {
  static thread_local std::string buf;
  std::lock_guard<std::mutex> lock(g_tune.mu);
  g_tune.load_env();
  const int i = g_tune.find(key);
  if (i < 0) return nullptr;
  buf = g_tune.value[i];
  return buf.c_str();
}

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

extern "C" int lcr_set_tuning(const char* key, const char* value) {
  using namespace lcr;
  if (!key || strncmp(key, "LCR_", 4) != 0) return LCR_ERR_INVALID_ARG;
  std::lock_guard<std::mutex> lock(g_tune.mu);
  g_tune.load_env();
  int i = g_tune.find(key);
  if (!value) {  // unset: back to the built-in default
    if (i >= 0) {
      g_tune.name[i] = g_tune.name[g_tune.n - 1];
      g_tune.value[i] = g_tune.value[g_tune.n - 1];
      --g_tune.n;
    }
    return LCR_OK;
  }
  if (i < 0) {
    if (g_tune.n >= TuneStore::kMax) return LCR_ERR_CAPACITY;
    i = g_tune.n++;
    g_tune.name[i] = key;
  }
  g_tune.value[i] = value;
  return LCR_OK;
}
