// Library-level entry points of the C ABI (include/selfocc_b200.h).
#include "common.cuh"
#include <atomic>

namespace so {
thread_local int g_last_cuda_error = 0;
static std::atomic<long long> g_launches{0};
void note_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
}  // namespace so

extern "C" int so_abi_version(void) { return SO_ABI_VERSION; }
extern "C" int so_last_cuda_error(void) { return so::g_last_cuda_error; }
extern "C" int64_t so_launch_count(void) { return so::g_launches.load(std::memory_order_relaxed); }
extern "C" const char* so_error_string(int code) {
  switch (code) {
    case SO_OK: return "ok";
    case SO_ERR_INVALID_ARG: return "invalid argument";
    case SO_ERR_UNSUPPORTED: return "unsupported configuration";
    case SO_ERR_CUDA: return "CUDA runtime error";
    case SO_ERR_NO_DEVICE: return "no CUDA device";
    default: return "unknown error code";
  }
}
