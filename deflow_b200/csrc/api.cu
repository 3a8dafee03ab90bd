// Library-level plumbing of the C ABI: error strings, launch accounting, device properties.
#include "common.cuh"
#include "../../include/deflow_b200.h"

#include <atomic>
#include <stdarg.h>

namespace dfb {

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

void add_launches(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    return DFB_ERR_CUDA;
  }
  return DFB_OK;
}

int sm_count() {
  static int cached[64] = {0};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64) dev = 0;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

}  // namespace dfb

extern "C" const char* dfb_last_error(void) { return dfb::g_err; }
extern "C" int dfb_version(void) { return 100; }
extern "C" long long dfb_launch_count(void) { return dfb::g_launches.load(std::memory_order_relaxed); }
