// Error reporting, launch accounting and device check for libjl_b200.so.
#include <atomic>
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

namespace jl {

static thread_local char g_err[512] = "";
static std::atomic<int64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

int check_device() {
  static thread_local int cached_dev = -1;
  static thread_local int cached_ok = 0;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) {
    set_error("no CUDA device (libjl_b200 has no CPU fallback)");
    return JL_ECUDA;
  }
  if (dev == cached_dev) return cached_ok ? JL_OK : JL_EUNSUPPORTED;
  int major = 0;
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  cached_dev = dev;
  cached_ok = (major == 10);
  if (!cached_ok) {
    set_error("device %d has compute capability %d.x; libjl_b200 is built for sm_100a only", dev, major);
    return JL_EUNSUPPORTED;
  }
  return JL_OK;
}

}  // namespace jl

extern "C" {
int jl_version(void) { return JL_VERSION; }
const char* jl_last_error(void) { return jl::g_err; }
int64_t jl_launch_count(void) { return jl::g_launches.load(); }
void jl_launch_count_reset(void) { jl::g_launches.store(0); }
}
