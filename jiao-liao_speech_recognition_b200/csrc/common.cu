// Error reporting, launch accounting and device check for libjl_b200.so.
#include <atomic>
#include <cudaTypedefs.h>
#include <mutex>
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

int g_use_pdl = 1;   // measured on B200 (profiles/README.md): programmatic edges with the trigger at the END of a CTA's loads make the 383-kernel step 5 % faster (6.85 vs 7.22 ms); a trigger at CTA start made it 4.8 % slower

int num_sms() {
  static thread_local int cached_dev = -1, cached = 0;
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev != cached_dev) {
    cudaDeviceGetAttribute(&cached, cudaDevAttrMultiProcessorCount, dev);
    cached_dev = dev;
  }
  return cached;
}

static PFN_cuTensorMapEncodeTiled get_encode_fn() {
  static PFN_cuTensorMapEncodeTiled fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled>(p);
  });
  return fn;
}

int make_tma_map_2d_bf16(CUtensorMap* map, const void* ptr, int64_t inner, int64_t outer, int64_t ld, int box_rows) {
  PFN_cuTensorMapEncodeTiled enc = get_encode_fn();
  JL_REQUIRE(enc != nullptr, JL_ECUDA, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t dims[2] = {static_cast<cuuint64_t>(inner), static_cast<cuuint64_t>(outer)};
  cuuint64_t strides[1] = {static_cast<cuuint64_t>(ld) * 2};
  cuuint32_t box[2] = {64u, static_cast<cuuint32_t>(box_rows)};
  cuuint32_t estr[2] = {1u, 1u};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  JL_REQUIRE(r == CUDA_SUCCESS, JL_ECUDA, "cuTensorMapEncodeTiled failed (%d): ptr=%p inner=%lld outer=%lld ld=%lld", (int)r, ptr,
             (long long)inner, (long long)outer, (long long)ld);
  return JL_OK;
}

}  // namespace jl

extern "C" {
int jl_version(void) { return JL_VERSION; }
const char* jl_last_error(void) { return jl::g_err; }
int64_t jl_launch_count(void) { return jl::g_launches.load(); }
void jl_launch_count_reset(void) { jl::g_launches.store(0); }
void jl_debug_set_pdl(int on) { jl::g_use_pdl = on ? 1 : 0; }
}
