// Shared host/device helpers for libjl_b200.so (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/jl_b200.h"

namespace jl {

void set_error(const char* fmt, ...);
void count_launch(int n = 1);
int check_device();   // JL_OK when the current device is sm_100

#define JL_REQUIRE(cond, code, ...)       \
  do {                                    \
    if (!(cond)) {                        \
      jl::set_error(__VA_ARGS__);         \
      return (code);                      \
    }                                     \
  } while (0)

#define JL_CHECK_LAUNCH(name)                                                   \
  do {                                                                          \
    cudaError_t e__ = cudaGetLastError();                                       \
    if (e__ != cudaSuccess) {                                                   \
      jl::set_error("%s: launch failed: %s", name, cudaGetErrorString(e__));    \
      return JL_ECUDA;                                                          \
    }                                                                           \
    jl::count_launch();                                                         \
  } while (0)

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

__device__ __forceinline__ float bf16_bits_to_float(uint32_t bits16) { return __uint_as_float(bits16 << 16); }
__device__ __forceinline__ float2 unpack_bf16x2(uint32_t v) {
  return make_float2(__uint_as_float(v << 16), __uint_as_float(v & 0xffff0000u));
}
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&t);
}

__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f)); }
__device__ __forceinline__ float gelu_erf_grad(float x) {
  const float cdf = 0.5f * (1.0f + erff(x * 0.70710678118654752f));
  const float pdf = 0.39894228040143268f * expf(-0.5f * x * x);
  return cdf + x * pdf;
}

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

}  // namespace jl
