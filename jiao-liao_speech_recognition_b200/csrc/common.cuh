// Shared host/device helpers for libjl_b200.so (sm_100a only).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/jl_b200.h"

namespace jl {

void set_error(const char* fmt, ...);
void count_launch(int n = 1);
int check_device();   // JL_OK when the current device is sm_100
int num_sms();
// 2-D bf16 tensor map: `inner` contiguous elements × `outer` rows (row stride ld elements), 128-byte swizzle, box = 64 × box_rows.
int make_tma_map_2d_bf16(CUtensorMap* map, const void* ptr, int64_t inner, int64_t outer, int64_t ld, int box_rows);

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

// Programmatic dependent launch: every kernel is launched with programmatic stream serialisation, lets its successor start
// launching at once, and waits for its predecessor's memory before touching global data.  The launch latency and CTA ramp of
// kernel N+1 then overlap the tail of kernel N (≈ 380 launches per fine-tune step).
#ifndef JL_PDL_EARLY_TRIGGER
#define JL_PDL_EARLY_TRIGGER 0   // 1: let the dependent grid launch as soon as every CTA has started (measured slower); 0: at CTA exit
#endif
__device__ __forceinline__ void pdl_launch_dependents() {
#if JL_PDL_EARLY_TRIGGER
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
#endif
}
// Late trigger: issued once a CTA has requested its last operand tile, so the dependent grid's launch overlaps the tail
// (last MMAs + epilogue) of this one.  Compile with -DJL_PDL_LATE_TRIGGER=0 to leave the trigger to CTA exit.
#ifndef JL_PDL_LATE_TRIGGER
#define JL_PDL_LATE_TRIGGER 1
#endif
__device__ __forceinline__ void pdl_trigger_late() {
#if JL_PDL_LATE_TRIGGER
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
#endif
}
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_prologue() {
  pdl_launch_dependents();
  pdl_wait();
}
extern int g_use_pdl;

template <typename... KArgs, typename... Args>
inline cudaError_t launch(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = g_use_pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

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

// 256-bit global accesses (sm_100): one request per 32-byte sector instead of two 16-byte halves.  32-byte aligned only.
__device__ __forceinline__ void st_global_v8(void* p, const uint32_t (&v)[8]) {
  asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]),
               "r"(v[5]), "r"(v[6]), "r"(v[7])
               : "memory");
}
__device__ __forceinline__ void ld_global_nc_v8(const void* p, uint32_t (&v)[8]) {
  asm volatile("ld.global.nc.v8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
               : "l"(p));
}

__device__ __forceinline__ float bf16_bits_to_float(uint32_t bits16) { return __uint_as_float(bits16 << 16); }
__device__ __forceinline__ float2 unpack_bf16x2(uint32_t v) {
  return make_float2(__uint_as_float(v << 16), __uint_as_float(v & 0xffff0000u));
}
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&t);
}

// erf-form GELU (SP/transformers/activations.py "gelu") and its derivative.  erf via Abramowitz & Stegun 7.1.26
// (|error| <= 1.5e-7 — far below the bf16 rounding of the stored result).  With s = |v|·√(log2(e)/2):
//   q(v) = 0.5 · (1 − erf(|v|/√2)) = 0.5 · poly(t) · 2^(−s²),  t = 1 / (1 + (p/√log2(e))·s)
//   Φ(v) = v >= 0 ? 1 − q : q;   gelu(v) = v · Φ(v);   gelu'(v) = Φ(v) + v · exp(−v²/2) / √(2π)
// The two MUFU operations are the raw rcp.approx / ex2.approx (their arguments are >= 1 and <= 0: none of the denormal
// guards of __fdividef / exp2f is needed), which leaves 12 issue slots per element — the GEMM epilogue has to stay under
// the ≈ 40 slots per output element a K = 768 mainloop gives it (profiles/r1m_gemm_epilogues.md).
__device__ __forceinline__ float mufu_rcp(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float mufu_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void gelu_parts(float v, float& cdf, float& e) {
  const float s = fabsf(v) * 0.8493218002880191f;
  const float t = mufu_rcp(fmaf(0.2727374808792225f, s, 1.0f));
  float poly = fmaf(1.061405429f, t, -1.453152027f);
  poly = fmaf(poly, t, 1.421413741f);
  poly = fmaf(poly, t, -0.284496736f);
  poly = fmaf(poly, t, 0.254829592f);
  poly *= t;
  e = mufu_ex2(-s * s);                            // exp(−v²/2)
  const float q = 0.5f * poly * e;
  cdf = (v >= 0.0f) ? 1.0f - q : q;
}
__device__ __forceinline__ float gelu_erf(float x) {
  float cdf, e;
  gelu_parts(x, cdf, e);
  return x * cdf;
}
__device__ __forceinline__ float gelu_erf_grad(float x) {
  float cdf, e;
  gelu_parts(x, cdf, e);
  return fmaf(x * 0.39894228040143268f, e, cdf);
}
// value and derivative from one evaluation (the forward FFN epilogue that saves gelu' for the backward pass)
__device__ __forceinline__ void gelu_erf_both(float x, float& y, float& dy) {
  float cdf, e;
  gelu_parts(x, cdf, e);
  y = x * cdf;
  dy = fmaf(x * 0.39894228040143268f, e, cdf);
}
// logistic function with one MUFU.EX2 and one MUFU.RCP (GLU gate)
__device__ __forceinline__ float sigmoid_fast(float x) {
  return mufu_rcp(1.0f + mufu_ex2(-1.4426950408889634f * x));      // x → −∞: ex2 = +inf, rcp(+inf) = 0
}

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

}  // namespace jl
