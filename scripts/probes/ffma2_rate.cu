// Issue-rate probe (tuning aid): dependent-chain-free FFMA vs FFMA2 (fma.rn.f32x2) throughput per SM on sm_100a.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) {
  float2 d;
  asm("{ .reg .b64 ra, rb, rc, rd; mov.b64 ra, {%2, %3}; mov.b64 rb, {%4, %5}; mov.b64 rc, {%6, %7}; fma.rn.f32x2 rd, ra, rb, rc; mov.b64 {%0, %1}, rd; }"
      : "=f"(d.x), "=f"(d.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
  return d;
}
template <int MODE>
__global__ void __launch_bounds__(256) probe(float* out, int iters, float seed) {
  float2 a[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = make_float2(seed + i + threadIdx.x, seed - i);
  const float2 m = make_float2(1.0001f, 0.9999f), c = make_float2(0.5f, 0.25f);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0) { a[i].x = fmaf(a[i].x, m.x, c.x); a[i].y = fmaf(a[i].y, m.y, c.y); }
      else a[i] = fma2(a[i], m, c);
    }
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += a[i].x + a[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
  float* out; cudaMalloc(&out, 148 * 8 * 256 * 4);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 20000;
  for (int mode = 0; mode < 2; ++mode) {
    for (int rep = 0; rep < 2; ++rep) {
      cudaEventRecord(e0);
      if (mode == 0) probe<0><<<148 * 8, 256>>>(out, iters, 1.0f); else probe<1><<<148 * 8, 256>>>(out, iters, 1.0f);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      const double fma = 148.0 * 8 * 256 * 16.0 * iters;
      printf("%s: %.3f ms, %.1f TFLOP/s fp32 (2 flop per fma)\n", mode ? "FFMA2" : "FFMA ", ms, 2 * fma / ms * 1e-9);
    }
  }
  return 0;
}
