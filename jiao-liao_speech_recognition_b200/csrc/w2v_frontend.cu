// Raw-waveform (wav2vec2 / XLS-R) front end, data-movement side: the convolutions themselves run on the tcgen05
// GEMM; these kernels build its operands.
//   wave_stats_kernel     per-utterance mean and 1/sqrt(var + 1e-7) over the valid samples
//                         (SP/transformers/models/wav2vec2/feature_extraction_wav2vec2.py:78-97), two-pass, one CTA per utterance
//   wave_im2col_kernel    layer 0 of the feature encoder (Conv1d(1 → C, k = 10, s = 5), modeling_wav2vec2.py:275-299):
//                         normalises on the fly and writes the [B·T0, 16] bf16 im2col matrix (taps 10..15 zero) — the
//                         normalised waveform is never materialised
//   im2col_1d_kernel      [B, T_in, C] bf16 → [B·T_out, k·cg] for Conv1d(k, stride, pad) over the channel group [c0, c0 + cg):
//                         feature-encoder layers 1-6 (k = 3 / 2, s = 2) and the grouped positional convolution
//                         (k = 128, pad = 64, groups = 16, modeling_wav2vec2.py:326-368), tap-major columns
#include <algorithm>

#include "common.cuh"

namespace jl {

constexpr int WS_THREADS = 1024;

__device__ __forceinline__ float block_sum_1024(float v, float* red) {
  v = warp_sum(v);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  if (warp == 0) {
    float t = (lane < WS_THREADS / 32) ? red[lane] : 0.0f;
    t = warp_sum(t);
    if (lane == 0) red[32] = t;
  }
  __syncthreads();
  return red[32];
}

__global__ void __launch_bounds__(WS_THREADS) wave_stats_kernel(const float* __restrict__ wave, int64_t stride, const int32_t* __restrict__ num_samples,
                                                                int max_samples, float* __restrict__ stats) {
  jl::pdl_prologue();
  __shared__ float red[33];
  const int b = blockIdx.x;
  const int n = min(num_samples[b], max_samples);
  const float* x = wave + static_cast<int64_t>(b) * stride;
  float s = 0.0f;
  for (int i = threadIdx.x; i < n; i += WS_THREADS) s += x[i];
  const float mean = (n > 0) ? block_sum_1024(s, red) / static_cast<float>(n) : 0.0f;
  float q = 0.0f;
  for (int i = threadIdx.x; i < n; i += WS_THREADS) {
    const float d = x[i] - mean;
    q = fmaf(d, d, q);
  }
  const float var = (n > 0) ? block_sum_1024(q, red) / static_cast<float>(n) : 0.0f;
  if (threadIdx.x == 0) {
    stats[2 * b] = mean;
    stats[2 * b + 1] = 1.0f / sqrtf(var + 1e-7f);
  }
}

// out[(b, t), j] = (x[b, s·t + j] - mean_b) · rstd_b for j < k and s·t + j < n_b, else 0;   16 columns per row (k <= 16)
__global__ void wave_im2col_kernel(const float* __restrict__ wave, int64_t stride, const int32_t* __restrict__ num_samples, int max_samples,
                                   const float* __restrict__ stats, __nv_bfloat16* __restrict__ out, int batch, int t_out, int k, int s) {
  jl::pdl_prologue();
  const int64_t total = static_cast<int64_t>(batch) * t_out;
  for (int64_t r = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; r < total; r += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int b = static_cast<int>(r / t_out), t = static_cast<int>(r - static_cast<int64_t>(b) * t_out);
    const int n = min(num_samples[b], max_samples);
    const float mean = stats[2 * b], rstd = stats[2 * b + 1];
    const float* x = wave + static_cast<int64_t>(b) * stride + static_cast<int64_t>(t) * s;
    uint32_t w[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int i0 = 2 * j, i1 = 2 * j + 1;
      const float v0 = (i0 < k && t * s + i0 < n) ? (x[i0] - mean) * rstd : 0.0f;
      const float v1 = (i1 < k && t * s + i1 < n) ? (x[i1] - mean) * rstd : 0.0f;
      w[j] = pack_bf16x2(v0, v1);
    }
    st_global_v8(out + r * 16, w);
  }
}

// out[(b, t), j · cg + c] = x[b, stride·t - pad + j, c0 + c]   (zero outside [0, t_in)); 8 channels (16 B) per thread
__global__ void im2col_1d_kernel(const uint4* __restrict__ x, uint4* __restrict__ out, int batch, int t_in, int c8_total, int t_out, int k, int stride,
                                 int pad, int c0_8, int cg8) {
  jl::pdl_prologue();
  const int64_t total = static_cast<int64_t>(batch) * t_out * k * cg8;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int ch = static_cast<int>(i % cg8);
    int64_t r = i / cg8;
    const int tap = static_cast<int>(r % k);
    r /= k;
    const int t = static_cast<int>(r % t_out);
    const int b = static_cast<int>(r / t_out);
    const int ti = stride * t - pad + tap;
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (ti >= 0 && ti < t_in) v = __ldg(x + (static_cast<int64_t>(b) * t_in + ti) * c8_total + c0_8 + ch);
    out[i] = v;
  }
}

}  // namespace jl

extern "C" {

int jl_wave_stats(const float* wave, int64_t wave_stride, const int32_t* num_samples, int32_t batch, int32_t max_samples, float* stats, void* stream) {
  JL_REQUIRE(wave && num_samples && stats, JL_EINVAL, "wave_stats: null pointer");
  JL_REQUIRE(batch > 0 && max_samples > 0 && wave_stride >= max_samples, JL_EINVAL, "wave_stats: bad sizes");
  if (int rc = jl::check_device()) return rc;
  jl::launch(jl::wave_stats_kernel, batch, jl::WS_THREADS, 0, static_cast<cudaStream_t>(stream), wave, wave_stride, num_samples, max_samples, stats);
  JL_CHECK_LAUNCH("wave_stats");
  return JL_OK;
}

int jl_wave_im2col(const float* wave, int64_t wave_stride, const int32_t* num_samples, int32_t batch, int32_t max_samples, const float* stats,
                   void* out, int32_t t_out, int32_t kernel, int32_t stride, void* stream) {
  JL_REQUIRE(wave && num_samples && stats && out, JL_EINVAL, "wave_im2col: null pointer");
  JL_REQUIRE(batch > 0 && t_out > 0 && kernel > 0 && kernel <= 16 && stride > 0, JL_EINVAL, "wave_im2col: kernel must be 1..16, sizes positive");
  JL_REQUIRE((reinterpret_cast<uintptr_t>(out) & 31) == 0, JL_EINVAL, "wave_im2col: out must be 32-byte aligned");
  if (int rc = jl::check_device()) return rc;
  const int64_t rows = static_cast<int64_t>(batch) * t_out;
  int blocks = static_cast<int>(std::min<int64_t>((rows + 255) / 256, 148 * 16));
  jl::launch(jl::wave_im2col_kernel, blocks, 256, 0, static_cast<cudaStream_t>(stream), wave, wave_stride, num_samples, max_samples, stats,
             reinterpret_cast<__nv_bfloat16*>(out), batch, t_out, kernel, stride);
  JL_CHECK_LAUNCH("wave_im2col");
  return JL_OK;
}

int jl_im2col_1d(const void* x, void* out, int32_t batch, int32_t t_in, int32_t channels, int32_t t_out, int32_t kernel, int32_t stride, int32_t pad,
                 int32_t c0, int32_t cg, void* stream) {
  JL_REQUIRE(x && out, JL_EINVAL, "im2col_1d: null pointer");
  JL_REQUIRE(batch > 0 && t_in > 0 && t_out > 0 && kernel > 0 && stride > 0 && pad >= 0, JL_EINVAL, "im2col_1d: bad sizes");
  JL_REQUIRE((channels & 7) == 0 && (c0 & 7) == 0 && (cg & 7) == 0 && cg > 0 && c0 >= 0 && c0 + cg <= channels, JL_EINVAL,
             "im2col_1d: channels, c0 and cg must be multiples of 8 with c0 + cg <= channels");
  JL_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0, JL_EINVAL, "im2col_1d: 16-byte alignment");
  if (int rc = jl::check_device()) return rc;
  const int64_t total = static_cast<int64_t>(batch) * t_out * kernel * (cg >> 3);
  int blocks = static_cast<int>(std::min<int64_t>((total + 255) / 256, 148 * 16));
  jl::launch(jl::im2col_1d_kernel, blocks, 256, 0, static_cast<cudaStream_t>(stream), reinterpret_cast<const uint4*>(x), reinterpret_cast<uint4*>(out), batch,
             t_in, channels >> 3, t_out, kernel, stride, pad, c0 >> 3, cg >> 3);
  JL_CHECK_LAUNCH("im2col_1d");
  return JL_OK;
}

}  // extern "C"
