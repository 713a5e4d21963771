// Data-movement / elementwise helpers on the path (all HBM-bound, 16-byte vectorised, coalesced):
//   im2col for the two Conv1d(k=5, s=2, p=2) subsampling layers (SP/transformers/models/speech_to_text/
//   modeling_speech_to_text.py:82-99) so they run on the tcgen05 GEMM; ×sqrt(d) + sinusoidal positions + pad-row
//   zeroing (:542,568-579,123-139); bf16 transpose; column sums for bias gradients; casts; fused AdamW over the
//   flat adapter bucket (torch.optim.AdamW update order).
#include "common.cuh"
#include "ptx_sm100.cuh"

namespace jl {

// out[(b, t), tap * c + ch] = x[b, 2 t - 2 + tap, ch]   (zero outside [0, t_in))
__global__ void im2col_k5s2_kernel(const uint4* __restrict__ x, uint4* __restrict__ out, int batch, int t_in, int c8, int t_out) {
  jl::pdl_prologue();
  const int64_t total = static_cast<int64_t>(batch) * t_out * 5 * c8;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int ch = static_cast<int>(i % c8);
    int64_t r = i / c8;
    const int tap = static_cast<int>(r % 5);
    r /= 5;
    const int t = static_cast<int>(r % t_out);
    const int b = static_cast<int>(r / t_out);
    const int ti = 2 * t - 2 + tap;
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (ti >= 0 && ti < t_in) v = __ldg(x + (static_cast<int64_t>(b) * t_in + ti) * c8 + ch);
    out[i] = v;
  }
}

// h[b,t,:] = h[b,t,:] * scale + pos[t + 2, :]  for t < len_b;  0 for t >= len_b
__global__ void embed_positions_kernel(__nv_bfloat16* __restrict__ h, float scale, const float* __restrict__ pos,
                                       const int32_t* __restrict__ lengths, int batch, int seq, int d) {
  jl::pdl_prologue();
  const int d8 = d >> 3;
  const int64_t total = static_cast<int64_t>(batch) * seq * d8;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int ch = static_cast<int>(i % d8);
    const int64_t row = i / d8;
    const int t = static_cast<int>(row % seq);
    const int b = static_cast<int>(row / seq);
    uint4* p = reinterpret_cast<uint4*>(h) + i;
    uint4 o = make_uint4(0u, 0u, 0u, 0u);
    if (t < lengths[b]) {
      const uint4 v = *p;
      const float4 p0 = __ldg(reinterpret_cast<const float4*>(pos + static_cast<int64_t>(t + 2) * d + ch * 8));
      const float4 p1 = __ldg(reinterpret_cast<const float4*>(pos + static_cast<int64_t>(t + 2) * d + ch * 8) + 1);
      const float2 f0 = unpack_bf16x2(v.x), f1 = unpack_bf16x2(v.y), f2 = unpack_bf16x2(v.z), f3 = unpack_bf16x2(v.w);
      o.x = pack_bf16x2(fmaf(f0.x, scale, p0.x), fmaf(f0.y, scale, p0.y));
      o.y = pack_bf16x2(fmaf(f1.x, scale, p0.z), fmaf(f1.y, scale, p0.w));
      o.z = pack_bf16x2(fmaf(f2.x, scale, p1.x), fmaf(f2.y, scale, p1.y));
      o.w = pack_bf16x2(fmaf(f3.x, scale, p1.z), fmaf(f3.y, scale, p1.w));
    }
    *p = o;
  }
}

// out[cu[b] + t, :] = h[b, t, :] * scale + pos[t + 2, :] for t < cu[b+1] - cu[b]: the embedding and the entry into the packed layout
__global__ void embed_positions_packed_kernel(const __nv_bfloat16* __restrict__ h, __nv_bfloat16* __restrict__ out, float scale,
                                              const float* __restrict__ pos, const int32_t* __restrict__ cu, int batch, int seq, int d) {
  jl::pdl_prologue();
  const int d8 = d >> 3;
  const int64_t total = static_cast<int64_t>(batch) * seq * d8;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int ch = static_cast<int>(i % d8);
    const int64_t row = i / d8;
    const int t = static_cast<int>(row % seq);
    const int b = static_cast<int>(row / seq);
    const int r0 = cu[b];
    if (t >= cu[b + 1] - r0) continue;
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(h) + i);
    const float4 p0 = __ldg(reinterpret_cast<const float4*>(pos + static_cast<int64_t>(t + 2) * d + ch * 8));
    const float4 p1 = __ldg(reinterpret_cast<const float4*>(pos + static_cast<int64_t>(t + 2) * d + ch * 8) + 1);
    const float2 f0 = unpack_bf16x2(v.x), f1 = unpack_bf16x2(v.y), f2 = unpack_bf16x2(v.z), f3 = unpack_bf16x2(v.w);
    uint4 o;
    o.x = pack_bf16x2(fmaf(f0.x, scale, p0.x), fmaf(f0.y, scale, p0.y));
    o.y = pack_bf16x2(fmaf(f1.x, scale, p0.z), fmaf(f1.y, scale, p0.w));
    o.z = pack_bf16x2(fmaf(f2.x, scale, p1.x), fmaf(f2.y, scale, p1.y));
    o.w = pack_bf16x2(fmaf(f3.x, scale, p1.z), fmaf(f3.y, scale, p1.w));
    reinterpret_cast<uint4*>(out)[(static_cast<int64_t>(r0) + t) * d8 + ch] = o;
  }
}

__global__ void transpose_bf16_kernel(const __nv_bfloat16* __restrict__ in, int64_t ld_in, __nv_bfloat16* __restrict__ out, int64_t ld_out,
                                      int rows, int cols) {
  jl::pdl_prologue();
  __shared__ __nv_bfloat16 tile[32][34];
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int r = r0 + j, c = c0 + threadIdx.x;
    if (r < rows && c < cols) tile[j][threadIdx.x] = in[static_cast<int64_t>(r) * ld_in + c];
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int c = c0 + j, r = r0 + threadIdx.x;
    if (r < rows && c < cols) out[static_cast<int64_t>(c) * ld_out + r] = tile[threadIdx.x][j];
  }
}

// Column sums (bias gradients) in ONE launch.  A cluster of 8 CTAs owns 32 adjacent columns: four threads cover the 64
// contiguous bytes a row has in those columns (full 32-byte sectors, where the previous 8-column CTAs used half of every
// sector they fetched), 64 rows per CTA pass, the row groups dealt round-robin to the CTAs of the cluster, four passes in
// flight per thread.  Each CTA combines its 64 row-threads by a fixed-order shared-memory tree; CTA 0 then adds the eight
// partials in rank order through distributed shared memory — deterministic, no atomics, no second kernel, and 8× as many
// CTAs as column blocks (192 for d = 768, where the 8-column version had 96).
constexpr int CS_CLUSTER = 8;
constexpr int CS_COLS = 32;
__global__ void __cluster_dims__(CS_CLUSTER, 1, 1) __launch_bounds__(256)
colsum_kernel(const __nv_bfloat16* __restrict__ x, int64_t ldx, int rows, int cols, float* __restrict__ out) {
  jl::pdl_prologue();
  __shared__ float s_acc[4][64][9];
  __shared__ float s_part[CS_COLS];
  const int tid = threadIdx.x;
  const int rank = static_cast<int>(ptx::cluster_ctarank());
  const int grp = tid & 3, rix = tid >> 2;
  const int col = (blockIdx.x / CS_CLUSTER) * CS_COLS + grp * 8;
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.0f;
  const bool vec = ((cols & 7) == 0) && ((ldx & 7) == 0) && ((reinterpret_cast<uintptr_t>(x) & 15) == 0);
  if (vec) {
    if (col < cols) {
      int r = rank * 64 + rix;
      constexpr int STEP = CS_CLUSTER * 64;
      for (; r + 3 * STEP < rows; r += 4 * STEP) {
        uint4 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) v[u] = __ldg(reinterpret_cast<const uint4*>(x + static_cast<int64_t>(r + STEP * u) * ldx + col));
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const float2 f0 = unpack_bf16x2(v[u].x), f1 = unpack_bf16x2(v[u].y), f2 = unpack_bf16x2(v[u].z), f3 = unpack_bf16x2(v[u].w);
          acc[0] += f0.x; acc[1] += f0.y; acc[2] += f1.x; acc[3] += f1.y;
          acc[4] += f2.x; acc[5] += f2.y; acc[6] += f3.x; acc[7] += f3.y;
        }
      }
      for (; r < rows; r += STEP) {
        const uint4 v = __ldg(reinterpret_cast<const uint4*>(x + static_cast<int64_t>(r) * ldx + col));
        const float2 f0 = unpack_bf16x2(v.x), f1 = unpack_bf16x2(v.y), f2 = unpack_bf16x2(v.z), f3 = unpack_bf16x2(v.w);
        acc[0] += f0.x; acc[1] += f0.y; acc[2] += f1.x; acc[3] += f1.y;
        acc[4] += f2.x; acc[5] += f2.y; acc[6] += f3.x; acc[7] += f3.y;
      }
    }
  } else {
    for (int r = rank * 64 + rix; r < rows; r += CS_CLUSTER * 64)
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (col + j < cols) acc[j] += __bfloat162float(x[static_cast<int64_t>(r) * ldx + col + j]);
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) s_acc[grp][rix][j] = acc[j];
  __syncthreads();
  for (int stride = 32; stride >= 1; stride >>= 1) {
    if (rix < stride) {
#pragma unroll
      for (int j = 0; j < 8; ++j) s_acc[grp][rix][j] += s_acc[grp][rix + stride][j];
    }
    __syncthreads();
  }
  if (tid < CS_COLS) s_part[tid] = s_acc[tid >> 3][0][tid & 7];
  ptx::cluster_sync_all();
  if (rank == 0 && tid < CS_COLS) {
    const int c = (blockIdx.x / CS_CLUSTER) * CS_COLS + tid;
    if (c < cols) {
      const uint32_t local = ptx::smem_u32(&s_part[tid]);
      float tot = 0.0f;
#pragma unroll
      for (int q = 0; q < CS_CLUSTER; ++q) tot += ptx::ld_shared_cluster_f32(ptx::mapa_shared(local, static_cast<uint32_t>(q)));
      out[c] = tot;
    }
  }
  ptx::cluster_sync_all();      // the partials of every CTA stay mapped until CTA 0 has read them
}

__global__ void cast_f32_to_bf16_kernel(const float* __restrict__ in, __nv_bfloat16* __restrict__ out, int64_t n) {
  jl::pdl_prologue();
  const int64_t n4 = n >> 2;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  const int64_t tid = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (((reinterpret_cast<uintptr_t>(in) & 15) == 0) && ((reinterpret_cast<uintptr_t>(out) & 7) == 0)) {
    for (int64_t i = tid; i < n4; i += stride) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(in) + i);
      uint2 o;
      o.x = pack_bf16x2(v.x, v.y);
      o.y = pack_bf16x2(v.z, v.w);
      reinterpret_cast<uint2*>(out)[i] = o;
    }
    for (int64_t i = n4 * 4 + tid; i < n; i += stride) out[i] = __float2bfloat16_rn(in[i]);
  } else {
    for (int64_t i = tid; i < n; i += stride) out[i] = __float2bfloat16_rn(in[i]);
  }
}

__global__ void add_bf16_kernel(const __nv_bfloat16* __restrict__ a, const __nv_bfloat16* __restrict__ b, __nv_bfloat16* __restrict__ out, int64_t n) {
  jl::pdl_prologue();
  const int64_t n8 = n >> 3;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  const int64_t tid = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const bool vec = ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b) | reinterpret_cast<uintptr_t>(out)) & 15) == 0;
  if (vec) {
    for (int64_t i = tid; i < n8; i += stride) {
      const uint4 x = __ldg(reinterpret_cast<const uint4*>(a) + i);
      const uint4 y = __ldg(reinterpret_cast<const uint4*>(b) + i);
      const float2 x0 = unpack_bf16x2(x.x), x1 = unpack_bf16x2(x.y), x2 = unpack_bf16x2(x.z), x3 = unpack_bf16x2(x.w);
      const float2 y0 = unpack_bf16x2(y.x), y1 = unpack_bf16x2(y.y), y2 = unpack_bf16x2(y.z), y3 = unpack_bf16x2(y.w);
      uint4 o;
      o.x = pack_bf16x2(x0.x + y0.x, x0.y + y0.y);
      o.y = pack_bf16x2(x1.x + y1.x, x1.y + y1.y);
      o.z = pack_bf16x2(x2.x + y2.x, x2.y + y2.y);
      o.w = pack_bf16x2(x3.x + y3.x, x3.y + y3.y);
      reinterpret_cast<uint4*>(out)[i] = o;
    }
    for (int64_t i = n8 * 8 + tid; i < n; i += stride) out[i] = __float2bfloat16_rn(__bfloat162float(a[i]) + __bfloat162float(b[i]));
  } else {
    for (int64_t i = tid; i < n; i += stride) out[i] = __float2bfloat16_rn(__bfloat162float(a[i]) + __bfloat162float(b[i]));
  }
}

// torch.optim.AdamW single-tensor update order: decay, moments, bias corrections, addcdiv.
__global__ void adamw_kernel(const jl_adamw_params p, float bc1, float bc2_sqrt) {
  jl::pdl_prologue();
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  float lr = p.lr;
  if (p.hyper_dev != nullptr) {      // graph-captured launch: the step-dependent scalars come from device memory
    lr = __ldg(p.hyper_dev);
    bc1 = __ldg(p.hyper_dev + 1);
    bc2_sqrt = __ldg(p.hyper_dev + 2);
  }
  const float step_size = lr / bc1;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < p.n; i += stride) {
    const float g = p.grad[i] * p.grad_scale;
    float w = p.param[i];
    w *= (1.0f - lr * p.weight_decay);
    const float m = p.exp_avg[i] + (g - p.exp_avg[i]) * (1.0f - p.beta1);
    const float v = p.exp_avg_sq[i] * p.beta2 + (1.0f - p.beta2) * g * g;
    const float denom = sqrtf(v) / bc2_sqrt + p.eps;
    w -= step_size * (m / denom);
    p.param[i] = w;
    p.exp_avg[i] = m;
    p.exp_avg_sq[i] = v;
    if (p.param_bf16 != nullptr) reinterpret_cast<__nv_bfloat16*>(p.param_bf16)[i] = __float2bfloat16_rn(w);
  }
}

// the optimizer clock of graph-captured steps (see jl_adamw_advance)
__global__ void adamw_advance_kernel(float* __restrict__ hyper) {
  jl::pdl_prologue();
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    const float step = hyper[3] + 1.0f;
    hyper[3] = step;
    hyper[1] = 1.0f - powf(hyper[4], step);
    hyper[2] = sqrtf(1.0f - powf(hyper[5], step));
  }
}

static int grid_for(int64_t work, int threads) {
  int64_t blocks = (work + threads - 1) / threads;
  const int64_t cap = 148 * 16;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return static_cast<int>(blocks);
}

}  // namespace jl

extern "C" {

int jl_im2col_k5s2(const void* x, void* out, int32_t batch, int32_t t_in, int32_t c, int32_t t_out, void* stream) {
  JL_REQUIRE(x && out, JL_EINVAL, "im2col: null pointer");
  JL_REQUIRE(batch > 0 && t_in > 0 && c > 0 && t_out > 0, JL_EINVAL, "im2col: dims must be positive");
  JL_REQUIRE((c & 7) == 0, JL_EUNSUPPORTED_SHAPE, "im2col: channels must be a multiple of 8 (got %d)", c);
  JL_REQUIRE(t_out == (t_in - 1) / 2 + 1, JL_EINVAL, "im2col: t_out %d != (t_in - 1) / 2 + 1 for t_in %d", t_out, t_in);
  JL_REQUIRE(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(out)) & 15) == 0, JL_EINVAL, "im2col: pointers must be 16-byte aligned");
  int rc = jl::check_device();
  if (rc != JL_OK) return rc;
  const int64_t total = static_cast<int64_t>(batch) * t_out * 5 * (c / 8);
  jl::launch(jl::im2col_k5s2_kernel, jl::grid_for(total, 256), 256, 0, reinterpret_cast<cudaStream_t>(stream), 
      reinterpret_cast<const uint4*>(x), reinterpret_cast<uint4*>(out), batch, t_in, c / 8, t_out);
  JL_CHECK_LAUNCH("im2col_k5s2");
  return JL_OK;
}

int jl_embed_positions(void* h, float scale, const float* pos_table, const int32_t* lengths, int32_t batch, int32_t seq, int32_t d,
                       void* stream) {
  JL_REQUIRE(h && pos_table && lengths, JL_EINVAL, "embed_positions: null pointer");
  JL_REQUIRE(batch > 0 && seq > 0 && d > 0 && (d & 7) == 0, JL_EINVAL, "embed_positions: bad dims (d must be a multiple of 8)");
  JL_REQUIRE(((reinterpret_cast<uintptr_t>(h) | reinterpret_cast<uintptr_t>(pos_table)) & 15) == 0, JL_EINVAL,
             "embed_positions: pointers must be 16-byte aligned");
  int rc = jl::check_device();
  if (rc != JL_OK) return rc;
  const int64_t total = static_cast<int64_t>(batch) * seq * (d / 8);
  jl::launch(jl::embed_positions_kernel, jl::grid_for(total, 256), 256, 0, reinterpret_cast<cudaStream_t>(stream), 
      reinterpret_cast<__nv_bfloat16*>(h), scale, pos_table, lengths, batch, seq, d);
  JL_CHECK_LAUNCH("embed_positions");
  return JL_OK;
}

int jl_embed_positions_packed(const void* h, void* out, float scale, const float* pos_table, const int32_t* cu_seqlens, int32_t batch,
                              int32_t seq, int32_t d, void* stream) {
  JL_REQUIRE(h && out && pos_table && cu_seqlens, JL_EINVAL, "embed_positions_packed: null pointer");
  JL_REQUIRE(batch > 0 && seq > 0 && d > 0 && (d & 7) == 0, JL_EINVAL, "embed_positions_packed: bad dims (d must be a multiple of 8)");
  JL_REQUIRE(((reinterpret_cast<uintptr_t>(h) | reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(pos_table)) & 15) == 0, JL_EINVAL,
             "embed_positions_packed: pointers must be 16-byte aligned");
  int rc = jl::check_device();
  if (rc != JL_OK) return rc;
  const int64_t total = static_cast<int64_t>(batch) * seq * (d / 8);
  jl::launch(jl::embed_positions_packed_kernel, jl::grid_for(total, 256), 256, 0, reinterpret_cast<cudaStream_t>(stream),
             reinterpret_cast<const __nv_bfloat16*>(h), reinterpret_cast<__nv_bfloat16*>(out), scale, pos_table, cu_seqlens, batch, seq, d);
  JL_CHECK_LAUNCH("embed_positions_packed");
  return JL_OK;
}

int jl_transpose_bf16(const void* in, int64_t ld_in, void* out, int64_t ld_out, int32_t rows, int32_t cols, void* stream) {
  JL_REQUIRE(in && out, JL_EINVAL, "transpose: null pointer");
  JL_REQUIRE(rows > 0 && cols > 0 && ld_in >= cols && ld_out >= rows, JL_EINVAL, "transpose: bad dims");
  int rc = jl::check_device();
  if (rc != JL_OK) return rc;
  dim3 grid(jl::ceil_div(cols, 32), jl::ceil_div(rows, 32));
  jl::launch(jl::transpose_bf16_kernel, grid, dim3(32, 8), 0, reinterpret_cast<cudaStream_t>(stream), 
      reinterpret_cast<const __nv_bfloat16*>(in), ld_in, reinterpret_cast<__nv_bfloat16*>(out), ld_out, rows, cols);
  JL_CHECK_LAUNCH("transpose_bf16");
  return JL_OK;
}

int jl_colsum_workspace_bytes(int32_t rows, int32_t cols, size_t* out) {
  JL_REQUIRE(out != nullptr && rows > 0 && cols > 0, JL_EINVAL, "colsum_workspace_bytes: bad argument");
  *out = 0;   // single-pass kernel: no scratch needed (the argument is kept for ABI stability)
  return JL_OK;
}

int jl_colsum_bf16(const void* x, int64_t ldx, float* out, int32_t rows, int32_t cols, float* partial, void* stream) {
  (void)partial;
  JL_REQUIRE(x && out, JL_EINVAL, "colsum: null pointer");
  JL_REQUIRE(rows > 0 && cols > 0 && ldx >= cols, JL_EINVAL, "colsum: bad dims");
  int rc = jl::check_device();
  if (rc != JL_OK) return rc;
  jl::launch(jl::colsum_kernel, jl::ceil_div(cols, jl::CS_COLS) * jl::CS_CLUSTER, 256, 0, reinterpret_cast<cudaStream_t>(stream), reinterpret_cast<const __nv_bfloat16*>(x), ldx, rows,
                                                                                             cols, out);
  JL_CHECK_LAUNCH("colsum");
  return JL_OK;
}

int jl_cast_f32_to_bf16(const float* in, void* out, int64_t n, void* stream) {
  JL_REQUIRE(in && out && n > 0, JL_EINVAL, "cast: bad argument");
  int rc = jl::check_device();
  if (rc != JL_OK) return rc;
  jl::launch(jl::cast_f32_to_bf16_kernel, jl::grid_for(n / 4 + 1, 256), 256, 0, reinterpret_cast<cudaStream_t>(stream), 
      in, reinterpret_cast<__nv_bfloat16*>(out), n);
  JL_CHECK_LAUNCH("cast_f32_to_bf16");
  return JL_OK;
}

int jl_add_bf16(const void* a, const void* b, void* out, int64_t n, void* stream) {
  JL_REQUIRE(a && b && out && n > 0, JL_EINVAL, "add: bad argument");
  int rc = jl::check_device();
  if (rc != JL_OK) return rc;
  jl::launch(jl::add_bf16_kernel, jl::grid_for(n / 8 + 1, 256), 256, 0, reinterpret_cast<cudaStream_t>(stream), 
      reinterpret_cast<const __nv_bfloat16*>(a), reinterpret_cast<const __nv_bfloat16*>(b), reinterpret_cast<__nv_bfloat16*>(out), n);
  JL_CHECK_LAUNCH("add_bf16");
  return JL_OK;
}

int jl_adamw_advance(float* hyper_dev, void* stream) {
  JL_REQUIRE(hyper_dev != nullptr && (reinterpret_cast<uintptr_t>(hyper_dev) & 3) == 0, JL_EINVAL, "adamw_advance: null / misaligned pointer");
  int rc = jl::check_device();
  if (rc != JL_OK) return rc;
  jl::launch(jl::adamw_advance_kernel, 1, 32, 0, reinterpret_cast<cudaStream_t>(stream), hyper_dev);
  JL_CHECK_LAUNCH("adamw_advance");
  return JL_OK;
}

int jl_adamw_bucket(const jl_adamw_params* p, void* stream) {
  JL_REQUIRE(p && p->param && p->grad && p->exp_avg && p->exp_avg_sq, JL_EINVAL, "adamw: null pointer");
  JL_REQUIRE(p->n > 0 && (p->step >= 1 || p->hyper_dev != nullptr), JL_EINVAL, "adamw: n must be positive and step >= 1 (or hyper_dev given)");
  int rc = jl::check_device();
  if (rc != JL_OK) return rc;
  const float stepf = static_cast<float>(p->step >= 1 ? p->step : 1);
  const float bc1 = 1.0f - powf(p->beta1, stepf);
  const float bc2_sqrt = sqrtf(1.0f - powf(p->beta2, stepf));
  jl::launch(jl::adamw_kernel, jl::grid_for(p->n, 256), 256, 0, reinterpret_cast<cudaStream_t>(stream), *p, bc1, bc2_sqrt);
  JL_CHECK_LAUNCH("adamw_bucket");
  return JL_OK;
}

}  // extern "C"
