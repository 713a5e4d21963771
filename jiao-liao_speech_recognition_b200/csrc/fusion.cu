// f4: AdapterFusion-style AttAdapter over the K source-dialect adapters of a slot (SURVEY §8c ambiguity (ii), §8f f4;
// /root/reference/README.md:1 "multi-dialect knowledge transfer", "adapter with attention").  Per frame:
//     α = softmax_k(q · key_k · scale),   out = h + Σ_k α_k y_k
// with y_k the update of the k-th dialect's WFAdapter, q a projection of LN(h) and key_k a projection of y_k (both produced by
// the tcgen05 GEMM).  These two kernels are the only part that is not a GEMM: the K dot products of a frame, the softmax over
// K (K <= 8: registers) and the weighted sum — pure streaming work, (K + 1) · rows · d · 2 B read and rows · d · 2 B written
// forward; the backward pass re-reads y once and writes the K gradients α_k · dout.
// One warp per frame row, 16-byte loads, every reduction in a fixed order (deterministic).
#include <math_constants.h>

#include "common.cuh"

namespace jl {

constexpr int FUSE_MAX_K = 8;

__device__ __forceinline__ void fuse_unpack8(const uint4& v, float (&f)[8]) {
  const float2 a = unpack_bf16x2(v.x), b = unpack_bf16x2(v.y), c = unpack_bf16x2(v.z), e = unpack_bf16x2(v.w);
  f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y; f[4] = c.x; f[5] = c.y; f[6] = e.x; f[7] = e.y;
}
__device__ __forceinline__ uint4 fuse_pack8(const float (&f)[8]) {
  return make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]), pack_bf16x2(f[6], f[7]));
}

// q · key_k over b columns (b a multiple of 8, <= 256): lanes 0 .. b/8-1 hold 8 columns each
__device__ __forceinline__ float fuse_dot_b(const __nv_bfloat16* q, const __nv_bfloat16* key, int b, int lane) {
  float acc = 0.0f;
  if (lane * 8 < b) {
    float fq[8], fk[8];
    fuse_unpack8(__ldg(reinterpret_cast<const uint4*>(q) + lane), fq);
    fuse_unpack8(__ldg(reinterpret_cast<const uint4*>(key) + lane), fk);
#pragma unroll
    for (int j = 0; j < 8; ++j) acc = fmaf(fq[j], fk[j], acc);
  }
  return warp_sum(acc);
}

__global__ void __launch_bounds__(256) fusion_combine_fwd_kernel(const jl_fusion_params p) {
  jl::pdl_prologue();
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= p.rows) return;
  const int K = p.num_adapters;
  const __nv_bfloat16* q = reinterpret_cast<const __nv_bfloat16*>(p.q) + static_cast<int64_t>(row) * p.ldq;
  float s[FUSE_MAX_K];
  float mx = -CUDART_INF_F;
#pragma unroll
  for (int k = 0; k < FUSE_MAX_K; ++k) {
    s[k] = -CUDART_INF_F;
    if (k < K) {
      const __nv_bfloat16* key = reinterpret_cast<const __nv_bfloat16*>(p.key) + k * p.key_stride + static_cast<int64_t>(row) * p.ldkey;
      s[k] = fuse_dot_b(q, key, p.b, lane) * p.scale;
      mx = fmaxf(mx, s[k]);
    }
  }
  float sum = 0.0f;
#pragma unroll
  for (int k = 0; k < FUSE_MAX_K; ++k) {
    s[k] = (k < K) ? expf(s[k] - mx) : 0.0f;
    sum += s[k];
  }
  const float inv = 1.0f / sum;
#pragma unroll
  for (int k = 0; k < FUSE_MAX_K; ++k) s[k] *= inv;
  if (p.alpha != nullptr && lane < K) {
    float a = 0.0f;
#pragma unroll
    for (int k = 0; k < FUSE_MAX_K; ++k) a = (lane == k) ? s[k] : a;
    p.alpha[static_cast<int64_t>(row) * K + lane] = a;
  }
  const uint4* hrow = reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(p.h) + static_cast<int64_t>(row) * p.ldh);
  uint4* orow = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.out) + static_cast<int64_t>(row) * p.ldo);
  const int nv = p.d >> 3;
  if (p.row_lengths != nullptr) {                  // padded layout: rows past the utterance's last frame stay exactly zero
    const int bi = row / p.rows_per_seq;
    if (row - bi * p.rows_per_seq >= __ldg(p.row_lengths + bi)) {
      for (int c = lane; c < nv; c += 32) orow[c] = make_uint4(0u, 0u, 0u, 0u);
      return;
    }
  }
  for (int c = lane; c < nv; c += 32) {
    float acc[8];
    fuse_unpack8(__ldg(hrow + c), acc);
#pragma unroll
    for (int k = 0; k < FUSE_MAX_K; ++k) {
      if (k < K) {
        const uint4* yrow = reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(p.y) + k * p.y_stride + static_cast<int64_t>(row) * p.ldy);
        float fy[8];
        fuse_unpack8(__ldg(yrow + c), fy);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = fmaf(s[k], fy[j], acc[j]);
      }
    }
    orow[c] = fuse_pack8(acc);
  }
}

// dα_k = dout · y_k;  ds_k = α_k (dα_k − Σ_j α_j dα_j) · scale;  dq = Σ_k ds_k key_k;  dkey_k = ds_k q;  dy_k = α_k dout
__global__ void __launch_bounds__(256) fusion_combine_bwd_kernel(const jl_fusion_params p) {
  jl::pdl_prologue();
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= p.rows) return;
  const int K = p.num_adapters;
  const int nv = p.d >> 3;
  const uint4* drow = reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(p.dout) + static_cast<int64_t>(row) * p.lddout);
  float al[FUSE_MAX_K], da[FUSE_MAX_K];
#pragma unroll
  for (int k = 0; k < FUSE_MAX_K; ++k) {
    al[k] = (k < K) ? __ldg(p.alpha + static_cast<int64_t>(row) * K + k) : 0.0f;
    da[k] = 0.0f;
  }
  // pass over the row: dα_k partials and dy_k = α_k · dout
  for (int c = lane; c < nv; c += 32) {
    float fd[8];
    fuse_unpack8(__ldg(drow + c), fd);
#pragma unroll
    for (int k = 0; k < FUSE_MAX_K; ++k) {
      if (k < K) {
        const int64_t off = k * p.y_stride + static_cast<int64_t>(row) * p.ldy;
        float fy[8], o[8];
        fuse_unpack8(__ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(p.y) + off) + c), fy);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          da[k] = fmaf(fd[j], fy[j], da[k]);
          o[j] = al[k] * fd[j];
        }
        reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.dy) + k * p.dy_stride + static_cast<int64_t>(row) * p.lddy)[c] = fuse_pack8(o);
      }
    }
  }
  float dot = 0.0f;
#pragma unroll
  for (int k = 0; k < FUSE_MAX_K; ++k) {
    da[k] = warp_sum(da[k]);
    dot = fmaf(al[k], da[k], dot);
  }
  float ds[FUSE_MAX_K];
#pragma unroll
  for (int k = 0; k < FUSE_MAX_K; ++k) ds[k] = al[k] * (da[k] - dot) * p.scale;
  if (lane * 8 < p.b) {
    const __nv_bfloat16* q = reinterpret_cast<const __nv_bfloat16*>(p.q) + static_cast<int64_t>(row) * p.ldq;
    float fq[8], dq[8];
    fuse_unpack8(__ldg(reinterpret_cast<const uint4*>(q) + lane), fq);
#pragma unroll
    for (int j = 0; j < 8; ++j) dq[j] = 0.0f;
#pragma unroll
    for (int k = 0; k < FUSE_MAX_K; ++k) {
      if (k < K) {
        const __nv_bfloat16* key = reinterpret_cast<const __nv_bfloat16*>(p.key) + k * p.key_stride + static_cast<int64_t>(row) * p.ldkey;
        float fk[8], o[8];
        fuse_unpack8(__ldg(reinterpret_cast<const uint4*>(key) + lane), fk);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          dq[j] = fmaf(ds[k], fk[j], dq[j]);
          o[j] = ds[k] * fq[j];
        }
        reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.dkey) + k * p.dkey_stride + static_cast<int64_t>(row) * p.lddkey)[lane] = fuse_pack8(o);
      }
    }
    reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.dq) + static_cast<int64_t>(row) * p.lddq)[lane] = fuse_pack8(dq);
  }
}

static int fusion_validate(const jl_fusion_params* p, bool bwd) {
  JL_REQUIRE(p != nullptr, JL_EINVAL, "fusion: null params");
  JL_REQUIRE(p->y && p->q && p->key && p->alpha, JL_EINVAL, "fusion: null y / q / key / alpha");
  JL_REQUIRE(p->rows > 0 && p->d > 0 && (p->d & 7) == 0 && p->b > 0 && (p->b & 7) == 0 && p->b <= 256, JL_EUNSUPPORTED_SHAPE,
             "fusion: d and b must be multiples of 8, b <= 256 (got d %d, b %d)", p->d, p->b);
  JL_REQUIRE(p->num_adapters >= 1 && p->num_adapters <= FUSE_MAX_K, JL_EUNSUPPORTED_SHAPE, "fusion: 1 <= num_adapters <= %d (got %d)", FUSE_MAX_K,
             p->num_adapters);
  JL_REQUIRE(((p->ldy | p->ldq | p->ldkey | p->y_stride | p->key_stride) & 7) == 0, JL_EINVAL, "fusion: strides must be multiples of 8 elements");
  JL_REQUIRE(((reinterpret_cast<uintptr_t>(p->y) | reinterpret_cast<uintptr_t>(p->q) | reinterpret_cast<uintptr_t>(p->key)) & 15) == 0, JL_EINVAL,
             "fusion: pointers must be 16-byte aligned");
  if (!bwd) {
    JL_REQUIRE(p->row_lengths == nullptr || p->rows_per_seq > 0, JL_EINVAL, "fusion_fwd: row_lengths needs rows_per_seq > 0");
    JL_REQUIRE(p->h && p->out && ((p->ldh | p->ldo) & 7) == 0, JL_EINVAL, "fusion_fwd: null h / out or bad stride");
    JL_REQUIRE(((reinterpret_cast<uintptr_t>(p->h) | reinterpret_cast<uintptr_t>(p->out)) & 15) == 0, JL_EINVAL, "fusion_fwd: pointers must be 16-byte aligned");
  } else {
    JL_REQUIRE(p->dout && p->dy && p->dq && p->dkey, JL_EINVAL, "fusion_bwd: null dout / dy / dq / dkey");
    JL_REQUIRE(((p->lddout | p->lddy | p->lddq | p->lddkey | p->dy_stride | p->dkey_stride) & 7) == 0, JL_EINVAL, "fusion_bwd: strides must be multiples of 8");
    JL_REQUIRE(((reinterpret_cast<uintptr_t>(p->dout) | reinterpret_cast<uintptr_t>(p->dy) | reinterpret_cast<uintptr_t>(p->dq) |
                 reinterpret_cast<uintptr_t>(p->dkey)) & 15) == 0, JL_EINVAL, "fusion_bwd: pointers must be 16-byte aligned");
  }
  return check_device();
}

}  // namespace jl

extern "C" {

int jl_fusion_combine_fwd(const jl_fusion_params* p, void* stream) {
  int rc = jl::fusion_validate(p, false);
  if (rc != JL_OK) return rc;
  jl::launch(jl::fusion_combine_fwd_kernel, jl::ceil_div(p->rows, 8), 256, 0, reinterpret_cast<cudaStream_t>(stream), *p);
  JL_CHECK_LAUNCH("fusion_combine_fwd");
  return JL_OK;
}

int jl_fusion_combine_bwd(const jl_fusion_params* p, void* stream) {
  int rc = jl::fusion_validate(p, true);
  if (rc != JL_OK) return rc;
  jl::launch(jl::fusion_combine_bwd_kernel, jl::ceil_div(p->rows, 8), 256, 0, reinterpret_cast<cudaStream_t>(stream), *p);
  JL_CHECK_LAUNCH("fusion_combine_bwd");
  return JL_OK;
}

}  // extern "C"
