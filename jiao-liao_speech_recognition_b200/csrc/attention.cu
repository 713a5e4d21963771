// Self-attention softmax(Q Kᵀ·scale + keymask) V per (utterance, head), head_dim 64, forward and backward.
// Replaces the eager attention math of SP/transformers/models/wav2vec2/modeling_wav2vec2.py:438-463 (called from
// :500-549) without materialising the [B, H, T, T] scores, and serves the AttAdapter's own attention (heads = 1).
//
// Flash-style tiling: 64 query rows × 64 keys per step, 4 warps × 16 rows, bf16 mma.sync.m16n8k16 with fp32
// accumulation, operands staged in XOR-swizzled shared memory with cp.async and read with ldmatrix (conflict free).
// On this path attention is ≈ 5 % of the FLOPs at T' = 250; the dense projections around it run on
// tcgen05.  Backward is split into a dQ kernel (one CTA per query tile, also produces delta = rowsum(dO ∘ O)) and a
// dK/dV kernel (one CTA per key tile) so that every gradient element has exactly one writer — deterministic, no
// atomics.  Key-padding mask from `lengths`; query rows >= length produce zeros and receive no gradient.
#include <math_constants.h>

#include <atomic>

#include "common.cuh"

namespace jl {

constexpr int ATT_D = 64;
constexpr int ATT_B = 64;          // rows per tile (queries and keys)
constexpr int ATT_THREADS = 128;
constexpr float LOG2E = 1.4426950408889634f;

__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}
__device__ __forceinline__ void mma_bf16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, int src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory"); }

// Tile = 64 rows × 64 bf16 (128 B per row, 8 chunks of 16 B); physical chunk = chunk ^ (row & 7).
__device__ __forceinline__ uint32_t tile_addr(uint32_t base, int row, int chunk) {
  return base + static_cast<uint32_t>(row * 128 + ((chunk ^ (row & 7)) << 4));
}

// Load rows [row0, row0 + 64) × 64 columns starting at `src` (row stride ld elements); rows >= row_limit are zero-filled.
__device__ __forceinline__ void load_tile(uint32_t base, const __nv_bfloat16* src, int64_t ld, int row0, int row_limit) {
#pragma unroll
  for (int i = 0; i < (ATT_B * 8) / ATT_THREADS; ++i) {
    const int idx = threadIdx.x + i * ATT_THREADS;
    const int r = idx >> 3, c = idx & 7;
    const bool ok = (row0 + r) < row_limit;
    const __nv_bfloat16* p = src + static_cast<int64_t>(ok ? (row0 + r) : 0) * ld + c * 8;
    cp_async16(tile_addr(base, r, c), p, ok ? 16 : 0);
  }
}

// A fragments (16 rows starting at row0, all 64 k) of a row-major tile.
__device__ __forceinline__ void load_a_frags(uint32_t base, int row0, int lane, uint32_t (&f)[4][4]) {
#pragma unroll
  for (int kk = 0; kk < 4; ++kk) ldsm_x4(f[kk], tile_addr(base, row0 + (lane & 7) + ((lane >> 3) & 1) * 8, 2 * kk + (lane >> 4)));
}
// B fragments from a tile stored [n rows][k]: n-block pair np (16 n), k-block kk (16 k) → {b0,b1} of n-block 2np, {b0,b1} of 2np+1.
__device__ __forceinline__ void load_b_nk(uint32_t base, int np, int kk, int lane, uint32_t (&r)[4]) {
  ldsm_x4(r, tile_addr(base, np * 16 + (lane & 7) + (lane >> 4) * 8, 2 * kk + ((lane >> 3) & 1)));
}
// B fragments from a tile stored [k rows][n]: k-block kk, n-block pair np (transposing load).
__device__ __forceinline__ void load_b_kn(uint32_t base, int kk, int np, int lane, uint32_t (&r)[4]) {
  ldsm_x4_t(r, tile_addr(base, kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8, 2 * np + (lane >> 4)));
}

// acc[16 × 64] += A(16 × 64, regs) · Bᵀ with B tile stored [64 n][64 k]
__device__ __forceinline__ void mma_a_bnk(float (&acc)[8][4], const uint32_t (&a)[4][4], uint32_t b_base, int lane) {
#pragma unroll
  for (int kk = 0; kk < 4; ++kk)
#pragma unroll
    for (int np = 0; np < 4; ++np) {
      uint32_t r[4];
      load_b_nk(b_base, np, kk, lane, r);
      mma_bf16(acc[2 * np], a[kk], r[0], r[1]);
      mma_bf16(acc[2 * np + 1], a[kk], r[2], r[3]);
    }
}
// acc[16 × 64] += P(16 × 64, fp32 C-layout regs, converted to bf16 A fragments) · B with B tile stored [64 k][64 n]
__device__ __forceinline__ void mma_p_bkn(float (&acc)[8][4], const float (&p)[8][4], uint32_t b_base, int lane) {
#pragma unroll
  for (int kk = 0; kk < 4; ++kk) {
    uint32_t a[4];
    a[0] = pack_bf16x2(p[2 * kk][0], p[2 * kk][1]);
    a[1] = pack_bf16x2(p[2 * kk][2], p[2 * kk][3]);
    a[2] = pack_bf16x2(p[2 * kk + 1][0], p[2 * kk + 1][1]);
    a[3] = pack_bf16x2(p[2 * kk + 1][2], p[2 * kk + 1][3]);
#pragma unroll
    for (int np = 0; np < 4; ++np) {
      uint32_t r[4];
      load_b_kn(b_base, kk, np, lane, r);
      mma_bf16(acc[2 * np], a, r[0], r[1]);
      mma_bf16(acc[2 * np + 1], a, r[2], r[3]);
    }
  }
}

__device__ __forceinline__ void zero_acc(float (&a)[8][4]) {
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) a[i][j] = 0.0f;
}

// Store a warp's 16 × 64 fp32 C-layout tile as bf16 rows (row_lo = first row of the warp), masking rows >= row_limit.
__device__ __forceinline__ void store_c_tile(__nv_bfloat16* dst, int64_t ld, int row_lo, int row_limit, int lane, const float (&acc)[8][4], float s_lo,
                                             float s_hi) {
  const int g = lane >> 2, t = lane & 3;
  const int r0 = row_lo + g, r1 = r0 + 8;
#pragma unroll
  for (int nb = 0; nb < 8; ++nb) {
    const int col = nb * 8 + 2 * t;
    if (r0 < row_limit) *reinterpret_cast<uint32_t*>(dst + static_cast<int64_t>(r0) * ld + col) = pack_bf16x2(acc[nb][0] * s_lo, acc[nb][1] * s_lo);
    if (r1 < row_limit) *reinterpret_cast<uint32_t*>(dst + static_cast<int64_t>(r1) * ld + col) = pack_bf16x2(acc[nb][2] * s_hi, acc[nb][3] * s_hi);
  }
}

// ------------------------------------------------------------------------------------------------ forward
__global__ void __launch_bounds__(ATT_THREADS) attn_fwd_kernel(const jl_attn_fwd_params p) {
  jl::pdl_prologue();
  __shared__ __align__(128) __nv_bfloat16 s_q[ATT_B * ATT_D];
  __shared__ __align__(128) __nv_bfloat16 s_k[ATT_B * ATT_D];
  __shared__ __align__(128) __nv_bfloat16 s_v[ATT_B * ATT_D];
  const int b = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * ATT_B;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int len = min(p.lengths ? p.lengths[b] : p.seq, p.seq);
  const int64_t row_base = static_cast<int64_t>(b) * p.seq;
  const __nv_bfloat16* q = reinterpret_cast<const __nv_bfloat16*>(p.q) + row_base * p.ld_qkv + h * ATT_D;
  const __nv_bfloat16* k = reinterpret_cast<const __nv_bfloat16*>(p.k) + row_base * p.ld_qkv + h * ATT_D;
  const __nv_bfloat16* v = reinterpret_cast<const __nv_bfloat16*>(p.v) + row_base * p.ld_qkv + h * ATT_D;
  __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(p.o) + row_base * p.ld_o + h * ATT_D;
  float* lse = p.lse ? p.lse + (static_cast<int64_t>(b) * p.heads + h) * p.seq : nullptr;
  const int row_lo = q0 + warp * 16;
  float acc_o[8][4];
  zero_acc(acc_o);

  if (q0 >= len) {   // the whole query tile is padding
    store_c_tile(o, p.ld_o, row_lo, p.seq, lane, acc_o, 0.0f, 0.0f);
    if (lse && t == 0) {
      if (row_lo + g < p.seq) lse[row_lo + g] = 0.0f;
      if (row_lo + g + 8 < p.seq) lse[row_lo + g + 8] = 0.0f;
    }
    return;
  }
  const uint32_t q_base = static_cast<uint32_t>(__cvta_generic_to_shared(s_q));
  const uint32_t k_base = static_cast<uint32_t>(__cvta_generic_to_shared(s_k));
  const uint32_t v_base = static_cast<uint32_t>(__cvta_generic_to_shared(s_v));
  load_tile(q_base, q, p.ld_qkv, q0, len);
  cp_async_wait_all();
  __syncthreads();
  uint32_t qf[4][4];
  load_a_frags(q_base, warp * 16, lane, qf);

  const float sl2 = p.scale * LOG2E;
  float m_lo = -CUDART_INF_F, m_hi = -CUDART_INF_F, l_lo = 0.0f, l_hi = 0.0f;
  const int nkb = (len + ATT_B - 1) / ATT_B;
  for (int kb = 0; kb < nkb; ++kb) {
    __syncthreads();
    load_tile(k_base, k, p.ld_qkv, kb * ATT_B, len);
    load_tile(v_base, v, p.ld_qkv, kb * ATT_B, len);
    cp_async_wait_all();
    __syncthreads();
    float s[8][4];
    zero_acc(s);
    mma_a_bnk(s, qf, k_base, lane);
    float mx_lo = -CUDART_INF_F, mx_hi = -CUDART_INF_F;
#pragma unroll
    for (int nb = 0; nb < 8; ++nb)
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int key = kb * ATT_B + nb * 8 + 2 * t + (c & 1);
        if (key >= len) s[nb][c] = -CUDART_INF_F;
        if (c < 2) mx_lo = fmaxf(mx_lo, s[nb][c]);
        else mx_hi = fmaxf(mx_hi, s[nb][c]);
      }
    mx_lo = fmaxf(mx_lo, __shfl_xor_sync(0xffffffffu, mx_lo, 1));
    mx_lo = fmaxf(mx_lo, __shfl_xor_sync(0xffffffffu, mx_lo, 2));
    mx_hi = fmaxf(mx_hi, __shfl_xor_sync(0xffffffffu, mx_hi, 1));
    mx_hi = fmaxf(mx_hi, __shfl_xor_sync(0xffffffffu, mx_hi, 2));
    const float mn_lo = fmaxf(m_lo, mx_lo), mn_hi = fmaxf(m_hi, mx_hi);   // finite: every key tile holds >= 1 valid key
    const float corr_lo = exp2f((m_lo - mn_lo) * sl2), corr_hi = exp2f((m_hi - mn_hi) * sl2);
    float rs_lo = 0.0f, rs_hi = 0.0f;
#pragma unroll
    for (int nb = 0; nb < 8; ++nb) {
      s[nb][0] = exp2f((s[nb][0] - mn_lo) * sl2);
      s[nb][1] = exp2f((s[nb][1] - mn_lo) * sl2);
      s[nb][2] = exp2f((s[nb][2] - mn_hi) * sl2);
      s[nb][3] = exp2f((s[nb][3] - mn_hi) * sl2);
      rs_lo += s[nb][0] + s[nb][1];
      rs_hi += s[nb][2] + s[nb][3];
    }
    rs_lo += __shfl_xor_sync(0xffffffffu, rs_lo, 1);
    rs_lo += __shfl_xor_sync(0xffffffffu, rs_lo, 2);
    rs_hi += __shfl_xor_sync(0xffffffffu, rs_hi, 1);
    rs_hi += __shfl_xor_sync(0xffffffffu, rs_hi, 2);
    l_lo = l_lo * corr_lo + rs_lo;
    l_hi = l_hi * corr_hi + rs_hi;
    m_lo = mn_lo;
    m_hi = mn_hi;
#pragma unroll
    for (int nb = 0; nb < 8; ++nb) {
      acc_o[nb][0] *= corr_lo; acc_o[nb][1] *= corr_lo;
      acc_o[nb][2] *= corr_hi; acc_o[nb][3] *= corr_hi;
    }
    mma_p_bkn(acc_o, s, v_base, lane);
  }
  const int r0 = row_lo + g, r1 = r0 + 8;
  const float inv_lo = (r0 < len) ? 1.0f / l_lo : 0.0f;
  const float inv_hi = (r1 < len) ? 1.0f / l_hi : 0.0f;
  store_c_tile(o, p.ld_o, row_lo, p.seq, lane, acc_o, inv_lo, inv_hi);
  if (lse && t == 0) {
    if (r0 < p.seq) lse[r0] = (r0 < len) ? m_lo * p.scale + logf(l_lo) : 0.0f;
    if (r1 < p.seq) lse[r1] = (r1 < len) ? m_hi * p.scale + logf(l_hi) : 0.0f;
  }
}

// ------------------------------------------------------------------------------------------------ backward: dQ (+ delta)
__global__ void __launch_bounds__(ATT_THREADS) attn_bwd_dq_kernel(const jl_attn_bwd_params p) {
  jl::pdl_prologue();
  __shared__ __align__(128) __nv_bfloat16 s_q[ATT_B * ATT_D];
  __shared__ __align__(128) __nv_bfloat16 s_do[ATT_B * ATT_D];
  __shared__ __align__(128) __nv_bfloat16 s_k[ATT_B * ATT_D];
  __shared__ __align__(128) __nv_bfloat16 s_v[ATT_B * ATT_D];
  __shared__ float s_delta[ATT_B];
  const int b = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * ATT_B;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int len = min(p.lengths ? p.lengths[b] : p.seq, p.seq);
  const int64_t row_base = static_cast<int64_t>(b) * p.seq;
  const __nv_bfloat16* q = reinterpret_cast<const __nv_bfloat16*>(p.q) + row_base * p.ld_qkv + h * ATT_D;
  const __nv_bfloat16* k = reinterpret_cast<const __nv_bfloat16*>(p.k) + row_base * p.ld_qkv + h * ATT_D;
  const __nv_bfloat16* v = reinterpret_cast<const __nv_bfloat16*>(p.v) + row_base * p.ld_qkv + h * ATT_D;
  const __nv_bfloat16* o = reinterpret_cast<const __nv_bfloat16*>(p.o) + row_base * p.ld_o + h * ATT_D;
  const __nv_bfloat16* d_o = reinterpret_cast<const __nv_bfloat16*>(p.d_o) + row_base * p.ld_o + h * ATT_D;
  __nv_bfloat16* dq = reinterpret_cast<__nv_bfloat16*>(p.dq) + row_base * p.ld_dqkv + h * ATT_D;
  const float* lse = p.lse + (static_cast<int64_t>(b) * p.heads + h) * p.seq;
  float* delta = p.delta + (static_cast<int64_t>(b) * p.heads + h) * p.seq;
  const int row_lo = q0 + warp * 16;
  float acc[8][4];
  zero_acc(acc);

  // delta[q] = Σ_d dO[q, d] · O[q, d]   (two threads per row)
  {
    const int r = threadIdx.x >> 1, half = threadIdx.x & 1;
    const int row = q0 + r;
    float d = 0.0f;
    if (row < len) {
      const uint4* po = reinterpret_cast<const uint4*>(o + static_cast<int64_t>(row) * p.ld_o + half * 32);
      const uint4* pd = reinterpret_cast<const uint4*>(d_o + static_cast<int64_t>(row) * p.ld_o + half * 32);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const uint4 a = __ldg(po + i), c = __ldg(pd + i);
        const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, cw[4] = {c.x, c.y, c.z, c.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 fa = unpack_bf16x2(aw[j]), fc = unpack_bf16x2(cw[j]);
          d = fmaf(fa.x, fc.x, d);
          d = fmaf(fa.y, fc.y, d);
        }
      }
    }
    d += __shfl_xor_sync(0xffffffffu, d, 1);
    if (half == 0) {
      s_delta[r] = d;
      if (row < p.seq) delta[row] = d;
    }
  }
  if (q0 >= len) {
    store_c_tile(dq, p.ld_dqkv, row_lo, p.seq, lane, acc, 0.0f, 0.0f);
    return;
  }
  const uint32_t q_base = static_cast<uint32_t>(__cvta_generic_to_shared(s_q));
  const uint32_t do_base = static_cast<uint32_t>(__cvta_generic_to_shared(s_do));
  const uint32_t k_base = static_cast<uint32_t>(__cvta_generic_to_shared(s_k));
  const uint32_t v_base = static_cast<uint32_t>(__cvta_generic_to_shared(s_v));
  load_tile(q_base, q, p.ld_qkv, q0, len);
  load_tile(do_base, d_o, p.ld_o, q0, len);
  cp_async_wait_all();
  __syncthreads();
  uint32_t qf[4][4], dof[4][4];
  load_a_frags(q_base, warp * 16, lane, qf);
  load_a_frags(do_base, warp * 16, lane, dof);
  const int r0 = row_lo + g, r1 = r0 + 8;
  const float sl2 = p.scale * LOG2E;
  const float lse_lo = (r0 < len) ? lse[r0] * LOG2E : 0.0f, lse_hi = (r1 < len) ? lse[r1] * LOG2E : 0.0f;
  const float dl_lo = s_delta[warp * 16 + g], dl_hi = s_delta[warp * 16 + g + 8];

  const int nkb = (len + ATT_B - 1) / ATT_B;
  for (int kb = 0; kb < nkb; ++kb) {
    __syncthreads();
    load_tile(k_base, k, p.ld_qkv, kb * ATT_B, len);
    load_tile(v_base, v, p.ld_qkv, kb * ATT_B, len);
    cp_async_wait_all();
    __syncthreads();
    float s[8][4], dp[8][4];
    zero_acc(s);
    zero_acc(dp);
    mma_a_bnk(s, qf, k_base, lane);
    mma_a_bnk(dp, dof, v_base, lane);
#pragma unroll
    for (int nb = 0; nb < 8; ++nb)
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int key = kb * ATT_B + nb * 8 + 2 * t + (c & 1);
        const bool lo = c < 2;
        const bool valid = key < len && (lo ? r0 : r1) < len;
        const float pr = valid ? exp2f(s[nb][c] * sl2 - (lo ? lse_lo : lse_hi)) : 0.0f;
        s[nb][c] = pr * (dp[nb][c] - (lo ? dl_lo : dl_hi)) * p.scale;     // dS
      }
    mma_p_bkn(acc, s, k_base, lane);                                       // dQ += dS · K
  }
  store_c_tile(dq, p.ld_dqkv, row_lo, p.seq, lane, acc, 1.0f, 1.0f);
}

// ------------------------------------------------------------------------------------------------ backward: dK, dV
__global__ void __launch_bounds__(ATT_THREADS) attn_bwd_dkv_kernel(const jl_attn_bwd_params p) {
  jl::pdl_prologue();
  __shared__ __align__(128) __nv_bfloat16 s_k[ATT_B * ATT_D];
  __shared__ __align__(128) __nv_bfloat16 s_v[ATT_B * ATT_D];
  __shared__ __align__(128) __nv_bfloat16 s_q[ATT_B * ATT_D];
  __shared__ __align__(128) __nv_bfloat16 s_do[ATT_B * ATT_D];
  __shared__ float s_lse[ATT_B];
  __shared__ float s_delta[ATT_B];
  const int b = blockIdx.z, h = blockIdx.y, k0 = blockIdx.x * ATT_B;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int len = min(p.lengths ? p.lengths[b] : p.seq, p.seq);
  const int64_t row_base = static_cast<int64_t>(b) * p.seq;
  const __nv_bfloat16* q = reinterpret_cast<const __nv_bfloat16*>(p.q) + row_base * p.ld_qkv + h * ATT_D;
  const __nv_bfloat16* k = reinterpret_cast<const __nv_bfloat16*>(p.k) + row_base * p.ld_qkv + h * ATT_D;
  const __nv_bfloat16* v = reinterpret_cast<const __nv_bfloat16*>(p.v) + row_base * p.ld_qkv + h * ATT_D;
  const __nv_bfloat16* d_o = reinterpret_cast<const __nv_bfloat16*>(p.d_o) + row_base * p.ld_o + h * ATT_D;
  __nv_bfloat16* dk = reinterpret_cast<__nv_bfloat16*>(p.dk) + row_base * p.ld_dqkv + h * ATT_D;
  __nv_bfloat16* dv = reinterpret_cast<__nv_bfloat16*>(p.dv) + row_base * p.ld_dqkv + h * ATT_D;
  const float* lse = p.lse + (static_cast<int64_t>(b) * p.heads + h) * p.seq;
  const float* delta = p.delta + (static_cast<int64_t>(b) * p.heads + h) * p.seq;
  const int row_lo = k0 + warp * 16;
  float acc_dk[8][4], acc_dv[8][4];
  zero_acc(acc_dk);
  zero_acc(acc_dv);
  if (k0 >= len) {
    store_c_tile(dk, p.ld_dqkv, row_lo, p.seq, lane, acc_dk, 0.0f, 0.0f);
    store_c_tile(dv, p.ld_dqkv, row_lo, p.seq, lane, acc_dv, 0.0f, 0.0f);
    return;
  }
  const uint32_t k_base = static_cast<uint32_t>(__cvta_generic_to_shared(s_k));
  const uint32_t v_base = static_cast<uint32_t>(__cvta_generic_to_shared(s_v));
  const uint32_t q_base = static_cast<uint32_t>(__cvta_generic_to_shared(s_q));
  const uint32_t do_base = static_cast<uint32_t>(__cvta_generic_to_shared(s_do));
  load_tile(k_base, k, p.ld_qkv, k0, len);
  load_tile(v_base, v, p.ld_qkv, k0, len);
  cp_async_wait_all();
  __syncthreads();
  uint32_t kf[4][4], vf[4][4];
  load_a_frags(k_base, warp * 16, lane, kf);
  load_a_frags(v_base, warp * 16, lane, vf);
  const int key_lo = row_lo + g, key_hi = key_lo + 8;
  const float sl2 = p.scale * LOG2E;

  const int nqb = (len + ATT_B - 1) / ATT_B;
  for (int qb = 0; qb < nqb; ++qb) {
    __syncthreads();
    load_tile(q_base, q, p.ld_qkv, qb * ATT_B, len);
    load_tile(do_base, d_o, p.ld_o, qb * ATT_B, len);
    if (threadIdx.x < ATT_B) {
      const int r = qb * ATT_B + threadIdx.x;
      s_lse[threadIdx.x] = (r < len) ? lse[r] * LOG2E : 0.0f;
      s_delta[threadIdx.x] = (r < len) ? delta[r] : 0.0f;
    }
    cp_async_wait_all();
    __syncthreads();
    // Sᵀ[key, query] = K_w · Q_iᵀ ;  dPᵀ[key, query] = V_w · dO_iᵀ
    float st[8][4], dpt[8][4];
    zero_acc(st);
    zero_acc(dpt);
    mma_a_bnk(st, kf, q_base, lane);
    mma_a_bnk(dpt, vf, do_base, lane);
    float ds[8][4];
#pragma unroll
    for (int nb = 0; nb < 8; ++nb)
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int ql = nb * 8 + 2 * t + (c & 1);
        const int qrow = qb * ATT_B + ql;
        const int key = (c < 2) ? key_lo : key_hi;
        const bool valid = key < len && qrow < len;
        const float pr = valid ? exp2f(st[nb][c] * sl2 - s_lse[ql]) : 0.0f;
        st[nb][c] = pr;                                                   // Pᵀ
        ds[nb][c] = pr * (dpt[nb][c] - s_delta[ql]) * p.scale;            // dSᵀ
      }
    mma_p_bkn(acc_dv, st, do_base, lane);                                 // dV += Pᵀ · dO
    mma_p_bkn(acc_dk, ds, q_base, lane);                                  // dK += dSᵀ · Q
  }
  store_c_tile(dk, p.ld_dqkv, row_lo, p.seq, lane, acc_dk, 1.0f, 1.0f);
  store_c_tile(dv, p.ld_dqkv, row_lo, p.seq, lane, acc_dv, 1.0f, 1.0f);
}

static int attn_check(const void* q, const void* k, const void* v, int64_t ld_qkv, int batch, int seq, int heads) {
  JL_REQUIRE(q && k && v, JL_EINVAL, "attn: null q/k/v");
  JL_REQUIRE(batch > 0 && seq > 0 && heads > 0, JL_EINVAL, "attn: batch, seq, heads must be positive");
  JL_REQUIRE(batch <= 65535 && heads <= 65535, JL_EUNSUPPORTED_SHAPE, "attn: batch / heads exceed the grid limits");
  JL_REQUIRE((ld_qkv & 7) == 0, JL_EINVAL, "attn: ld_qkv must be a multiple of 8");
  JL_REQUIRE(((reinterpret_cast<uintptr_t>(q) | reinterpret_cast<uintptr_t>(k) | reinterpret_cast<uintptr_t>(v)) & 15) == 0, JL_EINVAL,
             "attn: q/k/v must be 16-byte aligned");
  return JL_OK;
}

int attn_fwd_tc(const jl_attn_fwd_params* p, cudaStream_t stream);   // attention_tc.cu
int attn_bwd_tc(const jl_attn_bwd_params* p, cudaStream_t stream);
extern int g_attn_fwd_ctas;
extern int g_attn_short;
static std::atomic<int> g_attn_impl{0};   // 0 = tcgen05 kernels, 1 = mma.sync kernels

}  // namespace jl

extern "C" {

void jl_debug_set_attn_impl(int impl) {
  // 0 = tcgen05 kernels (short-sequence forward for <= 256 frames, else the key-block forward compiled for 2 CTAs/SM),
  // 1 = mma.sync kernels, 2 = tcgen05 with the 3-CTAs/SM key-block forward, 3 = tcgen05 with the key-block forward at every length
  jl::g_attn_fwd_ctas = (impl == 2) ? 3 : 2;
  jl::g_attn_short = (impl == 0) ? 1 : 0;
  jl::g_attn_impl.store(impl == 1 ? 1 : 0);
}

int jl_attn_fwd(const jl_attn_fwd_params* p, void* stream) {
  JL_REQUIRE(p != nullptr, JL_EINVAL, "attn_fwd: null params");
  int rc = jl::attn_check(p->q, p->k, p->v, p->ld_qkv, p->batch, p->seq, p->heads);
  if (rc != JL_OK) return rc;
  JL_REQUIRE(p->o && (p->ld_o & 7) == 0 && (reinterpret_cast<uintptr_t>(p->o) & 15) == 0, JL_EINVAL, "attn_fwd: bad output pointer / stride");
  rc = jl::check_device();
  if (rc != JL_OK) return rc;
  if (p->cu_seqlens != nullptr) JL_REQUIRE(p->total_rows > 0, JL_EINVAL, "attn_fwd: packed layout needs total_rows > 0");
  if (jl::g_attn_impl.load() == 0) return jl::attn_fwd_tc(p, reinterpret_cast<cudaStream_t>(stream));
  JL_REQUIRE(p->cu_seqlens == nullptr, JL_EUNSUPPORTED_SHAPE, "attn_fwd: the packed (cu_seqlens) layout is implemented by the tcgen05 kernels only");
  dim3 grid(jl::ceil_div(p->seq, jl::ATT_B), p->heads, p->batch);
  jl::launch(jl::attn_fwd_kernel, grid, jl::ATT_THREADS, 0, reinterpret_cast<cudaStream_t>(stream), *p);
  JL_CHECK_LAUNCH("attn_fwd");
  return JL_OK;
}

int jl_attn_bwd(const jl_attn_bwd_params* p, void* stream) {
  JL_REQUIRE(p != nullptr, JL_EINVAL, "attn_bwd: null params");
  int rc = jl::attn_check(p->q, p->k, p->v, p->ld_qkv, p->batch, p->seq, p->heads);
  if (rc != JL_OK) return rc;
  JL_REQUIRE(p->o && p->d_o && p->lse && p->dq && p->dk && p->dv && p->delta, JL_EINVAL, "attn_bwd: null pointer");
  JL_REQUIRE((p->ld_o & 7) == 0 && (p->ld_dqkv & 7) == 0, JL_EINVAL, "attn_bwd: strides must be multiples of 8");
  JL_REQUIRE(((reinterpret_cast<uintptr_t>(p->o) | reinterpret_cast<uintptr_t>(p->d_o) | reinterpret_cast<uintptr_t>(p->dq) |
               reinterpret_cast<uintptr_t>(p->dk) | reinterpret_cast<uintptr_t>(p->dv)) & 15) == 0, JL_EINVAL,
             "attn_bwd: pointers must be 16-byte aligned");
  rc = jl::check_device();
  if (rc != JL_OK) return rc;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  if (p->cu_seqlens != nullptr) JL_REQUIRE(p->total_rows > 0, JL_EINVAL, "attn_bwd: packed layout needs total_rows > 0");
  if (jl::g_attn_impl.load() == 0) return jl::attn_bwd_tc(p, s);
  JL_REQUIRE(p->cu_seqlens == nullptr, JL_EUNSUPPORTED_SHAPE, "attn_bwd: the packed (cu_seqlens) layout is implemented by the tcgen05 kernels only");
  dim3 grid(jl::ceil_div(p->seq, jl::ATT_B), p->heads, p->batch);
  jl::launch(jl::attn_bwd_dq_kernel, grid, jl::ATT_THREADS, 0, s, *p);
  JL_CHECK_LAUNCH("attn_bwd_dq");
  jl::launch(jl::attn_bwd_dkv_kernel, grid, jl::ATT_THREADS, 0, s, *p);
  JL_CHECK_LAUNCH("attn_bwd_dkv");
  return JL_OK;
}

}  // extern "C"
