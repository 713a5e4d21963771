// Attention on the 5th-generation tensor cores: softmax(Q Kᵀ·scale + keymask) V per (utterance, head), head_dim 64,
// forward and backward, scores never leave the SM.
// Replaces the eager attention math of SP/transformers/models/wav2vec2/modeling_wav2vec2.py:438-463 (+ its autograd
// backward) and serves the AttAdapter's one-head attention.
//
// Common structure (256 threads, up to 2 CTAs per SM so one CTA's softmax overlaps the other's MMAs):
//   warp 0      TMA producer — 128B-swizzled [rows × 64] bf16 boxes straight out of the [B·T, ld] q|k|v / dO matrices
//   warp 1      one thread issues tcgen05.mma (M = 128 rows = the CTA's "outer" tile, N = 64, K = 64), accumulators in TMEM
//   warp 2      TMEM allocator (256 columns)
//   warps 4-11  two warpgroups × 128 threads: thread (quadrant, lane) of warpgroup g owns columns [32g, 32g+32) of TMEM
//               lane = outer row: tcgen05.ld its half of the score row, ex2.approx / mask / scale in registers, write
//               the bf16 operand tile (P, dS, Pᵀ, dSᵀ) back to shared memory in the K-major 128B-swizzle layout the next
//               MMA reads, tcgen05.st for the online-softmax rescale of O; the two halves of a row exchange their
//               running max through shared memory (one named barrier per key tile)
// The same 64-row TMA tile is used as a K-major B operand (scores) and as an MN-major B operand (value / gradient
// products) — no transposed copies.
//   forward   outer = 128 queries, inner = 64 keys:  S = Q·Kᵀ → P → O += P·V            (online softmax, O rescaled in TMEM)
//   bwd dQ    outer = 128 queries, inner = 64 keys:  S = Q·Kᵀ, dP = dO·Vᵀ → dS → dQ += dS·K   (also writes delta = rowsum(dO∘O))
//   bwd dKV   outer = 128 keys,    inner = 64 queries: Sᵀ = K·Qᵀ, dPᵀ = V·dOᵀ → Pᵀ, dSᵀ → dV += Pᵀ·dO, dK += dSᵀ·Q
// Every gradient element has exactly one writer (deterministic, no atomics).
#include <cuda.h>
#include <math_constants.h>

#include "common.cuh"
#include "ptx_sm100.cuh"

namespace jl {

constexpr int TC_THREADS = 384;   // 4 service warps (TMA, MMA, TMEM alloc, spare) + 8 softmax warps
constexpr int TC_OUTER = 128;
constexpr int TC_INNER = 64;
constexpr uint32_t TC_T128 = 128 * 128;   // bytes of a [128 × 64] bf16 tile
constexpr uint32_t TC_T64 = 64 * 128;     // bytes of a [ 64 × 64] bf16 tile
constexpr float TC_LOG2E = 1.4426950408889634f;
constexpr uint32_t TC_IDESC_KK = (1u << 4) | (1u << 7) | (1u << 10) | ((64u >> 3) << 17) | ((128u >> 4) << 24);   // M128 N64, A,B K-major
constexpr uint32_t TC_IDESC_KMN = TC_IDESC_KK | (1u << 16);                                                        // B MN-major

__device__ __forceinline__ float tc_exp2(float x) {      // single MUFU.EX2 (flush-to-zero); arguments here are <= 8
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// D[128 × 64] (+)= A[128 × 64] · B, A K-major at a_addr; B = 64-row tile at b_addr read K-major (Bᵀ: rows are N) or
// MN-major (rows are K).  4 MMAs of K = 16.
__device__ __forceinline__ void tc_mma_64(uint32_t d_tmem, uint32_t a_addr, uint32_t b_addr, bool b_mn, bool accumulate) {
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const uint64_t da = ptx::make_sw128_desc(a_addr + k * 32, 16, 1024);
    const uint64_t db = b_mn ? ptx::make_sw128_desc(b_addr + k * 2048, 8192, 1024) : ptx::make_sw128_desc(b_addr + k * 32, 16, 1024);
    ptx::umma_bf16(d_tmem, da, db, b_mn ? TC_IDESC_KMN : TC_IDESC_KK, (accumulate || k > 0) ? 1u : 0u);
  }
}

// Row r of a [128 × 64] bf16 K-major 128B-swizzled tile: 8 chunks of 16 B, physical chunk = c ^ (r & 7).
__device__ __forceinline__ void tc_store_row(uint8_t* tile, int r, const uint32_t (&packed)[32]) {
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    uint4 v = make_uint4(packed[4 * c], packed[4 * c + 1], packed[4 * c + 2], packed[4 * c + 3]);
    *reinterpret_cast<uint4*>(tile + r * 128 + ((c ^ (r & 7)) << 4)) = v;
  }
}

// 16 consecutive columns (two 16-byte chunks, first chunk index c0) of row r
__device__ __forceinline__ void tc_store_cols16(uint8_t* tile, int r, int c0, const uint32_t (&pk)[8]) {
  *reinterpret_cast<uint4*>(tile + r * 128 + (((c0) ^ (r & 7)) << 4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
  *reinterpret_cast<uint4*>(tile + r * 128 + (((c0 + 1) ^ (r & 7)) << 4)) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
}

__device__ __forceinline__ void tc_store_global_row32(__nv_bfloat16* dst, const float (&v)[32], float s) {
  if ((reinterpret_cast<uintptr_t>(dst) & 31) == 0) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      uint32_t w[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) w[j] = pack_bf16x2(v[16 * i + 2 * j] * s, v[16 * i + 2 * j + 1] * s);
      st_global_v8(dst + 16 * i, w);
    }
    return;
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    uint4 o;
    o.x = pack_bf16x2(v[8 * i + 0] * s, v[8 * i + 1] * s);
    o.y = pack_bf16x2(v[8 * i + 2] * s, v[8 * i + 3] * s);
    o.z = pack_bf16x2(v[8 * i + 4] * s, v[8 * i + 5] * s);
    o.w = pack_bf16x2(v[8 * i + 6] * s, v[8 * i + 7] * s);
    reinterpret_cast<uint4*>(dst)[i] = o;
  }
}

__device__ __forceinline__ void tc_zero_rows(__nv_bfloat16* base, int64_t ld, int row0, int seq) {
  for (int idx = threadIdx.x; idx < TC_OUTER * 8; idx += TC_THREADS) {
    const int r = idx >> 3, c = idx & 7;
    if (row0 + r < seq) *reinterpret_cast<uint4*>(base + static_cast<int64_t>(row0 + r) * ld + c * 8) = make_uint4(0u, 0u, 0u, 0u);
  }
}

// Where an utterance lives.  Padded layout: rows [b·seq, (b+1)·seq) of the [B·seq, ld] matrices, `lengths[b]` of them valid, the
// others kept zero (they are written).  Packed layout (cu_seqlens != NULL, SURVEY §5 "varlen packing"): rows [cu[b], cu[b+1]) of
// the [total_rows, ld] matrices, no padding rows at all — nothing past the utterance's last row may be written (it is the next
// utterance's first row); per-row statistics (lse, delta) are laid out [heads, total_rows].
struct TcSeq {
  int64_t row_base;     // first row of the utterance in the q|k|v / o matrices
  int64_t stat_base;    // first element of the utterance's (lse, delta) run for this head
  int len;              // valid frames
  int lim;              // rows this utterance owns in the matrices (padded: seq, packed: len)
};
template <typename P>
__device__ __forceinline__ TcSeq tc_seq(const P& p, int b, int h) {
  TcSeq q;
  if (p.cu_seqlens != nullptr) {
    const int r0 = p.cu_seqlens[b];
    q.row_base = r0;
    q.len = min(p.cu_seqlens[b + 1] - r0, p.seq);
    q.lim = q.len;
    q.stat_base = static_cast<int64_t>(h) * p.total_rows + r0;
  } else {
    q.row_base = static_cast<int64_t>(b) * p.seq;
    q.len = min(p.lengths ? p.lengths[b] : p.seq, p.seq);
    q.lim = p.seq;
    q.stat_base = (static_cast<int64_t>(b) * p.heads + h) * p.seq;
  }
  return q;
}

struct TcBars {
  uint64_t x_full;         // outer tile(s) landed
  uint64_t y_full[2];      // inner tile stage landed
  uint64_t y_empty[2];     // inner tile stage consumed by the accumulate MMAs
  uint64_t s_full[2];      // score MMAs complete (fwd: two score buffers; bwd uses [0])
  uint64_t s_empty[2];     // score rows read into registers by the 4 softmax warps
  uint64_t p_full;         // operand tile(s) written to shared memory by the 4 softmax warps
  uint64_t acc_done;       // accumulate MMAs of the current inner tile complete
  uint32_t tmem_slot;
};

// ================================================================================================ forward
struct __align__(1024) AttnFwdSmem {
  uint8_t q[TC_T128];
  uint8_t k[2][TC_T64];
  uint8_t v[2][TC_T64];
  uint8_t p[TC_T128];
  float red_max[2][2][TC_OUTER];   // [key tile parity][column half][row]
  float red_sum[2][TC_OUTER];
  TcBars bars;
};

template <int MIN_CTAS>
__global__ void __launch_bounds__(TC_THREADS, MIN_CTAS)
attn_fwd_tc_kernel(const __grid_constant__ CUtensorMap tq, const __grid_constant__ CUtensorMap tk, const __grid_constant__ CUtensorMap tv,
                   const jl_attn_fwd_params p) {
  // Everything up to griddepcontrol.wait touches no global memory, so under programmatic dependent launch it overlaps the
  // tail of the preceding kernel: barrier init, tensor-map prefetch, TMEM allocation.
  const int b = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * TC_OUTER;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  extern __shared__ uint8_t tc_smem_raw[];
  AttnFwdSmem& s = *reinterpret_cast<AttnFwdSmem*>(tc_smem_raw + ((1024u - (ptx::smem_u32(tc_smem_raw) & 1023u)) & 1023u));
  TcBars& B = s.bars;
  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&tq);
    ptx::prefetch_tensormap(&tk);
    ptx::prefetch_tensormap(&tv);
  }
  if (warp == 1 && lane == 0) {
    ptx::mbar_init(&B.x_full, 1);
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(&B.y_full[i], 1);
      ptx::mbar_init(&B.y_empty[i], 2);     // K slot freed by the score MMA commit, V slot by the PV commit
      ptx::mbar_init(&B.s_full[i], 1);
      ptx::mbar_init(&B.s_empty[i], 8);
    }
    ptx::mbar_init(&B.p_full, 8);
    ptx::mbar_init(&B.acc_done, 1);
    ptx::fence_barrier_init();
  }
  // 128 TMEM columns (64 scores + 64 output) per CTA: the score rows are pulled into registers with a single wait and the
  // buffer is released at once, so S_{j+1} still overlaps softmax_j.
  if (warp == 2) {
    ptx::tmem_alloc(&B.tmem_slot, 128);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  jl::pdl_prologue();   // `lengths`, q, k, v may be produced by the preceding kernel: wait before the first global access
  const TcSeq sq = tc_seq(p, b, h);
  const int len = sq.len, lim = sq.lim;
  const int64_t row_base = sq.row_base;
  __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(p.o) + row_base * p.ld_o + h * 64;
  float* lse = p.lse ? p.lse + sq.stat_base : nullptr;
  const int nkb = (len + TC_INNER - 1) / TC_INNER;
  const bool skip = q0 >= len;          // the whole query tile is padding
  if (skip) {
    tc_zero_rows(o, p.ld_o, q0, lim);
    if (lse && threadIdx.x < TC_OUTER && q0 + threadIdx.x < lim) lse[q0 + threadIdx.x] = 0.0f;
  }
  const uint32_t tmem = B.tmem_slot;
  const uint32_t t_s[2] = {tmem, tmem};
  const uint32_t t_o = tmem + 64;
  const int grow = static_cast<int>(row_base);        // row of the utterance's first frame in the [B·T, ld] matrices

  if (skip) {
    // nothing to compute
  } else if (warp == 0) {
    if (lane == 0) {
      ptx::mbar_expect_tx(&B.x_full, TC_T128);
      ptx::tma_load_2d(s.q, &tq, &B.x_full, h * 64, grow + q0);
      for (int j = 0; j < nkb; ++j) {
        const int st = j & 1;
        ptx::mbar_wait(&B.y_empty[st], ((j >> 1) & 1) ^ 1u);
        ptx::mbar_expect_tx(&B.y_full[st], 2 * TC_T64);
        ptx::tma_load_2d(s.k[st], &tk, &B.y_full[st], h * 64, grow + j * TC_INNER);
        ptx::tma_load_2d(s.v[st], &tv, &B.y_full[st], h * 64, grow + j * TC_INNER);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t q_addr = ptx::smem_u32(s.q), p_addr = ptx::smem_u32(s.p);
      ptx::mbar_wait(&B.x_full, 0);
      auto issue_scores = [&](int j) {
        const int st = j & 1;
        ptx::mbar_wait(&B.y_full[st], (j >> 1) & 1);
        ptx::mbar_wait(&B.s_empty[0], (j & 1) ^ 1u);                               // scores of block j-1 are in registers
        ptx::tc_fence_after();
        tc_mma_64(t_s[st], q_addr, ptx::smem_u32(s.k[st]), false, false);          // S_j = Q · K_jᵀ
        ptx::umma_commit(&B.y_empty[st]);
        ptx::umma_commit(&B.s_full[0]);
      };
      issue_scores(0);
      for (int j = 0; j < nkb; ++j) {
        if (j + 1 < nkb) issue_scores(j + 1);
        ptx::mbar_wait(&B.p_full, j & 1);
        ptx::tc_fence_after();
        tc_mma_64(t_o, p_addr, ptx::smem_u32(s.v[j & 1]), true, j > 0);              // O += P_j · V_j
        ptx::umma_commit(&B.y_empty[j & 1]);
        ptx::umma_commit(&B.acc_done);
      }
    }
  } else if (warp >= 4) {
    const int r = (warp & 3) * 32 + lane;                  // TMEM lane = query row of the tile
    const int half = (warp - 4) >> 2;                      // which 32 of the 64 key columns this thread owns
    const uint32_t lane_off = static_cast<uint32_t>((warp & 3) * 32) << 16;
    const float sl2 = p.scale * TC_LOG2E;
    float m = -CUDART_INF_F, l = 0.0f;                     // m: reference max of the row (same in both halves), l: partial sum
    for (int j = 0; j < nkb; ++j) {
      const int st = j & 1;
      ptx::mbar_wait(&B.s_full[0], j & 1);
      ptx::tc_fence_after();
      const int kbase = j * TC_INNER + half * 32;
      uint32_t sv[32];
      ptx::tmem_ld_32x32(t_s[st] + lane_off + half * 32, sv);
      ptx::tmem_ld_wait();
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&B.s_empty[0]);       // score buffer may be overwritten by block j + 1
      const bool partial = j * TC_INNER + TC_INNER > len;   // only the last key tile needs the padding mask
      float mloc = -CUDART_INF_F;
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        if (partial && kbase + i >= len) sv[i] = 0xff800000u;     // -inf
        mloc = fmaxf(mloc, __uint_as_float(sv[i]));
      }
      s.red_max[st][half][r] = mloc;
      asm volatile("bar.sync 2, 256;" ::: "memory");
      const float mx = fmaxf(m, fmaxf(mloc, s.red_max[st][half ^ 1][r]));
      // Lazy rescale: keep the old reference max unless the row max grew by more than 2^8 (exp2 arguments stay <= 8);
      // O and l are only rescaled then, which removes most TMEM round trips of the online softmax.
      const bool grow = (mx - m) * sl2 > 8.0f;              // true on the first tile (m = -inf); identical in both halves
      const float m_new = grow ? mx : m;
      const float corr = grow ? tc_exp2((m - m_new) * sl2) : 1.0f;
      const float mxs = m_new * sl2;
      float sum0 = 0.0f, sum1 = 0.0f;
      uint32_t packed[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const float a = tc_exp2(fmaf(__uint_as_float(sv[2 * i]), sl2, -mxs));
        const float e = tc_exp2(fmaf(__uint_as_float(sv[2 * i + 1]), sl2, -mxs));
        sum0 += a;
        sum1 += e;
        packed[i] = pack_bf16x2(a, e);
      }
      l = l * corr + (sum0 + sum1);
      m = m_new;
      if (j > 0) {
        ptx::mbar_wait(&B.acc_done, (j - 1) & 1);          // PV_{j-1} complete: O may be rescaled, P may be overwritten
        ptx::tc_fence_after();
        if (__any_sync(0xffffffffu, grow)) {
          uint32_t ov[32];
          ptx::tmem_ld_32x32(t_o + lane_off + half * 32, ov);
          ptx::tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) ov[i] = __float_as_uint(__uint_as_float(ov[i]) * corr);
          ptx::tmem_st_32x32(t_o + lane_off + half * 32, ov);
          ptx::tmem_st_wait();
          ptx::tc_fence_before();
        }
      }
#pragma unroll
      for (int c2 = 0; c2 < 2; ++c2) {
        uint32_t t0[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) t0[i] = packed[8 * c2 + i];
        tc_store_cols16(s.p, r, 4 * half + 2 * c2, t0);
      }
      ptx::fence_proxy_async();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&B.p_full);
    }
    ptx::mbar_wait(&B.acc_done, (nkb - 1) & 1);
    ptx::tc_fence_after();
    s.red_sum[half][r] = l;
    asm volatile("bar.sync 2, 256;" ::: "memory");
    const float l_tot = l + s.red_sum[half ^ 1][r];
    const int row = q0 + r;
    const float inv = (row < len) ? 1.0f / l_tot : 0.0f;
    {
      uint32_t ov[32];
      ptx::tmem_ld_32x32(t_o + lane_off + half * 32, ov);
      ptx::tmem_ld_wait();
      float of[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) of[i] = __uint_as_float(ov[i]);
      if (row < lim) tc_store_global_row32(o + static_cast<int64_t>(row) * p.ld_o + half * 32, of, inv);
    }
    if (lse && half == 0 && row < lim) lse[row] = (row < len) ? m * p.scale + logf(l_tot) : 0.0f;
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem, 128);
  }
}

// ================================================================================================ forward, seq <= 256
// Utterances of at most 256 frames (10.2 s at 40 ms): one CTA per (utterance, head, 128-query tile) computes the whole score
// tile S = Q · Kᵀ (M 128 × N 256) with one group of MMAs.  The softmax sees whole rows, so it is the plain two-pass form (max,
// then exp / sum while writing P) — no online rescale, no key-block loop and none of its hand-offs; O = P · V runs as up to 16
// MMAs (K = 256 keys) into TMEM columns S no longer needs, and P overwrites the Q / K staging area once the score MMAs have
// completed.  The dependency chain of a CTA is load → S → softmax → PV → store, once, instead of once per key block; 256 TMEM
// columns and 98 KB of shared memory keep two CTAs per SM so that one CTA's chain overlaps the other's.
constexpr uint32_t TS_IDESC_S = (1u << 4) | (1u << 7) | (1u << 10) | ((256u >> 3) << 17) | ((128u >> 4) << 24);   // M128 N256, A,B K-major

struct __align__(1024) AttnFwdShortSmem {
  uint8_t q[TC_T128];              // ┐
  uint8_t k[4][TC_T64];            // │ 64 KB: Q, K (one [256 × 64] K-major B operand) — later P (4 key tiles of [128 × 64])
  uint8_t p_tail[TC_T128];         // ┘
  uint8_t v[4][TC_T64];            // V, 64-key tiles, read as MN-major B operands
  float red_max[2][TC_OUTER];      // [column half][row]
  float red_sum[2][TC_OUTER];
  uint64_t qk_full, v_full, s_full, p_full, o_full;
  uint32_t tmem_slot;
};

__global__ void __launch_bounds__(TC_THREADS, 2)
attn_fwd_short_kernel(const __grid_constant__ CUtensorMap tq, const __grid_constant__ CUtensorMap tk, const __grid_constant__ CUtensorMap tv,
                      const jl_attn_fwd_params p) {
  const int g = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  extern __shared__ uint8_t tc_smem_raw[];
  AttnFwdShortSmem& s = *reinterpret_cast<AttnFwdShortSmem*>(tc_smem_raw + ((1024u - (ptx::smem_u32(tc_smem_raw) & 1023u)) & 1023u));
  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&tq);
    ptx::prefetch_tensormap(&tk);
    ptx::prefetch_tensormap(&tv);
  }
  if (warp == 1 && lane == 0) {
    ptx::mbar_init(&s.qk_full, 1);
    ptx::mbar_init(&s.v_full, 1);
    ptx::mbar_init(&s.s_full, 1);
    ptx::mbar_init(&s.p_full, 8);
    ptx::mbar_init(&s.o_full, 1);
    ptx::fence_barrier_init();
  }
  if (warp == 2) {
    ptx::tmem_alloc(&s.tmem_slot, 256);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  jl::pdl_prologue();     // everything above overlaps the preceding kernel; `lengths` and q / k / v may come from it
  const uint32_t tmem = s.tmem_slot;
  const TcSeq sq = tc_seq(p, b, h);
  const int len = sq.len, lim = sq.lim;
  const int64_t row_base = sq.row_base;
  const int grow = static_cast<int>(row_base);
  __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(p.o) + row_base * p.ld_o + h * 64;
  float* lse = p.lse ? p.lse + sq.stat_base : nullptr;
  const bool active = g * TC_OUTER < len;                     // the query tile has at least one valid row
  const int nkt = (len + TC_INNER - 1) / TC_INNER;            // 64-key tiles with at least one valid key: 0..4

  if (warp == 0) {
    if (lane == 0 && active) {
      ptx::mbar_expect_tx(&s.qk_full, TC_T128 + nkt * TC_T64);
      ptx::tma_load_2d(s.q, &tq, &s.qk_full, h * 64, grow + g * TC_OUTER);
      for (int j = 0; j < nkt; ++j) ptx::tma_load_2d(s.k[j], &tk, &s.qk_full, h * 64, grow + j * TC_INNER);
      ptx::mbar_expect_tx(&s.v_full, nkt * TC_T64);
      for (int j = 0; j < nkt; ++j) ptx::tma_load_2d(s.v[j], &tv, &s.v_full, h * 64, grow + j * TC_INNER);
    }
  } else if (warp == 1) {
    if (lane == 0 && active) {
      ptx::mbar_wait(&s.qk_full, 0);
      ptx::tc_fence_after();
      const uint32_t q_addr = ptx::smem_u32(s.q), k_addr = ptx::smem_u32(s.k[0]);
#pragma unroll
      for (int kk = 0; kk < 4; ++kk)
        ptx::umma_bf16(tmem, ptx::make_sw128_desc(q_addr + kk * 32, 16, 1024), ptx::make_sw128_desc(k_addr + kk * 32, 16, 1024), TS_IDESC_S,
                       kk > 0 ? 1u : 0u);
      ptx::umma_commit(&s.s_full);
      ptx::mbar_wait(&s.v_full, 0);
      ptx::mbar_wait(&s.p_full, 0);
      ptx::tc_fence_after();
      const uint32_t p_base = ptx::smem_u32(s.q);
      for (int kt = 0; kt < nkt; ++kt) tc_mma_64(tmem, p_base + kt * TC_T128, ptx::smem_u32(s.v[kt]), true, kt > 0);      // O += P[:, kt] · V_kt
      ptx::umma_commit(&s.o_full);
    }
  } else if (warp >= 4) {
    const int half = (warp - 4) >> 2;                         // which 128 of the 256 key columns
    const int r = (warp & 3) * 32 + lane;                     // TMEM lane = query row of the tile
    const int row = g * TC_OUTER + r;                         // row of the utterance
    const uint32_t lane_off = static_cast<uint32_t>((warp & 3) * 32) << 16;
    if (!active) {
      // the whole query tile is padding (or the utterance is empty): zero rows, nothing to compute
      if (row < lim) {
        uint4* dst = reinterpret_cast<uint4*>(o + static_cast<int64_t>(row) * p.ld_o + half * 32);
#pragma unroll
        for (int i = 0; i < 4; ++i) dst[i] = make_uint4(0u, 0u, 0u, 0u);
        if (lse && half == 0) lse[row] = 0.0f;
      }
    } else {
      const uint32_t t_s = tmem + lane_off + 128u * half;
      const float sl2 = p.scale * TC_LOG2E;
      const int kbase = half * 128;
      ptx::mbar_wait(&s.s_full, 0);
      ptx::tc_fence_after();
      // ---- pass 1: row maximum over this thread's 128 columns (two TMEM loads in flight)
      float mloc = -CUDART_INF_F;
#pragma unroll 1
      for (int c = 0; c < 4; c += 2) {
        uint32_t sa[32], sb[32];
        ptx::tmem_ld_32x32(t_s + 32u * c, sa);
        ptx::tmem_ld_32x32(t_s + 32u * (c + 1), sb);
        ptx::tmem_ld_wait();
        const int k0 = kbase + 32 * c;
        if (k0 + 64 <= len) {
#pragma unroll
          for (int i = 0; i < 32; ++i) mloc = fmaxf(mloc, fmaxf(__uint_as_float(sa[i]), __uint_as_float(sb[i])));
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            if (k0 + i < len) mloc = fmaxf(mloc, __uint_as_float(sa[i]));
            if (k0 + 32 + i < len) mloc = fmaxf(mloc, __uint_as_float(sb[i]));
          }
        }
      }
      s.red_max[half][r] = mloc;
      asm volatile("bar.sync 2, 256;" ::: "memory");
      const float mx = fmaxf(mloc, s.red_max[half ^ 1][r]);         // finite: len >= 1 valid key
      const float mxs = mx * sl2;
      // ---- pass 2: exp, row sum, P as the K-major 128B-swizzled A operand of the PV MMAs (over the Q / K staging area: the
      //      score MMAs, whose completion s_full signalled, were its last readers)
      float sum0 = 0.0f, sum1 = 0.0f;
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        uint32_t sv[32];
        ptx::tmem_ld_32x32(t_s + 32u * c, sv);
        ptx::tmem_ld_wait();
        const int k0 = kbase + 32 * c;
        uint32_t packed[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          float a = tc_exp2(fmaf(__uint_as_float(sv[2 * i]), sl2, -mxs));
          float e = tc_exp2(fmaf(__uint_as_float(sv[2 * i + 1]), sl2, -mxs));
          if (k0 + 2 * i >= len) a = 0.0f;
          if (k0 + 2 * i + 1 >= len) e = 0.0f;
          sum0 += a;
          sum1 += e;
          packed[i] = pack_bf16x2(a, e);
        }
        uint8_t* tile = s.q + ((k0 >> 6) * TC_T128);                // key tile of 64 columns
        const int chunk0 = ((k0 & 63) >> 3);                        // first 16-byte chunk inside the tile row: 0 or 4
#pragma unroll
        for (int c2 = 0; c2 < 2; ++c2) {
          uint32_t t0[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) t0[i] = packed[8 * c2 + i];
          tc_store_cols16(tile, r, chunk0 + 2 * c2, t0);
        }
      }
      ptx::tc_fence_before();
      ptx::fence_proxy_async();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&s.p_full);
      s.red_sum[half][r] = sum0 + sum1;
      // ---- epilogue: O / l  (O occupies the first 64 of the score columns)
      ptx::mbar_wait(&s.o_full, 0);
      ptx::tc_fence_after();
      asm volatile("bar.sync 2, 256;" ::: "memory");
      const float l_tot = (sum0 + sum1) + s.red_sum[half ^ 1][r];
      const float inv = (row < len) ? 1.0f / l_tot : 0.0f;
      uint32_t ov[32];
      ptx::tmem_ld_32x32(tmem + lane_off + 32u * half, ov);
      ptx::tmem_ld_wait();
      float of[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) of[i] = __uint_as_float(ov[i]);
      if (row < lim) tc_store_global_row32(o + static_cast<int64_t>(row) * p.ld_o + half * 32, of, inv);
      if (lse && half == 0 && row < lim) lse[row] = (row < len) ? mx * p.scale + logf(l_tot) : 0.0f;
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem, 256);
  }
}

// ================================================================================================ backward
// MODE 0: dQ   (outer rows = queries; X1 = Q, X2 = dO; inner: Y1 = K_j, Y2 = V_j;  acc0 = dQ += dS · K_j)
// MODE 1: dKV  (outer rows = keys;    X1 = K, X2 = V;  inner: Y1 = Q_i, Y2 = dO_i; acc0 = dV += Pᵀ · dO_i, acc1 = dK += dSᵀ · Q_i)
constexpr int TC_BWD_PREFETCH_SEQ = 1024;
template <int MODE>
struct __align__(1024) AttnBwdSmem {
  uint8_t x1[TC_T128];
  uint8_t x2[TC_T128];
  uint8_t y1[2][TC_T64];
  uint8_t y2[2][TC_T64];
  uint8_t op0[TC_T128];                       // dS (MODE 0) / Pᵀ (MODE 1)
  uint8_t op1[MODE == 1 ? TC_T128 : 16];      // dSᵀ (MODE 1)
  float col_lse[2][TC_INNER];                 // MODE 1: per-query log-sum-exp (×log2e) and delta of the inner tile
  float col_delta[2][TC_INNER];
  float col_all[MODE == 1 ? 2 * TC_BWD_PREFETCH_SEQ : 4];   // MODE 1, seq <= TC_BWD_PREFETCH_SEQ: lse·log2e and delta of every query, loaded once
  TcBars bars;
};

template <int MODE>
__global__ void __launch_bounds__(TC_THREADS, 2)
attn_bwd_tc_kernel(const __grid_constant__ CUtensorMap tx1, const __grid_constant__ CUtensorMap tx2, const __grid_constant__ CUtensorMap ty1,
                   const __grid_constant__ CUtensorMap ty2, const jl_attn_bwd_params p) {
  // As in the forward kernel, the set-up below needs no global memory and overlaps the preceding kernel's tail (PDL).
  const int b = blockIdx.z, h = blockIdx.y, r0 = blockIdx.x * TC_OUTER;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  extern __shared__ uint8_t tc_smem_raw[];
  AttnBwdSmem<MODE>& s = *reinterpret_cast<AttnBwdSmem<MODE>*>(tc_smem_raw + ((1024u - (ptx::smem_u32(tc_smem_raw) & 1023u)) & 1023u));
  TcBars& B = s.bars;
  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&tx1);
    ptx::prefetch_tensormap(&tx2);
    ptx::prefetch_tensormap(&ty1);
    ptx::prefetch_tensormap(&ty2);
  }
  if (warp == 1 && lane == 0) {
    ptx::mbar_init(&B.x_full, 1);
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(&B.y_full[i], 1);
      ptx::mbar_init(&B.y_empty[i], 1);
      ptx::mbar_init(&B.s_full[i], 1);
      ptx::mbar_init(&B.s_empty[i], 8);
    }
    ptx::mbar_init(&B.p_full, 8);
    ptx::mbar_init(&B.acc_done, 1);
    ptx::fence_barrier_init();
  }
  if (warp == 2) {
    ptx::tmem_alloc(&B.tmem_slot, 256);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  jl::pdl_prologue();   // `lengths` and every operand may be produced by the preceding kernel: wait before the first global access
  const TcSeq sq = tc_seq(p, b, h);
  const int len = sq.len, lim = sq.lim;
  const int64_t row_base = sq.row_base;
  const float* lse = p.lse + sq.stat_base;
  float* delta = p.delta + sq.stat_base;
  __nv_bfloat16* out0 = reinterpret_cast<__nv_bfloat16*>(MODE == 0 ? p.dq : p.dv) + row_base * p.ld_dqkv + h * 64;
  __nv_bfloat16* out1 = reinterpret_cast<__nv_bfloat16*>(p.dk) + row_base * p.ld_dqkv + h * 64;
  const int nib = (len + TC_INNER - 1) / TC_INNER;
  const bool skip = r0 >= len;          // the whole outer tile is padding
  if (skip) {
    tc_zero_rows(out0, p.ld_dqkv, r0, lim);
    if (MODE == 1) tc_zero_rows(out1, p.ld_dqkv, r0, lim);
    if (MODE == 0 && threadIdx.x < TC_OUTER && r0 + threadIdx.x < lim) delta[r0 + threadIdx.x] = 0.0f;
  }
  const uint32_t tmem = B.tmem_slot;
  const uint32_t t_s = tmem, t_dp = tmem + 64, t_acc0 = tmem + 128, t_acc1 = tmem + 192;
  const int grow = static_cast<int>(row_base);

  if (skip) {
    // nothing to compute
  } else if (warp == 0) {
    if (lane == 0) {
      ptx::mbar_expect_tx(&B.x_full, 2 * TC_T128);
      ptx::tma_load_2d(s.x1, &tx1, &B.x_full, h * 64, grow + r0);
      ptx::tma_load_2d(s.x2, &tx2, &B.x_full, h * 64, grow + r0);
      for (int j = 0; j < nib; ++j) {
        const int st = j & 1;
        ptx::mbar_wait(&B.y_empty[st], ((j >> 1) & 1) ^ 1u);
        ptx::mbar_expect_tx(&B.y_full[st], 2 * TC_T64);
        ptx::tma_load_2d(s.y1[st], &ty1, &B.y_full[st], h * 64, grow + j * TC_INNER);
        ptx::tma_load_2d(s.y2[st], &ty2, &B.y_full[st], h * 64, grow + j * TC_INNER);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t x1 = ptx::smem_u32(s.x1), x2 = ptx::smem_u32(s.x2);
      const uint32_t op0 = ptx::smem_u32(s.op0), op1 = ptx::smem_u32(s.op1);
      ptx::mbar_wait(&B.x_full, 0);
      auto issue_scores = [&](int j) {
        const int st = j & 1;
        ptx::mbar_wait(&B.y_full[st], (j >> 1) & 1);
        ptx::mbar_wait(&B.s_empty[0], (j & 1) ^ 1u);                     // score rows of block j-1 are in registers
        ptx::tc_fence_after();
        tc_mma_64(t_s, x1, ptx::smem_u32(s.y1[st]), false, false);         // S  = X1 · Y1ᵀ
        tc_mma_64(t_dp, x2, ptx::smem_u32(s.y2[st]), false, false);        // dP = X2 · Y2ᵀ
        ptx::umma_commit(&B.s_full[0]);
      };
      issue_scores(0);
      for (int j = 0; j < nib; ++j) {
        if (j + 1 < nib) issue_scores(j + 1);
        const int st = j & 1;
        ptx::mbar_wait(&B.p_full, j & 1);
        ptx::tc_fence_after();
        if (MODE == 0) {
          tc_mma_64(t_acc0, op0, ptx::smem_u32(s.y1[st]), true, j > 0);    // dQ += dS · K_j
        } else {
          tc_mma_64(t_acc0, op0, ptx::smem_u32(s.y2[st]), true, j > 0);    // dV += Pᵀ · dO_i
          tc_mma_64(t_acc1, op1, ptx::smem_u32(s.y1[st]), true, j > 0);    // dK += dSᵀ · Q_i
        }
        ptx::umma_commit(&B.y_empty[st]);
        ptx::umma_commit(&B.acc_done);
      }
    }
  } else if (warp >= 4) {
    const int tid = threadIdx.x - 128;
    const int r = (warp & 3) * 32 + lane;
    const int hh = (warp - 4) >> 2;                          // which 32 of the 64 inner columns this thread owns
    const uint32_t lane_off = static_cast<uint32_t>((warp & 3) * 32) << 16;
    const int row = r0 + r;                                  // query (MODE 0) / key (MODE 1) index
    const bool row_ok = row < len;
    const float sl2 = p.scale * TC_LOG2E;
    float row_lse = 0.0f, row_delta = 0.0f;
    if (MODE == 0) {
      // delta[q] = Σ_d dO[q, d] · O[q, d]
      if (row_ok) {
        const uint4* po = reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(p.o) + (row_base + row) * p.ld_o + h * 64);
        const uint4* pd = reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(p.d_o) + (row_base + row) * p.ld_o + h * 64);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const uint4 a = __ldg(po + i), c = __ldg(pd + i);
          const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, cw[4] = {c.x, c.y, c.z, c.w};
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float2 fa = unpack_bf16x2(aw[q]), fc = unpack_bf16x2(cw[q]);
            row_delta = fmaf(fa.x, fc.x, row_delta);
            row_delta = fmaf(fa.y, fc.y, row_delta);
          }
        }
        row_lse = lse[row] * TC_LOG2E;
      }
      if (hh == 0 && row < lim) delta[row] = row_delta;
    }
    // MODE 1 needs lse and delta of the inner tile's queries (the score columns).  For utterances of up to 1024 frames they
    // are all fetched here, once; otherwise per inner tile — a global-load latency plus a barrier in front of every tile.
    const bool pre = (MODE == 1) && p.seq <= TC_BWD_PREFETCH_SEQ;
    if (pre) {
      for (int q = tid; q < nib * TC_INNER; q += 256) {
        s.col_all[q] = (q < len) ? lse[q] * TC_LOG2E : 0.0f;
        s.col_all[TC_BWD_PREFETCH_SEQ + q] = (q < len) ? delta[q] : 0.0f;
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");
    }
    for (int j = 0; j < nib; ++j) {
      const int st = j & 1;
      const int cbase = j * TC_INNER;                        // first key (MODE 0) / query (MODE 1) of the inner tile
      const float* cl = pre ? s.col_all + cbase : s.col_lse[st];
      const float* cd = pre ? s.col_all + TC_BWD_PREFETCH_SEQ + cbase : s.col_delta[st];
      if (MODE == 1 && !pre) {
        if (tid < TC_INNER) {
          const int q = cbase + tid;
          s.col_lse[st][tid] = (q < len) ? lse[q] * TC_LOG2E : 0.0f;
        } else if (tid < 2 * TC_INNER) {
          const int q = cbase + tid - TC_INNER;
          s.col_delta[st][tid - TC_INNER] = (q < len) ? delta[q] : 0.0f;
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");
      }
      ptx::mbar_wait(&B.s_full[0], j & 1);
      ptx::tc_fence_after();
      {
        uint32_t sv[32], dv[32], pk0[16], pk1[16];
        ptx::tmem_ld_32x32(t_s + lane_off + hh * 32, sv);
        ptx::tmem_ld_32x32(t_dp + lane_off + hh * 32, dv);
        ptx::tmem_ld_wait();
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&B.s_empty[0]);     // this warp's share of S / dP is in registers
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          float pr[2], ds[2];
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int c = hh * 32 + 2 * i + e;
            const bool ok = row_ok && (cbase + c < len);
            const float lse_c = (MODE == 0) ? row_lse : cl[c];
            const float dl_c = (MODE == 0) ? row_delta : cd[c];
            const float pe = ok ? tc_exp2(fmaf(__uint_as_float(sv[2 * i + e]), sl2, -lse_c)) : 0.0f;
            pr[e] = pe;
            ds[e] = pe * (__uint_as_float(dv[2 * i + e]) - dl_c) * p.scale;
          }
          pk0[i] = (MODE == 0) ? pack_bf16x2(ds[0], ds[1]) : pack_bf16x2(pr[0], pr[1]);
          pk1[i] = pack_bf16x2(ds[0], ds[1]);
        }
        if (j > 0) ptx::mbar_wait(&B.acc_done, (j - 1) & 1);   // accumulate MMAs of block j-1 have read the operand tiles
#pragma unroll
        for (int c2 = 0; c2 < 2; ++c2) {
          uint32_t t0[8], t1[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) { t0[i] = pk0[8 * c2 + i]; t1[i] = pk1[8 * c2 + i]; }
          tc_store_cols16(s.op0, r, 4 * hh + 2 * c2, t0);
          if (MODE == 1) tc_store_cols16(s.op1, r, 4 * hh + 2 * c2, t1);
        }
      }
      ptx::fence_proxy_async();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&B.p_full);
    }
    ptx::mbar_wait(&B.acc_done, (nib - 1) & 1);
    ptx::tc_fence_after();
#pragma unroll
    for (int a = 0; a < (MODE == 1 ? 2 : 1); ++a) {
      __nv_bfloat16* dst = (a == 0) ? out0 : out1;
      uint32_t ov[32];
      ptx::tmem_ld_32x32((a == 0 ? t_acc0 : t_acc1) + lane_off + hh * 32, ov);
      ptx::tmem_ld_wait();
      float of[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) of[i] = __uint_as_float(ov[i]);
      if (row < lim) tc_store_global_row32(dst + static_cast<int64_t>(row) * p.ld_dqkv + hh * 32, of, 1.0f);
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem, 256);
  }
}

// ================================================================================================ backward, utterances of <= 256 frames
// One CTA per (utterance, head) computes dQ, dK and dV from ONE evaluation of the scores: the 256 × 256 score matrix is
// walked in four 128 × 128 blocks (key half j outer, query half i inner),
//     S = Q_i·K_jᵀ, dP = dO_i·V_jᵀ  →  P = exp2(S·scale·log2e − lse), dS = P ∘ (dP − delta) · scale   (thread = query row)
//     dQ_i += dS·K_j,   dV_j += Pᵀ·dO_i,   dK_j += dSᵀ·Q_i
// where the two-kernel path (dQ, then dKV on transposed scores) evaluates S, dP and the exponentials twice and loads every
// operand twice.  P and dS are written once, as K-major [128 q × 64 k] tiles; the tensor core reads the same bytes K-major
// (A = dS for dQ) and MN-major (A = Pᵀ, dSᵀ for dV, dK) — the major bit of the instruction descriptor and the descriptor
// strides are all that changes, as for the MN-major B operands of the other kernels.
// TMEM (512 columns): S 0-127, dP 128-255, dQ_0 256-319, dQ_1 320-383, dK_j 384-447, dV_j 448-511.
// Shared memory: Q, K, V, dO [256 × 64] (128 KB, one TMA box of 128 rows each half) + P, dS (64 KB): one CTA per SM.
constexpr uint32_t TC_IDESC_S128 = (1u << 4) | (1u << 7) | (1u << 10) | ((128u >> 3) << 17) | ((128u >> 4) << 24);   // M128 N128, A,B K-major
constexpr uint32_t TC_IDESC_DQ = TC_IDESC_KMN;                                    // M128 N64, A K-major, B MN-major
constexpr uint32_t TC_IDESC_DKV = TC_IDESC_KK | (1u << 15) | (1u << 16);          // M128 N64, A MN-major, B MN-major

constexpr int FB_SOFTMAX_WARPS = 16;                       // 4 per TMEM lane quadrant: thread = (query row, quarter of the 128 key columns)
constexpr int FB_THREADS = 128 + FB_SOFTMAX_WARPS * 32;     // + TMA, MMA, TMEM-allocator and one spare warp

struct FusedBwdBars {
  uint64_t x_full;         // Q, K, V, dO landed
  uint64_t s_full;         // score MMAs (S, dP) of the current block complete
  uint64_t s_empty;        // S / dP of the current block are in registers (16 warps)
  uint64_t p_full;         // P, dS of the current block are in shared memory (16 warps)
  uint64_t acc_done;       // accumulate MMAs of the current block complete: P / dS may be overwritten, dK_j / dV_j read after i = last
  uint64_t dkv_empty;      // dK_j / dV_j accumulators read out (16 warps)
  uint32_t tmem_slot;
};
struct __align__(1024) AttnBwdFusedSmem {
  uint8_t q[2][TC_T128];
  uint8_t k[2][TC_T128];
  uint8_t v[2][TC_T128];
  uint8_t d_o[2][TC_T128];
  uint8_t p[2][TC_T128];       // [key-column tile of 64][128 q rows × 128 B]
  uint8_t ds[2][TC_T128];
  float row_lse[2 * TC_OUTER];   // lse · log2e per query
  float row_delta[2 * TC_OUTER]; // Σ_d dO · O per query
  FusedBwdBars bars;
};

// 16 fp32 → 16 bf16 (32 bytes) to global memory
__device__ __forceinline__ void tc_store_global_row16(__nv_bfloat16* dst, const uint32_t (&v)[16]) {
  uint32_t w[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) w[j] = pack_bf16x2(__uint_as_float(v[2 * j]), __uint_as_float(v[2 * j + 1]));
  if ((reinterpret_cast<uintptr_t>(dst) & 31) == 0) {
    st_global_v8(dst, w);
  } else {
    reinterpret_cast<uint4*>(dst)[0] = make_uint4(w[0], w[1], w[2], w[3]);
    reinterpret_cast<uint4*>(dst)[1] = make_uint4(w[4], w[5], w[6], w[7]);
  }
}

__global__ void __launch_bounds__(FB_THREADS, 1)
attn_bwd_fused_kernel(const __grid_constant__ CUtensorMap tq, const __grid_constant__ CUtensorMap tk, const __grid_constant__ CUtensorMap tv,
                      const __grid_constant__ CUtensorMap tdo, const jl_attn_bwd_params p) {
  const int b = blockIdx.y, h = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  extern __shared__ uint8_t tc_smem_raw[];
  AttnBwdFusedSmem& s = *reinterpret_cast<AttnBwdFusedSmem*>(tc_smem_raw + ((1024u - (ptx::smem_u32(tc_smem_raw) & 1023u)) & 1023u));
  FusedBwdBars& B = s.bars;
  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&tq);
    ptx::prefetch_tensormap(&tk);
    ptx::prefetch_tensormap(&tv);
    ptx::prefetch_tensormap(&tdo);
  }
  if (warp == 1 && lane == 0) {
    ptx::mbar_init(&B.x_full, 1);
    ptx::mbar_init(&B.s_full, 1);
    ptx::mbar_init(&B.s_empty, FB_SOFTMAX_WARPS);
    ptx::mbar_init(&B.p_full, FB_SOFTMAX_WARPS);
    ptx::mbar_init(&B.acc_done, 1);
    ptx::mbar_init(&B.dkv_empty, FB_SOFTMAX_WARPS);
    ptx::fence_barrier_init();
  }
  if (warp == 2) {
    ptx::tmem_alloc(&B.tmem_slot, 512);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  jl::pdl_prologue();
  const TcSeq sq = tc_seq(p, b, h);
  const int len = sq.len, lim = sq.lim;
  const int64_t row_base = sq.row_base;
  const int grow = static_cast<int>(row_base);
  const int nh = (len + TC_OUTER - 1) / TC_OUTER;          // 128-row halves that hold valid frames (0, 1 or 2)
  const int nblocks = nh * nh;                            // block n: key half j = n / nh, query half i = n % nh
  const uint32_t tmem = B.tmem_slot;
  const uint32_t t_s = tmem, t_dp = tmem + 128, t_dq = tmem + 256, t_dk = tmem + 384, t_dv = tmem + 448;

  if (warp == 0) {
    if (lane == 0 && nh > 0) {
      ptx::mbar_expect_tx(&B.x_full, static_cast<uint32_t>(nh) * 4 * TC_T128);
      for (int i = 0; i < nh; ++i) {
        ptx::tma_load_2d(s.q[i], &tq, &B.x_full, h * 64, grow + i * TC_OUTER);
        ptx::tma_load_2d(s.k[i], &tk, &B.x_full, h * 64, grow + i * TC_OUTER);
        ptx::tma_load_2d(s.v[i], &tv, &B.x_full, h * 64, grow + i * TC_OUTER);
        ptx::tma_load_2d(s.d_o[i], &tdo, &B.x_full, h * 64, grow + i * TC_OUTER);
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && nh > 0) {
      const uint32_t aq = ptx::smem_u32(s.q[0]), ak = ptx::smem_u32(s.k[0]), av = ptx::smem_u32(s.v[0]), ado = ptx::smem_u32(s.d_o[0]);
      const uint32_t ap = ptx::smem_u32(s.p[0]), ads = ptx::smem_u32(s.ds[0]);
      ptx::mbar_wait(&B.x_full, 0);
      auto issue_scores = [&](int n) {
        const int j = n / nh, i = n - j * nh;
        if (n > 0) ptx::mbar_wait(&B.s_empty, (n - 1) & 1);               // S / dP of block n-1 are in registers
        ptx::tc_fence_after();
#pragma unroll
        for (int k = 0; k < 4; ++k) {                                      // S = Q_i · K_jᵀ   (M 128, N 128, K 64)
          const uint64_t da = ptx::make_sw128_desc(aq + i * TC_T128 + k * 32, 16, 1024);
          const uint64_t db = ptx::make_sw128_desc(ak + j * TC_T128 + k * 32, 16, 1024);
          ptx::umma_bf16(t_s, da, db, TC_IDESC_S128, k > 0 ? 1u : 0u);
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {                                      // dP = dO_i · V_jᵀ
          const uint64_t da = ptx::make_sw128_desc(ado + i * TC_T128 + k * 32, 16, 1024);
          const uint64_t db = ptx::make_sw128_desc(av + j * TC_T128 + k * 32, 16, 1024);
          ptx::umma_bf16(t_dp, da, db, TC_IDESC_S128, k > 0 ? 1u : 0u);
        }
        ptx::umma_commit(&B.s_full);
      };
      issue_scores(0);
      for (int n = 0; n < nblocks; ++n) {
        const int j = n / nh, i = n - j * nh;
        if (n + 1 < nblocks) issue_scores(n + 1);
        ptx::mbar_wait(&B.p_full, n & 1);
        if (i == 0 && j > 0) ptx::mbar_wait(&B.dkv_empty, (j - 1) & 1);    // dK / dV of the previous key half have been read out
        ptx::tc_fence_after();
        // dV_j += Pᵀ · dO_i, dK_j += dSᵀ · Q_i : A MN-major (M = 128 keys = two 64-wide tiles 16 KB apart, K = 128 queries)
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {
          const uint64_t da = ptx::make_sw128_desc(ap + kk * 2048, TC_T128, 1024);
          const uint64_t db = ptx::make_sw128_desc(ado + i * TC_T128 + kk * 2048, 8192, 1024);
          ptx::umma_bf16(t_dv, da, db, TC_IDESC_DKV, (i > 0 || kk > 0) ? 1u : 0u);
        }
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {
          const uint64_t da = ptx::make_sw128_desc(ads + kk * 2048, TC_T128, 1024);
          const uint64_t db = ptx::make_sw128_desc(aq + i * TC_T128 + kk * 2048, 8192, 1024);
          ptx::umma_bf16(t_dk, da, db, TC_IDESC_DKV, (i > 0 || kk > 0) ? 1u : 0u);
        }
        // dQ_i += dS · K_j : A = dS K-major (two 64-key tiles × 4 k-steps), B = K_j rows as K (MN-major, N = 64 dims)
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {
          const uint64_t da = ptx::make_sw128_desc(ads + (kk >> 2) * TC_T128 + (kk & 3) * 32, 16, 1024);
          const uint64_t db = ptx::make_sw128_desc(ak + j * TC_T128 + kk * 2048, 8192, 1024);
          ptx::umma_bf16(t_dq + i * 64, da, db, TC_IDESC_DQ, (j > 0 || kk > 0) ? 1u : 0u);
        }
        ptx::umma_commit(&B.acc_done);
      }
    }
  } else if (warp >= 4) {
    const int st = threadIdx.x - 128;                        // 0..511
    const int r = (warp & 3) * 32 + lane;                    // TMEM lane = row of the 128-row half
    const int qq = (warp - 4) >> 2;                          // which 32 of the block's 128 key columns this thread owns
    const uint32_t lane_off = static_cast<uint32_t>((warp & 3) * 32) << 16;
    const float sl2 = p.scale * TC_LOG2E;
    const float* lse = p.lse + sq.stat_base;
    __nv_bfloat16* dq = reinterpret_cast<__nv_bfloat16*>(p.dq) + row_base * p.ld_dqkv + h * 64;
    __nv_bfloat16* dk = reinterpret_cast<__nv_bfloat16*>(p.dk) + row_base * p.ld_dqkv + h * 64;
    __nv_bfloat16* dv = reinterpret_cast<__nv_bfloat16*>(p.dv) + row_base * p.ld_dqkv + h * 64;
    // per query row: lse·log2e and delta = Σ_d dO·O.  Two threads share a row (64 contiguous bytes of O and dO each).
    {
      const int row = st >> 1, half = st & 1;
      float acc = 0.0f;
      if (row < len) {
        const uint4* po = reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(p.o) + (row_base + row) * p.ld_o + h * 64) + half * 4;
        const uint4* pd = reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(p.d_o) + (row_base + row) * p.ld_o + h * 64) + half * 4;
        uint4 a[4], d[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) { a[c] = __ldg(po + c); d[c] = __ldg(pd + c); }
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const uint32_t aw[4] = {a[c].x, a[c].y, a[c].z, a[c].w}, dw[4] = {d[c].x, d[c].y, d[c].z, d[c].w};
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float2 fa = unpack_bf16x2(aw[q]), fd = unpack_bf16x2(dw[q]);
            acc = fmaf(fa.x, fd.x, acc);
            acc = fmaf(fa.y, fd.y, acc);
          }
        }
      }
      acc += __shfl_xor_sync(0xffffffffu, acc, 1);
      if (half == 0) {
        s.row_delta[row] = acc;
        s.row_lse[row] = (row < len) ? lse[row] * TC_LOG2E : 0.0f;
        if (p.delta != nullptr && row < lim) p.delta[sq.stat_base + row] = acc;
      }
      asm volatile("bar.sync 1, %0;" ::"n"(FB_SOFTMAX_WARPS * 32) : "memory");
    }
    for (int n = 0; n < nblocks; ++n) {
      const int j = n / nh, i = n - j * nh;
      const int qrow = i * TC_OUTER + r;
      const bool row_ok = qrow < len;
      const float lse_r = s.row_lse[qrow], dl_r = s.row_delta[qrow];
      const int kbase = j * TC_OUTER + qq * 32;              // first key of this thread's 32 columns
      uint8_t* ptile = s.p[qq >> 1];
      uint8_t* dstile = s.ds[qq >> 1];
      ptx::mbar_wait(&B.s_full, n & 1);
      ptx::tc_fence_after();
#pragma unroll
      for (int c = 0; c < 2; ++c) {                          // two chunks of 16 key columns
        uint32_t sv[16], dvv[16], pkp[8], pkd[8];
        ptx::tmem_ld_32x16(t_s + lane_off + qq * 32 + c * 16, sv);
        ptx::tmem_ld_32x16(t_dp + lane_off + qq * 32 + c * 16, dvv);
        ptx::tmem_ld_wait();
        if (c == 1) {
          ptx::tc_fence_before();
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive(&B.s_empty);     // this warp's share of S / dP is in registers
        }
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          float pr[2], dsv[2];
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            const int col = c * 16 + 2 * e + u;
            const bool ok = row_ok && (kbase + col < len);
            const float pe = ok ? tc_exp2(fmaf(__uint_as_float(sv[2 * e + u]), sl2, -lse_r)) : 0.0f;
            pr[u] = pe;
            dsv[u] = pe * (__uint_as_float(dvv[2 * e + u]) - dl_r) * p.scale;
          }
          pkp[e] = pack_bf16x2(pr[0], pr[1]);
          pkd[e] = pack_bf16x2(dsv[0], dsv[1]);
        }
        if (c == 0 && n > 0) ptx::mbar_wait(&B.acc_done, (n - 1) & 1);   // the accumulate MMAs of block n-1 have read P / dS
        tc_store_cols16(ptile, r, (qq & 1) * 4 + c * 2, pkp);
        tc_store_cols16(dstile, r, (qq & 1) * 4 + c * 2, pkd);
      }
      ptx::fence_proxy_async();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&B.p_full);
      if (i == nh - 1) {
        // dK_j, dV_j are complete once this block's accumulate MMAs are: read them out (row = key), free the accumulators
        ptx::mbar_wait(&B.acc_done, n & 1);
        ptx::tc_fence_after();
        const int krow = j * TC_OUTER + r;
        uint32_t ok_[16], ov_[16];
        ptx::tmem_ld_32x16(t_dk + lane_off + qq * 16, ok_);
        ptx::tmem_ld_32x16(t_dv + lane_off + qq * 16, ov_);
        ptx::tmem_ld_wait();
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&B.dkv_empty);
        if (krow < lim) {
          tc_store_global_row16(dk + static_cast<int64_t>(krow) * p.ld_dqkv + qq * 16, ok_);
          tc_store_global_row16(dv + static_cast<int64_t>(krow) * p.ld_dqkv + qq * 16, ov_);
        }
      }
    }
    // dQ of both halves (complete after the last block); rows of halves without valid frames are zero
    if (nblocks > 0) {
      ptx::mbar_wait(&B.acc_done, (nblocks - 1) & 1);
      ptx::tc_fence_after();
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int qrow = i * TC_OUTER + r;
      uint32_t ov[16];
      if (i < nh) {
        ptx::tmem_ld_32x16(t_dq + i * 64 + lane_off + qq * 16, ov);
        ptx::tmem_ld_wait();
      } else {
#pragma unroll
        for (int e = 0; e < 16; ++e) ov[e] = 0u;
      }
      if (qrow < lim) {
        tc_store_global_row16(dq + static_cast<int64_t>(qrow) * p.ld_dqkv + qq * 16, ov);
        if (i >= nh) {                                       // key rows of an all-padding half: dK = dV = 0
          tc_store_global_row16(dk + static_cast<int64_t>(qrow) * p.ld_dqkv + qq * 16, ov);
          tc_store_global_row16(dv + static_cast<int64_t>(qrow) * p.ld_dqkv + qq * 16, ov);
        }
      }
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem, 512);
  }
}

// ------------------------------------------------------------------------------------------------ host
template <typename K>
static int tc_set_smem(K kern, size_t bytes, const char* name) {
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(bytes));
  JL_REQUIRE(e == cudaSuccess, JL_ECUDA, "%s: cannot reserve %zu B of shared memory: %s", name, bytes, cudaGetErrorString(e));
  return JL_OK;
}

int g_attn_fwd_ctas = 2;   // resident CTAs per SM the forward kernel is compiled for
int g_attn_short = 1;      // 1: utterances of <= 256 frames take the whole-row forward kernel and the fused dQ/dK/dV backward kernel

int attn_fwd_tc(const jl_attn_fwd_params* p, cudaStream_t stream) {
  const int64_t rows = p->cu_seqlens ? static_cast<int64_t>(p->total_rows) : static_cast<int64_t>(p->batch) * p->seq;
  const int64_t inner = static_cast<int64_t>(p->heads) * 64;
  CUtensorMap tq, tk, tv;
  int rc = make_tma_map_2d_bf16(&tq, p->q, inner, rows, p->ld_qkv, TC_OUTER);
  if (rc == JL_OK) rc = make_tma_map_2d_bf16(&tk, p->k, inner, rows, p->ld_qkv, TC_INNER);
  if (rc == JL_OK) rc = make_tma_map_2d_bf16(&tv, p->v, inner, rows, p->ld_qkv, TC_INNER);
  if (rc != JL_OK) return rc;
  if (g_attn_short && p->seq <= 2 * TC_OUTER) {
    const size_t smem_s = sizeof(AttnFwdShortSmem) + 1024;
    static thread_local int configured_short = -1;
    int dev_s = 0;
    cudaGetDevice(&dev_s);
    if (configured_short != dev_s) {
      rc = tc_set_smem(attn_fwd_short_kernel, smem_s, "attn_fwd_short");
      if (rc != JL_OK) return rc;
      configured_short = dev_s;
    }
    jl::launch(attn_fwd_short_kernel, dim3(ceil_div(p->seq, TC_OUTER), p->heads, p->batch), TC_THREADS, smem_s, stream, tq, tk, tv, *p);
    JL_CHECK_LAUNCH("attn_fwd_short");
    return JL_OK;
  }
  const size_t smem = sizeof(AttnFwdSmem) + 1024;
  static thread_local int configured_dev = -1;
  int dev = 0;
  cudaGetDevice(&dev);
  if (configured_dev != dev) {
    rc = tc_set_smem(attn_fwd_tc_kernel<3>, smem, "attn_fwd_tc");
    if (rc == JL_OK) rc = tc_set_smem(attn_fwd_tc_kernel<2>, smem, "attn_fwd_tc");
    if (rc != JL_OK) return rc;
    configured_dev = dev;
  }
  dim3 grid(ceil_div(p->seq, TC_OUTER), p->heads, p->batch);
  if (g_attn_fwd_ctas == 3) jl::launch(attn_fwd_tc_kernel<3>, grid, TC_THREADS, smem, stream, tq, tk, tv, *p);
  else jl::launch(attn_fwd_tc_kernel<2>, grid, TC_THREADS, smem, stream, tq, tk, tv, *p);
  JL_CHECK_LAUNCH("attn_fwd_tc");
  return JL_OK;
}

int attn_bwd_tc(const jl_attn_bwd_params* p, cudaStream_t stream) {
  const int64_t rows = p->cu_seqlens ? static_cast<int64_t>(p->total_rows) : static_cast<int64_t>(p->batch) * p->seq;
  const int64_t inner = static_cast<int64_t>(p->heads) * 64;
  CUtensorMap q128, do128, k64, v64, k128, v128, q64, do64;
  int rc = make_tma_map_2d_bf16(&q128, p->q, inner, rows, p->ld_qkv, TC_OUTER);
  if (rc == JL_OK) rc = make_tma_map_2d_bf16(&do128, p->d_o, inner, rows, p->ld_o, TC_OUTER);
  if (rc == JL_OK) rc = make_tma_map_2d_bf16(&k64, p->k, inner, rows, p->ld_qkv, TC_INNER);
  if (rc == JL_OK) rc = make_tma_map_2d_bf16(&v64, p->v, inner, rows, p->ld_qkv, TC_INNER);
  if (rc == JL_OK) rc = make_tma_map_2d_bf16(&k128, p->k, inner, rows, p->ld_qkv, TC_OUTER);
  if (rc == JL_OK) rc = make_tma_map_2d_bf16(&v128, p->v, inner, rows, p->ld_qkv, TC_OUTER);
  if (rc == JL_OK) rc = make_tma_map_2d_bf16(&q64, p->q, inner, rows, p->ld_qkv, TC_INNER);
  if (rc == JL_OK) rc = make_tma_map_2d_bf16(&do64, p->d_o, inner, rows, p->ld_o, TC_INNER);
  if (rc != JL_OK) return rc;
  if (g_attn_short && p->seq <= 2 * TC_OUTER) {
    // whole utterance in one CTA: scores, exponentials and operand loads once for dQ, dK and dV
    const size_t smem_f = sizeof(AttnBwdFusedSmem) + 1024;
    static thread_local int configured_dev_f = -1;
    int devf = 0;
    cudaGetDevice(&devf);
    if (configured_dev_f != devf) {
      rc = tc_set_smem(attn_bwd_fused_kernel, smem_f, "attn_bwd_fused");
      if (rc != JL_OK) return rc;
      configured_dev_f = devf;
    }
    jl::launch(attn_bwd_fused_kernel, dim3(p->heads, p->batch), FB_THREADS, smem_f, stream, q128, k128, v128, do128, *p);
    JL_CHECK_LAUNCH("attn_bwd_fused");
    return JL_OK;
  }
  const size_t smem0 = sizeof(AttnBwdSmem<0>) + 1024, smem1 = sizeof(AttnBwdSmem<1>) + 1024;
  static thread_local int configured_dev = -1;
  int dev = 0;
  cudaGetDevice(&dev);
  if (configured_dev != dev) {
    rc = tc_set_smem(attn_bwd_tc_kernel<0>, smem0, "attn_bwd_tc<dq>");
    if (rc == JL_OK) rc = tc_set_smem(attn_bwd_tc_kernel<1>, smem1, "attn_bwd_tc<dkv>");
    if (rc != JL_OK) return rc;
    configured_dev = dev;
  }
  dim3 grid(ceil_div(p->seq, TC_OUTER), p->heads, p->batch);
  jl::launch(attn_bwd_tc_kernel<0>, grid, TC_THREADS, smem0, stream, q128, do128, k64, v64, *p);
  JL_CHECK_LAUNCH("attn_bwd_tc_dq");
  jl::launch(attn_bwd_tc_kernel<1>, grid, TC_THREADS, smem1, stream, k128, v128, q64, do64, *p);
  JL_CHECK_LAUNCH("attn_bwd_tc_dkv");
  return JL_OK;
}

}  // namespace jl
