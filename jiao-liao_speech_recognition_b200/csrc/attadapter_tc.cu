// a7: the AttAdapter forward as ONE kernel (north_star: "AttAdapter's small attention is a single fused kernel"):
//
//   z = LN(h);  q|k|v = z W_qkvᵀ + b  (∈ R^64 each);  a = softmax(q kᵀ / 8 + keymask) v  over the utterance's own frames;
//   out = h + a W_oᵀ + b_o                                            (/root/reference/README.md:1)
//
// for utterances of at most 256 frames (10.2 s at 40 ms — the benchmark's case; longer ones take the composed path: LayerNorm →
// GEMM → jl_attn_fwd → GEMM).  It replaces four launches (and their three HBM round trips of z, q|k|v and a) by one:
// h is read once for the projections and once more (L2) for the residual, out is written once.
//
// One CTA per (utterance, 128-query tile), 384 threads:
//   warp 0   TMA producer: 64-wide k-chunks of h and of W' = W_qkv ⊙ γ through a 3-stage ring, W_o in 256-row chunks
//   warp 1   tcgen05.mma issuer, accumulators in TMEM (512 columns)
//   warp 2   TMEM allocator
//   warps 4-11  row statistics (4 warps), then every epilogue: LayerNorm fold + bias → bf16 operand tiles, softmax, a, output
// Phases:
//   A  acc_own[128, 192] = h_own · W'ᵀ  (q, k, v of the CTA's own 128 frames);  when the utterance has frames in the OTHER
//      128-row half, acc_oth[128, 128] = h_oth · W'_kvᵀ (their k, v: the attention needs every key; recomputing them here
//      costs one more pass over 196 KB of L2-resident h and avoids a cluster exchange).  The LayerNorm is folded into the
//      projection, (LN(h) Wᵀ)[i, j] = rstd_i (h W'ᵀ[i, j] − μ_i s_j) + t_j, so the tensor cores consume the raw h tiles
//      while the row threads accumulate Σx, Σx² from the same shared-memory tiles (as in wfadapter_tc.cu).
//   B  fold + bias → q, k, v as K-major 128B-swizzled bf16 operand tiles in shared memory (+ q|k|v, statistics to HBM when training)
//   C  S[128, 256] = q · kᵀ;  two-pass softmax over whole rows (8 warps);  P → shared memory;  O[128, 64] = P · v
//   D  a = O / l → bf16 operand tile (+ a, lse to HBM when training);  out[128, d] = a · W_oᵀ + b_o + h in chunks of 256 columns,
//      accumulators double-buffered in TMEM against the store
#include <cuda.h>
#include <math_constants.h>

#include "common.cuh"
#include "ptx_sm100.cuh"

namespace jl {

constexpr int AA_THREADS = 384;
constexpr int AA_STAGES = 3;                        // 3 x 40 KB ring + 64 KB of W_o chunks + the barriers fit the 227 KB of a CTA
constexpr uint32_t AA_T128 = 128 * 128;             // bytes of a [128 x 64] bf16 tile
constexpr uint32_t AA_WQKV = 192 * 128;             // bytes of a [192 x 64] bf16 tile
constexpr uint32_t AA_STAGE = AA_T128 + AA_WQKV;    // 40 KB
constexpr float AA_LOG2E = 1.4426950408889634f;

struct __align__(1024) AaSmem {
  uint8_t ring[AA_STAGES * AA_STAGE];   // phase A ring; afterwards: q | k[2] | (P tail) | v[2] | a operand tiles (see offsets below)
  uint8_t wo[2][256 * 128];             // W_o chunks [256 rows x 64]
  float mu[2][128], rs[2][128];         // LayerNorm statistics: [0] own rows, [1] rows of the other half
  float red_max[2][128], red_sum[2][128];
  uint64_t full[AA_STAGES], empty[AA_STAGES];
  uint64_t acc_full;                    // phase A accumulators complete
  uint64_t kv_ready;                    // q, k, v operand tiles written (8 warps)
  uint64_t s_full, p_full, o_full, a_ready;
  uint64_t wo_full[2], wo_empty[2], out_full[2], out_empty[2];
  uint32_t tmem_slot;
};
// operand tiles inside `ring` once phase A is over
constexpr uint32_t AA_OFF_Q = 0;                    // [128 x 64]            later P tile 0
constexpr uint32_t AA_OFF_K = AA_T128;              // [256 x 64] (2 tiles)  later P tiles 1, 2
constexpr uint32_t AA_OFF_PT = 3 * AA_T128;         //                       P tile 3
constexpr uint32_t AA_OFF_V = 4 * AA_T128;          // [256 x 64] = four 64-key tiles, read as MN-major B operands
constexpr uint32_t AA_OFF_A = 6 * AA_T128;          // [128 x 64]

__device__ __forceinline__ float aa_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void aa_store_chunk(uint8_t* tile, int r, int c, const uint32_t* pk) {      // 8 bf16 = 16 B, chunk c of row r
  *reinterpret_cast<uint4*>(tile + r * 128 + ((c ^ (r & 7)) << 4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
}
// D[128 x n] (+)= A[128 x 64] · B[n x 64]ᵀ, both K-major 128B-swizzled; 4 MMAs of K = 16
__device__ __forceinline__ void aa_mma_kk(uint32_t d_tmem, uint32_t a_addr, uint32_t b_addr, int n, bool accumulate) {
  const uint32_t idesc = ptx::make_idesc_bf16_f32(128, n);
#pragma unroll
  for (int k = 0; k < 4; ++k)
    ptx::umma_bf16(d_tmem, ptx::make_sw128_desc(a_addr + k * 32, 16, 1024), ptx::make_sw128_desc(b_addr + k * 32, 16, 1024), idesc,
                   (accumulate || k > 0) ? 1u : 0u);
}
// O[128 x 64] (+)= P[128 x 64 keys] · V[64 keys x 64], V tile read MN-major (rows are K)
__device__ __forceinline__ void aa_mma_pv(uint32_t d_tmem, uint32_t p_addr, uint32_t v_addr, bool accumulate) {
  const uint32_t idesc = ptx::make_idesc_bf16_f32(128, 64) | (1u << 16);
#pragma unroll
  for (int k = 0; k < 4; ++k)
    ptx::umma_bf16(d_tmem, ptx::make_sw128_desc(p_addr + k * 32, 16, 1024), ptx::make_sw128_desc(v_addr + k * 2048, 8192, 1024), idesc,
                   (accumulate || k > 0) ? 1u : 0u);
}

#ifdef JL_AA_TIMING
#define AA_T(i) do { if (threadIdx.x == 128 && blockIdx.x == 0 && blockIdx.y == 0) aa_ts[i] = clock64(); } while (0)
#else
#define AA_T(i) do { } while (0)
#endif

__global__ void __launch_bounds__(AA_THREADS, 1)
attadapter_fwd_kernel(const __grid_constant__ CUtensorMap t_h, const __grid_constant__ CUtensorMap t_w, const __grid_constant__ CUtensorMap t_wkv,
                      const __grid_constant__ CUtensorMap t_wo, const jl_attadapter_fwd_params p) {
  extern __shared__ uint8_t aa_smem_raw[];
  AaSmem& s = *reinterpret_cast<AaSmem*>(aa_smem_raw + ((1024u - (ptx::smem_u32(aa_smem_raw) & 1023u)) & 1023u));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = blockIdx.x, b = blockIdx.y;         // query half, utterance
  const int nk = p.d / 64;                          // k-chunks of the projections
  const int nc = (p.d + 255) / 256;                 // 256-column chunks of the output projection

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&t_h);
    ptx::prefetch_tensormap(&t_w);
    ptx::prefetch_tensormap(&t_wkv);
    ptx::prefetch_tensormap(&t_wo);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < AA_STAGES; ++i) {
      ptx::mbar_init(&s.full[i], 1);
      ptx::mbar_init(&s.empty[i], 5);         // MMA commit + the 4 statistics warps
    }
    ptx::mbar_init(&s.acc_full, 1);
    ptx::mbar_init(&s.kv_ready, 8);
    ptx::mbar_init(&s.s_full, 1);
    ptx::mbar_init(&s.p_full, 8);
    ptx::mbar_init(&s.o_full, 1);
    ptx::mbar_init(&s.a_ready, 8);
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(&s.wo_full[i], 1);
      ptx::mbar_init(&s.wo_empty[i], 1);
      ptx::mbar_init(&s.out_full[i], 1);
      ptx::mbar_init(&s.out_empty[i], 8);
    }
    ptx::fence_barrier_init();
  }
  if (warp == 2) {
    ptx::tmem_alloc(&s.tmem_slot, 512);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = s.tmem_slot;
  jl::pdl_prologue();           // h, lengths and the packed weights may come from the preceding kernels
#ifdef JL_AA_TIMING
  long long aa_ts[12];
  for (int i = 0; i < 12; ++i) aa_ts[i] = 0;
#endif
  AA_T(0);

  // where the utterance lives (padded rows b·seq + t, or packed rows cu[b] + t)
  int64_t row_base;
  int len, lim;
  if (p.cu_seqlens != nullptr) {
    const int r0 = p.cu_seqlens[b];
    row_base = r0;
    len = min(p.cu_seqlens[b + 1] - r0, p.seq);
    lim = len;
  } else {
    row_base = static_cast<int64_t>(b) * p.seq;
    len = min(p.lengths ? p.lengths[b] : p.seq, p.seq);
    lim = p.seq;
  }
  const int grow = static_cast<int>(row_base);
  const bool active = g * 128 < len;                   // the CTA's query tile holds at least one valid frame
  const bool other = active && len > 128;              // the other 128-row half holds keys
  const int og = g ^ 1;
  const int npass = other ? 2 : 1;
  const int nkt = (len + 63) / 64;                     // 64-key tiles with valid keys
  const uint32_t t_acc_own = tmem, t_acc_oth = tmem + 192;
  const uint32_t t_s = tmem, t_o = tmem + 256;
  const uint32_t t_out[2] = {tmem, tmem + 256};
  uint8_t* R = s.ring;

  if (warp == 0) {
    if (lane == 0 && active) {
      // W_o chunks 0, 1 right away (their buffers are not shared with anything)
      for (int c = 0; c < min(nc, 2); ++c) {
        ptx::mbar_expect_tx(&s.wo_full[c], 256 * 128);
        ptx::tma_load_2d(s.wo[c], &t_wo, &s.wo_full[c], 0, c * 256);
      }
      int it = 0;
      for (int pass = 0; pass < npass; ++pass) {
        const int r0 = grow + (pass == 0 ? g : og) * 128;
        for (int kc = 0; kc < nk; ++kc, ++it) {
          const int st = it % AA_STAGES;
          ptx::mbar_wait(&s.empty[st], ((it / AA_STAGES) & 1) ^ 1u);
          uint8_t* base = R + st * AA_STAGE;
          if (pass == 0) {
            ptx::mbar_expect_tx(&s.full[st], AA_STAGE);
            ptx::tma_load_2d(base, &t_h, &s.full[st], kc * 64, r0);
            ptx::tma_load_2d(base + AA_T128, &t_w, &s.full[st], kc * 64, 0);
          } else {
            ptx::mbar_expect_tx(&s.full[st], 2 * AA_T128);
            ptx::tma_load_2d(base, &t_h, &s.full[st], kc * 64, r0);
            ptx::tma_load_2d(base + AA_T128, &t_wkv, &s.full[st], kc * 64, 0);
          }
        }
      }
      for (int c = 2; c < nc; ++c) {
        const int st = c & 1;
        ptx::mbar_wait(&s.wo_empty[st], ((c >> 1) & 1) ^ 1u);
        ptx::mbar_expect_tx(&s.wo_full[st], 256 * 128);
        ptx::tma_load_2d(s.wo[st], &t_wo, &s.wo_full[st], 0, c * 256);
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && active) {
      // ---- phase A
      int it = 0;
      for (int pass = 0; pass < npass; ++pass) {
        for (int kc = 0; kc < nk; ++kc, ++it) {
          const int st = it % AA_STAGES;
          ptx::mbar_wait(&s.full[st], (it / AA_STAGES) & 1);
          ptx::tc_fence_after();
          const uint32_t base = ptx::smem_u32(R + st * AA_STAGE);
          if (pass == 0) aa_mma_kk(t_acc_own, base, base + AA_T128, 192, kc > 0);
          else aa_mma_kk(t_acc_oth, base, base + AA_T128, 128, kc > 0);
          ptx::umma_commit(&s.empty[st]);
        }
      }
      ptx::umma_commit(&s.acc_full);
      // ---- phase C: S = q · kᵀ (N = 256 keys), then O = P · v over the key tiles that hold frames
      ptx::mbar_wait(&s.kv_ready, 0);
      ptx::tc_fence_after();
      const uint32_t rb = ptx::smem_u32(R);
      aa_mma_kk(t_s, rb + AA_OFF_Q, rb + AA_OFF_K, 256, false);
      ptx::umma_commit(&s.s_full);
      ptx::mbar_wait(&s.p_full, 0);
      ptx::tc_fence_after();
      for (int kt = 0; kt < nkt; ++kt) aa_mma_pv(t_o, rb + kt * AA_T128, rb + AA_OFF_V + kt * (64 * 128), kt > 0);
      ptx::umma_commit(&s.o_full);
      // ---- phase D: out chunk c = a · W_o[c]ᵀ
      ptx::mbar_wait(&s.a_ready, 0);
      ptx::tc_fence_after();
      for (int c = 0; c < nc; ++c) {
        const int st = c & 1;
        const int ncols = min(256, p.d - c * 256);
        ptx::mbar_wait(&s.wo_full[st], (c >> 1) & 1);
        ptx::mbar_wait(&s.out_empty[st], ((c >> 1) & 1) ^ 1u);
        ptx::tc_fence_after();
        aa_mma_kk(t_out[st], rb + AA_OFF_A, ptx::smem_u32(s.wo[st]), ncols, false);
        ptx::umma_commit(&s.wo_empty[st]);
        ptx::umma_commit(&s.out_full[st]);
      }
    }
  } else if (warp >= 4) {
    const int quad = warp & 3;
    const int grp = (warp - 4) >> 2;                     // 0: warps 4-7, 1: warps 8-11
    const int r = quad * 32 + lane;                      // TMEM lane = row of the 128-row tile
    const uint32_t lane_off = static_cast<uint32_t>(quad * 32) << 16;
    const int qrow = g * 128 + r;                        // frame index inside the utterance
    __nv_bfloat16* out_row = reinterpret_cast<__nv_bfloat16*>(p.out) + (row_base + qrow) * p.ldo;
    const __nv_bfloat16* h_row = reinterpret_cast<const __nv_bfloat16*>(p.h) + (row_base + qrow) * p.ldh;
    if (!active) {
      // no valid frame in this tile: the rows the layout still owns get what the composed path gives them — a = 0, so
      // out = b_o + h (or 0 when the caller wants padded rows zeroed); saved tensors are zero there
      if (qrow < lim) {
        for (int c = grp * (p.d / 16); c < (grp + 1) * (p.d / 16); ++c) {          // 8 columns per step, half the row per warp group
          float o[8];
          if (p.zero_padded_rows) {
#pragma unroll
            for (int j = 0; j < 8; ++j) o[j] = 0.0f;
          } else {
            const uint4 hv = __ldg(reinterpret_cast<const uint4*>(h_row) + c);
            const float2 f0 = unpack_bf16x2(hv.x), f1 = unpack_bf16x2(hv.y), f2 = unpack_bf16x2(hv.z), f3 = unpack_bf16x2(hv.w);
            const float hh[8] = {f0.x, f0.y, f1.x, f1.y, f2.x, f2.y, f3.x, f3.y};
#pragma unroll
            for (int j = 0; j < 8; ++j) o[j] = hh[j] + __ldg(p.bo + c * 8 + j);
          }
          reinterpret_cast<uint4*>(out_row)[c] = make_uint4(pack_bf16x2(o[0], o[1]), pack_bf16x2(o[2], o[3]), pack_bf16x2(o[4], o[5]), pack_bf16x2(o[6], o[7]));
        }
        if (grp == 0) {
          if (p.qkv_out != nullptr) {
            uint4* qd = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.qkv_out) + (row_base + qrow) * 192);
            for (int c = 0; c < 24; ++c) qd[c] = make_uint4(0u, 0u, 0u, 0u);
          }
          if (p.a_out != nullptr) {
            uint4* ad = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.a_out) + (row_base + qrow) * 64);
            for (int c = 0; c < 8; ++c) ad[c] = make_uint4(0u, 0u, 0u, 0u);
          }
          if (p.mean != nullptr) { p.mean[row_base + qrow] = 0.0f; p.rstd[row_base + qrow] = 0.0f; }
          if (p.lse != nullptr) p.lse[(p.cu_seqlens ? row_base : static_cast<int64_t>(b) * p.seq) + qrow] = 0.0f;
        }
      }
    } else {
      // ---- phase A (warps 4-7): LayerNorm statistics of the staged h tiles, own rows then the other half's
      if (grp == 0) {
        int it = 0;
        for (int pass = 0; pass < npass; ++pass) {
          float sx = 0.0f, sxx = 0.0f;
          for (int kc = 0; kc < nk; ++kc, ++it) {
            const int st = it % AA_STAGES;
            ptx::mbar_wait(&s.full[st], (it / AA_STAGES) & 1);
            const uint8_t* tile = R + st * AA_STAGE;
#pragma unroll
            for (int c = 0; c < 8; ++c) {
              const uint4 v = *reinterpret_cast<const uint4*>(tile + r * 128 + ((c ^ (r & 7)) << 4));
              const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                const float2 f = unpack_bf16x2(w[q]);
                sx += f.x + f.y;
                sxx = fmaf(f.x, f.x, fmaf(f.y, f.y, sxx));
              }
            }
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(&s.empty[st]);
          }
          const float inv_d = 1.0f / static_cast<float>(p.d);
          const float mu = sx * inv_d;
          const float var = fmaxf(sxx * inv_d - mu * mu, 0.0f);
          const float rstd = 1.0f / sqrtf(var + p.eps);
          s.mu[pass][r] = mu;
          s.rs[pass][r] = rstd;
          if (pass == 0 && p.mean != nullptr && qrow < lim) {
            p.mean[row_base + qrow] = mu;
            p.rstd[row_base + qrow] = rstd;
          }
        }
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");       // statistics visible to all 8 warps
      AA_T(1);
      // ---- phase B: fold + bias → q, k, v operand tiles (and q|k|v to HBM).  Own rows: 6 chunks of 32 columns (q0 q1 k0 k1 v0 v1),
      //      other rows: 4 chunks (k0 k1 v0 v1); warp group 0 takes own 0-2 + other 0-1, group 1 own 3-5 + other 2-3.
      ptx::mbar_wait(&s.acc_full, 0);
      ptx::tc_fence_after();
      AA_T(2);
      __nv_bfloat16* qkv_row = (p.qkv_out != nullptr && qrow < lim) ? reinterpret_cast<__nv_bfloat16*>(p.qkv_out) + (row_base + qrow) * 192 : nullptr;
#pragma unroll 1
      for (int item = 0; item < 5; ++item) {
        const bool own = item < 3;
        if (!own && !other) break;
        const int ch = own ? grp * 3 + item : grp * 2 + (item - 3);          // chunk index inside the accumulator
        const int col0 = own ? ch * 32 : 64 + ch * 32;                       // column of W_qkv (q 0-63, k 64-127, v 128-191)
        const float mu = s.mu[own ? 0 : 1][r], rstd = s.rs[own ? 0 : 1][r];
        uint32_t v[32];
        ptx::tmem_ld_32x32((own ? t_acc_own : t_acc_oth) + lane_off + ch * 32, v);
        ptx::tmem_ld_wait();
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const int j = col0 + 2 * i;
          const float x0 = fmaf(rstd, __uint_as_float(v[2 * i]) - mu * __ldg(p.s + j), __ldg(p.tb + j));
          const float x1 = fmaf(rstd, __uint_as_float(v[2 * i + 1]) - mu * __ldg(p.s + j + 1), __ldg(p.tb + j + 1));
          pk[i] = pack_bf16x2(x0, x1);
        }
        // destination tile: q; k / v of the half the rows belong to (utterance order: half 0 = frames 0-127)
        const int hidx = own ? g : og;
        uint8_t* tile;
        if (col0 < 64) tile = R + AA_OFF_Q;
        else if (col0 < 128) tile = R + AA_OFF_K + hidx * AA_T128;
        else tile = R + AA_OFF_V + hidx * AA_T128;
        const int c0 = ((col0 & 63) >> 3);                                   // first 16-byte chunk inside the 64-wide tile row: 0 or 4
#pragma unroll
        for (int c = 0; c < 4; ++c) aa_store_chunk(tile, r, c0 + c, pk + 4 * c);
        if (own && qkv_row != nullptr) {
#pragma unroll
          for (int c = 0; c < 4; ++c) reinterpret_cast<uint4*>(qkv_row + col0)[c] = make_uint4(pk[4 * c], pk[4 * c + 1], pk[4 * c + 2], pk[4 * c + 3]);
        }
      }
      ptx::tc_fence_before();
      ptx::fence_proxy_async();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&s.kv_ready);
      AA_T(3);
      // ---- phase C: softmax over the whole row of 256 scores (this thread: 128 of them), P → operand tiles over q / k
      const float sl2 = p.scale * AA_LOG2E;
      const int kbase = grp * 128;
      const uint32_t t_srow = t_s + lane_off + 128u * grp;
      ptx::mbar_wait(&s.s_full, 0);
      ptx::tc_fence_after();
      AA_T(4);
      float mloc = -CUDART_INF_F;
#pragma unroll 1
      for (int c = 0; c < 4; c += 2) {
        uint32_t sa[32], sb[32];
        ptx::tmem_ld_32x32(t_srow + 32u * c, sa);
        ptx::tmem_ld_32x32(t_srow + 32u * (c + 1), sb);
        ptx::tmem_ld_wait();
        const int k0 = kbase + 32 * c;
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          if (k0 + i < len) mloc = fmaxf(mloc, __uint_as_float(sa[i]));
          if (k0 + 32 + i < len) mloc = fmaxf(mloc, __uint_as_float(sb[i]));
        }
      }
      s.red_max[grp][r] = mloc;
      asm volatile("bar.sync 2, 256;" ::: "memory");
      const float mx = fmaxf(mloc, s.red_max[grp ^ 1][r]);          // finite: the utterance has at least one key
      const float mxs = mx * sl2;
      float sum = 0.0f;
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        uint32_t sv[32];
        ptx::tmem_ld_32x32(t_srow + 32u * c, sv);
        ptx::tmem_ld_wait();
        const int k0 = kbase + 32 * c;
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          float a0 = aa_exp2(fmaf(__uint_as_float(sv[2 * i]), sl2, -mxs));
          float a1 = aa_exp2(fmaf(__uint_as_float(sv[2 * i + 1]), sl2, -mxs));
          if (k0 + 2 * i >= len) a0 = 0.0f;
          if (k0 + 2 * i + 1 >= len) a1 = 0.0f;
          sum += a0 + a1;
          pk[i] = pack_bf16x2(a0, a1);
        }
        uint8_t* tile = R + (k0 >> 6) * AA_T128;                      // P tile of 64 keys
        const int c0 = ((k0 & 63) >> 3);
#pragma unroll
        for (int cc = 0; cc < 4; ++cc) aa_store_chunk(tile, r, c0 + cc, pk + 4 * cc);
      }
      ptx::tc_fence_before();
      ptx::fence_proxy_async();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&s.p_full);
      AA_T(5);
      s.red_sum[grp][r] = sum;
      // ---- phase D: a = O / l (this thread: 32 of the 64 dims) → operand tile, HBM; lse
      ptx::mbar_wait(&s.o_full, 0);
      ptx::tc_fence_after();
      AA_T(6);
      asm volatile("bar.sync 2, 256;" ::: "memory");
      const float l_tot = sum + s.red_sum[grp ^ 1][r];
      const bool valid = qrow < len;
      const float inv = valid ? 1.0f / l_tot : 0.0f;
      {
        uint32_t ov[32];
        ptx::tmem_ld_32x32(t_o + lane_off + 32u * grp, ov);
        ptx::tmem_ld_wait();
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) pk[i] = pack_bf16x2(__uint_as_float(ov[2 * i]) * inv, __uint_as_float(ov[2 * i + 1]) * inv);
#pragma unroll
        for (int c = 0; c < 4; ++c) aa_store_chunk(R + AA_OFF_A, r, grp * 4 + c, pk + 4 * c);
        if (p.a_out != nullptr && qrow < lim) {
          uint4* ad = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.a_out) + (row_base + qrow) * 64 + grp * 32);
#pragma unroll
          for (int c = 0; c < 4; ++c) ad[c] = make_uint4(pk[4 * c], pk[4 * c + 1], pk[4 * c + 2], pk[4 * c + 3]);
        }
        if (p.lse != nullptr && grp == 0 && qrow < lim)
          p.lse[(p.cu_seqlens ? row_base : static_cast<int64_t>(b) * p.seq) + qrow] = valid ? mx * p.scale + logf(l_tot) : 0.0f;
      }
      ptx::tc_fence_before();
      ptx::fence_proxy_async();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&s.a_ready);
      AA_T(7);
      // ---- output projection epilogue: out = acc + b_o + h, 32 columns per step; this warp group takes column groups grp, grp + 2, …
      const bool write = qrow < lim;
      const bool zero = p.zero_padded_rows && !valid;
      for (int c = 0; c < nc; ++c) {
        const int st = c & 1;
        const int ngrp = min(256, p.d - c * 256) / 32;
        ptx::mbar_wait(&s.out_full[st], (c >> 1) & 1);
        ptx::tc_fence_after();
        if (c == 0) AA_T(8);
#pragma unroll 1
        for (int q = grp; q < ngrp; q += 2) {
          const int col = c * 256 + q * 32;
          uint32_t hres[2][8];
          if (write && !zero) {
            ld_global_nc_v8(h_row + col, hres[0]);
            ld_global_nc_v8(h_row + col + 16, hres[1]);
          }
          uint32_t v[32];
          ptx::tmem_ld_32x32(t_out[st] + lane_off + q * 32, v);
          ptx::tmem_ld_wait();
          if (write) {
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
              uint32_t w[8];
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const int j = hh * 16 + 2 * i;
                float o0 = 0.0f, o1 = 0.0f;
                if (!zero) {
                  const float2 hr = unpack_bf16x2(hres[hh][i]);
                  o0 = __uint_as_float(v[j]) + __ldg(p.bo + col + j) + hr.x;
                  o1 = __uint_as_float(v[j + 1]) + __ldg(p.bo + col + j + 1) + hr.y;
                }
                w[i] = pack_bf16x2(o0, o1);
              }
              st_global_v8(out_row + col + hh * 16, w);
            }
          }
        }
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&s.out_empty[st]);
      }
      AA_T(9);
#ifdef JL_AA_TIMING
      if (threadIdx.x == 128 && blockIdx.x == 0 && blockIdx.y == 0)
        printf("aa phases (cycles): stats %lld  acc_wait %lld  fold %lld  s_wait %lld  softmax %lld  o_wait %lld  a %lld  out_wait %lld  out %lld  total %lld\n",
               aa_ts[1] - aa_ts[0], aa_ts[2] - aa_ts[1], aa_ts[3] - aa_ts[2], aa_ts[4] - aa_ts[3], aa_ts[5] - aa_ts[4], aa_ts[6] - aa_ts[5],
               aa_ts[7] - aa_ts[6], aa_ts[8] - aa_ts[7], aa_ts[9] - aa_ts[8], aa_ts[9] - aa_ts[0]);
#endif
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem, 512);
  }
}

// LayerNorm-fold packing of a projection that follows a LayerNorm (see jl_lnfold_pack): one CTA per output row j.
__global__ void __launch_bounds__(256) lnfold_pack_kernel(const jl_lnfold_pack_params p) {
  jl::pdl_prologue();
  __shared__ float red_s[8], red_t[8];
  const int j = blockIdx.x;
  const __nv_bfloat16* w = reinterpret_cast<const __nv_bfloat16*>(p.w) + static_cast<int64_t>(j) * p.d;
  __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(p.w_scaled) + static_cast<int64_t>(j) * p.d;
  float ss = 0.0f, tt = 0.0f;
  for (int c = threadIdx.x; c < p.d; c += blockDim.x) {
    const float x = __bfloat162float(w[c]);
    const __nv_bfloat16 xs = __float2bfloat16_rn(x * __ldg(p.gamma + c));
    out[c] = xs;
    ss += __bfloat162float(xs);
    tt = fmaf(x, __ldg(p.beta + c), tt);
  }
  ss = warp_sum(ss);
  tt = warp_sum(tt);
  if ((threadIdx.x & 31) == 0) { red_s[threadIdx.x >> 5] = ss; red_t[threadIdx.x >> 5] = tt; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float a = 0.0f, b = 0.0f;
    for (int q = 0; q < static_cast<int>(blockDim.x >> 5); ++q) { a += red_s[q]; b += red_t[q]; }
    p.s[j] = a;
    p.tb[j] = b + (p.bias != nullptr ? __ldg(p.bias + j) : 0.0f);
  }
}

}  // namespace jl

extern "C" {

int jl_lnfold_pack(const jl_lnfold_pack_params* p, void* stream) {
  JL_REQUIRE(p != nullptr && p->w && p->gamma && p->beta && p->w_scaled && p->s && p->tb, JL_EINVAL, "lnfold_pack: null pointer");
  JL_REQUIRE(p->n > 0 && p->d > 0, JL_EINVAL, "lnfold_pack: bad dims");
  int rc = jl::check_device();
  if (rc != JL_OK) return rc;
  jl::launch(jl::lnfold_pack_kernel, p->n, 256, 0, reinterpret_cast<cudaStream_t>(stream), *p);
  JL_CHECK_LAUNCH("lnfold_pack");
  return JL_OK;
}

int jl_attadapter_fwd(const jl_attadapter_fwd_params* p, void* stream) {
  JL_REQUIRE(p != nullptr, JL_EINVAL, "attadapter_fwd: null params");
  JL_REQUIRE(p->h && p->out && p->wqkv_scaled && p->s && p->tb && p->wo && p->bo, JL_EINVAL, "attadapter_fwd: null pointer");
  JL_REQUIRE(p->batch > 0 && p->seq > 0, JL_EINVAL, "attadapter_fwd: batch and seq must be positive");
  JL_REQUIRE(p->seq <= 256, JL_EUNSUPPORTED_SHAPE, "attadapter_fwd: utterances of at most 256 frames (got seq %d): use the composed path", p->seq);
  JL_REQUIRE(p->d >= 64 && (p->d % 64) == 0, JL_EUNSUPPORTED_SHAPE, "attadapter_fwd: d must be a multiple of 64 (got %d)", p->d);
  JL_REQUIRE((p->ldh % 16) == 0 && (p->ldo % 16) == 0, JL_EINVAL, "attadapter_fwd: row strides must be multiples of 16 elements");
  JL_REQUIRE(((reinterpret_cast<uintptr_t>(p->h) | reinterpret_cast<uintptr_t>(p->out)) & 31) == 0, JL_EINVAL, "attadapter_fwd: h / out must be 32-byte aligned");
  JL_REQUIRE(p->cu_seqlens == nullptr || p->total_rows > 0, JL_EINVAL, "attadapter_fwd: packed layout needs total_rows > 0");
  JL_REQUIRE((p->mean == nullptr) == (p->rstd == nullptr), JL_EINVAL, "attadapter_fwd: mean and rstd go together");
  for (const void* q : {p->qkv_out, p->a_out})
    JL_REQUIRE((reinterpret_cast<uintptr_t>(q) & 15) == 0, JL_EINVAL, "attadapter_fwd: qkv_out / a_out must be 16-byte aligned");
  int rc = jl::check_device();
  if (rc != JL_OK) return rc;
  const int64_t rows = p->cu_seqlens ? static_cast<int64_t>(p->total_rows) : static_cast<int64_t>(p->batch) * p->seq;
  CUtensorMap t_h, t_w, t_wkv, t_wo;
  rc = jl::make_tma_map_2d_bf16(&t_h, p->h, p->d, rows, p->ldh, 128);
  if (rc == JL_OK) rc = jl::make_tma_map_2d_bf16(&t_w, p->wqkv_scaled, p->d, 192, p->d, 192);
  if (rc == JL_OK)
    rc = jl::make_tma_map_2d_bf16(&t_wkv, reinterpret_cast<const __nv_bfloat16*>(p->wqkv_scaled) + static_cast<int64_t>(64) * p->d, p->d, 128, p->d, 128);
  if (rc == JL_OK) rc = jl::make_tma_map_2d_bf16(&t_wo, p->wo, 64, p->d, 64, 256);
  if (rc != JL_OK) return rc;
  const size_t smem = sizeof(jl::AaSmem) + 1024;
  static thread_local int configured_dev = -1;
  int dev = 0;
  cudaGetDevice(&dev);
  if (configured_dev != dev) {
    cudaError_t e = cudaFuncSetAttribute(jl::attadapter_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    JL_REQUIRE(e == cudaSuccess, JL_ECUDA, "attadapter_fwd: cannot reserve %zu B of shared memory: %s", smem, cudaGetErrorString(e));
    configured_dev = dev;
  }
  jl::launch(jl::attadapter_fwd_kernel, dim3(jl::ceil_div(p->seq, 128), p->batch), jl::AA_THREADS, smem, reinterpret_cast<cudaStream_t>(stream), t_h, t_w,
             t_wkv, t_wo, *p);
  JL_CHECK_LAUNCH("attadapter_fwd");
  return JL_OK;
}

}  // extern "C"
