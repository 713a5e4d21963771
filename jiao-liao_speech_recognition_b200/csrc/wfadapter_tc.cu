// WFAdapter forward as ONE kernel: LayerNorm + the four factorised low-rank projections + bias / ReLU + residual add.
//
//   z = LN(h);  u = relu((z B_dᵀ) A_dᵀ + c_d);  y = (u B_uᵀ) A_uᵀ + c_u;  out = h + y          (the
//   published analogue is the bottleneck adapter of SP/transformers/models/wav2vec2/modeling_wav2vec2.py:931-953)
//
// HBM-bound by design: h is read once for the projections (+ once more, L2-resident, for the residual), out is written
// once; the factors (≈ 170 KB in bf16) stay in shared memory / L2.  The LayerNorm is folded into the first projection:
//   (LN(h) B_dᵀ)[i, j] = rstd_i · ( (h B_d'ᵀ)[i, j] − μ_i · s_j ) + t_j ,   B_d' = B_d ⊙ γ,  s_j = Σ_k B_d'[j, k],  t_j = Σ_k B_d[j, k] β_k
// so the tensor cores consume the raw h tiles that TMA delivers, while the 128 row-threads accumulate Σx and Σx² of
// their row from the same shared-memory tiles.
//
// One CTA = 128 rows (TMEM lanes).  warp 0: TMA producer; warp 1: tcgen05.mma issuer; warp 2: TMEM allocator;
// warps 4-7: row statistics, the epilogue of every stage (TMEM → registers → bf16 operand tile of the next MMA in the
// K-major 128B-swizzle layout); warps 4-11: the final + bias + residual store (256-bit accesses, the residual segment of
// the next column group is prefetched while the current one is written).
//   G1  acc1[128, r]   = h[128, d] · B_d'ᵀ            12 × 64-wide k-chunks, 4-stage TMA ring
//   G2  acc2[128, b]   = t1[128, 64] · A_dᵀ            rank padded to 64 with zeros (host-side packing)
//   G3  acc3[128, r]   = u[128, b] · B_uᵀ
//   G4  acc4[128, 128] = t2[128, 64] · A_u,chunkᵀ      d / 128 chunks, accumulators double-buffered against the store
#include <cuda.h>
#include <math_constants.h>

#include "common.cuh"
#include "ptx_sm100.cuh"

namespace jl {

constexpr int WF_THREADS = 384;   // 4 service warps + 4 row warps (statistics, stage epilogues) + 4 more for the output stage
constexpr int WF_STAGES = 4;
constexpr uint32_t WF_T128 = 128 * 128;   // bytes of a [128 × 64] bf16 tile

struct __align__(1024) WfSmem {
  uint8_t hs[WF_STAGES][WF_T128];   // h k-chunks (G1 ring); the same 64 KB hold the u operand tiles [128 × 64] × b/64 afterwards
  uint8_t bds[WF_STAGES][64 * 128]; // B_d' k-chunks [r ≤ 64 rows × 64]
  uint8_t t[WF_T128];               // t1 / t2 operand tile [128 × 64], columns ≥ r are zero
  uint8_t ad[256 * 128];            // A_d padded [b ≤ 256 rows × 64]; after G2 the two A_u chunk stages [128 × 64]
  uint8_t bu[4][64 * 128];          // B_u k-chunks [r rows × 64] × b/64
  uint64_t full[WF_STAGES], empty[WF_STAGES];    // G1 ring
  uint64_t ad_free;              // G2 has consumed A_d: its shared memory may take A_u chunks
  uint64_t w_full;               // A_d, B_u resident
  uint64_t acc_full;             // G1 / G2 / G3 accumulators complete (one completion each)
  uint64_t op_full;              // operand tile written by the 4 row warps (one completion per stage)
  uint64_t au_full[2], au_empty[2];
  uint64_t acc4_full[2], acc4_empty[2];
  uint32_t tmem_slot;
};

__device__ __forceinline__ uint32_t wf_idesc(int n) { return ptx::make_idesc_bf16_f32(128, n); }

// D[128 × n] (+)= A[128 × 64] · B[n × 64]ᵀ, both K-major 128B-swizzled tiles; 4 MMAs of K = 16.
__device__ __forceinline__ void wf_mma_k64(uint32_t d_tmem, uint32_t a_addr, uint32_t b_addr, int n, bool accumulate) {
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const uint64_t da = ptx::make_sw128_desc(a_addr + k * 32, 16, 1024);
    const uint64_t db = ptx::make_sw128_desc(b_addr + k * 32, 16, 1024);
    ptx::umma_bf16(d_tmem, da, db, wf_idesc(n), (accumulate || k > 0) ? 1u : 0u);
  }
}

__device__ __forceinline__ void wf_store_chunk(uint8_t* tile, int r, int c, const uint32_t (&pk)[4]) {
  *reinterpret_cast<uint4*>(tile + r * 128 + ((c ^ (r & 7)) << 4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
}

// training: 32 columns (first column col0) of an intermediate's row → global memory, 8 columns per 16-byte store; `width` = r or b
__device__ __forceinline__ void wf_save_cols32(void* base, int64_t row, int width, int col0, const uint32_t (&pk)[16]) {
  __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(base) + row * width + col0;
#pragma unroll
  for (int c = 0; c < 4; ++c)
    if (col0 + c * 8 < width) reinterpret_cast<uint4*>(dst)[c] = make_uint4(pk[4 * c], pk[4 * c + 1], pk[4 * c + 2], pk[4 * c + 3]);
}

__global__ void __launch_bounds__(WF_THREADS, 1)
wfadapter_fwd_kernel(const __grid_constant__ CUtensorMap t_h, const __grid_constant__ CUtensorMap t_bd, const __grid_constant__ CUtensorMap t_ad,
                     const __grid_constant__ CUtensorMap t_bu, const __grid_constant__ CUtensorMap t_au, const jl_wfadapter_fwd_params p) {
  jl::pdl_launch_dependents();
  extern __shared__ uint8_t wf_smem_raw[];
  WfSmem& s = *reinterpret_cast<WfSmem*>(wf_smem_raw + ((1024u - (ptx::smem_u32(wf_smem_raw) & 1023u)) & 1023u));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * 128;
  const int nk = p.d / 64;          // G1 k-chunks
  const int nu = p.b / 64;          // u tiles / B_u chunks
  const int nc4 = p.d / 128;        // G4 column chunks

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&t_h);
    ptx::prefetch_tensormap(&t_bd);
    ptx::prefetch_tensormap(&t_ad);
    ptx::prefetch_tensormap(&t_bu);
    ptx::prefetch_tensormap(&t_au);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < WF_STAGES; ++i) {
      ptx::mbar_init(&s.full[i], 1);
      ptx::mbar_init(&s.empty[i], 5);          // MMA commit + the 4 row warps (statistics pass)
    }
    ptx::mbar_init(&s.ad_free, 1);
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(&s.au_full[i], 1);
      ptx::mbar_init(&s.au_empty[i], 1);
      ptx::mbar_init(&s.acc4_full[i], 1);
      ptx::mbar_init(&s.acc4_empty[i], 8);
    }
    ptx::mbar_init(&s.w_full, 1);
    ptx::mbar_init(&s.acc_full, 1);
    ptx::mbar_init(&s.op_full, 4);
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
  const uint32_t t_acc13 = tmem;              // [128 × r]   columns 0..63
  const uint32_t t_acc2 = tmem + 64;          // [128 × b]   columns 64..319
  const uint32_t t_acc4[2] = {tmem + 64, tmem + 192};   // [128 × 128] each, reuse of the acc2 region
  jl::pdl_wait();

  if (warp == 0) {
    if (lane == 0) {
      // resident factors
      ptx::mbar_expect_tx(&s.w_full, static_cast<uint32_t>(p.b * 128 + nu * p.r * 128));
      ptx::tma_load_2d(s.ad, &t_ad, &s.w_full, 0, 0);
      for (int c = 0; c < nu; ++c) ptx::tma_load_2d(s.bu[c], &t_bu, &s.w_full, c * 64, 0);
      // G1 ring
      for (int kc = 0; kc < nk; ++kc) {
        const int st = kc % WF_STAGES;
        ptx::mbar_wait(&s.empty[st], ((kc / WF_STAGES) & 1) ^ 1u);
        ptx::mbar_expect_tx(&s.full[st], static_cast<uint32_t>(WF_T128 + p.r * 128));
        ptx::tma_load_2d(s.hs[st], &t_h, &s.full[st], kc * 64, m0);
        ptx::tma_load_2d(s.bds[st], &t_bd, &s.full[st], kc * 64, 0);
      }
      // G4: A_u chunks stream through the A_d region once G2 has read it
      ptx::mbar_wait(&s.ad_free, 0);
      for (int c = 0; c < nc4; ++c) {
        const int st = c & 1;
        ptx::mbar_wait(&s.au_empty[st], ((c >> 1) & 1) ^ 1u);
        ptx::mbar_expect_tx(&s.au_full[st], WF_T128);
        ptx::tma_load_2d(s.ad + st * WF_T128, &t_au, &s.au_full[st], 0, c * 128);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t t_addr = ptx::smem_u32(s.t);
      // G1: acc1 = h · B_d'ᵀ
      for (int kc = 0; kc < nk; ++kc) {
        const int st = kc % WF_STAGES;
        ptx::mbar_wait(&s.full[st], (kc / WF_STAGES) & 1);
        ptx::tc_fence_after();
        wf_mma_k64(t_acc13, ptx::smem_u32(s.hs[st]), ptx::smem_u32(s.bds[st]), p.r, kc > 0);
        ptx::umma_commit(&s.empty[st]);
      }
      ptx::umma_commit(&s.acc_full);                                   // completion 0
      // G2: acc2 = t1 · A_dᵀ
      ptx::mbar_wait(&s.op_full, 0);
      ptx::mbar_wait(&s.w_full, 0);
      ptx::tc_fence_after();
      wf_mma_k64(t_acc2, t_addr, ptx::smem_u32(s.ad), p.b, false);
      ptx::umma_commit(&s.ad_free);
      ptx::umma_commit(&s.acc_full);                                   // completion 1
      // G3: acc3 = u · B_uᵀ
      ptx::mbar_wait(&s.op_full, 1);
      ptx::tc_fence_after();
      for (int c = 0; c < nu; ++c) wf_mma_k64(t_acc13, ptx::smem_u32(s.hs[c]), ptx::smem_u32(s.bu[c]), p.r, c > 0);
      ptx::umma_commit(&s.acc_full);                                   // completion 2
      // G4: acc4[c & 1] = t2 · A_u,cᵀ
      ptx::mbar_wait(&s.op_full, 0);                                   // completion 2 of op_full (parity 0 again)
      ptx::tc_fence_after();
      for (int c = 0; c < nc4; ++c) {
        const int st = c & 1;
        ptx::mbar_wait(&s.au_full[st], (c >> 1) & 1);
        ptx::mbar_wait(&s.acc4_empty[st], ((c >> 1) & 1) ^ 1u);
        ptx::tc_fence_after();
        wf_mma_k64(t_acc4[st], t_addr, ptx::smem_u32(s.ad + st * WF_T128), 128, false);
        ptx::umma_commit(&s.au_empty[st]);
        ptx::umma_commit(&s.acc4_full[st]);
      }
    }
  } else if (warp >= 4) {
    const int r = (warp & 3) * 32 + lane;
    const uint32_t lane_off = static_cast<uint32_t>((warp & 3) * 32) << 16;
    const int row = m0 + r;
    const int half = (warp - 4) >> 2;        // output stage: which two of the four 32-column groups of a chunk
    bool valid = row < p.rows;
    if (valid && p.row_lengths != nullptr) {
      const int bb = row / p.rows_per_seq;
      valid = (row - bb * p.rows_per_seq) < __ldg(p.row_lengths + bb);
    }
    if (half == 0) {
    // ---- row statistics from the staged h tiles
    float sx = 0.0f, sxx = 0.0f;
    for (int kc = 0; kc < nk; ++kc) {
      const int st = kc % WF_STAGES;
      ptx::mbar_wait(&s.full[st], (kc / WF_STAGES) & 1);
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const uint4 v = *reinterpret_cast<const uint4*>(s.hs[st] + r * 128 + ((c ^ (r & 7)) << 4));
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
    if (row < p.rows) {
      if (p.mean != nullptr) p.mean[row] = mu;
      if (p.rstd != nullptr) p.rstd[row] = rstd;
    }
    // ---- stage 1: t1 = rstd · (acc1 − μ s) + t  → operand tile (columns ≥ r zero)
    ptx::mbar_wait(&s.acc_full, 0);
    ptx::tc_fence_after();
#pragma unroll
    for (int hh = 0; hh < 2; ++hh) {
      uint32_t pk[16];
      if (hh * 32 < p.r) {
        uint32_t v[32];
        ptx::tmem_ld_32x32(t_acc13 + lane_off + hh * 32, v);
        ptx::tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const int j0 = hh * 32 + 2 * i;
          const float a = (j0 < p.r) ? fmaf(rstd, __uint_as_float(v[2 * i]) - mu * __ldg(p.s + j0), __ldg(p.t + j0)) : 0.0f;
          const float b = (j0 + 1 < p.r) ? fmaf(rstd, __uint_as_float(v[2 * i + 1]) - mu * __ldg(p.s + j0 + 1), __ldg(p.t + j0 + 1)) : 0.0f;
          pk[i] = pack_bf16x2(a, b);
        }
      } else {
#pragma unroll
        for (int i = 0; i < 16; ++i) pk[i] = 0u;
      }
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const uint32_t q4[4] = {pk[4 * c], pk[4 * c + 1], pk[4 * c + 2], pk[4 * c + 3]};
        wf_store_chunk(s.t, r, hh * 4 + c, q4);
      }
      if (p.t1_out != nullptr && row < p.rows && hh * 32 < p.r) wf_save_cols32(p.t1_out, row, p.r, hh * 32, pk);
    }
    ptx::tc_fence_before();
    ptx::fence_proxy_async();
    __syncwarp();
    if (lane == 0) ptx::mbar_arrive(&s.op_full);                        // completion 0
    // ---- stage 2: u = relu(acc2 + c_d) → operand tiles
    ptx::mbar_wait(&s.acc_full, 1);
    ptx::tc_fence_after();
    for (int cc = 0; cc < p.b / 32; ++cc) {
      uint32_t v[32];
      ptx::tmem_ld_32x32(t_acc2 + lane_off + cc * 32, v);
      ptx::tmem_ld_wait();
      uint32_t pk[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const int j0 = cc * 32 + 2 * i;
        const float a = fmaxf(__uint_as_float(v[2 * i]) + __ldg(p.c_d + j0), 0.0f);
        const float b = fmaxf(__uint_as_float(v[2 * i + 1]) + __ldg(p.c_d + j0 + 1), 0.0f);
        pk[i] = pack_bf16x2(a, b);
      }
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const uint32_t q4[4] = {pk[4 * c], pk[4 * c + 1], pk[4 * c + 2], pk[4 * c + 3]};
        wf_store_chunk(s.hs[cc >> 1], r, (cc & 1) * 4 + c, q4);
      }
      if (p.u_out != nullptr && row < p.rows) wf_save_cols32(p.u_out, row, p.b, cc * 32, pk);
    }
    ptx::tc_fence_before();
    ptx::fence_proxy_async();
    __syncwarp();
    if (lane == 0) ptx::mbar_arrive(&s.op_full);                        // completion 1
    // ---- stage 3: t2 = acc3 → operand tile
    ptx::mbar_wait(&s.acc_full, 0);                                     // completion 2 (parity 0 again)
    ptx::tc_fence_after();
#pragma unroll
    for (int hh = 0; hh < 2; ++hh) {
      uint32_t pk[16];
      if (hh * 32 < p.r) {
        uint32_t v[32];
        ptx::tmem_ld_32x32(t_acc13 + lane_off + hh * 32, v);
        ptx::tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const int j0 = hh * 32 + 2 * i;
          pk[i] = pack_bf16x2((j0 < p.r) ? __uint_as_float(v[2 * i]) : 0.0f, (j0 + 1 < p.r) ? __uint_as_float(v[2 * i + 1]) : 0.0f);
        }
      } else {
#pragma unroll
        for (int i = 0; i < 16; ++i) pk[i] = 0u;
      }
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const uint32_t q4[4] = {pk[4 * c], pk[4 * c + 1], pk[4 * c + 2], pk[4 * c + 3]};
        wf_store_chunk(s.t, r, hh * 4 + c, q4);
      }
      if (p.t2_out != nullptr && row < p.rows && hh * 32 < p.r) wf_save_cols32(p.t2_out, row, p.r, hh * 32, pk);
    }
    ptx::tc_fence_before();
    ptx::fence_proxy_async();
    __syncwarp();
    if (lane == 0) ptx::mbar_arrive(&s.op_full);                        // completion 2
    }   // half == 0
    // ---- stage 4 (8 warps): out = h + acc4 + c_u, padded rows := 0, 256-bit accesses; the residual row segment of the next
    //      column group is requested before the current one is finished
    const __nv_bfloat16* hrow = reinterpret_cast<const __nv_bfloat16*>(p.h) + static_cast<int64_t>(row) * p.ldh;
    __nv_bfloat16* orow = reinterpret_cast<__nv_bfloat16*>(p.out) + static_cast<int64_t>(row) * p.ldo;
    const bool in_range = row < p.rows;
    uint32_t hres[2][2][8];
    if (in_range) {
      ld_global_nc_v8(hrow + half * 64, hres[0][0]);
      ld_global_nc_v8(hrow + half * 64 + 16, hres[0][1]);
    }
    const int iters = nc4 * 2;
#pragma unroll 2
    for (int it = 0; it < iters; ++it) {
      const int c = it >> 1, q = half * 2 + (it & 1);
      const int st = c & 1;
      const int col = c * 128 + q * 32;
      const int cur = it & 1;
      if ((it & 1) == 0) {
        ptx::mbar_wait(&s.acc4_full[st], (c >> 1) & 1);
        ptx::tc_fence_after();
      }
      uint32_t v[32];
      ptx::tmem_ld_32x32(t_acc4[st] + lane_off + q * 32, v);
      if (in_range && it + 1 < iters) {
        const int c2 = (it + 1) >> 1, q2 = half * 2 + ((it + 1) & 1);
        ld_global_nc_v8(hrow + c2 * 128 + q2 * 32, hres[cur ^ 1][0]);
        ld_global_nc_v8(hrow + c2 * 128 + q2 * 32 + 16, hres[cur ^ 1][1]);
      }
      ptx::tmem_ld_wait();
      if (in_range) {
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          uint32_t w[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int j = hh * 16 + 2 * i;
            const float2 hr = unpack_bf16x2(hres[cur][hh][i]);
            const float a = valid ? __uint_as_float(v[j]) + __ldg(p.c_u + col + j) + hr.x : 0.0f;
            const float b = valid ? __uint_as_float(v[j + 1]) + __ldg(p.c_u + col + j + 1) + hr.y : 0.0f;
            w[i] = pack_bf16x2(a, b);
          }
          st_global_v8(orow + col + hh * 16, w);
        }
      }
      if (it & 1) {
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&s.acc4_empty[st]);
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

// Operand packing on the device (see jl_wfadapter_pack): one CTA per (factor-set, rank row j) for B_d' = B_d ⊙ γ with its row sums
// s, t; the zero-padded copies of A_d / A_u by a grid-stride loop of the same grid.
__global__ void __launch_bounds__(256) wfadapter_pack_kernel(const jl_wfadapter_pack_params p) {
  jl::pdl_prologue();
  __shared__ float red_s[8], red_t[8];
  const int k = blockIdx.y, j = blockIdx.x;            // j < r
  const __nv_bfloat16* bd = reinterpret_cast<const __nv_bfloat16*>(p.down_B) + (static_cast<int64_t>(k) * p.r + j) * p.d;
  __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(p.bd_scaled) + (static_cast<int64_t>(k) * p.r + j) * p.d;
  float ss = 0.0f, tt = 0.0f;
  for (int c = threadIdx.x; c < p.d; c += blockDim.x) {
    const float w = __bfloat162float(bd[c]);
    const __nv_bfloat16 ws = __float2bfloat16_rn(w * __ldg(p.gamma + c));
    out[c] = ws;
    ss += __bfloat162float(ws);                          // the sum of what the tensor core will actually multiply by
    tt = fmaf(w, __ldg(p.beta + c), tt);
  }
  ss = warp_sum(ss);
  tt = warp_sum(tt);
  if ((threadIdx.x & 31) == 0) { red_s[threadIdx.x >> 5] = ss; red_t[threadIdx.x >> 5] = tt; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float a = 0.0f, b = 0.0f;
    for (int w = 0; w < static_cast<int>(blockDim.x >> 5); ++w) { a += red_s[w]; b += red_t[w]; }     // fixed order
    p.s[k * p.r + j] = a;
    p.t[k * p.r + j] = b;
  }
  // zero-padded [rows, 64] copies of A_d [b, r] and A_u [d, r] of factor set k
  const int tid = j * blockDim.x + threadIdx.x, nthr = p.r * blockDim.x;
  const __nv_bfloat16* ad = reinterpret_cast<const __nv_bfloat16*>(p.down_A) + static_cast<int64_t>(k) * p.b * p.r;
  const __nv_bfloat16* au = reinterpret_cast<const __nv_bfloat16*>(p.up_A) + static_cast<int64_t>(k) * p.d * p.r;
  __nv_bfloat16* adp = reinterpret_cast<__nv_bfloat16*>(p.ad_pad) + static_cast<int64_t>(k) * p.b * 64;
  __nv_bfloat16* aup = reinterpret_cast<__nv_bfloat16*>(p.au_pad) + static_cast<int64_t>(k) * p.d * 64;
  const __nv_bfloat16 zero = __float2bfloat16_rn(0.0f);
  for (int i = tid; i < p.b * 64; i += nthr) adp[i] = ((i & 63) < p.r) ? ad[(i >> 6) * p.r + (i & 63)] : zero;
  for (int i = tid; i < p.d * 64; i += nthr) aup[i] = ((i & 63) < p.r) ? au[(i >> 6) * p.r + (i & 63)] : zero;
}

}  // namespace jl

extern "C" {

int jl_wfadapter_pack(const jl_wfadapter_pack_params* p, void* stream) {
  JL_REQUIRE(p != nullptr, JL_EINVAL, "wfadapter_pack: null params");
  JL_REQUIRE(p->down_B && p->down_A && p->up_A && p->gamma && p->beta && p->bd_scaled && p->s && p->t && p->ad_pad && p->au_pad, JL_EINVAL,
             "wfadapter_pack: null pointer");
  JL_REQUIRE(p->sets >= 1 && p->d > 0 && p->r >= 1 && p->r <= 64 && p->b >= 1, JL_EINVAL, "wfadapter_pack: bad dims (1 <= r <= 64)");
  int rc = jl::check_device();
  if (rc != JL_OK) return rc;
  jl::launch(jl::wfadapter_pack_kernel, dim3(p->r, p->sets), 256, 0, reinterpret_cast<cudaStream_t>(stream), *p);
  JL_CHECK_LAUNCH("wfadapter_pack");
  return JL_OK;
}

int jl_wfadapter_fwd(const jl_wfadapter_fwd_params* p, void* stream) {
  JL_REQUIRE(p != nullptr, JL_EINVAL, "wfadapter_fwd: null params");
  JL_REQUIRE(p->h && p->out && p->bd_scaled && p->s && p->t && p->ad_pad && p->c_d && p->bu && p->au_pad && p->c_u, JL_EINVAL,
             "wfadapter_fwd: null pointer");
  JL_REQUIRE(p->rows > 0, JL_EINVAL, "wfadapter_fwd: rows must be positive");
  JL_REQUIRE(p->d >= 128 && (p->d % 128) == 0, JL_EUNSUPPORTED_SHAPE, "wfadapter_fwd: d must be a multiple of 128 (got %d)", p->d);
  JL_REQUIRE(p->r >= 16 && p->r <= 64 && (p->r % 16) == 0, JL_EUNSUPPORTED_SHAPE, "wfadapter_fwd: rank must be 16, 32, 48 or 64 (got %d)", p->r);
  JL_REQUIRE(p->b >= 64 && p->b <= 256 && (p->b % 64) == 0, JL_EUNSUPPORTED_SHAPE, "wfadapter_fwd: bottleneck must be 64, 128, 192 or 256 (got %d)", p->b);
  JL_REQUIRE((p->ldh % 16) == 0 && (p->ldo % 16) == 0, JL_EINVAL, "wfadapter_fwd: row strides must be multiples of 16 elements");
  JL_REQUIRE(((reinterpret_cast<uintptr_t>(p->h) | reinterpret_cast<uintptr_t>(p->out)) & 31) == 0, JL_EINVAL, "wfadapter_fwd: h / out must be 32-byte aligned");
  if (p->row_lengths) JL_REQUIRE(p->rows_per_seq > 0, JL_EINVAL, "wfadapter_fwd: row_lengths needs rows_per_seq > 0");
  int rc = jl::check_device();
  if (rc != JL_OK) return rc;
  CUtensorMap t_h, t_bd, t_ad, t_bu, t_au;
  rc = jl::make_tma_map_2d_bf16(&t_h, p->h, p->d, p->rows, p->ldh, 128);
  if (rc == JL_OK) rc = jl::make_tma_map_2d_bf16(&t_bd, p->bd_scaled, p->d, p->r, p->d, p->r);
  if (rc == JL_OK) rc = jl::make_tma_map_2d_bf16(&t_ad, p->ad_pad, 64, p->b, 64, p->b);
  if (rc == JL_OK) rc = jl::make_tma_map_2d_bf16(&t_bu, p->bu, p->b, p->r, p->b, p->r);
  if (rc == JL_OK) rc = jl::make_tma_map_2d_bf16(&t_au, p->au_pad, 64, p->d, 64, 128);
  if (rc != JL_OK) return rc;
  const size_t smem = sizeof(jl::WfSmem) + 1024;
  static thread_local int configured_dev = -1;
  int dev = 0;
  cudaGetDevice(&dev);
  if (configured_dev != dev) {
    cudaError_t e = cudaFuncSetAttribute(jl::wfadapter_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    JL_REQUIRE(e == cudaSuccess, JL_ECUDA, "wfadapter_fwd: cannot reserve %zu B of shared memory: %s", smem, cudaGetErrorString(e));
    configured_dev = dev;
  }
  jl::launch(jl::wfadapter_fwd_kernel, jl::ceil_div(p->rows, 128), jl::WF_THREADS, smem, reinterpret_cast<cudaStream_t>(stream), t_h, t_bd, t_ad,
             t_bu, t_au, *p);
  JL_CHECK_LAUNCH("wfadapter_fwd");
  return JL_OK;
}

}  // extern "C"
