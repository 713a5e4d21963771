// Backward through "LayerNorm → narrow projection" in ONE kernel (the tail of the AttAdapter backward, a7):
//
//   forward:   y = LN(h) Wᵀ + b          W [n, d], n <= 192 (the AttAdapter's q | k | v projection: n = 192)
//   backward:  dz = dy · W               [rows, d]    (gradient of the LayerNorm output)
//              dh = LN'(dz) + dres       dh_ij = rstd_i (dz_ij γ_j − c1_i − x̂_ij c2_i) + dres_ij,  x̂ = (h − μ) rstd
//              c1_i = mean_j(dz_ij γ_j),  c2_i = mean_j(dz_ij γ_j x̂_ij)
//
// It replaces jl_gemm_bf16 (dy · W, 10 µs at 8000 × 768 × 192) + jl_layernorm_bwd (10.6 µs), and the HBM round trip of dz between
// them.  The two row means do NOT need the full dz row: with the LayerNorm-fold vectors of the forward kernel (jl_lnfold_pack:
// W' = bf16(W ⊙ γ), s_k = Σ_j W'_kj, t_k = Σ_j W_kj β_j + b_k)
//   c1_i = (1/d) Σ_k dy_ik s_k                       because Σ_j dz_ij γ_j = Σ_k dy_ik Σ_j W_kj γ_j
//   c2_i = (1/d) Σ_k dy_ik (y_ik − t_k)              because Σ_j W'_kj x̂_ij = y_ik − t_k  (the forward projection itself)
// are dot products over the n <= 192 columns of dy and the SAVED y — a prologue — so the product runs once, chunk by chunk, and
// every chunk's epilogue emits final dh.
//
// CTA = 128 rows, 384 threads: warp 0 TMA producer, warp 1 tcgen05.mma issuer, warp 2 TMEM allocator, warps 4-11 epilogue.
//   A = dy tile [128 x n] resident (K-major, n / 64 swizzled tiles); B = W[:, 64-column chunk] as it lies in memory ([n, 64] box,
//   read MN-major); per chunk the h and dres tiles arrive by TMA in staging buffers (3 deep), each thread folds its row of the
//   accumulator into the dres tile in place (and dz into the h tile when the caller wants it for dγ / dβ), and the tiles leave
//   with coalesced 16-byte stores.
#include <cuda.h>
#include <cstddef>

#include "common.cuh"
#include "ptx_sm100.cuh"

namespace jl {

constexpr int LP_THREADS = 384;
constexpr int LP_BUFS = 3;                          // B chunk / staging buffers
constexpr int LP_ACC = 8;                           // TMEM accumulators of 64 columns
constexpr uint32_t LP_T128 = 128 * 128;             // bytes of a [128 x 64] bf16 tile
constexpr uint32_t LP_BCH = 192 * 128;              // bytes of a [192 x 64] bf16 box
constexpr int LP_MAX_D = 1024;
extern int g_lnproj_split;

struct __align__(1024) LpSmem {
  uint8_t a[3][LP_T128];                // dy tiles (k = 64-column groups of dy)
  uint8_t b[LP_BUFS][LP_BCH];           // W[0:n, chunk]: n rows of 64 contiguous columns
  uint8_t hs[LP_BUFS][LP_T128];         // h tile of the chunk (first: the saved y tiles for the prologue); dz in place
  uint8_t rs[LP_BUFS][LP_T128];         // dres tile of the chunk; dh in place
  float gamma[LP_MAX_D];
  float fs[192], ftb[192];
  float red1[2][128], red2[2][128];
  uint64_t a_full, y_full, y_done;
  uint64_t b_full[LP_BUFS], b_empty[LP_BUFS], h_full[LP_BUFS], r_full[LP_BUFS], stg_empty[LP_BUFS];
  uint64_t acc_full[LP_ACC], acc_empty[LP_ACC];
  uint32_t tmem_slot;
};

static_assert(offsetof(LpSmem, ftb) == offsetof(LpSmem, fs) + 192 * 4 && offsetof(LpSmem, red1) == offsetof(LpSmem, ftb) + 192 * 4 &&
              offsetof(LpSmem, red2) == offsetof(LpSmem, red1) + 2 * 128 * 4, "the column-partial scratch aliases fs | ftb | red1 | red2");
__device__ __forceinline__ uint32_t lp_chunk_off(int r, int c) { return static_cast<uint32_t>(r * 128 + ((c ^ (r & 7)) << 4)); }
__device__ __forceinline__ uint4 lp_lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}
// coalesced copy of one [128 x 64] bf16 tile to global rows (256 threads; thread t: 16-byte chunk t & 7 of rows (t >> 3) + 32 j)
__device__ __forceinline__ void lp_tile_to_global(uint32_t tile, __nv_bfloat16* dst, int64_t ld, int rows_valid, int t) {
  const int ch = t & 7, r0 = t >> 3;
  const uint32_t off = static_cast<uint32_t>(r0 * 128 + ((ch ^ (r0 & 7)) << 4));
  uint4 v[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) v[j] = lp_lds128(tile + off + j * 4096);
#pragma unroll
  for (int j = 0; j < 4; ++j)
    if (r0 + 32 * j < rows_valid) *reinterpret_cast<uint4*>(dst + static_cast<int64_t>(r0 + 32 * j) * ld + ch * 8) = v[j];
}

// Column sums over the 32 rows a warp holds: in: a[j] = this lane's (row's) value in column j; out (return value): the sum over the
// warp's 32 rows of column `lane`.  Recursive halving — 31 shuffles instead of the 160 of 32 butterfly reductions; fixed order.
__device__ __forceinline__ float lp_warp_colsum(float (&a)[32], int lane) {
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
    const bool upper = (lane & off) != 0;
#pragma unroll
    for (int i = 0; i < off; ++i) {
      const float send = upper ? a[i] : a[i + off];
      const float keep = upper ? a[i + off] : a[i];
      a[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
  }
  return a[0];
}

__global__ void __launch_bounds__(LP_THREADS, 1)
lnproj_bwd_kernel(const __grid_constant__ CUtensorMap t_dy, const __grid_constant__ CUtensorMap t_y, const __grid_constant__ CUtensorMap t_w,
                  const __grid_constant__ CUtensorMap t_h, const __grid_constant__ CUtensorMap t_r, const jl_lnproj_bwd_params p) {
  extern __shared__ uint8_t lp_smem_raw[];
  LpSmem& s = *reinterpret_cast<LpSmem*>(lp_smem_raw + ((1024u - (ptx::smem_u32(lp_smem_raw) & 1023u)) & 1023u));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // row tile → rows [row0, row_end) and the factor set of its run (several row ranges, each with its own W / s / t block: the dialect
  // runs of a WFAdapter); without runs: tile i = rows 128 i … of the whole matrix, set 0
  int row0 = blockIdx.x * 128, row_end = p.rows, set = 0;
  if (p.num_runs > 0) {
    int tile = blockIdx.x, i = 0;
    for (; i < p.num_runs - 1; ++i) {
      const int nt = (p.run_start[i + 1] - p.run_start[i] + 127) / 128;
      if (tile < nt) break;
      tile -= nt;
    }
    row0 = p.run_start[i] + tile * 128;
    row_end = p.run_start[i + 1];
    set = p.run_set[i];
  }
  const int nkt = (p.n + 63) / 64;                  // 64-wide k tiles (columns of dy; a partial last tile is zero-filled by TMA)
  // the 64-column output chunks of the row tile may be split between gridDim.y CTAs (each repeats the small prologue): an SM takes
  // in ≈ 57 B/clk and stores ≈ 28 B/clk (see the host function for when)
  const int nc_all = p.d / 64;
  const int c_begin = static_cast<int>((static_cast<int64_t>(nc_all) * blockIdx.y) / gridDim.y);
  const int c_end = static_cast<int>((static_cast<int64_t>(nc_all) * (blockIdx.y + 1)) / gridDim.y);
  const int nc = c_end - c_begin;                   // chunks of this CTA: global chunk index c_begin + c
  const int rows_valid = min(128, row_end - row0);

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&t_dy);
    ptx::prefetch_tensormap(&t_y);
    ptx::prefetch_tensormap(&t_w);
    ptx::prefetch_tensormap(&t_h);
    ptx::prefetch_tensormap(&t_r);
  }
  if (warp == 1 && lane == 0) {
    ptx::mbar_init(&s.a_full, 1);
    ptx::mbar_init(&s.y_full, 1);
    ptx::mbar_init(&s.y_done, 8);
    for (int i = 0; i < LP_BUFS; ++i) {
      ptx::mbar_init(&s.b_full[i], 1);
      ptx::mbar_init(&s.b_empty[i], 1);
      ptx::mbar_init(&s.h_full[i], 1);
      ptx::mbar_init(&s.r_full[i], 1);
      ptx::mbar_init(&s.stg_empty[i], 1);
    }
    for (int i = 0; i < LP_ACC; ++i) {
      ptx::mbar_init(&s.acc_full[i], 1);
      ptx::mbar_init(&s.acc_empty[i], 8);
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
  jl::pdl_prologue();

  if (warp == 0) {
    if (lane == 0) {
      ptx::mbar_expect_tx(&s.a_full, nkt * LP_T128);
      for (int kt = 0; kt < nkt; ++kt) ptx::tma_load_2d(s.a[kt], &t_dy, &s.a_full, kt * 64, row0);
      ptx::mbar_expect_tx(&s.y_full, nkt * LP_T128);
      for (int kt = 0; kt < nkt; ++kt) ptx::tma_load_2d(s.hs[kt], &t_y, &s.y_full, kt * 64, row0);
      for (int c = 0; c < min(nc, LP_BUFS); ++c) {
        ptx::mbar_expect_tx(&s.b_full[c], nkt * 64 * 128);
        ptx::tma_load_2d(s.b[c], &t_w, &s.b_full[c], (c_begin + c) * 64, set * p.n);
        ptx::mbar_expect_tx(&s.r_full[c], LP_T128);
        ptx::tma_load_2d(s.rs[c], &t_r, &s.r_full[c], (c_begin + c) * 64, row0);
      }
      ptx::mbar_wait(&s.y_done, 0);                 // the prologue has read the saved y tiles: their buffers take h tiles now
      for (int c = 0; c < min(nc, LP_BUFS); ++c) {
        ptx::mbar_expect_tx(&s.h_full[c], LP_T128);
        ptx::tma_load_2d(s.hs[c], &t_h, &s.h_full[c], (c_begin + c) * 64, row0);
      }
      for (int c = LP_BUFS; c < nc; ++c) {
        const int bi = c % LP_BUFS;
        const uint32_t par = ((c / LP_BUFS) - 1) & 1;
        ptx::mbar_wait(&s.b_empty[bi], par);
        ptx::mbar_expect_tx(&s.b_full[bi], nkt * 64 * 128);
        ptx::tma_load_2d(s.b[bi], &t_w, &s.b_full[bi], (c_begin + c) * 64, set * p.n);
        ptx::mbar_wait(&s.stg_empty[bi], par);
        ptx::mbar_expect_tx(&s.h_full[bi], LP_T128);
        ptx::tma_load_2d(s.hs[bi], &t_h, &s.h_full[bi], (c_begin + c) * 64, row0);
        ptx::mbar_expect_tx(&s.r_full[bi], LP_T128);
        ptx::tma_load_2d(s.rs[bi], &t_r, &s.r_full[bi], (c_begin + c) * 64, row0);
        if (c == nc - 1) jl::pdl_trigger_late();
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      ptx::mbar_wait(&s.a_full, 0);
      const uint32_t idesc = ptx::make_idesc_bf16_f32(128, 64) | (1u << 16);      // B operand MN-major
      for (int c = 0; c < nc; ++c) {
        const int bi = c % LP_BUFS, ai = c % LP_ACC;
        ptx::mbar_wait(&s.b_full[bi], (c / LP_BUFS) & 1);
        ptx::mbar_wait(&s.acc_empty[ai], ((c / LP_ACC) & 1) ^ 1u);
        ptx::tc_fence_after();
        const uint32_t bb = ptx::smem_u32(s.b[bi]);
        for (int kt = 0; kt < nkt; ++kt) {
          const uint32_t ab = ptx::smem_u32(s.a[kt]);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            ptx::umma_bf16(tmem + ai * 64, ptx::make_sw128_desc(ab + k * 32, 16, 1024), ptx::make_sw128_desc(bb + kt * 8192 + k * 2048, 8192, 1024), idesc,
                           (kt > 0 || k > 0) ? 1u : 0u);
        }
        ptx::umma_commit(&s.b_empty[bi]);
        ptx::umma_commit(&s.acc_full[ai]);
      }
    }
  } else if (warp >= 4) {
    const int quad = warp & 3;
    const int grp = (warp - 4) >> 2;
    const int r = quad * 32 + lane;                      // TMEM lane = row of the tile
    const int et = threadIdx.x - 128;
    const uint32_t lane_off = static_cast<uint32_t>(quad * 32) << 16;
    const int row = row0 + r;
    for (int i = et; i < p.d; i += 256) s.gamma[i] = __ldg(p.gamma + i);
    if (et < 192) { s.fs[et] = et < p.n ? __ldg(p.s + set * p.n + et) : 0.0f; s.ftb[et] = et < p.n ? __ldg(p.tb + set * p.n + et) : 0.0f; }
    const float mu = row < row_end ? __ldg(p.mean + row) : 0.0f;
    const float rstd = row < row_end ? __ldg(p.rstd + row) : 0.0f;
    asm volatile("bar.sync 1, 256;" ::: "memory");
    // ---- prologue: c1, c2 from dy and the saved y (this thread: every other 16-byte chunk group of its row)
    ptx::mbar_wait(&s.a_full, 0);
    ptx::mbar_wait(&s.y_full, 0);
    float p1 = 0.0f, p2 = 0.0f;
    for (int kt = 0; kt < nkt; ++kt) {
      const uint32_t at = ptx::smem_u32(s.a[kt]), yt = ptx::smem_u32(s.hs[kt]);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int ch = grp * 4 + j;
        const uint4 dv = lp_lds128(at + lp_chunk_off(r, ch)), yv = lp_lds128(yt + lp_chunk_off(r, ch));
        const uint32_t dw[4] = {dv.x, dv.y, dv.z, dv.w}, yw[4] = {yv.x, yv.y, yv.z, yv.w};
        const float* fs = s.fs + kt * 64 + ch * 8;
        const float* ft = s.ftb + kt * 64 + ch * 8;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float2 d2 = unpack_bf16x2(dw[q]), y2 = unpack_bf16x2(yw[q]);
          p1 = fmaf(d2.x, fs[2 * q], p1);
          p1 = fmaf(d2.y, fs[2 * q + 1], p1);
          p2 = fmaf(d2.x, y2.x - ft[2 * q], p2);
          p2 = fmaf(d2.y, y2.y - ft[2 * q + 1], p2);
        }
      }
    }
    if (p.dy_scaled != nullptr) {
      // operands of the projection's weight gradient without LN(h):  dW = ((dy ⊙ rstd)ᵀ h − v 1ᵀ) ⊙ γ + cs βᵀ  with  v = (dy ⊙ rstd)ᵀ μ,
      // cs = Σ_rows dy.  Here: dy ⊙ rstd (bf16) to HBM and the column sums v, cs of this warp's 32 rows (finished by jl_lnproj_wgrad).
      for (int kt = 0; kt < nkt; ++kt) {
        const uint32_t at = ptx::smem_u32(s.a[kt]);
        float dv[32], dm[32];
        uint32_t pk[16];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const uint4 v4 = lp_lds128(at + lp_chunk_off(r, grp * 4 + j));
          const uint32_t w4[4] = {v4.x, v4.y, v4.z, v4.w};
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float2 d2 = unpack_bf16x2(w4[q]);
            dv[j * 8 + 2 * q] = d2.x;
            dv[j * 8 + 2 * q + 1] = d2.y;
            pk[j * 4 + q] = pack_bf16x2(d2.x * rstd, d2.y * rstd);
            const float2 s2 = unpack_bf16x2(pk[j * 4 + q]);          // the rounded operand the GEMM will see
            dm[j * 8 + 2 * q] = s2.x * mu;
            dm[j * 8 + 2 * q + 1] = s2.y * mu;
          }
        }
        if (row < row_end) {
          uint4* dst = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.dy_scaled) + static_cast<int64_t>(row) * p.lddys + kt * 64 + grp * 32);
#pragma unroll
          for (int j = 0; j < 4; ++j) dst[j] = make_uint4(pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
        }
        const float cs = lp_warp_colsum(dv, lane), vm = lp_warp_colsum(dm, lane);
        const int64_t prow = static_cast<int64_t>(blockIdx.x) * 4 + quad;
        const int64_t nrows = static_cast<int64_t>(gridDim.x) * 4;
        if (blockIdx.y == 0) {
          p.wgrad_partial[prow * p.n + kt * 64 + grp * 32 + lane] = cs;
          p.wgrad_partial[(nrows + prow) * p.n + kt * 64 + grp * 32 + lane] = vm;
        }
      }
    }
    s.red1[grp][r] = p1;
    s.red2[grp][r] = p2;
    __syncwarp();
    if (lane == 0) ptx::mbar_arrive(&s.y_done);
    asm volatile("bar.sync 1, 256;" ::: "memory");
    const float inv_d = 1.0f / static_cast<float>(p.d);
    const float c1 = (s.red1[0][r] + s.red1[1][r]) * inv_d;
    const float c2 = (s.red2[0][r] + s.red2[1][r]) * inv_d;
    const float nmr = -mu * rstd;
    const bool want_cols = p.col_partial != nullptr;
    float* colp = s.fs;                 // [3][4][64] floats over fs | ftb | red1 | red2 (3584 B): all dead after the prologue
    __nv_bfloat16* dx_tile = reinterpret_cast<__nv_bfloat16*>(p.dx) + static_cast<int64_t>(row0) * p.lddx;
    __nv_bfloat16* dz_tile = p.dz ? reinterpret_cast<__nv_bfloat16*>(p.dz) + static_cast<int64_t>(row0) * p.lddz : nullptr;
    // ---- chunks of 64 output columns: this thread = row r, columns grp * 32 … + 32
    for (int c = 0; c < nc; ++c) {
      const int bi = c % LP_BUFS, ai = c % LP_ACC;
      ptx::mbar_wait(&s.acc_full[ai], (c / LP_ACC) & 1);
      ptx::mbar_wait(&s.h_full[bi], (c / LP_BUFS) & 1);
      ptx::mbar_wait(&s.r_full[bi], (c / LP_BUFS) & 1);
      ptx::tc_fence_after();
      uint32_t v[32];
      float zx[32], ro[32];
      ptx::tmem_ld_32x32(tmem + ai * 64 + lane_off + grp * 32, v);
      ptx::tmem_ld_wait();
      const uint32_t ht = ptx::smem_u32(s.hs[bi]), rt = ptx::smem_u32(s.rs[bi]);
      const float* gam = s.gamma + (c_begin + c) * 64 + grp * 32;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const uint32_t off = lp_chunk_off(r, grp * 4 + j);
        const uint4 hv = lp_lds128(ht + off), rv = lp_lds128(rt + off);
        const uint32_t hw[4] = {hv.x, hv.y, hv.z, hv.w}, rw[4] = {rv.x, rv.y, rv.z, rv.w};
        const float4 g0 = *reinterpret_cast<const float4*>(gam + j * 8), g1 = *reinterpret_cast<const float4*>(gam + j * 8 + 4);
        const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
        uint32_t ox[4], oz[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float2 h2 = unpack_bf16x2(hw[q]), r2 = unpack_bf16x2(rw[q]);
          const float z0 = __uint_as_float(v[j * 8 + 2 * q]), z1 = __uint_as_float(v[j * 8 + 2 * q + 1]);
          const float x0 = fmaf(h2.x, rstd, nmr), x1 = fmaf(h2.y, rstd, nmr);
          if (want_cols) {
            zx[j * 8 + 2 * q] = z0 * x0;
            zx[j * 8 + 2 * q + 1] = z1 * x1;
            ro[j * 8 + 2 * q] = r2.x;
            ro[j * 8 + 2 * q + 1] = r2.y;
          }
          const float t0 = fmaf(-x0, c2, fmaf(z0, gg[2 * q], -c1)), t1 = fmaf(-x1, c2, fmaf(z1, gg[2 * q + 1], -c1));
          ox[q] = pack_bf16x2(fmaf(rstd, t0, r2.x), fmaf(rstd, t1, r2.y));
          oz[q] = pack_bf16x2(z0, z1);
        }
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(rt + off), "r"(ox[0]), "r"(ox[1]), "r"(ox[2]), "r"(ox[3]) : "memory");
        if (dz_tile != nullptr)
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(ht + off), "r"(oz[0]), "r"(oz[1]), "r"(oz[2]), "r"(oz[3]) : "memory");
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&s.acc_empty[ai]);
      if (want_cols) {
        // column sums of this warp's 32 rows: Σ dz (→ dβ), Σ dz x̂ (→ dγ), Σ dres (→ the bias gradient of the layer that produced dres)
        // (rows past the end of a run hold the next run's data — zero-filled by TMA only past the end of the matrix)
        const bool live = r < rows_valid;
        float zz[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          zz[i] = live ? __uint_as_float(v[i]) : 0.0f;
          zx[i] = live ? zx[i] : 0.0f;
          ro[i] = live ? ro[i] : 0.0f;
        }
        const float sb = lp_warp_colsum(zz, lane), sg = lp_warp_colsum(zx, lane), so = lp_warp_colsum(ro, lane);
        colp[(0 * 4 + quad) * 64 + grp * 32 + lane] = sb;
        colp[(1 * 4 + quad) * 64 + grp * 32 + lane] = sg;
        colp[(2 * 4 + quad) * 64 + grp * 32 + lane] = so;
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (want_cols && et < 192) {
        const int qn = et >> 6, col = et & 63;
        const float* cp = colp + qn * 256 + col;
        const float tot = ((cp[0] + cp[64]) + cp[128]) + cp[192];          // the four 32-row groups, fixed order
        p.col_partial[(static_cast<int64_t>(qn) * gridDim.x + blockIdx.x) * p.d + (c_begin + c) * 64 + col] = tot;
      }
      lp_tile_to_global(rt, dx_tile + (c_begin + c) * 64, p.lddx, rows_valid, et);
      if (dz_tile != nullptr) lp_tile_to_global(ht, dz_tile + (c_begin + c) * 64, p.lddz, rows_valid, et);
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (et == 0) {
        ptx::fence_proxy_async();
        ptx::mbar_arrive(&s.stg_empty[bi]);
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

// col_partial [3][nblk][d] → dbeta = Σ_blk q0, dgamma = Σ_blk q1, dbias = Σ_blk q2 (any output may be NULL); fixed order.
// CTA = 32 columns × 8 row groups (each sums a strided subset of the partial rows, coalesced 128 B reads).
__global__ void __launch_bounds__(256) lnproj_bwd_reduce_kernel(const float* __restrict__ partial, int nblk, int d, float* __restrict__ dgamma,
                                                                float* __restrict__ dbeta, float* __restrict__ dbias, int accumulate, int tile_offset,
                                                                int total_tiles) {
  jl::pdl_prologue();
  __shared__ float sm[3][8][33];
  const int lane = threadIdx.x & 31, grp = threadIdx.x >> 5;
  const int i = blockIdx.x * 32 + lane;
  float acc[3] = {0.0f, 0.0f, 0.0f};
  if (i < d) {
    for (int k = grp; k < nblk; k += 8) {
#pragma unroll
      for (int q = 0; q < 3; ++q) acc[q] += partial[(static_cast<int64_t>(q) * total_tiles + tile_offset + k) * d + i];
    }
  }
#pragma unroll
  for (int q = 0; q < 3; ++q) sm[q][grp][lane] = acc[q];
  __syncthreads();
  if (grp == 0 && i < d) {
#pragma unroll
    for (int q = 0; q < 3; ++q) {
      float t = acc[q];
#pragma unroll
      for (int w = 1; w < 8; ++w) t += sm[q][w][lane];
      float* out = q == 0 ? dbeta : (q == 1 ? dgamma : dbias);
      if (out != nullptr) out[i] = (accumulate && q < 2) ? out[i] + t : t;      // dgamma / dbeta may collect several row ranges
    }
  }
}

// dW[k, :] = (M0[k, :] − v_k) ⊙ γ + cs_k β,  db_k = cs_k, with cs = Σ partial[0], v = Σ partial[1] over the `prow` partial rows
// (fixed order).  One CTA per projection row k; M0 is overwritten in place.
__global__ void __launch_bounds__(256) lnproj_wgrad_kernel(float* __restrict__ m0, int64_t ldm, const float* __restrict__ partial, int prow, int n, int d,
                                                           const float* __restrict__ gamma, const float* __restrict__ beta, float* __restrict__ dbias) {
  jl::pdl_prologue();
  __shared__ float red[2][256];
  const int k = blockIdx.x, t = threadIdx.x;
  float cs = 0.0f, vm = 0.0f;
  for (int i = t; i < prow; i += 256) {
    cs += partial[static_cast<int64_t>(i) * n + k];
    vm += partial[(static_cast<int64_t>(prow) + i) * n + k];
  }
  red[0][t] = cs;
  red[1][t] = vm;
  __syncthreads();
  for (int off = 128; off > 0; off >>= 1) {
    if (t < off) {
      red[0][t] += red[0][t + off];
      red[1][t] += red[1][t + off];
    }
    __syncthreads();
  }
  cs = red[0][0];
  vm = red[1][0];
  float* row = m0 + static_cast<int64_t>(k) * ldm;
  for (int j = t; j < d; j += 256) row[j] = fmaf(row[j] - vm, __ldg(gamma + j), cs * __ldg(beta + j));
  if (t == 0 && dbias != nullptr) dbias[k] = cs;
}

// The operands of jl_lnproj_wgrad as a kernel of their own (for the weight-gradient branch): dy ⊙ rstd (bf16) and, per 32 rows, the
// column sums of dy and of (dy ⊙ rstd) μ.  CTA = 128 rows, 256 threads: thread = (row, half of every 64-column group).
__global__ void __launch_bounds__(256) lnproj_wgrad_prep_kernel(const __nv_bfloat16* __restrict__ dy, int64_t lddy, const float* __restrict__ mean,
                                                                const float* __restrict__ rstd, int rows, int n, __nv_bfloat16* __restrict__ dys, int64_t lddys,
                                                                float* __restrict__ partial) {
  jl::pdl_prologue();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int quad = warp & 3, grp = warp >> 2;
  const int row = blockIdx.x * 128 + quad * 32 + lane;
  const bool ok = row < rows;
  const float mu = ok ? __ldg(mean + row) : 0.0f, rs = ok ? __ldg(rstd + row) : 0.0f;
  const int64_t prow = static_cast<int64_t>(blockIdx.x) * 4 + quad, nrows = static_cast<int64_t>(gridDim.x) * 4;
  for (int c0 = grp * 32; c0 < n; c0 += 64) {
    float dv[32], dm[32];
    uint32_t pk[16];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const bool in = ok && c0 + j * 8 < n;
      const uint4 v4 = in ? __ldg(reinterpret_cast<const uint4*>(dy + static_cast<int64_t>(row) * lddy + c0) + j) : make_uint4(0u, 0u, 0u, 0u);
      const uint32_t w4[4] = {v4.x, v4.y, v4.z, v4.w};
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float2 d2 = unpack_bf16x2(w4[q]);
        dv[j * 8 + 2 * q] = d2.x;
        dv[j * 8 + 2 * q + 1] = d2.y;
        pk[j * 4 + q] = pack_bf16x2(d2.x * rs, d2.y * rs);
        const float2 s2 = unpack_bf16x2(pk[j * 4 + q]);
        dm[j * 8 + 2 * q] = s2.x * mu;
        dm[j * 8 + 2 * q + 1] = s2.y * mu;
      }
      if (in) reinterpret_cast<uint4*>(dys + static_cast<int64_t>(row) * lddys + c0)[j] = make_uint4(pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
    }
    const float cs = lp_warp_colsum(dv, lane), vm = lp_warp_colsum(dm, lane);
    if (c0 + lane < n) {
      partial[prow * n + c0 + lane] = cs;
      partial[(nrows + prow) * n + c0 + lane] = vm;
    }
  }
}

int g_lnproj_split = 0;     // 0 = automatic; tuning aid (jl_debug_set_lnproj_split)

}  // namespace jl

extern "C" {

void jl_debug_set_lnproj_split(int split) { jl::g_lnproj_split = split; }

int jl_lnproj_wgrad_prep(const void* dy, int64_t lddy, const float* mean, const float* rstd, int32_t rows, int32_t n, void* dy_scaled, int64_t lddys,
                         float* wgrad_partial, void* stream) {
  JL_REQUIRE(dy && mean && rstd && dy_scaled && wgrad_partial && rows > 0 && n > 0 && (n % 8) == 0, JL_EINVAL, "lnproj_wgrad_prep: bad arguments");
  JL_REQUIRE((lddy % 8) == 0 && (lddys % 8) == 0 && ((reinterpret_cast<uintptr_t>(dy) | reinterpret_cast<uintptr_t>(dy_scaled)) & 15) == 0, JL_EINVAL,
             "lnproj_wgrad_prep: 16-byte aligned pointers and row strides that are multiples of 8");
  int rc = jl::check_device();
  if (rc != JL_OK) return rc;
  jl::launch(jl::lnproj_wgrad_prep_kernel, jl::ceil_div(rows, 128), 256, 0, reinterpret_cast<cudaStream_t>(stream), reinterpret_cast<const __nv_bfloat16*>(dy), lddy,
             mean, rstd, static_cast<int>(rows), static_cast<int>(n), reinterpret_cast<__nv_bfloat16*>(dy_scaled), lddys, wgrad_partial);
  JL_CHECK_LAUNCH("lnproj_wgrad_prep");
  return JL_OK;
}

int jl_lnproj_wgrad(float* m0, int64_t ldm, const float* wgrad_partial, int32_t partial_rows, int32_t n, int32_t d, const float* gamma, const float* beta,
                    float* dbias, void* stream) {
  JL_REQUIRE(m0 && wgrad_partial && gamma && beta && partial_rows > 0 && n > 0 && d > 0 && ldm >= d, JL_EINVAL, "lnproj_wgrad: bad arguments");
  int rc = jl::check_device();
  if (rc != JL_OK) return rc;
  jl::launch(jl::lnproj_wgrad_kernel, n, 256, 0, reinterpret_cast<cudaStream_t>(stream), m0, ldm, wgrad_partial, partial_rows, n, d, gamma, beta, dbias);
  JL_CHECK_LAUNCH("lnproj_wgrad");
  return JL_OK;
}

int jl_lnproj_bwd_reduce(const float* col_partial, int32_t row_tiles, int32_t d, float* dgamma, float* dbeta, float* dbias, int32_t accumulate,
                         int32_t tile_offset, int32_t total_tiles, void* stream) {
  JL_REQUIRE(col_partial != nullptr && row_tiles > 0 && d > 0, JL_EINVAL, "lnproj_bwd_reduce: bad arguments");
  if (total_tiles <= 0) total_tiles = row_tiles;
  JL_REQUIRE(tile_offset >= 0 && tile_offset + row_tiles <= total_tiles, JL_EINVAL, "lnproj_bwd_reduce: tile range out of bounds");
  int rc = jl::check_device();
  if (rc != JL_OK) return rc;
  jl::launch(jl::lnproj_bwd_reduce_kernel, jl::ceil_div(d, 32), 256, 0, reinterpret_cast<cudaStream_t>(stream), col_partial, row_tiles, d, dgamma, dbeta, dbias,
             static_cast<int>(accumulate), static_cast<int>(tile_offset), static_cast<int>(total_tiles));
  JL_CHECK_LAUNCH("lnproj_bwd_reduce");
  return JL_OK;
}

int jl_lnproj_bwd(const jl_lnproj_bwd_params* p, void* stream) {
  JL_REQUIRE(p != nullptr, JL_EINVAL, "lnproj_bwd: null params");
  JL_REQUIRE(p->dy && p->y && p->w && p->s && p->tb && p->gamma && p->h && p->mean && p->rstd && p->dres && p->dx, JL_EINVAL, "lnproj_bwd: null pointer");
  JL_REQUIRE(p->rows > 0, JL_EINVAL, "lnproj_bwd: rows must be positive");
  JL_REQUIRE((p->dy_scaled == nullptr) == (p->wgrad_partial == nullptr), JL_EINVAL, "lnproj_bwd: dy_scaled and wgrad_partial go together");
  JL_REQUIRE(p->dy_scaled == nullptr || ((p->lddys % 8) == 0 && (reinterpret_cast<uintptr_t>(p->dy_scaled) & 15) == 0), JL_EINVAL,
             "lnproj_bwd: dy_scaled must be 16-byte aligned with a row stride that is a multiple of 8");
  JL_REQUIRE(p->n >= 8 && p->n <= 192 && (p->n % 8) == 0, JL_EUNSUPPORTED_SHAPE, "lnproj_bwd: n must be a multiple of 8, at most 192 (got %d)", p->n);
  JL_REQUIRE(p->dy_scaled == nullptr || (p->n % 64) == 0, JL_EUNSUPPORTED_SHAPE, "lnproj_bwd: dy_scaled needs n to be a multiple of 64 (got %d)", p->n);
  JL_REQUIRE(p->num_runs >= 0 && p->num_runs <= JL_LNPROJ_MAX_RUNS, JL_EINVAL, "lnproj_bwd: at most %d row runs", JL_LNPROJ_MAX_RUNS);
  JL_REQUIRE(p->num_runs == 0 || p->dy_scaled == nullptr, JL_EUNSUPPORTED_SHAPE, "lnproj_bwd: dy_scaled is not available with row runs");
  const int w_sets = p->w_sets > 0 ? p->w_sets : 1;
  int row_tiles = jl::ceil_div(p->rows, 128);
  if (p->num_runs > 0) {
    row_tiles = 0;
    for (int i = 0; i < p->num_runs; ++i) {
      JL_REQUIRE(p->run_start[i] >= 0 && p->run_start[i + 1] > p->run_start[i] && p->run_start[i + 1] <= p->rows, JL_EINVAL, "lnproj_bwd: run %d is empty or out of range", i);
      JL_REQUIRE(p->run_set[i] >= 0 && p->run_set[i] < w_sets, JL_EINVAL, "lnproj_bwd: run %d names factor set %d of %d", i, p->run_set[i], w_sets);
      row_tiles += jl::ceil_div(p->run_start[i + 1] - p->run_start[i], 128);
    }
  }
  JL_REQUIRE(p->d >= 64 && (p->d % 64) == 0 && p->d <= jl::LP_MAX_D, JL_EUNSUPPORTED_SHAPE, "lnproj_bwd: d must be a multiple of 64, at most %d (got %d)",
             jl::LP_MAX_D, p->d);
  JL_REQUIRE((p->lddy % 8) == 0 && (p->ldy % 8) == 0 && (p->ldh % 8) == 0 && (p->lddres % 8) == 0 && (p->lddx % 8) == 0 && (p->dz == nullptr || (p->lddz % 8) == 0),
             JL_EINVAL, "lnproj_bwd: row strides must be multiples of 8 elements");
  for (const void* q : {p->dy, p->y, p->w, p->h, p->dres, static_cast<const void*>(p->dx), static_cast<const void*>(p->dz)})
    JL_REQUIRE((reinterpret_cast<uintptr_t>(q) & 15) == 0, JL_EINVAL, "lnproj_bwd: pointers must be 16-byte aligned");
  int rc = jl::check_device();
  if (rc != JL_OK) return rc;
  CUtensorMap t_dy, t_y, t_w, t_h, t_r;
  rc = jl::make_tma_map_2d_bf16(&t_dy, p->dy, p->n, p->rows, p->lddy, 128);
  if (rc == JL_OK) rc = jl::make_tma_map_2d_bf16(&t_y, p->y, p->n, p->rows, p->ldy, 128);
  if (rc == JL_OK) rc = jl::make_tma_map_2d_bf16(&t_w, p->w, p->d, static_cast<int64_t>(w_sets) * p->n, p->d, (p->n + 63) / 64 * 64);
  if (rc == JL_OK) rc = jl::make_tma_map_2d_bf16(&t_h, p->h, p->d, p->rows, p->ldh, 128);
  if (rc == JL_OK) rc = jl::make_tma_map_2d_bf16(&t_r, p->dres, p->d, p->rows, p->lddres, 128);
  if (rc != JL_OK) return rc;
  const size_t smem = sizeof(jl::LpSmem) + 1024;
  static thread_local int configured_dev = -1;
  int dev = 0;
  cudaGetDevice(&dev);
  if (configured_dev != dev) {
    cudaError_t e = cudaFuncSetAttribute(jl::lnproj_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    JL_REQUIRE(e == cudaSuccess, JL_ECUDA, "lnproj_bwd: cannot reserve %zu B of shared memory: %s", smem, cudaGetErrorString(e));
    configured_dev = dev;
  }
  // Two CTAs per row tile (column halves) make the kernel itself faster (8000 x 768: 12.7 vs 19.9 µs with dz).  Whether the STEP gets
  // faster depends on what runs beside it: with a busy weight-gradient branch (it fills the SMs this kernel leaves idle) the split only
  // repeats the prologue (base config before the column partials: 6.20 vs 6.17 ms; 24-layer config with two adapters per layer: 18.48 vs
  // 18.30 ms); with a light one it pays (base config now: 5.95 vs 5.99 ms).  The caller knows and says so with col_split; automatic =
  // split only when the row tiles alone would leave most of the machine idle.
  int split = (row_tiles * 8 <= jl::num_sms() && p->d >= 128) ? 2 : 1;
  if (p->col_split > 0 && p->d >= 128 * p->col_split) split = p->col_split;
  if (jl::g_lnproj_split > 0 && p->d >= 128 * jl::g_lnproj_split) split = jl::g_lnproj_split;      // (col_partial rows are indexed by the row tile only: any split works)
  jl::launch(jl::lnproj_bwd_kernel, dim3(row_tiles, split), jl::LP_THREADS, smem, reinterpret_cast<cudaStream_t>(stream), t_dy, t_y, t_w, t_h, t_r, *p);
  JL_CHECK_LAUNCH("lnproj_bwd");
  return JL_OK;
}

}  // extern "C"
