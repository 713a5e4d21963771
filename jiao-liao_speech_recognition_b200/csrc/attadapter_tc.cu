// a7: the AttAdapter forward as ONE kernel (north_star: "AttAdapter's small attention is a single fused kernel"):
//
//   z = LN(h);  q|k|v = z W_qkvᵀ + b  (∈ R^64 each);  a = softmax(q kᵀ / 8 + keymask) v  over the utterance's own frames;
//   out = h + a W_oᵀ + b_o                                            (/root/reference/README.md:1)
//
// for utterances of at most 256 frames (10.2 s at 40 ms — the benchmark's case; longer ones take the composed path: LayerNorm →
// GEMM → jl_attn_fwd → GEMM).  It replaces four launches (and their three HBM round trips of z, q|k|v and a) by one:
// h is read once for the projections and once more (L2) for the residual, out is written once.
//
// One CLUSTER of two CTAs per utterance, one CTA per 128-frame half, 384 threads:
//   warp 0   TMA producer: 64-wide k-chunks of h (own 128 rows) and of W' = W_qkv ⊙ γ through a 4-stage ring — each CTA of an
//            active pair fetches HALF of every W' chunk and multicasts it to both —, W_o in 128-row chunks (3 buffers) and, in phase D,
//            the residual tiles
//   warp 1   tcgen05.mma issuer, accumulators in TMEM (512 columns)
//   warp 2   TMEM allocator
//   warps 4-11  row statistics, then every epilogue: LayerNorm fold + bias → bf16 operand tiles, softmax, a, output
// Phases:
//   A  acc[128, 192] = h_own · W'ᵀ (q, k, v of the CTA's own frames).  The LayerNorm is folded into the projection,
//      (LN(h) Wᵀ)[i, j] = rstd_i (h W'ᵀ[i, j] − μ_i s_j) + t_j, so the tensor cores consume the raw h tiles while the row threads
//      accumulate Σx, Σx² from the same shared-memory tiles (as in wfadapter_tc.cu).
//   B  fold + bias → q, k, v as K-major 128B-swizzled bf16 operand tiles in shared memory; the k and v tiles are ALSO written into
//      the peer CTA's shared memory (st.shared::cluster), so each CTA ends up with the keys and values of all 256 frames without
//      recomputing the other half's projections (the first version did: 883 KB of operand traffic per CTA, now 344 KB).
//   C  S[128, 256] = q · kᵀ;  two-pass softmax over whole rows (8 warps);  P → shared memory;  O[128, 64] = P · v
//   D  a = O / l → bf16 operand tile;  out[128, d] = a · W_oᵀ + b_o + h in chunks of 128 columns: the residual tile arrives by TMA
//      in a staging buffer, each thread adds its row of the accumulator in place, and the tile leaves with coalesced 16-byte
//      stores (a warp writes four 128-byte row segments per instruction; the first version's row-per-thread 32-byte accesses
//      cost 27 k of its 61 k cycles).
// Training also writes q|k|v, a (from the operand tiles, coalesced), the LayerNorm statistics and the row logsumexp.
#include <cuda.h>
#include <math_constants.h>

#include "common.cuh"
#include "ptx_sm100.cuh"

namespace jl {

constexpr int AA_THREADS = 384;
constexpr int AA_STAGES = 4;                        // 4 x 40 KB ring (the TMA round trip is ≈ 1400 cycles, a stage is consumed in ≈ 150)
constexpr uint32_t AA_T128 = 128 * 128;             // bytes of a [128 x 64] bf16 tile
constexpr uint32_t AA_WQKV = 192 * 128;             // bytes of a [192 x 64] bf16 tile
constexpr uint32_t AA_STAGE = AA_T128 + AA_WQKV;    // 40 KB
constexpr int AA_WO_BUFS = 3;                       // W_o chunks of 128 output columns (16 KB), cycled
constexpr int AA_OUT_BUFS = 4;                      // TMEM accumulators of 128 columns for the output projection
constexpr int AA_STG = 3;                           // 32 KB staging buffers (residual in, result out) over the dead operand tiles
constexpr int AA_MAX_D = 1024;
constexpr float AA_LOG2E = 1.4426950408889634f;

struct __align__(1024) AaSmem {
  uint8_t ring[AA_STAGES * AA_STAGE];   // phase A ring; afterwards the operand tiles / staging buffers (offsets below)
  uint8_t wo[AA_WO_BUFS][128 * 128];    // W_o chunks [128 output columns x 64]
  float bo[AA_MAX_D];
  float fs[192], ftb[192];              // LayerNorm-fold vectors s, t (see jl_lnfold_pack)
  float mu[128], rs[128];               // LayerNorm statistics of the CTA's rows
  float red_max[2][128], red_sum[2][128];
  uint64_t full[AA_STAGES], empty[AA_STAGES];
  uint64_t acc_full;                    // phase A accumulators complete
  uint64_t kv_ready;                    // q, k, v operand tiles written (8 warps, + the peer's 8 when the pair is active)
  uint64_t s_full, p_full, o_full, a_ready;
  uint64_t wo_full[AA_WO_BUFS], wo_empty[AA_WO_BUFS], out_full[AA_OUT_BUFS], out_empty[AA_OUT_BUFS], res_full[AA_STG], res_empty[AA_STG];
  uint32_t tmem_slot;
#ifdef JL_AA_TIMING
  long long ts_prod[16], ts_mma[16], ts_stat[16], ts_out[8][5];
#endif
};
// operand tiles inside `ring` once phase A is over
constexpr uint32_t AA_OFF_Q = 0;                    // [128 x 64]            later P tile 0
constexpr uint32_t AA_OFF_K = AA_T128;              // [256 x 64] (2 tiles)  later P tiles 1, 2
constexpr uint32_t AA_OFF_PT = 3 * AA_T128;         //                       P tile 3
constexpr uint32_t AA_OFF_V = 4 * AA_T128;          // [256 x 64] = four 64-key tiles, read as MN-major B operands
constexpr uint32_t AA_OFF_A = 6 * AA_T128;          // [128 x 64]
// staging buffer i = ring + i * 32 KB (i < 3): over q / k / P / v, which are dead once O = P · v is complete
static_assert(AA_STG * 2 * AA_T128 <= AA_OFF_A, "staging buffers must not reach the `a` operand tile");

__device__ __forceinline__ float aa_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ uint32_t aa_chunk_off(int r, int c) { return static_cast<uint32_t>(r * 128 + ((c ^ (r & 7)) << 4)); }
__device__ __forceinline__ void aa_store_chunk(uint8_t* tile, int r, int c, const uint32_t* pk) {      // 8 bf16 = 16 B, chunk c of row r
  *reinterpret_cast<uint4*>(tile + aa_chunk_off(r, c)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
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
// Coalesced copy of NT [128 x 64] bf16 operand tiles (shared-memory addresses tile[i]) to global rows: tile i → columns
// [i * 64, i * 64 + 64) of the row.  256 threads; thread t owns 16-byte chunk (t & 7) of rows (t >> 3) + 32 j: a warp writes four
// 128-byte row segments per instruction, and the swizzled chunk offset is the same for all of a thread's rows.
__device__ __forceinline__ uint4 aa_lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}
template <int NT>
__device__ __forceinline__ void aa_tiles_to_global(const uint32_t (&tile)[NT], __nv_bfloat16* dst, int64_t ld, int rows_valid, int t) {
  const int ch = t & 7, r0 = t >> 3;
  const uint32_t off = static_cast<uint32_t>(r0 * 128 + ((ch ^ (r0 & 7)) << 4));
  __nv_bfloat16* d0 = dst + static_cast<int64_t>(r0) * ld + ch * 8;
#pragma unroll
  for (int i = 0; i < NT; ++i) {
    uint4 v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) v[j] = aa_lds128(tile[i] + off + j * 4096);
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (r0 + 32 * j < rows_valid) *reinterpret_cast<uint4*>(d0 + static_cast<int64_t>(32 * j) * ld + i * 64) = v[j];
  }
}

#ifdef JL_AA_TIMING
#define AA_T(i) do { if (threadIdx.x == 128 && blockIdx.x == 0 && blockIdx.y == 0) aa_ts[i] = clock64(); } while (0)
#else
#define AA_T(i) do { } while (0)
#endif

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(AA_THREADS, 1)
attadapter_fwd_kernel(const __grid_constant__ CUtensorMap t_h, const __grid_constant__ CUtensorMap t_w, const __grid_constant__ CUtensorMap t_w96,
                      const __grid_constant__ CUtensorMap t_wo, const jl_attadapter_fwd_params p) {
  extern __shared__ uint8_t aa_smem_raw[];
  AaSmem& s = *reinterpret_cast<AaSmem*>(aa_smem_raw + ((1024u - (ptx::smem_u32(aa_smem_raw) & 1023u)) & 1023u));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = blockIdx.x, b = blockIdx.y;         // frame half (= rank in the cluster), utterance
  const uint32_t peer = static_cast<uint32_t>(g ^ 1);
  const int nk = p.d / 64;                          // k-chunks of the projections
  // 128-column chunks of the output projection; gridDim.z clusters per utterance share them (each repeats phases A-C: more SM-time,
  // a shorter phase D — for when the kernel's latency, not the machine's occupancy, bounds the step)
  const int nc_all = p.d / 128;
  const int c_begin = static_cast<int>((static_cast<int64_t>(nc_all) * blockIdx.z) / gridDim.z);
  const int nc = static_cast<int>((static_cast<int64_t>(nc_all) * (blockIdx.z + 1)) / gridDim.z) - c_begin;
  const bool first_z = blockIdx.z == 0;             // writes the tensors saved for the backward pass

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&t_h);
    ptx::prefetch_tensormap(&t_w);
    ptx::prefetch_tensormap(&t_w96);
    ptx::prefetch_tensormap(&t_wo);
  }
  if (warp == 2) {
    ptx::tmem_alloc(&s.tmem_slot, 512);
    ptx::tmem_relinquish();
  }
  jl::pdl_prologue();           // h, lengths and the packed weights may come from the preceding kernels
#ifdef JL_AA_TIMING
  long long aa_ts[12];
  for (int i = 0; i < 12; ++i) aa_ts[i] = 0;
#endif

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
  const bool active = g * 128 < len;                   // the CTA's frame half holds at least one valid frame
  const bool pair = len > 128;                         // both CTAs of the cluster are active: they share W' and exchange k, v
  const int nkt = (len + 63) / 64;                     // 64-key tiles with valid keys
  const uint32_t t_acc = 0, t_s = 0, t_o = 256;        // TMEM column offsets
  uint8_t* R = s.ring;

  if (warp == 1 && lane == 0) {
    for (int i = 0; i < AA_STAGES; ++i) {
      ptx::mbar_init(&s.full[i], 1);
      ptx::mbar_init(&s.empty[i], pair ? 18 : 9);     // MMA commit + the 8 statistics warps, of every CTA the stage's W' half goes to
    }
    ptx::mbar_init(&s.acc_full, 1);
    ptx::mbar_init(&s.kv_ready, pair ? 16 : 8);
    ptx::mbar_init(&s.s_full, 1);
    ptx::mbar_init(&s.p_full, 8);
    ptx::mbar_init(&s.o_full, 1);
    ptx::mbar_init(&s.a_ready, 8);
    for (int i = 0; i < AA_WO_BUFS; ++i) {
      ptx::mbar_init(&s.wo_full[i], 1);
      ptx::mbar_init(&s.wo_empty[i], 1);
    }
    for (int i = 0; i < AA_OUT_BUFS; ++i) {
      ptx::mbar_init(&s.out_full[i], 1);
      ptx::mbar_init(&s.out_empty[i], 8);
    }
    for (int i = 0; i < AA_STG; ++i) {
      ptx::mbar_init(&s.res_full[i], 1);
      ptx::mbar_init(&s.res_empty[i], 1);
    }
    ptx::fence_barrier_init();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = s.tmem_slot;
  ptx::cluster_sync_all();      // the peer's barriers exist before anything is multicast to it or arrives on them
  // Second cluster barrier phase = "phase A is over in both CTAs" (their rings may be overwritten with operand tiles): the
  // epilogue warps of an active CTA arrive once their accumulators are complete and their statistics are done; everybody else
  // has nothing to protect and arrives at once.  All waits pair up below.
  if (!active || warp < 4) ptx::cluster_arrive();
  AA_T(0);

  if (warp == 0) {
    if (lane == 0 && active) {
      // the first W_o chunks right away: dedicated buffers
      for (int c = 0; c < min(nc, AA_WO_BUFS); ++c) {
        ptx::mbar_expect_tx(&s.wo_full[c], AA_T128);
        ptx::tma_load_2d(s.wo[c], &t_wo, &s.wo_full[c], 0, (c_begin + c) * 128);
      }
      for (int kc = 0; kc < nk; ++kc) {
        const int st = kc % AA_STAGES;
        ptx::mbar_wait(&s.empty[st], ((kc / AA_STAGES) & 1) ^ 1u);
#ifdef JL_AA_TIMING
        if (kc < 16) s.ts_prod[kc] = clock64();
#endif
        uint8_t* base = R + st * AA_STAGE;
        ptx::mbar_expect_tx(&s.full[st], AA_STAGE);
        ptx::tma_load_2d(base, &t_h, &s.full[st], kc * 64, grow + g * 128);
        if (pair) ptx::tma_load_2d_mcast(base + AA_T128 + g * (96 * 128), &t_w96, &s.full[st], kc * 64, g * 96, static_cast<uint16_t>(3));
        else ptx::tma_load_2d(base + AA_T128, &t_w, &s.full[st], kc * 64, 0);
      }
      jl::pdl_trigger_late();
      // phase D: the residual tile of output chunk c lands in staging buffer c % 3 (over q / k / P / v: dead once O = P · v is
      // complete; afterwards a buffer is free when the epilogue has stored the chunk it held), and the remaining W_o chunks
      ptx::mbar_wait(&s.o_full, 0);
      for (int c = 0; c < nc; ++c) {
        const int sb = c % AA_STG;
        if (c >= AA_STG) ptx::mbar_wait(&s.res_empty[sb], ((c / AA_STG) - 1) & 1);
        uint8_t* stg = R + sb * (2 * AA_T128);
        ptx::mbar_expect_tx(&s.res_full[sb], 2 * AA_T128);
        ptx::tma_load_2d(stg, &t_h, &s.res_full[sb], (c_begin + c) * 128, grow + g * 128);
        ptx::tma_load_2d(stg + AA_T128, &t_h, &s.res_full[sb], (c_begin + c) * 128 + 64, grow + g * 128);
        if (c >= AA_WO_BUFS) {
          const int wb = c % AA_WO_BUFS;
          ptx::mbar_wait(&s.wo_empty[wb], ((c / AA_WO_BUFS) - 1) & 1);
          ptx::mbar_expect_tx(&s.wo_full[wb], AA_T128);
          ptx::tma_load_2d(s.wo[wb], &t_wo, &s.wo_full[wb], 0, (c_begin + c) * 128);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && active) {
      // ---- phase A
      for (int kc = 0; kc < nk; ++kc) {
        const int st = kc % AA_STAGES;
        ptx::mbar_wait(&s.full[st], (kc / AA_STAGES) & 1);
        ptx::tc_fence_after();
#ifdef JL_AA_TIMING
        if (kc < 16) s.ts_mma[kc] = clock64();
#endif
        const uint32_t base = ptx::smem_u32(R + st * AA_STAGE);
        aa_mma_kk(tmem + t_acc, base, base + AA_T128, 192, kc > 0);
        if (pair) ptx::umma_commit_mcast(&s.empty[st], static_cast<uint16_t>(3));
        else ptx::umma_commit(&s.empty[st]);
      }
      ptx::umma_commit(&s.acc_full);
      // ---- phase C: S = q · kᵀ (N = 256 keys), then O = P · v over the key tiles that hold frames
      ptx::mbar_wait_cluster(&s.kv_ready, 0);
      ptx::tc_fence_after();
      const uint32_t rb = ptx::smem_u32(R);
      aa_mma_kk(tmem + t_s, rb + AA_OFF_Q, rb + AA_OFF_K, 256, false);
      ptx::umma_commit(&s.s_full);
      ptx::mbar_wait(&s.p_full, 0);
      ptx::tc_fence_after();
      for (int kt = 0; kt < nkt; ++kt) aa_mma_pv(tmem + t_o, rb + kt * AA_T128, rb + AA_OFF_V + kt * (64 * 128), kt > 0);
      ptx::umma_commit(&s.o_full);
      // ---- phase D: out chunk c = a · W_o[c]ᵀ
      ptx::mbar_wait(&s.a_ready, 0);
      ptx::tc_fence_after();
      for (int c = 0; c < nc; ++c) {
        const int wb = c % AA_WO_BUFS, ob = c % AA_OUT_BUFS;
        ptx::mbar_wait(&s.wo_full[wb], (c / AA_WO_BUFS) & 1);
        ptx::mbar_wait(&s.out_empty[ob], ((c / AA_OUT_BUFS) & 1) ^ 1u);
        ptx::tc_fence_after();
        aa_mma_kk(tmem + ob * 128, rb + AA_OFF_A, ptx::smem_u32(s.wo[wb]), 128, false);
        ptx::umma_commit(&s.wo_empty[wb]);
        ptx::umma_commit(&s.out_full[ob]);
      }
    }
  } else if (warp >= 4) {
    const int quad = warp & 3;
    const int grp = (warp - 4) >> 2;                     // 0: warps 4-7, 1: warps 8-11
    const int r = quad * 32 + lane;                      // TMEM lane = row of the 128-row tile
    const int et = threadIdx.x - 128;                    // 0..255 among the epilogue threads
    const uint32_t lane_off = static_cast<uint32_t>(quad * 32) << 16;
    const int qrow = g * 128 + r;                        // frame index inside the utterance
    const int rows_owned = min(128, lim - g * 128);      // rows of this tile the layout gives to this utterance (may be <= 0)
    if (!active) {
      // no valid frame in this half: the rows the layout still owns get what the composed path gives them — a = 0, so
      // out = b_o + h (or 0 when the caller wants padded rows zeroed); saved tensors are zero there
      if (qrow < lim && first_z) {
        __nv_bfloat16* out_row = reinterpret_cast<__nv_bfloat16*>(p.out) + (row_base + qrow) * p.ldo;
        const __nv_bfloat16* h_row = reinterpret_cast<const __nv_bfloat16*>(p.h) + (row_base + qrow) * p.ldh;
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
      for (int i = et; i < p.d; i += 256) s.bo[i] = __ldg(p.bo + i);
      if (et < 192) { s.fs[et] = __ldg(p.s + et); s.ftb[et] = __ldg(p.tb + et); }
      // ---- phase A (all 8 warps): LayerNorm statistics of the staged h tiles — warp group 0 sums the first 32 columns of every
      //      64-column chunk, group 1 the other 32; four independent accumulator chains per thread
      {
        float sx[4] = {0.0f, 0.0f, 0.0f, 0.0f}, sq[4] = {0.0f, 0.0f, 0.0f, 0.0f};
        for (int kc = 0; kc < nk; ++kc) {
          const int st = kc % AA_STAGES;
          ptx::mbar_wait(&s.full[st], (kc / AA_STAGES) & 1);
          const uint8_t* tile = R + st * AA_STAGE;
          uint4 v[4];
#pragma unroll
          for (int c = 0; c < 4; ++c) v[c] = *reinterpret_cast<const uint4*>(tile + aa_chunk_off(r, grp * 4 + c));
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const uint32_t w[4] = {v[c].x, v[c].y, v[c].z, v[c].w};
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const float2 f = unpack_bf16x2(w[q]);
              sx[q] += f.x;
              sq[q] = fmaf(f.x, f.x, sq[q]);
              sx[q] += f.y;
              sq[q] = fmaf(f.y, f.y, sq[q]);
            }
          }
          __syncwarp();
#ifdef JL_AA_TIMING
          if (threadIdx.x == 128 && kc < 16) s.ts_stat[kc] = clock64();
#endif
          if (lane == 0) {
            ptx::mbar_arrive(&s.empty[st]);
            if (pair) ptx::mbar_arrive_remote(ptx::mapa_shared(ptx::smem_u32(&s.empty[st]), peer));
          }
        }
        s.red_sum[grp][r] = (sx[0] + sx[1]) + (sx[2] + sx[3]);
        s.red_max[grp][r] = (sq[0] + sq[1]) + (sq[2] + sq[3]);
        asm volatile("bar.sync 2, 256;" ::: "memory");
        if (grp == 0) {
          const float tx = s.red_sum[0][r] + s.red_sum[1][r], txx = s.red_max[0][r] + s.red_max[1][r];
          const float inv_d = 1.0f / static_cast<float>(p.d);
          const float mu = tx * inv_d;
          const float var = fmaxf(txx * inv_d - mu * mu, 0.0f);
          const float rstd = 1.0f / sqrtf(var + p.eps);
          s.mu[r] = mu;
          s.rs[r] = rstd;
          if (p.mean != nullptr && qrow < lim && first_z) {
            p.mean[row_base + qrow] = mu;
            p.rstd[row_base + qrow] = rstd;
          }
        }
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");       // statistics (and b_o) visible to all 8 warps
      AA_T(1);
      ptx::mbar_wait(&s.acc_full, 0);
      ptx::tc_fence_after();
      AA_T(2);
      ptx::cluster_arrive();                               // this CTA's ring is free …
      ptx::cluster_wait();                                 // … and so is the peer's
      // ---- phase B: fold + bias → q, k, v operand tiles: 6 chunks of 32 columns (q0 q1 k0 | k1 v0 v1), three per warp group;
      //      the k and v chunks also go to the same place in the peer's shared memory
      const float mu = s.mu[r], rstd = s.rs[r];
#pragma unroll 1
      for (int item = 0; item < 3; ++item) {
        const int ch = grp * 3 + item;
        const int col0 = ch * 32;                                            // column of W_qkv (q 0-63, k 64-127, v 128-191)
        uint32_t v[32];
        ptx::tmem_ld_32x32(tmem + t_acc + lane_off + ch * 32, v);
        ptx::tmem_ld_wait();
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const int j = col0 + 2 * i;
          const float2 fs = *reinterpret_cast<const float2*>(s.fs + j), ft = *reinterpret_cast<const float2*>(s.ftb + j);
          const float x0 = fmaf(rstd, __uint_as_float(v[2 * i]) - mu * fs.x, ft.x);
          const float x1 = fmaf(rstd, __uint_as_float(v[2 * i + 1]) - mu * fs.y, ft.y);
          pk[i] = pack_bf16x2(x0, x1);
        }
        // destination tile: q; k / v of this CTA's frame half (utterance order: half 0 = frames 0-127)
        uint8_t* tile;
        if (col0 < 64) tile = R + AA_OFF_Q;
        else if (col0 < 128) tile = R + AA_OFF_K + g * AA_T128;
        else tile = R + AA_OFF_V + g * AA_T128;
        const int c0 = ((col0 & 63) >> 3);                                   // first 16-byte chunk inside the 64-wide tile row: 0 or 4
#pragma unroll
        for (int c = 0; c < 4; ++c) aa_store_chunk(tile, r, c0 + c, pk + 4 * c);
        if (pair && col0 >= 64) {
          const uint32_t remote = ptx::mapa_shared(ptx::smem_u32(tile), peer);
#pragma unroll
          for (int c = 0; c < 4; ++c)
            ptx::st_shared_cluster_v4(remote + aa_chunk_off(r, c0 + c), pk[4 * c], pk[4 * c + 1], pk[4 * c + 2], pk[4 * c + 3]);
        }
      }
      ptx::tc_fence_before();
      ptx::fence_proxy_async_all();
      __syncwarp();
      if (lane == 0) {
        ptx::mbar_arrive(&s.kv_ready);
        if (pair) ptx::mbar_arrive_cluster(ptx::mapa_shared(ptx::smem_u32(&s.kv_ready), peer));
      }
      AA_T(3);
      if (p.qkv_out != nullptr && first_z) {
        // q | k | v of the CTA's rows, from the operand tiles (they stay until the softmax overwrites them with P)
        asm volatile("bar.sync 2, 256;" ::: "memory");
        const uint32_t rb = ptx::smem_u32(R);
        const uint32_t tiles[3] = {rb + AA_OFF_Q, rb + AA_OFF_K + g * AA_T128, rb + AA_OFF_V + g * AA_T128};
        aa_tiles_to_global<3>(tiles, reinterpret_cast<__nv_bfloat16*>(p.qkv_out) + (row_base + g * 128) * 192, 192, rows_owned, et);
      }
      // ---- phase C: softmax over the whole row of 256 scores (this thread: 128 of them), P → operand tiles over q / k
      const float sl2 = p.scale * AA_LOG2E;
      const int kbase = grp * 128;
      const uint32_t t_srow = tmem + t_s + lane_off + 128u * grp;
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
      // ---- phase D: a = O / l (this thread: 32 of the 64 dims) → operand tile; lse
      ptx::mbar_wait(&s.o_full, 0);
      ptx::tc_fence_after();
      AA_T(6);
      asm volatile("bar.sync 2, 256;" ::: "memory");
      const float l_tot = sum + s.red_sum[grp ^ 1][r];
      const bool valid = qrow < len;
      const float inv = valid ? 1.0f / l_tot : 0.0f;
      {
        uint32_t ov[32];
        ptx::tmem_ld_32x32(tmem + t_o + lane_off + 32u * grp, ov);
        ptx::tmem_ld_wait();
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) pk[i] = pack_bf16x2(__uint_as_float(ov[2 * i]) * inv, __uint_as_float(ov[2 * i + 1]) * inv);
#pragma unroll
        for (int c = 0; c < 4; ++c) aa_store_chunk(R + AA_OFF_A, r, grp * 4 + c, pk + 4 * c);
        if (p.lse != nullptr && grp == 0 && qrow < lim && first_z)
          p.lse[(p.cu_seqlens ? row_base : static_cast<int64_t>(b) * p.seq) + qrow] = valid ? mx * p.scale + logf(l_tot) : 0.0f;
      }
      ptx::tc_fence_before();
      ptx::fence_proxy_async();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&s.a_ready);
      AA_T(7);
      if (p.a_out != nullptr && first_z) {
        asm volatile("bar.sync 2, 256;" ::: "memory");
        const uint32_t tiles[1] = {ptx::smem_u32(R) + AA_OFF_A};
        aa_tiles_to_global<1>(tiles, reinterpret_cast<__nv_bfloat16*>(p.a_out) + (row_base + g * 128) * 64, 64, rows_owned, et);
      }
      // ---- output projection epilogue, 128 columns per chunk: this thread adds its row's 64 accumulator columns (+ b_o) to the
      //      residual tile in place, then the 256 threads store the tile with coalesced 16-byte accesses
      const bool zero = p.zero_padded_rows && !valid;
      __nv_bfloat16* out_tile = reinterpret_cast<__nv_bfloat16*>(p.out) + (row_base + g * 128) * p.ldo;
      for (int c = 0; c < nc; ++c) {
        const int sb = c % AA_STG, ob = c % AA_OUT_BUFS;
        uint8_t* stg = R + sb * (2 * AA_T128);
        ptx::mbar_wait(&s.res_full[sb], (c / AA_STG) & 1);
        ptx::mbar_wait(&s.out_full[ob], (c / AA_OUT_BUFS) & 1);
        ptx::tc_fence_after();
        if (c == 0) AA_T(8);
#ifdef JL_AA_TIMING
#define AA_TO(i) do { if (threadIdx.x == 128 && c < 8) s.ts_out[c][i] = clock64(); } while (0)
#else
#define AA_TO(i) do { } while (0)
#endif
        AA_TO(0);
        uint32_t va[32], vb[32];
        ptx::tmem_ld_32x32(tmem + ob * 128 + lane_off + grp * 64, va);
        ptx::tmem_ld_32x32(tmem + ob * 128 + lane_off + grp * 64 + 32, vb);
        ptx::tmem_ld_wait();
        uint8_t* mytile = stg + grp * AA_T128;
        const float* bias = s.bo + (c_begin + c) * 128 + grp * 64;
#pragma unroll
        for (int ch = 0; ch < 8; ++ch) {
          uint4* cell = reinterpret_cast<uint4*>(mytile + aa_chunk_off(r, ch));
          const uint4 hv = *cell;
          const uint32_t hw[4] = {hv.x, hv.y, hv.z, hv.w};
          const float4 b0 = *reinterpret_cast<const float4*>(bias + ch * 8), b1 = *reinterpret_cast<const float4*>(bias + ch * 8 + 4);
          const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
          uint32_t w[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int j = (ch & 3) * 8 + 2 * i;
            const float a0 = __uint_as_float(ch < 4 ? va[j] : vb[j]), a1 = __uint_as_float(ch < 4 ? va[j + 1] : vb[j + 1]);
            const float2 hr = unpack_bf16x2(hw[i]);
            w[i] = zero ? 0u : pack_bf16x2(a0 + bb[2 * i] + hr.x, a1 + bb[2 * i + 1] + hr.y);
          }
          *cell = make_uint4(w[0], w[1], w[2], w[3]);
        }
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&s.out_empty[ob]);
        AA_TO(1);
        asm volatile("bar.sync 2, 256;" ::: "memory");
        AA_TO(2);
        {
          const uint32_t tiles[2] = {ptx::smem_u32(stg), ptx::smem_u32(stg) + AA_T128};
          aa_tiles_to_global<2>(tiles, out_tile + (c_begin + c) * 128, p.ldo, rows_owned, et);
        }
        AA_TO(3);
        asm volatile("bar.sync 2, 256;" ::: "memory");
        AA_TO(4);
        if (et == 0) {
          ptx::fence_proxy_async();
          ptx::mbar_arrive(&s.res_empty[sb]);
        }
      }
      AA_T(9);
#ifdef JL_AA_TIMING
      if (threadIdx.x == 128 && blockIdx.x == 0 && blockIdx.y == 0)
        printf("aa phases (cycles): stats %lld  acc_wait %lld  fold %lld  s_wait %lld  softmax %lld  o_wait %lld  a %lld  out_wait %lld  out %lld  total %lld\n",
               aa_ts[1] - aa_ts[0], aa_ts[2] - aa_ts[1], aa_ts[3] - aa_ts[2], aa_ts[4] - aa_ts[3], aa_ts[5] - aa_ts[4], aa_ts[6] - aa_ts[5],
               aa_ts[7] - aa_ts[6], aa_ts[8] - aa_ts[7], aa_ts[9] - aa_ts[8], aa_ts[9] - aa_ts[0]);
      if (threadIdx.x == 128 && blockIdx.x == 0 && blockIdx.y == 0)
        for (int kc = 0; kc < min(nk, 16); ++kc)
          printf("  kc %2d  producer-issue %6lld  mma-sees-full %6lld  stats-done %6lld\n", kc, s.ts_prod[kc] - aa_ts[0], s.ts_mma[kc] - aa_ts[0], s.ts_stat[kc] - aa_ts[0]);
      if (threadIdx.x == 128 && blockIdx.x == 0 && blockIdx.y == 0)
        for (int c = 0; c < min(nc, 8); ++c)
          printf("  out chunk %d  start %6lld  compute %5lld  bar %5lld  store %5lld  bar %5lld\n", c, s.ts_out[c][0] - aa_ts[8], s.ts_out[c][1] - s.ts_out[c][0],
                 s.ts_out[c][2] - s.ts_out[c][1], s.ts_out[c][3] - s.ts_out[c][2], s.ts_out[c][4] - s.ts_out[c][3]);
#endif
    }
  }
  if (!active || warp < 4) {
    __syncwarp();
    ptx::cluster_wait();
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

// Several projections in one launch (all the AttAdapters of a model at the start of a training step): blockIdx.y = job.
struct LnFoldJobs {
  jl_lnfold_pack_params job[JL_LNFOLD_MAX_JOBS];
};
__global__ void __launch_bounds__(256) lnfold_pack_multi_kernel(const __grid_constant__ LnFoldJobs jobs) {
  jl::pdl_prologue();
  const jl_lnfold_pack_params& p = jobs.job[blockIdx.y];
  if (static_cast<int>(blockIdx.x) >= p.n) return;
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

int jl_lnfold_pack_multi(const jl_lnfold_pack_params* jobs, int32_t count, void* stream) {
  JL_REQUIRE(jobs != nullptr && count >= 1 && count <= JL_LNFOLD_MAX_JOBS, JL_EINVAL, "lnfold_pack_multi: 1..%d jobs", JL_LNFOLD_MAX_JOBS);
  jl::LnFoldJobs js;
  int nmax = 0;
  for (int i = 0; i < count; ++i) {
    const jl_lnfold_pack_params& p = jobs[i];
    JL_REQUIRE(p.w && p.gamma && p.beta && p.w_scaled && p.s && p.tb && p.n > 0 && p.d > 0, JL_EINVAL, "lnfold_pack_multi: job %d: null pointer or bad dims", i);
    js.job[i] = p;
    nmax = p.n > nmax ? p.n : nmax;
  }
  for (int i = count; i < JL_LNFOLD_MAX_JOBS; ++i) js.job[i] = jobs[0];
  int rc = jl::check_device();
  if (rc != JL_OK) return rc;
  jl::launch(jl::lnfold_pack_multi_kernel, dim3(nmax, count), 256, 0, reinterpret_cast<cudaStream_t>(stream), js);
  JL_CHECK_LAUNCH("lnfold_pack_multi");
  return JL_OK;
}

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
  JL_REQUIRE(p->d >= 128 && (p->d % 128) == 0 && p->d <= jl::AA_MAX_D, JL_EUNSUPPORTED_SHAPE,
             "attadapter_fwd: d must be a multiple of 128, at most %d (got %d): use the composed path", jl::AA_MAX_D, p->d);
  JL_REQUIRE((p->ldh % 16) == 0 && (p->ldo % 16) == 0, JL_EINVAL, "attadapter_fwd: row strides must be multiples of 16 elements");
  JL_REQUIRE(((reinterpret_cast<uintptr_t>(p->h) | reinterpret_cast<uintptr_t>(p->out)) & 31) == 0, JL_EINVAL, "attadapter_fwd: h / out must be 32-byte aligned");
  JL_REQUIRE(p->cu_seqlens == nullptr || p->total_rows > 0, JL_EINVAL, "attadapter_fwd: packed layout needs total_rows > 0");
  JL_REQUIRE((p->mean == nullptr) == (p->rstd == nullptr), JL_EINVAL, "attadapter_fwd: mean and rstd go together");
  for (const void* q : {p->qkv_out, p->a_out})
    JL_REQUIRE((reinterpret_cast<uintptr_t>(q) & 15) == 0, JL_EINVAL, "attadapter_fwd: qkv_out / a_out must be 16-byte aligned");
  int rc = jl::check_device();
  if (rc != JL_OK) return rc;
  const int64_t rows = p->cu_seqlens ? static_cast<int64_t>(p->total_rows) : static_cast<int64_t>(p->batch) * p->seq;
  CUtensorMap t_h, t_w, t_w96, t_wo;
  rc = jl::make_tma_map_2d_bf16(&t_h, p->h, p->d, rows, p->ldh, 128);
  if (rc == JL_OK) rc = jl::make_tma_map_2d_bf16(&t_w, p->wqkv_scaled, p->d, 192, p->d, 192);
  if (rc == JL_OK) rc = jl::make_tma_map_2d_bf16(&t_w96, p->wqkv_scaled, p->d, 192, p->d, 96);
  if (rc == JL_OK) rc = jl::make_tma_map_2d_bf16(&t_wo, p->wo, 64, p->d, 64, 128);
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
  const int zs = (p->col_split >= 2 && p->d >= 256) ? 2 : 1;
  jl::launch(jl::attadapter_fwd_kernel, dim3(2, p->batch, zs), jl::AA_THREADS, smem, reinterpret_cast<cudaStream_t>(stream), t_h, t_w, t_w96, t_wo, *p);
  JL_CHECK_LAUNCH("attadapter_fwd");
  return JL_OK;
}

}  // extern "C"
