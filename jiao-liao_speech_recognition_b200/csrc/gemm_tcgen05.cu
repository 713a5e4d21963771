// bf16 GEMM on the 5th-generation tensor cores (tcgen05.mma, accumulators in TMEM) fed by TMA.
//
//   C[M,N] = epilogue(alpha * A · Bᵀ + bias)          A: [M,K] (or [K,M]),  B: [N,K] (or [K,N]),  fp32 accumulate
//
// Replaces every dense contraction of the path: q/k/v/out projections, FFN, lm_head
// (SP/transformers/models/wav2vec2/modeling_wav2vec2.py:495-498,524-528,547,557-573,1708), the two
// Conv1d+GLU subsampling layers through im2col (SP/transformers/models/speech_to_text/
// modeling_speech_to_text.py:82-99), the adapter projections, and — with the MN-major operand modes —
// the dgrad (dY·W) and wgrad (dYᵀ·X) products of the adapter-only backward without transposed copies.
//
// Structure (one persistent CTA per SM, 512 threads, warp-specialised):
//   warp 0      TMA producer: cp.async.bulk.tensor 128B-swizzled boxes → STAGES-deep smem ring (mbarrier full/empty)
//   warp 1      MMA issuer: one elected lane issues 4 × tcgen05.mma (128 × BN × 16) per 64-wide k-block,
//               tcgen05.commit releases the smem slot / publishes the accumulator
//   warp 2      TMEM allocator (2 accumulator stages × BN fp32 columns) so the epilogue of tile i overlaps the
//               mainloop of tile i+1
//   warps 4-15  epilogue: tcgen05.ld 32 lanes × 32 columns → registers → bias / GELU / ReLU / GLU / residual /
//               activation-gradient / row masking → 16-byte global stores
#include <cuda.h>
#include <math_constants.h>

#include <atomic>
#include <mutex>

#include "common.cuh"
#include "ptx_sm100.cuh"

namespace jl {

constexpr int GEMM_BM = 128;
constexpr int GEMM_BK = 64;
#ifndef JL_GEMM_EPI_WARPS
#define JL_GEMM_EPI_WARPS 12
#endif
constexpr int GEMM_EPI_WARPS = JL_GEMM_EPI_WARPS;   // a multiple of 4 (one warp per TMEM lane quadrant and column group)
constexpr int GEMM_THREADS = 128 + GEMM_EPI_WARPS * 32;
constexpr int GEMM_A_BYTES = GEMM_BM * GEMM_BK * 2;

struct GemmDev {
  void* c;
  int64_t ldc;
  const float* bias;
  const __nv_bfloat16* residual;
  int64_t ldr;
  const __nv_bfloat16* aux;
  int64_t ldaux;
  __nv_bfloat16* aux_out;
  int64_t ldaux_out;
  const int32_t* row_lengths;
  int32_t rows_per_seq;
  int32_t m, n, k;
  int32_t epilogue;
  int32_t out_dtype;
  float alpha;
  int32_t num_m_tiles, num_n_tiles;
  // split-K (weight-gradient shapes: few output tiles, very long K): unit = (tile, split); raw fp32 partials go to ws
  int32_t split_k, kb_per_split;
  float* ws;
  int64_t ldw;
  // tail split (pair kernel): tiles [tail_first, num_tiles) are the partial last wave; each is cut into tail_split K ranges that
  // run on otherwise idle CTA pairs, raw partial accumulators go to tail_ws and the LAST range to arrive (per epilogue warp,
  // counted in tail_cnt) adds them in range order and applies the epilogue — no CTA ever waits for another one.
  // N split of the tail wave (single-CTA kernel): tiles [nsplit_first, tiles) — the partial last wave — are cut into nsplit
  // column slices of BN / nsplit, each an independent (shorter) work unit: the tail wave then takes 1 / nsplit of a tile time
  // and needs no partial sums.
  int32_t nsplit_first, nsplit;
  int32_t tail_first, tail_split, tail_kb_per;
  int64_t tail_stride;     // floats per K range in tail_ws = tail tiles · 256 · BN
  float* tail_ws;
  int32_t* tail_cnt;
};

// One schedulable piece of the pair kernel: a whole output tile, or one K range of a tail tile.
struct PairUnit {
  int tile, kb0, kb1, sidx;
  bool split;
};
__device__ __forceinline__ int pair_num_units(const GemmDev& g, int num_tiles) {
  return g.tail_first + (num_tiles - g.tail_first) * g.tail_split;
}
__device__ __forceinline__ PairUnit pair_unit(const GemmDev& g, int u, int num_kb) {
  PairUnit r;
  if (u < g.tail_first) {
    r.tile = u; r.kb0 = 0; r.kb1 = num_kb; r.sidx = 0; r.split = false;
  } else {
    const int v = u - g.tail_first;
    r.tile = g.tail_first + v / g.tail_split;
    r.sidx = v - (r.tile - g.tail_first) * g.tail_split;
    r.kb0 = r.sidx * g.tail_kb_per;
    r.kb1 = min(num_kb, r.kb0 + g.tail_kb_per);
    r.split = g.tail_split > 1;
  }
  return r;
}

// ------------------------------------------------------------------------------------------------
// Epilogue for one row × 32 consecutive accumulator columns (shared by the tcgen05 kernel and the
// test-only SIMT reference so that both apply bit-identical post-processing).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void load_bf16_row32(const __nv_bfloat16* p, bool vec, int nvalid, float (&out)[32]) {
  if (vec) {
    const uint4* q = reinterpret_cast<const uint4*>(p);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      uint4 v = __ldg(q + i);
      float2 f0 = unpack_bf16x2(v.x), f1 = unpack_bf16x2(v.y), f2 = unpack_bf16x2(v.z), f3 = unpack_bf16x2(v.w);
      out[8 * i + 0] = f0.x; out[8 * i + 1] = f0.y; out[8 * i + 2] = f1.x; out[8 * i + 3] = f1.y;
      out[8 * i + 4] = f2.x; out[8 * i + 5] = f2.y; out[8 * i + 6] = f3.x; out[8 * i + 7] = f3.y;
    }
  } else {
#pragma unroll
    for (int j = 0; j < 32; ++j) out[j] = (j < nvalid) ? __bfloat162float(p[j]) : 0.0f;
  }
}

template <int COUNT>
__device__ __forceinline__ void store_bf16_row(__nv_bfloat16* p, bool vec, int nvalid, const float (&v)[COUNT]) {
  if (vec && (reinterpret_cast<uintptr_t>(p) & 31) == 0) {      // 32-byte sectors written by single requests
#pragma unroll
    for (int i = 0; i < COUNT / 16; ++i) {
      uint32_t w[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) w[j] = pack_bf16x2(v[16 * i + 2 * j], v[16 * i + 2 * j + 1]);
      st_global_v8(p + 16 * i, w);
    }
  } else if (vec) {
    uint4* q = reinterpret_cast<uint4*>(p);
#pragma unroll
    for (int i = 0; i < COUNT / 8; ++i) {
      uint4 o;
      o.x = pack_bf16x2(v[8 * i + 0], v[8 * i + 1]);
      o.y = pack_bf16x2(v[8 * i + 2], v[8 * i + 3]);
      o.z = pack_bf16x2(v[8 * i + 4], v[8 * i + 5]);
      o.w = pack_bf16x2(v[8 * i + 6], v[8 * i + 7]);
      q[i] = o;
    }
  } else {
#pragma unroll
    for (int j = 0; j < COUNT; ++j)
      if (j < nvalid) p[j] = __float2bfloat16_rn(v[j]);
  }
}

template <int COUNT>
__device__ __forceinline__ void store_f32_row(float* p, bool vec, int nvalid, const float (&v)[COUNT]) {
  if (vec && (reinterpret_cast<uintptr_t>(p) & 31) == 0) {
#pragma unroll
    for (int i = 0; i < COUNT / 8; ++i) {
      uint32_t w[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) w[j] = __float_as_uint(v[8 * i + j]);
      st_global_v8(p + 8 * i, w);
    }
  } else if (vec) {
    float4* q = reinterpret_cast<float4*>(p);
#pragma unroll
    for (int i = 0; i < COUNT / 4; ++i) q[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
  } else {
#pragma unroll
    for (int j = 0; j < COUNT; ++j)
      if (j < nvalid) p[j] = v[j];
  }
}

// Epilogue inputs that do not depend on the accumulator (residual row, activation-gradient input) are requested before
// the epilogue waits on its TMEM load, so their L2 / HBM latency overlaps the TMEM read instead of following it.
struct EpiPrefetch {
  uint4 res[4];
  uint4 aux[4];
};
__device__ __forceinline__ void epi_load_row(const __nv_bfloat16* p, uint4 (&dst)[4]) {      // 32 bf16 = 64 B
  if ((reinterpret_cast<uintptr_t>(p) & 31) == 0) {
    uint32_t w0[8], w1[8];
    ld_global_nc_v8(p, w0);
    ld_global_nc_v8(p + 16, w1);
    dst[0] = make_uint4(w0[0], w0[1], w0[2], w0[3]);
    dst[1] = make_uint4(w0[4], w0[5], w0[6], w0[7]);
    dst[2] = make_uint4(w1[0], w1[1], w1[2], w1[3]);
    dst[3] = make_uint4(w1[4], w1[5], w1[6], w1[7]);
  } else {
    const uint4* q = reinterpret_cast<const uint4*>(p);
#pragma unroll
    for (int i = 0; i < 4; ++i) dst[i] = __ldg(q + i);
  }
}
__device__ __forceinline__ bool epi_res_vec(const GemmDev& g, int row, int col0) {
  return g.residual != nullptr && row < g.m && col0 + 32 <= g.n && (g.ldr & 7) == 0 && g.epilogue != JL_EPI_GLU;
}
__device__ __forceinline__ bool epi_aux_vec(const GemmDev& g, int row, int col0) {
  return (g.epilogue == JL_EPI_GELU_BWD || g.epilogue == JL_EPI_RELU_BWD || g.epilogue == JL_EPI_MUL_AUX) && row < g.m && col0 + 32 <= g.n && (g.ldaux & 7) == 0;
}
__device__ __forceinline__ void epi_prefetch(const GemmDev& g, int row, int col0, EpiPrefetch& pf) {
  if (epi_res_vec(g, row, col0)) epi_load_row(g.residual + static_cast<int64_t>(row) * g.ldr + col0, pf.res);
  if (epi_aux_vec(g, row, col0)) epi_load_row(g.aux + static_cast<int64_t>(row) * g.ldaux + col0, pf.aux);
}
__device__ __forceinline__ void unpack_row32(const uint4 (&q)[4], float (&out)[32]) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 f0 = unpack_bf16x2(q[i].x), f1 = unpack_bf16x2(q[i].y), f2 = unpack_bf16x2(q[i].z), f3 = unpack_bf16x2(q[i].w);
    out[8 * i + 0] = f0.x; out[8 * i + 1] = f0.y; out[8 * i + 2] = f1.x; out[8 * i + 3] = f1.y;
    out[8 * i + 4] = f2.x; out[8 * i + 5] = f2.y; out[8 * i + 6] = f3.x; out[8 * i + 7] = f3.y;
  }
}

__device__ __forceinline__ void gemm_epilogue_row32(const GemmDev& g, int row, int col0, float (&acc)[32], const EpiPrefetch& pf) {
  if (row >= g.m || col0 >= g.n) return;
  const int nvalid = min(32, g.n - col0);
  const bool full = (nvalid == 32);

  // alpha, bias
  if (g.bias != nullptr) {
    if (full) {
      const float4* bq = reinterpret_cast<const float4*>(g.bias + col0);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        float4 b = __ldg(bq + i);
        acc[4 * i + 0] = fmaf(acc[4 * i + 0], g.alpha, b.x);
        acc[4 * i + 1] = fmaf(acc[4 * i + 1], g.alpha, b.y);
        acc[4 * i + 2] = fmaf(acc[4 * i + 2], g.alpha, b.z);
        acc[4 * i + 3] = fmaf(acc[4 * i + 3], g.alpha, b.w);
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j) acc[j] = fmaf(acc[j], g.alpha, (j < nvalid) ? __ldg(g.bias + col0 + j) : 0.0f);
    }
  } else {
#pragma unroll
    for (int j = 0; j < 32; ++j) acc[j] *= g.alpha;
  }

  bool zero_row = false;
  if (g.row_lengths != nullptr) {
    const int b = row / g.rows_per_seq;
    const int t = row - b * g.rows_per_seq;
    zero_row = (t >= __ldg(g.row_lengths + b));
  }

  if (g.epilogue == JL_EPI_ARGMAX) {
    // f1 (inference head): no output matrix — (max, first argmax) of this row over the chunk's columns, chunk-major so that the
    // 32 rows of a warp store 128 contiguous bytes
    float best = -CUDART_INF_F;
    int besti = 0x7fffffff;
#pragma unroll
    for (int j = 0; j < 32; ++j)
      if (j < nvalid && acc[j] > best) { best = acc[j]; besti = col0 + j; }      // strict >: the first maximum wins
    const int64_t chunk = col0 >> 5;
    reinterpret_cast<float*>(g.c)[chunk * g.ldc + row] = best;
    reinterpret_cast<int32_t*>(g.aux_out)[chunk * g.ldaux_out + row] = besti;
    return;
  }

  if (g.epilogue == JL_EPI_GLU) {
    // interleaved (value, gate) column pairs → 16 outputs
    float o[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const float val = acc[2 * j], gate = acc[2 * j + 1];
      o[j] = zero_row ? 0.0f : val * sigmoid_fast(gate);
    }
    const int ocol = col0 >> 1;
    const int on = g.n >> 1;
    const int ovalid = min(16, on - ocol);
    if (g.out_dtype == JL_DT_BF16) {
      __nv_bfloat16* cp = reinterpret_cast<__nv_bfloat16*>(g.c) + static_cast<int64_t>(row) * g.ldc + ocol;
      const bool vec = (ovalid == 16) && ((g.ldc & 7) == 0);
      store_bf16_row<16>(cp, vec, ovalid, o);
    } else {
      float* cp = reinterpret_cast<float*>(g.c) + static_cast<int64_t>(row) * g.ldc + ocol;
      const bool vec = (ovalid == 16) && ((g.ldc & 3) == 0);
      store_f32_row<16>(cp, vec, ovalid, o);
    }
    return;
  }

  if (g.epilogue == JL_EPI_GELU) {
    if (g.aux_out != nullptr) {
      __nv_bfloat16* ap = g.aux_out + static_cast<int64_t>(row) * g.ldaux_out + col0;
      store_bf16_row<32>(ap, full && ((g.ldaux_out & 7) == 0), nvalid, acc);
    }
#pragma unroll
    for (int j = 0; j < 32; ++j) acc[j] = gelu_erf(acc[j]);
  } else if (g.epilogue == JL_EPI_GELU_DGELU) {
    // value to C, derivative to aux_out: the backward GEMM then only multiplies (JL_EPI_MUL_AUX) — the erf is evaluated once
    float d[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) gelu_erf_both(acc[j], acc[j], d[j]);
    __nv_bfloat16* ap = g.aux_out + static_cast<int64_t>(row) * g.ldaux_out + col0;
    store_bf16_row<32>(ap, full && ((g.ldaux_out & 7) == 0), nvalid, d);
  } else if (g.epilogue == JL_EPI_RELU) {
#pragma unroll
    for (int j = 0; j < 32; ++j) acc[j] = fmaxf(acc[j], 0.0f);
  } else if (g.epilogue == JL_EPI_GELU_BWD || g.epilogue == JL_EPI_RELU_BWD || g.epilogue == JL_EPI_MUL_AUX) {
    float a[32];
    if (epi_aux_vec(g, row, col0)) unpack_row32(pf.aux, a);
    else load_bf16_row32(g.aux + static_cast<int64_t>(row) * g.ldaux + col0, false, nvalid, a);
    if (g.epilogue == JL_EPI_GELU_BWD) {
#pragma unroll
      for (int j = 0; j < 32; ++j) acc[j] *= gelu_erf_grad(a[j]);
    } else if (g.epilogue == JL_EPI_MUL_AUX) {
#pragma unroll
      for (int j = 0; j < 32; ++j) acc[j] *= a[j];
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j) acc[j] = (a[j] > 0.0f) ? acc[j] : 0.0f;
    }
  }

  if (g.residual != nullptr) {
    float r[32];
    if (epi_res_vec(g, row, col0)) unpack_row32(pf.res, r);
    else load_bf16_row32(g.residual + static_cast<int64_t>(row) * g.ldr + col0, false, nvalid, r);
#pragma unroll
    for (int j = 0; j < 32; ++j) acc[j] += r[j];
  }
  if (zero_row) {
#pragma unroll
    for (int j = 0; j < 32; ++j) acc[j] = 0.0f;
  }

  if (g.out_dtype == JL_DT_BF16) {
    __nv_bfloat16* cp = reinterpret_cast<__nv_bfloat16*>(g.c) + static_cast<int64_t>(row) * g.ldc + col0;
    store_bf16_row<32>(cp, full && ((g.ldc & 7) == 0), nvalid, acc);
  } else {
    float* cp = reinterpret_cast<float*>(g.c) + static_cast<int64_t>(row) * g.ldc + col0;
    store_f32_row<32>(cp, full && ((g.ldc & 3) == 0), nvalid, acc);
  }
}

// ------------------------------------------------------------------------------------------------
// The tcgen05 kernel
// ------------------------------------------------------------------------------------------------
template <int BN, int STAGES>
struct GemmSmem {
  static constexpr int B_BYTES = BN * GEMM_BK * 2;
  static constexpr int STAGE_BYTES = GEMM_A_BYTES + B_BYTES;
  static constexpr int BAR_OFFSET = STAGES * STAGE_BYTES;
  static constexpr int TOTAL = BAR_OFFSET + (2 * STAGES + 4) * 8 + 16 + 1024;   // + alignment slack
};

// One schedulable piece of the single-CTA kernel: (output tile, K split) or (tail tile, column slice).
struct CtaUnit {
  int tile, split, n0, bn;
};
template <int BN>
__device__ __forceinline__ CtaUnit cta_unit(const GemmDev& g, int u) {
  CtaUnit r;
  int sub = 0;
  r.bn = BN;
  if (g.nsplit > 1) {
    r.split = 0;
    if (u < g.nsplit_first) {
      r.tile = u;
    } else {
      const int v = u - g.nsplit_first;
      r.tile = g.nsplit_first + v / g.nsplit;
      sub = v - (r.tile - g.nsplit_first) * g.nsplit;
      if (g.nsplit == 3) {                 // 256 columns as 96 + 96 + 64 (UMMA N is a multiple of 16, the epilogue works in chunks of 32)
        r.bn = (sub < 2) ? 96 : 64;
        r.n0 = (r.tile % g.num_n_tiles) * BN + sub * 96;
        return r;
      }
      r.bn = BN / g.nsplit;
    }
  } else {
    r.tile = u / g.split_k;
    r.split = u - r.tile * g.split_k;
  }
  r.n0 = (r.tile % g.num_n_tiles) * BN + sub * r.bn;
  return r;
}

template <int BN, int STAGES, bool A_MN, bool B_MN>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b, const __grid_constant__ CUtensorMap tma_b_sub,
                    const __grid_constant__ CUtensorMap tma_b_sub2, const GemmDev g) {
  jl::pdl_launch_dependents();
  using L = GemmSmem<BN, STAGES>;
  constexpr uint32_t TMEM_COLS = (2 * BN < 32) ? 32 : 2 * BN;
  static_assert(BN == 32 || BN == 64 || BN == 128 || BN == 256, "BN must be a power of two in [32, 256]");
  static_assert(!B_MN || BN >= 64, "MN-major B needs 64-wide swizzle blocks");

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* s_a = smem;
  uint8_t* s_b = smem + STAGES * GEMM_A_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::BAR_OFFSET);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tfull_bar = empty_bar + STAGES;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int out_tiles = g.num_m_tiles * g.num_n_tiles;
  // work units: (output tile, K split), or whole tiles followed by the column slices of the tail tiles
  const int num_tiles = (g.nsplit > 1) ? g.nsplit_first + (out_tiles - g.nsplit_first) * g.nsplit : out_tiles * g.split_k;
  const int num_kb_total = (g.k + GEMM_BK - 1) / GEMM_BK;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&tma_a);
    ptx::prefetch_tensormap(&tma_b);
    if (g.nsplit > 1) ptx::prefetch_tensormap(&tma_b_sub);
    if (g.nsplit == 3) ptx::prefetch_tensormap(&tma_b_sub2);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      ptx::mbar_init(full_bar + s, 1);
      ptx::mbar_init(empty_bar + s, 1);
    }
    for (int s = 0; s < 2; ++s) {
      ptx::mbar_init(tfull_bar + s, 1);
      ptx::mbar_init(tempty_bar + s, GEMM_EPI_WARPS);
    }
    ptx::fence_barrier_init();
  }
  if (warp == 2) {
    ptx::tmem_alloc(tmem_slot, TMEM_COLS);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  jl::pdl_wait();      // everything above overlapped the previous kernel's tail; its results are visible from here on

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int unit = blockIdx.x; unit < num_tiles; unit += gridDim.x) {
        const CtaUnit un = cta_unit<BN>(g, unit);
        const int tile = un.tile, split = un.split;
        const int m0 = (tile / g.num_n_tiles) * GEMM_BM;
        const int n0 = un.n0;
        const bool sliced = un.bn != BN;                       // column slice of a tail tile (K-major B only)
        const int kb_begin = split * g.kb_per_split;
        const int kb_end = min(num_kb_total, kb_begin + g.kb_per_split);
        for (int kb = kb_begin; kb < kb_end; ++kb) {
          ptx::mbar_wait(empty_bar + stage, phase ^ 1u);
          ptx::mbar_expect_tx(full_bar + stage, sliced ? GEMM_A_BYTES + un.bn * GEMM_BK * 2 : L::STAGE_BYTES);
          uint8_t* a_dst = s_a + stage * GEMM_A_BYTES;
          uint8_t* b_dst = s_b + stage * L::B_BYTES;
          const int k0 = kb * GEMM_BK;
          if (A_MN) {
#pragma unroll
            for (int i = 0; i < GEMM_BM / 64; ++i) ptx::tma_load_2d(a_dst + i * 8192, &tma_a, full_bar + stage, m0 + 64 * i, k0);
          } else {
            ptx::tma_load_2d(a_dst, &tma_a, full_bar + stage, k0, m0);
          }
          if (B_MN) {
#pragma unroll
            for (int i = 0; i < BN / 64; ++i) ptx::tma_load_2d(b_dst + i * 8192, &tma_b, full_bar + stage, n0 + 64 * i, k0);
          } else {
            ptx::tma_load_2d(b_dst, !sliced ? &tma_b : ((g.nsplit == 3 && un.bn == 64) ? &tma_b_sub2 : &tma_b_sub), full_bar + stage, k0, n0);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
      }
      jl::pdl_trigger_late();      // all operand loads requested: the next kernel may start launching
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc_full = ptx::make_idesc_bf16_f32(GEMM_BM, BN) | (A_MN ? (1u << 15) : 0u) | (B_MN ? (1u << 16) : 0u);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int unit = blockIdx.x; unit < num_tiles; unit += gridDim.x, ++it) {
        const int as = it & 1;
        const uint32_t aphase = (it >> 1) & 1;
        const CtaUnit un = cta_unit<BN>(g, unit);
        const int split = un.split;
        // the N field (bits 17-22, N >> 3) of the instruction descriptor is the only thing a column slice changes
        const uint32_t idesc = (un.bn == BN) ? idesc_full : ((idesc_full & ~(0x3Fu << 17)) | (static_cast<uint32_t>(un.bn >> 3) << 17));
        const int num_kb = min(num_kb_total, (split + 1) * g.kb_per_split) - split * g.kb_per_split;
        ptx::mbar_wait(tempty_bar + as, aphase ^ 1u);
        ptx::tc_fence_after();
        const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(as * BN);
        for (int kb = 0; kb < num_kb; ++kb) {
          ptx::mbar_wait(full_bar + stage, phase);
          ptx::tc_fence_after();
          const uint32_t a_addr = ptx::smem_u32(s_a + stage * GEMM_A_BYTES);
          const uint32_t b_addr = ptx::smem_u32(s_b + stage * L::B_BYTES);
#pragma unroll
          for (int k = 0; k < GEMM_BK / 16; ++k) {
            // K-major: 32 B along the 128 B swizzled row per 16-element k-step; 8-row groups 1024 B apart.
            // MN-major: 16 k-rows (2 × 1024 B swizzle atoms) per k-step; 64-wide MN blocks 8192 B apart.
            const uint64_t da = A_MN ? ptx::make_sw128_desc(a_addr + k * 2048, 8192, 1024) : ptx::make_sw128_desc(a_addr + k * 32, 16, 1024);
            const uint64_t db = B_MN ? ptx::make_sw128_desc(b_addr + k * 2048, 8192, 1024) : ptx::make_sw128_desc(b_addr + k * 32, 16, 1024);
            ptx::umma_bf16(d_tmem, da, db, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          ptx::umma_commit(empty_bar + stage);   // smem slot free once these MMAs have read it
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
        ptx::umma_commit(tfull_bar + as);         // accumulator complete
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue =====================
    const int quad = warp & 3;                    // TMEM lane quadrant this warp may read
    const int half = (warp - 4) >> 2;             // which 32-column chunks: c ≡ half (mod 3)
    int it = 0;
    for (int unit = blockIdx.x; unit < num_tiles; unit += gridDim.x, ++it) {
      const int as = it & 1;
      const uint32_t aphase = (it >> 1) & 1;
      const CtaUnit un = cta_unit<BN>(g, unit);
      const int tile = un.tile, split = un.split;
      const int m0 = (tile / g.num_n_tiles) * GEMM_BM;
      const int n0 = un.n0;
      const int nchunks = un.bn / 32;
      ptx::mbar_wait(tfull_bar + as, aphase);
      ptx::tc_fence_after();
      const int row = m0 + quad * 32 + lane;
#pragma unroll 1
      for (int c = half; c < nchunks; c += GEMM_EPI_WARPS / 4) {
        uint32_t v[32];
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + static_cast<uint32_t>(as * BN + c * 32);
        ptx::tmem_ld_32x32(taddr, v);
        EpiPrefetch pf;
        if (g.split_k == 1) epi_prefetch(g, row, n0 + c * 32, pf);
        ptx::tmem_ld_wait();
        if (c + GEMM_EPI_WARPS / 4 >= nchunks) {
          // this warp's last chunk is in registers: hand the accumulator stage back BEFORE the epilogue math and stores (the
          // release semantics of the arrive would otherwise wait for this tile's global stores to be acknowledged)
          ptx::tc_fence_before();
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive(tempty_bar + as);
        }
        float acc[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) acc[j] = __uint_as_float(v[j]);
        if (g.split_k > 1) {
          const int col0 = n0 + c * 32;
          if (row < g.m && col0 < g.n)
            store_f32_row<32>(g.ws + (static_cast<int64_t>(split) * g.m + row) * g.ldw + col0, col0 + 32 <= g.ldw, g.n - col0, acc);
        } else {
          gemm_epilogue_row32(g, row, n0 + c * 32, acc, pf);
        }
      }
      if (half >= nchunks) {          // a warp without a chunk of this tile still owes its arrival
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(tempty_bar + as);
      }
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, TMEM_COLS);
  }
}


// ------------------------------------------------------------------------------------------------
// CTA-pair kernel (tcgen05 cta_group::2): one 256 × BN output tile per cluster of two CTAs.  Each CTA stages its own
// 128 rows of A and half (BN/2 rows) of B per k-block, the leader CTA's single MMA thread issues
// tcgen05.mma.cta_group::2 (M = 256) which reads B from both CTAs' shared memory, and each CTA drains its own 128
// accumulator lanes.  Per k-block a pair moves (256 + BN) · 128 B for 256 · BN · 64 MACs — 128 FLOP/B at BN = 256
// against 85 for the single-CTA 128 × 256 tile — which is what lifts the L2 → SM operand traffic (≈ 6.3 kB/clk chip-wide)
// out of the way of the tensor pipe.
// ------------------------------------------------------------------------------------------------
// K range of a tail tile (pair kernel): publish the raw fp32 partial, count the arrival, and — in the warp that arrives last for
// its slice — add the partials in range order and apply the epilogue.  Only compiled into the TAIL instantiation of the pair
// kernel so that its registers stay out of the default one (nvcc 12.9 crashes on a __noinline__ function holding tcgen05.ld).
template <int BN>
__device__ __forceinline__ void pair_tail_unit(const GemmDev& g, const PairUnit& un, uint32_t tmem_base, int as, int quad, int half, int lane, int warp,
                                            int rank, uint32_t leader_tempty, int row, int n0) {
  const int tile = un.tile;
  // ---- K range of a tail tile: publish the raw partial, then the last range to arrive (per warp slice) finishes the tile
  const int tidx = tile - g.tail_first;
  const int64_t slice_rows = (static_cast<int64_t>(tidx) * 2 + rank) * GEMM_BM + quad * 32 + lane;       // row slot of this thread
  const int64_t range_stride = g.tail_stride;                                                            // floats per K range
  float* mine = g.tail_ws + static_cast<int64_t>(un.sidx) * range_stride + slice_rows * BN;
#pragma unroll 1
  for (int c = half; c < BN / 32; c += GEMM_EPI_WARPS / 4) {
    uint32_t v[32];
    ptx::tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + static_cast<uint32_t>(as * BN + c * 32), v);
    ptx::tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      uint32_t w[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) w[j] = v[8 * i + j];
      st_global_v8(mine + c * 32 + 8 * i, w);
    }
  }
  ptx::tc_fence_before();
  __syncwarp();
  if (lane == 0) ptx::mbar_arrive_cluster(leader_tempty + static_cast<uint32_t>(as * 8));   // accumulator stage free
  __threadfence();                                                                            // partial visible before it is counted
  __syncwarp();
  int32_t* cnt = g.tail_cnt + (tidx * 2 + rank) * GEMM_EPI_WARPS + (warp - 4);
  int arrived = 0;
  if (lane == 0) arrived = atomicAdd(cnt, 1);
  arrived = __shfl_sync(0xffffffffu, arrived, 0);
  if (arrived == g.tail_split - 1) {
    __threadfence();
    const float* base = g.tail_ws + slice_rows * BN;
#pragma unroll 1
    for (int c = half; c < BN / 32; c += GEMM_EPI_WARPS / 4) {
      EpiPrefetch pf;
      epi_prefetch(g, row, n0 + c * 32, pf);
      float acc[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) acc[j] = 0.0f;
      for (int sp = 0; sp < g.tail_split; ++sp) {            // fixed order → bit-reproducible
        const float4* src = reinterpret_cast<const float4*>(base + static_cast<int64_t>(sp) * range_stride + c * 32);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float4 q = __ldcg(src + i);                  // L2: written by other SMs during this launch
          acc[4 * i + 0] += q.x; acc[4 * i + 1] += q.y; acc[4 * i + 2] += q.z; acc[4 * i + 3] += q.w;
        }
      }
      gemm_epilogue_row32(g, row, n0 + c * 32, acc, pf);
    }
    __syncwarp();
    if (lane == 0) *cnt = 0;                                 // leave the counters zero for the next launch
  }
}

template <int BN, int STAGES>
struct GemmSmem2 {
  static constexpr int BNH = BN / 2;
  static constexpr int B_BYTES = BNH * GEMM_BK * 2;
  static constexpr int STAGE_BYTES = GEMM_A_BYTES + B_BYTES;
  static constexpr int BAR_OFFSET = STAGES * STAGE_BYTES;
  static constexpr int TOTAL = BAR_OFFSET + (2 * STAGES + 4) * 8 + 16 + 1024;
};

// PAIRS = 2: a cluster of four CTAs = two MMA pairs that share the same N tile (rows m0 … m0+511).  Each CTA fetches only a
// quarter of the B tile and TMA-multicasts it to the CTA at the same position of the other pair, so a k-block moves
// 4·16 KB of A + BN·128 B of B for 512 × BN × 64 MACs (175 FLOP/B at BN = 256 instead of 128).
template <int BN, int STAGES, bool A_MN, bool B_MN, int PAIRS, bool TAIL = false>
__global__ void __cluster_dims__(2 * PAIRS, 1, 1) __launch_bounds__(GEMM_THREADS, 1)
gemm_tcgen05_2cta_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b, const GemmDev g) {
  jl::pdl_launch_dependents();
  using L = GemmSmem2<BN, STAGES>;
  constexpr int BNH = BN / 2;
  constexpr uint32_t TMEM_COLS = (2 * BN <= 256) ? 256 : 512;
  static_assert(BN == 128 || BN == 192 || BN == 256, "pair tile N must be 128, 192 or 256");
  constexpr int BQ = BNH / PAIRS;               // B rows (K-major) / columns (MN-major) this CTA fetches itself
  static_assert(!B_MN || (BQ % 64) == 0, "MN-major B needs 64-wide swizzle blocks per fetched slice");
  static_assert(PAIRS == 1 || PAIRS == 2, "one or two MMA pairs per cluster");
  static_assert((BQ * 128) % 1024 == 0, "fetched B slice must keep the 1024-byte swizzle atom alignment");

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* s_a = smem;
  uint8_t* s_b = smem + STAGES * GEMM_A_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::BAR_OFFSET);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tfull_bar = empty_bar + STAGES;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t crank = ptx::cluster_ctarank();                // rank in the cluster
  const uint32_t rank = crank & 1u;                             // position in the MMA pair
  const uint32_t pic = crank >> 1;                              // pair index in the cluster
  const uint32_t leader_rank = crank & ~1u;
  const bool leader = (rank == 0);
  const int pair = blockIdx.x / (2 * PAIRS);                    // cluster index
  const int num_pairs = gridDim.x / (2 * PAIRS);
  const int num_tiles = g.num_m_tiles * g.num_n_tiles;          // (256·PAIRS) × BN tiles
  const int num_kb = (g.k + GEMM_BK - 1) / GEMM_BK;
  const int num_units = pair_num_units(g, num_tiles);           // whole tiles + K ranges of the tail tiles

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&tma_a);
    ptx::prefetch_tensormap(&tma_b);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      ptx::mbar_init(full_bar + s, 1);       // leader's: one arrive.expect_tx covering both CTAs' bytes
      ptx::mbar_init(empty_bar + s, PAIRS);  // each CTA's own: multicast commits from every pair's MMA thread
    }
    for (int s = 0; s < 2; ++s) {
      ptx::mbar_init(tfull_bar + s, 1);                       // each CTA's own: multicast commit
      ptx::mbar_init(tempty_bar + s, 2 * GEMM_EPI_WARPS);     // leader's: epilogue warps of both CTAs
    }
    ptx::fence_barrier_init();
  }
  if (warp == 2) {
    ptx::tmem_alloc_2sm(tmem_slot, TMEM_COLS);
    ptx::tmem_relinquish_2sm();
  }
  ptx::tc_fence_before();
  ptx::cluster_sync_all();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  jl::pdl_wait();      // everything above overlapped the previous kernel's tail; its results are visible from here on

  if (warp == 0) {
    // ===================== TMA producer (both CTAs) =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int u = pair; u < num_units; u += num_pairs) {
        const PairUnit un = pair_unit(g, u, num_kb);
        const int tile = un.tile;
        const int m0 = (tile / g.num_n_tiles) * (256 * PAIRS) + static_cast<int>(pic) * 256 + static_cast<int>(rank) * GEMM_BM;
        const int n0 = (tile % g.num_n_tiles) * BN + static_cast<int>(rank) * BNH + static_cast<int>(pic) * BQ * (PAIRS - 1);
        const uint16_t bmask = static_cast<uint16_t>((1u << rank) | (1u << (rank + 2)));   // same pair position in both pairs
        for (int kb = un.kb0; kb < un.kb1; ++kb) {
          ptx::mbar_wait(empty_bar + stage, phase ^ 1u);
          if (leader) ptx::mbar_expect_tx(full_bar + stage, 2 * L::STAGE_BYTES);
          const uint32_t bar = ptx::mapa_shared(ptx::smem_u32(full_bar + stage), leader_rank);
          uint8_t* a_dst = s_a + stage * GEMM_A_BYTES;
          uint8_t* b_dst = s_b + stage * L::B_BYTES + (PAIRS - 1) * static_cast<int>(pic) * BQ * 128;
          const int k0 = kb * GEMM_BK;
          if (A_MN) {
#pragma unroll
            for (int i = 0; i < GEMM_BM / 64; ++i) ptx::tma_load_2d_2sm(a_dst + i * 8192, &tma_a, bar, m0 + 64 * i, k0);
          } else {
            ptx::tma_load_2d_2sm(a_dst, &tma_a, bar, k0, m0);
          }
          if (PAIRS == 1) {
            if (B_MN) {
#pragma unroll
              for (int i = 0; i < BQ / 64; ++i) ptx::tma_load_2d_2sm(b_dst + i * 8192, &tma_b, bar, n0 + 64 * i, k0);
            } else {
              ptx::tma_load_2d_2sm(b_dst, &tma_b, bar, k0, n0);
            }
          } else {
            if (B_MN) {
#pragma unroll
              for (int i = 0; i < BQ / 64; ++i) ptx::tma_load_2d_2sm_mcast(b_dst + i * 8192, &tma_b, bar, n0 + 64 * i, k0, bmask);
            } else {
              ptx::tma_load_2d_2sm_mcast(b_dst, &tma_b, bar, k0, n0, bmask);
            }
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
      }
      jl::pdl_trigger_late();      // all operand loads requested: the next kernel may start launching
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only) =====================
    if (leader && lane == 0) {
      constexpr uint32_t idesc = ptx::make_idesc_bf16_f32(256, BN) | (A_MN ? (1u << 15) : 0u) | (B_MN ? (1u << 16) : 0u);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int u = pair; u < num_units; u += num_pairs, ++it) {
        const PairUnit un = pair_unit(g, u, num_kb);
        const int as = it & 1;
        const uint32_t aphase = (it >> 1) & 1;
        ptx::mbar_wait(tempty_bar + as, aphase ^ 1u);
        ptx::tc_fence_after();
        const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(as * BN);
        for (int kb = un.kb0; kb < un.kb1; ++kb) {
          ptx::mbar_wait(full_bar + stage, phase);
          ptx::tc_fence_after();
          const uint32_t a_addr = ptx::smem_u32(s_a + stage * GEMM_A_BYTES);
          const uint32_t b_addr = ptx::smem_u32(s_b + stage * L::B_BYTES);
#pragma unroll
          for (int k = 0; k < GEMM_BK / 16; ++k) {
            const uint64_t da = A_MN ? ptx::make_sw128_desc(a_addr + k * 2048, 8192, 1024) : ptx::make_sw128_desc(a_addr + k * 32, 16, 1024);
            const uint64_t db = B_MN ? ptx::make_sw128_desc(b_addr + k * 2048, 8192, 1024) : ptx::make_sw128_desc(b_addr + k * 32, 16, 1024);
            ptx::umma_bf16_2sm(d_tmem, da, db, idesc, (kb != un.kb0 || k != 0) ? 1u : 0u);
          }
          ptx::umma_commit_2sm(empty_bar + stage, PAIRS == 1 ? 0x3 : 0xF);   // every CTA that writes into these slots
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
        ptx::umma_commit_2sm(tfull_bar + as, static_cast<uint16_t>(0x3u << (2 * pic)));   // accumulators complete in both CTAs of the pair
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue (both CTAs, own 128 accumulator lanes) =====================
    const int quad = warp & 3;
    const int half = (warp - 4) >> 2;
    const uint32_t leader_tempty = ptx::mapa_shared(ptx::smem_u32(tempty_bar), leader_rank);
    int it = 0;
    for (int u = pair; u < num_units; u += num_pairs, ++it) {
      const PairUnit un = pair_unit(g, u, num_kb);
      const int tile = un.tile;
      const int as = it & 1;
      const uint32_t aphase = (it >> 1) & 1;
      const int m0 = (tile / g.num_n_tiles) * (256 * PAIRS) + static_cast<int>(pic) * 256 + static_cast<int>(rank) * GEMM_BM;
      const int n0 = (tile % g.num_n_tiles) * BN;
      ptx::mbar_wait(tfull_bar + as, aphase);
      ptx::tc_fence_after();
      const int row = m0 + quad * 32 + lane;
      if (TAIL && un.split) {
        pair_tail_unit<BN>(g, un, tmem_base, as, quad, half, lane, warp, static_cast<int>(rank), leader_tempty, row, n0);
        continue;
      }
#pragma unroll 1
      for (int c = half; c < BN / 32; c += GEMM_EPI_WARPS / 4) {
        uint32_t v[32];
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + static_cast<uint32_t>(as * BN + c * 32);
        ptx::tmem_ld_32x32(taddr, v);
        EpiPrefetch pf;
        epi_prefetch(g, row, n0 + c * 32, pf);
        ptx::tmem_ld_wait();
        if (c + GEMM_EPI_WARPS / 4 >= BN / 32) {
          // last chunk of this warp is in registers: free the accumulator stage before the epilogue math and the stores — the
          // cluster-scope release of the arrive otherwise waits until this tile's global stores are acknowledged (ncu: 10 % of
          // the samples of the GELU-gradient GEMM sat in that membar), which delays the MMAs of the tile after next
          ptx::tc_fence_before();
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive_cluster(leader_tempty + static_cast<uint32_t>(as * 8));
        }
        float acc[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) acc[j] = __uint_as_float(v[j]);
        gemm_epilogue_row32(g, row, n0 + c * 32, acc, pf);
      }
      if (half >= BN / 32) {
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive_cluster(leader_tempty + static_cast<uint32_t>(as * 8));
      }
    }
  }

  ptx::tc_fence_before();
  ptx::cluster_sync_all();        // the peer may still be arriving on / reading from this CTA's shared memory
  if (warp == 2) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc_2sm(tmem_base, TMEM_COLS);
  }
}

// C = alpha * Σ_s ws[s]  (fixed order → deterministic), fp32 or bf16 out.
__global__ void splitk_reduce_kernel(const float* __restrict__ ws, int64_t ldw, int splits, int m, int n, float alpha, void* __restrict__ c,
                                     int64_t ldc, int out_dtype) {
  jl::pdl_prologue();
  const int64_t total = static_cast<int64_t>(m) * n;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int r = static_cast<int>(i / n), col = static_cast<int>(i - static_cast<int64_t>(r) * n);
    float a = 0.0f;
    for (int sidx = 0; sidx < splits; ++sidx) a += ws[(static_cast<int64_t>(sidx) * m + r) * ldw + col];
    a *= alpha;
    if (out_dtype == JL_DT_BF16) reinterpret_cast<__nv_bfloat16*>(c)[static_cast<int64_t>(r) * ldc + col] = __float2bfloat16_rn(a);
    else reinterpret_cast<float*>(c)[static_cast<int64_t>(r) * ldc + col] = a;
  }
}

// ------------------------------------------------------------------------------------------------
// Test-only SIMT reference (fp32 accumulate, same epilogue code).
// ------------------------------------------------------------------------------------------------
struct GemmRefOperands {
  const __nv_bfloat16* a; int64_t lda; int a_mn;
  const __nv_bfloat16* b; int64_t ldb; int b_mn;
};

__global__ void gemm_ref_kernel(const GemmRefOperands o, const GemmDev g) {
  jl::pdl_prologue();
  const int chunks = (g.n + 31) / 32;
  const int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= static_cast<int64_t>(g.m) * chunks) return;
  const int row = static_cast<int>(idx / chunks);
  const int col0 = static_cast<int>(idx % chunks) * 32;
  float acc[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) acc[j] = 0.0f;
  for (int kk = 0; kk < g.k; ++kk) {
    const float av = __bfloat162float(o.a_mn ? o.a[static_cast<int64_t>(kk) * o.lda + row] : o.a[static_cast<int64_t>(row) * o.lda + kk]);
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      const int col = col0 + j;
      if (col < g.n) {
        const float bv = __bfloat162float(o.b_mn ? o.b[static_cast<int64_t>(kk) * o.ldb + col] : o.b[static_cast<int64_t>(col) * o.ldb + kk]);
        acc[j] = fmaf(av, bv, acc[j]);
      }
    }
  }
  EpiPrefetch pf;
  epi_prefetch(g, row, col0, pf);
  gemm_epilogue_row32(g, row, col0, acc, pf);
}

// ------------------------------------------------------------------------------------------------
// Host side
// ------------------------------------------------------------------------------------------------
static int make_map(CUtensorMap* map, const void* ptr, int64_t inner, int64_t outer, int64_t ld, int box_rows) {
  return make_tma_map_2d_bf16(map, ptr, inner, outer, ld, box_rows);
}

static int validate(const jl_gemm_params* p) {
  JL_REQUIRE(p != nullptr, JL_EINVAL, "gemm: null params");
  JL_REQUIRE(p->a && p->b && p->c, JL_EINVAL, "gemm: null operand pointer");
  JL_REQUIRE(p->m > 0 && p->n > 0 && p->k > 0, JL_EINVAL, "gemm: m, n, k must be positive (got %d, %d, %d)", p->m, p->n, p->k);
  JL_REQUIRE((reinterpret_cast<uintptr_t>(p->a) & 15) == 0 && (reinterpret_cast<uintptr_t>(p->b) & 15) == 0, JL_EINVAL,
             "gemm: A and B must be 16-byte aligned");
  JL_REQUIRE((p->lda & 7) == 0 && (p->ldb & 7) == 0, JL_EINVAL, "gemm: lda/ldb must be multiples of 8 elements (got %lld, %lld)",
             (long long)p->lda, (long long)p->ldb);
  JL_REQUIRE(p->a_layout == JL_LAYOUT_K || p->a_layout == JL_LAYOUT_MN, JL_EINVAL, "gemm: bad a_layout");
  JL_REQUIRE(p->b_layout == JL_LAYOUT_K || p->b_layout == JL_LAYOUT_MN, JL_EINVAL, "gemm: bad b_layout");
  // K-major A may have lda < K: rows that overlap in memory (sliding windows over a [T, C] activation — a Conv1d without im2col)
  JL_REQUIRE(p->lda >= (p->a_layout == JL_LAYOUT_K ? 8 : p->m), JL_EINVAL, "gemm: lda too small");
  JL_REQUIRE(p->ldb >= (p->b_layout == JL_LAYOUT_K ? p->k : p->n), JL_EINVAL, "gemm: ldb too small");
  JL_REQUIRE(p->epilogue >= JL_EPI_NONE && p->epilogue <= JL_EPI_ARGMAX, JL_EINVAL, "gemm: unknown epilogue %d", p->epilogue);
  if (p->epilogue == JL_EPI_ARGMAX) {
    JL_REQUIRE(p->aux_out != nullptr && p->ldc >= p->m && p->ldaux_out >= p->m, JL_EINVAL, "gemm: JL_EPI_ARGMAX needs aux_out and ldc, ldaux_out >= M");
    JL_REQUIRE(p->residual == nullptr && p->row_lengths == nullptr, JL_EINVAL, "gemm: JL_EPI_ARGMAX takes no residual / row_lengths");
  }
  JL_REQUIRE(p->out_dtype == JL_DT_BF16 || p->out_dtype == JL_DT_F32, JL_EINVAL, "gemm: unknown out_dtype %d", p->out_dtype);
  JL_REQUIRE((reinterpret_cast<uintptr_t>(p->c) & 15) == 0, JL_EINVAL, "gemm: C must be 16-byte aligned");
  if (p->epilogue == JL_EPI_GLU) JL_REQUIRE((p->n & 1) == 0, JL_EINVAL, "gemm: GLU needs an even N");
  if (p->epilogue == JL_EPI_GELU_BWD || p->epilogue == JL_EPI_RELU_BWD || p->epilogue == JL_EPI_MUL_AUX)
    JL_REQUIRE(p->aux != nullptr, JL_EINVAL, "gemm: activation-gradient epilogue needs aux");
  if (p->epilogue == JL_EPI_GELU_DGELU) JL_REQUIRE(p->aux_out != nullptr, JL_EINVAL, "gemm: JL_EPI_GELU_DGELU needs aux_out");
  if (p->bias) JL_REQUIRE((reinterpret_cast<uintptr_t>(p->bias) & 15) == 0, JL_EINVAL, "gemm: bias must be 16-byte aligned");
  if (p->residual) JL_REQUIRE((reinterpret_cast<uintptr_t>(p->residual) & 15) == 0, JL_EINVAL, "gemm: residual must be 16-byte aligned");
  if (p->aux) JL_REQUIRE((reinterpret_cast<uintptr_t>(p->aux) & 15) == 0, JL_EINVAL, "gemm: aux must be 16-byte aligned");
  if (p->aux_out) JL_REQUIRE((reinterpret_cast<uintptr_t>(p->aux_out) & 15) == 0, JL_EINVAL, "gemm: aux_out must be 16-byte aligned");
  if (p->row_lengths) JL_REQUIRE(p->rows_per_seq > 0, JL_EINVAL, "gemm: row_lengths needs rows_per_seq > 0");
  return JL_OK;
}

static int pick_bn(const jl_gemm_params* p);

// K splits for a plain product (no bias / activation / residual) whose output tiles cannot fill the SMs.
static int pick_split(const jl_gemm_params* p, int bn, int* kb_per_split) {
  const int num_kb = ceil_div(p->k, GEMM_BK);
  *kb_per_split = num_kb;
  if (p->epilogue != JL_EPI_NONE || p->bias || p->residual || p->row_lengths) return 1;
  const int tiles = ceil_div(p->m, GEMM_BM) * ceil_div(p->n, bn);
  const int sms = num_sms();
  if (tiles * 2 > sms || num_kb < 16) return 1;
  int want = sms / tiles;
  if (want > num_kb / 4) want = num_kb / 4;
  if (want < 2) return 1;
  const int per = ceil_div(num_kb, want);
  *kb_per_split = per;
  return ceil_div(num_kb, per);
}

static size_t splitk_ws_bytes(const jl_gemm_params* p) {
  int per = 0;
  const int s = pick_split(p, pick_bn(p), &per);
  if (s <= 1) return 0;
  const size_t ldw = (static_cast<size_t>(p->n) + 3) & ~static_cast<size_t>(3);
  return static_cast<size_t>(s) * p->m * ldw * sizeof(float);
}

// bit 0: K-range split of the pair kernel's tail wave (needs a workspace; measured slower, off); bit 1: column slices of the
// single-CTA kernel's tail wave (on); bit 2: never cut a 256-wide tail tile three ways (96 + 96 + 64) — on: the three-way cut is
// 1-2 % faster per GEMM timed alone (8000x768x3072: 37.0 → 36.8 us, 8000x768x768: 16.0 → 15.1) but 1.1 % SLOWER for the whole step
// (6.36 vs 6.29 ms, three A/B jobs): with 123 instead of 82 CTAs busy in the tail wave the weight-gradient branch finds fewer
// idle SMs and finishes later (without that branch the three-way cut wins: 6.04 vs 6.09 ms)
static std::atomic<int> g_gemm_tail{2 | 4};
static GemmDev to_dev(const jl_gemm_params* p, int bn) {
  GemmDev g;
  g.split_k = 1; g.kb_per_split = ceil_div(p->k, GEMM_BK); g.ws = nullptr; g.ldw = 0;
  g.nsplit_first = 0; g.nsplit = 1;
  g.tail_first = 0; g.tail_split = 1; g.tail_kb_per = g.kb_per_split; g.tail_stride = 0; g.tail_ws = nullptr; g.tail_cnt = nullptr;
  g.c = p->c; g.ldc = p->ldc;
  g.bias = p->bias;
  g.residual = reinterpret_cast<const __nv_bfloat16*>(p->residual); g.ldr = p->ldr;
  g.aux = reinterpret_cast<const __nv_bfloat16*>(p->aux); g.ldaux = p->ldaux;
  g.aux_out = reinterpret_cast<__nv_bfloat16*>(p->aux_out); g.ldaux_out = p->ldaux_out;
  g.row_lengths = p->row_lengths; g.rows_per_seq = p->rows_per_seq;
  g.m = p->m; g.n = p->n; g.k = p->k;
  g.epilogue = p->epilogue; g.out_dtype = p->out_dtype; g.alpha = p->alpha;
  g.num_m_tiles = ceil_div(p->m, GEMM_BM);
  g.num_n_tiles = ceil_div(p->n, bn);
  return g;
}

template <int BN, int STAGES, bool A_MN, bool B_MN>
static int launch_gemm(const jl_gemm_params* p, cudaStream_t stream) {
  using L = GemmSmem<BN, STAGES>;
  auto kern = gemm_tcgen05_kernel<BN, STAGES, A_MN, B_MN>;
  static thread_local int configured_dev = -1;
  int dev = 0;
  cudaGetDevice(&dev);
  if (configured_dev != dev) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL);
    JL_REQUIRE(e == cudaSuccess, JL_ECUDA, "gemm: cannot reserve %d B of shared memory: %s", L::TOTAL, cudaGetErrorString(e));
    configured_dev = dev;
  }
  CUtensorMap ma, mb;
  int rc;
  if (A_MN) rc = make_map(&ma, p->a, p->m, p->k, p->lda, 64);
  else rc = make_map(&ma, p->a, p->k, p->m, p->lda, GEMM_BM);
  if (rc != JL_OK) return rc;
  if (B_MN) rc = make_map(&mb, p->b, p->n, p->k, p->ldb, 64);
  else rc = make_map(&mb, p->b, p->k, p->n, p->ldb, BN);
  if (rc != JL_OK) return rc;
  GemmDev g = to_dev(p, BN);
  if (p->workspace != nullptr) {
    int per = 0;
    const int s = pick_split(p, BN, &per);
    const size_t need = splitk_ws_bytes(p);
    if (s > 1 && static_cast<size_t>(p->workspace_bytes) >= need && (reinterpret_cast<uintptr_t>(p->workspace) & 15) == 0) {
      g.split_k = s;
      g.kb_per_split = per;
      g.ws = reinterpret_cast<float*>(p->workspace);
      g.ldw = (static_cast<int64_t>(p->n) + 3) & ~static_cast<int64_t>(3);
    }
  }
  int units = g.num_m_tiles * g.num_n_tiles * g.split_k;
  CUtensorMap mbs = mb, mbs2 = mb;
  if (!B_MN && BN >= 128 && g.split_k == 1 && (g_gemm_tail.load() & 2)) {
    // the last, partial wave of tiles as column slices: [tail · nsplit <= SMs] → it takes 1 / nsplit of a tile time (3 slices of a
    // 256-wide tile: 96 + 96 + 64 columns → 0.375; bit 2 of the tail mode switches the three-way split off)
    const int sms = num_sms();
    const int first = (units / sms) * sms, tail = units - first;
    const bool three = BN == 256 && (g_gemm_tail.load() & 4) == 0;
    const int ns = (tail == 0) ? 1 : (BN >= 256 && tail * 4 <= sms) ? 4 : (three && tail * 3 <= sms) ? 3 : (tail * 2 <= sms ? 2 : 1);
    if (ns > 1) {
      rc = make_map(&mbs, p->b, p->k, p->n, p->ldb, ns == 3 ? 96 : BN / ns);
      if (rc == JL_OK && ns == 3) rc = make_map(&mbs2, p->b, p->k, p->n, p->ldb, 64);
      if (rc != JL_OK) return rc;
      g.nsplit_first = first;
      g.nsplit = ns;
      units = first + tail * ns;
    }
  }
  const int grid = units < num_sms() ? units : num_sms();
  jl::launch(kern, grid, GEMM_THREADS, L::TOTAL, stream, ma, mb, mbs, mbs2, g);
  JL_CHECK_LAUNCH("gemm_tcgen05");
  if (g.split_k > 1) {
    const int64_t total = static_cast<int64_t>(p->m) * p->n;
    int blocks = static_cast<int>((total + 255) / 256);
    if (blocks > 148 * 8) blocks = 148 * 8;
    jl::launch(splitk_reduce_kernel, blocks, 256, 0, stream, g.ws, g.ldw, g.split_k, p->m, p->n, p->alpha, p->c, p->ldc, p->out_dtype);
    JL_CHECK_LAUNCH("gemm_splitk_reduce");
  }
  return JL_OK;
}

template <int BN, int STAGES>
static int dispatch_layout(const jl_gemm_params* p, cudaStream_t s) {
  const bool amn = p->a_layout == JL_LAYOUT_MN, bmn = p->b_layout == JL_LAYOUT_MN;
  if constexpr (BN >= 64) {
    if (amn && bmn) return launch_gemm<BN, STAGES, true, true>(p, s);
    if (bmn) return launch_gemm<BN, STAGES, false, true>(p, s);
  }
  if (amn) return launch_gemm<BN, STAGES, true, false>(p, s);
  return launch_gemm<BN, STAGES, false, false>(p, s);
}

// Pick the N tile: fewest (waves × per-tile cost) over the candidates; per-tile cost ∝ BN plus a fixed part.
static std::atomic<int> g_gemm_bn{0};     // tuning hook: 0 = automatic N tile, else the N tile forced on the kernel the mode selects
static int pick_bn(const jl_gemm_params* p) {
  const bool bmn = p->b_layout == JL_LAYOUT_MN;
  const int forced = g_gemm_bn.load();
  if (forced == 256 || forced == 128 || forced == 64 || (forced == 32 && !bmn)) return forced;
  if (!bmn && p->n <= 32) return 32;
  if (p->n <= 64) return 64;
  if (p->k <= GEMM_BK && p->m >= 1024) return 64;      // one k-block: epilogue-bound, small tiles spread it over all SMs
  if (p->n <= 256 && p->m >= 1024) return 128;
  if (!bmn && (p->n % 256) == 0 && p->m >= 4096) return 256;   // measured: the 128 x 256 tile is the fastest on every encoder shape
  const int sms = num_sms();
  const int mt = ceil_div(p->m, GEMM_BM);
  int best = 64;
  long best_cost = -1;
  const int cands[3] = {256, 128, 64};
  for (int i = 0; i < 3; ++i) {
    const int bn = cands[i];
    if (bn > 64 && bn / 2 >= p->n) continue;
    const long tiles = static_cast<long>(mt) * ceil_div(p->n, bn);
    const long waves = (tiles + sms - 1) / sms;
    const long cost = waves * (bn + 32);
    if (best_cost < 0 || cost < best_cost) { best_cost = cost; best = bn; }
  }
  return best;
}


// Tail split of the pair kernel: the last, partial wave of 256 × bn tiles is cut into K ranges so that it fills the idle pairs.
constexpr size_t GEMM_TAIL_ZERO_BYTES = 8192;      // arrival counters at the head of the workspace (zero on entry, left zero)
struct TailPlan {
  int first = 0, split = 1, kb_per = 0, tail_tiles = 0;
  double waves = 0.0;       // full waves + the shortened tail wave
};
static TailPlan plan_tail(const jl_gemm_params* p, int bn, bool allow_split) {
  TailPlan t;
  const int pairs = num_sms() / 2;
  const int tiles = ceil_div(p->m, 256) * ceil_div(p->n, bn);
  const int num_kb = ceil_div(p->k, GEMM_BK);
  const int tail = tiles % pairs;
  t.first = tiles; t.kb_per = num_kb;
  t.waves = static_cast<double>(ceil_div(tiles, pairs));
  if (!allow_split || tail == 0 || (g_gemm_tail.load() & 1) == 0 || p->a_layout == JL_LAYOUT_MN) return t;
  int want = pairs / tail;
  if (want > 4) want = 4;
  if (want > num_kb / 4) want = num_kb / 4;
  if (want < 2) return t;
  const int per = ceil_div(num_kb, want);
  const int split = ceil_div(num_kb, per);
  if (split < 2 || static_cast<size_t>(tail) * 2 * GEMM_EPI_WARPS * sizeof(int32_t) > GEMM_TAIL_ZERO_BYTES) return t;
  t.first = tiles - tail; t.split = split; t.kb_per = per; t.tail_tiles = tail;
  t.waves = static_cast<double>(tiles / pairs) + static_cast<double>(per) / num_kb + 0.08;   // + fix-up traffic
  return t;
}
static size_t tail_ws_bytes(const TailPlan& t, int bn) {
  if (t.split <= 1) return 0;
  return GEMM_TAIL_ZERO_BYTES + static_cast<size_t>(t.split) * t.tail_tiles * 256 * bn * sizeof(float);
}

template <int BN, int STAGES, bool A_MN, bool B_MN, int PAIRS>
static int launch_gemm_2cta(const jl_gemm_params* p, cudaStream_t stream) {
  using L = GemmSmem2<BN, STAGES>;
  constexpr bool CAN_TAIL = (PAIRS == 1) && !A_MN;     // the tail-split instantiation exists for K-major A only
  auto kern = gemm_tcgen05_2cta_kernel<BN, STAGES, A_MN, B_MN, PAIRS, false>;
  auto kern_tail = gemm_tcgen05_2cta_kernel<BN, STAGES, A_MN, B_MN, PAIRS, CAN_TAIL>;
  static thread_local int configured_dev = -1;
  int dev = 0;
  cudaGetDevice(&dev);
  if (configured_dev != dev) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL);
    if (e == cudaSuccess && CAN_TAIL) e = cudaFuncSetAttribute(kern_tail, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL);
    JL_REQUIRE(e == cudaSuccess, JL_ECUDA, "gemm(2cta): cannot reserve %d B of shared memory: %s", L::TOTAL, cudaGetErrorString(e));
    configured_dev = dev;
  }
  CUtensorMap ma, mb;
  int rc;
  if (A_MN) rc = make_map(&ma, p->a, p->m, p->k, p->lda, 64);
  else rc = make_map(&ma, p->a, p->k, p->m, p->lda, GEMM_BM);
  if (rc != JL_OK) return rc;
  if (B_MN) rc = make_map(&mb, p->b, p->n, p->k, p->ldb, 64);
  else rc = make_map(&mb, p->b, p->k, p->n, p->ldb, BN / 2 / PAIRS);
  if (rc != JL_OK) return rc;
  GemmDev g = to_dev(p, BN);
  g.num_m_tiles = ceil_div(p->m, 256 * PAIRS);
  int units = g.num_m_tiles * g.num_n_tiles;
  g.tail_first = units;
  if (CAN_TAIL && p->workspace != nullptr && (reinterpret_cast<uintptr_t>(p->workspace) & 255) == 0) {
    const TailPlan t = plan_tail(p, BN, true);
    if (t.split > 1 && static_cast<size_t>(p->workspace_bytes) >= tail_ws_bytes(t, BN)) {
      g.tail_first = t.first; g.tail_split = t.split; g.tail_kb_per = t.kb_per;
      g.tail_stride = static_cast<int64_t>(t.tail_tiles) * 256 * BN;
      g.tail_cnt = reinterpret_cast<int32_t*>(p->workspace);
      g.tail_ws = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(p->workspace) + GEMM_TAIL_ZERO_BYTES);
      units = t.first + t.tail_tiles * t.split;
    }
  }
  const int tiles = units;
  const int pairs_max = num_sms() / (2 * PAIRS);
  const int pairs = tiles < pairs_max ? tiles : pairs_max;
  jl::launch(g.tail_split > 1 ? kern_tail : kern, 2 * PAIRS * pairs, GEMM_THREADS, L::TOTAL, stream, ma, mb, g);
  JL_CHECK_LAUNCH("gemm_tcgen05_2cta");
  return JL_OK;
}

template <int BN, int STAGES, int PAIRS>
static int dispatch_layout_2cta(const jl_gemm_params* p, cudaStream_t s) {
  const bool amn = p->a_layout == JL_LAYOUT_MN, bmn = p->b_layout == JL_LAYOUT_MN;
  if constexpr ((BN / 2 / PAIRS) % 64 == 0) {
    if (amn && bmn) return launch_gemm_2cta<BN, STAGES, true, true, PAIRS>(p, s);
    if (bmn) return launch_gemm_2cta<BN, STAGES, false, true, PAIRS>(p, s);
  }
  if (amn) return launch_gemm_2cta<BN, STAGES, true, false, PAIRS>(p, s);
  return launch_gemm_2cta<BN, STAGES, false, false, PAIRS>(p, s);
}

static std::atomic<int> g_gemm_mode{0};   // 0 = auto, 1 = single-CTA kernel only, 2 = CTA-pair kernel wherever legal, 3 = + B multicast across two pairs

// N tile of the CTA-pair kernel, or 0 when the product should run on the single-CTA kernel.  `assume_ws`: plan as if the caller
// will supply the workspace jl_gemm_workspace_bytes asks for.
static int pick_bn_2cta(const jl_gemm_params* p, bool assume_ws = false) {
  const int mode = g_gemm_mode.load();
  if (mode == 1) return 0;
  const int forced = g_gemm_bn.load();
  if (mode >= 2 && (forced == 128 || forced == 192 || forced == 256) && !(forced == 192 && p->b_layout == JL_LAYOUT_MN)) return forced;
  if (p->n < 128) return 0;
  if ((mode == 0) && (p->m < 1024 || static_cast<int64_t>(p->m) * p->n < 512 * 1024)) return 0;
  // Measured on B200 (profiles/r1g_gemm_tile_sweep.md): narrow outputs (N <= 256: the AttAdapter q|k|v projection and its
  // dgrad) and one-k-block products (K <= 64: the adapter output projections, pure epilogue work) run faster on the
  // single-CTA kernel, whose 128-row tiles put twice as many CTAs on the machine.
  if (mode == 0 && (p->n <= 256 || p->k <= GEMM_BK)) return 0;
  // ... and so does a short-K product with an MN-major B (the AttAdapter's dz = dqkv · W_qkv, K = 192): 8.6 us on the single-CTA
  // kernel against 10.7 us on 256 x 128 pair tiles (profiles/r1g_gemm_tile_sweep.md, row "att dz").
  if (mode == 0 && p->b_layout == JL_LAYOUT_MN && p->k <= 256) return 0;
  // ... and so do the large K-major products whose N is a multiple of 256 (q|k|v, attention / FFN output projections, their
  // dgrads): 128 x 256 single-CTA tiles need 1.28-3.8 waves of 148 CTAs where 256 x 192 pair tiles need 1.7-5.2 waves of 74
  // pairs, and measure 2-8 % faster (profiles/r1g_gemm_tile_sweep.md).  With the activation epilogues of the FFN input
  // projection and its dgrad (N = 3072) the single-CTA tile is equal or up to 6 % faster as well
  // (profiles/r1m_gemm_epilogues.md: 52.1 vs 55.4 us with the GELU + gelu' epilogue); the GLU convolutions stay on the pair kernel.
  if (mode == 0 && p->b_layout == JL_LAYOUT_K && (p->n % 256) == 0 && p->m >= 4096 && p->epilogue != JL_EPI_GLU) return 0;
  int per = 0;
  const bool have_ws = assume_ws || p->workspace != nullptr;
  if (have_ws && pick_split(p, pick_bn(p), &per) > 1) return 0;
  const bool bmn = p->b_layout == JL_LAYOUT_MN;
  int best = 0;
  double best_cost = -1.0;
  const int cands[3] = {256, 192, 128};
  for (int i = 0; i < 3; ++i) {
    const int bn = cands[i];
    if (bn == 192 && bmn) continue;
    if (bn > 128 && bn / 2 >= p->n) continue;
    const TailPlan t = plan_tail(p, bn, have_ws && (assume_ws || p->workspace_bytes >= static_cast<int64_t>(tail_ws_bytes(plan_tail(p, bn, true), bn))));
    const double cost = t.waves * (bn + 32);
    if (best_cost < 0 || cost < best_cost) { best_cost = cost; best = bn; }
  }
  // Long-K products whose N is a multiple of 256 (FFN output projection, dX of the q|k|v projection): the 256-wide tile moves
  // 15 % fewer operand bytes per FLOP through L2 and measures 1-6 % faster than the 192-wide one although both need two waves.
  if (mode == 0 && best == 192 && (p->n % 256) == 0 && p->k >= 2048) best = 256;
  return best;
}

// B multicast across two pairs (clusters of four CTAs): only when requested (mode 3) until validated on hardware.
static bool use_multicast(const jl_gemm_params* p, int bn) {
  (void)bn;
  return g_gemm_mode.load() == 3 && p->m >= 512;
}

}  // namespace jl

extern "C" {

int jl_gemm_bf16(const jl_gemm_params* p, void* stream) {
  int rc = jl::validate(p);
  if (rc != JL_OK) return rc;
  rc = jl::check_device();
  if (rc != JL_OK) return rc;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const int bn2 = jl::pick_bn_2cta(p);
  if (bn2 != 0 && jl::use_multicast(p, bn2)) {
    const bool bmn = p->b_layout == JL_LAYOUT_MN;
    if (bn2 == 256) return jl::dispatch_layout_2cta<256, 6, 2>(p, s);
    if (bn2 == 192 && !bmn) return jl::dispatch_layout_2cta<192, 6, 2>(p, s);
    if (bn2 == 128 && !bmn) return jl::dispatch_layout_2cta<128, 8, 2>(p, s);
  }
  switch (bn2) {
    case 256: return jl::dispatch_layout_2cta<256, 6, 1>(p, s);
    case 192: return jl::dispatch_layout_2cta<192, 6, 1>(p, s);
    case 128: return jl::dispatch_layout_2cta<128, 8, 1>(p, s);
    default: break;
  }
  switch (jl::pick_bn(p)) {
    case 256: return jl::dispatch_layout<256, 4>(p, s);
    case 128: return jl::dispatch_layout<128, 6>(p, s);
    case 64: return jl::dispatch_layout<64, 8>(p, s);
    default: return jl::dispatch_layout<32, 8>(p, s);
  }
}

void jl_debug_set_gemm_mode(int mode) { jl::g_gemm_mode.store(mode); }
void jl_debug_set_gemm_bn(int bn) { jl::g_gemm_bn.store(bn); }

int jl_gemm_workspace_bytes(const jl_gemm_params* p, size_t* out) {
  JL_REQUIRE(out != nullptr, JL_EINVAL, "gemm_workspace_bytes: null out");
  int rc = jl::validate(p);
  if (rc != JL_OK) return rc;
  *out = jl::splitk_ws_bytes(p);
  if (*out == 0) {
    const int bn2 = jl::pick_bn_2cta(p, true);
    if (bn2 != 0 && !jl::use_multicast(p, bn2)) *out = jl::tail_ws_bytes(jl::plan_tail(p, bn2, true), bn2);
  }
  return JL_OK;
}

int jl_gemm_workspace_zero_bytes(const jl_gemm_params* p, size_t* out) {
  JL_REQUIRE(out != nullptr, JL_EINVAL, "gemm_workspace_zero_bytes: null out");
  int rc = jl::validate(p);
  if (rc != JL_OK) return rc;
  *out = 0;
  if (jl::splitk_ws_bytes(p) != 0) return JL_OK;
  const int bn2 = jl::pick_bn_2cta(p, true);
  if (bn2 != 0 && !jl::use_multicast(p, bn2) && jl::plan_tail(p, bn2, true).split > 1) *out = jl::GEMM_TAIL_ZERO_BYTES;
  return JL_OK;
}

void jl_debug_set_gemm_tail(int mode) { jl::g_gemm_tail.store(mode & 7); }

int jl_debug_gemm_ref(const jl_gemm_params* p, void* stream) {
  int rc = jl::validate(p);
  if (rc != JL_OK) return rc;
  rc = jl::check_device();
  if (rc != JL_OK) return rc;
  jl::GemmRefOperands o;
  o.a = reinterpret_cast<const __nv_bfloat16*>(p->a); o.lda = p->lda; o.a_mn = p->a_layout == JL_LAYOUT_MN;
  o.b = reinterpret_cast<const __nv_bfloat16*>(p->b); o.ldb = p->ldb; o.b_mn = p->b_layout == JL_LAYOUT_MN;
  const jl::GemmDev g = jl::to_dev(p, 32);
  const int64_t total = static_cast<int64_t>(p->m) * ((p->n + 31) / 32);
  const int threads = 128;
  const int64_t blocks = (total + threads - 1) / threads;
  jl::launch(jl::gemm_ref_kernel, static_cast<unsigned>(blocks), threads, 0, reinterpret_cast<cudaStream_t>(stream), o, g);
  JL_CHECK_LAUNCH("gemm_ref");
  return JL_OK;
}

}  // extern "C"
