// LayerNorm over the last dim, bf16 in / bf16 out, fp32 statistics (eps 1e-5 on the path).
// Replaces nn.LayerNorm at SP/transformers/models/wav2vec2/modeling_wav2vec2.py:623,625,639,645 (pre-LN layer),
// :792 (final encoder norm) and :941 (adapter norm), forward and backward (dx; dgamma/dbeta only for the
// trainable adapter norms — the backbone is frozen).
//
// HBM-bound: one warp per row, the whole row lives in registers (16-byte loads, two-pass variance in registers),
// one read + one write of the activation.  Backward keeps per-lane dgamma/dbeta accumulators in registers across
// the rows a CTA owns, reduces them through shared memory and leaves one partial row per CTA; a second small
// kernel sums the partials in fixed order (deterministic, no atomics).
#include "common.cuh"
#include "ptx_sm100.cuh"

namespace jl {

constexpr int LN_WARPS = 8;
constexpr int LN_THREADS = LN_WARPS * 32;
// CTA sizes of the two streaming kernels that run 72 times per step (forward; backward of the frozen norms): small CTAs spread
// the 8000 rows evenly over the 148 SMs (250 CTAs of 32 rows leave 46 SMs with half the work of the others): 9.8 -> 9.3 us and 10.6 -> 10.0 us at 8000 x 768
// with 2 warps per CTA (profiles/README.md).
#ifndef JL_LN_FWD_WARPS
#define JL_LN_FWD_WARPS 2
#endif
#ifndef JL_LN_FWD_HOIST
#define JL_LN_FWD_HOIST 0
#endif
#ifndef JL_LN_BWD_WARPS
#define JL_LN_BWD_WARPS 2
#endif

template <int NCH>
__device__ __forceinline__ void ln_load_row(const __nv_bfloat16* row, int nchunks, int lane, float (&x)[NCH][8]) {
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
    const int ch = c * 32 + lane;
    if (ch < nchunks) {
      const uint4 v = __ldg(reinterpret_cast<const uint4*>(row) + ch);
      const float2 f0 = unpack_bf16x2(v.x), f1 = unpack_bf16x2(v.y), f2 = unpack_bf16x2(v.z), f3 = unpack_bf16x2(v.w);
      x[c][0] = f0.x; x[c][1] = f0.y; x[c][2] = f1.x; x[c][3] = f1.y;
      x[c][4] = f2.x; x[c][5] = f2.y; x[c][6] = f3.x; x[c][7] = f3.y;
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) x[c][j] = 0.0f;
    }
  }
}

__device__ __forceinline__ void ln_load8_f32(const float* p, float (&o)[8]) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(p));
  const float4 b = __ldg(reinterpret_cast<const float4*>(p) + 1);
  o[0] = a.x; o[1] = a.y; o[2] = a.z; o[3] = a.w; o[4] = b.x; o[5] = b.y; o[6] = b.z; o[7] = b.w;
}

__device__ __forceinline__ void ln_store8_bf16(__nv_bfloat16* p, const float (&v)[8]) {
  uint4 o;
  o.x = pack_bf16x2(v[0], v[1]); o.y = pack_bf16x2(v[2], v[3]); o.z = pack_bf16x2(v[4], v[5]); o.w = pack_bf16x2(v[6], v[7]);
  *reinterpret_cast<uint4*>(p) = o;
}

__device__ __forceinline__ void ln_unpack8(const uint4& v, float (&o)[8]) {
  const float2 f0 = unpack_bf16x2(v.x), f1 = unpack_bf16x2(v.y), f2 = unpack_bf16x2(v.z), f3 = unpack_bf16x2(v.w);
  o[0] = f0.x; o[1] = f0.y; o[2] = f1.x; o[3] = f1.y; o[4] = f2.x; o[5] = f2.y; o[6] = f3.x; o[7] = f3.y;
}

// A warp owns LN_ROWS consecutive rows and requests all of them before it reduces the first one, so 12 (d = 768) 16-byte
// loads per lane are in flight instead of 3 — the kernel is a pure stream and lives on memory-level parallelism.
constexpr int LN_ROWS = 4;
template <int NCH>
__global__ void __launch_bounds__(LN_THREADS) layernorm_fwd_kernel(const jl_layernorm_fwd_params p) {
  jl::pdl_prologue();
  const int lane = threadIdx.x & 31;
  const int row0 = (blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * LN_ROWS;
  if (row0 >= p.rows) return;
  const int nchunks = p.d >> 3;
  const float inv_d = 1.0f / static_cast<float>(p.d);
  uint4 raw[LN_ROWS][NCH];
#pragma unroll
  for (int r = 0; r < LN_ROWS; ++r) {
    const __nv_bfloat16* xr = reinterpret_cast<const __nv_bfloat16*>(p.x) + static_cast<int64_t>(min(row0 + r, p.rows - 1)) * p.ldx;
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      const int ch = c * 32 + lane;
      raw[r][c] = (ch < nchunks) ? __ldg(reinterpret_cast<const uint4*>(xr) + ch) : make_uint4(0u, 0u, 0u, 0u);
    }
  }
#if JL_LN_FWD_HOIST
  // γ, β of this lane's chunks once per warp (4 rows), not once per row: a quarter of the L1 traffic of the stream
  float gq[NCH][8], bq[NCH][8];
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
    const int ch = min(c * 32 + lane, nchunks - 1);
    ln_load8_f32(p.gamma + ch * 8, gq[c]);
    ln_load8_f32(p.beta + ch * 8, bq[c]);
  }
#endif
#pragma unroll
  for (int r = 0; r < LN_ROWS; ++r) {
    const int row = row0 + r;
    if (row >= p.rows) break;
    float x[NCH][8];
    float sum = 0.0f;
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      ln_unpack8(raw[r][c], x[c]);
#pragma unroll
      for (int j = 0; j < 8; ++j) sum += x[c][j];
    }
    const float mean = warp_sum(sum) * inv_d;
    float sq = 0.0f;
#pragma unroll
    for (int c = 0; c < NCH; ++c)
      if (c * 32 + lane < nchunks) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float d = x[c][j] - mean;
          sq = fmaf(d, d, sq);
        }
      }
    const float var = warp_sum(sq) * inv_d;
    const float rstd = 1.0f / sqrtf(var + p.eps);
    if (lane == 0) {
      if (p.mean != nullptr) p.mean[row] = mean;
      if (p.rstd != nullptr) p.rstd[row] = rstd;
    }
    __nv_bfloat16* yr = reinterpret_cast<__nv_bfloat16*>(p.y) + static_cast<int64_t>(row) * p.ldy;
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      const int ch = c * 32 + lane;
      if (ch < nchunks) {
        float o[8];
#if JL_LN_FWD_HOIST
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = fmaf((x[c][j] - mean) * rstd, gq[c][j], bq[c][j]);
#else
        float g[8], b[8];
        ln_load8_f32(p.gamma + ch * 8, g);
        ln_load8_f32(p.beta + ch * 8, b);
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = fmaf((x[c][j] - mean) * rstd, g[j], b[j]);
#endif
        if (p.act == JL_EPI_GELU) {
#pragma unroll
          for (int j = 0; j < 8; ++j) o[j] = gelu_erf(o[j]);
        }
        ln_store8_bf16(yr + ch * 8, o);
      }
    }
  }
}

// WGRAD = false (frozen norms of the backbone: dx only) keeps the dγ/dβ accumulators out of the register file, which
// doubles the resident warps of this HBM-bound kernel.
template <int NCH, bool WGRAD>
__global__ void __launch_bounds__(LN_THREADS) layernorm_bwd_kernel(const jl_layernorm_bwd_params p) {
  jl::pdl_prologue();
  __shared__ float s_red[WGRAD ? NCH * 256 : 1];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nchunks = p.d >> 3;
  const float inv_d = 1.0f / static_cast<float>(p.d);
  constexpr int NW = WGRAD ? NCH : 1;
  float dg[NW][8], db[NW][8];
#pragma unroll
  for (int c = 0; c < NW; ++c)
#pragma unroll
    for (int j = 0; j < 8; ++j) { dg[c][j] = 0.0f; db[c][j] = 0.0f; }

  constexpr int RB = WGRAD ? 1 : 2;       // rows per warp-iteration (the lean dx-only variant keeps two rows in flight)
  const int nwarps = blockDim.x >> 5;
  for (int row0 = (blockIdx.x * nwarps + warp) * RB; row0 < p.rows; row0 += gridDim.x * nwarps * RB) {
    uint4 rx[RB][NCH], rdy[RB][NCH], rdr[RB][NCH];
#pragma unroll
    for (int r = 0; r < RB; ++r) {
      const int64_t rr = min(row0 + r, p.rows - 1);
#pragma unroll
      for (int c = 0; c < NCH; ++c) {
        const int ch = c * 32 + lane;
        const bool ok = ch < nchunks;
        rx[r][c] = ok ? __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(p.x) + rr * p.ldx) + ch) : make_uint4(0u, 0u, 0u, 0u);
        rdy[r][c] = ok ? __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(p.dy) + rr * p.lddy) + ch) : make_uint4(0u, 0u, 0u, 0u);
        rdr[r][c] = (ok && p.dres != nullptr) ? __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(p.dres) + rr * p.lddres) + ch)
                                             : make_uint4(0u, 0u, 0u, 0u);
      }
    }
#pragma unroll
    for (int r = 0; r < RB; ++r) {
      const int row = row0 + r;
      if (row >= p.rows) break;
      float x[NCH][8], dy[NCH][8];
#pragma unroll
      for (int c = 0; c < NCH; ++c) {
        ln_unpack8(rx[r][c], x[c]);
        ln_unpack8(rdy[r][c], dy[c]);
      }
      const float mean = __ldg(p.mean + row), rstd = __ldg(p.rstd + row);
      float s1 = 0.0f, s2 = 0.0f;
#pragma unroll
      for (int c = 0; c < NCH; ++c) {
        const int ch = c * 32 + lane;
        if (ch < nchunks) {
          float g[8];
          ln_load8_f32(p.gamma + ch * 8, g);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float xh = (x[c][j] - mean) * rstd;
            if constexpr (WGRAD) { dg[c][j] = fmaf(dy[c][j], xh, dg[c][j]); db[c][j] += dy[c][j]; }
            const float gy = dy[c][j] * g[j];
            x[c][j] = xh;
            dy[c][j] = gy;
            s1 += gy;
            s2 = fmaf(gy, xh, s2);
          }
        }
      }
      const float c1 = warp_sum(s1) * inv_d, c2 = warp_sum(s2) * inv_d;
      __nv_bfloat16* dxr = reinterpret_cast<__nv_bfloat16*>(p.dx) + static_cast<int64_t>(row) * p.lddx;
#pragma unroll
      for (int c = 0; c < NCH; ++c) {
        const int ch = c * 32 + lane;
        if (ch < nchunks) {
          float o[8], dr[8];
          ln_unpack8(rdr[r][c], dr);
#pragma unroll
          for (int j = 0; j < 8; ++j) o[j] = rstd * (dy[c][j] - c1 - x[c][j] * c2) + dr[j];
          ln_store8_bf16(dxr + ch * 8, o);
        }
      }
    }
  }

  if constexpr (WGRAD) {
    // CTA reduction of the per-warp accumulators (warps add in fixed order) → one partial row per CTA
    for (int pass = 0; pass < 2; ++pass) {
      for (int w = 0; w < LN_WARPS; ++w) {
        __syncthreads();
        if (warp == w) {
#pragma unroll
          for (int c = 0; c < NCH; ++c)
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float v = pass == 0 ? dg[c][j] : db[c][j];
              float* dst = &s_red[(c * 32 + lane) * 8 + j];
              *dst = (w == 0) ? v : (*dst + v);
            }
        }
      }
      __syncthreads();
      float* dst = p.partial + (static_cast<int64_t>(pass) * gridDim.x + blockIdx.x) * p.d;
      for (int i = threadIdx.x; i < p.d; i += LN_THREADS) dst[i] = s_red[i];
    }
  }
}

// partial = [2][nblk][d] (dgamma partials, then dbeta partials).  CTA = 32 columns × 8 row groups; each group sums a
// strided subset of the partial rows (coalesced 128 B reads), then the 8 groups are combined in fixed order.
__global__ void __launch_bounds__(256) layernorm_bwd_reduce_kernel(const float* __restrict__ partial, int nblk, int d,
                                                                   float* __restrict__ dgamma, float* __restrict__ dbeta) {
  jl::pdl_prologue();
  __shared__ float s_g[8][33], s_b[8][33];
  const int lane = threadIdx.x & 31, grp = threadIdx.x >> 5;
  const int i = blockIdx.x * 32 + lane;
  float g = 0.0f, b = 0.0f;
  if (i < d) {
    for (int k = grp; k < nblk; k += 8) {
      g += partial[static_cast<int64_t>(k) * d + i];
      b += partial[(static_cast<int64_t>(nblk) + k) * d + i];
    }
  }
  s_g[grp][lane] = g;
  s_b[grp][lane] = b;
  __syncthreads();
  if (grp == 0 && i < d) {
#pragma unroll
    for (int w = 1; w < 8; ++w) { g += s_g[w][lane]; b += s_b[w][lane]; }
    dgamma[i] = g;
    if (dbeta != nullptr) dbeta[i] = b;
  }
}

// dγ = Σ_rows dy ∘ x̂, dβ = Σ_rows dy in ONE launch, laid out like colsum_kernel (elementwise.cu): a cluster of 8 CTAs owns 32
// adjacent columns, four threads cover a row's 64 contiguous bytes, the row groups of 64 are dealt round-robin to the CTAs,
// fixed-order shared-memory tree per CTA, CTA 0 adds the eight partials in rank order through distributed shared memory.
// Lets the dX part of an adapter norm's backward use the lean kernel on the critical path while this runs on the
// weight-gradient branch.
constexpr int LNW_CLUSTER = 8;
constexpr int LNW_COLS = 32;
__global__ void __cluster_dims__(LNW_CLUSTER, 1, 1) __launch_bounds__(256)
layernorm_wgrad_kernel(const __nv_bfloat16* __restrict__ dy, int64_t lddy, const __nv_bfloat16* __restrict__ x, int64_t ldx,
                       const float* __restrict__ mean, const float* __restrict__ rstd, int rows, int d, float* __restrict__ dgamma,
                       float* __restrict__ dbeta) {
  jl::pdl_prologue();
  __shared__ float s_g[4][64][9], s_b[4][64][9];
  __shared__ float s_part[2][LNW_COLS];
  const int tid = threadIdx.x;
  const int rank = static_cast<int>(ptx::cluster_ctarank());
  const int grp = tid & 3, rix = tid >> 2;
  const int col = (blockIdx.x / LNW_CLUSTER) * LNW_COLS + grp * 8;
  float g[8], b[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { g[j] = 0.0f; b[j] = 0.0f; }
  constexpr int STEP = LNW_CLUSTER * 64;
  if (col < d) {
    for (int r0 = rank * 64 + rix; r0 < rows; r0 += 4 * STEP) {
      uint4 vy[4], vx[4];
      float mu[4], rs[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {                 // four independent rows in flight per thread
        const int r = min(r0 + STEP * u, rows - 1);
        vy[u] = __ldg(reinterpret_cast<const uint4*>(dy + static_cast<int64_t>(r) * lddy + col));
        vx[u] = __ldg(reinterpret_cast<const uint4*>(x + static_cast<int64_t>(r) * ldx + col));
        mu[u] = __ldg(mean + r);
        rs[u] = __ldg(rstd + r);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (r0 + STEP * u >= rows) break;
        const uint32_t wy[4] = {vy[u].x, vy[u].y, vy[u].z, vy[u].w}, wx[4] = {vx[u].x, vx[u].y, vx[u].z, vx[u].w};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float2 fy = unpack_bf16x2(wy[q]), fx = unpack_bf16x2(wx[q]);
          g[2 * q] = fmaf(fy.x, (fx.x - mu[u]) * rs[u], g[2 * q]);
          g[2 * q + 1] = fmaf(fy.y, (fx.y - mu[u]) * rs[u], g[2 * q + 1]);
          b[2 * q] += fy.x;
          b[2 * q + 1] += fy.y;
        }
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) { s_g[grp][rix][j] = g[j]; s_b[grp][rix][j] = b[j]; }
  __syncthreads();
  for (int stride = 32; stride >= 1; stride >>= 1) {
    if (rix < stride) {
#pragma unroll
      for (int j = 0; j < 8; ++j) { s_g[grp][rix][j] += s_g[grp][rix + stride][j]; s_b[grp][rix][j] += s_b[grp][rix + stride][j]; }
    }
    __syncthreads();
  }
  if (tid < LNW_COLS) {
    s_part[0][tid] = s_g[tid >> 3][0][tid & 7];
    s_part[1][tid] = s_b[tid >> 3][0][tid & 7];
  }
  ptx::cluster_sync_all();
  if (rank == 0 && tid < LNW_COLS) {
    const int c = (blockIdx.x / LNW_CLUSTER) * LNW_COLS + tid;
    if (c < d) {
      const uint32_t lg = ptx::smem_u32(&s_part[0][tid]), lb = ptx::smem_u32(&s_part[1][tid]);
      float tg = 0.0f, tb = 0.0f;
#pragma unroll
      for (int q = 0; q < LNW_CLUSTER; ++q) {
        tg += ptx::ld_shared_cluster_f32(ptx::mapa_shared(lg, static_cast<uint32_t>(q)));
        tb += ptx::ld_shared_cluster_f32(ptx::mapa_shared(lb, static_cast<uint32_t>(q)));
      }
      dgamma[c] = tg;
      if (dbeta != nullptr) dbeta[c] = tb;
    }
  }
  ptx::cluster_sync_all();      // the partials of every CTA stay mapped until CTA 0 has read them
}

// Several column reductions in ONE launch (the weight-gradient branch of an adapter's backward pass needs three: the bias gradients
// Σ_r dy[r, :] of its two projections and the dγ / dβ of its LayerNorm).  A job is Σ_r dy[r, c] (out_sum) and, when x is given,
// Σ_r dy[r, c] · (x[r, c] − mean[r]) · rstd[r] (out_dot).  Same scheme as layernorm_wgrad_kernel — a cluster of 8 CTAs per 32
// columns, fixed-order trees, the 8 partials added in rank order through distributed shared memory — but the clusters of all jobs
// share one grid: 54 clusters instead of three launches of 6-24, and one launch latency instead of three.
struct ColReduceDev {
  jl_colreduce_job job[JL_COLREDUCE_MAX_JOBS];
  int32_t first_cluster[JL_COLREDUCE_MAX_JOBS + 1];     // prefix sums of the jobs' cluster counts
  int32_t num_jobs;
};
__global__ void __cluster_dims__(LNW_CLUSTER, 1, 1) __launch_bounds__(256) colreduce_multi_kernel(const ColReduceDev p) {
  jl::pdl_prologue();
  __shared__ float s_g[4][64][9], s_b[4][64][9];
  __shared__ float s_part[2][LNW_COLS];
  const int tid = threadIdx.x;
  const int rank = static_cast<int>(ptx::cluster_ctarank());
  const int cluster = blockIdx.x / LNW_CLUSTER;
  int ji = 0;
  while (ji + 1 < p.num_jobs && cluster >= p.first_cluster[ji + 1]) ++ji;
  const jl_colreduce_job& jb = p.job[ji];
  const int grp = tid & 3, rix = tid >> 2;
  const int col0 = (cluster - p.first_cluster[ji]) * LNW_COLS;
  const int col = col0 + grp * 8;
  const __nv_bfloat16* dy = reinterpret_cast<const __nv_bfloat16*>(jb.dy);
  const __nv_bfloat16* x = reinterpret_cast<const __nv_bfloat16*>(jb.x);
  const bool dot = x != nullptr;
  const int rows = jb.rows;
  float g[8], b[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { g[j] = 0.0f; b[j] = 0.0f; }
  constexpr int STEP = LNW_CLUSTER * 64;
  if (col < jb.cols) {
    for (int r0 = rank * 64 + rix; r0 < rows; r0 += 4 * STEP) {
      uint4 vy[4], vx[4];
      float mu[4], rs[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {                 // four independent rows in flight per thread
        const int r = min(r0 + STEP * u, rows - 1);
        vy[u] = __ldg(reinterpret_cast<const uint4*>(dy + static_cast<int64_t>(r) * jb.lddy + col));
        if (dot) {
          vx[u] = __ldg(reinterpret_cast<const uint4*>(x + static_cast<int64_t>(r) * jb.ldx + col));
          mu[u] = __ldg(jb.mean + r);
          rs[u] = __ldg(jb.rstd + r);
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (r0 + STEP * u >= rows) break;
        const uint32_t wy[4] = {vy[u].x, vy[u].y, vy[u].z, vy[u].w};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float2 fy = unpack_bf16x2(wy[q]);
          b[2 * q] += fy.x;
          b[2 * q + 1] += fy.y;
        }
        if (dot) {
          const uint32_t wx[4] = {vx[u].x, vx[u].y, vx[u].z, vx[u].w};
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float2 fy = unpack_bf16x2(wy[q]), fx = unpack_bf16x2(wx[q]);
            g[2 * q] = fmaf(fy.x, (fx.x - mu[u]) * rs[u], g[2 * q]);
            g[2 * q + 1] = fmaf(fy.y, (fx.y - mu[u]) * rs[u], g[2 * q + 1]);
          }
        }
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) { s_g[grp][rix][j] = g[j]; s_b[grp][rix][j] = b[j]; }
  __syncthreads();
  for (int stride = 32; stride >= 1; stride >>= 1) {
    if (rix < stride) {
#pragma unroll
      for (int j = 0; j < 8; ++j) { s_g[grp][rix][j] += s_g[grp][rix + stride][j]; s_b[grp][rix][j] += s_b[grp][rix + stride][j]; }
    }
    __syncthreads();
  }
  if (tid < LNW_COLS) {
    s_part[0][tid] = s_g[tid >> 3][0][tid & 7];
    s_part[1][tid] = s_b[tid >> 3][0][tid & 7];
  }
  ptx::cluster_sync_all();
  if (rank == 0 && tid < LNW_COLS) {
    const int c = col0 + tid;
    if (c < jb.cols) {
      const uint32_t lg = ptx::smem_u32(&s_part[0][tid]), lb = ptx::smem_u32(&s_part[1][tid]);
      float tg = 0.0f, tb = 0.0f;
#pragma unroll
      for (int q = 0; q < LNW_CLUSTER; ++q) {
        tg += ptx::ld_shared_cluster_f32(ptx::mapa_shared(lg, static_cast<uint32_t>(q)));
        tb += ptx::ld_shared_cluster_f32(ptx::mapa_shared(lb, static_cast<uint32_t>(q)));
      }
      if (jb.out_sum != nullptr) jb.out_sum[c] = tb;
      if (dot && jb.out_dot != nullptr) jb.out_dot[c] = tg;
    }
  }
  ptx::cluster_sync_all();      // the partials of every CTA stay mapped until CTA 0 has read them
}

static int ln_bwd_blocks(int rows) {
  int blocks = ceil_div(rows, LN_WARPS);
  return blocks < 296 ? blocks : 296;   // 2 CTAs per SM on 148 SMs
}

static int ln_pick(int d) { return d <= 768 ? 3 : (d <= 1024 ? 4 : 8); }

}  // namespace jl

extern "C" {

int jl_layernorm_fwd(const jl_layernorm_fwd_params* p, void* stream) {
  JL_REQUIRE(p && p->x && p->y && p->gamma && p->beta, JL_EINVAL, "layernorm_fwd: null pointer");
  JL_REQUIRE(p->rows > 0 && p->d > 0, JL_EINVAL, "layernorm_fwd: rows and d must be positive");
  JL_REQUIRE((p->d & 7) == 0 && p->d <= 2048, JL_EUNSUPPORTED_SHAPE, "layernorm_fwd: d must be a multiple of 8 and <= 2048 (got %d)", p->d);
  JL_REQUIRE((p->ldx & 7) == 0 && (p->ldy & 7) == 0, JL_EINVAL, "layernorm_fwd: row strides must be multiples of 8");
  JL_REQUIRE(p->act == JL_EPI_NONE || p->act == JL_EPI_GELU, JL_EINVAL, "layernorm_fwd: act must be JL_EPI_NONE or JL_EPI_GELU (got %d)", p->act);
  JL_REQUIRE(((reinterpret_cast<uintptr_t>(p->x) | reinterpret_cast<uintptr_t>(p->y) | reinterpret_cast<uintptr_t>(p->gamma) |
               reinterpret_cast<uintptr_t>(p->beta)) & 15) == 0, JL_EINVAL, "layernorm_fwd: pointers must be 16-byte aligned");
  int rc = jl::check_device();
  if (rc != JL_OK) return rc;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const int blocks = jl::ceil_div(p->rows, JL_LN_FWD_WARPS * jl::LN_ROWS);
  switch (jl::ln_pick(p->d)) {
    case 3: jl::launch(jl::layernorm_fwd_kernel<3>, blocks, JL_LN_FWD_WARPS * 32, 0, s, *p); break;
    case 4: jl::launch(jl::layernorm_fwd_kernel<4>, blocks, JL_LN_FWD_WARPS * 32, 0, s, *p); break;
    default: jl::launch(jl::layernorm_fwd_kernel<8>, blocks, JL_LN_FWD_WARPS * 32, 0, s, *p); break;
  }
  JL_CHECK_LAUNCH("layernorm_fwd");
  return JL_OK;
}

int jl_layernorm_wgrad(const jl_layernorm_bwd_params* p, void* stream) {
  JL_REQUIRE(p && p->dy && p->x && p->mean && p->rstd && p->dgamma, JL_EINVAL, "layernorm_wgrad: null pointer");
  JL_REQUIRE(p->rows > 0 && p->d > 0 && (p->d & 7) == 0, JL_EINVAL, "layernorm_wgrad: d must be a positive multiple of 8");
  JL_REQUIRE((p->ldx & 7) == 0 && (p->lddy & 7) == 0, JL_EINVAL, "layernorm_wgrad: row strides must be multiples of 8");
  JL_REQUIRE(((reinterpret_cast<uintptr_t>(p->x) | reinterpret_cast<uintptr_t>(p->dy)) & 15) == 0, JL_EINVAL, "layernorm_wgrad: pointers must be 16-byte aligned");
  int rc = jl::check_device();
  if (rc != JL_OK) return rc;
  jl::launch(jl::layernorm_wgrad_kernel, jl::ceil_div(p->d, jl::LNW_COLS) * jl::LNW_CLUSTER, 256, 0, reinterpret_cast<cudaStream_t>(stream), reinterpret_cast<const __nv_bfloat16*>(p->dy),
             p->lddy, reinterpret_cast<const __nv_bfloat16*>(p->x), p->ldx, p->mean, p->rstd, p->rows, p->d, p->dgamma, p->dbeta);
  JL_CHECK_LAUNCH("layernorm_wgrad");
  return JL_OK;
}

int jl_colreduce_multi(const jl_colreduce_job* jobs, int32_t num_jobs, void* stream) {
  JL_REQUIRE(jobs != nullptr && num_jobs >= 1 && num_jobs <= JL_COLREDUCE_MAX_JOBS, JL_EINVAL, "colreduce_multi: 1..%d jobs", JL_COLREDUCE_MAX_JOBS);
  jl::ColReduceDev d;
  d.num_jobs = num_jobs;
  int clusters = 0;
  for (int i = 0; i < num_jobs; ++i) {
    const jl_colreduce_job& j = jobs[i];
    JL_REQUIRE(j.dy != nullptr && j.rows > 0 && j.cols > 0, JL_EINVAL, "colreduce_multi: job %d: null dy or empty shape", i);
    JL_REQUIRE((j.cols & 7) == 0 && (j.lddy & 7) == 0 && (reinterpret_cast<uintptr_t>(j.dy) & 15) == 0, JL_EINVAL,
               "colreduce_multi: job %d: cols / lddy must be multiples of 8 and dy 16-byte aligned", i);
    if (j.x != nullptr)
      JL_REQUIRE(j.mean && j.rstd && (j.ldx & 7) == 0 && (reinterpret_cast<uintptr_t>(j.x) & 15) == 0, JL_EINVAL,
                 "colreduce_multi: job %d: x needs mean, rstd, ldx %% 8 == 0 and 16-byte alignment", i);
    JL_REQUIRE(j.out_sum != nullptr || (j.x != nullptr && j.out_dot != nullptr), JL_EINVAL, "colreduce_multi: job %d has no output", i);
    d.job[i] = j;
    d.first_cluster[i] = clusters;
    clusters += jl::ceil_div(j.cols, jl::LNW_COLS);
  }
  for (int i = num_jobs; i <= JL_COLREDUCE_MAX_JOBS; ++i) d.first_cluster[i] = clusters;
  int rc = jl::check_device();
  if (rc != JL_OK) return rc;
  jl::launch(jl::colreduce_multi_kernel, clusters * jl::LNW_CLUSTER, 256, 0, reinterpret_cast<cudaStream_t>(stream), d);
  JL_CHECK_LAUNCH("colreduce_multi");
  return JL_OK;
}

int jl_layernorm_bwd_workspace_bytes(const jl_layernorm_bwd_params* p, size_t* out) {
  JL_REQUIRE(p && out, JL_EINVAL, "layernorm_bwd_workspace_bytes: null argument");
  *out = (p->dgamma != nullptr) ? static_cast<size_t>(2) * jl::ln_bwd_blocks(p->rows) * p->d * sizeof(float) : 0;
  return JL_OK;
}

int jl_layernorm_bwd(const jl_layernorm_bwd_params* p, void* stream) {
  JL_REQUIRE(p && p->dy && p->x && p->gamma && p->mean && p->rstd && p->dx, JL_EINVAL, "layernorm_bwd: null pointer");
  JL_REQUIRE(p->rows > 0 && p->d > 0, JL_EINVAL, "layernorm_bwd: rows and d must be positive");
  JL_REQUIRE((p->d & 7) == 0 && p->d <= 2048, JL_EUNSUPPORTED_SHAPE, "layernorm_bwd: d must be a multiple of 8 and <= 2048 (got %d)", p->d);
  JL_REQUIRE((p->ldx & 7) == 0 && (p->lddy & 7) == 0 && (p->lddx & 7) == 0 && (p->dres == nullptr || (p->lddres & 7) == 0), JL_EINVAL,
             "layernorm_bwd: row strides must be multiples of 8");
  JL_REQUIRE(p->dgamma == nullptr || p->partial != nullptr, JL_EINVAL, "layernorm_bwd: dgamma needs the partial workspace");
  int rc = jl::check_device();
  if (rc != JL_OK) return rc;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const int blocks = (p->dgamma != nullptr) ? jl::ln_bwd_blocks(p->rows) : jl::ceil_div(p->rows, JL_LN_BWD_WARPS * 2);
  const bool wg = p->dgamma != nullptr;
  switch (jl::ln_pick(p->d)) {
    case 3:
      if (wg) jl::launch(jl::layernorm_bwd_kernel<3, true>, blocks, jl::LN_THREADS, 0, s, *p);
      else jl::launch(jl::layernorm_bwd_kernel<3, false>, blocks, JL_LN_BWD_WARPS * 32, 0, s, *p);
      break;
    case 4:
      if (wg) jl::launch(jl::layernorm_bwd_kernel<4, true>, blocks, jl::LN_THREADS, 0, s, *p);
      else jl::launch(jl::layernorm_bwd_kernel<4, false>, blocks, JL_LN_BWD_WARPS * 32, 0, s, *p);
      break;
    default:
      if (wg) jl::launch(jl::layernorm_bwd_kernel<8, true>, blocks, jl::LN_THREADS, 0, s, *p);
      else jl::launch(jl::layernorm_bwd_kernel<8, false>, blocks, JL_LN_BWD_WARPS * 32, 0, s, *p);
      break;
  }
  JL_CHECK_LAUNCH("layernorm_bwd");
  if (p->dgamma != nullptr) {
    jl::launch(jl::layernorm_bwd_reduce_kernel, jl::ceil_div(p->d, 32), 256, 0, s, p->partial, blocks, p->d, p->dgamma, p->dbeta);
    JL_CHECK_LAUNCH("layernorm_bwd_reduce");
  }
  return JL_OK;
}

}  // extern "C"
