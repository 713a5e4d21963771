// a9 + a10: fused log-softmax + CTC loss/gradient, and the greedy collapse decoder.
//
// Replaces SP/transformers/models/wav2vec2/modeling_wav2vec2.py:1711-1736 (labels >= 0 selection, fp32 log_softmax,
// ctc_loss with blank = pad_token_id, reduction, zero_infinity) → SP/torch/nn/functional.py:3042-3115, and the
// argmax + groupby collapse + pad strip of SP/transformers/models/wav2vec2/tokenization_wav2vec2.py:310-317.
//
// HBM-bound design: the [B·T, V] logits are streamed once per pass with 16-byte loads, one warp per frame:
//   ctc_prep        compact the labels (>= 0) per utterance, target lengths
//   ctc_row_stats   online max / log-sum-exp over V by warp shuffle (+ first-max argmax for greedy), then gathers the
//                   2S+1 extended-label log-probs of the frame while the row is still hot in L1/L2
//   ctc_lattice     one CTA per (utterance, direction): the alpha and the beta recursion of an utterance run
//                   on different SMs over the gathered log-probs, one barrier per frame, rows double-buffered in smem
//   ctc_grad        one CTA per frame: grad = softmax − Σ_{s: l'_s = v} exp(α+β−lp+nll); duplicate labels are combined in
//                   fixed order in shared memory (deterministic — no float atomics), the row is written once
//   ctc_reduce      reduction "sum" | "mean" and zero_infinity
#include <math_constants.h>

#include <algorithm>

#include "common.cuh"

namespace jl {

constexpr int CTC_LATTICE_HALF = 512;
constexpr float TC_LOG2E_F = 1.4426950408889634f;
constexpr float TC_LN2_F = 0.6931471805599453f;
constexpr int CTC_LAT_AHEAD = 8;          // frames of log-probs kept in flight by the lattice recursion

__device__ __forceinline__ void cp_async_f32(float* smem_dst, const float* gmem_src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(static_cast<uint32_t>(__cvta_generic_to_shared(smem_dst))), "l"(gmem_src)
               : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
constexpr int CTC_GRAD_THREADS = 256;

struct CtcWs {
  int32_t* labels;   // [B, Smax] compacted
  int32_t* nxt;      // [B, Smax] next position holding the same label, -1 at the end of the chain
  int32_t* fst;      // [B, Smax] 1 where the position is the first occurrence of its label
  int32_t* nblank;   // [B] number of labels equal to the blank index (degenerate targets)
  int32_t* tlen;     // [B]
  float* lse;        // [B*T]
  float* lpx;        // [B, T, L]
  float* alpha;      // [B, T, L]
  float* beta;       // [B, T, L]
};

__host__ __device__ inline size_t align256(size_t x) { return (x + 255) & ~static_cast<size_t>(255); }

static CtcWs carve_ws(const jl_ctc_params* p, void* ws, size_t* total) {
  const size_t B = p->batch, T = p->seq, S = p->max_label_len > 0 ? p->max_label_len : 1, L = 2 * S + 1;
  size_t off = 0;
  uint8_t* base = reinterpret_cast<uint8_t*>(ws);
  CtcWs w;
  w.labels = reinterpret_cast<int32_t*>(base + off); off += align256(B * S * 4);
  w.nxt = reinterpret_cast<int32_t*>(base + off); off += align256(B * S * 4);
  w.fst = reinterpret_cast<int32_t*>(base + off); off += align256(B * S * 4);
  w.nblank = reinterpret_cast<int32_t*>(base + off); off += align256(B * 4);
  w.tlen = reinterpret_cast<int32_t*>(base + off); off += align256(B * 4);
  w.lse = reinterpret_cast<float*>(base + off); off += align256(B * T * 4);
  w.lpx = reinterpret_cast<float*>(base + off); off += align256(B * T * L * 4);
  w.alpha = reinterpret_cast<float*>(base + off); off += align256(B * T * L * 4);
  w.beta = reinterpret_cast<float*>(base + off); off += align256(B * T * L * 4);
  if (total) *total = off;
  return w;
}

// One CTA per utterance: compact the labels, then link repeated labels (first-occurrence flag + next-occurrence index) so
// that the gradient kernel combines duplicates in position order with O(S) work per frame instead of O(S²).
constexpr int CTC_PREP_THREADS = 128;
__global__ void __launch_bounds__(CTC_PREP_THREADS) ctc_prep_kernel(const int32_t* __restrict__ labels, int smax, int vocab, int blank,
                                                                    int32_t* __restrict__ out_labels, int32_t* __restrict__ nxt,
                                                                    int32_t* __restrict__ fst, int32_t* __restrict__ nblank,
                                                                    int32_t* __restrict__ tlen) {
  jl::pdl_prologue();
  __shared__ int s_n, s_nb;
  __shared__ int s_warp_cnt[CTC_PREP_THREADS / 32];
  const int b = blockIdx.x;
  int32_t* lab = out_labels + static_cast<int64_t>(b) * smax;
  // order-preserving compaction of the non-negative labels: coalesced loads, ballot prefix inside a warp, running offset
  // across the rounds of 128 labels (one dependent global load per label made this kernel a 14 us latency chain)
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  int base = 0, nb_local = 0;
  for (int j0 = 0; j0 < smax; j0 += CTC_PREP_THREADS) {
    const int j = j0 + threadIdx.x;
    const int v = (j < smax) ? labels[static_cast<int64_t>(b) * smax + j] : -1;
    const bool keep = v >= 0;
    const unsigned m = __ballot_sync(0xffffffffu, keep);
    if (lane == 0) s_warp_cnt[wid] = __popc(m);
    __syncthreads();
    int off = base;
    for (int w = 0; w < wid; ++w) off += s_warp_cnt[w];
    int total = 0;
    for (int w = 0; w < CTC_PREP_THREADS / 32; ++w) total += s_warp_cnt[w];
    if (keep) {
      const int c = min(v, vocab - 1);     // range is validated on the host
      lab[off + __popc(m & ((1u << lane) - 1u))] = c;
      nb_local += (c == blank);
    }
    base += total;
    __syncthreads();
  }
  if (threadIdx.x == 0) { s_n = base; s_nb = 0; }
  __syncthreads();
  if (nb_local) atomicAdd(&s_nb, nb_local);      // integer count: order-independent
  __syncthreads();
  if (threadIdx.x == 0) {
    tlen[b] = s_n;
    nblank[b] = s_nb;
  }
  const int n = s_n;
  for (int j = threadIdx.x; j < n; j += CTC_PREP_THREADS) {
    const int c = lab[j];
    int first = 1, next = -1;
    for (int i = 0; i < j; ++i)
      if (lab[i] == c) { first = 0; break; }
    for (int i = j + 1; i < n; ++i)
      if (lab[i] == c) { next = i; break; }
    fst[static_cast<int64_t>(b) * smax + j] = first;
    nxt[static_cast<int64_t>(b) * smax + j] = next;
  }
}

template <typename T>
__device__ __forceinline__ float ld_logit(const T* p, int64_t i);
template <>
__device__ __forceinline__ float ld_logit<float>(const float* p, int64_t i) { return __ldg(p + i); }
template <>
__device__ __forceinline__ float ld_logit<__nv_bfloat16>(const __nv_bfloat16* p, int64_t i) { return __bfloat162float(p[i]); }

__device__ __forceinline__ void online_update(float x, int idx, float& m, float& s, float& best, int& besti) {
  if (x > best) { best = x; besti = idx; }
  if (x > m) {
    s = s * expf(m - x) + 1.0f;
    m = x;
  } else {
    s += expf(x - m);
  }
}

// One warp per frame row.  lse / frame_ids / lpx may each be null.
template <typename T>
__global__ void __launch_bounds__(256) ctc_row_stats_kernel(const T* __restrict__ logits, int64_t ld, int rows, int seq, int vocab,
                                                            const int32_t* __restrict__ lengths, float* __restrict__ lse_out,
                                                            int32_t* __restrict__ frame_ids, const int32_t* __restrict__ labels, int smax,
                                                            const int32_t* __restrict__ tlen, int blank, float* __restrict__ lpx,
                                                            const int32_t* __restrict__ cu) {
  jl::pdl_prologue();
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);       // (b, t) in the padded index space: lse / frame_ids / lpx are [B, seq]
  if (row >= rows) return;
  const int b = row / seq, t = row - b * seq;
  if (t >= lengths[b]) {
    if (lane == 0) {
      if (lse_out) lse_out[row] = 0.0f;
      if (frame_ids) frame_ids[row] = -1;
    }
    return;
  }
  // packed layout: the frame's logits are row cu[b] + t (no padding rows in the logits matrix)
  const T* x = logits + (cu != nullptr ? static_cast<int64_t>(cu[b]) + t : static_cast<int64_t>(row)) * ld;
  // The 2S+1 label logits of the row are gathered BEFORE the streaming pass (raw, lse subtracted at the end): the sectors
  // they touch are then re-read by the stream within microseconds, from L2.  Gathered after the pass, a fifth of them had
  // already been evicted by the other rows in flight (ncu: 192 MB of DRAM reads for 160 MB of logits).
  if (lpx != nullptr) {
    const int L = 2 * tlen[b] + 1;
    float* dst = lpx + static_cast<int64_t>(row) * (2 * smax + 1);
    for (int sidx = lane; sidx < L; sidx += 32) {
      const int c = (sidx & 1) ? labels[static_cast<int64_t>(b) * smax + (sidx >> 1)] : blank;
      dst[sidx] = ld_logit<T>(x, c);
    }
  }
  float m = -CUDART_INF_F, s = 0.0f, best = -CUDART_INF_F;
  int besti = 0x7fffffff;
  constexpr int VEC = 16 / sizeof(T);
  const bool vec_ok = ((ld % VEC) == 0) && ((reinterpret_cast<uintptr_t>(logits) & 15) == 0);
  if (vec_ok) {
    const int nvec = vocab / VEC;
    const uint4* xv = reinterpret_cast<const uint4*>(x);
    // Four 16-byte loads per lane in flight; per group of 4·VEC values: one running-max update (a rescale of the sum only
    // when the maximum moves — a handful of times per row), then one FADD + FMUL + MUFU.EX2 + FADD per value.  The argmax
    // is searched element by element only in the groups that raise it.
    constexpr int U = 4;
    int i = lane;
    for (; i + 32 * (U - 1) < nvec; i += 32 * U) {
      uint4 raw[U];
#pragma unroll
      for (int u = 0; u < U; ++u) raw[u] = __ldg(xv + i + 32 * u);
      float val[U][VEC];
      float cm = -CUDART_INF_F;
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if constexpr (sizeof(T) == 4) {
          val[u][0] = __uint_as_float(raw[u].x); val[u][1] = __uint_as_float(raw[u].y);
          val[u][2] = __uint_as_float(raw[u].z); val[u][3] = __uint_as_float(raw[u].w);
        } else {
          const uint32_t w[4] = {raw[u].x, raw[u].y, raw[u].z, raw[u].w};
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float2 f = unpack_bf16x2(w[q]);
            val[u][2 * q] = f.x; val[u][2 * q + 1] = f.y;
          }
        }
#pragma unroll
        for (int q = 0; q < VEC; ++q) cm = fmaxf(cm, val[u][q]);
      }
      if (cm > best) {                       // indices grow with u and q: strict > keeps the first maximum
#pragma unroll
        for (int u = 0; u < U; ++u)
#pragma unroll
          for (int q = 0; q < VEC; ++q)
            if (val[u][q] > best) { best = val[u][q]; besti = (i + 32 * u) * VEC + q; }
      }
      if (cm > m) {                          // cm is finite here; m = -inf on the first group gives ex2(-inf) = 0
        s *= mufu_ex2((m - cm) * TC_LOG2E_F);
        m = cm;
      }
      if (m != -CUDART_INF_F) {              // a group of -inf only (fully masked logits) adds nothing
#pragma unroll
        for (int u = 0; u < U; ++u)
#pragma unroll
          for (int q = 0; q < VEC; ++q) s += mufu_ex2((val[u][q] - m) * TC_LOG2E_F);
      }
    }
    for (; i < nvec; i += 32) {
      const uint4 v = __ldg(xv + i);
      if constexpr (sizeof(T) == 4) {
        online_update(__uint_as_float(v.x), i * 4 + 0, m, s, best, besti);
        online_update(__uint_as_float(v.y), i * 4 + 1, m, s, best, besti);
        online_update(__uint_as_float(v.z), i * 4 + 2, m, s, best, besti);
        online_update(__uint_as_float(v.w), i * 4 + 3, m, s, best, besti);
      } else {
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float2 f = unpack_bf16x2(w[q]);
          online_update(f.x, i * 8 + 2 * q, m, s, best, besti);
          online_update(f.y, i * 8 + 2 * q + 1, m, s, best, besti);
        }
      }
    }
    for (int i = nvec * VEC + lane; i < vocab; i += 32) online_update(ld_logit<T>(x, i), i, m, s, best, besti);
  } else {
    for (int i = lane; i < vocab; i += 32) online_update(ld_logit<T>(x, i), i, m, s, best, besti);
  }
  // warp combine: (m, s) by rescaling; argmax with "first max wins"
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float m2 = __shfl_xor_sync(0xffffffffu, m, o);
    const float s2 = __shfl_xor_sync(0xffffffffu, s, o);
    const float b2 = __shfl_xor_sync(0xffffffffu, best, o);
    const int i2 = __shfl_xor_sync(0xffffffffu, besti, o);
    const float mm = fmaxf(m, m2);
    const float sa = (m == -CUDART_INF_F) ? 0.0f : s * expf(m - mm);
    const float sb = (m2 == -CUDART_INF_F) ? 0.0f : s2 * expf(m2 - mm);
    s = sa + sb;
    m = mm;
    if (b2 > best || (b2 == best && i2 < besti)) { best = b2; besti = i2; }
  }
  const float lse = m + logf(s);
  if (lane == 0) {
    if (lse_out) lse_out[row] = lse;
    if (frame_ids) frame_ids[row] = besti;
  }
  if (lpx != nullptr) {
    const int S = tlen[b];
    const int L = 2 * S + 1;
    const int Lmax = 2 * smax + 1;
    float* dst = lpx + static_cast<int64_t>(row) * Lmax;
    for (int sidx = lane; sidx < L; sidx += 32) dst[sidx] -= lse;      // same lane wrote dst[sidx] above
  }
}

__device__ __forceinline__ float lse3(float a, float b, float c) {
  const float m = fmaxf(a, fmaxf(b, c));
  if (m == -CUDART_INF_F) return -CUDART_INF_F;
  return m + logf(expf(a - m) + expf(b - m) + expf(c - m));
}

// log(e^a + e^b + e^c) for the lattice recursion with one MUFU per exp / log (ex2.approx, lg2.approx).  The values are
// log-probabilities of magnitude 10²–10³ whose fp32 spacing (≈ 6e-5) is far coarser than the 2⁻²² error of the
// approximations, so the recursion's accuracy is unchanged; the recursion is serial, and its step time is the length of
// this dependent instruction chain.
__device__ __forceinline__ float lse3_fast(float a, float b, float c) {
  const float m = fmaxf(a, fmaxf(b, c));
  const float ms = (m == -CUDART_INF_F) ? 0.0f : m;          // differences first: (x - m) is exact or nearly so, x·log2e - m·log2e is not
  float ea, eb, ec, lg;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(ea) : "f"((a - ms) * TC_LOG2E_F));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(eb) : "f"((b - ms) * TC_LOG2E_F));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(ec) : "f"((c - ms) * TC_LOG2E_F));
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(lg) : "f"(ea + eb + ec));
  return (m == -CUDART_INF_F) ? -CUDART_INF_F : fmaf(lg, TC_LN2_F, m);
}

// One CTA per (utterance, direction): blockIdx.y = 0 runs the alpha recursion forward in time, blockIdx.y = 1 the beta
// recursion backward, so the two directions of an utterance run on different SMs.  The recursion is serial in t and its
// step time is the dependent instruction chain of one state update (few warps per SM: nothing hides ALU latency), so
// the common case — every thread owns one state (SINGLE) — keeps its neighbour indices, the skip-transition predicate and
// all pointers in registers and does no per-step index arithmetic.
template <bool SINGLE>
__global__ void __launch_bounds__(CTC_LATTICE_HALF) ctc_lattice_kernel(const int32_t* __restrict__ labels, int smax,
                                                                       const int32_t* __restrict__ tlen,
                                                                       const int32_t* __restrict__ lengths, int seq, int blank,
                                                                       const float* __restrict__ lpx, float* __restrict__ alpha,
                                                                       float* __restrict__ beta, float* __restrict__ nll) {
  jl::pdl_prologue();
  extern __shared__ float lat_smem[];
  const int b = blockIdx.x;
  const bool is_beta = blockIdx.y != 0;
  const int S = tlen[b];
  const int L = 2 * S + 1;
  const int Lmax = 2 * smax + 1;
  const int T = min(lengths[b], seq);
  float* buf = lat_smem;                                      // [2][Lmax]
  int* ext = reinterpret_cast<int*>(lat_smem + 2 * Lmax);     // [Lmax]
  float* lpring = lat_smem + 3 * Lmax;                        // [CTC_LAT_AHEAD][Lmax]
  const int nthr = blockDim.x;                                // the host sizes the CTA to the longest extended label sequence
  const int gtid = threadIdx.x;
  for (int s = gtid; s < L; s += nthr) ext[s] = (s & 1) ? labels[static_cast<int64_t>(b) * smax + (s >> 1)] : blank;
  __syncthreads();
  if (T == 0) {
    if (!is_beta && gtid == 0) nll[b] = (L > 1) ? CUDART_INF_F : 0.0f;
    return;
  }
  const int64_t base = static_cast<int64_t>(b) * seq * Lmax;
  const int dir = is_beta ? -Lmax : Lmax;                     // one frame along the recursion (elements)
  const int t_first = is_beta ? T - 1 : 0;
  const float* lp_row = lpx + base + static_cast<int64_t>(t_first) * Lmax;      // row of the current frame
  float* out_row = (is_beta ? beta : alpha) + base + static_cast<int64_t>(t_first) * Lmax;

  // t = 0 (alpha) / t = T-1 (beta)
  for (int s = gtid; s < L; s += nthr) {
    float v = -CUDART_INF_F;
    if (!is_beta) {
      if (s <= 1) v = lp_row[s];
    } else {
      if (s >= L - 2) v = lp_row[s];
    }
    buf[s] = v;
    out_row[s] = v;
  }
  __syncthreads();
  // A frame's log-probs must already be on the SM when its step starts: they are streamed into a shared-memory ring
  // CTC_LAT_AHEAD frames ahead with cp.async (each thread copies exactly the elements it will read itself, so its own
  // wait_group is the only synchronisation the ring needs).
#pragma unroll 1
  for (int u = 0; u < CTC_LAT_AHEAD; ++u) {
    if (1 + u < T)
      for (int s = gtid; s < L; s += nthr) cp_async_f32(lpring + u * Lmax + s, lp_row + (1 + u) * dir + s);
    cp_async_commit();
  }
  const int nb = is_beta ? 1 : -1;                            // neighbour offset along the recursion

  if constexpr (SINGLE) {
    const int s = gtid;
    const bool active = s < L;
    const int s1 = s + nb, s2 = s + 2 * nb;
    const bool has1 = active && s1 >= 0 && s1 < L;
    const bool has2 = active && s2 >= 0 && s2 < L && (s & 1) && ext[s] != ext[s2];
    // shared-memory byte addresses of this thread's cells in the two row buffers and in the ring
    const float* p0 = buf + s;
    const float* p1 = buf + (has1 ? s1 : s);
    const float* p2 = buf + (has2 ? s2 : s);
    const float* lp_src = lp_row + CTC_LAT_AHEAD * dir + s;   // + dir per step → row (step + AHEAD)
    float* out_ptr = out_row + s;
    int slot = 0, par = 0;                                     // par: buffer holding the previous row
    for (int step = 1; step < T; ++step) {
      lp_src += dir;
      out_ptr += dir;
      cp_async_wait<CTC_LAT_AHEAD - 1>();
      if (active) {
        const int po = par * Lmax;
        const float x0 = p0[po];
        const float x1 = has1 ? p1[po] : -CUDART_INF_F;
        const float x2 = has2 ? p2[po] : -CUDART_INF_F;
        float* lps = lpring + slot * Lmax + s;
        const float v = lse3_fast(x0, x1, x2) + *lps;
        buf[(par ^ 1) * Lmax + s] = v;
        *out_ptr = v;
        if (step + CTC_LAT_AHEAD < T) cp_async_f32(lps, lp_src);
      }
      cp_async_commit();
      slot = (slot + 1 == CTC_LAT_AHEAD) ? 0 : slot + 1;
      par ^= 1;
      __syncthreads();
    }
  } else {
    int slot = 0;
    for (int step = 1; step < T; ++step) {
      lp_row += dir;
      out_row += dir;
      const float* prev = buf + ((step - 1) & 1) * Lmax;
      float* cur = buf + (step & 1) * Lmax;
      float* lps = lpring + slot * Lmax;
      cp_async_wait<CTC_LAT_AHEAD - 1>();                     // this step's frame has landed (groups complete in order)
      for (int s = gtid; s < L; s += nthr) {
        const int s1 = s + nb, s2 = s + 2 * nb;
        float x0 = prev[s], x1 = -CUDART_INF_F, x2 = -CUDART_INF_F;
        if (s1 >= 0 && s1 < L) x1 = prev[s1];
        if (s2 >= 0 && s2 < L && (s & 1) && ext[s] != ext[s2]) x2 = prev[s2];
        const float v = lse3_fast(x0, x1, x2) + lps[s];
        cur[s] = v;
        out_row[s] = v;
      }
      if (step + CTC_LAT_AHEAD < T)                           // refill the slot just consumed (same thread, same elements)
        for (int s = gtid; s < L; s += nthr) cp_async_f32(lps + s, lp_row + CTC_LAT_AHEAD * dir + s);
      cp_async_commit();
      slot = (slot + 1 == CTC_LAT_AHEAD) ? 0 : slot + 1;
      __syncthreads();
    }
  }
  if (!is_beta && gtid == 0) {
    const float* last = buf + ((T - 1) & 1) * Lmax;
    const float l1 = last[L - 1];
    const float l2 = (L > 1) ? last[L - 2] : -CUDART_INF_F;
    nll[b] = -lse3(l1, l2, -CUDART_INF_F);
  }
}

template <typename TG>
__device__ __forceinline__ void st_grad(TG* p, int64_t i, float v);
template <>
__device__ __forceinline__ void st_grad<float>(float* p, int64_t i, float v) { p[i] = v; }
template <>
__device__ __forceinline__ void st_grad<__nv_bfloat16>(__nv_bfloat16* p, int64_t i, float v) { p[i] = __float2bfloat16_rn(v); }

// One CTA per frame row.
template <typename T, typename TG>
__global__ void __launch_bounds__(CTC_GRAD_THREADS) ctc_grad_kernel(const T* __restrict__ logits, int64_t ld, TG* __restrict__ grad, int64_t ldg,
                                                                   int seq, int vocab, int batch, const int32_t* __restrict__ lengths,
                                                                   const int32_t* __restrict__ labels, const int32_t* __restrict__ nxt,
                                                                   const int32_t* __restrict__ fst, const int32_t* __restrict__ nblank,
                                                                   int smax, const int32_t* __restrict__ tlen, int blank,
                                                                   const float* __restrict__ lse_all, const float* __restrict__ lpx,
                                                                   const float* __restrict__ alpha, const float* __restrict__ beta,
                                                                   const float* __restrict__ nll, int reduction, int zero_infinity,
                                                                   const int32_t* __restrict__ cu) {
  jl::pdl_prologue();
  extern __shared__ float grad_smem[];
  const int row = blockIdx.x;
  const int b = row / seq, t = row - b * seq;
  const int tid = threadIdx.x;
  if (cu != nullptr && t >= lengths[b]) return;               // packed layout: a padded (b, t) has no row at all
  const int64_t mrow = (cu != nullptr) ? static_cast<int64_t>(cu[b]) + t : static_cast<int64_t>(row);   // row in logits / grad
  TG* g = grad + mrow * ldg;
  const float nll_b = nll[b];
  const bool infeasible = isinf(nll_b);
  if (t >= lengths[b] || (infeasible && zero_infinity)) {
    for (int v = tid; v < vocab; v += CTC_GRAD_THREADS) st_grad<TG>(g, v, 0.0f);
    return;
  }
  if (infeasible) {
    for (int v = tid; v < vocab; v += CTC_GRAD_THREADS) st_grad<TG>(g, v, CUDART_NAN_F);
    return;
  }
  const int S = tlen[b];
  const int L = 2 * S + 1;
  const int Lmax = 2 * smax + 1;
  float* occ = grad_smem;                               // [Lmax]
  float* blank_part = grad_smem + Lmax;                 // [32]
  const float scale = (reduction == JL_CTC_MEAN) ? 1.0f / (static_cast<float>(max(S, 1)) * static_cast<float>(batch)) : 1.0f;
  const int64_t lat = static_cast<int64_t>(row) * Lmax;
  for (int s = tid; s < L; s += CTC_GRAD_THREADS) {
    const float e = alpha[lat + s] + beta[lat + s] - lpx[lat + s] + nll_b;
    const float o = expf(e);
    occ[s] = (isfinite(o)) ? o : 0.0f;
  }
  __syncthreads();
  const T* x = logits + mrow * ld;
  const float lse = lse_all[row];
  // softmax part, written once (4 elements per thread per step when the rows are 16-byte aligned)
  if ((ld & 3) == 0 && (ldg & 3) == 0 && (reinterpret_cast<uintptr_t>(logits) & 15) == 0 && (reinterpret_cast<uintptr_t>(grad) & 15) == 0) {
    const int nv = vocab >> 2;
    for (int i = tid; i < nv; i += CTC_GRAD_THREADS) {
      float a[4];
      if constexpr (sizeof(T) == 4) {
        const float4 q = __ldg(reinterpret_cast<const float4*>(x) + i);
        a[0] = q.x; a[1] = q.y; a[2] = q.z; a[3] = q.w;
      } else {
        const uint2 q = __ldg(reinterpret_cast<const uint2*>(x) + i);
        const float2 f0 = unpack_bf16x2(q.x), f1 = unpack_bf16x2(q.y);
        a[0] = f0.x; a[1] = f0.y; a[2] = f1.x; a[3] = f1.y;
      }
#pragma unroll
      for (int e = 0; e < 4; ++e) a[e] = __expf(a[e] - lse) * scale;     // arguments <= 0; 2-ulp ex2.approx is far inside the tolerance
      if constexpr (sizeof(TG) == 4) {
        reinterpret_cast<float4*>(g)[i] = make_float4(a[0], a[1], a[2], a[3]);
      } else {
        uint2 o;
        o.x = pack_bf16x2(a[0], a[1]);
        o.y = pack_bf16x2(a[2], a[3]);
        reinterpret_cast<uint2*>(g)[i] = o;
      }
    }
    for (int v = nv * 4 + tid; v < vocab; v += CTC_GRAD_THREADS) st_grad<TG>(g, v, expf(ld_logit<T>(x, v) - lse) * scale);
  } else {
    for (int v = tid; v < vocab; v += CTC_GRAD_THREADS) st_grad<TG>(g, v, expf(ld_logit<T>(x, v) - lse) * scale);
  }
  // blank: fixed-order sum over the even states (warp 0: strided partials in lane order, then a shuffle tree)
  if (tid < 32) {
    float part = 0.0f;
    for (int s = 2 * tid; s < L; s += 64) part += occ[s];
    part = warp_sum(part);
    if (tid == 0) blank_part[0] = part;
  }
  __syncthreads();   // also orders the softmax row writes before the fix-ups below
  // non-blank labels: the first occurrence of each label owns the sum over its repeats, taken in position order along the
  // chain ctc_prep linked (deterministic; one writer per vocabulary column)
  const int64_t lrow = static_cast<int64_t>(b) * smax;
  for (int j = tid; j < S; j += CTC_GRAD_THREADS) {
    const int c = labels[lrow + j];
    if (!fst[lrow + j] || c == blank) continue;
    float tot = 0.0f;
    for (int jj = j; jj >= 0; jj = nxt[lrow + jj]) tot += occ[2 * jj + 1];
    st_grad<TG>(g, c, (expf(ld_logit<T>(x, c) - lse) - tot) * scale);
  }
  if (tid == 0) {
    // a label equal to the blank index folds into the blank column (degenerate but well defined)
    float tot = blank_part[0];
    if (nblank[b] > 0)
      for (int s2 = 1; s2 < L; s2 += 2)
        if (labels[lrow + (s2 >> 1)] == blank) tot += occ[s2];
    st_grad<TG>(g, blank, (expf(ld_logit<T>(x, blank) - lse) - tot) * scale);
  }
}

__global__ void ctc_reduce_kernel(float* __restrict__ nll, const int32_t* __restrict__ tlen, int batch, int reduction, int zero_infinity,
                                  float* __restrict__ loss) {
  jl::pdl_prologue();
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  float acc = 0.0f;
  for (int b = 0; b < batch; ++b) {
    float v = nll[b];
    if (zero_infinity && isinf(v)) {
      v = 0.0f;
      nll[b] = 0.0f;
    }
    acc += (reduction == JL_CTC_MEAN) ? v / static_cast<float>(max(tlen[b], 1)) : v;
  }
  if (loss) *loss = (reduction == JL_CTC_MEAN) ? acc / static_cast<float>(batch) : acc;
}

// Greedy: collapse consecutive repeats of the per-frame argmax, drop blank, compact.  One warp per utterance.
__global__ void ctc_collapse_kernel(const int32_t* __restrict__ frame_ids, const int32_t* __restrict__ lengths, int seq, int blank,
                                    int32_t* __restrict__ out_ids, int32_t* __restrict__ out_lengths) {
  jl::pdl_prologue();
  const int b = blockIdx.x, lane = threadIdx.x;
  const int T = min(lengths[b], seq);
  const int32_t* ids = frame_ids + static_cast<int64_t>(b) * seq;
  int32_t* out = out_ids + static_cast<int64_t>(b) * seq;
  int base = 0;
  for (int t0 = 0; t0 < T; t0 += 32) {
    const int t = t0 + lane;
    bool keep = false;
    int id = -1;
    if (t < T) {
      id = ids[t];
      keep = (id != blank) && (t == 0 || ids[t - 1] != id);
    }
    const unsigned mask = __ballot_sync(0xffffffffu, keep);
    if (keep) out[base + __popc(mask & ((1u << lane) - 1u))] = id;
    base += __popc(mask);
  }
  for (int i = base + lane; i < seq; i += 32) out[i] = -1;
  if (lane == 0) out_lengths[b] = base;
}

// f1: frame argmax from the per-chunk (max, argmax) pairs of a JL_EPI_ARGMAX lm_head GEMM.  Thread = frame (b, t); the chunk-major
// layout makes every load of a warp coalesced.  Chunks are visited in column order and only a strictly larger maximum replaces
// the current one, so ties go to the lowest column, as torch.argmax.
__global__ void __launch_bounds__(256) ctc_argmax_reduce_kernel(const float* __restrict__ pmax, const int32_t* __restrict__ pidx, int64_t ld,
                                                                int num_chunks, const int32_t* __restrict__ lengths,
                                                                const int32_t* __restrict__ cu, int batch, int seq,
                                                                int32_t* __restrict__ frame_ids) {
  jl::pdl_prologue();
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= batch * seq) return;
  const int b = idx / seq, t = idx - b * seq;
  if (t >= lengths[b]) {
    frame_ids[idx] = -1;
    return;
  }
  const int64_t row = (cu != nullptr) ? static_cast<int64_t>(cu[b]) + t : static_cast<int64_t>(idx);
  float best = -CUDART_INF_F;
  int besti = 0x7fffffff;
  for (int c = 0; c < num_chunks; ++c) {
    const float v = __ldg(pmax + c * ld + row);
    if (v > best) {
      best = v;
      besti = __ldg(pidx + c * ld + row);
    }
  }
  frame_ids[idx] = besti;
}

static int ctc_validate_common(const void* logits, const int32_t* lengths, int batch, int seq, int vocab, int blank, int dtype) {
  JL_REQUIRE(logits && lengths, JL_EINVAL, "ctc: null logits / input_lengths");
  JL_REQUIRE(batch > 0 && seq > 0 && vocab > 1, JL_EINVAL, "ctc: batch, seq must be positive and vocab > 1");
  JL_REQUIRE(blank >= 0 && blank < vocab, JL_EINVAL, "ctc: blank %d outside [0, %d)", blank, vocab);
  JL_REQUIRE(dtype == JL_DT_BF16 || dtype == JL_DT_F32, JL_EINVAL, "ctc: unknown logits dtype %d", dtype);
  return JL_OK;
}

}  // namespace jl

extern "C" {

int jl_ctc_workspace_bytes(const jl_ctc_params* p, size_t* out) {
  JL_REQUIRE(p && out, JL_EINVAL, "ctc_workspace_bytes: null argument");
  JL_REQUIRE(p->batch > 0 && p->seq > 0 && p->max_label_len >= 0, JL_EINVAL, "ctc_workspace_bytes: bad shape");
  jl::carve_ws(p, nullptr, out);
  return JL_OK;
}

int jl_ctc_fwd(const jl_ctc_params* p, void* workspace, void* stream) {
  JL_REQUIRE(p != nullptr, JL_EINVAL, "ctc: null params");
  int rc = jl::ctc_validate_common(p->logits, p->input_lengths, p->batch, p->seq, p->vocab, p->blank, p->logits_dtype);
  if (rc != JL_OK) return rc;
  JL_REQUIRE(p->labels != nullptr || p->max_label_len == 0, JL_EINVAL, "ctc: null labels");
  JL_REQUIRE(p->nll != nullptr && workspace != nullptr, JL_EINVAL, "ctc: nll and workspace are required");
  JL_REQUIRE(p->reduction == JL_CTC_SUM || p->reduction == JL_CTC_MEAN, JL_EINVAL, "ctc: unknown reduction %d", p->reduction);
  JL_REQUIRE(p->max_label_len >= 0 && 2 * p->max_label_len + 1 <= 8192, JL_EUNSUPPORTED_SHAPE, "ctc: max_label_len %d too large", p->max_label_len);
  if (p->grad) JL_REQUIRE(p->grad_dtype == JL_DT_BF16 || p->grad_dtype == JL_DT_F32, JL_EINVAL, "ctc: unknown grad dtype");
  rc = jl::check_device();
  if (rc != JL_OK) return rc;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const jl::CtcWs w = jl::carve_ws(p, workspace, nullptr);
  const int smax = p->max_label_len > 0 ? p->max_label_len : 1;
  const int Lmax = 2 * smax + 1;
  const int rows = p->batch * p->seq;

  if (p->max_label_len > 0) {
    jl::launch(jl::ctc_prep_kernel, p->batch, jl::CTC_PREP_THREADS, 0, s, p->labels, smax, p->vocab, p->blank, w.labels, w.nxt, w.fst, w.nblank,
               w.tlen);
    JL_CHECK_LAUNCH("ctc_prep");
  } else {
    cudaMemsetAsync(w.tlen, 0, sizeof(int32_t) * p->batch, s);
  }
  const int sblocks = jl::ceil_div(rows, 8);
  if (p->logits_dtype == JL_DT_F32)
    jl::launch(jl::ctc_row_stats_kernel<float>, sblocks, 256, 0, s, reinterpret_cast<const float*>(p->logits), p->ld_logits, rows, p->seq, p->vocab,
                                                           p->input_lengths, w.lse, nullptr, w.labels, smax, w.tlen, p->blank, w.lpx, p->cu_seqlens);
  else
    jl::launch(jl::ctc_row_stats_kernel<__nv_bfloat16>, sblocks, 256, 0, s, reinterpret_cast<const __nv_bfloat16*>(p->logits), p->ld_logits, rows, p->seq,
                                                                   p->vocab, p->input_lengths, w.lse, nullptr, w.labels, smax, w.tlen,
                                                                   p->blank, w.lpx, p->cu_seqlens);
  JL_CHECK_LAUNCH("ctc_row_stats");
  const size_t lat_smem = static_cast<size_t>(3 + jl::CTC_LAT_AHEAD) * Lmax * sizeof(float);
  const bool lat_single = Lmax <= jl::CTC_LATTICE_HALF;       // every thread owns one state of the extended label sequence
  auto lat_kernel = lat_single ? jl::ctc_lattice_kernel<true> : jl::ctc_lattice_kernel<false>;
  if (lat_smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(lat_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(lat_smem));
    JL_REQUIRE(e == cudaSuccess, JL_ECUDA, "ctc: cannot reserve lattice shared memory: %s", cudaGetErrorString(e));
  }
  const int lat_half = std::min(jl::CTC_LATTICE_HALF, ((2 * smax + 1 + 31) / 32) * 32);
  jl::launch(lat_kernel, dim3(p->batch, 2), lat_half, lat_smem, s, w.labels, smax, w.tlen, p->input_lengths, p->seq, p->blank, w.lpx,
             w.alpha, w.beta, p->nll);
  JL_CHECK_LAUNCH("ctc_lattice");
  if (p->grad != nullptr) {
    const size_t gsm = static_cast<size_t>(Lmax + 32) * sizeof(float);
#define JL_CTC_GRAD(TL, TGR)                                                                                                           \
  jl::launch(jl::ctc_grad_kernel<TL, TGR>, rows, jl::CTC_GRAD_THREADS, gsm, s, reinterpret_cast<const TL*>(p->logits), p->ld_logits,              \
                                                                        reinterpret_cast<TGR*>(p->grad), p->ld_grad, p->seq, p->vocab,   \
                                                                        p->batch, p->input_lengths, w.labels, w.nxt, w.fst, w.nblank,    \
                                                                        smax, w.tlen, p->blank,                                          \
                                                                        w.lse, w.lpx, w.alpha, w.beta, p->nll, p->reduction,             \
                                                                        p->zero_infinity, p->cu_seqlens)
    if (p->logits_dtype == JL_DT_F32 && p->grad_dtype == JL_DT_F32) JL_CTC_GRAD(float, float);
    else if (p->logits_dtype == JL_DT_F32) JL_CTC_GRAD(float, __nv_bfloat16);
    else if (p->grad_dtype == JL_DT_F32) JL_CTC_GRAD(__nv_bfloat16, float);
    else JL_CTC_GRAD(__nv_bfloat16, __nv_bfloat16);
#undef JL_CTC_GRAD
    JL_CHECK_LAUNCH("ctc_grad");
  }
  jl::launch(jl::ctc_reduce_kernel, 1, 32, 0, s, p->nll, w.tlen, p->batch, p->reduction, p->zero_infinity, p->loss);
  JL_CHECK_LAUNCH("ctc_reduce");
  return JL_OK;
}

int jl_ctc_greedy_from_partials(const float* pmax, const int32_t* pidx, int64_t ld, int32_t num_chunks, const int32_t* input_lengths,
                                const int32_t* cu_seqlens, int32_t batch, int32_t seq, int32_t blank, int32_t* frame_ids,
                                int32_t* out_ids, int32_t* out_lengths, void* stream) {
  JL_REQUIRE(pmax && pidx && input_lengths && frame_ids && out_ids && out_lengths, JL_EINVAL, "ctc_greedy_from_partials: null pointer");
  JL_REQUIRE(batch > 0 && seq > 0 && num_chunks > 0 && ld > 0 && blank >= 0, JL_EINVAL, "ctc_greedy_from_partials: bad dims");
  int rc = jl::check_device();
  if (rc != JL_OK) return rc;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  jl::launch(jl::ctc_argmax_reduce_kernel, jl::ceil_div(batch * seq, 256), 256, 0, s, pmax, pidx, ld, num_chunks, input_lengths, cu_seqlens, batch,
             seq, frame_ids);
  JL_CHECK_LAUNCH("ctc_argmax_reduce");
  jl::launch(jl::ctc_collapse_kernel, batch, 32, 0, s, frame_ids, input_lengths, seq, blank, out_ids, out_lengths);
  JL_CHECK_LAUNCH("ctc_collapse");
  return JL_OK;
}

int jl_ctc_greedy(const jl_ctc_greedy_params* p, void* stream) {
  JL_REQUIRE(p != nullptr, JL_EINVAL, "ctc_greedy: null params");
  int rc = jl::ctc_validate_common(p->logits, p->input_lengths, p->batch, p->seq, p->vocab, p->blank, p->logits_dtype);
  if (rc != JL_OK) return rc;
  JL_REQUIRE(p->frame_ids && p->out_ids && p->out_lengths, JL_EINVAL, "ctc_greedy: null output pointer");
  rc = jl::check_device();
  if (rc != JL_OK) return rc;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const int rows = p->batch * p->seq;
  const int sblocks = jl::ceil_div(rows, 8);
  if (p->logits_dtype == JL_DT_F32)
    jl::launch(jl::ctc_row_stats_kernel<float>, sblocks, 256, 0, s, reinterpret_cast<const float*>(p->logits), p->ld_logits, rows, p->seq, p->vocab,
                                                           p->input_lengths, nullptr, p->frame_ids, nullptr, 1, nullptr, p->blank, nullptr, p->cu_seqlens);
  else
    jl::launch(jl::ctc_row_stats_kernel<__nv_bfloat16>, sblocks, 256, 0, s, reinterpret_cast<const __nv_bfloat16*>(p->logits), p->ld_logits, rows, p->seq,
                                                                   p->vocab, p->input_lengths, nullptr, p->frame_ids, nullptr, 1, nullptr,
                                                                   p->blank, nullptr, p->cu_seqlens);
  JL_CHECK_LAUNCH("ctc_argmax");
  jl::launch(jl::ctc_collapse_kernel, p->batch, 32, 0, s, p->frame_ids, p->input_lengths, p->seq, p->blank, p->out_ids, p->out_lengths);
  JL_CHECK_LAUNCH("ctc_collapse");
  return JL_OK;
}

}  // extern "C"
