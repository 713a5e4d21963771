// a1 + a2: Kaldi-compatible 80-bin log-mel filterbank + utterance CMVN, waveform resident in HBM.
//
// Replaces SP/torchaudio/compliance/kaldi.py:514-645 (fbank: frames 400/160 snip_edges :63-67, DC removal
// :183-186, pre-emphasis :193-198, povey window :98-100,201-204, zero-pad to 512 :207-211, |rfft|² :616-618,
// mel projection :630, log(max(·, eps)) :633) as called with ×2^15 scaling from
// SP/transformers/models/speech_to_text/feature_extraction_speech_to_text.py:104-120, and utterance_cmvn + padding +
// attention mask at :142-163, :275-303.
//
// Kernel 1 (mel_fbank_kernel): one CTA = 32 consecutive frames of one utterance.  The 5360 samples the frames
//   share are staged once in shared memory with coalesced loads (each sample is read from HBM exactly once).
//   The 512-point real FFT of a frame is a 256-point complex FFT held in REGISTERS by a half-warp: lane l owns
//   z[l + 16 j], j = 0..15 (z[n] = y[2n] + i·y[2n+1]); radix-16 butterflies in registers → lane twiddles
//   W256^(l·k) (registers) → ONE 16 × 16 transpose through shared memory → radix-16 again → the lane holds
//   Z[l + 16 j].  The real-FFT untangling needs Z[256 − k], which lives in lane 16 − l: one shuffle per bin.
//   A warp carries two frames (f, f + 16) at a time; the 257 power bins of every frame go to shared memory and
//   the mel projection then runs with lane = frame: the filter weight is a warp-uniform load and the power bins
//   are read conflict-free (row stride 257), i.e. 2 shared-memory wavefronts per tap for 32 frames.
//   Shared-memory wavefronts per frame: ≈ 115 (the shared-memory Stockham version this replaces: ≈ 500, and the
//   LSU was its limiter — profiles/README.md §5).
//   The CTA writes its [32, 80] tile coalesced and leaves per-bin (mean, M2) partials for the CMVN statistics.
// Kernel 2 (cmvn_kernel): merges the utterance's partials in fixed order (Chan's parallel variance update —
//   deterministic, no atomics), normalises the tile, zeroes padded frames, optionally emits a bf16 copy.
#include "common.cuh"

namespace jl {

constexpr int MEL_FRAME_LEN = 400;
constexpr int MEL_FRAME_SHIFT = 160;
constexpr int MEL_FPC = JL_MEL_FRAMES_PER_CTA;                              // frames per CTA
constexpr int MEL_SAMPLES_PER_CTA = (MEL_FPC - 1) * MEL_FRAME_SHIFT + MEL_FRAME_LEN;   // 5360
constexpr int MEL_WARPS = 8;
constexpr int MEL_THREADS = MEL_WARPS * 32;
constexpr float MEL_PREEMPH = 0.97f;
constexpr float MEL_FLT_EPS = 1.1920928955078125e-07f;
constexpr int MEL_TROW = 17;                    // float2 per row of the 16 x 16 transpose tile (17: conflict-free both ways)
constexpr int MEL_PSTRIDE = 257;                // floats per frame of power bins (odd: lane = frame reads hit 32 banks)
constexpr int MEL_OSTRIDE = JL_MEL_BINS + 1;    // floats per frame of log-mel outputs
static_assert(MEL_FPC == 32, "mel_fbank_kernel maps one lane to one frame in the mel projection");

struct MelSmem {
  float wave[MEL_SAMPLES_PER_CTA];
  float window[MEL_FRAME_LEN];
  float2 tbuf[MEL_WARPS][2][16 * MEL_TROW];     // per warp, per frame of the pair: transpose tile
  float pw[MEL_FPC][MEL_PSTRIDE];               // power spectrum of every frame of the tile
  float out[MEL_FPC][MEL_OSTRIDE];
};

__device__ __forceinline__ float2 cmul(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }

__device__ __forceinline__ int utt_frames(int n) { return n < MEL_FRAME_LEN ? 0 : 1 + (n - MEL_FRAME_LEN) / MEL_FRAME_SHIFT; }

// 4-point DFT (forward, e^{-2 pi i nk/4}): no multiplications
__device__ __forceinline__ void bfly4(float2 a0, float2 a1, float2 a2, float2 a3, float2& y0, float2& y1, float2& y2, float2& y3) {
  const float2 s02 = cadd(a0, a2), d02 = csub(a0, a2), s13 = cadd(a1, a3), d13 = csub(a1, a3);
  y0 = cadd(s02, s13);
  y1 = make_float2(d02.x + d13.y, d02.y - d13.x);     // d02 - i d13
  y2 = csub(s02, s13);
  y3 = make_float2(d02.x - d13.y, d02.y + d13.x);     // d02 + i d13
}
// a · W16^E, W16 = e^{-2 pi i / 16}
template <int E>
__device__ __forceinline__ float2 mul_w16(float2 a) {
  constexpr float C1 = 0.92387953251128674f, S1 = 0.38268343236508977f, R = 0.70710678118654752f;
  if constexpr (E == 0) return a;
  else if constexpr (E == 1) return make_float2(a.x * C1 + a.y * S1, a.y * C1 - a.x * S1);
  else if constexpr (E == 2) return make_float2((a.x + a.y) * R, (a.y - a.x) * R);
  else if constexpr (E == 3) return make_float2(a.x * S1 + a.y * C1, a.y * S1 - a.x * C1);
  else if constexpr (E == 4) return make_float2(a.y, -a.x);
  else if constexpr (E == 6) return make_float2((a.y - a.x) * R, -(a.x + a.y) * R);
  else { static_assert(E == 9, "unused twiddle"); return make_float2(-(a.x * C1 + a.y * S1), a.x * S1 - a.y * C1); }   // W16^9 = (-C1, +S1)
}
// 16-point DFT in registers, natural order in and out: n = n0 + 4 n1, k = k1 + 4 k0
__device__ __forceinline__ void fft16(float2 (&v)[16]) {
  float2 b[4][4];
#pragma unroll
  for (int n0 = 0; n0 < 4; ++n0) bfly4(v[n0], v[n0 + 4], v[n0 + 8], v[n0 + 12], b[n0][0], b[n0][1], b[n0][2], b[n0][3]);
  bfly4(b[0][0], b[1][0], b[2][0], b[3][0], v[0], v[4], v[8], v[12]);
  bfly4(b[0][1], mul_w16<1>(b[1][1]), mul_w16<2>(b[2][1]), mul_w16<3>(b[3][1]), v[1], v[5], v[9], v[13]);
  bfly4(b[0][2], mul_w16<2>(b[1][2]), mul_w16<4>(b[2][2]), mul_w16<6>(b[3][2]), v[2], v[6], v[10], v[14]);
  bfly4(b[0][3], mul_w16<3>(b[1][3]), mul_w16<6>(b[2][3]), mul_w16<9>(b[3][3]), v[3], v[7], v[11], v[15]);
}

// W32^k = e^{-2 pi i k / 32}, k = 0..15 (compile-time indices only)
__device__ __forceinline__ float2 w32(int k) {
  constexpr float c[16] = {1.0f, 0.98078528040323043f, 0.92387953251128674f, 0.83146961230254524f, 0.70710678118654752f, 0.55557023301960218f,
                           0.38268343236508977f, 0.19509032201612825f, 0.0f, -0.19509032201612825f, -0.38268343236508977f,
                           -0.55557023301960218f, -0.70710678118654752f, -0.83146961230254524f, -0.92387953251128674f, -0.98078528040323043f};
  constexpr float s[16] = {0.0f, 0.19509032201612825f, 0.38268343236508977f, 0.55557023301960218f, 0.70710678118654752f, 0.83146961230254524f,
                           0.92387953251128674f, 0.98078528040323043f, 1.0f, 0.98078528040323043f, 0.92387953251128674f, 0.83146961230254524f,
                           0.70710678118654752f, 0.55557023301960218f, 0.38268343236508977f, 0.19509032201612825f};
  return make_float2(c[k], -s[k]);
}

__global__ void __launch_bounds__(MEL_THREADS, 2) mel_fbank_kernel(const jl_mel_cmvn_params p, float* __restrict__ partials, int nblk) {
  jl::pdl_prologue();
  extern __shared__ __align__(16) uint8_t mel_smem_raw[];
  MelSmem& s = *reinterpret_cast<MelSmem*>(mel_smem_raw);

  const int b = blockIdx.y;
  const int blk = blockIdx.x;
  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const int n = p.num_samples[b];
  const int frames_b = min(utt_frames(n), p.max_frames);
  const int f0 = blk * MEL_FPC;
  const int nv = max(0, min(MEL_FPC, frames_b - f0));   // valid frames in this tile

  if (blk == 0 && tid == 0 && p.frame_lengths != nullptr) p.frame_lengths[b] = frames_b;
  if (p.attention_mask != nullptr && tid < MEL_FPC && f0 + tid < p.max_frames)
    p.attention_mask[static_cast<int64_t>(b) * p.max_frames + f0 + tid] = (tid < nv) ? 1 : 0;

  if (nv > 0) {
    // ---- stage the window and the shared span of samples
    for (int i = tid; i < MEL_FRAME_LEN; i += MEL_THREADS) s.window[i] = __ldg(p.window + i);
    const float* wave = p.wave + static_cast<int64_t>(b) * p.wave_stride;
    const int s0 = f0 * MEL_FRAME_SHIFT;
    const int span = (nv - 1) * MEL_FRAME_SHIFT + MEL_FRAME_LEN;     // all < n by construction
    if ((p.wave_stride & 3) == 0 && (reinterpret_cast<uintptr_t>(p.wave) & 15) == 0) {
      const float4* w4 = reinterpret_cast<const float4*>(wave + s0);  // s0 is a multiple of 160 → 16 B aligned
      float4* d4 = reinterpret_cast<float4*>(s.wave);
      for (int i = tid; i < span / 4; i += MEL_THREADS) {
        const float4 v = __ldg(w4 + i);
        d4[i] = make_float4(v.x * 32768.0f, v.y * 32768.0f, v.z * 32768.0f, v.w * 32768.0f);
      }
    } else {
      for (int i = tid; i < span; i += MEL_THREADS) s.wave[i] = __ldg(wave + s0 + i) * 32768.0f;
    }
    // lane constants: hl = position in the half-warp = residue of the lane's FFT points
    const int hl = lane & 15, fr = lane >> 4;
    const float2* tw512 = reinterpret_cast<const float2*>(p.twiddle);   // e^{-2 pi i k / 512}
    float2 twl[16];                                                      // W256^(hl·k1)
#pragma unroll
    for (int k1 = 1; k1 < 16; ++k1) twl[k1] = __ldg(tw512 + ((2 * hl * k1) & 511));
    const float2 wl = __ldg(tw512 + hl);                                 // W512^hl
    __syncthreads();

    float2* tb = s.tbuf[warp][fr];
    const int src_lane = (16 - hl) & 15;
#pragma unroll 1
    for (int it = 0; it < 2; ++it) {
      const int fa = warp + 8 * it;                  // the warp's frames: fa (lanes 0-15) and fa + 16 (lanes 16-31)
      if (fa >= nv) break;                           // warp-uniform
      const int fl = fa + 16 * fr;
      const bool valid = fl < nv;
      const float* x = s.wave + (valid ? fl : fa) * MEL_FRAME_SHIFT;
      // ---- samples 2n, 2n+1 and 2n-1 of the lane's points n = hl + 16 j; frame mean (kaldi.py:183-186)
      float2 raw[13];
      float prev[13];
      float sum = 0.0f;
#pragma unroll
      for (int j = 0; j < 13; ++j) {
        const int nn = hl + 16 * j;
        if (j < 12 || hl < 8) {
          raw[j] = *reinterpret_cast<const float2*>(x + 2 * nn);
          prev[j] = x[max(2 * nn - 1, 0)];
          sum += raw[j].x + raw[j].y;
        } else {
          raw[j] = make_float2(0.0f, 0.0f);
          prev[j] = 0.0f;
        }
      }
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
      const float mean = sum * (1.0f / MEL_FRAME_LEN);
      // ---- pre-emphasis with replicated first sample (kaldi.py:193-198), povey window, zero-pad to 512
      float2 v[16];
#pragma unroll
      for (int j = 0; j < 13; ++j) {
        const int nn = hl + 16 * j;
        if (j < 12 || hl < 8) {
          const float2 w = *reinterpret_cast<const float2*>(s.window + 2 * nn);
          const float c0 = raw[j].x - mean, c1 = raw[j].y - mean, pm = prev[j] - mean;
          v[j] = make_float2((c0 - MEL_PREEMPH * pm) * w.x, (c1 - MEL_PREEMPH * c0) * w.y);
        } else {
          v[j] = make_float2(0.0f, 0.0f);
        }
      }
#pragma unroll
      for (int j = 13; j < 16; ++j) v[j] = make_float2(0.0f, 0.0f);
      // ---- 256-point complex FFT: radix-16 over j, lane twiddles, transpose, radix-16 over the lanes' index
      fft16(v);
#pragma unroll
      for (int k1 = 1; k1 < 16; ++k1) v[k1] = cmul(v[k1], twl[k1]);
#pragma unroll
      for (int k1 = 0; k1 < 16; ++k1) tb[k1 * MEL_TROW + hl] = v[k1];
      __syncwarp();
#pragma unroll
      for (int n1 = 0; n1 < 16; ++n1) v[n1] = tb[hl * MEL_TROW + n1];
      __syncwarp();
      fft16(v);                                       // v[k2] = Z[hl + 16 k2]
      // ---- untangle the real transform: X[k] = E[k] + W512^k O[k]; power = |X|² (kaldi.py:616-618)
      float* pwf = s.pw[valid ? fl : fa];
#pragma unroll
      for (int k2 = 0; k2 < 16; ++k2) {
        float2 zp;
        zp.x = __shfl_sync(0xffffffffu, v[15 - k2].x, src_lane, 16);
        zp.y = __shfl_sync(0xffffffffu, v[15 - k2].y, src_lane, 16);
        if (hl == 0) zp = v[(16 - k2) & 15];          // lane 0 is its own partner: Z[256 - 16 k2]
        const float2 zk = v[k2];
        const float2 e = make_float2(0.5f * (zk.x + zp.x), 0.5f * (zk.y - zp.y));
        const float2 o = make_float2(0.5f * (zk.y + zp.y), -0.5f * (zk.x - zp.x));   // (zk - conj(zp)) / (2i)
        const float2 wo = cmul(cmul(wl, w32(k2)), o);                                  // W512^(hl + 16 k2)
        const float re = e.x + wo.x, im = e.y + wo.y;
        if (valid) pwf[hl + 16 * k2] = re * re + im * im;
      }
      if (valid && hl == 0) {
        const float d = v[0].x - v[0].y;              // X[256] = Re Z[0] - Im Z[0]
        pwf[256] = d * d;
      }
    }
    __syncthreads();
    // ---- mel projection + log (kaldi.py:630-633): lane = frame, the warp walks its bins; the weight load is warp-uniform
    {
      const float* pwl = s.pw[lane];
      for (int m = warp; m < JL_MEL_BINS; m += MEL_WARPS) {
        const int lo = __ldg(p.mel_lo + m), cnt = __ldg(p.mel_cnt + m);
        const float* wrow = p.mel_w + m * JL_MEL_MAXW;
        float acc = 0.0f;
        for (int j = 0; j < cnt; ++j) acc = fmaf(__ldg(wrow + j), pwl[lo + j], acc);
        s.out[lane][m] = logf(fmaxf(acc, MEL_FLT_EPS));
      }
    }
  }
  __syncthreads();

  // ---- coalesced tile store (raw log-mel; padded frames = 0) and CMVN partials
  float* feats = p.feats + (static_cast<int64_t>(b) * p.max_frames + f0) * JL_MEL_BINS;
  const int rows = min(MEL_FPC, p.max_frames - f0);
  for (int i = tid; i < rows * JL_MEL_BINS; i += MEL_THREADS) {
    const int fl = i / JL_MEL_BINS;
    const float v = (fl < nv) ? s.out[fl][i - fl * JL_MEL_BINS] : 0.0f;
    feats[i] = v;
    if (!p.apply_cmvn && p.feats_bf16 != nullptr)
      reinterpret_cast<__nv_bfloat16*>(p.feats_bf16)[(static_cast<int64_t>(b) * p.max_frames + f0) * JL_MEL_BINS + i] = __float2bfloat16_rn(v);
  }
  if (partials != nullptr && tid < JL_MEL_BINS) {
    float mean = 0.0f, m2 = 0.0f;
    if (nv > 0) {
      float sum = 0.0f;
      for (int fl = 0; fl < nv; ++fl) sum += s.out[fl][tid];
      mean = sum / static_cast<float>(nv);
      for (int fl = 0; fl < nv; ++fl) {
        const float d = s.out[fl][tid] - mean;
        m2 = fmaf(d, d, m2);
      }
    }
    float* dst = partials + (static_cast<int64_t>(b) * nblk + blk) * 2 * JL_MEL_BINS;
    dst[tid] = mean;
    dst[JL_MEL_BINS + tid] = m2;
  }
}

__global__ void __launch_bounds__(MEL_THREADS) cmvn_kernel(const jl_mel_cmvn_params p, const float* __restrict__ partials, int nblk) {
  jl::pdl_prologue();
  extern __shared__ __align__(16) float cmvn_part[];          // [used tiles][2 · 80] partials of this utterance
  __shared__ __align__(16) float s_mean[JL_MEL_BINS];
  __shared__ __align__(16) float s_std[JL_MEL_BINS];
  const int b = blockIdx.y, blk = blockIdx.x, tid = threadIdx.x;
  const int frames_b = min(utt_frames(p.num_samples[b]), p.max_frames);
  const int f0 = blk * MEL_FPC;
  const int nv = max(0, min(MEL_FPC, frames_b - f0));
  const int used = (frames_b + MEL_FPC - 1) / MEL_FPC;
  // all partials of the utterance with coalesced, independent loads (the merge below is a serial chain: it must not wait on
  // one global load per step)
  {
    const float* src = partials + static_cast<int64_t>(b) * nblk * 2 * JL_MEL_BINS;
    for (int i = tid; i < used * 2 * JL_MEL_BINS; i += MEL_THREADS) cmvn_part[i] = src[i];
  }
  __syncthreads();
  if (tid < JL_MEL_BINS) {
    float n = 0.0f, mean = 0.0f, m2 = 0.0f;
    for (int c = 0; c < used; ++c) {
      const float nc = static_cast<float>(min(MEL_FPC, frames_b - c * MEL_FPC));
      const float mc = cmvn_part[c * 2 * JL_MEL_BINS + tid], m2c = cmvn_part[c * 2 * JL_MEL_BINS + JL_MEL_BINS + tid];
      const float tot = n + nc;
      const float delta = mc - mean;
      mean += delta * (nc / tot);
      m2 += m2c + delta * delta * (n * nc / tot);
      n = tot;
    }
    s_mean[tid] = mean;
    // population std, no epsilon (feature_extraction_speech_to_text.py:152-156); clamped only where HF would divide by 0
    s_std[tid] = (n > 0.0f) ? fmaxf(sqrtf(m2 / n), 1e-10f) : 1.0f;
  }
  __syncthreads();
  const int64_t base = (static_cast<int64_t>(b) * p.max_frames + f0) * JL_MEL_BINS;      // multiple of 2560 elements: 16-byte aligned
  const int rows = min(MEL_FPC, p.max_frames - f0);
  __nv_bfloat16* out16 = reinterpret_cast<__nv_bfloat16*>(p.feats_bf16);
  float4* f4 = reinterpret_cast<float4*>(p.feats + base);
  const bool vec16 = out16 != nullptr && (reinterpret_cast<uintptr_t>(out16) & 7) == 0;
  if ((reinterpret_cast<uintptr_t>(p.feats) & 15) == 0) {
    for (int i = tid; i < rows * (JL_MEL_BINS / 4); i += MEL_THREADS) {
      const int fl = i / (JL_MEL_BINS / 4);
      const int m = (i - fl * (JL_MEL_BINS / 4)) * 4;
      float4 v = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
      if (fl < nv) {
        const float4 x = f4[i];
        const float4 mu = *reinterpret_cast<const float4*>(s_mean + m);
        const float4 sd = *reinterpret_cast<const float4*>(s_std + m);
        v = make_float4((x.x - mu.x) / sd.x, (x.y - mu.y) / sd.y, (x.z - mu.z) / sd.z, (x.w - mu.w) / sd.w);
      }
      f4[i] = v;
      if (vec16) {
        uint2 o;
        o.x = pack_bf16x2(v.x, v.y);
        o.y = pack_bf16x2(v.z, v.w);
        *reinterpret_cast<uint2*>(out16 + base + 4 * i) = o;
      } else if (out16 != nullptr) {
        out16[base + 4 * i + 0] = __float2bfloat16_rn(v.x); out16[base + 4 * i + 1] = __float2bfloat16_rn(v.y);
        out16[base + 4 * i + 2] = __float2bfloat16_rn(v.z); out16[base + 4 * i + 3] = __float2bfloat16_rn(v.w);
      }
    }
  } else {
    for (int i = tid; i < rows * JL_MEL_BINS; i += MEL_THREADS) {
      const int fl = i / JL_MEL_BINS;
      const int m = i - fl * JL_MEL_BINS;
      float v = 0.0f;
      if (fl < nv) v = (p.feats[base + i] - s_mean[m]) / s_std[m];
      p.feats[base + i] = v;
      if (out16 != nullptr) out16[base + i] = __float2bfloat16_rn(v);
    }
  }
}

static int mel_validate(const jl_mel_cmvn_params* p) {
  JL_REQUIRE(p != nullptr, JL_EINVAL, "mel_cmvn: null params");
  JL_REQUIRE(p->wave && p->num_samples && p->window && p->twiddle && p->mel_lo && p->mel_cnt && p->mel_w && p->feats, JL_EINVAL,
             "mel_cmvn: null pointer in params");
  JL_REQUIRE(p->batch > 0 && p->max_frames > 0, JL_EINVAL, "mel_cmvn: batch and max_frames must be positive");
  JL_REQUIRE(p->batch <= 65535, JL_EUNSUPPORTED_SHAPE, "mel_cmvn: batch %d exceeds 65535", p->batch);
  return JL_OK;
}

}  // namespace jl

extern "C" {

int jl_mel_cmvn_workspace_bytes(const jl_mel_cmvn_params* p, size_t* out) {
  JL_REQUIRE(p != nullptr && out != nullptr, JL_EINVAL, "mel_cmvn_workspace_bytes: null argument");
  JL_REQUIRE(p->batch > 0 && p->max_frames > 0, JL_EINVAL, "mel_cmvn: batch and max_frames must be positive");
  const size_t nblk = static_cast<size_t>(jl::ceil_div(p->max_frames, jl::MEL_FPC));
  *out = static_cast<size_t>(p->batch) * nblk * 2 * JL_MEL_BINS * sizeof(float);
  return JL_OK;
}

int jl_mel_cmvn_fwd(const jl_mel_cmvn_params* p, void* workspace, void* stream) {
  int rc = jl::mel_validate(p);
  if (rc != JL_OK) return rc;
  rc = jl::check_device();
  if (rc != JL_OK) return rc;
  JL_REQUIRE(!p->apply_cmvn || workspace != nullptr, JL_EINVAL, "mel_cmvn: CMVN needs the workspace");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const int nblk = jl::ceil_div(p->max_frames, jl::MEL_FPC);
  static thread_local int configured_dev = -1;
  int dev = 0;
  cudaGetDevice(&dev);
  if (configured_dev != dev) {
    cudaError_t e = cudaFuncSetAttribute(jl::mel_fbank_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(jl::MelSmem));
    JL_REQUIRE(e == cudaSuccess, JL_ECUDA, "mel_cmvn: cannot reserve shared memory: %s", cudaGetErrorString(e));
    configured_dev = dev;
  }
  dim3 grid(nblk, p->batch);
  jl::launch(jl::mel_fbank_kernel, grid, jl::MEL_THREADS, sizeof(jl::MelSmem), s, *p, p->apply_cmvn ? reinterpret_cast<float*>(workspace) : nullptr, nblk);
  JL_CHECK_LAUNCH("mel_fbank");
  if (p->apply_cmvn) {
    const size_t part_bytes = static_cast<size_t>(nblk) * 2 * JL_MEL_BINS * sizeof(float);
    JL_REQUIRE(part_bytes <= 200 * 1024, JL_EUNSUPPORTED_SHAPE, "mel_cmvn: max_frames %d exceeds the %d frames one utterance may have", p->max_frames,
               (200 * 1024 / (2 * JL_MEL_BINS * 4)) * jl::MEL_FPC);
    if (part_bytes > 48 * 1024) {
      cudaError_t e = cudaFuncSetAttribute(jl::cmvn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(part_bytes));
      JL_REQUIRE(e == cudaSuccess, JL_ECUDA, "mel_cmvn: cannot reserve shared memory: %s", cudaGetErrorString(e));
    }
    jl::launch(jl::cmvn_kernel, grid, jl::MEL_THREADS, part_bytes, s, *p, reinterpret_cast<const float*>(workspace), nblk);
    JL_CHECK_LAUNCH("cmvn");
  }
  return JL_OK;
}

}  // extern "C"
