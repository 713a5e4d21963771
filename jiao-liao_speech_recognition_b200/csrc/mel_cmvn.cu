// a1 + a2: Kaldi-compatible 80-bin log-mel filterbank + utterance CMVN, waveform resident in HBM.
//
// Replaces SP/torchaudio/compliance/kaldi.py:514-645 (fbank: frames 400/160 snip_edges :63-67, DC removal
// :183-186, pre-emphasis :193-198, povey window :98-100,201-204, zero-pad to 512 :207-211, |rfft|² :616-618,
// mel projection :630, log(max(·, eps)) :633) as called with ×2^15 scaling from
// SP/transformers/models/speech_to_text/feature_extraction_speech_to_text.py:104-120, and utterance_cmvn + padding +
// attention mask at :142-163, :275-303.
//
// Kernel 1 (mel_fbank_kernel): one CTA = 32 consecutive frames of one utterance.  The 5360 samples the frames
//   share are staged once in shared memory with coalesced loads (each sample is read from HBM exactly once);
//   each warp then owns a frame at a time: frame mean by warp shuffle, pre-emphasis + window written as 256
//   complex points, a 256-point radix-4 Stockham FFT in a per-warp shared ping-pong buffer, the real-FFT
//   untangling pass to 257 power bins, and the sparse triangular mel projection (≤ 32 taps per bin) + log.
//   The CTA writes its [32, 80] tile coalesced and leaves per-bin (mean, M2) partials for the CMVN statistics.
// Kernel 2 (cmvn_kernel): merges the utterance's partials in fixed order (Chan's parallel variance update —
//   deterministic, no atomics), normalises the tile, zeroes padded frames, optionally emits a bf16 copy.
#include "common.cuh"

namespace jl {

constexpr int MEL_FRAME_LEN = 400;
constexpr int MEL_FRAME_SHIFT = 160;
constexpr int MEL_NFFT = 512;
constexpr int MEL_FPC = JL_MEL_FRAMES_PER_CTA;                              // frames per CTA
constexpr int MEL_SAMPLES_PER_CTA = (MEL_FPC - 1) * MEL_FRAME_SHIFT + MEL_FRAME_LEN;   // 5360
constexpr int MEL_WARPS = 8;
constexpr int MEL_THREADS = MEL_WARPS * 32;
constexpr float MEL_PREEMPH = 0.97f;
constexpr float MEL_FLT_EPS = 1.1920928955078125e-07f;

// Shared-memory layouts are chosen for the LSU, which bounds this kernel (ncu: 73 % of peak shared wavefronts, half of
// them bank conflicts in the first version): the FFT buffers are padded (one float2 after every four), the per-pass
// twiddles are stored contiguously in the butterfly index, and the mel weights are transposed so that the 32 lanes of a
// tap load hit 32 different banks.
constexpr int MEL_FFT_PAD = 256 + 64;
__device__ __forceinline__ int fft_phys(int idx) { return idx + (idx >> 2); }

struct MelSmem {
  float wave[MEL_SAMPLES_PER_CTA];
  float2 fft[MEL_WARPS][2][MEL_FFT_PAD];
  float window[MEL_FRAME_LEN];
  float2 tw512[257];                              // exp(-2πi k / 512), k = 0..256 (real-FFT untangling)
  float2 tw_pass[3 * (4 + 16 + 64)];              // per radix-4 pass p ∈ {4, 16, 64}: [r-1][k] = exp(-2πi k r / (4p))
  float mel_wt[JL_MEL_MAXW][JL_MEL_BINS + 1];     // transposed: [tap][bin]
  int mel_lo[JL_MEL_BINS];
  int mel_cnt[JL_MEL_BINS];
  float out[MEL_FPC][JL_MEL_BINS];
};

__device__ __forceinline__ float2 cmul(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }

__device__ __forceinline__ int utt_frames(int n) { return n < MEL_FRAME_LEN ? 0 : 1 + (n - MEL_FRAME_LEN) / MEL_FRAME_SHIFT; }

// One radix-4 Stockham pass over 256 complex points held in shared memory; a warp does the 64 butterflies.
// p = size of the sub-transforms already computed (1, 4, 16, 64); tw = this pass's twiddles [3][p].
__device__ __forceinline__ void fft256_pass(const float2* __restrict__ src, float2* __restrict__ dst, const float2* __restrict__ tw, int p,
                                            int lane) {
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int i = lane + 32 * h;
    const int k = i & (p - 1);
    const int j = ((i - k) << 2) + k;
    float2 u0 = src[fft_phys(i)], u1 = src[fft_phys(i + 64)], u2 = src[fft_phys(i + 128)], u3 = src[fft_phys(i + 192)];
    if (p > 1) {
      u1 = cmul(u1, tw[k]);
      u2 = cmul(u2, tw[p + k]);
      u3 = cmul(u3, tw[2 * p + k]);
    }
    const float2 a = make_float2(u0.x + u2.x, u0.y + u2.y);
    const float2 b = make_float2(u0.x - u2.x, u0.y - u2.y);
    const float2 c = make_float2(u1.x + u3.x, u1.y + u3.y);
    const float2 d = make_float2(u1.x - u3.x, u1.y - u3.y);   // (u1 - u3); multiplied by -i → (d.y, -d.x)
    dst[fft_phys(j)] = make_float2(a.x + c.x, a.y + c.y);
    dst[fft_phys(j + p)] = make_float2(b.x + d.y, b.y - d.x);
    dst[fft_phys(j + 2 * p)] = make_float2(a.x - c.x, a.y - c.y);
    dst[fft_phys(j + 3 * p)] = make_float2(b.x - d.y, b.y + d.x);
  }
  __syncwarp();
}

__global__ void __launch_bounds__(MEL_THREADS) mel_fbank_kernel(const jl_mel_cmvn_params p, float* __restrict__ partials, int nblk) {
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
    // ---- stage constants and the shared span of samples
    for (int i = tid; i < MEL_FRAME_LEN; i += MEL_THREADS) s.window[i] = __ldg(p.window + i);
    for (int i = tid; i < 257; i += MEL_THREADS) s.tw512[i] = __ldg(reinterpret_cast<const float2*>(p.twiddle) + i);
    for (int i = tid; i < 3 * (4 + 16 + 64); i += MEL_THREADS) {
      // pass tables: offsets 0 (p = 4), 12 (p = 16), 60 (p = 64); entry [r-1][k] = W256^(k r 64/p) = tw512[2 k r 64/p]
      const int pp = (i < 12) ? 4 : (i < 60 ? 16 : 64);
      const int base = (i < 12) ? 0 : (i < 60 ? 12 : 60);
      const int rr = (i - base) / pp + 1, kk = (i - base) % pp;
      s.tw_pass[i] = __ldg(reinterpret_cast<const float2*>(p.twiddle) + ((2 * kk * rr * (64 / pp)) & 511));
    }
    for (int i = tid; i < JL_MEL_BINS * JL_MEL_MAXW; i += MEL_THREADS) s.mel_wt[i % JL_MEL_MAXW][i / JL_MEL_MAXW] = __ldg(p.mel_w + i);
    if (tid < JL_MEL_BINS) {
      s.mel_lo[tid] = __ldg(p.mel_lo + tid);
      s.mel_cnt[tid] = __ldg(p.mel_cnt + tid);
    }
    const float* wave = p.wave + static_cast<int64_t>(b) * p.wave_stride;
    const int s0 = f0 * MEL_FRAME_SHIFT;
    const int span = (nv - 1) * MEL_FRAME_SHIFT + MEL_FRAME_LEN;     // all < n by construction
    if ((p.wave_stride & 3) == 0 && (reinterpret_cast<uintptr_t>(p.wave) & 15) == 0) {
      const float4* w4 = reinterpret_cast<const float4*>(wave + s0);  // s0 is a multiple of 160 → 16 B aligned
      for (int i = tid; i < span / 4; i += MEL_THREADS) {
        float4 v = __ldg(w4 + i);
        s.wave[4 * i + 0] = v.x * 32768.0f;
        s.wave[4 * i + 1] = v.y * 32768.0f;
        s.wave[4 * i + 2] = v.z * 32768.0f;
        s.wave[4 * i + 3] = v.w * 32768.0f;
      }
    } else {
      for (int i = tid; i < span; i += MEL_THREADS) s.wave[i] = __ldg(wave + s0 + i) * 32768.0f;
    }
    __syncthreads();

    float2* bufA = s.fft[warp][0];
    float2* bufB = s.fft[warp][1];
    float* bufA_f = reinterpret_cast<float*>(bufA);
    float* pw = reinterpret_cast<float*>(bufB);                       // 257 power bins, after the last pass
    for (int fl = warp; fl < nv; fl += MEL_WARPS) {
      const float* x = s.wave + fl * MEL_FRAME_SHIFT;
      // frame mean (kaldi.py:183-186)
      float sum = 0.0f;
      for (int i = lane; i < MEL_FRAME_LEN; i += 32) sum += x[i];
      const float mean = warp_sum(sum) * (1.0f / MEL_FRAME_LEN);
      // pre-emphasis with replicated first sample, povey window, zero-pad to 512
      for (int i = lane; i < MEL_NFFT; i += 32) {
        float y = 0.0f;
        if (i < MEL_FRAME_LEN) {
          const float cur = x[i] - mean;
          const float prev = x[i > 0 ? i - 1 : 0] - mean;
          y = (cur - MEL_PREEMPH * prev) * s.window[i];
        }
        bufA_f[2 * fft_phys(i >> 1) + (i & 1)] = y;                   // z[j] = (y[2j], y[2j+1])
      }
      __syncwarp();
      fft256_pass(bufA, bufB, s.tw_pass, 1, lane);
      fft256_pass(bufB, bufA, s.tw_pass, 4, lane);
      fft256_pass(bufA, bufB, s.tw_pass + 12, 16, lane);
      fft256_pass(bufB, bufA, s.tw_pass + 60, 64, lane);
      // untangle the real transform: X[k] = E[k] + W512^k O[k], k = 0..256; power = |X|²
      float pk[9];
#pragma unroll
      for (int j = 0; j < 9; ++j) {
        const int k = lane + 32 * j;
        pk[j] = 0.0f;
        if (k <= 256) {
          const float2 zk = bufA[fft_phys(k & 255)];
          const float2 zn = bufA[fft_phys((256 - k) & 255)];
          const float2 e = make_float2(0.5f * (zk.x + zn.x), 0.5f * (zk.y - zn.y));
          const float2 o = make_float2(0.5f * (zk.y + zn.y), -0.5f * (zk.x - zn.x));   // (zk - conj(zn)) / (2i)
          const float2 wo = cmul(s.tw512[k], o);
          const float re = e.x + wo.x, im = e.y + wo.y;
          pk[j] = re * re + im * im;
        }
      }
      __syncwarp();
#pragma unroll
      for (int j = 0; j < 9; ++j) {
        const int k = lane + 32 * j;
        if (k <= 256) pw[k] = pk[j];
      }
      __syncwarp();
      // sparse mel projection + log (kaldi.py:630-633)
      for (int m = lane; m < JL_MEL_BINS; m += 32) {
        const int lo = s.mel_lo[m], cnt = s.mel_cnt[m];
        float acc = 0.0f;
        for (int j = 0; j < cnt; ++j) acc = fmaf(s.mel_wt[j][m], pw[lo + j], acc);
        s.out[fl][m] = logf(fmaxf(acc, MEL_FLT_EPS));
      }
      __syncwarp();
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
  __shared__ float s_mean[JL_MEL_BINS];
  __shared__ float s_std[JL_MEL_BINS];
  const int b = blockIdx.y, blk = blockIdx.x, tid = threadIdx.x;
  const int frames_b = min(utt_frames(p.num_samples[b]), p.max_frames);
  const int f0 = blk * MEL_FPC;
  const int nv = max(0, min(MEL_FPC, frames_b - f0));
  if (tid < JL_MEL_BINS) {
    float n = 0.0f, mean = 0.0f, m2 = 0.0f;
    const int used = (frames_b + MEL_FPC - 1) / MEL_FPC;
    for (int c = 0; c < used; ++c) {
      const float nc = static_cast<float>(min(MEL_FPC, frames_b - c * MEL_FPC));
      const float* src = partials + (static_cast<int64_t>(b) * nblk + c) * 2 * JL_MEL_BINS;
      const float mc = src[tid], m2c = src[JL_MEL_BINS + tid];
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
  const int64_t base = (static_cast<int64_t>(b) * p.max_frames + f0) * JL_MEL_BINS;
  const int rows = min(MEL_FPC, p.max_frames - f0);
  __nv_bfloat16* out16 = reinterpret_cast<__nv_bfloat16*>(p.feats_bf16);
  for (int i = tid; i < rows * JL_MEL_BINS; i += MEL_THREADS) {
    const int fl = i / JL_MEL_BINS;
    const int m = i - fl * JL_MEL_BINS;
    float v = 0.0f;
    if (fl < nv) v = (p.feats[base + i] - s_mean[m]) / s_std[m];
    p.feats[base + i] = v;
    if (out16 != nullptr) out16[base + i] = __float2bfloat16_rn(v);
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
    jl::launch(jl::cmvn_kernel, grid, jl::MEL_THREADS, 0, s, *p, reinterpret_cast<const float*>(workspace), nblk);
    JL_CHECK_LAUNCH("cmvn");
  }
  return JL_OK;
}

}  // extern "C"
