/*
 * jl_b200.h — C ABI of libjl_b200.so: the B200 (sm_100a) kernels behind the
 * Jiao-Liao ASR forward + adapter fine-tune hot path.
 *
 * The reference (mixxs/Jiao-Liao_Speech_Recognition) publishes no code and
 * therefore no FFI of its own (/root/reference/README.md:3); its hot path runs
 * inside the Python dependencies it pins (/root/reference/requirements.txt:75,78,81).
 * Each entry point below names the dependency function it replaces
 * (SP = site-packages of the build container).
 *
 * Conventions:
 *  - plain C, POD parameter structs, raw DEVICE pointers, sizes and element strides;
 *  - the caller owns every buffer including workspace; the library never allocates or
 *    frees device memory and keeps no pointer after the call returns;
 *  - every call enqueues its work on `stream` (a cudaStream_t passed as void*) and
 *    returns without synchronising;
 *  - return 0 (JL_OK) or a negative error code; text via jl_last_error() (thread-local);
 *  - no CPU fallback: a device that is not sm_100 gives JL_EUNSUPPORTED.
 */
#ifndef JL_B200_H
#define JL_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define JL_VERSION 1

enum {
  JL_OK = 0,
  JL_EINVAL = -1,            /* bad argument (null pointer, negative size, misalignment) */
  JL_EUNSUPPORTED_SHAPE = -2,/* shape outside what the kernel was built for */
  JL_ECUDA = -3,             /* CUDA runtime / driver error at launch */
  JL_EUNSUPPORTED = -4       /* not an sm_100 device */
};

enum { JL_DT_BF16 = 0, JL_DT_F32 = 1 };

int jl_version(void);
const char* jl_last_error(void);
/* number of kernels this library has launched in the calling process (bench.py's gpu_launches) */
int64_t jl_launch_count(void);
void jl_launch_count_reset(void);

/* ------------------------------------------------------------------------------------------
 * a1 + a2: 80-bin Kaldi log-mel filterbank + utterance CMVN.
 * Replaces SP/torchaudio/compliance/kaldi.py:514-645 (fbank) as called from
 * SP/transformers/models/speech_to_text/feature_extraction_speech_to_text.py:104-120, and
 * utterance_cmvn / pad / attention-mask at :142-163, :275-303.
 * ------------------------------------------------------------------------------------------ */
#define JL_MEL_BINS 80
#define JL_MEL_MAXW 32       /* widest triangular filter in FFT bins */
#define JL_MEL_FRAMES_PER_CTA 32
typedef struct {
  const float* wave;          /* [batch, wave_stride] fp32 in [-1, 1], 16 kHz */
  int64_t wave_stride;        /* elements between utterances */
  const int32_t* num_samples; /* [batch] valid samples per utterance */
  int32_t batch;
  int32_t max_frames;         /* F of the output (>= frames of the longest utterance) */
  const float* window;        /* [400] povey window (host-built, fp32) */
  const float* twiddle;       /* [512, 2] (cos, -sin)(2*pi*k/512), k = 0..511 (host-built from float64) */
  const int32_t* mel_lo;      /* [80] first FFT bin of each filter */
  const int32_t* mel_cnt;     /* [80] number of FFT bins (<= JL_MEL_MAXW) */
  const float* mel_w;         /* [80, JL_MEL_MAXW] weights, zero padded */
  float* feats;               /* [batch, max_frames, 80] fp32 out (CMVN'd, padded frames = 0) */
  void* feats_bf16;           /* optional bf16 copy of feats, same shape (may be NULL) */
  int32_t* attention_mask;    /* [batch, max_frames] int32 out (may be NULL) */
  int32_t* frame_lengths;     /* [batch] int32 out: valid frames per utterance */
  int32_t apply_cmvn;         /* 1 = CMVN (default path); 0 = raw log-mel (tests) */
} jl_mel_cmvn_params;
int jl_mel_cmvn_workspace_bytes(const jl_mel_cmvn_params* p, size_t* out);
int jl_mel_cmvn_fwd(const jl_mel_cmvn_params* p, void* workspace, void* stream);

/* ------------------------------------------------------------------------------------------
 * bf16 GEMM on tcgen05 tensor cores:  C[M,N] = epilogue(alpha * A[M,K] · B[N,K]^T + bias)
 * A and B are K-contiguous by default (nn.Linear layout); see a_layout / b_layout.  Replaces torch.nn.functional.linear as used by
 * SP/transformers/models/wav2vec2/modeling_wav2vec2.py:495-498,524-528,547 (q/k/v/out),
 * :557-573 (FFN), :1708 (lm_head), and (through im2col) the Conv1d at
 * SP/transformers/models/speech_to_text/modeling_speech_to_text.py:82-99.
 * ------------------------------------------------------------------------------------------ */
enum {
  JL_EPI_NONE = 0,
  JL_EPI_GELU = 1,      /* erf GELU; if aux_out != NULL the pre-activation is stored there (bf16) */
  JL_EPI_RELU = 2,
  JL_EPI_GELU_BWD = 3,  /* C = acc * gelu'(aux)          (aux = saved pre-activation, bf16) */
  JL_EPI_RELU_BWD = 4,  /* C = acc * (aux > 0)           (aux = saved post-activation, bf16) */
  JL_EPI_GLU = 5,       /* C[:, j] = v[2j] * sigmoid(v[2j+1]); C has N/2 columns (interleaved weight rows) */
  JL_EPI_GELU_DGELU = 6,/* erf GELU to C and its derivative gelu'(pre-activation) to aux_out (bf16, required): the training
                         * forward of the FFN input projection — erf is evaluated once per element, not again in backward */
  JL_EPI_MUL_AUX = 7,   /* C = acc * aux                   (aux = saved activation derivative, bf16) */
  JL_EPI_ARGMAX = 8     /* SURVEY §8 f1 (inference): the output matrix is never written.  Per row and per chunk of 32 output columns
                         * the epilogue emits (max, first argmax) of alpha*acc + bias:  c = float [ceil(N/32), ldc] maxima,
                         * aux_out = int32 [ceil(N/32), ldaux_out] column indices (chunk-major, ld >= M; out_dtype ignored).
                         * lm_head + greedy CTC decode without a [B*T', V] logits tensor: jl_ctc_greedy_from_partials finishes it. */
};
enum {
  JL_LAYOUT_K = 0,      /* operand stored with K contiguous:  A[M, K] / B[N, K]  (nn.Linear layout) */
  JL_LAYOUT_MN = 1      /* operand stored with M (or N) contiguous:  A as [K, M] / B as [K, N]  */
};
typedef struct {
  const void* a; int64_t lda;        /* bf16, row stride lda elements (multiple of 8); [M, K] or [K, M] per a_layout.  A K-major A may
                                        have lda < K: rows then overlap in memory (sliding windows of k frames over a [T, C]
                                        activation, lda = stride·C, K = k·C — a Conv1d without an im2col copy) */
  const void* b; int64_t ldb;        /* bf16, [N, K] or [K, N] per b_layout */
  int32_t a_layout, b_layout;        /* JL_LAYOUT_K | JL_LAYOUT_MN — MN lets dgrad (dY·W) and wgrad (dYᵀ·X) run without transposed copies */
  void* c; int64_t ldc;              /* [M, N] (N/2 columns for GLU), dtype out_dtype */
  const float* bias;                 /* [N] fp32 or NULL */
  const void* residual; int64_t ldr; /* bf16 [M, N] added after the epilogue op, or NULL */
  const void* aux; int64_t ldaux;    /* bf16 [M, N] epilogue input (GELU_BWD / RELU_BWD) or NULL */
  void* aux_out; int64_t ldaux_out;  /* bf16 [M, N] pre-activation out (GELU) or NULL */
  const int32_t* row_lengths;        /* optional [M / rows_per_seq]: rows t >= length are written as 0 */
  int32_t rows_per_seq;
  int32_t m, n, k;
  int32_t epilogue;
  int32_t out_dtype;                 /* JL_DT_BF16 | JL_DT_F32 */
  float alpha;
  void* workspace;                   /* optional K-split scratch (jl_gemm_workspace_bytes), 256-byte aligned; NULL = never split */
  int64_t workspace_bytes;
} jl_gemm_params;
int jl_gemm_bf16(const jl_gemm_params* p, void* stream);
/* Bytes of scratch this product would use, 0 when it runs unsplit.  Two uses:
 *  - split-K: plain products (no bias / activation / residual) whose output tiles cannot fill the SMs and whose K is long —
 *    the weight-gradient shapes dYᵀ·X; partials are reduced in fixed order by a second launch;
 *  - tail split: any product on the CTA-pair kernel whose last wave of 256-row tiles is partial — those tiles are cut into
 *    K ranges that run on the otherwise idle pairs; the last range to arrive adds the partials in range order and applies
 *    the epilogue inside the same launch (no CTA waits for another).  Its arrival counters live in the first
 *    jl_gemm_workspace_zero_bytes() bytes of the workspace, which must be ZERO on entry and are left zero; a workspace must
 *    not be shared by products that may run concurrently (one per stream). */
int jl_gemm_workspace_bytes(const jl_gemm_params* p, size_t* out);
int jl_gemm_workspace_zero_bytes(const jl_gemm_params* p, size_t* out);

/* ------------------------------------------------------------------------------------------
 * LayerNorm over the last dim (eps 1e-5).  Replaces nn.LayerNorm at
 * SP/transformers/models/wav2vec2/modeling_wav2vec2.py:623,625,639,645,792,941.
 * ------------------------------------------------------------------------------------------ */
typedef struct {
  const void* x; int64_t ldx;     /* bf16 [rows, d] */
  const float* gamma; const float* beta;
  void* y; int64_t ldy;           /* bf16 [rows, d] */
  float* mean; float* rstd;       /* [rows] fp32 out (NULL in inference) */
  int32_t rows, d;
  float eps;
  int32_t act;                    /* JL_EPI_NONE, or JL_EPI_GELU applied to the normalised row: the Conv1d → LayerNorm → GELU layers
                                     of the wav2vec2 feature encoder (modeling_wav2vec2.py:291-299) */
} jl_layernorm_fwd_params;
int jl_layernorm_fwd(const jl_layernorm_fwd_params* p, void* stream);

typedef struct {
  const void* dy; int64_t lddy;   /* bf16 [rows, d] */
  const void* x; int64_t ldx;     /* bf16 [rows, d] forward input */
  const float* gamma;
  const float* mean; const float* rstd;
  const void* dres; int64_t lddres; /* optional bf16 [rows, d] added to dx (the residual branch grad) */
  void* dx; int64_t lddx;         /* bf16 [rows, d] */
  float* dgamma; float* dbeta;    /* optional fp32 [d] (trainable adapter norms); NULL for frozen */
  float* partial;                 /* workspace fp32 [2, nblk, d] when dgamma != NULL */
  int32_t rows, d;
} jl_layernorm_bwd_params;
int jl_layernorm_bwd_workspace_bytes(const jl_layernorm_bwd_params* p, size_t* out);
int jl_layernorm_bwd(const jl_layernorm_bwd_params* p, void* stream);
/* only dγ / dβ (uses dy, x, mean, rstd, dgamma, dbeta, rows, d of the struct): one launch, deterministic; lets the caller run
 * the dx part (jl_layernorm_bwd with dgamma = NULL) on the critical path and the weight gradients on a side stream */
int jl_layernorm_wgrad(const jl_layernorm_bwd_params* p, void* stream);

/* Several column reductions in one launch — what the weight-gradient branch of an adapter's backward pass needs: the bias
 * gradients sum_r dy[r, :] of its projections (autograd's sum over rows of nn.Linear's grad_output) and the d gamma / d beta of its
 * LayerNorm.  Per job: out_sum[c] = sum_r dy[r, c] and, when x != NULL, out_dot[c] = sum_r dy[r, c] * (x[r, c] - mean[r]) * rstd[r].
 * Deterministic (fixed-order trees, no atomics).  cols, lddy, ldx multiples of 8. */
#define JL_COLREDUCE_MAX_JOBS 4
typedef struct {
  const void* dy; int64_t lddy;       /* bf16 [rows, cols] */
  const void* x; int64_t ldx;         /* bf16 [rows, cols] or NULL (plain column sum) */
  const float* mean; const float* rstd;   /* [rows] fp32 (with x) */
  int32_t rows, cols;
  float* out_sum;                     /* [cols] fp32 or NULL */
  float* out_dot;                     /* [cols] fp32 or NULL */
} jl_colreduce_job;
int jl_colreduce_multi(const jl_colreduce_job* jobs, int32_t num_jobs, void* stream);

/* ------------------------------------------------------------------------------------------
 * a6: WFAdapter forward in one kernel — LayerNorm + factorised down / up projections + bias / ReLU + residual.
 *   out = h + (relu((LN(h) B_d^T) A_d^T + c_d) B_u^T) A_u^T + c_u            (/root/reference/README.md:1;
 * bottleneck-adapter analogue SP/transformers/models/wav2vec2/modeling_wav2vec2.py:931-953, hook :647-648).
 * The LayerNorm is folded into the first projection, so the caller passes  bd_scaled = bf16(B_d * gamma)  [r, d],
 * s[j] = sum_k bd_scaled[j, k]  and  t[j] = sum_k B_d[j, k] * beta[k];  the rank dimension of A_d and A_u is zero-padded to 64.
 * ------------------------------------------------------------------------------------------ */
typedef struct {
  const void* h; int64_t ldh;       /* bf16 [rows, d] */
  void* out; int64_t ldo;           /* bf16 [rows, d] */
  const void* bd_scaled;            /* bf16 [r, d] */
  const float* s; const float* t;   /* fp32 [r] */
  const void* ad_pad;               /* bf16 [b, 64]   (A_d [b, r], zero padded) */
  const float* c_d;                 /* fp32 [b] */
  const void* bu;                   /* bf16 [r, b] */
  const void* au_pad;               /* bf16 [d, 64]   (A_u [d, r], zero padded) */
  const float* c_u;                 /* fp32 [d] */
  const int32_t* row_lengths;       /* optional: rows t >= length of their utterance are written as 0 */
  int32_t rows_per_seq;
  float* mean; float* rstd;         /* optional fp32 [rows] LayerNorm statistics out */
  int32_t rows, d, r, b;
  float eps;
  /* training: the intermediates the backward pass needs, written by the same launch (all optional, bf16, contiguous rows):
   * t1 = LN(h) B_d^T [rows, r], u = relu(t1 A_d^T + c_d) [rows, b], t2 = u B_u^T [rows, r] */
  void* t1_out; void* u_out; void* t2_out;
} jl_wfadapter_fwd_params;
int jl_wfadapter_fwd(const jl_wfadapter_fwd_params* p, void* stream);
/* The operands jl_wfadapter_fwd reads, derived from the adapter's parameters ON THE DEVICE (so that a captured training step
 * re-derives them from the freshly updated factors at every replay): for each of the `sets` dialect factor sets
 *   bd_scaled = bf16(B_d * gamma) [r, d],  s[j] = sum_k bd_scaled[j, k],  t[j] = sum_k B_d[j, k] * beta[k],
 *   ad_pad = A_d zero-padded to [b, 64],  au_pad = A_u zero-padded to [d, 64].
 * down_B [sets, r, d], down_A [sets, b, r], up_A [sets, d, r]: bf16 (the shadow copies the GEMMs read); gamma, beta: fp32 [d]. */
typedef struct {
  const void* down_B; const void* down_A; const void* up_A;
  const float* gamma; const float* beta;
  void* bd_scaled; float* s; float* t; void* ad_pad; void* au_pad;     /* outputs, [sets, ...] */
  int32_t sets, d, r, b;
} jl_wfadapter_pack_params;
int jl_wfadapter_pack(const jl_wfadapter_pack_params* p, void* stream);

/* ------------------------------------------------------------------------------------------
 * a7: the AttAdapter forward in one kernel, for utterances of at most 256 frames —
 *   out = h + softmax(q k^T * scale + keymask) v W_o^T + b_o,   q|k|v = LN(h) W_qkv^T + b_qkv  (one head of 64)
 * (/root/reference/README.md:1 "adapter with attention").  Replaces LayerNorm → GEMM → jl_attn_fwd → GEMM.
 * The LayerNorm is folded into the q|k|v projection, so the caller passes the operands jl_lnfold_pack derives on the device:
 * wqkv_scaled = bf16(W_qkv * gamma) [192, d], s[j] = sum_k wqkv_scaled[j, k], tb[j] = sum_k W_qkv[j, k] beta[k] + b_qkv[j].
 * Rows: padded layout b*seq + t with `lengths`, or packed layout cu_seqlens[b] + t (see jl_attn_fwd_params).
 * Training: qkv_out [rows, 192], a_out [rows, 64] (bf16), mean / rstd [rows], lse ([B, seq] padded, [total_rows] packed) receive
 * what the backward pass (jl_attn_bwd + GEMMs) needs.
 * ------------------------------------------------------------------------------------------ */
typedef struct {
  const void* h; int64_t ldh;          /* bf16 [rows, d] */
  void* out; int64_t ldo;              /* bf16 [rows, d] */
  const void* wqkv_scaled;             /* bf16 [192, d] */
  const float* s; const float* tb;     /* fp32 [192] */
  const void* wo; const float* bo;     /* bf16 [d, 64], fp32 [d] */
  const int32_t* lengths;              /* [B] valid frames (padded layout) or NULL */
  const int32_t* cu_seqlens; int32_t total_rows;   /* packed layout or NULL */
  int32_t batch, seq, d;               /* seq <= 256; d a multiple of 128, at most 1024 */
  float scale, eps;
  int32_t zero_padded_rows;            /* padded layout: rows t >= length are written as 0 (else b_o + h, as the composed path) */
  void* qkv_out; void* a_out; float* mean; float* rstd; float* lse;      /* optional (training) */
  int32_t col_split;                   /* 2: two clusters per utterance share the output columns (each repeats the attention): lower
                                          latency, more SM-time; 0 / 1: one */
} jl_attadapter_fwd_params;
int jl_attadapter_fwd(const jl_attadapter_fwd_params* p, void* stream);
/* LayerNorm folding of a projection that follows a LayerNorm: LN(h) W^T + bias = rstd (h W'^T - mean s) + tb with
 * W' = bf16(W * gamma) [n, d], s[j] = sum_k W'[j, k], tb[j] = sum_k W[j, k] beta[k] + bias[j] (bias may be NULL).  Derived on the
 * device so that a captured training step follows the optimizer. */
typedef struct {
  const void* w; const float* bias; const float* gamma; const float* beta;
  void* w_scaled; float* s; float* tb;
  int32_t n, d;
} jl_lnfold_pack_params;
int jl_lnfold_pack(const jl_lnfold_pack_params* p, void* stream);
/* the same for up to JL_LNFOLD_MAX_JOBS projections in one launch (every AttAdapter of a model at the start of a training step) */
#define JL_LNFOLD_MAX_JOBS 48
int jl_lnfold_pack_multi(const jl_lnfold_pack_params* jobs, int32_t count, void* stream);

/* ------------------------------------------------------------------------------------------
 * a7 (backward): backward through "LayerNorm → narrow projection" in one kernel — the tail of the AttAdapter backward:
 *   forward  y = LN(h) W^T + b   (W [n, d], n a multiple of 8, at most 192: the AttAdapter's q|k|v projection, n = 192; the WFAdapter's
 *   first low-rank projection, n = rank)
 *   dz = dy W;   dx = LayerNorm'(dz; h, mean, rstd, gamma) + dres
 * Replaces jl_gemm_bf16 (dy · W) + jl_layernorm_bwd (SP/transformers/models/wav2vec2/modeling_wav2vec2.py:941 analogue under
 * autograd).  s / tb are the LayerNorm-fold vectors of jl_lnfold_pack for this projection; y is the projection output saved by
 * the forward pass.  dz (optional) receives dy W in bf16 for the LayerNorm weight gradients (jl_layernorm_wgrad).
 * ------------------------------------------------------------------------------------------ */
typedef struct {
  const void* dy; int64_t lddy;        /* bf16 [rows, n] */
  const void* y; int64_t ldy;          /* bf16 [rows, n] saved forward output (bias included) */
  const void* w;                       /* bf16 [n, d] contiguous (unscaled projection weight) */
  const float* s; const float* tb;     /* fp32 [n] */
  const float* gamma;                  /* fp32 [d] */
  const void* h; int64_t ldh;          /* bf16 [rows, d] LayerNorm input */
  const float* mean; const float* rstd;/* fp32 [rows] */
  const void* dres; int64_t lddres;    /* bf16 [rows, d] gradient of the residual branch (added) */
  void* dx; int64_t lddx;              /* bf16 [rows, d] out */
  void* dz; int64_t lddz;              /* optional bf16 [rows, d] out */
  int32_t rows, n, d;                  /* d a multiple of 64, at most 1024 */
  float* col_partial;                  /* optional fp32 [3][ceil(rows / 128)][d]: per-row-tile column sums of dz (→ dbeta), dz * xhat
                                          (→ dgamma) and dres (→ the bias gradient of the layer whose output gradient dres is), to be
                                          summed by jl_lnproj_bwd_reduce — the LayerNorm weight gradients without a dz tensor */
  void* dy_scaled; int64_t lddys;      /* optional bf16 [rows, n] out: dy * rstd (row-scaled), the A operand of the projection's weight gradient */
  float* wgrad_partial;                /* with dy_scaled: fp32 [2][4 * ceil(rows / 128)][n] column sums of dy and of dy_scaled * mean per 32 rows */
  int32_t col_split;                   /* CTAs per 128-row tile (column ranges): 0 = automatic, 1, 2 — 2 lowers the kernel's latency, 1 its SM-time */
  /* Row runs (the dialect runs of a WFAdapter): rows [run_start[i], run_start[i+1]) use factor set run_set[i], i.e. rows set*n … of w
   * ([w_sets * n, d]) and entries set*n … of s / tb.  Row tiles never straddle a run; col_partial then has one row per tile of every
   * run, in run order.  num_runs = 0: one run, all rows, set 0. */
  int32_t num_runs, w_sets;
  int32_t run_start[9];
  int32_t run_set[8];
} jl_lnproj_bwd_params;
#define JL_LNPROJ_MAX_RUNS 8
int jl_lnproj_bwd(const jl_lnproj_bwd_params* p, void* stream);
/* The projection's weight / bias gradient WITHOUT LN(h):  m0 = dy_scaled^T h  (a jl_gemm_bf16 with MN-major operands, fp32 [n, d]) is
 * finished in place:  dW = (m0 - v 1^T) * gamma + cs beta^T,  dbias = cs,  with cs / v = the column sums in wgrad_partial
 * (partial_rows = 4 * ceil(rows / 128)).  Replaces the LayerNorm recomputation + jl_colsum_bf16 of the two-kernel path. */
/* dy_scaled / wgrad_partial (see jl_lnproj_bwd_params) produced by a kernel of their own, for the weight-gradient branch */
int jl_lnproj_wgrad_prep(const void* dy, int64_t lddy, const float* mean, const float* rstd, int32_t rows, int32_t n, void* dy_scaled, int64_t lddys,
                         float* wgrad_partial, void* stream);
int jl_lnproj_wgrad(float* m0, int64_t ldm, const float* wgrad_partial, int32_t partial_rows, int32_t n, int32_t d, const float* gamma, const float* beta,
                    float* dbias, void* stream);
/* dgamma / dbeta / dbias [d] (any may be NULL) = fixed-order sums over the row tiles of col_partial; accumulate != 0: dgamma and dbeta
 * are added to (row ranges of one LayerNorm handled by several calls — the dialect runs of a WFAdapter), dbias is always overwritten */
int jl_lnproj_bwd_reduce(const float* col_partial, int32_t row_tiles, int32_t d, float* dgamma, float* dbeta, float* dbias, int32_t accumulate,
                         int32_t tile_offset, int32_t total_tiles, void* stream);    /* sums tiles [tile_offset, tile_offset + row_tiles) of total_tiles (0 = row_tiles) */

/* ------------------------------------------------------------------------------------------
 * f4: AdapterFusion-style AttAdapter over the K source-dialect adapters of a slot (SURVEY.md §8c ambiguity (ii), §8f f4;
 * /root/reference/README.md:1 "multi-dialect knowledge transfer" + "adapter with attention").  Per frame:
 *   alpha = softmax_k(q . key_k * scale),   out = h + sum_k alpha_k y_k
 * y_k = update of the k-th dialect's WFAdapter, q / key_k = projections of LN(h) / y_k (all produced by jl_gemm_bf16).  These
 * entry points are the part that is not a GEMM: the K dot products, the softmax over K and the weighted sum (forward); and
 * d alpha, the softmax backward, dq, dkey_k and dy_k = alpha_k dout (backward; the caller adds dkey_k W_k to dy_k with a GEMM).
 * One struct serves both directions; a tensor of K matrices is addressed as base + k * stride (elements).
 * ------------------------------------------------------------------------------------------ */
typedef struct {
  const void* h; int64_t ldh;                          /* fwd: bf16 [rows, d] */
  const void* y; int64_t ldy; int64_t y_stride;        /* bf16 K x [rows, d] */
  const void* q; int64_t ldq;                          /* bf16 [rows, b] */
  const void* key; int64_t ldkey; int64_t key_stride;  /* bf16 K x [rows, b] */
  void* out; int64_t ldo;                              /* fwd: bf16 [rows, d] */
  float* alpha;                                        /* fp32 [rows, K]: written by fwd, read by bwd */
  const void* dout; int64_t lddout;                    /* bwd: bf16 [rows, d] */
  void* dy; int64_t lddy; int64_t dy_stride;           /* bwd out: bf16 K x [rows, d] = alpha_k * dout */
  void* dq; int64_t lddq;                              /* bwd out: bf16 [rows, b] */
  void* dkey; int64_t lddkey; int64_t dkey_stride;     /* bwd out: bf16 K x [rows, b] */
  int32_t rows, d, b, num_adapters;                    /* d, b multiples of 8; b <= 256; num_adapters <= 8 */
  float scale;                                         /* 1 / sqrt(b) */
  const int32_t* row_lengths; int32_t rows_per_seq;    /* fwd, optional (padded layout): rows t >= length of their utterance := 0 */
} jl_fusion_params;
int jl_fusion_combine_fwd(const jl_fusion_params* p, void* stream);
int jl_fusion_combine_bwd(const jl_fusion_params* p, void* stream);

/* ------------------------------------------------------------------------------------------
 * Self-attention softmax(Q K^T * scale + keymask) V per (utterance, head), head_dim 64.
 * Replaces the eager math at SP/transformers/models/wav2vec2/modeling_wav2vec2.py:438-463
 * (never materialises the [B,H,T,T] scores); also the AttAdapter attention (heads = 1).
 * q/k/v/o rows are (b*seq + t); head h occupies columns [h*64, h*64+64).
 * ------------------------------------------------------------------------------------------ */
typedef struct {
  const void* q; const void* k; const void* v; int64_t ld_qkv;  /* bf16, common row stride */
  void* o; int64_t ld_o;                                         /* bf16 [B*seq, heads*64] */
  float* lse;                     /* [B, heads, seq] fp32 (log-sum-exp of scaled scores) or NULL */
  const int32_t* lengths;         /* [B] valid frames (keys >= length masked; query rows >= length output 0) */
  int32_t batch, seq, heads;
  float scale;
  /* Packed ("varlen") layout — SURVEY.md §5: utterance b occupies rows [cu_seqlens[b], cu_seqlens[b+1]) of [total_rows, ld]
   * matrices, there are no padding rows; `seq` is then only the upper bound on an utterance's length (grid sizing), `lengths` is
   * ignored, and lse is laid out [heads, total_rows].  NULL = padded [B*seq] layout. */
  const int32_t* cu_seqlens;      /* [B + 1] int32 device, or NULL */
  int32_t total_rows;
} jl_attn_fwd_params;
int jl_attn_fwd(const jl_attn_fwd_params* p, void* stream);

typedef struct {
  const void* q; const void* k; const void* v; int64_t ld_qkv;
  const void* o; const void* d_o; int64_t ld_o;
  const float* lse;
  void* dq; void* dk; void* dv; int64_t ld_dqkv;   /* bf16 */
  float* delta;                   /* workspace [B, heads, seq] fp32 ([heads, total_rows] in the packed layout) */
  const int32_t* lengths;
  int32_t batch, seq, heads;
  float scale;
  const int32_t* cu_seqlens;      /* packed layout, see jl_attn_fwd_params; NULL = padded */
  int32_t total_rows;
} jl_attn_bwd_params;
int jl_attn_bwd(const jl_attn_bwd_params* p, void* stream);

/* ------------------------------------------------------------------------------------------
 * a9: fused log-softmax + CTC loss (+ gradient w.r.t. the logits).  Replaces
 * SP/transformers/models/wav2vec2/modeling_wav2vec2.py:1711-1736 →
 * SP/torch/nn/functional.py:3042-3115 (ctc_loss) incl. the log_softmax pass.
 * ------------------------------------------------------------------------------------------ */
enum { JL_CTC_SUM = 0, JL_CTC_MEAN = 1 };
typedef struct {
  const void* logits; int64_t ld_logits;  /* [B*seq, V] rows, dtype logits_dtype */
  int32_t logits_dtype;
  const int32_t* labels; int32_t max_label_len;   /* [B, max_label_len], negative = padding */
  const int32_t* input_lengths;           /* [B] valid frames */
  int32_t batch, seq, vocab, blank;
  int32_t reduction, zero_infinity;
  float* nll;                             /* [B] fp32 out: per-utterance −log p (0 if zero_infinity hit) */
  float* loss;                            /* [1] fp32 out: reduced loss */
  void* grad; int64_t ld_grad;            /* optional [B*seq, V] d loss / d logits, dtype grad_dtype */
  int32_t grad_dtype;
  const int32_t* cu_seqlens;              /* optional [B + 1]: packed layout — frame t of utterance b is row cu_seqlens[b] + t of
                                             logits and grad (no padding rows; `seq` = upper bound on the lengths); NULL = row b*seq + t */
} jl_ctc_params;
int jl_ctc_workspace_bytes(const jl_ctc_params* p, size_t* out);
int jl_ctc_fwd(const jl_ctc_params* p, void* workspace, void* stream);

/* ------------------------------------------------------------------------------------------
 * a10: greedy CTC decode — argmax over V (first max wins), drop frames >= length, collapse
 * consecutive repeats, drop blank.  Replaces torch.argmax + the Python groupby at
 * SP/transformers/models/wav2vec2/tokenization_wav2vec2.py:310-317.
 * ------------------------------------------------------------------------------------------ */
typedef struct {
  const void* logits; int64_t ld_logits; int32_t logits_dtype;
  const int32_t* input_lengths;
  int32_t batch, seq, vocab, blank;
  int32_t* frame_ids;     /* workspace/out [B, seq] per-frame argmax */
  int32_t* out_ids;       /* [B, seq] compacted ids (tail = -1) */
  int32_t* out_lengths;   /* [B] */
  const int32_t* cu_seqlens; /* optional [B + 1]: packed logits rows (see jl_ctc_params); frame_ids / out_ids stay [B, seq] */
} jl_ctc_greedy_params;
int jl_ctc_greedy(const jl_ctc_greedy_params* p, void* stream);
/* f1: greedy decode from the per-chunk (max, argmax) pairs a JL_EPI_ARGMAX lm_head GEMM produced instead of logits.
 * pmax / pidx: [num_chunks, ld] (row = frame row of the hidden-state matrix: b*seq + t, or cu_seqlens[b] + t when packed).
 * Ties go to the lowest column index, as torch.argmax.  Outputs as jl_ctc_greedy ([B, seq] frame_ids / out_ids, [B] out_lengths). */
int jl_ctc_greedy_from_partials(const float* pmax, const int32_t* pidx, int64_t ld, int32_t num_chunks, const int32_t* input_lengths,
                                const int32_t* cu_seqlens, int32_t batch, int32_t seq, int32_t blank, int32_t* frame_ids,
                                int32_t* out_ids, int32_t* out_lengths, void* stream);

/* ------------------------------------------------------------------------------------------
 * Data-movement / elementwise helpers on the path.
 * ------------------------------------------------------------------------------------------ */
/* im2col for Conv1d(k=5, s=2, p=2) over [B, t_in, c] bf16 → [B*t_out, 5*c] bf16 (tap-major). */
int jl_im2col_k5s2(const void* x, void* out, int32_t batch, int32_t t_in, int32_t c, int32_t t_out, void* stream);
/* h[b,t,:] = h[b,t,:] * scale + pos[t+2,:] for t < len_b;  h[b,t,:] = 0 for t >= len_b
 * (modeling_speech_to_text.py:542,568-579; pos_table is the host-built [seq + 2, d] fp32 sinusoid table, :123-139) */
int jl_embed_positions(void* h, float scale, const float* pos_table, const int32_t* lengths, int32_t batch, int32_t seq,
                       int32_t d, void* stream);
/* The same embedding that also PACKS the rows: out[cu_seqlens[b] + t, :] = h[b, t, :] * scale + pos[t + 2, :] for t < len_b
 * (len_b = cu_seqlens[b+1] - cu_seqlens[b]); padded frames of h are dropped.  h: [batch, seq, d] bf16, out: [total, d] bf16.
 * Entry into the packed (varlen) layout of the encoder: no padding rows in any GEMM / LayerNorm / attention after it. */
int jl_embed_positions_packed(const void* h, void* out, float scale, const float* pos_table, const int32_t* cu_seqlens,
                              int32_t batch, int32_t seq, int32_t d, void* stream);
/* Raw-waveform (wav2vec2 / XLS-R) front end — the second front end beside mel.  The convolutions run on
 * jl_gemm_bf16; these entry points build its operands.
 *  jl_wave_stats   stats[b] = (mean, 1/sqrt(var + 1e-7)) over the valid samples of utterance b — zero_mean_unit_var_norm,
 *                  SP/transformers/models/wav2vec2/feature_extraction_wav2vec2.py:78-97
 *  jl_wave_im2col  feature-encoder layer 0, Conv1d(1 → C, kernel <= 16, stride) (modeling_wav2vec2.py:275-299): normalises on
 *                  the fly and writes the bf16 im2col matrix [batch · t_out, 16] (unused taps and samples past the utterance = 0)
 *  jl_im2col_1d    x [batch, t_in, channels] bf16 → [batch · t_out, kernel · cg] (tap-major) for Conv1d(kernel, stride, pad)
 *                  over channels [c0, c0 + cg): encoder layers 1-6 and the grouped positional convolution (:326-368) */
int jl_wave_stats(const float* wave, int64_t wave_stride, const int32_t* num_samples, int32_t batch, int32_t max_samples, float* stats,
                  void* stream);
int jl_wave_im2col(const float* wave, int64_t wave_stride, const int32_t* num_samples, int32_t batch, int32_t max_samples, const float* stats,
                   void* out, int32_t t_out, int32_t kernel, int32_t stride, void* stream);
int jl_im2col_1d(const void* x, void* out, int32_t batch, int32_t t_in, int32_t channels, int32_t t_out, int32_t kernel, int32_t stride,
                 int32_t pad, int32_t c0, int32_t cg, void* stream);
int jl_transpose_bf16(const void* in, int64_t ld_in, void* out, int64_t ld_out, int32_t rows, int32_t cols, void* stream);
/* out[n] (+)= sum_m x[m, n]  (bias gradients) */
int jl_colsum_bf16(const void* x, int64_t ldx, float* out, int32_t rows, int32_t cols, float* partial, void* stream);
int jl_colsum_workspace_bytes(int32_t rows, int32_t cols, size_t* out);
int jl_cast_f32_to_bf16(const float* in, void* out, int64_t n, void* stream);
int jl_add_bf16(const void* a, const void* b, void* out, int64_t n, void* stream);

/* Fused AdamW over one flat fp32 bucket (adapter + lm_head parameters); also refreshes the bf16
 * shadow copy the kernels read.  grad is multiplied by grad_scale first (1/world after allreduce).
 * hyper_dev (optional): DEVICE pointer to 3 floats {lr, 1 - beta1^step, sqrt(1 - beta2^step)} read by the kernel instead of
 * lr / step — the launch then carries no step-dependent scalar and can live inside a captured CUDA graph (the host refreshes
 * the 12 bytes before each replay). */
typedef struct {
  float* param; const float* grad; float* exp_avg; float* exp_avg_sq; void* param_bf16;
  int64_t n;
  float lr, beta1, beta2, eps, weight_decay, grad_scale;
  int32_t step;
  const float* hyper_dev;
} jl_adamw_params;
int jl_adamw_bucket(const jl_adamw_params* p, void* stream);
/* Advance the device-side optimizer clock that `hyper_dev` launches read: hyper = {lr, bc1, bc2_sqrt, step, beta1, beta2}
 * (6 fp32 in device memory; the host writes lr / beta1 / beta2 and step = 0 once).  One thread: step += 1,
 * bc1 = 1 - beta1^step, bc2_sqrt = sqrt(1 - beta2^step).  Enqueued once per training step, before the AdamW launches, it makes
 * the whole step (forward, backward, all-reduce, AdamW) one replayable CUDA graph with no host-computed scalar in it. */
int jl_adamw_advance(float* hyper_dev, void* stream);

/* ------------------------------------------------------------------------------------------
 * a11: the one collective of the fine-tune step — sum of the flat fp32 adapter + lm_head gradient
 * bucket over the data-parallel ranks (what DDP's reducer does for the reference's trainable set,
 * /root/reference/requirements.txt:1,75).  One process per GPU; NCCL over NVLink 5 / NVSwitch.
 * libnccl.so.2 is resolved with dlopen at the first call (the copy the host process already has
 * loaded wins), so the library carries no link-time dependency on NCCL; without it every
 * jl_comm_* call returns JL_EUNSUPPORTED.
 *   rank 0: jl_comm_unique_id(id) → ship the JL_COMM_ID_BYTES bytes to every rank out of band →
 *   every rank: jl_comm_init(id, rank, world, &comm) on its own current device →
 *   per step: jl_comm_allreduce(comm, bucket, n, stream) (in place, enqueue only; graph-capturable) →
 *   jl_comm_destroy(comm).
 * ------------------------------------------------------------------------------------------ */
#define JL_COMM_ID_BYTES 128
typedef struct jl_comm jl_comm;
int jl_comm_unique_id(void* id_out);
int jl_comm_init(const void* id, int32_t rank, int32_t world, jl_comm** out);
int jl_comm_allreduce(jl_comm* comm, float* buf, size_t n_f32, void* stream);
int jl_comm_rank(const jl_comm* comm, int32_t* rank, int32_t* world);
int jl_comm_destroy(jl_comm* comm);

/* test / tuning hook: 1 = launch every kernel with programmatic dependent launch, 0 (default) = plain stream order */
void jl_debug_set_pdl(int on);

/* test / tuning hook: attention implementation — 0 = tcgen05 kernels (TMEM scores, TMA operands; utterances of <= 256 frames
 * take the whole-row forward kernel and the fused dQ/dK/dV backward kernel), 1 = mma.sync flash kernels, 2 / 3 = tcgen05
 * with the key-block forward kernel and the two-kernel (dQ, dKV) backward at every length (2: the forward's 3-CTAs/SM
 * build).  All are the product's own kernels. */
void jl_debug_set_attn_impl(int impl);

/* test / tuning hook: 0 = automatic kernel choice (default), 1 = single-CTA tcgen05 kernel only, 2 = CTA-pair
 * (cta_group::2) kernel wherever it is legal.  Both are the product's own kernels; results are identical up to fp32
 * accumulation order. */
void jl_debug_set_gemm_mode(int mode);
/* test / tuning hook: force the N tile (32/64/128/256 single-CTA kernel, 128/192/256 pair kernel with mode 2); 0 = automatic */
void jl_debug_set_gemm_bn(int bn);
/* test / tuning hook — what happens to the last, partial wave of output tiles (bit mask, default 2 | 4):
 *   bit 1 (2): single-CTA kernel: the tail tiles are cut into 2, 3 (96 + 96 + 64 columns of a 256-wide tile) or 4 column slices
 *              (independent, shorter work units);  bit 2 (4): never use the three-way cut;
 *   bit 0 (1): CTA-pair kernel: the tail tiles are cut into K ranges with an in-kernel fix-up (needs the workspace;
 *              validated, measured slower, off by default) */
void jl_debug_set_gemm_tail(int mode);

/* test-only device reference GEMM (SIMT fp32 accumulate) used by tests/ to check jl_gemm_bf16 at
 * sizes the CPU oracle cannot reach; never called by the product path. */
/* CTAs per row tile of jl_lnproj_bwd (column ranges): 0 = automatic, 1, 2, ... */
void jl_debug_set_lnproj_split(int split);
int jl_debug_gemm_ref(const jl_gemm_params* p, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* JL_B200_H */
