"""Drop-in PyTorch modules for the Jiao-Liao ASR path: ``JLEncoder`` (conv subsampler + pre-LN transformer with
``WFAdapter`` / ``AttAdapter`` slots), ``JLForCTC`` (encoder + CTC head).

The modules are ordinary ``nn.Module``s — parameters, ``state_dict``, ``.to()``, optimizers all work — but their
``forward`` never runs a PyTorch op on activations: ``JLEngine`` walks the layers and calls the C ABI
(``libjl_b200.so``: tcgen05 GEMMs, fused LayerNorm / attention / CTC kernels).  Training is adapter-only: the
backbone is frozen (``freeze_base_model``), the backward pass propagates dX through the frozen layers with the same
GEMM kernel in its MN-major operand modes and produces weight gradients only for the adapters and ``lm_head``.

Interfaces mirrored (SP = site-packages of the build container):
  * encoder forward           SP/transformers/models/speech_to_text/modeling_speech_to_text.py:561-608
  * pre-LN layer + adapter    SP/transformers/models/wav2vec2/modeling_wav2vec2.py:612-655 (hook :627-630,647-648)
  * CTC model forward / loss  SP/transformers/models/wav2vec2/modeling_wav2vec2.py:1675-1744
  * adapter management        :1046-1060 (_get_adapters), :1062-1073 (init_adapter_layers), :1075-1248 (load_adapter),
                              :1666-1672 (freeze_base_model)
WFAdapter / AttAdapter definitions: SURVEY.md §8c (from /root/reference/README.md:1 + BASELINE.json north_star).
"""
from __future__ import annotations

import math
import os
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn as nn

from . import _lib as L
from . import hf_compat, ops
from .configuration import JLConfig

BF16, F32, I32 = torch.bfloat16, torch.float32, torch.int32
_DEBUG_SKIP_SIDE = os.environ.get("JL_DEBUG_SKIP_SIDE") == "1"
# where the dγ / dβ of a trainable adapter LayerNorm are computed: "side" = a kernel of its own on the weight-gradient branch
# (re-reads dz and h), "main" = inside the LayerNorm backward kernel of the main chain (per-CTA partials + fixed-order reduce)
_LN_WGRAD = os.environ.get("JL_LN_WGRAD", "side")
# CUDA stream priority of the weight-gradient branch (0 = default, -1 = high: its kernel nodes outrank the main chain's when an SM frees up)
_SIDE_PRIORITY = int(os.environ.get("JL_SIDE_PRIORITY", "0"))
# the column reductions of an adapter's backward pass (bias gradients, LayerNorm dγ / dβ) as ONE launch (jl_colreduce_multi) instead
# of one kernel each.  Measured on B200: 24 fewer launches per step but SLOWER (6.38 vs 6.29 ms; 20.0 vs 19.8 ms on the 24-layer
# config) — the merged launch can only start when its last operand (dz) exists and then competes with the main chain as one large
# grid, where the three small kernels slip into the gaps as their operands appear.  Off by default.
_MERGED_REDUCE = os.environ.get("JL_MERGED_REDUCE", "0") == "1"


def subsampled_length(n, num_convs: int = 2):
    """(L - 1) // 2 + 1 per stride-2 conv (modeling_speech_to_text.py:489-496).  Works on ints and tensors."""
    for _ in range(num_convs):
        n = (n - 1) // 2 + 1
    return n


def sinusoid_table(num_rows: int, dim: int) -> torch.Tensor:
    """modeling_speech_to_text.py:123-139: [sin | cos] halves, log(10000)/(half-1) spacing, padding row 1 zeroed.
    Built once per device with the same fp32 operation sequence HF uses, so the table matches the reference's."""
    half = dim // 2
    e = math.log(10000) / (half - 1)
    e = torch.exp(torch.arange(half, dtype=torch.int64).float() * -e)
    e = torch.arange(num_rows, dtype=torch.int64).float().unsqueeze(1) * e.unsqueeze(0)
    tab = torch.cat([torch.sin(e), torch.cos(e)], dim=1).view(num_rows, -1)
    if dim % 2 == 1:
        tab = torch.cat([tab, torch.zeros(num_rows, 1)], dim=1)
    tab[1, :] = 0
    return tab


# =============================================================================================== modules
class WFAdapter(nn.Module):
    """Bottleneck adapter whose projections exist only as low-rank factors (weight factorisation):
    z = LN(h); u = relu((z B_dᵀ) A_dᵀ + c_d); y = (u B_uᵀ) A_uᵀ + c_u; out = h + y.
    Every factor carries a leading dialect dimension K; ``dialect`` selects the set used by a call."""

    kind = "wf"

    def __init__(self, hidden_size: int, bottleneck: int = 256, rank: int = 32, num_dialects: int = 1, eps: float = 1e-5):
        super().__init__()
        k, d, b, r = num_dialects, hidden_size, bottleneck, rank
        self.hidden_size, self.bottleneck, self.rank, self.num_dialects = d, b, r, k
        self.norm = nn.LayerNorm(d, eps=eps)
        self.down_B = nn.Parameter(torch.empty(k, r, d))
        self.down_A = nn.Parameter(torch.empty(k, b, r))
        self.down_bias = nn.Parameter(torch.empty(k, b))
        self.up_B = nn.Parameter(torch.empty(k, r, b))
        self.up_A = nn.Parameter(torch.empty(k, d, r))
        self.up_bias = nn.Parameter(torch.empty(k, d))
        self.reset_parameters()

    def reset_parameters(self, std: float = 0.02, generator: Optional[torch.Generator] = None):
        with torch.no_grad():
            for p in (self.down_B, self.down_A, self.up_B, self.up_A):
                p.copy_(torch.randn(p.shape, generator=generator) * std)
            self.down_bias.zero_()
            self.up_bias.zero_()
            self.norm.weight.fill_(1.0)
            self.norm.bias.zero_()


class AttAdapter(nn.Module):
    """Adapter that attends over the utterance's own hidden states: z = LN(h); q,k,v = z W_{q,k,v}ᵀ + b ∈ R^64;
    a = softmax(q kᵀ / 8 + keymask) v; out = h + a W_oᵀ + b_o.  One head of dim 64."""

    kind = "att"

    def __init__(self, hidden_size: int, att_dim: int = 64, eps: float = 1e-5):
        super().__init__()
        if att_dim != 64:
            raise ValueError("AttAdapter: att_dim must be 64")
        self.hidden_size, self.att_dim = hidden_size, att_dim
        self.norm = nn.LayerNorm(hidden_size, eps=eps)
        self.q_proj = nn.Linear(hidden_size, att_dim)
        self.k_proj = nn.Linear(hidden_size, att_dim)
        self.v_proj = nn.Linear(hidden_size, att_dim)
        self.o_proj = nn.Linear(att_dim, hidden_size)
        self.reset_parameters()

    def reset_parameters(self, std: float = 0.02, generator: Optional[torch.Generator] = None):
        with torch.no_grad():
            for lin in (self.q_proj, self.k_proj, self.v_proj, self.o_proj):
                lin.weight.copy_(torch.randn(lin.weight.shape, generator=generator) * std)
                lin.bias.zero_()
            self.norm.weight.fill_(1.0)
            self.norm.bias.zero_()


class FusionAdapter(nn.Module):
    """AdapterFusion-style AttAdapter over the K source-dialect adapters of the slot (SURVEY §8c ambiguity (ii), §8f f4): every
    frame runs all K factor sets of ``source`` (a WFAdapter, one set per source dialect), then attends over their K updates:
    y_k = WFAdapter_k(h) − h;  q = LN_f(h) W_qᵀ + b_q;  key_k = y_k W_kᵀ + b_k;  α = softmax_k(q·key_k / √b);  out = h + Σ_k α_k y_k.
    No dialect id is needed.  For knowledge transfer, train the source sets on the neighbouring dialects (kind "wf"), load them
    here, freeze ``source`` (``requires_grad_(False)``) and fine-tune norm / q_proj / k_proj on the target dialect."""

    kind = "fuse"

    def __init__(self, hidden_size: int, att_dim: int = 64, bottleneck: int = 256, rank: int = 32, num_dialects: int = 2, eps: float = 1e-5):
        super().__init__()
        if not 1 <= num_dialects <= 8:
            raise ValueError("FusionAdapter: 1 <= num_dialects <= 8")
        if att_dim % 8 or att_dim > 256:
            raise ValueError("FusionAdapter: att_dim must be a multiple of 8, <= 256")
        self.hidden_size, self.att_dim, self.num_dialects = hidden_size, att_dim, num_dialects
        self.source = WFAdapter(hidden_size, bottleneck, rank, num_dialects, eps)
        self.norm = nn.LayerNorm(hidden_size, eps=eps)
        self.q_proj = nn.Linear(hidden_size, att_dim)
        self.k_proj = nn.Linear(hidden_size, att_dim)
        self.reset_parameters()

    def reset_parameters(self, std: float = 0.02, generator: Optional[torch.Generator] = None):
        self.source.reset_parameters(std, generator)
        with torch.no_grad():
            for lin in (self.q_proj, self.k_proj):
                lin.weight.copy_(torch.randn(lin.weight.shape, generator=generator) * std)
                lin.bias.zero_()
            self.norm.weight.fill_(1.0)
            self.norm.bias.zero_()


def _make_adapter(kind: Optional[str], cfg: JLConfig) -> Optional[nn.Module]:
    if kind is None:
        return None
    if kind == "fuse":
        return FusionAdapter(cfg.hidden_size, cfg.att_dim, cfg.wf_bottleneck, cfg.wf_rank, cfg.num_dialects, cfg.layer_norm_eps)
    if kind == "wf":
        return WFAdapter(cfg.hidden_size, cfg.wf_bottleneck, cfg.wf_rank, cfg.num_dialects, cfg.layer_norm_eps)
    if kind == "att":
        return AttAdapter(cfg.hidden_size, cfg.att_dim, cfg.layer_norm_eps)
    raise ValueError(kind)


class JLAttention(nn.Module):
    def __init__(self, d: int):
        super().__init__()
        self.q_proj, self.k_proj, self.v_proj, self.out_proj = (nn.Linear(d, d) for _ in range(4))


class JLFeedForward(nn.Module):
    def __init__(self, d: int, inner: int):
        super().__init__()
        self.intermediate_dense = nn.Linear(d, inner)
        self.output_dense = nn.Linear(inner, d)


class JLEncoderLayer(nn.Module):
    def __init__(self, cfg: JLConfig):
        super().__init__()
        d = cfg.hidden_size
        self.layer_norm = nn.LayerNorm(d, eps=cfg.layer_norm_eps)
        self.attention = JLAttention(d)
        self.adapter_attn = _make_adapter(cfg.adapter_attn, cfg)
        self.final_layer_norm = nn.LayerNorm(d, eps=cfg.layer_norm_eps)
        self.feed_forward = JLFeedForward(d, cfg.intermediate_size)
        self.adapter_ffn = _make_adapter(cfg.adapter_ffn, cfg)


class JLWav2Vec2FrontEnd(nn.Module):
    """Parameters of the raw-waveform front end (XLS-R / MMS / wav2vec2-large: "layer" feature-extractor norm, conv bias):
    7 × [Conv1d → LayerNorm → GELU] (modeling_wav2vec2.py:275-299, 382-420), feature projection LayerNorm → Linear
    (:422-434), weight-normed grouped positional Conv1d (:326-368).  Frozen on this path."""

    def __init__(self, cfg: JLConfig):
        super().__init__()
        c, d = cfg.conv_dim, cfg.hidden_size
        self.conv = nn.ModuleList([nn.Conv1d(1 if i == 0 else c, c, k, stride=s) for i, (k, s) in enumerate(zip(cfg.conv_kernel, cfg.conv_stride))])
        self.conv_norm = nn.ModuleList([nn.LayerNorm(c, eps=1e-5) for _ in cfg.conv_kernel])
        self.proj_norm = nn.LayerNorm(c, eps=cfg.layer_norm_eps)
        self.proj = nn.Linear(c, d)
        self.pos_conv = nn.Module()
        kp, gp = cfg.num_conv_pos_embeddings, cfg.num_conv_pos_embedding_groups
        self.pos_conv.weight_v = nn.Parameter(torch.empty(d, d // gp, kp))
        self.pos_conv.weight_g = nn.Parameter(torch.empty(1, 1, kp))
        self.pos_conv.bias = nn.Parameter(torch.zeros(d))

    def reset_pos_conv(self, generator: Optional[torch.Generator] = None) -> None:
        """modeling_wav2vec2.py:966-973: N(0, 2·sqrt(1 / (k · in_channels))) weight, zero bias; weight_norm g = ‖v‖ per tap."""
        v = self.pos_conv.weight_v
        with torch.no_grad():
            v.copy_(torch.randn(v.shape, generator=generator) * (2.0 * math.sqrt(1.0 / (v.shape[2] * v.shape[0]))))
            self.pos_conv.weight_g.copy_(v.norm(dim=(0, 1), keepdim=True))
            self.pos_conv.bias.zero_()


def _sample_lengths(attention_mask, frame_lengths):
    """Valid length per utterance as the callers give it: explicit lengths win, else the mask's row sums, else None (= full)."""
    if frame_lengths is not None:
        return frame_lengths
    return attention_mask.sum(-1) if attention_mask is not None else None


def wav2vec2_lengths(num_samples, kernels, strides):
    """floor((L - k) / s) + 1 per conv layer (modeling_wav2vec2.py:1005-1020); works on ints and integer tensors."""
    n = num_samples
    for k, s in zip(kernels, strides):
        n = (n - k) // s + 1
    return n


class JLEncoder(nn.Module):
    """Front end (mel: Conv1d(k5,s2)+GLU ×2 → ×√d + sinusoid positions; wav2vec2: raw-waveform conv stack → projection →
    positional conv) → N pre-LN layers with adapter slots → final LayerNorm."""

    def __init__(self, cfg: JLConfig):
        super().__init__()
        self.config = cfg
        d = cfg.hidden_size
        if cfg.front_end == "wav2vec2":
            self.w2v = JLWav2Vec2FrontEnd(cfg)
            self.conv = nn.ModuleList()
        else:
            self.conv = nn.ModuleList([
                nn.Conv1d(cfg.input_feat_per_channel, cfg.conv_channels, 5, stride=2, padding=2),
                nn.Conv1d(cfg.conv_channels // 2, 2 * d, 5, stride=2, padding=2),
            ])
        self.layers = nn.ModuleList([JLEncoderLayer(cfg) for _ in range(cfg.num_hidden_layers)])
        self.layer_norm = nn.LayerNorm(d, eps=cfg.layer_norm_eps)
        self._engine: Optional["JLEngine"] = None

    def engine(self, lm_head: Optional[nn.Linear] = None) -> "JLEngine":
        if self._engine is None:
            self._engine = JLEngine(self, lm_head)
        if lm_head is not None:
            self._engine.lm_head = lm_head
        return self._engine

    def forward(self, input_features: torch.Tensor, attention_mask: Optional[torch.Tensor] = None,
                frame_lengths: Optional[torch.Tensor] = None, dialect: int = 0):
        """input_features [B, F, 80] (fp32 or bf16, CUDA) → last_hidden_state [B, T', d] bf16 (padded rows zeroed before the
        final LayerNorm, as the reference).  Also returns nothing else: use ``output_lengths`` for T' lengths."""
        eng = self.engine()
        lengths = eng.output_lengths(input_features, attention_mask, frame_lengths)
        st = eng.forward(input_features, lengths, training=False, dialect=dialect, want_logits=False,
                         sample_lengths=_sample_lengths(attention_mask, frame_lengths))
        b = input_features.shape[0]
        return st.h_final.view(b, st.t, self.config.hidden_size)


# =============================================================================================== engine
class PackedLayout:
    """Row layout of a packed ("varlen") batch: the T'_b valid frames of utterance b are rows [cu[b], cu[b+1]) of every
    [total, C] activation matrix — no padding rows, so the GEMMs, LayerNorms, adapters and the CTC head do no work on padding.
    ``cu`` is the device copy the attention / CTC kernels read; ``cu_host`` drives host-side slicing (dialect runs);
    ``seq_bound`` (a multiple of 128 >= the longest utterance) sizes the attention grid and the CTC workspace."""

    def __init__(self, lengths_host, device, cu: Optional[torch.Tensor] = None):
        self.lens = [int(x) for x in lengths_host]
        self.cu_host = [0]
        for n in self.lens:
            if n < 0:
                raise ValueError("negative utterance length")
            self.cu_host.append(self.cu_host[-1] + n)
        self.total = self.cu_host[-1]
        self.batch = len(self.lens)
        self.seq_bound = max(128, (max(self.lens + [1]) + 127) // 128 * 128)
        if cu is None:
            cu = torch.tensor(self.cu_host, dtype=I32, device=device)
        elif cu.numel() != self.batch + 1 or cu.dtype != I32:
            raise ValueError("cu must be an int32 tensor of batch + 1 entries")
        self.cu = cu

    def key(self):
        """What a captured CUDA graph bakes in (shapes and grids); the contents of ``cu`` may change between replays."""
        return (self.batch, self.total, self.seq_bound)


class _State:
    """Activations kept between forward and backward of one step."""
    pass


class _SideBranch:
    """Weight-gradient products do not feed the dX chain.  Two ways to keep them off the critical path:

    * ``defer=False`` — issue each on a second stream as soon as its operands exist: inside the captured CUDA graph they become
      parallel branches beside the main chain.  Measured cost on B200: the main chain's persistent GEMMs occupy every SM (128
      registers x 512 threads), so a branch kernel only ever runs in the gaps, and the ~86 small launches of a step (split-K
      products + their reduce kernels, column sums, LayerNorm dγ/dβ) cost the step 3-5 % (profiles/README.md, round 2).
    * ``defer=True`` — collect them and issue them all after the main chain's last kernel, round-robin over a pool of
      streams: every product then runs UNSPLIT on its few CTAs (K = all B·T' rows) with the products of all layers in flight at the
      same time — no split-K partials, no reduce kernels — at the price of an exposed tail.  Measured SLOWER (6.42 vs 6.33 ms on
      the headline config, 20.3 vs 19.8 ms on the 24-layer config): the unsplit K = 8000 products take ~40 µs each on their few
      CTAs and the tail is longer than what the overlap costs.  Kept as an option (``engine.defer_wgrads``), off by default.

    ``now=True`` issues immediately in both modes (the lm_head gradient, which the first half of the gradient exchange waits
    for).  Tensors a branch reads are kept alive until ``join`` so the caching allocator cannot recycle them early."""

    def __init__(self, enabled: bool = True, stream: Optional["torch.cuda.Stream"] = None, defer: bool = False, pool=None):
        self.enabled = enabled
        self.side = (stream if stream is not None else torch.cuda.Stream()) if enabled else None
        self.defer_mode = defer and enabled
        self.pool = pool or []
        self.deferred = []
        self.keep = []

    def run(self, fn, *tensors, now: bool = False) -> None:
        if _DEBUG_SKIP_SIDE:          # timing experiment only (wrong gradients): what the weight-gradient branch costs the step
            return
        if not self.enabled:
            fn()
            return
        self.keep.extend(tensors)
        if self.defer_mode and not now:
            self.deferred.append(fn)
            return
        self.side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(self.side):
            fn()

    def join(self) -> None:
        if self.enabled and not _DEBUG_SKIP_SIDE:
            main = torch.cuda.current_stream()
            if self.deferred:
                streams = self.pool if self.pool else [self.side]
                used = streams[: min(len(streams), len(self.deferred))]
                for st in used:
                    st.wait_stream(main)
                ops.GEMM_NO_SPLIT = True          # each product keeps its whole K on its own few CTAs: all layers run side by side
                try:
                    for i, fn in enumerate(self.deferred):
                        with torch.cuda.stream(used[i % len(used)]):
                            fn()
                finally:
                    ops.GEMM_NO_SPLIT = False
                for st in used:
                    main.wait_stream(st)
                self.deferred = []
            main.wait_stream(self.side)
        self.keep.clear()


class JLEngine:
    """Walks the module tree and issues the C-ABI calls.  Holds bf16 copies of the weights in the layouts the kernels
    read (q/k/v concatenated; conv weights tap-major with GLU rows interleaved)."""

    def __init__(self, encoder: JLEncoder, lm_head: Optional[nn.Linear] = None):
        L.load()
        self.enc = encoder
        self.cfg = encoder.config
        self.lm_head = lm_head
        self._frozen = None
        self._frozen_key = None
        self._shadow: Dict[int, Tuple[int, int, torch.Tensor]] = {}
        self._pos: Dict[Tuple[str, int], torch.Tensor] = {}
        self.flat = None   # set by training.FlatAdapterParams
        self.side_branch = True   # issue weight-gradient products on a second stream (parallel graph branches)
        # adapter weight gradients: all at the end of the backward pass, unsplit and concurrent (True), or each as soon as its
        # operands exist, beside the main chain (False) — see _SideBranch
        self.defer_wgrads = os.environ.get("JL_DEFER_WGRADS", "0") == "1"     # measured slower on B200 (profiles/README.md): off
        self._side_pools: Dict[int, list] = {}
        self._side_streams: Dict[int, "torch.cuda.Stream"] = {}   # one side stream per device, created once
        self.fused_wf = True      # inference: WFAdapter as one kernel (jl_wfadapter_fwd)
        # training: the same kernel, which then also writes the intermediates the backward pass needs (t1, u, t2, LayerNorm
        # statistics); its LayerNorm-folded operands are re-derived on the device at every step (jl_wfadapter_pack).  Parity-tested,
        # but measured no faster than LN + 4 GEMMs on B200 (24-layer config 19.96 vs 20.02 ms; mixed-length config with four dialect
        # runs per layer 15.49 vs 15.15 ms: one CTA per SM and a four-stage dependency chain per 128 rows) — off by default.
        self.fused_wf_train = os.environ.get("JL_FUSED_WF_TRAIN", "0") == "1"
        # AttAdapter forward as one kernel (jl_attadapter_fwd) for utterances of <= 256 frames, inference and training
        self.fused_att = os.environ.get("JL_FUSED_ATT", "1") != "0"
        # tail of the AttAdapter backward (dqkv · W_qkv + LayerNorm backward) as one kernel (jl_lnproj_bwd)
        self.fused_att_bwd = os.environ.get("JL_FUSED_ATT_BWD", "1") != "0"
        # dW_qkv without LN(h) (jl_lnproj_wgrad): 0 = off (LN(h) recomputed on the weight-gradient branch), 1 = operands from the prologue of
        # jl_lnproj_bwd (measured slower in the step: 6.07 vs 6.03 ms), 2 = operands from jl_lnproj_wgrad_prep on the weight-gradient branch
        self.lp_wgrad = int(os.environ.get("JL_LP_WGRAD", "0"))
        # CTAs per row tile of jl_lnproj_bwd in the AttAdapter backward: 2 (lower latency) when the layer's weight-gradient branch is light
        # (one adapter per layer: base config 5.95 vs 5.99 ms), 1 (less SM-time) with two adapters per layer (24-layer config 18.30 vs 18.48 ms)
        two = any(l.adapter_attn is not None and l.adapter_ffn is not None for l in self.enc.layers)
        self.lp_col_split = int(os.environ.get("JL_LNPROJ_COL_SPLIT", "1" if two else "2"))
        # clusters per utterance of the fused AttAdapter forward (they share the output columns, each repeats the attention): 2 lowers the
        # kernel's latency (64 -> 128 CTAs at 32 utterances): base config 5.97 -> 5.92 ms, 24-layer config 18.76 -> 18.61 ms
        self.att_col_split = int(os.environ.get("JL_ATT_COL_SPLIT", "2"))
        self._att_bufs: Dict[int, dict] = {}
        self._vparams = None
        self._att_packed_step = False
        self._wf_fold: Dict[int, list] = {}
        # tail of the WFAdapter backward (dt1 · B_d + LayerNorm backward + column sums) through jl_lnproj_bwd
        self.fused_wf_bwd = os.environ.get("JL_FUSED_WF_BWD", "1") != "0"
        self.wf_multi_run = os.environ.get("JL_WF_MULTI_RUN", "1") != "0"     # several dialect runs per batch: one jl_lnproj_bwd launch with row runs
        self._wf_bufs: Dict[int, dict] = {}

    def _side_stream(self, device) -> "torch.cuda.Stream":
        idx = device.index if device.index is not None else torch.cuda.current_device()
        st = self._side_streams.get(idx)
        if st is None:
            st = torch.cuda.Stream(device=idx, priority=_SIDE_PRIORITY)
            self._side_streams[idx] = st
        return st

    def _side_pool(self, device, n: int = 8) -> list:
        idx = device.index if device.index is not None else torch.cuda.current_device()
        pool = self._side_pools.get(idx)
        if pool is None:
            pool = [torch.cuda.Stream(device=idx) for _ in range(n)]
            self._side_pools[idx] = pool
        return pool

    # ------------------------------------------------------------------ weights
    def weights_version(self, include_optimizer_steps: bool = True) -> int:
        """Changes whenever a parameter of the model (backbone, adapters, lm_head) is modified through torch (``copy_``,
        optimizer step, ``load_state_dict``, ``load_adapter``, ``init_adapter_layers`` …) or re-allocated.  Captured CUDA graphs
        bake in the pointers of the bf16 shadows / packed weights derived from the parameters, so ``AdapterTrainer`` and
        ``Transcriber`` compare this number before every replay and rebuild what is stale.  The fused AdamW kernel updates the flat
        bucket behind torch's back; it is counted through ``flat.generation`` (left out for the trainer's own check: its graph
        reads the bucket's bf16 shadow, which that kernel refreshes in place)."""
        v = self.flat.generation if (include_optimizer_steps and self.flat is not None) else 0
        # the module walk (≈ 0.8 ms for 320 parameters) is done once; the per-call cost is one pass over the cached Parameter
        # objects.  Parameter objects are replaced only when a module is (lm_head on a vocabulary change: tracked by identity).
        key = id(self.lm_head)
        if self._vparams is None or self._vparams[0] != key:
            ps = list(self.enc.parameters())
            if self.lm_head is not None:
                ps += list(self.lm_head.parameters())
            self._vparams = (key, ps)
        return hash((v, tuple([(p._version, p.data_ptr()) for p in self._vparams[1]])))

    def _backbone_params(self):
        for n, p in self.enc.named_parameters():
            if ".adapter_attn." not in n and ".adapter_ffn." not in n:
                yield n, p

    def _frozen_pack(self):
        key = tuple((p.data_ptr(), p._version) for _, p in self._backbone_params())
        if self._frozen is not None and key == self._frozen_key:
            return self._frozen
        fz = {}
        with torch.no_grad():
            for i, conv in enumerate(self.enc.conv):
                cout, cin, k = conv.weight.shape
                half = cout // 2
                idx = torch.stack([torch.arange(half), torch.arange(half) + half], dim=1).reshape(-1).to(conv.weight.device)
                w = conv.weight.permute(0, 2, 1).reshape(cout, k * cin)[idx]
                fz[f"conv{i}.w"] = w.to(BF16).contiguous()
                fz[f"conv{i}.b"] = conv.bias[idx].to(F32).contiguous()
            if self.cfg.front_end == "wav2vec2":
                fe = self.enc.w2v
                for i, conv in enumerate(fe.conv):
                    cout, cin, k = conv.weight.shape
                    w = conv.weight.permute(0, 2, 1).reshape(cout, k * cin)               # tap-major columns
                    if i == 0:                                                              # K padded to 16 (taps k..15 are zero)
                        w = torch.cat([w, torch.zeros((cout, 16 - k), device=w.device, dtype=w.dtype)], 1)
                    fz[f"w2v.conv{i}.w"] = w.to(BF16).contiguous()
                    fz[f"w2v.conv{i}.b"] = conv.bias.to(F32).contiguous()
                fz["w2v.proj.w"] = fe.proj.weight.to(BF16).contiguous()
                fz["w2v.proj.b"] = fe.proj.bias.to(F32).contiguous()
                v, g = fe.pos_conv.weight_v.float(), fe.pos_conv.weight_g.float()
                wn = g * v / v.norm(dim=(0, 1), keepdim=True)                             # weight_norm(dim = 2)
                groups = self.cfg.num_conv_pos_embedding_groups
                cg = wn.shape[0] // groups
                for gi in range(groups):                                                    # per group: [cg_out, k · cg_in] tap-major
                    wg = wn[gi * cg:(gi + 1) * cg].permute(0, 2, 1).reshape(cg, -1)
                    fz[f"w2v.pos{gi}.w"] = wg.to(BF16).contiguous()
                fz["w2v.pos.b"] = fe.pos_conv.bias.to(F32).contiguous()
            for i, layer in enumerate(self.enc.layers):
                a = layer.attention
                fz[f"{i}.wqkv"] = torch.cat([a.q_proj.weight, a.k_proj.weight, a.v_proj.weight], 0).to(BF16).contiguous()
                fz[f"{i}.bqkv"] = torch.cat([a.q_proj.bias, a.k_proj.bias, a.v_proj.bias], 0).to(F32).contiguous()
                fz[f"{i}.wo"] = a.out_proj.weight.to(BF16).contiguous()
                fz[f"{i}.bo"] = a.out_proj.bias.to(F32).contiguous()
                fz[f"{i}.w1"] = layer.feed_forward.intermediate_dense.weight.to(BF16).contiguous()
                fz[f"{i}.b1"] = layer.feed_forward.intermediate_dense.bias.to(F32).contiguous()
                fz[f"{i}.w2"] = layer.feed_forward.output_dense.weight.to(BF16).contiguous()
                fz[f"{i}.b2"] = layer.feed_forward.output_dense.bias.to(F32).contiguous()
        self._frozen, self._frozen_key = fz, key
        self._frozen_t = None
        return fz

    def _frozen_pack_t(self):
        """Transposed bf16 copies of the frozen projection weights for the dgrad products dY · W: with W stored [in, out]
        the B operand is K-major, which the tcgen05 kernel streams ≈ 1.3-2× faster than the MN-major view of [out, in]
        (measured: profiles/).  Built once — the backbone is frozen."""
        fz = self._frozen_pack()
        if getattr(self, "_frozen_t", None) is None:
            ft = {}
            with torch.no_grad():
                for i in range(len(self.enc.layers)):
                    for k in ("wqkv", "wo", "w1", "w2"):
                        ft[f"{i}.{k}"] = fz[f"{i}.{k}"].t().contiguous()
            self._frozen_t = ft
        return self._frozen_t

    def _bf16(self, p: torch.Tensor) -> torch.Tensor:
        """bf16 copy of a trainable fp32 parameter, refreshed (by the cast kernel) when the parameter changed."""
        if self.flat is not None:
            v = self.flat.bf16_view(p)
            if v is not None:
                return v
        ent = self._shadow.get(id(p))
        if ent is not None and ent[0] == p.data_ptr() and ent[1] == p._version:
            return ent[2]
        sh = ops.cast_bf16(p.detach().contiguous())
        self._shadow[id(p)] = (p.data_ptr(), p._version, sh)
        return sh

    def _cat_bf16(self, params: List[torch.Tensor]) -> torch.Tensor:
        if self.flat is not None:
            v = self.flat.bf16_cat_view(params)
            if v is not None:
                return v
        key = tuple(id(p) for p in params)
        ver = tuple((p.data_ptr(), p._version) for p in params)
        ent = self._shadow.get(key)
        if ent is not None and ent[0] == ver:
            return ent[2]
        with torch.no_grad():
            sh = ops.cast_bf16(torch.cat([p.detach() for p in params], 0).contiguous())
        self._shadow[key] = (ver, None, sh)
        return sh

    def _cat_f32(self, params: List[torch.Tensor]) -> torch.Tensor:
        if self.flat is not None:
            v = self.flat.f32_cat_view(params)
            if v is not None:
                return v
        with torch.no_grad():
            return torch.cat([p.detach() for p in params], 0).contiguous()

    def _wf_fusable(self, ad) -> bool:
        return (ad.hidden_size % 128 == 0 and ad.rank % 16 == 0 and 16 <= ad.rank <= 64 and ad.bottleneck % 64 == 0
                and 64 <= ad.bottleneck <= 256)

    def _wf_pack(self, ad, k: int) -> dict:
        """Factors of dialect ``k`` in the layouts the fused kernel reads: B_d ⊙ γ (LayerNorm folded into the first
        projection) with its row sums s and the β term t, rank dimension of A_d / A_u zero-padded to 64."""
        ps = [ad.norm.weight, ad.norm.bias, ad.down_B, ad.down_A, ad.down_bias, ad.up_B, ad.up_A, ad.up_bias]
        ver = tuple((q.data_ptr(), q._version) for q in ps) + (k, self.flat.generation if self.flat is not None else 0)
        key = ("wf", id(ad), k)
        ent = self._shadow.get(key)
        if ent is not None and ent[0] == ver:
            return ent[2]
        with torch.no_grad():
            gamma, beta = ad.norm.weight.detach().float(), ad.norm.bias.detach().float()
            bd = (ad.down_B.detach()[k].float() * gamma[None, :]).to(BF16).contiguous()
            r, b, d = ad.rank, ad.bottleneck, ad.hidden_size
            adp = torch.zeros((b, 64), dtype=BF16, device=bd.device)
            adp[:, :r] = ad.down_A.detach()[k].to(BF16)
            aup = torch.zeros((d, 64), dtype=BF16, device=bd.device)
            aup[:, :r] = ad.up_A.detach()[k].to(BF16)
            pack = {"bd": bd, "s": bd.float().sum(1).contiguous(), "t": (ad.down_B.detach()[k].float() @ beta).contiguous(),
                    "ad": adp, "c_d": ad.down_bias.detach()[k].float().contiguous(), "bu": ad.up_B.detach()[k].to(BF16).contiguous(),
                    "au": aup, "c_u": ad.up_bias.detach()[k].float().contiguous(), "r": r, "b": b}
        self._shadow[key] = (ver, None, pack)
        return pack

    def _wf_pack_dev(self, ad) -> dict:
        """Kernel-layout operands of every factor set of ``ad``, derived ON THE DEVICE from the bf16 shadows by one launch into
        buffers that keep their addresses — inside a captured training step they follow the optimizer's updates at every replay."""
        bufs = ops.wfadapter_pack(self._bf16(ad.down_B), self._bf16(ad.down_A), self._bf16(ad.up_A), ad.norm.weight.detach(),
                                  ad.norm.bias.detach(), self._wf_bufs.get(id(ad)))
        self._wf_bufs[id(ad)] = bufs
        return bufs

    def _att_pack_dev(self, ad, training: bool, reuse: bool = False) -> dict:
        """LayerNorm-folded q|k|v projection of an AttAdapter (jl_lnfold_pack), derived on the device into buffers that keep their
        addresses.  Training: re-derived at every call (inside the captured step it follows the optimizer); inference: only when
        a parameter changed."""
        ps = [ad.norm.weight, ad.norm.bias, ad.q_proj.weight, ad.k_proj.weight, ad.v_proj.weight, ad.q_proj.bias, ad.k_proj.bias, ad.v_proj.bias]
        ver = tuple((q.data_ptr(), q._version) for q in ps) + (self.flat.generation if self.flat is not None else 0,)
        ent = self._att_bufs.get(id(ad))
        if ent is not None and training and self._att_packed_step:
            return ent[1]          # packed by _att_pack_all at the start of this step's forward pass
        if ent is not None and (reuse or (not training and ent[0] == ver)):
            return ent[1]          # reuse: the backward pass of the step whose forward pass derived the pack
        w = self._cat_bf16([ad.q_proj.weight, ad.k_proj.weight, ad.v_proj.weight])
        bq = self._cat_f32([ad.q_proj.bias, ad.k_proj.bias, ad.v_proj.bias])
        bufs = ops.lnfold_pack(w, bq, ad.norm.weight.detach(), ad.norm.bias.detach(), None if ent is None else ent[1])
        self._att_bufs[id(ad)] = (ver, bufs)
        return bufs

    def _wf_fold_packs(self, ad) -> dict:
        """LayerNorm-fold vectors (s, tb) of the first low-rank projection of every factor set of a WFAdapter (B_d[k] [r, d], no bias),
        for jl_lnproj_bwd: {"k": [per-set dicts], "all": {"s", "tb"} the same vectors as contiguous [K · r] arrays (multi-run calls)}.
        Derived on the device into buffers that keep their addresses (inside a captured step they follow the optimizer)."""
        w = self._bf16(ad.down_B)
        ent = self._wf_fold.get(id(ad))
        if ent is not None and self._att_packed_step:
            return ent             # derived by _att_pack_all at the start of this step's forward pass
        if ent is None:
            kk, r, d = w.shape
            s_all = torch.empty((kk, r), dtype=F32, device=w.device)
            tb_all = torch.empty((kk, r), dtype=F32, device=w.device)
            scratch = torch.empty((kk, r, d), dtype=BF16, device=w.device)
            ent = {"all": {"s": s_all.view(-1), "tb": tb_all.view(-1)}, "k": [{"w": scratch[k], "s": s_all[k], "tb": tb_all[k]} for k in range(kk)]}
            self._wf_fold[id(ad)] = ent
        ops.lnfold_pack_multi([(w[k], None, ad.norm.weight.detach(), ad.norm.bias.detach(), ent["k"][k]) for k in range(ad.num_dialects)])
        return ent

    def _att_pack_all(self) -> None:
        """Training step: the LayerNorm-folded q|k|v projections of EVERY AttAdapter in one launch at the start of the forward pass
        (they depend on the weights only; one launch per adapter cost 12 x 4.6 µs on the main chain).  The per-adapter calls of this
        step then reuse the buffers."""
        jobs = []
        for layer in self.enc.layers:
            for ad in (layer.adapter_attn, layer.adapter_ffn):
                if ad is None or ad.kind != "att":
                    continue
                ent = self._att_bufs.get(id(ad))
                if ent is None:
                    self._att_pack_dev(ad, True)             # first use: allocates the buffers (and packs)
                    continue
                w = self._cat_bf16([ad.q_proj.weight, ad.k_proj.weight, ad.v_proj.weight])
                bq = self._cat_f32([ad.q_proj.bias, ad.k_proj.bias, ad.v_proj.bias])
                jobs.append((w, bq, ad.norm.weight.detach(), ad.norm.bias.detach(), ent[1]))
        if self.fused_wf_bwd:
            # ... and the fold vectors of the WFAdapters' first projections (jl_lnproj_bwd in their backward pass)
            for layer in self.enc.layers:
                for ad in (layer.adapter_attn, layer.adapter_ffn):
                    if ad is None or ad.kind != "wf":
                        continue
                    ent = self._wf_fold.get(id(ad))
                    if ent is None:
                        continue                                 # first step: allocated (and packed) where they are first needed
                    w = self._bf16(ad.down_B)
                    for k in range(ad.num_dialects):
                        jobs.append((w[k], None, ad.norm.weight.detach(), ad.norm.bias.detach(), ent["k"][k]))
        if jobs:
            ops.lnfold_pack_multi(jobs)
        self._att_packed_step = True

    def pos_table(self, device, rows: int) -> torch.Tensor:
        key = (str(device), self.cfg.hidden_size)
        tab = self._pos.get(key)
        if tab is None or tab.shape[0] < rows:
            tab = sinusoid_table(max(rows, 1024), self.cfg.hidden_size).to(device).contiguous()
            self._pos[key] = tab
        return tab

    # ------------------------------------------------------------------ lengths
    def output_lengths(self, input_features, attention_mask=None, frame_lengths=None) -> torch.Tensor:
        """T' lengths (int32, device) after the two stride-2 convs (mel) / the conv stack (wav2vec2: ``input_features`` is the
        waveform batch [B, N], ``attention_mask`` its sample mask, ``frame_lengths`` the valid sample counts)."""
        if self.cfg.front_end == "wav2vec2":
            if frame_lengths is None:
                if attention_mask is None:
                    frame_lengths = torch.full((input_features.shape[0],), input_features.shape[1], dtype=I32, device=input_features.device)
                else:
                    frame_lengths = attention_mask.sum(-1)
            n = wav2vec2_lengths(frame_lengths.to(torch.int64), self.cfg.conv_kernel, self.cfg.conv_stride)
            return n.clamp_min(0).to(I32)
        if frame_lengths is None:
            if attention_mask is None:
                frame_lengths = torch.full((input_features.shape[0],), input_features.shape[1], dtype=I32, device=input_features.device)
            else:
                frame_lengths = attention_mask.sum(-1)
        return subsampled_length(frame_lengths.to(torch.int64)).to(I32)

    # ------------------------------------------------------------------ adapters
    @staticmethod
    def dialect_segments(dialect, b: int, num_dialects: int):
        """``dialect`` = one id for the whole batch or one id per utterance (SURVEY §8c ``dialect_ids [B]``) → runs
        ``[(k, b0, b1), …]`` of adjacent utterances that share a factor set.  Utterances of one dialect must be adjacent
        (sort the batch by dialect): every run is one slice of the [B·T', d] activation matrix, and every factor set
        has exactly one gradient writer."""
        if isinstance(dialect, int):
            ids = [dialect] * b
        else:
            ids = [int(k) for k in (dialect.tolist() if torch.is_tensor(dialect) else dialect)]
            if len(ids) != b:
                raise ValueError(f"dialect ids: expected {b} entries (one per utterance), got {len(ids)}")
        segs = []
        for i, k in enumerate(ids):
            if not 0 <= k < num_dialects:
                raise ValueError(f"dialect id {k} out of range for {num_dialects} factor set(s)")
            if segs and segs[-1][0] == k:
                segs[-1][2] = i + 1
            else:
                segs.append([k, i, i + 1])
        seen = [k for k, _, _ in segs]
        if len(set(seen)) != len(seen):
            raise ValueError("utterances of one dialect must be adjacent in the batch (sort the batch by dialect id)")
        return [tuple(x) for x in segs]

    @staticmethod
    def _seg_rows(b0: int, b1: int, t: int, pk: Optional[PackedLayout]) -> slice:
        """Rows of the utterances [b0, b1) in the [rows, C] activation matrices."""
        return slice(b0 * t, b1 * t) if pk is None else slice(pk.cu_host[b0], pk.cu_host[b1])

    def _adapter_fwd(self, ad: nn.Module, h: torch.Tensor, lengths, b: int, t: int, training: bool, dialect, zero_rows: bool,
                     pk: Optional[PackedLayout] = None):
        """Returns (out, saved).  out = h + adapter(h); padded rows zeroed when ``zero_rows`` (end of a layer; the packed layout
        ``pk`` has no padded rows)."""
        eps = ad.norm.eps
        zero_rows = zero_rows and pk is None
        cu = None if pk is None else pk.cu
        if ad.kind == "fuse":
            return self._fusion_fwd(ad, h, lengths, t, training, zero_rows)
        segs = self.dialect_segments(dialect, b, ad.num_dialects) if ad.kind == "wf" else None
        if ad.kind == "wf" and not training and self.fused_wf and self._wf_fusable(ad):
            # inference: the whole adapter is one kernel per dialect run (LN folded into the first projection); training
            # keeps the composed path because the backward needs every intermediate
            out = torch.empty_like(h)
            for k, b0, b1 in segs:
                rows = self._seg_rows(b0, b1, t, pk)
                if rows.stop == rows.start:
                    continue
                ops.wfadapter_fwd(h[rows], self._wf_pack(ad, k), eps, row_lengths=lengths[b0:b1] if zero_rows else None,
                                  rows_per_seq=t if zero_rows else 0, out=out[rows])
            return out, None
        if ad.kind == "wf" and training and self.fused_wf_train and self._wf_fusable(ad):
            # training: one kernel per dialect run as well; it also writes t1, u, t2 and the LayerNorm statistics.  LN(h) itself is
            # only needed by one weight-gradient product and is recomputed on the weight-gradient branch (``_adapter_bwd``).
            m, dev = h.shape[0], h.device
            bufs = self._wf_pack_dev(ad)
            bu16 = self._bf16(ad.up_B)
            mean = torch.empty((m,), dtype=F32, device=dev)
            rstd = torch.empty((m,), dtype=F32, device=dev)
            t1 = torch.empty((m, ad.rank), dtype=BF16, device=dev)
            u = torch.empty((m, ad.bottleneck), dtype=BF16, device=dev)
            t2 = torch.empty((m, ad.rank), dtype=BF16, device=dev)
            out = torch.empty_like(h)
            for k, b0, b1 in segs:
                rows = self._seg_rows(b0, b1, t, pk)
                if rows.stop == rows.start:
                    continue
                pack = {"bd": bufs["bd"][k], "s": bufs["s"][k], "t": bufs["t"][k], "ad": bufs["ad"][k], "au": bufs["au"][k], "bu": bu16[k],
                        "c_d": ad.down_bias.detach()[k], "c_u": ad.up_bias.detach()[k], "r": ad.rank, "b": ad.bottleneck}
                ops.wfadapter_fwd(h[rows], pack, eps, row_lengths=lengths[b0:b1] if zero_rows else None, rows_per_seq=t if zero_rows else 0,
                                  out=out[rows], mean=mean[rows], rstd=rstd[rows], t1=t1[rows], u=u[rows], t2=t2[rows])
            return out, (h, mean, rstd, None, t1, u, t2, segs)
        if ad.kind == "att" and self.fused_att and t <= 256 and ad.hidden_size % 128 == 0 and ad.hidden_size <= 1024:
            # the whole adapter in one kernel (LayerNorm folded into the q|k|v projection, attention, output projection, residual)
            out, sv = ops.attadapter_fwd(h, self._att_pack_dev(ad, training), self._bf16(ad.o_proj.weight), ad.o_proj.bias.detach(), lengths, b, t,
                                         eps, zero_padded_rows=zero_rows, training=training, cu_seqlens=cu, col_split=self.att_col_split)
            if not training:
                return out, None
            mean, rstd, qkv, a, lse = sv
            return out, (h, mean, rstd, None, qkv, a, lse)
        z, mean, rstd = ops.layernorm_fwd(h, ad.norm.weight.detach(), ad.norm.bias.detach(), eps, save_stats=training)
        rl = dict(row_lengths=lengths, rows_per_seq=t) if zero_rows else {}
        if ad.kind == "wf":
            m, dev = h.shape[0], h.device
            t1 = torch.empty((m, ad.rank), dtype=BF16, device=dev)
            u = torch.empty((m, ad.bottleneck), dtype=BF16, device=dev)
            t2 = torch.empty((m, ad.rank), dtype=BF16, device=dev)
            out = torch.empty_like(h)
            for k, b0, b1 in segs:
                rows = self._seg_rows(b0, b1, t, pk)
                if rows.stop == rows.start:
                    continue
                rls = dict(row_lengths=lengths[b0:b1], rows_per_seq=t) if zero_rows else {}
                ops.gemm(z[rows], self._bf16(ad.down_B)[k], out=t1[rows])
                ops.gemm(t1[rows], self._bf16(ad.down_A)[k], bias=ad.down_bias.detach()[k], epilogue=L.JL_EPI_RELU, out=u[rows])
                ops.gemm(u[rows], self._bf16(ad.up_B)[k], out=t2[rows])
                ops.gemm(t2[rows], self._bf16(ad.up_A)[k], bias=ad.up_bias.detach()[k], residual=h[rows], out=out[rows], **rls)
            saved = (h, mean, rstd, z, t1, u, t2, segs) if training else None
        else:
            wqkv = self._cat_bf16([ad.q_proj.weight, ad.k_proj.weight, ad.v_proj.weight])
            bqkv = self._cat_f32([ad.q_proj.bias, ad.k_proj.bias, ad.v_proj.bias])
            qkv = ops.gemm(z, wqkv, bias=bqkv)
            a, lse = ops.attn_fwd(qkv[:, 0:64], qkv[:, 64:128], qkv[:, 128:192], lengths, b, t, 1, 1.0 / 8.0, want_lse=training,
                                  cu_seqlens=cu)
            out = ops.gemm(a, self._bf16(ad.o_proj.weight), bias=ad.o_proj.bias.detach(), residual=h, **rl)
            saved = (h, mean, rstd, z, qkv, a, lse) if training else None
        return out, saved

    # ------------------------------------------------------------------ f4: AdapterFusion-style adapter
    def _fusion_fwd(self, ad: "FusionAdapter", h: torch.Tensor, lengths, t: int, training: bool, zero_rows: bool):
        """All K source-dialect factor sets on every row (the first projection of all sets is ONE GEMM with N = K·r, the key
        projection ONE GEMM over the K·M rows of y), then the fusion combine kernel.  Returns (out, saved)."""
        src = ad.source
        kk, r, bt, d, m = src.num_dialects, src.rank, src.bottleneck, src.hidden_size, h.shape[0]
        dev = h.device
        zs, mean_s, rstd_s = ops.layernorm_fwd(h, src.norm.weight.detach(), src.norm.bias.detach(), src.norm.eps, save_stats=training)
        t1 = ops.gemm(zs, self._bf16(src.down_B).view(kk * r, d))                                  # [m, K·r]
        u = torch.empty((kk, m, bt), dtype=BF16, device=dev)
        t2 = torch.empty((m, kk * r), dtype=BF16, device=dev)
        y = torch.empty((kk, m, d), dtype=BF16, device=dev)
        for k in range(kk):
            cs = slice(k * r, (k + 1) * r)
            ops.gemm(t1[:, cs], self._bf16(src.down_A)[k], bias=src.down_bias.detach()[k], epilogue=L.JL_EPI_RELU, out=u[k])
            ops.gemm(u[k], self._bf16(src.up_B)[k], out=t2[:, cs])
            ops.gemm(t2[:, cs], self._bf16(src.up_A)[k], bias=src.up_bias.detach()[k], out=y[k])      # the update only: no residual
        zf, mean_f, rstd_f = ops.layernorm_fwd(h, ad.norm.weight.detach(), ad.norm.bias.detach(), ad.norm.eps, save_stats=training)
        q = ops.gemm(zf, self._bf16(ad.q_proj.weight), bias=ad.q_proj.bias.detach())
        key = ops.gemm(y.view(kk * m, d), self._bf16(ad.k_proj.weight), bias=ad.k_proj.bias.detach()).view(kk, m, ad.att_dim)
        scale = 1.0 / math.sqrt(ad.att_dim)
        out, alpha = ops.fusion_combine_fwd(h, y, q, key, scale, row_lengths=lengths if zero_rows else None, rows_per_seq=t if zero_rows else 0)
        saved = (h, zs, mean_s, rstd_s, t1, u, t2, y, zf, mean_f, rstd_f, q, key, alpha, scale) if training else None
        return out, saved

    def _fusion_bwd(self, ad: "FusionAdapter", saved, dout: torch.Tensor, g: "GradSink", sb: "_SideBranch") -> torch.Tensor:
        MN = L.JL_LAYOUT_MN
        h, zs, mean_s, rstd_s, t1, u, t2, y, zf, mean_f, rstd_f, q, key, alpha, scale = saved
        src = ad.source
        kk, r, d, m, b = src.num_dialects, src.rank, src.hidden_size, h.shape[0], ad.att_dim
        dy, dq, dkey = ops.fusion_combine_bwd(dout, y, q, key, alpha, scale)
        dkey2, y2, dy2 = dkey.view(kk * m, b), y.view(kk * m, d), dy.view(kk * m, d)
        # knowledge transfer freezes the source-dialect adapters and trains the fusion only: their weight gradients are then skipped
        # (the gradient still flows THROUGH them to the layers below)
        train_src = src.down_B.requires_grad

        def w_fuse():
            ops.gemm(dkey2, y2, a_layout=MN, b_layout=MN, out=g.out(ad.k_proj.weight), out_dtype=F32)      # dkeyᵀ · y
            ops.colsum(dkey2, out=g.out(ad.k_proj.bias))
            ops.gemm(dq, zf, a_layout=MN, b_layout=MN, out=g.out(ad.q_proj.weight), out_dtype=F32)         # dqᵀ · LN_f(h)
            ops.colsum(dq, out=g.out(ad.q_proj.bias))
        sb.run(w_fuse, dkey2, y2, dq, zf)
        # the keys are projections of y: dy_k += dkey_k · W_k
        dyt = ops.gemm(dkey2, self._bf16(ad.k_proj.weight), b_layout=MN, residual=dy2)
        dzf = ops.gemm(dq, self._bf16(ad.q_proj.weight), b_layout=MN)
        sb.run(lambda: ops.layernorm_wgrad(dzf, h, mean_f, rstd_f, g.out(ad.norm.weight), g.out(ad.norm.bias)), dzf, h, mean_f, rstd_f)
        dh, _, _ = ops.layernorm_bwd(dzf, h, ad.norm.weight.detach(), mean_f, rstd_f, dres=dout)
        # the K source adapters, each on every row
        dt1 = torch.empty_like(t1)
        for k in range(kk):
            cs = slice(k * r, (k + 1) * r)
            dyk, t1k, t2k, uk = dyt[k * m:(k + 1) * m], t1[:, cs], t2[:, cs], u[k]

            def w_up(dyk=dyk, t2k=t2k, k=k):
                ops.gemm(dyk, t2k, a_layout=MN, b_layout=MN, out=g.out(src.up_A, k), out_dtype=F32)
                ops.colsum(dyk, out=g.out(src.up_bias, k))
            if train_src:
                sb.run(w_up, dyk, t2k)
            dt2 = ops.gemm(dyk, self._bf16(src.up_A)[k], b_layout=MN)
            if train_src:
                sb.run(lambda dt2=dt2, uk=uk, k=k: ops.gemm(dt2, uk, a_layout=MN, b_layout=MN, out=g.out(src.up_B, k), out_dtype=F32), dt2, uk)
            dpre = ops.gemm(dt2, self._bf16(src.up_B)[k], b_layout=MN, epilogue=L.JL_EPI_RELU_BWD, aux=uk)

            def w_down(dpre=dpre, t1k=t1k, k=k):
                ops.colsum(dpre, out=g.out(src.down_bias, k))
                ops.gemm(dpre, t1k, a_layout=MN, b_layout=MN, out=g.out(src.down_A, k), out_dtype=F32)
            if train_src:
                sb.run(w_down, dpre, t1k)
            ops.gemm(dpre, self._bf16(src.down_A)[k], b_layout=MN, out=dt1[:, cs])
            if train_src:
                sb.run(lambda k=k, cs=cs: ops.gemm(dt1[:, cs], zs, a_layout=MN, b_layout=MN, out=g.out(src.down_B, k), out_dtype=F32), dt1, zs)
        dzs = ops.gemm(dt1, self._bf16(src.down_B).view(kk * r, d), b_layout=MN)                  # Σ_k dt1_k · B_d[k]: one GEMM, K = K·r
        if train_src:
            sb.run(lambda: ops.layernorm_wgrad(dzs, h, mean_s, rstd_s, g.out(src.norm.weight), g.out(src.norm.bias)), dzs, h, mean_s, rstd_s)
        dh, _, _ = ops.layernorm_bwd(dzs, h, src.norm.weight.detach(), mean_s, rstd_s, dres=dh)
        return dh

    def _wf_bwd_rows(self, ad, k: int, rows: slice, dy, z, t1, u, t2, dz, g: "GradSink", sb: "_SideBranch", jobs: Optional[list] = None,
                     lp: Optional[dict] = None) -> None:
        """Backward of the WFAdapter projections for the utterances of dialect ``k`` (a row slice): weight gradients of
        factor set k on the side branch, dz[rows] = gradient at the adapter's LayerNorm output.  With ``jobs`` the bias-gradient
        column sums are appended to it (for one merged launch by the caller) instead of being launched here."""
        MN = L.JL_LAYOUT_MN
        dy, z, t1, u, t2 = dy[rows], z[rows], t1[rows], u[rows], t2[rows]

        def w_up():
            ops.gemm(dy, t2, a_layout=MN, b_layout=MN, out=g.out(ad.up_A, k), out_dtype=F32)             # dyᵀ · t2
            if jobs is None and lp is None:
                ops.colsum(dy, out=g.out(ad.up_bias, k))
        sb.run(w_up, dy, t2)
        if jobs is not None and lp is None:
            jobs.append(dict(dy=dy, out_sum=g.out(ad.up_bias, k)))
        dt2 = ops.gemm(dy, self._bf16(ad.up_A)[k], b_layout=MN)                                           # dy · A_u
        sb.run(lambda: ops.gemm(dt2, u, a_layout=MN, b_layout=MN, out=g.out(ad.up_B, k), out_dtype=F32), dt2, u)   # dt2ᵀ · u
        dpre = ops.gemm(dt2, self._bf16(ad.up_B)[k], b_layout=MN, epilogue=L.JL_EPI_RELU_BWD, aux=u)      # (dt2 · B_u) ∘ relu'

        def w_down():
            if jobs is None:
                ops.colsum(dpre, out=g.out(ad.down_bias, k))
            ops.gemm(dpre, t1, a_layout=MN, b_layout=MN, out=g.out(ad.down_A, k), out_dtype=F32)          # dpreᵀ · t1
        sb.run(w_down, dpre, t1)
        if jobs is not None:
            jobs.append(dict(dy=dpre, out_sum=g.out(ad.down_bias, k)))
        if lp is not None and lp.get("dt1_all") is not None:
            dt1 = ops.gemm(dpre, self._bf16(ad.down_A)[k], b_layout=MN, out=lp["dt1_all"][rows])           # dpre · A_d, into the shared matrix
        else:
            dt1 = ops.gemm(dpre, self._bf16(ad.down_A)[k], b_layout=MN)                                   # dpre · A_d
        sb.run(lambda: ops.gemm(dt1, z, a_layout=MN, b_layout=MN, out=g.out(ad.down_B, k), out_dtype=F32), dt1, z)  # dt1ᵀ · z
        if lp is None:
            ops.gemm(dt1, self._bf16(ad.down_B)[k], b_layout=MN, out=dz[rows])                            # dt1 · B_d
            return
        if lp.get("dt1_all") is not None:
            lp["runs"].append((rows.start, rows.stop, k))       # the tail of all runs is ONE jl_lnproj_bwd launch, issued by the caller
            return
        # dt1 · B_d and the LayerNorm backward of these rows in one kernel (jl_lnproj_bwd); its column sums give this factor set's
        # up-projection bias gradient (Σ dy) and this row range's share of the adapter LayerNorm's dγ / dβ
        _, _, cols = ops.lnproj_bwd(dt1, t1, self._bf16(ad.down_B)[k], lp["packs"]["k"][k], ad.norm.weight.detach(), lp["h"][rows], lp["mean"][rows],
                                    lp["rstd"][rows], dy, want_cols=True, out=lp["dh"][rows])
        acc = lp["seen"] > 0
        lp["seen"] += 1
        sb.run(lambda: ops.lnproj_bwd_reduce(cols, g.out(ad.norm.weight), g.out(ad.norm.bias), g.out(ad.up_bias, k), accumulate=acc), cols)

    def _adapter_bwd(self, ad: nn.Module, saved, dy: torch.Tensor, lengths, b: int, t: int, g: "GradSink", sb: "_SideBranch",
                     pk: Optional[PackedLayout] = None) -> torch.Tensor:
        """dy = grad of the adapter output → returns grad of the adapter input; weight grads go to ``g`` (issued on the
        side branch ``sb``)."""
        MN = L.JL_LAYOUT_MN
        cu = None if pk is None else pk.cu
        if ad.kind == "fuse":
            return self._fusion_bwd(ad, saved, dy, g, sb)
        if ad.kind == "wf":
            h, mean, rstd, z, t1, u, t2, segs = saved
            if z is None:
                # the fused forward kernel never wrote LN(h); only dB_d = dt1ᵀ · LN(h) needs it: recompute it on the weight-gradient
                # branch (the buffer is allocated here, on the main stream; the branch only writes into it)
                z = torch.empty_like(h)
                sb.run(lambda z=z: ops.layernorm_fwd(h, ad.norm.weight.detach(), ad.norm.bias.detach(), ad.norm.eps, out=z), h, z)
            present = set()
            jobs = [] if (_MERGED_REDUCE and _LN_WGRAD != "main") else None
            lp = None
            if self.fused_wf_bwd and _LN_WGRAD != "main" and ad.hidden_size % 64 == 0 and ad.hidden_size <= 1024 and ad.rank % 8 == 0 and ad.rank <= 192:
                lp = dict(h=h, mean=mean, rstd=rstd, seen=0)
                spans = [self._seg_rows(b0, b1, t, pk) for _, b0, b1 in segs]
                spans = [r_ for r_ in spans if r_.stop > r_.start]
                # one kernel per dialect run: with several runs per layer (the mixed-dialect batch: 4) the extra launches on the main chain
                # cost more than the side branch saves (14.81 vs 14.73 ms), with one run it pays (24-layer config 18.62 vs 18.96 ms)
                if sum(r_.stop - r_.start for r_ in spans) != h.shape[0] or len(spans) == 0 or len(spans) > 8:
                    lp = None
                elif len(spans) > 1:
                    if self.wf_multi_run:
                        # several dialect runs: their tails run as ONE launch over a shared dt1 matrix (row runs with their own factor set)
                        lp["dt1_all"], lp["runs"] = torch.empty((h.shape[0], ad.rank), dtype=BF16, device=h.device), []
                    else:
                        lp = None
            if lp is not None:
                lp["packs"], lp["dh"] = self._wf_fold_packs(ad), torch.empty_like(h)
            dz = torch.empty_like(h) if lp is None else None
            for k, b0, b1 in segs:
                rows = self._seg_rows(b0, b1, t, pk)
                if rows.stop == rows.start:
                    continue
                present.add(k)
                self._wf_bwd_rows(ad, k, rows, dy, z, t1, u, t2, dz, g, sb, jobs, lp)
            for k in range(ad.num_dialects):           # factor sets without utterances in this batch: zero gradient
                if k not in present:
                    for prm in (ad.up_A, ad.up_bias, ad.up_B, ad.down_A, ad.down_bias, ad.down_B):
                        g.out(prm, k).zero_()
            if lp is not None and lp.get("dt1_all") is not None:
                w_all = self._bf16(ad.down_B).view(ad.num_dialects * ad.rank, ad.hidden_size)
                _, _, cols = ops.lnproj_bwd(lp["dt1_all"], t1, w_all, lp["packs"]["all"], ad.norm.weight.detach(), h, mean, rstd, dy, want_cols=True,
                                            out=lp["dh"], runs=lp["runs"])
                runs = lp["runs"]

                def reduce_runs(cols=cols, runs=runs):
                    ops.lnproj_bwd_reduce(cols, g.out(ad.norm.weight), g.out(ad.norm.bias), None)
                    off = 0
                    for r0_, r1_, k_ in runs:
                        nt = (r1_ - r0_ + 127) // 128
                        ops.lnproj_bwd_reduce(cols, None, None, g.out(ad.up_bias, k_), tile_offset=off, num_tiles=nt)
                        off += nt
                sb.run(reduce_runs, cols)
                lp["seen"] = len(runs)
            if lp is not None:
                if lp["seen"] == 0:                # no utterance at all: the LayerNorm gradients are zero too
                    g.out(ad.norm.weight).zero_()
                    g.out(ad.norm.bias).zero_()
                    lp["dh"].copy_(dy)
                if jobs:
                    def reduce_wf(jobs=jobs):
                        for i in range(0, len(jobs), 4):
                            ops.colreduce_multi(jobs[i:i + 4])
                    sb.run(reduce_wf, *[j["dy"] for j in jobs])
                return lp["dh"]
        else:
            h, mean, rstd, z, qkv, a, lse = saved
            fused_fwd = z is None
            # tail of the backward as one kernel (jl_lnproj_bwd): it also leaves what the weight-gradient branch needs for dγ, dβ, db_o
            # and (without LN(h)) for dW_qkv, db_qkv
            use_lp = self.fused_att_bwd and _LN_WGRAD != "main" and ad.hidden_size % 64 == 0 and ad.hidden_size <= 1024
            if z is None and not use_lp:       # (with use_lp the variants below recompute it only if they need it)
                # the fused forward kernel never wrote LN(h); only dW_qkv = dqkvᵀ · LN(h) needs it: recomputed on the weight-gradient
                # branch into a buffer allocated here, on the main stream
                z = torch.empty_like(h)
                sb.run(lambda z=z: ops.layernorm_fwd(h, ad.norm.weight.detach(), ad.norm.bias.detach(), ad.norm.eps, out=z), h, z)
            jobs = [] if (_MERGED_REDUCE and _LN_WGRAD != "main") else None

            def w_o():
                ops.gemm(dy, a, a_layout=MN, b_layout=MN, out=g.out(ad.o_proj.weight), out_dtype=F32)         # dyᵀ · a
                if jobs is None and not use_lp:
                    ops.colsum(dy, out=g.out(ad.o_proj.bias))
            sb.run(w_o, dy, a)
            if jobs is not None and not use_lp:
                jobs.append(dict(dy=dy, out_sum=g.out(ad.o_proj.bias)))
            da = ops.gemm(dy, self._bf16(ad.o_proj.weight), b_layout=MN)                                      # dy · W_o
            dqkv = ops.attn_bwd(qkv[:, 0:64], qkv[:, 64:128], qkv[:, 128:192], a, da, lse, lengths, b, t, 1, 1.0 / 8.0, cu_seqlens=cu)
            ws = [ad.q_proj.weight, ad.k_proj.weight, ad.v_proj.weight]
            bs = [ad.q_proj.bias, ad.k_proj.bias, ad.v_proj.bias]

            gb_cat = g.out_cat(bs)
            if use_lp and self.lp_wgrad == 2:
                # dW_qkv without LN(h), its operands derived on the weight-gradient branch (jl_lnproj_wgrad_prep): nothing added to the main chain
                dh, _, cols = ops.lnproj_bwd(dqkv, qkv, self._cat_bf16(ws), self._att_pack_dev(ad, True, reuse=fused_fwd), ad.norm.weight.detach(),
                                             h, mean, rstd, dy, want_cols=True)
                sb.run(lambda: ops.lnproj_bwd_reduce(cols, g.out(ad.norm.weight), g.out(ad.norm.bias), g.out(ad.o_proj.bias)), cols)

                def w_qkv_prep():
                    dys, wpart = ops.lnproj_wgrad_prep(dqkv, mean, rstd)
                    gw = g.out_cat(ws)
                    ops.gemm(dys, h, a_layout=MN, b_layout=MN, out=gw, out_dtype=F32)                         # (dqkv ⊙ rstd)ᵀ · h
                    ops.lnproj_wgrad(gw, wpart, ad.norm.weight.detach(), ad.norm.bias.detach(), gb_cat)
                    g.scatter_cat(ws, gw)
                    g.scatter_cat(bs, gb_cat)
                sb.run(w_qkv_prep, dqkv, h, mean, rstd)
                return dh
            if use_lp and not self.lp_wgrad:
                # (variant kept for A/B: dW_qkv from LN(h) recomputed on the weight-gradient branch, as the two-kernel path does)
                if z is None:
                    z = torch.empty_like(h)
                    sb.run(lambda z=z: ops.layernorm_fwd(h, ad.norm.weight.detach(), ad.norm.bias.detach(), ad.norm.eps, out=z), h, z)
                dh, _, cols = ops.lnproj_bwd(dqkv, qkv, self._cat_bf16(ws), self._att_pack_dev(ad, True, reuse=fused_fwd), ad.norm.weight.detach(),
                                             h, mean, rstd, dy, want_cols=True, col_split=self.lp_col_split)
                sb.run(lambda: ops.lnproj_bwd_reduce(cols, g.out(ad.norm.weight), g.out(ad.norm.bias), g.out(ad.o_proj.bias)), cols)

                def w_qkv_z():
                    gw = g.out_cat(ws)
                    ops.gemm(dqkv, z, a_layout=MN, b_layout=MN, out=gw, out_dtype=F32)                        # dqkvᵀ · z
                    g.scatter_cat(ws, gw)
                    ops.colsum(dqkv, out=gb_cat)
                    g.scatter_cat(bs, gb_cat)
                sb.run(w_qkv_z, dqkv, z)
                return dh
            if use_lp:
                # dqkv · W_qkv and the LayerNorm backward in one kernel (its row means come from dqkv and the saved q|k|v); no dz, no
                # LN(h): the kernel leaves per-row-tile column sums and dqkv ⊙ rstd, from which the weight-gradient branch gets dγ, dβ,
                # db_o (jl_lnproj_bwd_reduce) and dW_qkv = ((dqkv ⊙ rstd)ᵀ h − v 1ᵀ) ⊙ γ + cs βᵀ, db_qkv = cs (GEMM + jl_lnproj_wgrad)
                dh, _, cols, dys, wpart = ops.lnproj_bwd(dqkv, qkv, self._cat_bf16(ws), self._att_pack_dev(ad, True, reuse=fused_fwd),
                                                         ad.norm.weight.detach(), h, mean, rstd, dy, want_cols=True, want_wgrad_operands=True)
                sb.run(lambda: ops.lnproj_bwd_reduce(cols, g.out(ad.norm.weight), g.out(ad.norm.bias), g.out(ad.o_proj.bias)), cols)

                def w_qkv_lp():
                    gw = g.out_cat(ws)
                    ops.gemm(dys, h, a_layout=MN, b_layout=MN, out=gw, out_dtype=F32)                         # (dqkv ⊙ rstd)ᵀ · h
                    ops.lnproj_wgrad(gw, wpart, ad.norm.weight.detach(), ad.norm.bias.detach(), gb_cat)
                    g.scatter_cat(ws, gw)
                    g.scatter_cat(bs, gb_cat)
                sb.run(w_qkv_lp, dys, h, wpart)
                if jobs:
                    def reduce_all(jobs=jobs):
                        for i in range(0, len(jobs), 4):
                            ops.colreduce_multi(jobs[i:i + 4])
                        for j in jobs:
                            if "scatter" in j:
                                g.scatter_cat(*j["scatter"])
                    sb.run(reduce_all, *[j["dy"] for j in jobs])
                return dh

            def w_qkv():
                gw = g.out_cat(ws)
                ops.gemm(dqkv, z, a_layout=MN, b_layout=MN, out=gw, out_dtype=F32)                            # dqkvᵀ · z
                g.scatter_cat(ws, gw)
                if jobs is None:
                    ops.colsum(dqkv, out=gb_cat)
                    g.scatter_cat(bs, gb_cat)
            sb.run(w_qkv, dqkv, z)
            if jobs is not None:
                jobs.append(dict(dy=dqkv, out_sum=gb_cat, scatter=(bs, gb_cat)))
            dz = ops.gemm(dqkv, self._cat_bf16(ws), b_layout=MN)                                              # dqkv · W_qkv
        if _LN_WGRAD == "main":
            dh, _, _ = ops.layernorm_bwd(dz, h, ad.norm.weight.detach(), mean, rstd, dres=dy, want_wgrad=True, dgamma=g.out(ad.norm.weight),
                                         dbeta=g.out(ad.norm.bias))
            return dh
        if jobs is not None:
            # one launch for every column reduction of this adapter: its bias gradients and the dγ / dβ of its LayerNorm
            jobs.append(dict(dy=dz, x=h, mean=mean, rstd=rstd, out_sum=g.out(ad.norm.bias), out_dot=g.out(ad.norm.weight)))

            def reduce_all(jobs=jobs):
                for i in range(0, len(jobs), 4):
                    ops.colreduce_multi(jobs[i:i + 4])
                for j in jobs:
                    if "scatter" in j:
                        g.scatter_cat(*j["scatter"])
            sb.run(reduce_all, dz, h, mean, rstd, *[j["dy"] for j in jobs])
        else:
            sb.run(lambda: ops.layernorm_wgrad(dz, h, mean, rstd, g.out(ad.norm.weight), g.out(ad.norm.bias)), dz, h, mean, rstd)
        dh, _, _ = ops.layernorm_bwd(dz, h, ad.norm.weight.detach(), mean, rstd, dres=dy)
        return dh

    # ------------------------------------------------------------------ forward
    def _wav2vec2_front_end(self, wave: torch.Tensor, sample_lengths: Optional[torch.Tensor], lengths: torch.Tensor, fz: dict):
        """Raw waveform [B, N] fp32 (un-normalised; padded with anything) + valid sample counts → hidden states entering layer 0
        [B·T, d] bf16 (padded rows zero) and T.  Utterance normalisation is folded into the layer-0 im2col; every convolution is
        an im2col + tcgen05 GEMM (bias epilogue) followed by the LayerNorm+GELU kernel; the grouped positional convolution is
        one im2col + GEMM (bias + GELU + residual epilogue) per group."""
        cfg = self.cfg
        if wave.dim() != 2 or wave.dtype != F32:
            raise ValueError("wav2vec2 front end expects the waveform batch [B, N] in fp32")
        wave = wave.contiguous()
        b, n = wave.shape
        if sample_lengths is None:
            sample_lengths = torch.full((b,), n, dtype=I32, device=wave.device)
        sample_lengths = sample_lengths.to(device=wave.device, dtype=I32)
        ks, ss = cfg.conv_kernel, cfg.conv_stride
        fe = self.enc.w2v
        c = cfg.conv_dim
        stats = ops.wave_stats(wave, sample_lengths)
        t_valid = (n - ks[0]) // ss[0] + 1
        if wav2vec2_lengths(n, ks, ss) <= 0:
            raise ValueError(f"waveform batch of {n} samples is shorter than the front end's receptive field")
        # No im2col for layers 1..: k consecutive frames of a [T, C] activation are contiguous, so the GEMM's A operand is a view
        # with row stride s·C and row length k·C (rows overlap in memory).  Every utterance gets the same number of rows per
        # layer; that needs T_0 to be a multiple of the product of the later strides (the extra windows compute don't-care rows).
        prod = 1
        for s_ in ss[1:]:
            prod *= s_
        t = (t_valid + prod - 1) // prod * prod
        a = ops.wave_im2col(wave, sample_lengths, stats, t, ks[0], ss[0])
        h = None
        for i in range(len(ks)):
            if i > 0:
                t = t // ss[i]
                a = torch.as_strided(h, (b * t, ks[i] * c), (ss[i] * c, 1))
            buf = torch.empty((b * t + 8, c), dtype=BF16, device=wave.device)     # slack rows: the last windows of the next layer
            buf[b * t:].zero_()
            y = ops.gemm(a, fz[f"w2v.conv{i}.w"], bias=fz[f"w2v.conv{i}.b"], out=buf[: b * t])
            h, _, _ = ops.layernorm_fwd(y, fe.conv_norm[i].weight.detach(), fe.conv_norm[i].bias.detach(), fe.conv_norm[i].eps, gelu=True, out=y)
        z, _, _ = ops.layernorm_fwd(h, fe.proj_norm.weight.detach(), fe.proj_norm.bias.detach(), fe.proj_norm.eps, out=h)
        hp = ops.gemm(z, fz["w2v.proj.w"], bias=fz["w2v.proj.b"], row_lengths=lengths, rows_per_seq=t)      # padded frames := 0
        d, groups, kp = cfg.hidden_size, cfg.num_conv_pos_embedding_groups, cfg.num_conv_pos_embeddings
        cg = d // groups
        out = torch.empty_like(hp)
        col = torch.empty((b * t, kp * cg), dtype=BF16, device=hp.device)
        hp3 = hp.view(b, t, d)
        for gi in range(groups):
            ops.im2col_1d(hp3, t, kp, 1, pad=kp // 2, c0=gi * cg, cg=cg, out=col)
            ops.gemm(col, fz[f"w2v.pos{gi}.w"], bias=fz["w2v.pos.b"][gi * cg:(gi + 1) * cg], epilogue=L.JL_EPI_GELU,
                     residual=hp[:, gi * cg:(gi + 1) * cg], out=out[:, gi * cg:(gi + 1) * cg], row_lengths=lengths, rows_per_seq=t)
        return out, t

    def forward(self, input_features: torch.Tensor, lengths: torch.Tensor, training: bool = False, dialect: int = 0,
                want_logits: bool = True, sample_lengths: Optional[torch.Tensor] = None, packed: Optional[PackedLayout] = None) -> _State:
        """``packed``: run the encoder in the packed row layout (mel front end only): the conv subsampler still sees the padded
        [B, F, 80] features, its output is packed by the position-embedding kernel, and from there on every matrix is
        [total valid frames, C].  ``st.logits`` / ``st.h_final`` are then packed too."""
        cfg = self.cfg
        d, heads = cfg.hidden_size, cfg.num_attention_heads
        fz = self._frozen_pack()
        if not input_features.is_cuda:
            raise RuntimeError("JLEngine.forward: input_features must be a CUDA tensor (no CPU fallback)")
        st = _State()
        st.training, st.dialect = training, dialect
        st.lengths = lengths
        st.packed = pk = packed
        cu = None if pk is None else pk.cu
        if pk is not None and cfg.front_end == "wav2vec2":
            raise NotImplementedError("the packed layout is implemented for the mel front end")
        if cfg.front_end == "wav2vec2":
            b = input_features.shape[0]
            h, t = self._wav2vec2_front_end(input_features, sample_lengths, lengths, fz)
        else:
            b, f, nmel = input_features.shape
            if nmel != cfg.input_feat_per_channel:
                raise ValueError(f"expected {cfg.input_feat_per_channel} mel bins, got {nmel}")
            x16 = input_features if input_features.dtype == BF16 else ops.cast_bf16(input_features.contiguous())
            x16 = x16.contiguous()
            # conv subsampler as two im2col GEMMs with fused bias + GLU
            a1, t1 = ops.im2col_k5s2(x16)
            c1 = ops.gemm(a1, fz["conv0.w"], bias=fz["conv0.b"], epilogue=L.JL_EPI_GLU)
            a2, t = ops.im2col_k5s2(c1.view(b, t1, cfg.conv_channels // 2))
            h = ops.gemm(a2, fz["conv1.w"], bias=fz["conv1.b"], epilogue=L.JL_EPI_GLU)
            if pk is None:
                ops.embed_positions_(h, math.sqrt(d), self.pos_table(h.device, t + 2), lengths, b, t)
            else:
                if pk.batch != b or max(pk.lens + [0]) > t:
                    raise ValueError("packed layout does not match the feature batch")
                h = ops.embed_positions_packed(h, math.sqrt(d), self.pos_table(h.device, t + 2), cu, b, t, pk.total)
                t = pk.seq_bound                      # from here on: upper bound on an utterance's length (attention grid)
        st.b, st.t = b, t
        scale = 1.0 / 8.0   # head_dim 64
        st.layers = []
        self._att_packed_step = False
        if training:
            self._att_pack_all()
        for i, layer in enumerate(self.enc.layers):
            sv = _State()
            sv.h_in = h
            x1, sv.mean1, sv.rstd1 = ops.layernorm_fwd(h, layer.layer_norm.weight.detach(), layer.layer_norm.bias.detach(),
                                                       layer.layer_norm.eps, save_stats=training)
            qkv = ops.gemm(x1, fz[f"{i}.wqkv"], bias=fz[f"{i}.bqkv"])
            o, lse = ops.attn_fwd(qkv[:, 0:d], qkv[:, d:2 * d], qkv[:, 2 * d:3 * d], lengths, b, t, heads, scale, want_lse=training,
                                  cu_seqlens=cu)
            h1 = ops.gemm(o, fz[f"{i}.wo"], bias=fz[f"{i}.bo"], residual=h)
            sv.qkv, sv.o, sv.lse = qkv, o, lse
            sv.ad_attn = None
            if layer.adapter_attn is not None:
                h1, sv.ad_attn = self._adapter_fwd(layer.adapter_attn, h1, lengths, b, t, training, dialect, zero_rows=False, pk=pk)
            sv.h1 = h1
            x2, sv.mean2, sv.rstd2 = ops.layernorm_fwd(h1, layer.final_layer_norm.weight.detach(), layer.final_layer_norm.bias.detach(),
                                                       layer.final_layer_norm.eps, save_stats=training)
            pre = torch.empty((h.shape[0], cfg.intermediate_size), dtype=BF16, device=h.device) if training else None
            # training: the epilogue leaves gelu'(pre-activation) for the backward GEMM (erf evaluated once per element)
            act = ops.gemm(x2, fz[f"{i}.w1"], bias=fz[f"{i}.b1"], epilogue=L.JL_EPI_GELU_DGELU if training else L.JL_EPI_GELU, aux_out=pre)
            sv.dgelu = pre
            last_is_ffn = layer.adapter_ffn is None
            rl = dict(row_lengths=lengths, rows_per_seq=t) if (last_is_ffn and pk is None) else {}
            h2 = ops.gemm(act, fz[f"{i}.w2"], bias=fz[f"{i}.b2"], residual=h1, **rl)
            sv.ad_ffn = None
            if layer.adapter_ffn is not None:
                h2, sv.ad_ffn = self._adapter_fwd(layer.adapter_ffn, h2, lengths, b, t, training, dialect, zero_rows=True, pk=pk)
            h = h2
            if training:
                st.layers.append(sv)
        st.h_last = h
        ln = self.enc.layer_norm
        st.h_final, st.mean_f, st.rstd_f = ops.layernorm_fwd(h, ln.weight.detach(), ln.bias.detach(), ln.eps, save_stats=training)
        st.logits = None
        st.argmax_partials = None
        if want_logits:
            if self.lm_head is None:
                raise RuntimeError("JLEngine.forward: no lm_head attached")
            if want_logits == "argmax":
                # lm_head ⊕ frame argmax — the [B·T', V] logits (5 MB fp32 per 10 s utterance) are never written
                st.argmax_partials = ops.lm_head_argmax(st.h_final, self._bf16(self.lm_head.weight), self.lm_head.bias.detach())
            else:
                out_dtype = F32 if cfg.logits_dtype == "float32" else BF16
                st.logits = ops.gemm(st.h_final, self._bf16(self.lm_head.weight), bias=self.lm_head.bias.detach(), out_dtype=out_dtype)
        return st

    # ------------------------------------------------------------------ backward (adapter-only)
    def lowest_adapter_layer(self) -> int:
        for i, layer in enumerate(self.enc.layers):
            if layer.adapter_attn is not None or layer.adapter_ffn is not None:
                return i
        return len(self.enc.layers)

    def backward(self, st: _State, dlogits: torch.Tensor, g: "GradSink", on_progress=None) -> None:
        """dlogits [B*T', V] bf16 (d loss / d logits, zero on padded rows) → adapter + lm_head gradients into ``g``.
        ``on_progress(i, side_stream)`` is called when every weight gradient of lm_head and of the layers above ``i`` has been
        issued (layers are visited top-down), so a trainer can start exchanging that part of the bucket while the backward of the
        lower layers is still running."""
        for n, p in self._backbone_params():
            if p.requires_grad:
                raise NotImplementedError(
                    f"backbone parameter {n} requires grad: this path implements the adapter-only backward of the reference "
                    "(call freeze_base_model())")
        MN = L.JL_LAYOUT_MN
        cfg = self.cfg
        d, heads = cfg.hidden_size, cfg.num_attention_heads
        b, t = st.b, st.t
        ft = self._frozen_pack_t()
        lengths = st.lengths
        pk = getattr(st, "packed", None)
        cu = None if pk is None else pk.cu
        sb = _SideBranch(enabled=self.side_branch, stream=self._side_stream(dlogits.device) if self.side_branch else None,
                         defer=self.defer_wgrads, pool=self._side_pool(dlogits.device) if (self.side_branch and self.defer_wgrads) else None)
        g.prepare(self)      # sinks that allocate do so here, on the main stream (the side branch only writes into them)

        # head: logits = h_final · Wᵀ + b
        def w_head():
            ops.gemm(dlogits, st.h_final, a_layout=MN, b_layout=MN, out=g.out(self.lm_head.weight), out_dtype=F32)   # dlogitsᵀ · h_final
            ops.colsum(dlogits, out=g.out(self.lm_head.bias))
        sb.run(w_head, dlogits, st.h_final, now=True)
        l0 = self.lowest_adapter_layer()
        if l0 >= len(self.enc.layers):
            sb.join()
            return
        dhf = ops.gemm(dlogits, self._bf16(self.lm_head.weight), b_layout=MN)                                   # dlogits · W
        ln = self.enc.layer_norm
        dh, _, _ = ops.layernorm_bwd(dhf, st.h_last, ln.weight.detach(), st.mean_f, st.rstd_f)
        for i in range(len(self.enc.layers) - 1, l0 - 1, -1):
            layer, sv = self.enc.layers[i], st.layers[i]
            if on_progress is not None:
                on_progress(i, None if _DEBUG_SKIP_SIDE else sb.side)
            if layer.adapter_ffn is not None:
                dh = self._adapter_bwd(layer.adapter_ffn, sv.ad_ffn, dh, lengths, b, t, g, sb, pk=pk)
                if i == l0 and layer.adapter_attn is None:
                    break
            # FFN: h2 = h1 + W2 · gelu(W1 · LN2(h1) + b1) + b2
            dpre = ops.gemm(dh, ft[f"{i}.w2"], epilogue=L.JL_EPI_MUL_AUX, aux=sv.dgelu)         # (dh · W2) ∘ gelu'   (B = W2ᵀ stored [4d, d])
            dx2 = ops.gemm(dpre, ft[f"{i}.w1"])                                                 # dpre · W1 (B = W1ᵀ stored [d, 4d])
            fl = layer.final_layer_norm
            dh1, _, _ = ops.layernorm_bwd(dx2, sv.h1, fl.weight.detach(), sv.mean2, sv.rstd2, dres=dh)
            if layer.adapter_attn is not None:
                dh1 = self._adapter_bwd(layer.adapter_attn, sv.ad_attn, dh1, lengths, b, t, g, sb, pk=pk)
                if i == l0:
                    break
            # attention: h1 = h + Wo · attn(LN1(h) Wqkvᵀ) + bo
            d_o = ops.gemm(dh1, ft[f"{i}.wo"])
            qkv = sv.qkv
            dqkv = ops.attn_bwd(qkv[:, 0:d], qkv[:, d:2 * d], qkv[:, 2 * d:3 * d], sv.o, d_o, sv.lse, lengths, b, t, heads, 1.0 / 8.0,
                                cu_seqlens=cu)
            dx1 = ops.gemm(dqkv, ft[f"{i}.wqkv"])
            l1 = layer.layer_norm
            dh, _, _ = ops.layernorm_bwd(dx1, sv.h_in, l1.weight.detach(), sv.mean1, sv.rstd1, dres=dh1)
        sb.join()


class GradSink:
    """Where the backward pass writes weight gradients: fp32 tensors keyed by parameter.  The default sink allocates a
    tensor per parameter; ``training.FlatAdapterParams`` hands out views of one contiguous bucket instead."""

    def __init__(self):
        self.grads: Dict[int, torch.Tensor] = {}
        self.params: Dict[int, torch.Tensor] = {}

    def prepare(self, engine: "JLEngine") -> None:
        """Allocate (zeroed) gradient tensors for every trainable parameter on the CURRENT stream, before the backward pass
        starts issuing weight-gradient products on its side stream: the caching allocator then never sees a block that was
        allocated on one stream and freed on another."""
        ps = [p for p in engine.enc.parameters() if p.requires_grad]
        if engine.lm_head is not None:
            ps += [p for p in engine.lm_head.parameters() if p.requires_grad]
        for p in ps:
            self._full(p)

    def _full(self, p: torch.Tensor) -> torch.Tensor:
        gt = self.grads.get(id(p))
        if gt is None:
            gt = torch.zeros(p.shape, dtype=F32, device=p.device)
            self.grads[id(p)] = gt
            self.params[id(p)] = p
        return gt

    def out(self, p: torch.Tensor, k: Optional[int] = None) -> torch.Tensor:
        gt = self._full(p)
        return gt if k is None else gt[k]

    def out_cat(self, ps: List[torch.Tensor]) -> torch.Tensor:
        rows = sum(p.shape[0] for p in ps)
        return torch.empty((rows,) + tuple(ps[0].shape[1:]), dtype=F32, device=ps[0].device)

    def scatter_cat(self, ps: List[torch.Tensor], cat: torch.Tensor) -> None:
        off = 0
        for p in ps:
            self._full(p).copy_(cat[off: off + p.shape[0]])
            off += p.shape[0]


# =============================================================================================== CTC model
class _CTCStep(torch.autograd.Function):
    """Autograd bridge: forward ran in the engine; backward runs the engine's adapter-only backward and returns the
    adapter / lm_head gradients so that ``loss.backward()`` + any torch optimizer work unchanged."""

    @staticmethod
    def forward(ctx, model, st, dlogits, loss, *params):
        ctx.model, ctx.st, ctx.dlogits, ctx.n = model, st, dlogits, len(params)
        ctx.params = params
        return loss.clone()

    @staticmethod
    def backward(ctx, gloss):
        sink = GradSink()
        eng = ctx.model.encoder.engine(ctx.model.lm_head)
        b, t = ctx.st.b, ctx.st.t
        eng.backward(ctx.st, ctx.dlogits.view(-1, ctx.dlogits.shape[-1]), sink)
        grads = []
        for p in ctx.params:
            gt = sink.grads.get(id(p))
            grads.append(None if gt is None else gt * gloss)
        return (None, None, None, None, *grads)


class JLForCTC(nn.Module):
    """Encoder + ``lm_head`` + CTC loss, HF ``Wav2Vec2ForCTC``-style interface."""

    def __init__(self, config: JLConfig):
        super().__init__()
        self.config = config
        self.encoder = JLEncoder(config)
        self.lm_head = nn.Linear(config.hidden_size, config.vocab_size)
        self.init_weights(seed=0)

    # ---- init (modeling_wav2vec2.py:990-1003: Linear N(0, 0.02) / bias 0, LayerNorm 1 / 0, Conv1d kaiming-normal)
    def init_weights(self, seed: int = 0) -> None:
        gen = torch.Generator().manual_seed(seed)
        std = self.config.initializer_range
        with torch.no_grad():
            for m in self.modules():
                if isinstance(m, nn.Linear):
                    m.weight.copy_(torch.randn(m.weight.shape, generator=gen) * std)
                    m.bias.zero_()
                elif isinstance(m, nn.LayerNorm):
                    m.weight.fill_(1.0)
                    m.bias.zero_()
                elif isinstance(m, nn.Conv1d):
                    fan_in = m.in_channels * m.kernel_size[0]
                    m.weight.copy_(torch.randn(m.weight.shape, generator=gen) * math.sqrt(2.0 / fan_in))
                    bound = math.sqrt(1.0 / fan_in)
                    m.bias.copy_((torch.rand(m.bias.shape, generator=gen) * 2 - 1) * bound)
                elif isinstance(m, (WFAdapter, AttAdapter)):
                    m.reset_parameters(std, gen)
                elif isinstance(m, JLWav2Vec2FrontEnd):
                    m.reset_pos_conv(gen)

    # ---- adapter management
    def freeze_base_model(self) -> None:
        """Frozen backbone, trainable adapters + lm_head (modeling_wav2vec2.py:1666-1672)."""
        for p in self.parameters():
            p.requires_grad = False
        for p in self._get_adapters().values():
            p.requires_grad = True

    def _get_adapters(self) -> Dict[str, nn.Parameter]:
        """name → Parameter for every adapter layer plus lm_head (modeling_wav2vec2.py:1046-1060)."""
        out = {}
        for n, p in self.named_parameters():
            if ".adapter_attn." in n or ".adapter_ffn." in n or n.startswith("lm_head."):
                out[n] = p
        return out

    def init_adapter_layers(self, seed: Optional[int] = None) -> None:
        """Re-initialise adapters and lm_head (modeling_wav2vec2.py:1062-1073)."""
        gen = torch.Generator().manual_seed(seed) if seed is not None else None
        std = self.config.initializer_range
        with torch.no_grad():
            for m in self.modules():
                if isinstance(m, (WFAdapter, AttAdapter, FusionAdapter)):
                    m.reset_parameters(std, gen)
            self.lm_head.weight.copy_((torch.randn(self.lm_head.weight.shape, generator=gen) * std).to(self.lm_head.weight.device))
            self.lm_head.bias.zero_()

    def save_adapter(self, path: str) -> None:
        """Write the adapter + lm_head tensors: ``*.safetensors`` (HF's ``adapter.<lang>.safetensors`` format,
        modeling_wav2vec2.py:1152-1168) or a ``torch.save`` file for any other suffix."""
        hf_compat.write_tensor_file({n: p for n, p in self._get_adapters().items()}, path)

    def load_adapter(self, path: str, strict: bool = True, model_dir: Optional[str] = None, dialect: int = 0) -> None:
        """Swap a per-dialect adapter in place (local files only).  ``path`` is a file, or — with ``model_dir`` — a
        language/dialect name resolved to ``adapter.<name>.safetensors`` / ``adapter.<name>.bin`` in that directory
        (modeling_wav2vec2.py:1152-1225).  Files with HF's bottleneck-adapter names (``…adapter_layer.{norm,linear_1,
        linear_2}``) load into the WFAdapter of the ``adapter_ffn`` slot (factor set ``dialect``).  lm_head is resized to
        the file's vocabulary like modeling_wav2vec2.py:1230-1244."""
        if model_dir is not None:
            path = hf_compat.adapter_file(model_dir, path)
        if not os.path.isfile(path):
            raise EnvironmentError(f"adapter file {path} not found (local files only)")
        sd = hf_compat.read_tensor_file(path)
        hf_named = any(".adapter_layer." in k for k in sd)
        new_vocab = sd["lm_head.weight"].shape[0] if "lm_head.weight" in sd else self.config.vocab_size
        if new_vocab != self.config.vocab_size:
            eng = self.encoder._engine
            if eng is not None and eng.flat is not None:
                raise RuntimeError(
                    f"load_adapter: the file's vocabulary ({new_vocab}) differs from the model's ({self.config.vocab_size}) and an "
                    "AdapterTrainer is attached: its flat parameter bucket was laid out for the current lm_head.  Load the adapter "
                    "before building the trainer (or build a new trainer afterwards).")
            dev = self.lm_head.weight.device
            self.lm_head = nn.Linear(self.config.hidden_size, new_vocab).to(dev)
            self.config.vocab_size = new_vocab
        if hf_named:
            missing, skipped = hf_compat.load_hf_state_dict(self, sd, strict=False, dialect=dialect)
            if strict and skipped:
                raise ValueError(f"adapter weights do not match: unexpected {sorted(skipped)}")
            return
        mine = self._get_adapters()
        unexpected = set(sd) - set(mine)
        missing = set(mine) - set(sd)
        if strict and (unexpected or missing):
            raise ValueError(f"adapter weights do not match: unexpected {sorted(unexpected)}, missing {sorted(missing)}")
        with torch.no_grad():
            for n, v in sd.items():
                if n in mine:
                    mine[n].copy_(v.to(mine[n].device, mine[n].dtype))

    def load_hf_state_dict(self, sd, strict: bool = False, dialect: int = 0, hf_config=None):
        """Load a ``Wav2Vec2ForCTC`` / ``Speech2Text`` encoder state dict by its HF names (see ``hf_compat``).
        Returns (missing model keys, skipped checkpoint keys).  Post-LN / group-norm wav2vec2-base checkpoints are refused."""
        return hf_compat.load_hf_state_dict(self, sd, strict=strict, dialect=dialect, hf_config=hf_config)

    # ---- forward
    def forward(self, input_features: Optional[torch.Tensor] = None, attention_mask: Optional[torch.Tensor] = None,
                labels: Optional[torch.Tensor] = None, frame_lengths: Optional[torch.Tensor] = None, dialect=0,
                input_values: Optional[torch.Tensor] = None, packed: bool = False):
        """→ (loss | None, logits [B, T', V]).  ``labels`` [B, S] padded with -100 (any negative value).  ``dialect``: the
        WFAdapter factor set — one id, or one id per utterance (utterances of one dialect adjacent).  With
        ``front_end="wav2vec2"`` the input is the waveform batch [B, N] fp32 (``input_values``, HF's name; raw or already
        normalised — the utterance normalisation is idempotent), ``attention_mask`` its sample mask.
        ``packed=True`` (mel front end): the encoder runs on the packed row layout (no work on padded frames — mixed-length
        batches); the returned logits are then [total valid frames, V] with utterance b at rows ``self.last_packed.cu_host[b]:
        cu_host[b+1]`` (``unpack_logits`` restores [B, T', V]).  The lengths are read back to the host once to lay the rows out."""
        cfg = self.config
        if input_features is None:
            input_features = input_values
        if input_features is None:
            raise ValueError("forward needs input_features (mel features) or input_values (waveforms)")
        eng = self.encoder.engine(self.lm_head)
        lengths = eng.output_lengths(input_features, attention_mask, frame_lengths)
        train = labels is not None and torch.is_grad_enabled() and any(p.requires_grad for p in self._get_adapters().values())
        pk = PackedLayout(lengths.tolist(), lengths.device) if packed else None
        self.last_packed = pk
        st = eng.forward(input_features, lengths, training=train, dialect=dialect, want_logits=True,
                         sample_lengths=_sample_lengths(attention_mask, frame_lengths), packed=pk)
        b, t = st.b, st.t
        logits = st.logits if pk is not None else st.logits.view(b, t, cfg.vocab_size)
        if labels is None:
            return None, logits
        if int(labels.max()) >= cfg.vocab_size:
            raise ValueError(f"Label values must be <= vocab_size: {cfg.vocab_size}")
        lab = labels.to(device=logits.device, dtype=I32)
        loss, nll, grad = ops.ctc_loss(logits, lab, lengths, blank=cfg.pad_token_id, reduction=cfg.ctc_loss_reduction,
                                       zero_infinity=cfg.ctc_zero_infinity, want_grad=train, grad_dtype=BF16,
                                       cu_seqlens=None if pk is None else pk.cu, max_len=0 if pk is None else pk.seq_bound)
        loss = loss.view(())
        if train:
            params = [p for p in self._get_adapters().values() if p.requires_grad]
            loss = _CTCStep.apply(self, st, grad, loss, *params)
        return loss, logits

    def output_lengths(self, input_features, attention_mask=None, frame_lengths=None) -> torch.Tensor:
        return self.encoder.engine(self.lm_head).output_lengths(input_features, attention_mask, frame_lengths)

    @staticmethod
    def unpack_logits(logits: torch.Tensor, pk: PackedLayout) -> torch.Tensor:
        """[total, V] packed logits → [B, max T', V] (padded frames zero): a row gather, no arithmetic."""
        tmax = max(pk.lens + [1])
        out = logits.new_zeros((pk.batch, tmax, logits.shape[-1]))
        for b, n in enumerate(pk.lens):
            out[b, :n] = logits[pk.cu_host[b]: pk.cu_host[b + 1]]
        return out

    @torch.no_grad()
    def greedy_decode(self, logits: torch.Tensor, lengths: torch.Tensor, packed: Optional[PackedLayout] = None) -> List[List[int]]:
        """argmax → collapse repeats → strip blank (tokenization_wav2vec2.py:310-317) on the GPU; only the compacted
        ids cross to the host.  ``packed``: the layout of [total, V] logits produced with ``forward(packed=True)``."""
        if packed is not None:
            ids, n, _ = ops.ctc_greedy(logits, lengths.to(I32), blank=self.config.pad_token_id, cu_seqlens=packed.cu, max_len=packed.seq_bound)
            ids, n = ids.cpu(), n.cpu()
            return [ids[i, : int(n[i])].tolist() for i in range(ids.shape[0])]
        ids, n, _ = ops.ctc_greedy(logits, lengths.to(I32), blank=self.config.pad_token_id)
        ids, n = ids.cpu(), n.cpu()
        return [ids[i, : int(n[i])].tolist() for i in range(ids.shape[0])]

    @torch.no_grad()
    def transcribe(self, input_features, attention_mask=None, frame_lengths=None, dialect=0, fused_head: bool = True) -> List[List[int]]:
        """Features → token ids.  ``fused_head`` (default): the lm_head GEMM's epilogue emits per-chunk (max, argmax) pairs instead
        of logits and the greedy decoder reads those; the ids are identical to decoding the materialised logits."""
        if not fused_head:
            _, logits = self.forward(input_features, attention_mask, None, frame_lengths, dialect=dialect)
            return self.greedy_decode(logits, self.output_lengths(input_features, attention_mask, frame_lengths))
        eng = self.encoder.engine(self.lm_head)
        lengths = eng.output_lengths(input_features, attention_mask, frame_lengths)
        st = eng.forward(input_features, lengths, training=False, dialect=dialect, want_logits="argmax",
                         sample_lengths=_sample_lengths(attention_mask, frame_lengths))
        pmax, pidx = st.argmax_partials
        ids, n, _ = ops.ctc_greedy_from_partials(pmax, pidx, lengths, st.b, st.t, blank=self.config.pad_token_id)
        ids, n = ids.cpu(), n.cpu()
        return [ids[i, : int(n[i])].tolist() for i in range(ids.shape[0])]
