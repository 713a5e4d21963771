"""Oracle end-to-end path: waveform → features → encoder (+adapters) → logits →
CTC loss / greedy IDs, plus HF-style weight init.  fp32 CPU.  Test
infrastructure only (see ``oracle/__init__.py``).

Config field names follow ``SP/transformers/models/wav2vec2/configuration_wav2vec2.py:165-219``
(hidden_size, num_hidden_layers, num_attention_heads, intermediate_size,
vocab_size, pad_token_id, ctc_loss_reduction, ctc_zero_infinity) and
``SP/transformers/models/speech_to_text/configuration_speech_to_text.py``
(conv_channels, input_feat_per_channel).  Init follows
``SP/transformers/models/wav2vec2/modeling_wav2vec2.py:990-1003`` (Linear
N(0, 0.02), bias 0; LN 1/0; Conv1d kaiming-normal + uniform bias); adapter
factors N(0, 0.02).
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence

import torch

from . import ctc as octc
from . import encoder as oenc
from . import features as ofeat


@dataclass
class OracleConfig:
    hidden_size: int = 768
    num_hidden_layers: int = 12
    num_attention_heads: int = 12
    intermediate_size: int = 3072
    conv_channels: int = 1024
    input_feat_per_channel: int = 80
    vocab_size: int = 5000
    pad_token_id: int = 0
    ctc_loss_reduction: str = "sum"
    ctc_zero_infinity: bool = False
    adapter_attn: Optional[str] = None      # None | "wf" | "att" | "fuse"
    adapter_ffn: Optional[str] = None
    wf_bottleneck: int = 256
    wf_rank: int = 32
    att_dim: int = 64
    num_dialects: int = 1
    initializer_range: float = 0.02
    front_end: str = "mel"                  # "mel" (Speech2Text conv subsampler) | "wav2vec2" (raw-waveform conv stack, §8 f3)
    conv_dim: int = 512
    conv_kernel: tuple = (10, 3, 3, 3, 3, 2, 2)
    conv_stride: tuple = (5, 2, 2, 2, 2, 2, 2)
    num_conv_pos_embeddings: int = 128
    num_conv_pos_embedding_groups: int = 16


def init_weights(cfg: OracleConfig, seed: int = 0) -> Dict[str, torch.Tensor]:
    g = torch.Generator().manual_seed(seed)
    d, std = cfg.hidden_size, cfg.initializer_range
    w: Dict[str, torch.Tensor] = {}

    def normal(*shape):
        return torch.randn(*shape, generator=g) * std

    def ln(p):
        w[p + ".weight"] = torch.ones(d)
        w[p + ".bias"] = torch.zeros(d)

    def linear(p, out_f, in_f):
        w[p + ".weight"] = normal(out_f, in_f)
        w[p + ".bias"] = torch.zeros(out_f)

    def conv(p, out_c, in_c, k):
        fan_in = in_c * k
        w[p + ".weight"] = torch.randn(out_c, in_c, k, generator=g) * math.sqrt(2.0 / fan_in)   # kaiming_normal_
        bound = math.sqrt(1.0 / fan_in)
        w[p + ".bias"] = (torch.rand(out_c, generator=g) * 2 - 1) * bound

    def adapter(p, kind):
        if kind is None:
            return
        ln(p + ".norm")
        if kind == "wf":
            k, b, r = cfg.num_dialects, cfg.wf_bottleneck, cfg.wf_rank
            w[p + ".down_B"] = normal(k, r, d)
            w[p + ".down_A"] = normal(k, b, r)
            w[p + ".down_bias"] = torch.zeros(k, b)
            w[p + ".up_B"] = normal(k, r, b)
            w[p + ".up_A"] = normal(k, d, r)
            w[p + ".up_bias"] = torch.zeros(k, d)
        elif kind == "att":
            for n in ("q_proj", "k_proj", "v_proj"):
                linear(f"{p}.{n}", cfg.att_dim, d)
            linear(p + ".o_proj", d, cfg.att_dim)
        elif kind == "fuse":
            # parameter order = the product module's (FusionAdapter: source WFAdapter, then norm / q_proj / k_proj) — ``ln`` above
            # has already created <p>.norm; move it behind the source adapter's tensors to keep one naming scheme
            nw, nb = w.pop(p + ".norm.weight"), w.pop(p + ".norm.bias")
            adapter(p + ".source", "wf")
            w[p + ".norm.weight"], w[p + ".norm.bias"] = nw, nb
            linear(p + ".q_proj", cfg.att_dim, d)
            linear(p + ".k_proj", cfg.att_dim, d)
        else:
            raise ValueError(kind)

    if cfg.front_end == "wav2vec2":
        c = cfg.conv_dim
        for i, k in enumerate(cfg.conv_kernel):
            conv(f"w2v.conv.{i}", c, 1 if i == 0 else c, k)
            w[f"w2v.conv_norm.{i}.weight"], w[f"w2v.conv_norm.{i}.bias"] = torch.ones(c), torch.zeros(c)
        w["w2v.proj_norm.weight"], w["w2v.proj_norm.bias"] = torch.ones(c), torch.zeros(c)
        linear("w2v.proj", d, c)
        kp, gp = cfg.num_conv_pos_embeddings, cfg.num_conv_pos_embedding_groups
        # modeling_wav2vec2.py:966-973: N(0, 2·sqrt(1 / (k · in_channels))) weight, zero bias; weight_norm: g = ‖v‖ per tap
        v = torch.randn(d, d // gp, kp, generator=g) * (2.0 * math.sqrt(1.0 / (kp * d)))
        w["w2v.pos_conv.weight_v"] = v
        w["w2v.pos_conv.weight_g"] = v.norm(dim=(0, 1), keepdim=True)
        w["w2v.pos_conv.bias"] = torch.zeros(d)
    else:
        conv("conv.0", cfg.conv_channels, cfg.input_feat_per_channel, 5)
        conv("conv.1", 2 * d, cfg.conv_channels // 2, 5)
    for i in range(cfg.num_hidden_layers):
        p = f"layers.{i}"
        ln(p + ".layer_norm")
        for n in ("q_proj", "k_proj", "v_proj", "out_proj"):
            linear(f"{p}.attention.{n}", d, d)
        adapter(p + ".adapter_attn", cfg.adapter_attn)
        ln(p + ".final_layer_norm")
        linear(p + ".feed_forward.intermediate_dense", cfg.intermediate_size, d)
        linear(p + ".feed_forward.output_dense", d, cfg.intermediate_size)
        adapter(p + ".adapter_ffn", cfg.adapter_ffn)
    ln("layer_norm")
    linear("lm_head", cfg.vocab_size, d)
    return w


def is_trainable(name: str) -> bool:
    """Adapter parameters + lm_head (HF ``_get_adapters``, modeling_wav2vec2.py:1046-1060)."""
    return ".adapter_attn." in name or ".adapter_ffn." in name or name.startswith("lm_head.")


def from_product_state_dict(sd: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """Product ``JLForCTC.state_dict()`` ('encoder.*', 'lm_head.*') → oracle names, fp32 CPU copies."""
    out = {}
    for k, v in sd.items():
        k2 = k[len("encoder."):] if k.startswith("encoder.") else k
        out[k2] = v.detach().to("cpu", torch.float32).clone()
    return out


class _CTCFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, labels, input_lengths, blank, reduction, zero_infinity):
        loss, nll, grad = octc.ctc_loss_and_grad(logits, labels, input_lengths, blank, reduction, zero_infinity)
        ctx.save_for_backward(grad)
        return torch.tensor(loss, dtype=torch.float32)

    @staticmethod
    def backward(ctx, g):
        (grad,) = ctx.saved_tensors
        return grad * g, None, None, None, None, None


def ctc_loss(logits, labels, input_lengths, blank=0, reduction="sum", zero_infinity=False):
    return _CTCFn.apply(logits, labels, input_lengths, blank, reduction, zero_infinity)


def forward_from_features(w, cfg: OracleConfig, feats: torch.Tensor, frame_lengths: torch.Tensor,
                          labels: Optional[torch.Tensor] = None, dialect=0):
    h, lengths = oenc.encode(w, cfg, feats, frame_lengths, dialect)
    logits = oenc.lm_head(w, h)
    loss = None
    if labels is not None:
        if int(labels.max()) >= cfg.vocab_size:
            raise ValueError(f"Label values must be <= vocab_size: {cfg.vocab_size}")
        loss = ctc_loss(logits, labels, lengths, cfg.pad_token_id, cfg.ctc_loss_reduction, cfg.ctc_zero_infinity)
    return loss, logits, lengths


def forward_from_waveforms(w, cfg: OracleConfig, waveforms: Sequence[torch.Tensor],
                           labels: Optional[torch.Tensor] = None, dialect=0):
    if cfg.front_end == "wav2vec2":
        from . import w2v_frontend
        x, ns = w2v_frontend.normalize(waveforms)
        return forward_from_features(w, cfg, x, torch.tensor(ns), labels, dialect)
    feats, mask, flens = ofeat.extract(waveforms)
    return forward_from_features(w, cfg, feats, torch.tensor(flens), labels, dialect)


def transcribe(w, cfg: OracleConfig, waveforms: Sequence[torch.Tensor]) -> List[List[int]]:
    with torch.no_grad():
        _, logits, lengths = forward_from_waveforms(w, cfg, waveforms)
    return octc.greedy_decode(logits, lengths, cfg.pad_token_id)
