"""Oracle restatement of the wav2vec2 / XLS-R raw-waveform front end: the second front end beside the mel
path.  fp32 CPU; test infrastructure only (see ``oracle/__init__.py``).  "layer" feature-extractor norm (XLS-R, MMS,
wav2vec2-large-lv60) with conv bias, stable-layer-norm encoder arrangement.

Follows, stage by stage (SP = site-packages of the build container):
* utterance normalisation — ``SP/transformers/models/wav2vec2/feature_extraction_wav2vec2.py:78-97`` (zero mean, unit
  variance over the valid samples, eps 1e-7 inside the sqrt, padding := 0);
* 7 × [Conv1d(k, s, bias) → LayerNorm(512) → GELU] — ``modeling_wav2vec2.py:275-299`` (layer), ``:382-420`` (stack);
* output lengths ``floor((L - k) / s) + 1`` per layer — ``:1005-1020``;
* feature projection LayerNorm(512) → Linear(512 → d) — ``:422-434``;
* padded frames := 0, then ``h + GELU(pos_conv(h))`` with a weight-normed (dim = 2) grouped Conv1d(d, d, k = 128,
  padding = 64, groups = 16) whose last output frame is dropped — ``:326-379`` and ``:742-766``.
"""
from __future__ import annotations

from typing import Dict, List, Sequence, Tuple

import torch
import torch.nn.functional as F

W = Dict[str, torch.Tensor]


def normalize(waves: Sequence[torch.Tensor]) -> Tuple[torch.Tensor, List[int]]:
    """list of 1-D fp32 waveforms → ([B, N_max] zero-mean / unit-variance per utterance, zero padded; sample counts)."""
    lens = [int(x.shape[0]) for x in waves]
    out = torch.zeros((len(waves), max(lens)), dtype=torch.float32)
    for i, x in enumerate(waves):
        x = x.to(torch.float32)
        out[i, : lens[i]] = (x - x.mean()) / torch.sqrt(x.var(unbiased=False) + 1e-7)
    return out, lens


def conv_lengths(n, kernels: Sequence[int], strides: Sequence[int]):
    for k, s in zip(kernels, strides):
        n = torch.div(n - k, s, rounding_mode="floor") + 1 if torch.is_tensor(n) else (n - k) // s + 1
    return n


def feature_encoder(w: W, x: torch.Tensor, kernels: Sequence[int], strides: Sequence[int]) -> torch.Tensor:
    """[B, N] → [B, T, C]."""
    h = x[:, None, :]
    for i, s in enumerate(strides):
        h = F.conv1d(h, w[f"w2v.conv.{i}.weight"], w[f"w2v.conv.{i}.bias"], stride=s)
        h = F.layer_norm(h.transpose(1, 2), (h.shape[1],), w[f"w2v.conv_norm.{i}.weight"], w[f"w2v.conv_norm.{i}.bias"], 1e-5).transpose(1, 2)
        h = F.gelu(h)
    return h.transpose(1, 2)


def feature_projection(w: W, h: torch.Tensor, eps: float = 1e-5) -> torch.Tensor:
    h = F.layer_norm(h, (h.shape[-1],), w["w2v.proj_norm.weight"], w["w2v.proj_norm.bias"], eps)
    return F.linear(h, w["w2v.proj.weight"], w["w2v.proj.bias"])


def pos_conv_weight(w: W) -> torch.Tensor:
    """weight_norm(dim = 2): W[:, :, k] = g[k] · V[:, :, k] / ‖V[:, :, k]‖_F."""
    v, g = w["w2v.pos_conv.weight_v"], w["w2v.pos_conv.weight_g"]
    return g * v / v.norm(dim=(0, 1), keepdim=True)


def pos_conv(w: W, h: torch.Tensor, groups: int) -> torch.Tensor:
    """[B, T, d] → GELU(conv)[B, T, d]; even kernel → the extra last frame is removed."""
    weight = pos_conv_weight(w)
    k = weight.shape[2]
    y = F.conv1d(h.transpose(1, 2), weight, w["w2v.pos_conv.bias"], padding=k // 2, groups=groups)
    if k % 2 == 0:
        y = y[:, :, :-1]
    return F.gelu(y).transpose(1, 2)


def front_end(w: W, cfg, x: torch.Tensor, sample_lengths: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """normalised waveforms [B, N] + valid sample counts → (hidden states entering layer 0 [B, T, d], frame lengths)."""
    h = feature_encoder(w, x, cfg.conv_kernel, cfg.conv_stride)
    lengths = conv_lengths(sample_lengths.to(torch.int64), cfg.conv_kernel, cfg.conv_stride).clamp_min(0)
    h = feature_projection(w, h)
    valid = (torch.arange(h.shape[1]).unsqueeze(0) < lengths.unsqueeze(1)).unsqueeze(-1)
    h = h * valid
    h = h + pos_conv(w, h, cfg.num_conv_pos_embedding_groups)
    return h, lengths
