"""Oracle stages a3–a8: conv subsampling, positions, pre-LN transformer layers with
WFAdapter / AttAdapter slots, final LN, lm_head.  Plain fp32 PyTorch on CPU,
functional over a flat ``{name: tensor}`` weight dict (the same names as the
product modules' ``state_dict``) so that autograd gives the reference gradients.

Test infrastructure only (see ``oracle/__init__.py``).  Restates:

* conv subsampler  ``SP/transformers/models/speech_to_text/modeling_speech_to_text.py:68-100``
* length chain     ``…/modeling_speech_to_text.py:489-496``  ((L-1)//2+1 per conv)
* embed scale + sinusoid positions ``…/modeling_speech_to_text.py:542,568-579,104-139``
* pre-LN layer     ``SP/transformers/models/wav2vec2/modeling_wav2vec2.py:612-655``
  (``Wav2Vec2EncoderLayerStableLayerNorm``), attention ``:466-549`` with the
  eager math ``:438-463``, feed-forward ``:552-573``, erf-GELU
  (``SP/transformers/activations.py`` "gelu"), LN eps 1e-5
* adapter hook     ``…/modeling_wav2vec2.py:627-630,647-648``; bottleneck analogue ``:931-953``
* WFAdapter / AttAdapter: definitions fixed in SURVEY.md §8c from
  ``/root/reference/README.md:1`` + BASELINE.json north_star
* final LN ``:792``; lm_head ``:1630,1708``
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import torch
import torch.nn.functional as F

LN_EPS = 1e-5
W = Dict[str, torch.Tensor]


def subsampled_length(n: torch.Tensor | int, num_convs: int = 2):
    """(L-1)//2+1 per stride-2 conv, modeling_speech_to_text.py:489-496."""
    for _ in range(num_convs):
        n = (n - 1) // 2 + 1
    return n


def conv_subsample(w: W, feats: torch.Tensor) -> torch.Tensor:
    """[B, F, 80] → [B, T', d]: 2 × (Conv1d k5 s2 p2 → GLU over channels)."""
    h = feats.transpose(1, 2)
    for i in range(2):
        h = F.conv1d(h, w[f"conv.{i}.weight"], w[f"conv.{i}.bias"], stride=2, padding=2)
        h = F.glu(h, dim=1)
    return h.transpose(1, 2).contiguous()


def sinusoid_table(num_rows: int, dim: int) -> torch.Tensor:
    """tensor2tensor-style table, modeling_speech_to_text.py:123-139 (row 1 = padding row := 0)."""
    half = dim // 2
    e = math.log(10000) / (half - 1)
    e = torch.exp(torch.arange(half, dtype=torch.int64).float() * -e)
    e = torch.arange(num_rows, dtype=torch.int64).float().unsqueeze(1) * e.unsqueeze(0)
    tab = torch.cat([torch.sin(e), torch.cos(e)], dim=1).view(num_rows, -1)
    if dim % 2 == 1:
        tab = torch.cat([tab, torch.zeros(num_rows, 1)], dim=1)
    tab[1, :] = 0
    return tab


def embed(h: torch.Tensor, lengths: torch.Tensor) -> torch.Tensor:
    """×√d, + sinusoid row (t+2) on valid frames, zero row on padded frames
    (modeling_speech_to_text.py:542,568-579 with padding_idx = 1, offset = 2)."""
    b, t, d = h.shape
    tab = sinusoid_table(t + 2, d)
    valid = torch.arange(t).unsqueeze(0) < lengths.unsqueeze(1)
    pos_ids = torch.where(valid, torch.arange(t).unsqueeze(0) + 2, torch.ones(1, dtype=torch.long))
    return math.sqrt(d) * h + tab[pos_ids]


def layer_norm(x: torch.Tensor, w: W, prefix: str) -> torch.Tensor:
    return F.layer_norm(x, (x.shape[-1],), w[prefix + ".weight"], w[prefix + ".bias"], LN_EPS)


def key_bias(lengths: torch.Tensor, t: int) -> torch.Tensor:
    """Additive −inf on padded keys, [B, 1, 1, T]."""
    valid = torch.arange(t).unsqueeze(0) < lengths.unsqueeze(1)
    bias = torch.zeros(valid.shape, dtype=torch.float32)
    bias.masked_fill_(~valid, float("-inf"))
    return bias[:, None, None, :]


def self_attention(w: W, p: str, x: torch.Tensor, lengths: torch.Tensor, heads: int) -> torch.Tensor:
    """modeling_wav2vec2.py:500-549 / :438-463 (eager), scaling head_dim**-0.5."""
    b, t, d = x.shape
    dh = d // heads
    q = F.linear(x, w[p + ".q_proj.weight"], w[p + ".q_proj.bias"]).view(b, t, heads, dh).transpose(1, 2)
    k = F.linear(x, w[p + ".k_proj.weight"], w[p + ".k_proj.bias"]).view(b, t, heads, dh).transpose(1, 2)
    v = F.linear(x, w[p + ".v_proj.weight"], w[p + ".v_proj.bias"]).view(b, t, heads, dh).transpose(1, 2)
    s = torch.matmul(q, k.transpose(2, 3)) * (dh ** -0.5) + key_bias(lengths, t)
    a = torch.softmax(s, dim=-1)
    o = torch.matmul(a, v).transpose(1, 2).reshape(b, t, d)
    return F.linear(o, w[p + ".out_proj.weight"], w[p + ".out_proj.bias"])


def feed_forward(w: W, p: str, x: torch.Tensor) -> torch.Tensor:
    """Linear → erf-GELU → Linear, modeling_wav2vec2.py:566-573."""
    h = F.linear(x, w[p + ".intermediate_dense.weight"], w[p + ".intermediate_dense.bias"])
    h = F.gelu(h)
    return F.linear(h, w[p + ".output_dense.weight"], w[p + ".output_dense.bias"])


def wf_adapter(w: W, p: str, h: torch.Tensor, dialect=0) -> torch.Tensor:
    """WFAdapter: bottleneck adapter whose projections exist only as
    low-rank factors.  z = LN(h); u = relu((z B_dᵀ) A_dᵀ + c_d);
    y = (u B_uᵀ) A_uᵀ + c_u; out = h + y.  Leading dim of every factor = dialect;
    ``dialect`` is one id for the batch or one id per utterance (dialect_ids [B])."""
    if not isinstance(dialect, int):
        ids = [int(k) for k in dialect]
        if len(ids) != h.shape[0]:
            raise ValueError(f"dialect ids: expected {h.shape[0]} entries, got {len(ids)}")
        return torch.cat([wf_adapter(w, p, h[i:i + 1], ids[i]) for i in range(len(ids))], 0)
    z = layer_norm(h, w, p + ".norm")
    u = torch.relu(F.linear(F.linear(z, w[p + ".down_B"][dialect]), w[p + ".down_A"][dialect], w[p + ".down_bias"][dialect]))
    y = F.linear(F.linear(u, w[p + ".up_B"][dialect]), w[p + ".up_A"][dialect], w[p + ".up_bias"][dialect])
    return h + y


def att_adapter(w: W, p: str, h: torch.Tensor, lengths: torch.Tensor) -> torch.Tensor:
    """AttAdapter: z = LN(h); q,k,v = z W_{q,k,v}ᵀ + b ∈ R^b;
    a = softmax(q kᵀ/√b + keymask) v over the utterance's own frames;
    out = h + a W_oᵀ + b_o.  One head."""
    b, t, d = h.shape
    z = layer_norm(h, w, p + ".norm")
    q = F.linear(z, w[p + ".q_proj.weight"], w[p + ".q_proj.bias"])
    k = F.linear(z, w[p + ".k_proj.weight"], w[p + ".k_proj.bias"])
    v = F.linear(z, w[p + ".v_proj.weight"], w[p + ".v_proj.bias"])
    bd = q.shape[-1]
    s = torch.matmul(q, k.transpose(1, 2)) * (bd ** -0.5) + key_bias(lengths, t)[:, 0]
    a = torch.matmul(torch.softmax(s, dim=-1), v)
    return h + F.linear(a, w[p + ".o_proj.weight"], w[p + ".o_proj.bias"])


def fusion_adapter(w: W, p: str, h: torch.Tensor) -> torch.Tensor:
    """AdapterFusion-style AttAdapter over the K source-dialect adapters of the slot (SURVEY §8c ambiguity (ii), §8f f4;
    /root/reference/README.md:1 "multi-dialect knowledge transfer" + "adapter with attention"; Pfeiffer et al.'s AdapterFusion
    with the value projection fixed to the identity).  The K factor sets of ``<p>.source`` (a WFAdapter, one set per source
    dialect) are all applied to every frame; the frame then attends over their K outputs:
        y_k   = WFAdapter_k(h) - h                         (the k-th dialect adapter's update, k = 0..K-1)
        q     = LN_f(h) W_qᵀ + b_q            ∈ R^b
        key_k = y_k W_kᵀ + b_k                ∈ R^b
        α     = softmax_k(q · key_k / √b)                  (over the K dialects, per frame — no attention over time)
        out   = h + Σ_k α_k y_k
    No dialect id is needed: the fusion weights α pick the mixture of source dialects per frame."""
    src = p + ".source"
    kd = w[src + ".down_B"].shape[0]
    z = layer_norm(h, w, src + ".norm")
    ys = []
    for k in range(kd):
        u = torch.relu(F.linear(F.linear(z, w[src + ".down_B"][k]), w[src + ".down_A"][k], w[src + ".down_bias"][k]))
        ys.append(F.linear(F.linear(u, w[src + ".up_B"][k]), w[src + ".up_A"][k], w[src + ".up_bias"][k]))
    y = torch.stack(ys, 0)                                                    # [K, B, T, d]
    q = F.linear(layer_norm(h, w, p + ".norm"), w[p + ".q_proj.weight"], w[p + ".q_proj.bias"])      # [B, T, b]
    key = F.linear(y, w[p + ".k_proj.weight"], w[p + ".k_proj.bias"])        # [K, B, T, b]
    s = (q.unsqueeze(0) * key).sum(-1) * (q.shape[-1] ** -0.5)               # [K, B, T]
    alpha = torch.softmax(s, dim=0)
    return h + (alpha.unsqueeze(-1) * y).sum(0)


def apply_adapter(w: W, p: str, kind: Optional[str], h: torch.Tensor, lengths: torch.Tensor, dialect=0) -> torch.Tensor:
    if kind is None:
        return h
    if kind == "wf":
        return wf_adapter(w, p, h, dialect)
    if kind == "att":
        return att_adapter(w, p, h, lengths)
    if kind == "fuse":
        return fusion_adapter(w, p, h)
    raise ValueError(kind)


def zero_padded_rows(h: torch.Tensor, lengths: torch.Tensor) -> torch.Tensor:
    valid = (torch.arange(h.shape[1]).unsqueeze(0) < lengths.unsqueeze(1)).unsqueeze(-1)
    return h * valid


def encoder_layer(w: W, i: int, h: torch.Tensor, lengths: torch.Tensor, heads: int,
                  adapter_attn: Optional[str], adapter_ffn: Optional[str], dialect=0) -> torch.Tensor:
    """h += Attn(LN h); [adapter_attn]; h += FFN(LN h); [adapter_ffn]; padded rows := 0."""
    p = f"layers.{i}"
    h = h + self_attention(w, p + ".attention", layer_norm(h, w, p + ".layer_norm"), lengths, heads)
    h = apply_adapter(w, p + ".adapter_attn", adapter_attn, h, lengths, dialect)
    h = h + feed_forward(w, p + ".feed_forward", layer_norm(h, w, p + ".final_layer_norm"))
    h = apply_adapter(w, p + ".adapter_ffn", adapter_ffn, h, lengths, dialect)
    return zero_padded_rows(h, lengths)


def encode(w: W, cfg, feats: torch.Tensor, frame_lengths: torch.Tensor, dialect=0):
    """[B, F, 80] CMVN features + valid frame counts → (last_hidden_state [B, T', d], T' lengths)."""
    if getattr(cfg, "front_end", "mel") == "wav2vec2":
        # raw-waveform front end: feats = normalised waveforms [B, N], frame_lengths = valid sample counts
        from . import w2v_frontend
        h, lengths = w2v_frontend.front_end(w, cfg, feats, frame_lengths)
    else:
        h = conv_subsample(w, feats)
        lengths = subsampled_length(frame_lengths)
        h = embed(h, lengths)
    h = zero_padded_rows(h, lengths)
    for i in range(cfg.num_hidden_layers):
        h = encoder_layer(w, i, h, lengths, cfg.num_attention_heads, cfg.adapter_attn, cfg.adapter_ffn, dialect)
    h = layer_norm(h, w, "layer_norm")
    return h, lengths


def lm_head(w: W, h: torch.Tensor) -> torch.Tensor:
    return F.linear(h, w["lm_head.weight"], w["lm_head.bias"])
