"""Oracle stages a9/a10: CTC loss (log-softmax + alpha/beta lattice) and greedy
collapse decode.  Test infrastructure only (see ``oracle/__init__.py``).

Restates:

* ``SP/transformers/models/wav2vec2/modeling_wav2vec2.py:1711-1736`` — labels
  padded with −100, ``target_lengths = (labels >= 0).sum(-1)``,
  ``log_softmax(logits, dtype=float32)``, ``ctc_loss(blank=pad_token_id,
  reduction=…, zero_infinity=…)`` with cuDNN disabled.
* ``SP/torch/nn/functional.py:3042-3115`` (``ctc_loss``): reduction "sum" = Σ_b
  nll_b, "mean" = mean_b(nll_b / max(target_len_b, 1)); ``zero_infinity``
  replaces infinite losses (and their gradients) by 0.  The lattice itself is
  the published CTC recursion (Graves et al. 2006, eqs. 6-8, 10-11, 16) that
  ATen's ``LossCTC.cpp`` implements; ATen's source is not in the container, so
  the recursion is pinned numerically against ``F.ctc_loss`` in
  ``tests/test_oracle_ctc.py``.
* ``SP/transformers/models/wav2vec2/tokenization_wav2vec2.py:310-317`` — greedy:
  ``groupby`` collapse of consecutive repeats, then drop the pad (= blank) token.
"""
from __future__ import annotations

from typing import List, Tuple

import numpy as np
import torch

NEG_INF = float("-inf")


def extended_labels(labels_b: np.ndarray, blank: int) -> np.ndarray:
    """l' = blank, l1, blank, l2, …, blank  (2S+1 states)."""
    s = len(labels_b)
    ext = np.full(2 * s + 1, blank, dtype=np.int64)
    ext[1::2] = labels_b
    return ext


def _lse(*xs: np.ndarray) -> np.ndarray:
    m = np.maximum.reduce(xs)
    m_safe = np.where(np.isneginf(m), 0.0, m)
    with np.errstate(divide="ignore"):
        return np.where(np.isneginf(m), NEG_INF,
                        m_safe + np.log(sum(np.exp(x - m_safe) for x in xs)))


def _shift_right(a: np.ndarray, k: int) -> np.ndarray:
    out = np.full_like(a, NEG_INF)
    if k < len(a):
        out[k:] = a[: len(a) - k]
    return out


def _shift_left(a: np.ndarray, k: int) -> np.ndarray:
    out = np.full_like(a, NEG_INF)
    if k < len(a):
        out[: len(a) - k] = a[k:]
    return out


def ctc_alpha_beta(lp: np.ndarray, ext: np.ndarray) -> Tuple[np.ndarray, np.ndarray, float]:
    """lp [T, V] log-probs (fp32), ext [2S+1] → (alpha [T, 2S+1], beta [T, 2S+1], nll)."""
    t_len = lp.shape[0]
    n = len(ext)
    dt = lp.dtype
    can_skip = np.zeros(n, dtype=bool)           # s-2 → s allowed iff l'_s != blank and l'_s != l'_{s-2}
    can_skip[2:] = (ext[2:] != ext[:-2]) & (np.arange(2, n) % 2 == 1)
    alpha = np.full((t_len, n), NEG_INF, dtype=dt)
    beta = np.full((t_len, n), NEG_INF, dtype=dt)
    if t_len == 0:
        return alpha, beta, float("inf") if n > 1 else 0.0
    alpha[0, 0] = lp[0, ext[0]]
    if n > 1:
        alpha[0, 1] = lp[0, ext[1]]
    for t in range(1, t_len):
        a = alpha[t - 1]
        a1 = _shift_right(a, 1)
        a2 = np.where(can_skip, _shift_right(a, 2), NEG_INF).astype(dt)
        alpha[t] = (_lse(a, a1, a2) + lp[t, ext]).astype(dt)
    last = alpha[t_len - 1, n - 1]
    last2 = alpha[t_len - 1, n - 2] if n > 1 else np.array(NEG_INF, dtype=dt)
    nll = -float(_lse(np.asarray(last), np.asarray(last2)))
    beta[t_len - 1, n - 1] = lp[t_len - 1, ext[n - 1]]
    if n > 1:
        beta[t_len - 1, n - 2] = lp[t_len - 1, ext[n - 2]]
    can_skip_fwd = np.zeros(n, dtype=bool)       # s → s+2
    can_skip_fwd[:-2] = can_skip[2:]
    for t in range(t_len - 2, -1, -1):
        b = beta[t + 1]
        b1 = _shift_left(b, 1)
        b2 = np.where(can_skip_fwd, _shift_left(b, 2), NEG_INF).astype(dt)
        beta[t] = (_lse(b, b1, b2) + lp[t, ext]).astype(dt)
    return alpha, beta, nll


def ctc_loss_and_grad(logits: torch.Tensor, labels: torch.Tensor, input_lengths: torch.Tensor,
                      blank: int = 0, reduction: str = "sum", zero_infinity: bool = False):
    """logits [B, T, V] fp32, labels [B, Smax] (−100 padded), input_lengths [B]
    → (loss scalar, per-utterance nll [B], d loss / d logits [B, T, V]).

    grad wrt logits through log_softmax:  softmax(v) − Σ_{s: l'_s = v} exp(α_t(s) + β_t(s) − lp_t(v) + nll),
    scaled by the reduction weight; zero for t ≥ input_length."""
    logits = logits.detach().to(torch.float32)
    bsz, tmax, v = logits.shape
    lp_all = torch.log_softmax(logits, dim=-1).numpy()
    lab = labels.numpy()
    nll = np.zeros(bsz, dtype=np.float32)
    grad = np.zeros((bsz, tmax, v), dtype=np.float32)
    tlens = (lab >= 0).sum(-1)
    for b in range(bsz):
        t_len = int(input_lengths[b])
        ext = extended_labels(lab[b, : tlens[b]], blank)
        lp = lp_all[b, :t_len]
        alpha, beta, nll_b = ctc_alpha_beta(lp, ext)
        nll[b] = nll_b
        if reduction == "mean":
            scale = 1.0 / (max(int(tlens[b]), 1) * bsz)
        else:
            scale = 1.0
        if np.isinf(nll_b):
            if zero_infinity:
                nll[b] = 0.0
                continue                      # grad stays 0
            grad[b, :t_len] = np.nan          # ATen propagates NaN grads for an infeasible alignment
            continue
        g = np.exp(lp)
        with np.errstate(over="ignore", invalid="ignore"):
            occ = np.exp(alpha + beta - lp[:, ext] + np.float32(nll_b))   # [T, 2S+1]
        occ = np.where(np.isfinite(occ), occ, 0.0)
        for s, lab_s in enumerate(ext):
            g[:, lab_s] -= occ[:, s]
        grad[b, :t_len] = g * scale
    if reduction == "mean":
        loss = float(np.mean(nll / np.maximum(tlens, 1)))
    else:
        loss = float(np.sum(nll))
    return loss, torch.from_numpy(nll), torch.from_numpy(grad)


def greedy_decode(logits: torch.Tensor, lengths, blank: int = 0) -> List[List[int]]:
    """argmax (first max wins) over V for frames < length, collapse consecutive
    repeats, drop blank."""
    out = []
    ids = torch.argmax(logits, dim=-1)
    for b in range(logits.shape[0]):
        seq = ids[b, : int(lengths[b])].tolist()
        dec, prev = [], None
        for tok in seq:
            if tok != prev:
                if tok != blank:
                    dec.append(tok)
            prev = tok
        out.append(dec)
    return out
