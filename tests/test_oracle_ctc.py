"""Pins oracle a9/a10 against torch.nn.functional.ctc_loss and the HF collapse rule."""
import itertools

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import ctc as oc


def ref_ctc(logits, labels, input_lengths, blank, reduction, zero_infinity):
    logits = logits.clone().requires_grad_(True)
    tl = (labels >= 0).sum(-1)
    flat = labels.masked_select(labels >= 0)
    lp = F.log_softmax(logits, dim=-1, dtype=torch.float32).transpose(0, 1)
    with torch.backends.cudnn.flags(enabled=False):
        loss = F.ctc_loss(lp, flat, input_lengths, tl, blank=blank, reduction=reduction, zero_infinity=zero_infinity)
    loss.backward()
    return float(loss.detach()), logits.grad


def make_case(b, t, v, smax, seed, min_len=1):
    g = torch.Generator().manual_seed(seed)
    logits = torch.randn(b, t, v, generator=g) * 2
    in_len = torch.randint(max(min_len, t // 2), t + 1, (b,), generator=g)
    labels = torch.full((b, smax), -100, dtype=torch.long)
    for i in range(b):
        s = int(torch.randint(0, min(smax, int(in_len[i]) // 2) + 1, (1,), generator=g))
        labels[i, :s] = torch.randint(1, v, (s,), generator=g)
    return logits, labels, in_len


@pytest.mark.parametrize("reduction", ["sum", "mean"])
@pytest.mark.parametrize("seed", [0, 1, 2])
def test_matches_torch_ctc(reduction, seed):
    logits, labels, in_len = make_case(4, 30, 17, 9, seed)
    loss, nll, grad = oc.ctc_loss_and_grad(logits, labels, in_len, 0, reduction, False)
    rloss, rgrad = ref_ctc(logits, labels, in_len, 0, reduction, False)
    assert abs(loss - rloss) <= 1e-5 * max(1, abs(rloss))
    assert torch.allclose(grad, rgrad, atol=5e-5), float((grad - rgrad).abs().max())  # fp32 exp(alpha+beta) rounding


def test_repeated_labels_need_blank():
    # target "aa" needs at least 3 frames
    logits = torch.randn(1, 3, 5, generator=torch.Generator().manual_seed(3))
    labels = torch.tensor([[2, 2]])
    loss, _, grad = oc.ctc_loss_and_grad(logits, labels, torch.tensor([3]), 0, "sum", False)
    rloss, rgrad = ref_ctc(logits, labels, torch.tensor([3]), 0, "sum", False)
    lp = torch.log_softmax(logits[0], -1)
    hand = -(lp[0, 2] + lp[1, 0] + lp[2, 2])          # only path: a, blank, a
    assert abs(loss - float(hand)) < 1e-5
    assert abs(loss - rloss) < 1e-5
    assert torch.allclose(grad, rgrad, atol=5e-5)


def test_infeasible_and_zero_infinity():
    logits = torch.randn(2, 2, 5, generator=torch.Generator().manual_seed(4))
    labels = torch.tensor([[2, 2], [1, -100]])
    il = torch.tensor([2, 2])
    loss, nll, grad = oc.ctc_loss_and_grad(logits, labels, il, 0, "sum", False)
    assert np.isinf(loss)
    loss0, nll0, grad0 = oc.ctc_loss_and_grad(logits, labels, il, 0, "sum", True)
    rloss0, rgrad0 = ref_ctc(logits, labels, il, 0, "sum", True)
    assert abs(loss0 - rloss0) < 1e-5
    assert torch.allclose(grad0, rgrad0, atol=5e-5)
    assert float(grad0[0].abs().max()) == 0.0


def test_empty_target_and_single_frame():
    logits = torch.randn(1, 1, 4, generator=torch.Generator().manual_seed(5))
    labels = torch.full((1, 3), -100)
    loss, _, grad = oc.ctc_loss_and_grad(logits, labels, torch.tensor([1]), 0, "sum", False)
    lp = torch.log_softmax(logits[0, 0], -1)
    assert abs(loss + float(lp[0])) < 1e-6
    rloss, rgrad = ref_ctc(logits, labels, torch.tensor([1]), 0, "sum", False)
    assert abs(loss - rloss) < 1e-6 and torch.allclose(grad, rgrad, atol=1e-6)


def test_brute_force_small():
    # enumerate all alignments for T=4, V=3, target [1, 2]
    logits = torch.randn(1, 4, 3, generator=torch.Generator().manual_seed(6))
    lp = torch.log_softmax(logits[0], -1).double()
    target = [1, 2]
    tot = 0.0
    for path in itertools.product(range(3), repeat=4):
        col, prev = [], None
        for p in path:
            if p != prev and p != 0:
                col.append(p)
            prev = p
        if col == target:
            tot += float(torch.exp(sum(lp[t, p] for t, p in enumerate(path))))
    loss, _, _ = oc.ctc_loss_and_grad(logits, torch.tensor([target]), torch.tensor([4]), 0, "sum", False)
    assert abs(loss + np.log(tot)) < 1e-5


def test_greedy_matches_hf_rule():
    tr = pytest.importorskip("transformers")
    from itertools import groupby
    g = torch.Generator().manual_seed(7)
    logits = torch.randn(3, 40, 6, generator=g)
    # force runs and ties
    logits[0, 5:9] = logits[0, 5]
    logits[1, :, :] = 0.0          # all ties → argmax picks index 0 = blank everywhere
    lengths = [40, 40, 17]
    got = oc.greedy_decode(logits, lengths, blank=0)
    ids = logits.argmax(-1)
    for b in range(3):
        seq = ids[b, : lengths[b]].tolist()
        ref = [k for k, _ in groupby(seq) if k != 0]       # tokenization_wav2vec2.py:310-317
        assert got[b] == ref
    assert got[1] == []
