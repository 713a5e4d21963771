"""Error rates on decoded output ("error rates out") — what the reference gets from ``jiwer`` and, for
word segmentation of Chinese text, ``jieba`` (/root/reference/requirements.txt:28,26; neither is installed here).
Host-side by nature: it consumes the compacted token ids / strings that ``jl_ctc_greedy`` hands back.

Definitions are jiwer's: rate = (substitutions + deletions + insertions) / reference length, accumulated over the whole
corpus (not a mean of per-utterance rates); CER counts characters with whitespace removed, WER counts tokens of a
segmenter (whitespace split by default; pass ``segment=jieba.lcut`` where jieba exists)."""
from __future__ import annotations

from typing import Callable, Iterable, List, Sequence, Tuple

import numpy as np


def edit_distance(ref: Sequence, hyp: Sequence) -> int:
    """Levenshtein distance (unit costs) between two token sequences; O(|ref|·|hyp|) time, O(|hyp|) memory, one numpy
    row update per reference token (the insertion chain is a running minimum, done with ``np.minimum.accumulate``)."""
    n, m = len(ref), len(hyp)
    if n == 0 or m == 0:
        return n + m
    vocab = {}
    r = np.fromiter((vocab.setdefault(t, len(vocab)) for t in ref), dtype=np.int64, count=n)
    h = np.fromiter((vocab.setdefault(t, len(vocab)) for t in hyp), dtype=np.int64, count=m)
    prev = np.arange(m + 1, dtype=np.int64)
    idx = np.arange(m + 1, dtype=np.int64)
    for i in range(1, n + 1):
        cur = np.empty(m + 1, dtype=np.int64)
        cur[0] = i
        # substitution / match and deletion candidates for every column
        cur[1:] = np.minimum(prev[:-1] + (h != r[i - 1]), prev[1:] + 1)
        # insertions: cur[j] = min(cur[j], cur[j-1] + 1)  ⇔  cur[j] - j = running min of (cur[j] - j)
        cur = np.minimum.accumulate(cur - idx) + idx
        prev = cur
    return int(prev[m])


def error_rate(refs: Iterable[Sequence], hyps: Iterable[Sequence]) -> Tuple[float, int, int]:
    """Corpus-level (edits / reference tokens, edits, reference tokens) over paired token sequences."""
    refs, hyps = list(refs), list(hyps)
    if len(refs) != len(hyps):
        raise ValueError(f"{len(refs)} references but {len(hyps)} hypotheses")
    edits = sum(edit_distance(r, h) for r, h in zip(refs, hyps))
    total = sum(len(r) for r in refs)
    if total == 0:
        raise ValueError("references are empty")
    return edits / total, edits, total


def _chars(s: str) -> List[str]:
    return [c for c in s if not c.isspace()]


def cer(refs: Iterable[str], hyps: Iterable[str]) -> float:
    """Character error rate (whitespace removed): the metric for Chinese dialect transcripts."""
    return error_rate([_chars(r) for r in refs], [_chars(h) for h in hyps])[0]


def wer(refs: Iterable[str], hyps: Iterable[str], segment: Callable[[str], List[str]] = str.split) -> float:
    """Word error rate over ``segment(text)`` tokens (whitespace split by default; ``jieba.lcut`` for Chinese words)."""
    return error_rate([segment(r) for r in refs], [segment(h) for h in hyps])[0]


def token_error_rate(ref_ids: Iterable[Sequence[int]], hyp_ids: Iterable[Sequence[int]]) -> float:
    """Error rate directly on token ids — ``JLForCTC.greedy_decode`` output against label sequences (negative = padding)."""
    strip = lambda seq: [int(t) for t in seq if int(t) >= 0]
    return error_rate([strip(r) for r in ref_ids], [strip(h) for h in hyp_ids])[0]
