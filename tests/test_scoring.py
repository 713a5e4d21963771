"""CPU tests of the error-rate host code against a brute-force recursion and hand-worked examples."""
import functools
import random

import pytest

from helpers import pkg


def _brute(a, b):
    @functools.lru_cache(None)
    def d(i, j):
        if i == 0 or j == 0:
            return i + j
        return min(d(i - 1, j) + 1, d(i, j - 1) + 1, d(i - 1, j - 1) + (a[i - 1] != b[j - 1]))
    return d(len(a), len(b))


def test_edit_distance_matches_brute_force():
    S = pkg().scoring
    rng = random.Random(0)
    for _ in range(200):
        a = tuple(rng.randrange(4) for _ in range(rng.randrange(0, 9)))
        b = tuple(rng.randrange(4) for _ in range(rng.randrange(0, 9)))
        assert S.edit_distance(a, b) == _brute(a, b), (a, b)
    assert S.edit_distance("kitten", "sitting") == 3
    assert S.edit_distance([], [1, 2]) == 2 and S.edit_distance([1, 2], []) == 2 and S.edit_distance([], []) == 0


def test_cer_wer_and_token_rates():
    S = pkg().scoring
    assert S.cer(["今天 天气 很好"], ["今天天气真好"]) == pytest.approx(1 / 6)
    assert S.wer(["the cat sat", "on the mat"], ["the cat sat", "on mat"]) == pytest.approx(1 / 6)
    assert S.wer(["今天 天气 很好"], ["今天 天气 真好"], segment=str.split) == pytest.approx(1 / 3)
    # corpus-level accumulation, not a mean of per-utterance rates
    assert S.cer(["ab", "cdefgh"], ["xx", "cdefgh"]) == pytest.approx(2 / 8)
    assert S.token_error_rate([[5, 6, 7, -100, -100]], [[5, 7]]) == pytest.approx(1 / 3)
    with pytest.raises(ValueError):
        S.error_rate([[1]], [])
    with pytest.raises(ValueError):
        S.error_rate([[]], [[1]])
