"""Pins oracle a1/a2 against the installed dependency functions it restates."""
import numpy as np
import pytest
import torch

from oracle import features as of


def synth_wave(n, seed):
    g = torch.Generator().manual_seed(seed)
    t = torch.arange(n, dtype=torch.float32) / 16000.0
    w = 0.1 * torch.randn(n, generator=g)
    for _ in range(5):
        f = 100 + 3900 * torch.rand(1, generator=g)
        ph = 2 * np.pi * torch.rand(1, generator=g)
        w = w + 0.05 * torch.sin(2 * np.pi * f * t + ph)
    return w.clamp(-1, 1)


def test_num_frames():
    assert of.num_frames(160000) == 998
    assert of.num_frames(400) == 1
    assert of.num_frames(399) == 0
    assert of.num_frames(559) == 1
    assert of.num_frames(560) == 2


def test_fbank_matches_torchaudio_kaldi():
    ta = pytest.importorskip("torchaudio.compliance.kaldi")
    for n, seed in [(16000, 0), (160000, 1), (4321, 2)]:
        w = synth_wave(n, seed)
        ref = ta.fbank((w * 2 ** 15).unsqueeze(0), num_mel_bins=80, sample_frequency=16000, dither=0.0)
        got = of.fbank80(w)
        assert got.shape == ref.shape
        assert torch.allclose(got, ref, atol=2e-5, rtol=1e-6), float((got - ref).abs().max())


def test_window_and_mel_banks_match_torchaudio():
    ta = pytest.importorskip("torchaudio.compliance.kaldi")
    win = ta._feature_window_function("povey", 400, 0.42, torch.device("cpu"), torch.float32)
    assert torch.equal(win, of.povey_window())
    banks, _ = ta.get_mel_banks(80, 512, 16000.0, 20.0, 0.0, 100.0, -500.0, 1.0)
    assert torch.equal(torch.nn.functional.pad(banks, (0, 1)), of.mel_banks())


def test_extract_matches_hf_feature_extractor():
    tr = pytest.importorskip("transformers")
    fe = tr.Speech2TextFeatureExtractor(feature_size=80, num_mel_bins=80, sampling_rate=16000)
    waves = [synth_wave(32000, 3), synth_wave(20000, 4), synth_wave(48000, 5)]
    ref = fe([w.numpy() for w in waves], sampling_rate=16000, padding=True, return_tensors="pt",
             return_attention_mask=True)
    feats, mask, lens = of.extract(waves)
    assert lens == [198, 123, 298]
    assert torch.equal(mask.long(), ref["attention_mask"].long())
    assert torch.allclose(feats, ref["input_features"], atol=5e-5), float((feats - ref["input_features"]).abs().max())
    # padded frames are exactly zero
    assert float(feats[1, 123:].abs().max()) == 0.0


def test_pure_tone_lands_in_right_mel_bin():
    n = 4000
    t = torch.arange(n, dtype=torch.float32) / 16000.0
    f0 = 1000.0
    w = 0.5 * torch.sin(2 * np.pi * f0 * t)
    fb = of.fbank80(w)
    banks = of.mel_banks()
    k = int(round(f0 / (16000 / 512)))
    expect = int(torch.argmax(banks[:, k]))
    assert int(torch.argmax(fb.mean(0))) == expect


def test_cmvn_stats():
    x = np.random.RandomState(0).randn(50, 80).astype(np.float32) * 3 + 7
    y = of.utterance_cmvn(x.copy(), 40)
    assert np.allclose(y[:40].mean(0), 0, atol=1e-5)
    assert np.allclose(y[:40].std(0), 1, atol=1e-5)
    assert np.all(y[40:] == 0)
