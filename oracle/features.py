"""Oracle stage a1/a2: Kaldi-compatible 80-bin log-mel filterbank + utterance CMVN.

Test infrastructure only (see ``oracle/__init__.py``).  Restates, in fp32:

* ``SP/torchaudio/compliance/kaldi.py:514-645`` (``fbank``) as called with the
  arguments at ``SP/transformers/models/speech_to_text/feature_extraction_speech_to_text.py:112-120``
  (num_mel_bins=80, sample_frequency=16000, dither=0, every other argument
  default: 25 ms / 10 ms frames, snip_edges, remove_dc_offset, preemphasis 0.97,
  povey window, round_to_power_of_two, use_power, low_freq 20, high_freq 0 →
  Nyquist, use_log_fbank, no energy column, no mean subtraction);
* ``feature_extraction_speech_to_text.py:142-163`` (``utterance_cmvn``) and the
  pad/attention-mask logic at ``:275-303``.
"""
from __future__ import annotations

import math
from typing import List, Sequence, Tuple

import numpy as np
import torch

SAMPLE_RATE = 16000
FRAME_LEN = 400          # 25 ms  (kaldi.py:140-141)
FRAME_SHIFT = 160        # 10 ms
NFFT = 512               # round_to_power_of_two (kaldi.py:142)
NUM_MEL = 80
LOW_FREQ = 20.0
PREEMPH = 0.97
FLT_EPS = float(torch.finfo(torch.float32).eps)   # kaldi.py:22 EPSILON


def num_frames(num_samples: int) -> int:
    """snip_edges frame count, kaldi.py:63-67."""
    if num_samples < FRAME_LEN:
        return 0
    return 1 + (num_samples - FRAME_LEN) // FRAME_SHIFT


def povey_window() -> torch.Tensor:
    """hann(400, symmetric) ** 0.85, kaldi.py:98-100."""
    n = torch.arange(FRAME_LEN, dtype=torch.float32)
    # torch.hann_window(N, periodic=False) = 0.5 - 0.5 cos(2 pi n / (N-1))
    hann = torch.hann_window(FRAME_LEN, periodic=False, dtype=torch.float32)
    del n
    return hann.pow(0.85)


def mel_scale(freq: torch.Tensor) -> torch.Tensor:
    """kaldi.py:266-267."""
    return 1127.0 * (1.0 + freq / 700.0).log()


def mel_banks() -> torch.Tensor:
    """[80, 257] fp32 triangular filters in mel space, kaldi.py:436-511, with the
    Nyquist column zero-padded as at kaldi.py:627."""
    nyquist = 0.5 * SAMPLE_RATE
    high_freq = nyquist
    fft_bin_width = SAMPLE_RATE / NFFT
    mel_low = 1127.0 * math.log(1.0 + LOW_FREQ / 700.0)      # mel_scale_scalar, kaldi.py:262-263
    mel_high = 1127.0 * math.log(1.0 + high_freq / 700.0)
    delta = (mel_high - mel_low) / (NUM_MEL + 1)
    b = torch.arange(NUM_MEL).unsqueeze(1)
    left = mel_low + b * delta
    center = mel_low + (b + 1.0) * delta
    right = mel_low + (b + 2.0) * delta
    mel = mel_scale(fft_bin_width * torch.arange(NFFT / 2)).unsqueeze(0)
    up = (mel - left) / (center - left)
    down = (right - mel) / (right - center)
    bins = torch.max(torch.zeros(1), torch.min(up, down))
    return torch.nn.functional.pad(bins, (0, 1), value=0.0).to(torch.float32)


def frame_signal(wave_i16scale: torch.Tensor) -> torch.Tensor:
    """[N] → [F, 400] strided frames (kaldi.py:54-82, snip_edges=True)."""
    n = wave_i16scale.shape[0]
    f = num_frames(n)
    if f == 0:
        return wave_i16scale.new_zeros((0, FRAME_LEN))
    return wave_i16scale.as_strided((f, FRAME_LEN), (FRAME_SHIFT, 1))


def fbank80(waveform: torch.Tensor) -> torch.Tensor:
    """[N] float32 waveform in [-1, 1] → [F, 80] fp32 log-mel.

    Steps: ×2^15 (feature_extraction_speech_to_text.py:111); frames; −frame mean
    (kaldi.py:183-186); pre-emphasis with replicate-left (:193-198); povey
    window (:201-204); zero-pad to 512 (:207-211); |rfft|² (:616-618);
    mel projection (:630); log(max(·, eps)) (:633).
    """
    assert waveform.dim() == 1
    x = waveform.to(torch.float32) * (2 ** 15)
    frames = frame_signal(x)
    if frames.shape[0] == 0:
        return torch.zeros((0, NUM_MEL), dtype=torch.float32)
    frames = frames - frames.mean(dim=1, keepdim=True)
    prev = torch.cat([frames[:, :1], frames[:, :-1]], dim=1)
    frames = frames - PREEMPH * prev
    frames = frames * povey_window().unsqueeze(0)
    frames = torch.nn.functional.pad(frames, (0, NFFT - FRAME_LEN))
    spec = torch.fft.rfft(frames).abs().pow(2.0)
    mel = torch.mm(spec, mel_banks().T)
    return torch.max(mel, torch.tensor(FLT_EPS)).log()


def utterance_cmvn(x: np.ndarray, input_length: int) -> np.ndarray:
    """Per-bin mean / population-std normalisation over the valid frames; padded
    frames := 0 (feature_extraction_speech_to_text.py:142-163).  No epsilon."""
    x = np.asarray(x, dtype=np.float32)
    mean = x[:input_length].mean(axis=0)
    x = np.subtract(x, mean)
    std = x[:input_length].std(axis=0)
    x = np.divide(x, std)
    if input_length < x.shape[0]:
        x[input_length:] = 0.0
    return x.astype(np.float32)


def extract(waveforms: Sequence[torch.Tensor]) -> Tuple[torch.Tensor, torch.Tensor, List[int]]:
    """List of [N_i] waveforms → (input_features [B, Fmax, 80] fp32,
    attention_mask [B, Fmax] int32, frame lengths).  Pad-to-longest then CMVN,
    as feature_extraction_speech_to_text.py:258-303."""
    feats = [fbank80(w).numpy() for w in waveforms]
    lens = [f.shape[0] for f in feats]
    fmax = max(lens) if lens else 0
    out = np.zeros((len(feats), fmax, NUM_MEL), dtype=np.float32)
    mask = np.zeros((len(feats), fmax), dtype=np.int32)
    for i, f in enumerate(feats):
        padded = np.zeros((fmax, NUM_MEL), dtype=np.float32)
        padded[: lens[i]] = f
        out[i] = utterance_cmvn(padded, lens[i])
        mask[i, : lens[i]] = 1
    return torch.from_numpy(out), torch.from_numpy(mask), lens
