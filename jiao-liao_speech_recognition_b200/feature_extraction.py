"""``JLFeatureExtractor`` — drop-in for HF ``Speech2TextFeatureExtractor.__call__``
(``SP/transformers/models/speech_to_text/feature_extraction_speech_to_text.py:174-303``) whose arithmetic
(Kaldi fbank ``SP/torchaudio/compliance/kaldi.py:514-645`` + utterance CMVN ``:142-163``) runs in the fused
``jl_mel_cmvn_fwd`` kernels.  The host only pads the waveforms into one pinned buffer and copies it to the
device once; features, mask and lengths stay in HBM.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Union

import numpy as np
import torch

from . import _lib as L
from . import ops

SAMPLE_RATE = 16000
FRAME_LEN, FRAME_SHIFT, NFFT, NUM_MEL = 400, 160, 512, 80
LOW_FREQ = 20.0

_TABLES: Dict[str, dict] = {}


def num_frames(num_samples: int) -> int:
    """snip_edges frame count (kaldi.py:63-67)."""
    return 0 if num_samples < FRAME_LEN else 1 + (num_samples - FRAME_LEN) // FRAME_SHIFT


def _host_tables() -> dict:
    """Window, FFT twiddles and the sparse mel filterbank.  The window and the filter weights are computed with the
    same fp32 operation sequence torchaudio uses (kaldi.py:98-100 and :436-511) so that the table the kernel
    multiplies by is the table the reference multiplies by; twiddles come from float64."""
    window = torch.hann_window(FRAME_LEN, periodic=False, dtype=torch.float32).pow(0.85).numpy()      # povey
    kk = np.arange(NFFT, dtype=np.float64)
    twiddle = np.stack([np.cos(2.0 * np.pi * kk / NFFT), -np.sin(2.0 * np.pi * kk / NFFT)], axis=1)
    # triangular filters in mel space (vtln off); FFT bin NFFT/2 carries no weight (kaldi.py:627)
    mel_low = 1127.0 * math.log(1.0 + LOW_FREQ / 700.0)
    mel_high = 1127.0 * math.log(1.0 + (0.5 * SAMPLE_RATE) / 700.0)
    delta = (mel_high - mel_low) / (NUM_MEL + 1)
    bins = torch.arange(NUM_MEL).unsqueeze(1)
    left = mel_low + bins * delta
    center = mel_low + (bins + 1.0) * delta
    right = mel_low + (bins + 2.0) * delta
    fft_mel = (1127.0 * (1.0 + ((SAMPLE_RATE / NFFT) * torch.arange(NFFT / 2)) / 700.0).log()).unsqueeze(0)
    up = (fft_mel - left) / (center - left)
    down = (right - fft_mel) / (right - center)
    tri_all = torch.max(torch.zeros(1), torch.min(up, down)).numpy()                                   # [80, 256] fp32
    lo = np.zeros(NUM_MEL, dtype=np.int32)
    cnt = np.zeros(NUM_MEL, dtype=np.int32)
    w = np.zeros((NUM_MEL, L.JL_MEL_MAXW), dtype=np.float32)
    for b in range(NUM_MEL):
        tri = tri_all[b]
        nz = np.nonzero(tri > 0.0)[0]
        if len(nz) == 0:
            continue
        first, last = int(nz[0]), int(nz[-1])
        if last - first + 1 > L.JL_MEL_MAXW:
            raise RuntimeError("mel filter wider than JL_MEL_MAXW")
        lo[b], cnt[b] = first, last - first + 1
        w[b, : cnt[b]] = tri[first: last + 1]
    return {"window": window.astype(np.float32), "twiddle": twiddle.astype(np.float32), "mel_lo": lo, "mel_cnt": cnt, "mel_w": w}


def device_tables(device: torch.device) -> dict:
    key = str(device)
    if key not in _TABLES:
        _TABLES[key] = {k: torch.from_numpy(v).to(device).contiguous() for k, v in _host_tables().items()}
    return _TABLES[key]


class JLFeatureExtractor:
    """80-bin log-mel + utterance CMVN on the GPU, HF call signature.

    ``__call__(raw_speech, sampling_rate=16000, padding=True, return_tensors="pt", return_attention_mask=True)``
    returns ``{"input_features": [B, F, 80] fp32 CUDA, "attention_mask": [B, F] int32 CUDA, "frame_lengths": [B]}``
    (plus ``"input_features_bf16"`` when ``return_bf16``).  Raises ``ValueError`` for a wrong sampling rate
    (feature_extraction_speech_to_text.py:238-244) and for multi-channel input (:252-253).
    Fast path: ``extract_device(wave [B, N] fp32 CUDA, num_samples [B] int32 CUDA)``.
    """

    model_input_names = ["input_features", "attention_mask"]

    def __init__(self, feature_size: int = 80, sampling_rate: int = SAMPLE_RATE, num_mel_bins: int = 80, padding_value: float = 0.0,
                 do_ceptral_normalize: bool = True, normalize_means: bool = True, normalize_vars: bool = True,
                 device: Union[str, torch.device] = "cuda"):
        if feature_size != 80 or num_mel_bins != 80:
            raise ValueError("JLFeatureExtractor supports 80 mel bins (the reference path's configuration)")
        if not (normalize_means and normalize_vars) and do_ceptral_normalize:
            raise ValueError("partial CMVN (means only / vars only) is not on the reference path")
        self.feature_size, self.sampling_rate, self.num_mel_bins = feature_size, sampling_rate, num_mel_bins
        self.padding_value = padding_value
        self.do_ceptral_normalize = do_ceptral_normalize
        self.device = torch.device(device)
        L.load()   # fail loudly at construction if the CUDA library is missing

    # ---- device fast path
    def extract_device(self, wave: torch.Tensor, num_samples: torch.Tensor, max_frames: Optional[int] = None,
                       return_bf16: bool = False) -> dict:
        if max_frames is None:
            max_frames = max(num_frames(wave.shape[1]), 1)
        feats, feats16, mask, flen = ops.mel_cmvn(wave, num_samples, device_tables(wave.device), max_frames,
                                                  apply_cmvn=self.do_ceptral_normalize, want_bf16=return_bf16)
        out = {"input_features": feats, "attention_mask": mask, "frame_lengths": flen}
        if return_bf16:
            out["input_features_bf16"] = feats16
        return out

    # ---- HF-style host entry point
    def __call__(self, raw_speech, sampling_rate: Optional[int] = None, padding: Union[bool, str] = True,
                 return_tensors: Optional[str] = "pt", return_attention_mask: Optional[bool] = True, return_bf16: bool = False,
                 **kwargs) -> dict:
        if sampling_rate is not None and sampling_rate != self.sampling_rate:
            raise ValueError(
                f"The model corresponding to this feature extractor was trained using a sampling rate of {self.sampling_rate}. "
                f"Please make sure that the provided `raw_speech` input was sampled with {self.sampling_rate} and not {sampling_rate}.")
        if return_tensors not in (None, "pt"):
            raise ValueError("JLFeatureExtractor returns CUDA torch tensors (return_tensors='pt')")
        if padding not in (True, "longest"):
            raise ValueError("only padding=True / 'longest' is on the reference path")
        waves = self._as_list(raw_speech)
        lens = [int(w.shape[0]) for w in waves]
        nmax = max(max(lens), FRAME_LEN)
        nmax = (nmax + 3) // 4 * 4                                   # 16-byte aligned rows for the vector loads
        host = torch.zeros((len(waves), nmax), dtype=torch.float32).pin_memory()
        for i, w in enumerate(waves):
            host[i, : lens[i]] = torch.as_tensor(w, dtype=torch.float32)
        wave = host.to(self.device, non_blocking=True)
        nsamp = torch.tensor(lens, dtype=torch.int32).to(self.device, non_blocking=True)
        max_frames = max(max(num_frames(n) for n in lens), 1)
        out = self.extract_device(wave, nsamp, max_frames, return_bf16=return_bf16)
        if not return_attention_mask:
            out.pop("attention_mask")
        return out

    @staticmethod
    def _as_list(raw_speech) -> List:
        if isinstance(raw_speech, torch.Tensor):
            raw_speech = raw_speech.detach().cpu().numpy()
        if isinstance(raw_speech, np.ndarray):
            if raw_speech.ndim > 2:
                raise ValueError("Only mono-channel audio is supported for input to JLFeatureExtractor")
            return [raw_speech] if raw_speech.ndim == 1 else [r for r in raw_speech]
        if isinstance(raw_speech, (list, tuple)):
            if len(raw_speech) and isinstance(raw_speech[0], (float, int)):
                return [np.asarray(raw_speech, dtype=np.float32)]
            out = []
            for r in raw_speech:
                a = r.detach().cpu().numpy() if isinstance(r, torch.Tensor) else np.asarray(r, dtype=np.float32)
                if a.ndim != 1:
                    raise ValueError("Only mono-channel audio is supported for input to JLFeatureExtractor")
                out.append(a)
            return out
        raise ValueError("raw_speech must be a numpy array, a torch tensor or a list of them")


class JLWaveformFeatureExtractor:
    """Input side of the raw-waveform front end (``JLConfig.front_end = "wav2vec2"``), HF ``Wav2Vec2FeatureExtractor`` call
    signature (SP/transformers/models/wav2vec2/feature_extraction_wav2vec2.py:102-240): pads a list of mono 16 kHz waveforms
    to the longest, moves them to the GPU through pinned memory and returns ``{"input_values": [B, N] fp32 CUDA,
    "attention_mask": [B, N] int32 CUDA, "num_samples": [B] int32 CUDA}``.

    ``do_normalize`` (zero mean / unit variance per utterance, :78-97) is applied by the model's first kernels
    (``jl_wave_stats`` + ``jl_wave_im2col``), not here: ``input_values`` are the padded raw samples.  The normalisation is
    idempotent, so already-normalised input (HF's own extractor output) gives the same logits."""

    model_input_names = ["input_values", "attention_mask"]

    def __init__(self, feature_size: int = 1, sampling_rate: int = SAMPLE_RATE, padding_value: float = 0.0, do_normalize: bool = True,
                 return_attention_mask: bool = True, device: Union[str, torch.device] = "cuda"):
        if feature_size != 1:
            raise ValueError("JLWaveformFeatureExtractor takes mono audio (feature_size = 1)")
        if not do_normalize:
            raise ValueError("the reference path (XLS-R / MMS / wav2vec2-large) normalises every utterance: do_normalize must be True")
        self.feature_size, self.sampling_rate, self.padding_value = feature_size, sampling_rate, padding_value
        self.do_normalize, self.return_attention_mask = do_normalize, return_attention_mask
        self.device = torch.device(device)
        L.load()   # fail loudly at construction if the CUDA library is missing

    def __call__(self, raw_speech, sampling_rate: Optional[int] = None, padding: Union[bool, str] = True,
                 return_tensors: Optional[str] = "pt", return_attention_mask: Optional[bool] = None, **kwargs) -> dict:
        if sampling_rate is not None and sampling_rate != self.sampling_rate:
            raise ValueError(
                f"The model corresponding to this feature extractor was trained using a sampling rate of {self.sampling_rate}. "
                f"Please make sure that the provided `raw_speech` input was sampled with {self.sampling_rate} and not {sampling_rate}.")
        if return_tensors not in (None, "pt"):
            raise ValueError("JLWaveformFeatureExtractor returns CUDA torch tensors (return_tensors='pt')")
        if padding not in (True, "longest"):
            raise ValueError("only padding=True / 'longest' is on the reference path")
        waves = JLFeatureExtractor._as_list(raw_speech)
        lens = [int(w.shape[0]) for w in waves]
        nmax = (max(lens) + 3) // 4 * 4
        host = torch.full((len(waves), nmax), float(self.padding_value), dtype=torch.float32).pin_memory()
        for i, w in enumerate(waves):
            host[i, : lens[i]] = torch.as_tensor(w, dtype=torch.float32)
        out = {"input_values": host.to(self.device, non_blocking=True),
               "num_samples": torch.tensor(lens, dtype=torch.int32).to(self.device, non_blocking=True)}
        want_mask = self.return_attention_mask if return_attention_mask is None else return_attention_mask
        if want_mask:
            mask = torch.zeros((len(waves), nmax), dtype=torch.int32)
            for i, n in enumerate(lens):
                mask[i, :n] = 1
            out["attention_mask"] = mask.to(self.device, non_blocking=True)
        return out
