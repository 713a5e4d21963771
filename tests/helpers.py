"""Shared helpers for the test-suite (imports the product package through importlib: its directory name has a hyphen)."""
import importlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

PKG = "jiao-liao_speech_recognition_b200"


def pkg():
    return importlib.import_module(PKG)


def has_gpu() -> bool:
    return torch.cuda.is_available()


def synth_wave(num_samples: int, seed: int) -> torch.Tensor:
    """SURVEY §8d synthetic utterance: 0.1·randn + 0.05·Σ_5 sin(2π f_k t + φ_k), clipped to [-1, 1], 16 kHz fp32."""
    g = torch.Generator().manual_seed(seed)
    t = torch.arange(num_samples, dtype=torch.float64) / 16000.0
    x = 0.1 * torch.randn(num_samples, generator=g, dtype=torch.float64)
    for _ in range(5):
        f = 100.0 + 3900.0 * float(torch.rand(1, generator=g))
        ph = 2.0 * 3.141592653589793 * float(torch.rand(1, generator=g))
        x = x + 0.05 * torch.sin(2.0 * 3.141592653589793 * f * t + ph)
    return x.clamp(-1.0, 1.0).to(torch.float32)


def rel_err(a: torch.Tensor, b: torch.Tensor) -> float:
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def max_rel(a: torch.Tensor, b: torch.Tensor) -> float:
    """max |Δ| / max(|ref|, 1) elementwise (the mel tolerance of SURVEY §8d)."""
    a, b = a.double().cpu(), b.double().cpu()
    return float(((a - b).abs() / b.abs().clamp_min(1.0)).max())


def round_bf16_(model) -> None:
    """Make every parameter bf16-representable so the fp32 oracle and the bf16 kernels see identical weights."""
    with torch.no_grad():
        for p in model.parameters():
            p.copy_(p.to(torch.bfloat16).to(torch.float32))
