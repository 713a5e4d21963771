"""GPU parity of the fused log-mel + CMVN kernels against the oracle (and the committed golden vectors)."""
import os

import numpy as np
import pytest
import torch

from helpers import ROOT, max_rel, pkg, synth_wave

pytestmark = pytest.mark.gpu


def _extract(waves, cmvn=True):
    P = pkg()
    fe = P.JLFeatureExtractor(device="cuda", do_ceptral_normalize=cmvn)
    out = fe([w.numpy() for w in waves], sampling_rate=16000, return_bf16=True)
    torch.cuda.synchronize()
    return out


def test_raw_fbank_matches_oracle_1e4():
    from oracle import features as of
    # frame counts whose last 32-frame tile is 6, 5, 1, 10, 20, 16 and 17 frames long: a warp carries frames f and f + 16
    waves = [synth_wave(160000, 1234), synth_wave(52345, 7), synth_wave(400, 9), synth_wave(16000 * 3 + 77, 11),
             synth_wave(400 + 160 * 51, 13), synth_wave(400 + 160 * 47 + 159, 14), synth_wave(400 + 160 * 48, 15)]
    out = _extract(waves, cmvn=False)
    feats = out["input_features"].cpu()
    for i, w in enumerate(waves):
        ref = of.fbank80(w)
        n = ref.shape[0]
        assert int(out["frame_lengths"][i]) == n
        err = max_rel(feats[i, :n], ref)
        assert err < 1e-4, f"utt {i}: mel rel err {err}"
        assert float(feats[i, n:].abs().max()) == 0.0 if n < feats.shape[1] else True


def test_cmvn_features_match_oracle_1e4_and_mask():
    from oracle import features as of
    waves = [synth_wave(160000, 1), synth_wave(100000, 2), synth_wave(31999, 3), synth_wave(160000, 4)]
    out = _extract(waves, cmvn=True)
    ref, mask, lens = of.extract(waves)
    got = out["input_features"].cpu()
    assert got.shape == ref.shape
    assert torch.equal(out["attention_mask"].cpu(), mask)
    assert out["frame_lengths"].cpu().tolist() == lens
    err = max_rel(got, ref)
    assert err < 1e-4, f"CMVN feature rel err {err}"
    got16 = out["input_features_bf16"].float().cpu()
    assert float((got16 - got).abs().max()) < 0.05           # bf16 copy of the same values


def test_too_short_and_empty_utterances():
    waves = [synth_wave(399, 5), synth_wave(1000, 6)]
    out = _extract(waves)
    assert out["frame_lengths"].cpu().tolist() == [0, 4]
    assert float(out["input_features"][0].abs().max()) == 0.0
    assert out["attention_mask"][0].sum().item() == 0


def test_golden_fbank_fixture():
    path = os.path.join(ROOT, "tests", "golden", "golden.npz")
    gold = np.load(path)
    waves = [torch.from_numpy(gold["wave0"]), torch.from_numpy(gold["wave1"])]
    out = _extract(waves, cmvn=True)
    ref = torch.from_numpy(gold["hf_input_features"])
    got = out["input_features"].cpu()
    assert got.shape == ref.shape
    assert max_rel(got, ref) < 1e-4


def test_wrong_sampling_rate_and_channels_raise():
    P = pkg()
    fe = P.JLFeatureExtractor(device="cuda")
    with pytest.raises(ValueError):
        fe([np.zeros(1000, dtype=np.float32)], sampling_rate=8000)
    with pytest.raises(ValueError):
        fe(np.zeros((2, 3, 1000), dtype=np.float32), sampling_rate=16000)


def test_full_size_linearity_property():
    """BASELINE config size (32 × 10 s): scaling the waveform by c shifts raw log-mel by 2·log(c) (size-independent check)."""
    P = pkg()
    fe = P.JLFeatureExtractor(device="cuda", do_ceptral_normalize=False)
    g = torch.Generator(device="cuda").manual_seed(0)
    wave = (0.1 * torch.randn(32, 160000, device="cuda", generator=g)).clamp(-1, 1)
    ns = torch.full((32,), 160000, dtype=torch.int32, device="cuda")
    a = fe.extract_device(wave, ns)["input_features"]
    b = fe.extract_device(wave * 0.5, ns)["input_features"]
    torch.cuda.synchronize()
    assert a.shape == (32, 998, 80)
    d = (a - b - 2.0 * float(np.log(2.0))).abs().max()
    assert float(d) < 1e-3
