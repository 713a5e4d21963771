"""CPU tests of host-side logic that needs no GPU: dialect runs, graph keys, oracle per-utterance dialect selection."""
import pytest
import torch

from helpers import pkg


def test_dialect_segments_runs_and_errors():
    E = pkg().JLEngine
    assert E.dialect_segments(0, 4, 1) == [(0, 0, 4)]
    assert E.dialect_segments([0, 0, 1, 1], 4, 2) == [(0, 0, 2), (1, 2, 4)]
    assert E.dialect_segments(torch.tensor([1, 0]), 2, 2) == [(1, 0, 1), (0, 1, 2)]
    with pytest.raises(ValueError, match="adjacent"):
        E.dialect_segments([0, 1, 0], 3, 2)
    with pytest.raises(ValueError, match="out of range"):
        E.dialect_segments([0, 2], 2, 2)
    with pytest.raises(ValueError, match="one per utterance"):
        E.dialect_segments([0], 2, 2)
    with pytest.raises(ValueError, match="out of range"):
        E.dialect_segments(1, 2, 1)


def test_dialect_key_is_hashable():
    T = pkg().training
    assert T._dialect_key(2) == 2
    assert T._dialect_key([0, 1]) == (0, 1)
    assert T._dialect_key(torch.tensor([3, 3])) == (3, 3)
    assert hash(T._dialect_key([0, 1])) == hash((0, 1))


def test_oracle_per_utterance_dialects_equal_per_utterance_calls():
    from oracle import encoder as oe
    g = torch.Generator().manual_seed(0)
    d, b, r, K = 32, 16, 4, 3
    w = {"a.norm.weight": torch.ones(d), "a.norm.bias": torch.zeros(d),
         "a.down_B": torch.randn(K, r, d, generator=g) * 0.1, "a.down_A": torch.randn(K, b, r, generator=g) * 0.1,
         "a.down_bias": torch.randn(K, b, generator=g) * 0.1, "a.up_B": torch.randn(K, r, b, generator=g) * 0.1,
         "a.up_A": torch.randn(K, d, r, generator=g) * 0.1, "a.up_bias": torch.randn(K, d, generator=g) * 0.1}
    h = torch.randn(3, 5, d, generator=g)
    out = oe.wf_adapter(w, "a", h, dialect=[2, 0, 0])
    for i, k in enumerate([2, 0, 0]):
        assert torch.equal(out[i], oe.wf_adapter(w, "a", h[i:i + 1], dialect=k)[0])
    assert not torch.allclose(out[0], oe.wf_adapter(w, "a", h[0:1], dialect=0)[0])
    with pytest.raises(ValueError):
        oe.wf_adapter(w, "a", h, dialect=[0, 1])


def test_weights_version_sees_parameter_writes_and_a_replaced_head_without_walking_the_modules_every_call():
    """JLEngine.weights_version (what AdapterTrainer / Transcriber compare before every graph replay): changes on an in-place write
    to any parameter and when lm_head is replaced; the Parameter list is cached (the module walk costs ~1 ms per call)."""
    import torch.nn as nn
    P = pkg()
    cfg = P.JLConfig(hidden_size=128, num_hidden_layers=2, num_attention_heads=2, intermediate_size=256, conv_channels=64, vocab_size=40,
                     adapter_attn="att", adapter_ffn="wf", wf_bottleneck=32, wf_rank=8)
    model = P.JLForCTC(cfg)
    eng = model.encoder.engine(model.lm_head)
    v0 = eng.weights_version(False)
    assert eng.weights_version(False) == v0 and eng._vparams is not None
    cached = eng._vparams[1]
    with torch.no_grad():
        model.encoder.layers[1].adapter_ffn.up_bias.add_(1.0)
    v1 = eng.weights_version(False)
    assert v1 != v0 and eng._vparams[1] is cached                     # same cached list, new version
    with torch.no_grad():
        model.encoder.layers[0].layer_norm.weight.mul_(1.0)            # a backbone write counts too
    assert eng.weights_version(False) != v1
    v2 = eng.weights_version(False)
    model.lm_head = nn.Linear(cfg.hidden_size, 48)
    eng = model.encoder.engine(model.lm_head)
    assert eng.weights_version(False) != v2 and eng._vparams[1] is not cached
