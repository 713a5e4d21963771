"""Pins oracle a3-a8 against the installed HF blocks they restate."""
import math

import pytest
import torch

from oracle import encoder as oe
from oracle import model as om

tr = pytest.importorskip("transformers")


def small_cfg(**kw):
    return om.OracleConfig(hidden_size=64, num_hidden_layers=2, num_attention_heads=4, intermediate_size=128,
                           conv_channels=96, vocab_size=31, **kw)


def test_conv_subsampler_and_positions_match_hf_speech2text():
    from transformers import Speech2TextConfig
    from transformers.models.speech_to_text.modeling_speech_to_text import Speech2TextEncoder
    cfg = Speech2TextConfig(d_model=64, encoder_layers=0, conv_channels=96, input_feat_per_channel=80,
                            encoder_attention_heads=4, encoder_ffn_dim=128, dropout=0.0, max_source_positions=400)
    enc = Speech2TextEncoder(cfg).eval()
    # drop the final layer norm so we can compare the embedding output
    enc.layer_norm = torch.nn.Identity()
    w = {f"conv.{i}.{n}": getattr(enc.conv.conv_layers[i], n).detach() for i in range(2) for n in ("weight", "bias")}
    g = torch.Generator().manual_seed(0)
    feats = torch.randn(2, 57, 80, generator=g)
    flen = torch.tensor([57, 30])
    mask = (torch.arange(57)[None] < flen[:, None]).long()
    feats = feats * mask[..., None]
    with torch.no_grad():
        ref = enc(feats, attention_mask=mask).last_hidden_state
    h = oe.conv_subsample(w, feats)
    lens = oe.subsampled_length(flen)
    assert lens.tolist() == [15, 8]
    got = oe.embed(h, lens)
    assert got.shape == ref.shape
    assert torch.allclose(got, ref, atol=1e-5), float((got - ref).abs().max())


def test_encoder_layer_matches_hf_wav2vec2_stable_ln():
    from transformers import Wav2Vec2Config
    from transformers.models.wav2vec2.modeling_wav2vec2 import Wav2Vec2EncoderLayerStableLayerNorm
    hc = Wav2Vec2Config(hidden_size=64, num_attention_heads=4, intermediate_size=128, hidden_dropout=0.0,
                        attention_dropout=0.0, activation_dropout=0.0, adapter_attn_dim=None)
    hc._attn_implementation = "eager"
    layer = Wav2Vec2EncoderLayerStableLayerNorm(hc).eval()
    w = {"layers.0." + k: v.detach() for k, v in layer.state_dict().items()}
    g = torch.Generator().manual_seed(1)
    x = torch.randn(2, 11, 64, generator=g)
    lens = torch.tensor([11, 7])
    bias = oe.key_bias(lens, 11).expand(2, 1, 11, 11).clone()
    bias[bias == float("-inf")] = torch.finfo(torch.float32).min
    with torch.no_grad():
        ref = layer(x, attention_mask=bias)[0]
    got = oe.encoder_layer(w, 0, x, lens, 4, None, None)
    # compare valid rows only (oracle zeroes padded rows)
    assert torch.allclose(got[0], ref[0], atol=2e-5)
    assert torch.allclose(got[1, :7], ref[1, :7], atol=2e-5)
    assert float(got[1, 7:].abs().max()) == 0.0


def test_wf_adapter_equals_dense_bottleneck_adapter():
    """With W = A·B materialised, WFAdapter must equal HF's bottleneck adapter (modeling_wav2vec2.py:931-953)."""
    from transformers import Wav2Vec2Config
    from transformers.models.wav2vec2.modeling_wav2vec2 import Wav2Vec2AttnAdapterLayer
    cfg = small_cfg(adapter_ffn="wf", wf_bottleneck=16, wf_rank=4)
    w = om.init_weights(cfg, seed=3)
    p = "layers.0.adapter_ffn"
    w[p + ".down_bias"] = torch.randn(1, 16) * 0.1
    w[p + ".up_bias"] = torch.randn(1, 64) * 0.1
    hc = Wav2Vec2Config(hidden_size=64, adapter_attn_dim=16)
    hf = Wav2Vec2AttnAdapterLayer(hc).eval()
    with torch.no_grad():
        hf.linear_1.weight.copy_(w[p + ".down_A"][0] @ w[p + ".down_B"][0])
        hf.linear_1.bias.copy_(w[p + ".down_bias"][0])
        hf.linear_2.weight.copy_(w[p + ".up_A"][0] @ w[p + ".up_B"][0])
        hf.linear_2.bias.copy_(w[p + ".up_bias"][0])
        hf.norm.weight.copy_(w[p + ".norm.weight"]); hf.norm.bias.copy_(w[p + ".norm.bias"])
    h = torch.randn(2, 9, 64)
    with torch.no_grad():
        ref = h + hf(h)
    got = oe.wf_adapter(w, p, h)
    assert torch.allclose(got, ref, atol=1e-5)


def test_att_adapter_matches_sdpa_and_masks_padding():
    cfg = small_cfg(adapter_ffn="att", att_dim=8)
    w = om.init_weights(cfg, seed=4)
    p = "layers.1.adapter_ffn"
    h = torch.randn(2, 10, 64)
    lens = torch.tensor([10, 6])
    got = oe.att_adapter(w, p, h, lens)
    # independent check with torch SDPA on the valid prefix of utterance 1
    z = torch.nn.functional.layer_norm(h[1:2, :6], (64,), w[p + ".norm.weight"], w[p + ".norm.bias"], 1e-5)
    q = z @ w[p + ".q_proj.weight"].T + w[p + ".q_proj.bias"]
    k = z @ w[p + ".k_proj.weight"].T + w[p + ".k_proj.bias"]
    v = z @ w[p + ".v_proj.weight"].T + w[p + ".v_proj.bias"]
    a = torch.nn.functional.scaled_dot_product_attention(q, k, v)
    ref = h[1:2, :6] + a @ w[p + ".o_proj.weight"].T + w[p + ".o_proj.bias"]
    assert torch.allclose(got[1:2, :6], ref, atol=1e-5)
    # changing padded frames must not change valid outputs
    h2 = h.clone(); h2[1, 6:] = 123.0
    got2 = oe.att_adapter(w, p, h2, lens)
    assert torch.allclose(got2[1, :6], got[1, :6], atol=1e-6)


def test_full_forward_shapes_grad_and_trainable_set():
    cfg = small_cfg(adapter_attn="att", adapter_ffn="wf", wf_bottleneck=16, wf_rank=4, att_dim=8)
    w = om.init_weights(cfg, seed=0)
    for k, v in w.items():
        v.requires_grad_(om.is_trainable(k))
    g = torch.Generator().manual_seed(5)
    feats = torch.randn(2, 61, 80, generator=g)
    flen = torch.tensor([61, 45])
    labels = torch.tensor([[3, 4, 5, -100], [7, 7, -100, -100]])
    loss, logits, lens = om.forward_from_features(w, cfg, feats, flen, labels)
    assert logits.shape == (2, 16, 31) and lens.tolist() == [16, 12]
    loss.backward()
    n_train = 0
    for k, v in w.items():
        if om.is_trainable(k):
            assert v.grad is not None and torch.isfinite(v.grad).all(), k
            n_train += 1
        else:
            assert v.grad is None
    assert n_train > 0
    # autograd CTC == torch CTC through the same graph
    w2 = {k: v.detach().clone().requires_grad_(om.is_trainable(k)) for k, v in w.items()}
    h, lens2 = oe.encode(w2, cfg, feats, flen)
    lg = oe.lm_head(w2, h)
    lp = torch.log_softmax(lg, -1).transpose(0, 1)
    ref = torch.nn.functional.ctc_loss(lp, labels[labels >= 0], lens2, (labels >= 0).sum(-1), blank=0, reduction="sum")
    ref.backward()
    assert abs(float(ref) - float(loss)) < 1e-4
    for k in w:
        if om.is_trainable(k):
            assert torch.allclose(w[k].grad, w2[k].grad, atol=1e-5, rtol=1e-4), k


def test_labels_out_of_vocab_raise():
    cfg = small_cfg()
    w = om.init_weights(cfg)
    with pytest.raises(ValueError):
        om.forward_from_features(w, cfg, torch.randn(1, 20, 80), torch.tensor([20]), torch.tensor([[31]]))


def test_relu_mask_flips_bound_bf16_wfadapter_down_gradients():
    """Why the WFAdapter's down-path gradients cannot meet a 3e-2 relative-Frobenius bound in bf16, independent of any kernel:
    with exact (fp64) arithmetic everywhere and ONE bf16 rounding — of LN(h), the operand every bf16 tensor-core implementation
    feeds to the first projection — the gradient of down_B moves by > 2 %, because the rounding flips the ReLU mask of the
    pre-activations that sit within ~0.3 % of zero (zero-bias N(0, 0.02) factors centre them on zero) and flipping a fraction f of
    dpre is a relative L2 error of sqrt(f).  The up-path gradients (no ReLU behind them) stay at the 0.3 % rounding level."""
    import torch.nn.functional as F
    torch.manual_seed(0)
    m, d, b, r = 2000, 256, 128, 16
    f64 = torch.float64

    def bf(x):
        return x.to(torch.bfloat16).to(f64)

    h = torch.randn(m, d, dtype=f64) * 1.5
    dy = torch.randn(m, d, dtype=f64) * 1e-3
    bd, ad = bf(torch.randn(r, d, dtype=f64) * 0.02), bf(torch.randn(b, r, dtype=f64) * 0.02)
    bu, au = bf(torch.randn(r, b, dtype=f64) * 0.02), bf(torch.randn(d, r, dtype=f64) * 0.02)
    z = F.layer_norm(h, (d,))

    def grads(zz):
        t1 = zz @ bd.T
        u = torch.relu(t1 @ ad.T)
        t2 = u @ bu.T
        dt2 = dy @ au
        dpre = (dt2 @ bu) * (u > 0)
        dt1 = dpre @ ad
        return {"up_A": dy.T @ t2, "up_B": dt2.T @ u, "down_A": dpre.T @ t1, "down_B": dt1.T @ zz}, (u > 0)

    exact, mask = grads(z)
    rounded, mask_r = grads(bf(z))
    flipped = float((mask != mask_r).double().mean())
    err = {k: float((rounded[k] - exact[k]).norm() / exact[k].norm()) for k in exact}
    assert 1e-4 < flipped < 1e-2, flipped                      # a fraction of a percent of the masks flip …
    assert err["down_B"] > 2e-2 and err["down_A"] > 2e-2, err   # … which alone costs the down path more than 2 %
    assert err["down_B"] < 8e-2 and err["down_A"] < 8e-2, err
    assert err["up_A"] < 5e-3 and err["up_B"] < 5e-3, err       # no ReLU behind the up path: plain rounding level
    assert abs(err["down_B"] - (2 * flipped) ** 0.5) < 0.6 * err["down_B"], (err, flipped)   # ≈ sqrt(flipped / active fraction)
