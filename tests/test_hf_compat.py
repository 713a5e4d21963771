"""CPU tests of checkpoint-name compatibility with the reference's pinned HF stack: a real
``Wav2Vec2ForCTC`` transformer stack with HF's per-language bottleneck adapter loads into ``JLForCTC`` by name and the
oracle reproduces HF's own layer outputs from the loaded tensors; Speech2Text names round-trip; adapter files."""
import os

import pytest
import torch

from helpers import pkg


def _hf_model(adapter_dim=16):
    from transformers import Wav2Vec2Config, Wav2Vec2ForCTC
    cfg = Wav2Vec2Config(hidden_size=128, num_hidden_layers=2, num_attention_heads=2, intermediate_size=256, conv_dim=(8,), conv_kernel=(10,),
                         conv_stride=(5,), num_feat_extract_layers=1, num_conv_pos_embeddings=4, num_conv_pos_embedding_groups=2,
                         vocab_size=24, adapter_attn_dim=adapter_dim, do_stable_layer_norm=True, hidden_dropout=0.0, attention_dropout=0.0,
                         activation_dropout=0.0, layerdrop=0.0, final_dropout=0.0, feat_proj_dropout=0.0)
    torch.manual_seed(3)
    m = Wav2Vec2ForCTC(cfg).eval()
    with torch.no_grad():                          # HF initialises the adapter's last projection to ~0: make it matter
        for layer in m.wav2vec2.encoder.layers:
            layer.adapter_layer.linear_1.weight.normal_(0, 0.3)
            layer.adapter_layer.linear_2.weight.normal_(0, 0.3)
            layer.adapter_layer.linear_1.bias.normal_(0, 0.1)
            layer.adapter_layer.linear_2.bias.normal_(0, 0.1)
    return m


def _jl_model(P, num_dialects=1):
    cfg = P.JLConfig(hidden_size=128, num_hidden_layers=2, num_attention_heads=2, intermediate_size=256, conv_channels=64, vocab_size=24,
                     adapter_ffn="wf", wf_bottleneck=16, wf_rank=16, num_dialects=num_dialects)
    return P.JLForCTC(cfg)


def test_wav2vec2_stack_and_hf_adapter_load_by_name_and_reproduce_hf_layer_outputs():
    P = pkg()
    from oracle import encoder as oe
    from oracle import model as om
    hf = _hf_model()
    jl = _jl_model(P)
    missing, skipped = jl.load_hf_state_dict(hf.state_dict())
    # the raw-waveform front end has no counterpart; the mel conv subsampler is not in a wav2vec2 checkpoint
    assert all(k.startswith(("wav2vec2.feature_", "wav2vec2.encoder.pos_conv_embed", "wav2vec2.masked_spec_embed")) for k in skipped), skipped
    assert all(k.startswith("encoder.conv.") for k in missing), missing
    w = om.from_product_state_dict(jl.state_dict())
    g = torch.Generator().manual_seed(0)
    h = torch.randn(2, 20, 128, generator=g)
    lengths = torch.tensor([20, 20])
    ref = h
    with torch.no_grad():
        for layer in hf.wav2vec2.encoder.layers:
            ref = layer(ref)[0]
        ref = hf.wav2vec2.encoder.layer_norm(ref)
        ref_logits = hf.lm_head(ref)
        out = h
        for i in range(2):
            out = oe.encoder_layer(w, i, out, lengths, 2, None, "wf")
        out = oe.layer_norm(out, w, "layer_norm")
        logits = oe.lm_head(w, out)
    assert torch.allclose(out, ref, atol=2e-5, rtol=1e-5), float((out - ref).abs().max())
    assert torch.allclose(logits, ref_logits, atol=2e-5, rtol=1e-5)
    # and the adapter is not a no-op in this check
    with torch.no_grad():
        plain = h
        for i in range(2):
            plain = oe.encoder_layer(w, i, plain, lengths, 2, None, None)
    assert float((oe.layer_norm(plain, w, "layer_norm") - ref).abs().max()) > 1e-2


def test_hf_adapter_file_loads_into_a_dialect_slot(tmp_path):
    """HF's ``adapter.<lang>.safetensors`` (the tensors of ``Wav2Vec2ForCTC._get_adapters()``) → factor set k of the WFAdapter."""
    from safetensors.torch import save_file
    P = pkg()
    hf = _hf_model()
    sd = {k: v.detach().clone().contiguous() for k, v in hf._get_adapters().items()}
    save_file(sd, str(tmp_path / "adapter.jiaoliao.safetensors"), metadata={"format": "pt"})
    jl = _jl_model(P, num_dialects=3)
    before = jl.encoder.layers[0].adapter_ffn.down_B.detach().clone()
    jl.load_adapter("jiaoliao", model_dir=str(tmp_path), dialect=2)
    ad = jl.encoder.layers[0].adapter_ffn
    hf_ad = hf.wav2vec2.encoder.layers[0].adapter_layer
    assert torch.equal(ad.down_B[2], hf_ad.linear_1.weight)
    assert torch.equal(ad.down_A[2], torch.eye(16))
    assert torch.equal(ad.up_A[2], hf_ad.linear_2.weight)
    assert torch.equal(ad.up_B[2], torch.eye(16))
    assert torch.equal(ad.down_bias[2], hf_ad.linear_1.bias) and torch.equal(ad.up_bias[2], hf_ad.linear_2.bias)
    assert torch.equal(ad.down_B[:2], before[:2]), "other dialects' factor sets must stay untouched"
    assert torch.equal(jl.lm_head.weight, hf.lm_head.weight)
    with pytest.raises(EnvironmentError):
        jl.load_adapter("cantonese", model_dir=str(tmp_path))
    with pytest.raises(ValueError):
        bad = dict(sd)
        bad["wav2vec2.encoder.layers.0.adapter_layer.linear_1.weight"] = torch.zeros(8, 128)
        save_file(bad, str(tmp_path / "adapter.bad.safetensors"))
        jl.load_adapter("bad", model_dir=str(tmp_path))


def test_speech_to_text_names_round_trip():
    P = pkg()
    H = P.hf_compat
    jl = _jl_model(P)
    for style in ("speech_to_text", "wav2vec2"):
        hf_sd = H.to_hf_state_dict(jl, style=style)
        if style == "speech_to_text":
            assert "model.encoder.conv.conv_layers.0.weight" in hf_sd and "model.encoder.layers.1.self_attn.q_proj.weight" in hf_sd
            assert "model.encoder.layers.0.fc1.bias" in hf_sd and "model.encoder.layers.0.self_attn_layer_norm.weight" in hf_sd
        else:
            assert "wav2vec2.encoder.layers.1.feed_forward.output_dense.weight" in hf_sd
        back, skipped = H.convert_hf_state_dict(hf_sd)
        own = jl.state_dict()
        assert not skipped
        assert set(back) == set(own)
        for k in own:
            assert torch.equal(back[k], own[k]), k
    # a real Speech2TextEncoder state dict (bare names) maps onto the conv subsampler + layers
    from transformers import Speech2TextConfig
    from transformers.models.speech_to_text.modeling_speech_to_text import Speech2TextEncoder
    enc = Speech2TextEncoder(Speech2TextConfig(d_model=128, encoder_layers=2, encoder_attention_heads=2, encoder_ffn_dim=256, conv_channels=64,
                                               input_feat_per_channel=80, num_conv_layers=2))
    conv, skipped = H.convert_hf_state_dict(enc.state_dict())
    own = {k: v for k, v in jl.state_dict().items() if ".adapter_" not in k and not k.startswith("lm_head")}
    assert set(conv) == set(own), set(conv) ^ set(own)
    for k in own:
        assert tuple(conv[k].shape) == tuple(own[k].shape), k


def test_own_adapter_files_safetensors_and_bin(tmp_path):
    P = pkg()
    a, b = _jl_model(P, num_dialects=2), _jl_model(P, num_dialects=2)
    a.init_adapter_layers(seed=7)
    for name in ("adapter.x.safetensors", "adapter.x.bin"):
        path = str(tmp_path / name)
        a.save_adapter(path)
        b.init_adapter_layers(seed=9)
        b.load_adapter(path)
        for (n1, p1), (n2, p2) in zip(sorted(a._get_adapters().items()), sorted(b._get_adapters().items())):
            assert n1 == n2 and torch.equal(p1, p2), n1
    assert P.hf_compat.adapter_file(str(tmp_path), "x").endswith("adapter.x.safetensors")
    os.remove(str(tmp_path / "adapter.x.safetensors"))
    assert P.hf_compat.adapter_file(str(tmp_path), "x").endswith("adapter.x.bin")


def test_post_ln_group_norm_checkpoints_are_refused():
    """wav2vec2-base style checkpoints (do_stable_layer_norm=False, feat_extract_norm="group", no conv bias) share the transformer
    leaf names with XLS-R / MMS but not the arithmetic (post-LN layers): loading them must raise, by config or by key layout."""
    from transformers import Wav2Vec2Config, Wav2Vec2ForCTC
    P = pkg()
    cfg = Wav2Vec2Config(hidden_size=128, num_hidden_layers=2, num_attention_heads=2, intermediate_size=256, conv_dim=(8, 8), conv_kernel=(10, 3),
                         conv_stride=(5, 2), num_feat_extract_layers=2, num_conv_pos_embeddings=4, num_conv_pos_embedding_groups=2,
                         vocab_size=24, do_stable_layer_norm=False, feat_extract_norm="group", conv_bias=False)
    hf = Wav2Vec2ForCTC(cfg)
    jl = _jl_model(P)
    with pytest.raises(ValueError, match="group"):
        jl.load_hf_state_dict(hf.state_dict())                       # detected from the feature-extractor keys
    sd_no_fe = {k: v for k, v in hf.state_dict().items() if "feature_extractor" not in k}
    with pytest.raises(ValueError, match="do_stable_layer_norm"):
        jl.load_hf_state_dict(sd_no_fe, hf_config=cfg)               # detected from the config
    with pytest.raises(ValueError, match="do_stable_layer_norm"):
        jl.load_hf_state_dict(sd_no_fe, hf_config={"do_stable_layer_norm": False})
