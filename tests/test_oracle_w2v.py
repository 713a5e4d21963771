"""CPU pin of the raw-waveform (wav2vec2 / XLS-R) front end of the oracle against the INSTALLED Hugging Face model
(SURVEY §8 f3): a real ``Wav2Vec2ForCTC`` ("layer" feature-extractor norm, conv bias, stable layer norm, per-language
bottleneck adapter) is loaded by its HF names into ``JLForCTC(front_end="wav2vec2")`` and the oracle, run from that state
dict, reproduces HF's logits and CTC loss."""
import pytest
import torch

from helpers import pkg, synth_wave


def _hf(adapter_dim=16):
    from transformers import Wav2Vec2Config, Wav2Vec2ForCTC
    cfg = Wav2Vec2Config(hidden_size=128, num_hidden_layers=2, num_attention_heads=2, intermediate_size=256, conv_dim=(64,) * 7,
                         feat_extract_norm="layer", conv_bias=True, num_conv_pos_embeddings=16, num_conv_pos_embedding_groups=4,
                         do_stable_layer_norm=True, hidden_dropout=0.0, attention_dropout=0.0, activation_dropout=0.0, feat_proj_dropout=0.0,
                         final_dropout=0.0, layerdrop=0.0, apply_spec_augment=False, vocab_size=24, adapter_attn_dim=adapter_dim,
                         ctc_loss_reduction="sum", pad_token_id=0)
    torch.manual_seed(5)
    m = Wav2Vec2ForCTC(cfg).eval()
    with torch.no_grad():
        for layer in m.wav2vec2.encoder.layers:
            layer.adapter_layer.linear_2.weight.normal_(0, 0.2)
        m.lm_head.weight.normal_(0, 0.3)
    return m


def _jl(P):
    cfg = P.JLConfig(hidden_size=128, num_hidden_layers=2, num_attention_heads=2, intermediate_size=256, vocab_size=24, front_end="wav2vec2",
                     conv_dim=64, num_conv_pos_embeddings=16, num_conv_pos_embedding_groups=4, adapter_ffn="wf", wf_bottleneck=16, wf_rank=16)
    return P.JLForCTC(cfg), cfg


def test_oracle_from_hf_checkpoint_reproduces_hf_logits_and_loss():
    from transformers import Wav2Vec2FeatureExtractor
    from oracle import model as om
    P = pkg()
    hf = _hf()
    jl, cfg = _jl(P)
    missing, skipped = jl.load_hf_state_dict(hf.state_dict(), strict=True)
    assert missing == [], missing
    assert skipped == ["wav2vec2.masked_spec_embed"], skipped
    w = om.from_product_state_dict(jl.state_dict())
    ocfg = om.OracleConfig(**{k: v for k, v in cfg.to_dict().items() if k in om.OracleConfig.__dataclass_fields__})
    waves = [synth_wave(16000, 31), synth_wave(11111, 32), synth_wave(7000, 33)]
    labels = torch.full((3, 6), -100, dtype=torch.int64)
    g = torch.Generator().manual_seed(2)
    for i, s in enumerate((6, 4, 3)):
        labels[i, :s] = torch.randint(1, 24, (s,), generator=g)
    fe = Wav2Vec2FeatureExtractor(feature_size=1, sampling_rate=16000, padding_value=0.0, do_normalize=True, return_attention_mask=True)
    enc = fe([x.numpy() for x in waves], sampling_rate=16000, padding=True, return_tensors="pt")
    with torch.no_grad():
        ref = hf(enc["input_values"], attention_mask=enc["attention_mask"], labels=labels)
        oloss, ologits, olens = om.forward_from_waveforms(w, ocfg, waves, labels)
    ref_lens = hf._get_feat_extract_output_lengths(enc["attention_mask"].sum(-1))
    assert olens.tolist() == ref_lens.tolist()
    for i, t in enumerate(olens.tolist()):
        assert torch.allclose(ologits[i, :t], ref.logits[i, :t], atol=1e-4, rtol=1e-4), float((ologits[i, :t] - ref.logits[i, :t]).abs().max())
    assert abs(float(oloss) - float(ref.loss)) <= 1e-4 * abs(float(ref.loss))


def test_oracle_front_end_stages_match_hf_modules():
    from oracle import w2v_frontend as wf
    hf = _hf().wav2vec2
    sd = hf.state_dict()
    P = pkg()
    conv, _ = P.hf_compat.convert_hf_state_dict(sd, with_front_end=True)
    w = {k[len("encoder."):]: v.float() for k, v in conv.items() if k.startswith("encoder.w2v.")}
    g = torch.Generator().manual_seed(3)
    x = torch.randn(2, 4000, generator=g)
    with torch.no_grad():
        feats = hf.feature_extractor(x).transpose(1, 2)
        mine = wf.feature_encoder(w, x, hf.config.conv_kernel, hf.config.conv_stride)
        assert torch.allclose(mine, feats, atol=2e-5), float((mine - feats).abs().max())
        proj, _ = hf.feature_projection(feats)
        assert torch.allclose(wf.feature_projection(w, mine), proj, atol=2e-5)
        pos = hf.encoder.pos_conv_embed(proj)
        assert torch.allclose(wf.pos_conv(w, proj, hf.config.num_conv_pos_embedding_groups), pos, atol=2e-5)
    assert wf.conv_lengths(4000, hf.config.conv_kernel, hf.config.conv_stride) == feats.shape[1]
    norm, ns = wf.normalize([x[0], x[1, :3000]])
    assert ns == [4000, 3000] and abs(float(norm[1, :3000].mean())) < 1e-6 and float(norm[1, 3000:].abs().max()) == 0.0
    assert abs(float(norm[0].var(unbiased=False)) - 1.0) < 1e-4
