"""GPU parity of the raw-waveform (wav2vec2 / XLS-R) front end through the C ABI: operand-building kernels
against torch restatements, the whole model (front end + encoder + adapters + CTC) against the fp32 oracle, a real HF
``Wav2Vec2ForCTC`` checkpoint end to end, and the graph-captured trainer / transcriber."""
import pytest
import torch
import torch.nn.functional as F

from helpers import pkg, rel_err, round_bf16_, synth_wave

pytestmark = pytest.mark.gpu
BF16, F32, I32 = torch.bfloat16, torch.float32, torch.int32


def test_wave_stats_and_layer0_im2col():
    from oracle import w2v_frontend as wf
    ops = pkg().ops
    waves = [synth_wave(16000, 1) * 3.0 + 0.2, synth_wave(9001, 2), synth_wave(405, 3)]
    norm, ns = wf.normalize(waves)
    n = (norm.shape[1] + 3) // 4 * 4
    raw = torch.zeros((3, n))
    for i, w_ in enumerate(waves):
        raw[i, : ns[i]] = w_
        raw[i, ns[i]:] = 7.0                                   # garbage in the padding must not matter
    raw, nsd = raw.cuda(), torch.tensor(ns, dtype=I32).cuda()
    stats = ops.wave_stats(raw, nsd)
    for i, w_ in enumerate(waves):
        assert abs(float(stats[i, 0]) - float(w_.mean())) < 1e-6
        assert abs(float(stats[i, 1]) - float(1.0 / torch.sqrt(w_.var(unbiased=False) + 1e-7))) < 1e-4 * float(stats[i, 1])
    t_out = (n - 10) // 5 + 1
    col = ops.wave_im2col(raw, nsd, stats, t_out, 10, 5).float().cpu().view(3, t_out, 16)
    padded = torch.zeros((3, n))
    padded[:, : norm.shape[1]] = norm
    ref = padded.unfold(1, 10, 5)                              # [3, t_out, 10]
    assert torch.equal(col[:, :, 10:], torch.zeros_like(col[:, :, 10:]))
    assert float((col[:, :, :10] - ref.to(BF16).float()).abs().max()) <= 2e-2    # one bf16 ulp at |x| <= 4


@pytest.mark.parametrize("k,s,pad,c0,cg", [(3, 2, 0, 0, 64), (2, 2, 0, 0, 64), (16, 1, 8, 32, 32), (128, 1, 64, 48, 48)])
def test_im2col_1d_matches_unfold(k, s, pad, c0, cg):
    ops = pkg().ops
    g = torch.Generator().manual_seed(k)
    c = 64 if c0 + cg <= 64 else 96
    x = torch.randn(2, 37, c, generator=g).to(BF16)
    t_out = (37 + 2 * pad - k) // s + 1
    if pad:
        t_out = 37                                             # "same" convolution with the last frame dropped (even kernel)
    out = ops.im2col_1d(x.cuda(), t_out, k, s, pad=pad, c0=c0, cg=cg).cpu().view(2, t_out, k, cg)
    xp = F.pad(x[:, :, c0:c0 + cg].float(), (0, 0, pad, pad))
    ref = xp.unfold(1, k, s)[:, :t_out].permute(0, 1, 3, 2)    # [B, t_out, k, cg]
    assert torch.equal(out.float(), ref)


def test_layernorm_gelu():
    ops = pkg().ops
    g = torch.Generator().manual_seed(0)
    x = torch.randn(1000, 512, generator=g).to(BF16)
    gamma, beta = torch.rand(512, generator=g) + 0.5, torch.randn(512, generator=g) * 0.1
    y, _, _ = ops.layernorm_fwd(x.cuda(), gamma.cuda(), beta.cuda(), 1e-5, gelu=True)
    ref = F.gelu(F.layer_norm(x.float(), (512,), gamma, beta, 1e-5))
    assert rel_err(y.float(), ref) < 5e-3
    xin = x.cuda().clone()
    y2, _, _ = ops.layernorm_fwd(xin, gamma.cuda(), beta.cuda(), 1e-5, gelu=True, out=xin)      # in place
    assert torch.equal(y2, y)


def _cfg(P, **kw):
    base = dict(hidden_size=128, num_hidden_layers=2, num_attention_heads=2, intermediate_size=256, vocab_size=40, front_end="wav2vec2",
                conv_dim=64, num_conv_pos_embeddings=16, num_conv_pos_embedding_groups=4, wf_bottleneck=32, wf_rank=8)
    base.update(kw)
    return P.JLConfig(**base)


def _oracle(model, cfg):
    from oracle import model as om
    w = om.from_product_state_dict(model.state_dict())
    ocfg = om.OracleConfig(**{k: v for k, v in cfg.to_dict().items() if k in om.OracleConfig.__dataclass_fields__})
    return om, w, ocfg


def test_wav2vec2_model_logits_loss_and_adapter_grads_match_oracle():
    P = pkg()
    cfg = _cfg(P, adapter_attn="att", adapter_ffn="wf")
    model = P.JLForCTC(cfg)
    round_bf16_(model)
    model = model.cuda()
    model.freeze_base_model()
    waves = [synth_wave(16000, 41), synth_wave(12345, 42), synth_wave(6000, 43)]
    fe = P.JLWaveformFeatureExtractor(device="cuda")
    enc = fe([w_.numpy() for w_ in waves], sampling_rate=16000)
    lens = model.output_lengths(enc["input_values"], attention_mask=enc["attention_mask"]).cpu().tolist()
    labels = torch.full((3, 8), -100, dtype=torch.int64)
    g = torch.Generator().manual_seed(6)
    for i, t in enumerate(lens):
        s = min(8, int(0.3 * t))
        labels[i, :s] = torch.randint(1, cfg.vocab_size, (s,), generator=g)
    loss, logits = model(input_values=enc["input_values"], attention_mask=enc["attention_mask"], labels=labels.cuda())
    loss.backward()
    torch.cuda.synchronize()
    om, w, ocfg = _oracle(model, cfg)
    for k, v in w.items():
        v.requires_grad_(om.is_trainable(k))
    oloss, ologits, olens = om.forward_from_waveforms(w, ocfg, waves, labels)
    oloss.backward()
    assert olens.tolist() == lens
    for i, t in enumerate(lens):
        # seven bf16 conv + LayerNorm layers in front of the encoder: 3e-2 (SURVEY §8d allows 3e-2 for the deep stack)
        assert rel_err(logits[i, :t].float(), ologits[i, :t]) < 3e-2, f"logits utt {i}"
    assert abs(float(loss) - float(oloss)) <= 1e-2 * abs(float(oloss))
    for name, p in model._get_adapters().items():
        ref = w[name[len("encoder."):] if name.startswith("encoder.") else name].grad
        err = float((p.grad.float().cpu() - ref).norm())
        if name.endswith("k_proj.bias"):
            # analytically zero (a key bias shifts every score of a query equally): both sides hold rounding noise only — compare
            # it with the size of the query-bias gradient of the same adapter
            qn = float(dict(model._get_adapters())[name.replace("k_proj", "q_proj")].grad.float().norm())
            assert err <= 5e-2 * qn, f"grad {name}: {err} vs q_proj.bias grad norm {qn}"
            continue
        assert err <= 6e-2 * float(ref.norm()) + 2e-6 * ref.numel() ** 0.5, f"grad {name}: err {err} ref norm {float(ref.norm())}"
    # raw and already-normalised input give the same logits (the utterance normalisation is idempotent)
    from oracle import w2v_frontend as wf
    norm, ns = wf.normalize(waves)
    x = torch.zeros_like(enc["input_values"])
    x[:, : norm.shape[1]] = norm.cuda()
    with torch.no_grad():
        _, logits_n = model(input_values=x, attention_mask=enc["attention_mask"])
    for i, t in enumerate(lens):
        assert rel_err(logits_n[i, :t].float(), logits[i, :t].float()) < 1e-2


def test_hf_wav2vec2_checkpoint_end_to_end_on_gpu():
    """A real HF ``Wav2Vec2ForCTC`` (XLS-R-style front end + per-language adapter) loaded by its HF names: product logits
    vs HF's own forward on the CPU."""
    from transformers import Wav2Vec2FeatureExtractor
    from test_oracle_w2v import _hf, _jl
    P = pkg()
    hf = _hf()
    round_bf16_(hf)
    jl, cfg = _jl(P)
    missing, _ = jl.load_hf_state_dict(hf.state_dict(), strict=True)
    assert missing == []
    jl = jl.cuda().eval()
    waves = [synth_wave(16000, 51), synth_wave(9500, 52)]
    hfe = Wav2Vec2FeatureExtractor(feature_size=1, sampling_rate=16000, padding_value=0.0, do_normalize=True, return_attention_mask=True)
    enc = hfe([x.numpy() for x in waves], sampling_rate=16000, padding=True, return_tensors="pt")
    with torch.no_grad():
        ref = hf(enc["input_values"], attention_mask=enc["attention_mask"]).logits
        mine = P.JLWaveformFeatureExtractor(device="cuda")([x.numpy() for x in waves], sampling_rate=16000)
        _, logits = jl(input_values=mine["input_values"], attention_mask=mine["attention_mask"])
        lens = jl.output_lengths(mine["input_values"], attention_mask=mine["attention_mask"]).cpu().tolist()
    assert lens == hf._get_feat_extract_output_lengths(enc["attention_mask"].sum(-1)).tolist()
    for i, t in enumerate(lens):
        assert rel_err(logits[i, :t].float(), ref[i, :t]) < 3e-2, f"utt {i}"
        agree = float((logits[i, :t].argmax(-1).cpu() == ref[i, :t].argmax(-1)).float().mean())
        assert agree >= 0.9, (i, agree)


def test_wav2vec2_trainer_and_transcriber_graphs():
    P = pkg()
    cfg = _cfg(P, adapter_ffn="att")
    model = P.JLForCTC(cfg)
    round_bf16_(model)
    model = model.cuda()
    model.freeze_base_model()
    n = 16000
    waves = [synth_wave(n, 61), synth_wave(12000, 62)]
    wave = torch.zeros((2, n))
    for i, w_ in enumerate(waves):
        wave[i, : w_.shape[0]] = w_
    ns = torch.tensor([n, 12000], dtype=I32)
    labels = torch.full((2, 6), -100, dtype=I32)
    labels[0, :6] = torch.tensor([3, 4, 4, 9, 1, 2], dtype=I32)
    labels[1, :3] = torch.tensor([7, 7, 5], dtype=I32)
    # autograd path
    loss, logits = model(input_values=wave.cuda(), frame_lengths=ns.cuda(), labels=labels.cuda().long())
    loss.backward()
    ref = {k: p.grad.detach().clone() for k, p in model._get_adapters().items()}
    ids_ref = model.greedy_decode(logits.detach(), model.output_lengths(wave.cuda(), frame_lengths=ns.cuda()))
    # trainer (flat bucket + CUDA graph)
    tr = P.AdapterTrainer(model, lr=0.0, weight_decay=0.0, use_cuda_graph=True, comm=None)
    tl = tr.step(wave.pin_memory(), ns, labels).item()
    torch.cuda.synchronize()
    assert abs(tl - float(loss)) <= 1e-3 * abs(float(loss))
    for k, p in model._get_adapters().items():
        got = tr.flat.out(p)
        assert rel_err(got, ref[k]) < 1e-2 or float((got - ref[k]).abs().max()) < 1e-5, k
    # transcriber (CUDA graph): same ids as the module path
    ts = P.Transcriber(model)
    ids, cnt = ts(wave.pin_memory(), ns)
    ids, cnt = ids.cpu(), cnt.cpu()
    assert [ids[i, : int(cnt[i])].tolist() for i in range(2)] == ids_ref
