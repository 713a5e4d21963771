"""AdapterFusion-style AttAdapter over the K source-dialect adapters of a slot (oracle: ``oracle.encoder.fusion_adapter``).
Kernel-level: the fusion combine kernels vs a torch fp32 restatement.  Model-level: logits, loss and every gradient (source
adapters, fusion LayerNorm / query / key projections, lm_head) vs the oracle, padded and packed layouts, and the knowledge-transfer
set-up — source adapters frozen, only the fusion trained — through the captured-graph trainer."""
import pytest
import torch

from helpers import assert_grads_match, pkg, rel_err, round_bf16_, synth_wave

pytestmark = pytest.mark.gpu
BF16, F32, I32 = torch.bfloat16, torch.float32, torch.int32


@pytest.mark.parametrize("rows,d,b,kk", [(1000, 768, 64, 4), (77, 128, 64, 1), (513, 1024, 64, 8), (250, 256, 32, 3)])
def test_fusion_combine_kernels_match_torch(rows, d, b, kk):
    P = pkg()
    ops = P.ops
    g = torch.Generator(device="cuda").manual_seed(rows + kk)
    h = torch.randn((rows, d), device="cuda", generator=g).to(BF16)
    y = (torch.randn((kk, rows, d), device="cuda", generator=g) * 0.5).to(BF16)
    q = torch.randn((rows, b), device="cuda", generator=g).to(BF16)
    key = torch.randn((kk, rows, b), device="cuda", generator=g).to(BF16)
    dout = (torch.randn((rows, d), device="cuda", generator=g) * 0.1).to(BF16)
    scale = b ** -0.5
    out, alpha = ops.fusion_combine_fwd(h, y, q, key, scale)
    dy, dq, dkey = ops.fusion_combine_bwd(dout, y, q, key, alpha, scale)
    torch.cuda.synchronize()
    hf, yf, qf, kf = h.float().requires_grad_(True), y.float().requires_grad_(True), q.float().requires_grad_(True), key.float().requires_grad_(True)
    s = (qf.unsqueeze(0) * kf).sum(-1) * scale
    a = torch.softmax(s, dim=0)
    ref = hf + (a.unsqueeze(-1) * yf).sum(0)
    assert rel_err(alpha, a.t()) < 1e-5
    assert rel_err(out.float(), ref) < 4e-3                       # one bf16 rounding of the output
    # reference gradients: dy_k here is only the direct term α_k · dout (the key path is added by a GEMM in the engine)
    (ref * dout.float()).sum().backward()
    assert rel_err(dq.float(), qf.grad) < 1e-2
    assert rel_err(dkey.float(), kf.grad) < 1e-2
    assert rel_err(dy.float(), a.detach().unsqueeze(-1) * dout.float().unsqueeze(0)) < 4e-3


def test_fusion_combine_zeroes_padded_rows():
    P = pkg()
    ops = P.ops
    rows, d, b, kk, seq = 60, 128, 64, 2, 20
    g = torch.Generator(device="cuda").manual_seed(1)
    h = torch.randn((rows, d), device="cuda", generator=g).to(BF16)
    y = torch.randn((kk, rows, d), device="cuda", generator=g).to(BF16)
    q = torch.randn((rows, b), device="cuda", generator=g).to(BF16)
    key = torch.randn((kk, rows, b), device="cuda", generator=g).to(BF16)
    lengths = torch.tensor([20, 7, 0], dtype=I32, device="cuda")
    out, _ = ops.fusion_combine_fwd(h, y, q, key, 0.125, row_lengths=lengths, rows_per_seq=seq)
    full, _ = ops.fusion_combine_fwd(h, y, q, key, 0.125)
    torch.cuda.synchronize()
    o3, f3 = out.view(3, seq, d), full.view(3, seq, d)
    assert torch.equal(o3[0], f3[0]) and torch.equal(o3[1, :7], f3[1, :7])
    assert float(o3[1, 7:].abs().max()) == 0.0 and float(o3[2].abs().max()) == 0.0


def _setup(P, slots=(None, "fuse"), kk=3):
    cfg = P.JLConfig(hidden_size=128, num_hidden_layers=3, num_attention_heads=2, intermediate_size=256, conv_channels=64, vocab_size=48,
                     adapter_attn=slots[0], adapter_ffn=slots[1], wf_bottleneck=32, wf_rank=8, num_dialects=kk)
    model = P.JLForCTC(cfg)
    with torch.no_grad():                 # make the fusion matter: the N(0, 0.02) init leaves α ≈ uniform and the updates tiny
        for layer in model.encoder.layers:
            for ad in (layer.adapter_attn, layer.adapter_ffn):
                if ad is not None and ad.kind == "fuse":
                    ad.source.up_A.mul_(16.0)
                    ad.source.down_A.mul_(8.0)
                    ad.source.down_B.mul_(4.0)
                    ad.source.down_bias.add_(0.1)
                    ad.q_proj.weight.mul_(10.0)
                    ad.k_proj.weight.mul_(10.0)
    round_bf16_(model)
    model = model.cuda()
    model.freeze_base_model()
    return cfg, model


def _labels(lens, vocab, seed):
    g = torch.Generator().manual_seed(seed)
    smax = max(1, max(int(0.4 * t) for t in lens))
    lab = torch.full((len(lens), smax), -100, dtype=torch.int64)
    for i, t in enumerate(lens):
        lab[i, : int(0.4 * t)] = torch.randint(1, vocab, (int(0.4 * t),), generator=g)
    return lab


@pytest.mark.parametrize("slots,packed", [((None, "fuse"), False), ((None, "fuse"), True), (("fuse", "att"), False)])
def test_fusion_adapter_model_vs_oracle(slots, packed):
    from oracle import model as om
    P = pkg()
    cfg, model = _setup(P, slots)
    waves = [synth_wave(24000, 1), synth_wave(17321, 2), synth_wave(9000, 3)]
    fe = P.JLFeatureExtractor(device="cuda")
    feats = fe([w.numpy() for w in waves], sampling_rate=16000)
    lens = model.output_lengths(feats["input_features"], frame_lengths=feats["frame_lengths"]).cpu().tolist()
    labels = _labels(lens, cfg.vocab_size, 4)
    loss, logits = model(feats["input_features"], attention_mask=feats["attention_mask"], labels=labels.cuda(), packed=packed)
    loss.backward()
    torch.cuda.synchronize()
    if packed:
        logits = model.unpack_logits(logits, model.last_packed)
    w = om.from_product_state_dict(model.state_dict())
    ocfg = om.OracleConfig(**{k: v for k, v in cfg.to_dict().items() if k in om.OracleConfig.__dataclass_fields__})
    for k, v in w.items():
        v.requires_grad_(om.is_trainable(k))
    oloss, ologits, olens = om.forward_from_waveforms(w, ocfg, waves, labels)
    oloss.backward()
    assert olens.tolist() == lens
    for i, t in enumerate(lens):
        assert rel_err(logits[i, :t].float(), ologits[i, :t]) < 2e-2, f"logits utt {i}"
    assert abs(float(loss) - float(oloss)) <= 2e-3 * abs(float(oloss))        # amplified adapters: twice SURVEY §8d's 1e-3
    # 5e-2 instead of 3e-2: the adapters are amplified ~500x here so that the fusion weights are far from uniform (and the test
    # sensitive to them); the ReLU-path rule of helpers.grad_tolerance applies to the source sets, with 1.2e-1 (measured 8.4e-2:
    # the amplified, biased pre-activations flip more masks than the N(0, 0.02) init does)
    assert_grads_match(model, lambda name: w[name[len("encoder."):] if name.startswith("encoder.") else name].grad, 5e-2, relu_path=1.2e-1)


def test_fusion_knowledge_transfer_trainer_frozen_sources():
    """The transfer set-up: source-dialect adapters frozen, fusion (LayerNorm, query, key) + lm_head trained, through the
    captured-graph trainer (flat bucket without the source sets).  Gradients equal the module path's; the sources do not move."""
    P = pkg()
    cfg, model = _setup(P, (None, "fuse"), kk=2)
    for layer in model.encoder.layers:
        layer.adapter_ffn.source.requires_grad_(False)
    n = 24000
    wave = torch.stack([synth_wave(n, 11), synth_wave(n, 12)])
    ns = torch.tensor([n, 15000], dtype=I32)
    fe = P.JLFeatureExtractor(device="cuda")
    feats = fe([wave[0].numpy(), wave[1, :15000].numpy()], sampling_rate=16000)
    lens = model.output_lengths(feats["input_features"], frame_lengths=feats["frame_lengths"]).cpu().tolist()
    labels = _labels(lens, cfg.vocab_size, 6)
    loss, _ = model(feats["input_features"], labels=labels.cuda(), frame_lengths=feats["frame_lengths"])
    loss.backward()
    ref = {k: p.grad.clone() for k, p in model._get_adapters().items() if p.requires_grad}
    assert all(".source." not in k for k in ref) and any(".q_proj." in k for k in ref)
    src_before = model.encoder.layers[0].adapter_ffn.source.up_A.detach().clone()
    tr = P.AdapterTrainer(model, lr=1e-2, weight_decay=0.0, comm=None)
    assert tr.flat.num_params == sum(p.numel() for p in model._get_adapters().values() if p.requires_grad)
    l1 = float(tr.step(wave.pin_memory(), ns, labels.to(I32)).item())
    torch.cuda.synchronize()
    assert abs(l1 - float(loss)) <= 1e-4 * abs(float(loss))
    for k, p in model._get_adapters().items():
        if p.requires_grad:
            assert rel_err(tr.flat.out(p), ref[k]) < 1e-3 or float((tr.flat.out(p) - ref[k]).abs().max()) < 1e-6, k
    l2 = float(tr.step(wave.pin_memory(), ns, labels.to(I32)).item())
    l3 = float(tr.step(wave.pin_memory(), ns, labels.to(I32)).item())
    assert l3 < l1
    assert torch.equal(src_before, model.encoder.layers[0].adapter_ffn.source.up_A)
