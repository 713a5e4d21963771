"""End-to-end GPU parity: waveform → features → encoder (+ adapters) → logits → CTC loss → adapter-only gradients,
product modules (through the C ABI) vs the fp32 CPU oracle on bf16-rounded weights."""
import math

import pytest
import torch

from helpers import assert_grads_match, pkg, rel_err, round_bf16_, synth_wave

pytestmark = pytest.mark.gpu
I32 = torch.int32


def _oracle_setup(model, cfg):
    from oracle import model as om
    w = om.from_product_state_dict(model.state_dict())
    ocfg = om.OracleConfig(**{k: v for k, v in cfg.to_dict().items() if k in om.OracleConfig.__dataclass_fields__})
    return om, w, ocfg


def _labels(lengths, vocab, smax, seed):
    g = torch.Generator().manual_seed(seed)
    lab = torch.full((len(lengths), smax), -100, dtype=torch.int64)
    for i, t in enumerate(lengths):
        s = min(smax, int(0.4 * int(t)))
        lab[i, :s] = torch.randint(1, vocab, (s,), generator=g)
    return lab


def _small_cfg(P, **kw):
    base = dict(hidden_size=128, num_hidden_layers=3, num_attention_heads=2, intermediate_size=256, conv_channels=64, vocab_size=48,
                wf_bottleneck=32, wf_rank=8)
    base.update(kw)
    return P.JLConfig(**base)


@pytest.mark.parametrize("slots", [(None, None), (None, "wf"), (None, "att"), ("att", "wf"), ("wf", "att")])
def test_small_model_logits_loss_and_adapter_grads(slots):
    P = pkg()
    cfg = _small_cfg(P, adapter_attn=slots[0], adapter_ffn=slots[1])
    model = P.JLForCTC(cfg)
    round_bf16_(model)
    model = model.cuda()
    model.freeze_base_model()
    waves = [synth_wave(24000, 1), synth_wave(17321, 2), synth_wave(9000, 3)]
    fe = P.JLFeatureExtractor(device="cuda")
    feats = fe([w.numpy() for w in waves], sampling_rate=16000)
    lens = model.output_lengths(feats["input_features"], frame_lengths=feats["frame_lengths"]).cpu().tolist()
    labels = _labels(lens, cfg.vocab_size, 10, seed=4)
    loss, logits = model(feats["input_features"], attention_mask=feats["attention_mask"], labels=labels.cuda())
    loss.backward()
    torch.cuda.synchronize()
    om, w, ocfg = _oracle_setup(model, cfg)
    for k, v in w.items():
        v.requires_grad_(om.is_trainable(k))
    oloss, ologits, olens = om.forward_from_waveforms(w, ocfg, waves, labels)
    oloss.backward()
    assert olens.tolist() == lens
    for i, t in enumerate(lens):
        assert rel_err(logits[i, :t].float(), ologits[i, :t]) < 2e-2, f"logits utt {i}"
    assert abs(float(loss) - float(oloss)) <= 1e-3 * abs(float(oloss))          # CTC loss <= 1e-3 relative
    # adapter gradients <= 3e-2 relative Frobenius (8e-2 upstream of the WFAdapter's ReLU — helpers.WF_RELU_PATH —
    # and the analytically-zero AttAdapter key-bias gradient measured against the query-bias gradient)
    assert_grads_match(model, lambda name: w[name[len("encoder."):] if name.startswith("encoder.") else name].grad, 3e-2)


def test_base_config_forward_logits_and_greedy_ids():
    """BASELINE.json config 1: 12-layer d=768 encoder with WFAdapter, batch 4 × 10 s, greedy decode."""
    P = pkg()
    cfg = P.JLConfig.base(adapter_ffn="wf")
    model = P.JLForCTC(cfg)
    round_bf16_(model)
    model = model.cuda().eval()
    waves = [synth_wave(160000, 1234 + i) for i in range(4)]
    fe = P.JLFeatureExtractor(device="cuda")
    feats = fe([w.numpy() for w in waves], sampling_rate=16000)
    with torch.no_grad():
        _, logits = model(feats["input_features"], attention_mask=feats["attention_mask"])
    lens = model.output_lengths(feats["input_features"], feats["attention_mask"])
    torch.cuda.synchronize()
    assert logits.shape == (4, 250, 5000)
    om, w, ocfg = _oracle_setup(model, cfg)
    with torch.no_grad():
        _, ologits, olens = om.forward_from_waveforms(w, ocfg, waves)
    fro = rel_err(logits.float(), ologits)
    assert fro < 2e-2, f"relative Frobenius error {fro}"                       # stated bf16 tolerance, 12 layers
    assert float((logits.float().cpu() - ologits).abs().max()) <= 5e-2 * float(ologits.abs().max())
    # frame-argmax agreement: random-init logits are nearly flat, so flips are allowed only where the oracle's own margin
    # between the two candidates is inside the numerical error band
    # SURVEY §8d asks for >= 99 % frame-argmax agreement.  Random-init logits are nearly flat (top-2 margins of the order of the
    # bf16 error), so the 99 % is asserted on the frames whose oracle top-2 margin exceeds the numerical error band — there the
    # argmax must agree — and every flip must lie inside the band; the overall figure (measured 0.97-0.99) is checked loosely.
    mine, theirs = logits.float().cpu().argmax(-1), ologits.argmax(-1)
    agree = (mine == theirs).float().mean()
    assert float(agree) >= 0.95, float(agree)
    gap = ologits.gather(-1, theirs[..., None]) - ologits.gather(-1, mine[..., None])
    band = 2.0 * float((logits.float().cpu() - ologits).abs().max())
    assert float(gap.max()) <= band, (float(gap.max()), band)
    top2 = ologits.topk(2, dim=-1).values
    clear = (top2[..., 0] - top2[..., 1]) > band
    assert float((mine == theirs)[clear].float().mean()) >= 0.99
    # greedy ids are bit-exact when decoded from the same logits
    from oracle import ctc as oc
    assert model.greedy_decode(logits, lens) == oc.greedy_decode(logits.float().cpu(), lens.cpu(), 0)


def test_label_out_of_vocab_raises_and_inference_returns_no_loss():
    P = pkg()
    cfg = _small_cfg(P, adapter_ffn="wf")
    model = P.JLForCTC(cfg).cuda()
    fe = P.JLFeatureExtractor(device="cuda")
    feats = fe([synth_wave(8000, 1).numpy()], sampling_rate=16000)
    loss, logits = model(feats["input_features"], feats["attention_mask"])
    assert loss is None and logits.shape[-1] == cfg.vocab_size
    with pytest.raises(ValueError):
        model(feats["input_features"], feats["attention_mask"], labels=torch.tensor([[cfg.vocab_size]]).cuda())


def test_trainer_step_matches_autograd_path_and_updates_adapters():
    """Flat-bucket fast path (CUDA graph) == autograd path: same loss, same gradients; AdamW moves only adapters + lm_head."""
    P = pkg()
    cfg = _small_cfg(P, adapter_attn="att", adapter_ffn="wf")
    torch.manual_seed(0)
    m1 = P.JLForCTC(cfg)
    round_bf16_(m1)
    m2 = P.JLForCTC(cfg)
    m2.load_state_dict(m1.state_dict())
    m1, m2 = m1.cuda(), m2.cuda()
    m1.freeze_base_model()
    m2.freeze_base_model()
    n = 16000
    wave = torch.stack([synth_wave(n, 1), synth_wave(n, 2)])
    ns = torch.tensor([n, 12000], dtype=I32)
    fe = P.JLFeatureExtractor(device="cuda")
    feats = fe([wave[0].numpy(), wave[1, :12000].numpy()], sampling_rate=16000)
    lens = m1.output_lengths(feats["input_features"], frame_lengths=feats["frame_lengths"]).cpu().tolist()
    labels = _labels(lens, cfg.vocab_size, 8, seed=9)
    loss_a, _ = m1(feats["input_features"], labels=labels.cuda(), frame_lengths=feats["frame_lengths"])
    loss_a.backward()
    ref_grads = {k: p.grad.clone() for k, p in m1._get_adapters().items()}
    backbone_before = m2.encoder.layers[0].attention.q_proj.weight.clone()
    tr = P.AdapterTrainer(m2, lr=1e-3, use_cuda_graph=True)
    before = tr.flat.param.clone()
    loss_b = tr.step(wave.pin_memory(), ns, labels.to(I32)).item()      # the returned tensor is the graph's static output
    torch.cuda.synchronize()
    assert abs(float(loss_a) - loss_b) <= 1e-5 * abs(float(loss_a))
    for k, p in m2._get_adapters().items():
        assert rel_err(p.grad, ref_grads[k]) < 1e-5, k
    assert not torch.equal(before, tr.flat.param)
    assert torch.equal(backbone_before, m2.encoder.layers[0].attention.q_proj.weight)
    loss_c = tr.step(wave.pin_memory(), ns, labels.to(I32)).item()     # graph replay with refreshed bf16 shadows
    torch.cuda.synchronize()
    assert loss_c < loss_b
    assert tr.launches_per_step > 50
    # prefetching path: submit() stages the next batch on a copy stream, step_async() returns a handle whose item() waits for that
    # step's 4-byte device → host copy only; the losses read late (after the next launch) are the losses of their own steps
    wp, lp = wave.pin_memory(), labels.to(I32).pin_memory()
    tr.submit(wp, ns, lp)
    handles = []
    for i in range(3):
        hd = tr.step_async()
        if i < 2:
            tr.submit(wp, ns, lp)
        handles.append(hd)
    lagged = [hd.item() for hd in handles]
    assert lagged[0] < loss_c and lagged[1] < lagged[0] and lagged[2] < lagged[1], (loss_c, lagged)
    ref = tr.step(wp, ns, lp).item()
    assert ref < lagged[2]


@pytest.mark.parametrize("fused_wf_train", [False, True])
def test_large_config_both_adapters_mixed_lengths(fused_wf_train):
    """``fused_wf_train``: the WFAdapter forward of the training pass as LN + 4 GEMMs (default) or as the one fused kernel that also
    writes the intermediates (``JLEngine.fused_wf_train``).
    BASELINE.json configs 3 + 4 in miniature: d=1024 / 16 heads / 4096 FFN transformer stack (4 of the 24 layers to keep
    the CPU oracle quick) with AttAdapter after attention and WFAdapter after the FFN, mixed-length utterances (padded and
    masked), CTC 'mean' reduction: logits, loss and adapter gradients vs the oracle."""
    P = pkg()
    cfg = P.JLConfig.large(num_hidden_layers=4, adapter_attn="att", adapter_ffn="wf", vocab_size=600, ctc_loss_reduction="mean")
    model = P.JLForCTC(cfg)
    round_bf16_(model)
    model = model.cuda()
    model.freeze_base_model()
    model.encoder.engine(model.lm_head).fused_wf_train = fused_wf_train
    waves = [synth_wave(16000 * 3 + 123, 21), synth_wave(16000 * 2, 22), synth_wave(9000, 23), synth_wave(16000 * 3 + 123, 24)]
    fe = P.JLFeatureExtractor(device="cuda")
    feats = fe([w.numpy() for w in waves], sampling_rate=16000)
    lens = model.output_lengths(feats["input_features"], feats["attention_mask"]).cpu().tolist()
    labels = _labels(lens, cfg.vocab_size, 24, seed=5)
    loss, logits = model(feats["input_features"], attention_mask=feats["attention_mask"], labels=labels.cuda())
    loss.backward()
    torch.cuda.synchronize()
    om, w, ocfg = _oracle_setup(model, cfg)
    for k, v in w.items():
        v.requires_grad_(om.is_trainable(k))
    oloss, ologits, olens = om.forward_from_waveforms(w, ocfg, waves, labels)
    oloss.backward()
    assert olens.tolist() == lens
    for i, t in enumerate(lens):
        assert rel_err(logits[i, :t].float(), ologits[i, :t]) < 2e-2, f"logits utt {i}"
    assert abs(float(loss) - float(oloss)) <= 1e-3 * abs(float(oloss))
    assert_grads_match(model, lambda name: w[name[len("encoder."):] if name.startswith("encoder.") else name].grad, 3e-2)


def test_transcriber_matches_module_path_and_oracle_ids():
    """Inference fast path (CUDA graph: waveform → mel → encoder → greedy ids) == module path; ids bit-exact vs the oracle's
    greedy decode of the same logits."""
    from oracle import ctc as oc
    P = pkg()
    cfg = _small_cfg(P, adapter_ffn="wf")
    model = P.JLForCTC(cfg)
    round_bf16_(model)
    model = model.cuda().eval()
    n = 20000
    wave = torch.stack([synth_wave(n, 31), synth_wave(n, 32), synth_wave(n, 33)])
    ns = torch.tensor([n, 15000, 8000], dtype=I32)
    tr = P.Transcriber(model, use_cuda_graph=True)
    for _ in range(2):                                       # second call replays the graph
        ids, nid = tr(wave.pin_memory(), ns)
        torch.cuda.synchronize()
    fe = P.JLFeatureExtractor(device="cuda")
    feats = fe([wave[i, : int(ns[i])].numpy() for i in range(3)], sampling_rate=16000)
    with torch.no_grad():
        _, logits = model(feats["input_features"], attention_mask=feats["attention_mask"])
    lens = model.output_lengths(feats["input_features"], feats["attention_mask"])
    ref = oc.greedy_decode(logits.float().cpu(), lens.cpu(), cfg.pad_token_id)
    got = [ids[i, : int(nid[i])].cpu().tolist() for i in range(3)]
    assert got == ref


def test_per_utterance_dialect_ids_select_wfadapter_factor_sets():
    """K = 3 WFAdapter factor sets, ``dialect_ids`` per utterance (runs [2, 2, 0, 0]); logits, loss and
    per-dialect adapter gradients vs the oracle; the factor set with no utterance in the batch gets exactly zero gradient;
    the inference path (fused one-kernel adapter per run) agrees with the training path; the trainer's graph agrees too."""
    P = pkg()
    cfg = _small_cfg(P, adapter_attn="att", adapter_ffn="wf", num_dialects=3, wf_bottleneck=64, wf_rank=16)
    model = P.JLForCTC(cfg)
    with torch.no_grad():                       # make the adapters matter: the N(0, 0.02) init leaves them near-identity
        for layer in model.encoder.layers:
            layer.adapter_ffn.up_A.mul_(24.0)
            layer.adapter_ffn.down_A.mul_(8.0)
            layer.adapter_ffn.down_B.mul_(4.0)
    round_bf16_(model)
    model = model.cuda()
    model.freeze_base_model()
    dialects = [2, 2, 0, 0]
    waves = [synth_wave(24000, 11), synth_wave(21000, 12), synth_wave(24000, 13), synth_wave(9000, 14)]
    fe = P.JLFeatureExtractor(device="cuda")
    feats = fe([w.numpy() for w in waves], sampling_rate=16000)
    lens = model.output_lengths(feats["input_features"], frame_lengths=feats["frame_lengths"]).cpu().tolist()
    labels = _labels(lens, cfg.vocab_size, 10, seed=5)
    loss, logits = model(feats["input_features"], attention_mask=feats["attention_mask"], labels=labels.cuda(), dialect=dialects)
    loss.backward()
    torch.cuda.synchronize()
    om, w, ocfg = _oracle_setup(model, cfg)
    for k, v in w.items():
        v.requires_grad_(om.is_trainable(k))
    oloss, ologits, olens = om.forward_from_waveforms(w, ocfg, waves, labels, dialect=dialects)
    oloss.backward()
    for i, t in enumerate(lens):
        assert rel_err(logits[i, :t].float(), ologits[i, :t]) < 2e-2, f"logits utt {i}"
    assert abs(float(loss) - float(oloss)) <= 2e-3 * abs(float(oloss))      # amplified adapters (below): twice the 1e-3 of SURVEY §8d
    for name, p in model._get_adapters().items():
        ref = w[name[len("encoder."):] if name.startswith("encoder.") else name].grad
        err = float((p.grad.float().cpu() - ref).norm())
        # 8e-2 instead of SURVEY §8d's 3e-2, for this test only: the adapters are amplified ~770x here, and with them the bf16 rounding of everything that flows through them
        assert err <= 8e-2 * float(ref.norm()) + 2e-6 * ref.numel() ** 0.5, f"grad {name}: err {err} ref norm {float(ref.norm())}"
        if ".adapter_ffn." in name and p.dim() >= 2 and p.shape[0] == 3 and "norm" not in name:
            assert float(p.grad[1].abs().max()) == 0.0, f"{name}: unused dialect 1 must have zero gradient"
            assert float(p.grad[0].abs().max()) > 0.0 and float(p.grad[2].abs().max()) > 0.0
    # a different assignment gives different logits (the ids are really used) …
    with torch.no_grad():
        _, logits_inf = model(feats["input_features"], attention_mask=feats["attention_mask"], dialect=dialects)
        _, logits_other = model(feats["input_features"], attention_mask=feats["attention_mask"], dialect=[0, 0, 2, 2])
    for i, t in enumerate(lens):
        # … and the inference path (one fused kernel per run) matches the composed training path
        same = rel_err(logits_inf[i, :t].float(), logits[i, :t].float())
        other = rel_err(logits_other[i, :t].float(), logits[i, :t].float())
        assert same < 1e-2 and other > 3e-2 and other > 5 * same, (same, other)
    # utterances of one dialect must be adjacent
    with pytest.raises(ValueError):
        model(feats["input_features"], attention_mask=feats["attention_mask"], dialect=[0, 2, 0, 2])
    with pytest.raises(ValueError):
        model(feats["input_features"], attention_mask=feats["attention_mask"], dialect=[0, 0, 3, 3])
    # trainer (flat bucket + CUDA graph keyed by the dialect runs): same gradients as the autograd path
    ref_grads = {n: p.grad.detach().clone() for n, p in model._get_adapters().items()}
    n = 24000
    wave = torch.zeros((4, n))
    ns = torch.tensor([w_.shape[0] for w_ in waves], dtype=I32)
    for i, w_ in enumerate(waves):
        wave[i, : w_.shape[0]] = w_
    lab32 = labels.to(I32)
    tr = P.AdapterTrainer(model, lr=0.0, weight_decay=0.0, use_cuda_graph=True, comm=None)
    tl = tr.step(wave.pin_memory(), ns, lab32, dialect=dialects).item()
    torch.cuda.synchronize()
    assert abs(tl - float(loss)) <= 1e-3 * abs(float(loss))
    for name, p in model._get_adapters().items():
        got = tr.flat.out(p)
        assert rel_err(got, ref_grads[name]) < 1e-2 or float((got - ref_grads[name]).abs().max()) < 1e-5, name


def test_batch_composition_invariance_and_zero_frame_utterance():
    """An utterance's logits, its loss term and the adapter gradients do not depend on what else is in the batch: adding an
    utterance that is too short to yield a single frame (< 400 samples → T' = 0, no labels) changes nothing, every output stays
    finite, and the valid rows of the other utterances are bit-identical (each output row's summation order is fixed)."""
    P = pkg()
    cfg = _small_cfg(P, adapter_attn="att", adapter_ffn="wf", ctc_zero_infinity=True)
    model = P.JLForCTC(cfg)
    round_bf16_(model)
    model = model.cuda()
    model.freeze_base_model()
    fe = P.JLFeatureExtractor(device="cuda")
    wa, wb, wc = synth_wave(24000, 21), synth_wave(9000, 22), synth_wave(300, 23)

    def run(waves, labels):
        for p_ in model.parameters():
            p_.grad = None
        feats = fe([w.numpy() for w in waves], sampling_rate=16000)
        lens = model.output_lengths(feats["input_features"], frame_lengths=feats["frame_lengths"]).cpu().tolist()
        loss, logits = model(feats["input_features"], attention_mask=feats["attention_mask"], labels=labels.cuda())
        loss.backward()
        torch.cuda.synchronize()
        return lens, float(loss), logits.detach().clone(), {n: p_.grad.detach().clone() for n, p_ in model._get_adapters().items()}

    lab2 = _labels([37, 14], cfg.vocab_size, 8, seed=9)
    lab3 = torch.cat([lab2, torch.full((1, 8), -100, dtype=torch.int64)], 0)
    lens2, loss2, logits2, grads2 = run([wa, wb], lab2)
    lens3, loss3, logits3, grads3 = run([wa, wb, wc], lab3)
    assert lens3[:2] == lens2 and lens3[2] == 0
    assert torch.isfinite(logits3).all() and math.isfinite(loss3)
    for i, t in enumerate(lens2):
        assert torch.equal(logits3[i, :t], logits2[i, :t]), f"utterance {i}: logits depend on batch composition"
    assert abs(loss3 - loss2) <= 1e-6 * abs(loss2)
    for n in grads2:
        assert torch.isfinite(grads3[n]).all(), n
        assert rel_err(grads3[n], grads2[n]) < 1e-3 or float((grads3[n] - grads2[n]).abs().max()) < 1e-6, n
