"""GPU parity of LayerNorm, attention, CTC, greedy decode and the elementwise helpers (through the C ABI) against the
oracle / plain fp32 math on the same seeded inputs."""
import math
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from helpers import ROOT, pkg, rel_err

pytestmark = pytest.mark.gpu
BF16, F32, I32 = torch.bfloat16, torch.float32, torch.int32


def _g(seed=0):
    return torch.Generator(device="cuda").manual_seed(seed)


# --------------------------------------------------------------------------------------------- LayerNorm
@pytest.mark.parametrize("rows,d", [(1000, 768), (333, 1024), (64, 128), (17, 2048)])
def test_layernorm_fwd_bwd(rows, d):
    ops = pkg().ops
    g = _g(1)
    x = (torch.randn(rows, d, device="cuda", generator=g) * 2 + 0.3).to(BF16)
    gamma = torch.randn(d, device="cuda", generator=g) * 0.2 + 1.0
    beta = torch.randn(d, device="cuda", generator=g) * 0.1
    dy = torch.randn(rows, d, device="cuda", generator=g).to(BF16)
    dres = torch.randn(rows, d, device="cuda", generator=g).to(BF16)
    y, mean, rstd = ops.layernorm_fwd(x, gamma, beta, 1e-5, save_stats=True)
    xr = x.float().requires_grad_(True)
    gr, br = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    yr = F.layer_norm(xr, (d,), gr, br, 1e-5)
    assert rel_err(y.float(), yr) < 6e-3
    assert rel_err(mean, xr.mean(-1)) < 1e-5
    yr.backward(dy.float())
    dx, dgamma, dbeta = ops.layernorm_bwd(dy, x, gamma, mean, rstd, dres=dres, want_wgrad=True)
    assert rel_err(dx.float(), xr.grad + dres.float()) < 8e-3
    assert rel_err(dgamma, gr.grad) < 2e-3
    assert rel_err(dbeta, br.grad) < 2e-3
    dg2, db2 = torch.empty_like(dgamma), torch.empty_like(dbeta)
    ops.layernorm_wgrad(dy, x, mean, rstd, dg2, db2)                 # stand-alone dγ / dβ kernel (side-branch path)
    assert rel_err(dg2, gr.grad) < 2e-3 and rel_err(db2, br.grad) < 2e-3
    dx2, _, _ = ops.layernorm_bwd(dy, x, gamma, mean, rstd)          # frozen norm: dx only
    assert rel_err(dx2.float(), xr.grad) < 8e-3
    torch.cuda.synchronize()


@pytest.mark.parametrize("rows,cols", [(8000, 768), (8000, 192), (8000, 5000), (1, 64), (63, 8), (513, 40), (700, 333), (4097, 1000)])
def test_column_reductions_on_clusters(rows, cols):
    """colsum_kernel / layernorm_wgrad_kernel: clusters of 8 CTAs per 32 columns, DSMEM combine in rank order — odd row counts
    (fewer row groups than CTAs in the cluster), column counts that are not multiples of 32 or of 8 (scalar path), strided
    views, and bit-identical results from run to run."""
    ops = pkg().ops
    g = _g(7)
    x = torch.randn(rows, cols, device="cuda", generator=g).to(BF16)
    ref = x.float().sum(0)
    out1, out2 = ops.colsum(x), ops.colsum(x)
    torch.cuda.synchronize()
    assert torch.equal(out1, out2)
    assert float((out1 - ref).abs().max()) <= 1e-4 * max(1.0, float(ref.abs().max())) + 2e-5 * rows ** 0.5
    if cols % 8 == 0 and cols >= 16:
        view = x[:, 8:cols]                                      # column window of a wider matrix (row stride > columns)
        assert float((ops.colsum(view) - ref[8:]).abs().max()) <= 1e-4 * max(1.0, float(ref.abs().max())) + 2e-5 * rows ** 0.5
    if cols % 8 == 0:
        dy = torch.randn(rows, cols, device="cuda", generator=g).to(BF16)
        mean = x.float().mean(-1).contiguous()
        rstd = (1.0 / torch.sqrt(x.float().var(-1, unbiased=False) + 1e-5)).contiguous()
        dg, db = torch.empty(cols, device="cuda"), torch.empty(cols, device="cuda")
        dg2, db2 = torch.empty_like(dg), torch.empty_like(db)
        ops.layernorm_wgrad(dy, x, mean, rstd, dg, db)
        ops.layernorm_wgrad(dy, x, mean, rstd, dg2, db2)
        torch.cuda.synchronize()
        assert torch.equal(dg, dg2) and torch.equal(db, db2)
        xh = (x.float() - mean[:, None]) * rstd[:, None]
        gref, bref = (dy.float() * xh).sum(0), dy.float().sum(0)
        assert rel_err(dg, gref) < 1e-3 or float((dg - gref).abs().max()) < 1e-2
        assert rel_err(db, bref) < 1e-3 or float((db - bref).abs().max()) < 1e-2


# --------------------------------------------------------------------------------------------- attention
def _attn_ref(q, k, v, lengths, b, t, h, scale):
    q, k, v = (x.float().view(b, t, h, 64).transpose(1, 2) for x in (q, k, v))
    valid = torch.arange(t, device=q.device)[None, :] < lengths[:, None]
    s = q @ k.transpose(2, 3) * scale
    s = s.masked_fill(~valid[:, None, None, :], float("-inf"))
    p = torch.softmax(s, -1)
    p = torch.nan_to_num(p)
    o = (p @ v).transpose(1, 2).reshape(b * t, h * 64)
    return o * valid.reshape(b * t, 1)


@pytest.fixture(params=[0, 1, 3], ids=["tcgen05", "mma_sync", "tcgen05_keyblock_fwd"])
def attn_impl(request):
    lib = pkg()._lib.load()
    lib.jl_debug_set_attn_impl(request.param)
    yield request.param
    lib.jl_debug_set_attn_impl(pkg()._lib.DEFAULT_ATTN_IMPL)


@pytest.mark.parametrize("b,t,h,lens", [(3, 250, 12, [250, 131, 64]), (2, 70, 1, [70, 1]), (2, 750, 2, [750, 300]), (1, 33, 3, [20]),
                                        (4, 128, 2, [128, 65, 64, 0]), (4, 256, 3, [256, 129, 128, 193]), (2, 257, 1, [257, 200])])
def test_attention_fwd_bwd(attn_impl, b, t, h, lens):
    ops = pkg().ops
    g = _g(2)
    d = h * 64
    qkv = (torch.randn(b * t, 3 * d, device="cuda", generator=g)).to(BF16)
    lengths = torch.tensor(lens, dtype=I32, device="cuda")
    scale = 0.125
    q, k, v = qkv[:, :d], qkv[:, d:2 * d], qkv[:, 2 * d:]
    o, lse = ops.attn_fwd(q, k, v, lengths, b, t, h, scale, want_lse=True)
    qr = qkv.float().requires_grad_(True)
    ref = _attn_ref(qr[:, :d], qr[:, d:2 * d], qr[:, 2 * d:], lengths, b, t, h, scale)
    assert rel_err(o.float(), ref) < 1e-2
    valid = (torch.arange(t, device="cuda")[None, :] < lengths[:, None]).reshape(b * t, 1)
    assert float((o.float() * (~valid)).abs().max()) == 0.0
    d_o = (torch.randn(b * t, d, device="cuda", generator=g) * valid).to(BF16)
    ref.backward(d_o.float())
    dqkv = ops.attn_bwd(q, k, v, o, d_o, lse, lengths, b, t, h, scale)
    torch.cuda.synchronize()
    gref = qr.grad * valid            # padded rows carry no gradient
    for name, sl in (("dq", slice(0, d)), ("dk", slice(d, 2 * d)), ("dv", slice(2 * d, 3 * d))):
        assert rel_err(dqkv[:, sl].float(), gref[:, sl]) < 2e-2, name


@pytest.mark.parametrize("b,t,h", [(32, 250, 12), (32, 250, 1), (5, 256, 2), (3, 97, 4)])
def test_attention_bwd_fused_matches_two_kernel_path(b, t, h):
    """<= 256 frames: the one-CTA-per-(utterance, head) backward (scores evaluated once; P / dS read K- and MN-major) against
    the dQ + dKV kernels on the same inputs, full bench size included, ragged lengths with an empty and a one-frame utterance."""
    P = pkg()
    ops, lib = P.ops, P._lib.load()
    g = _g(5)
    d = h * 64
    qkv = torch.randn(b * t, 3 * d, device="cuda", generator=g).to(BF16)
    lens = torch.randint(1, t + 1, (b,), generator=torch.Generator().manual_seed(3)).tolist()
    lens[0], lens[1], lens[-1] = t, 0, 1
    if b > 3:
        lens[2], lens[3] = 128, 129
    lengths = torch.tensor(lens, dtype=I32, device="cuda")
    q, k, v = qkv[:, :d], qkv[:, d:2 * d], qkv[:, 2 * d:]
    o, lse = ops.attn_fwd(q, k, v, lengths, b, t, h, 0.125, want_lse=True)
    valid = (torch.arange(t, device="cuda")[None, :] < lengths[:, None]).reshape(b * t, 1)
    d_o = (torch.randn(b * t, d, device="cuda", generator=g) * valid).to(BF16)
    fused = ops.attn_bwd(q, k, v, o, d_o, lse, lengths, b, t, h, 0.125).clone()
    lib.jl_debug_set_attn_impl(3)
    try:
        two = ops.attn_bwd(q, k, v, o, d_o, lse, lengths, b, t, h, 0.125).clone()
    finally:
        lib.jl_debug_set_attn_impl(P._lib.DEFAULT_ATTN_IMPL)
    torch.cuda.synchronize()
    assert torch.isfinite(fused.float()).all()
    assert float((fused.float() * (~valid)).abs().max()) == 0.0            # padded rows carry no gradient
    for name, sl in (("dq", slice(0, d)), ("dk", slice(d, 2 * d)), ("dv", slice(2 * d, 3 * d))):
        assert rel_err(fused[:, sl].float(), two[:, sl].float()) < 6e-3, name   # both round P / dS to bf16, in different orders


# --------------------------------------------------------------------------------------------- CTC
def _ctc_case(b, t, v, smax, seed, feasible=True):
    g = torch.Generator().manual_seed(seed)
    logits = torch.randn(b, t, v, generator=g) * 2.0
    ilens = torch.randint(max(1, t // 2), t + 1, (b,), generator=g)
    ilens[0] = t
    labels = torch.full((b, smax), -100, dtype=torch.int64)
    for i in range(b):
        s = int(torch.randint(0, min(smax, int(ilens[i]) // 2) + 1, (1,), generator=g))
        lab = torch.randint(1, v, (s,), generator=g)
        if s >= 2:
            lab[1] = lab[0]                                  # a repeated label (needs a blank between)
        labels[i, :s] = lab
    return logits, labels, ilens


@pytest.mark.parametrize("reduction", ["sum", "mean"])
@pytest.mark.parametrize("b,t,v,smax", [(4, 50, 37, 12), (3, 250, 5000, 100), (2, 9, 8, 4)])
def test_ctc_loss_and_grad_match_oracle(b, t, v, smax, reduction):
    from oracle import ctc as oc
    ops = pkg().ops
    logits, labels, ilens = _ctc_case(b, t, v, smax, seed=b * 100 + t)
    oloss, onll, ograd = oc.ctc_loss_and_grad(logits, labels, ilens, 0, reduction, False)
    loss, nll, grad = ops.ctc_loss(logits.cuda(), labels.to(I32).cuda(), ilens.to(I32).cuda(), 0, reduction, False, want_grad=True,
                                   grad_dtype=F32)
    torch.cuda.synchronize()
    assert abs(float(loss) - oloss) <= 1e-4 * max(1.0, abs(oloss))          # north_star: 1e-3 relative
    assert rel_err(nll, onll) < 1e-5
    # fp32 log-space lattices of magnitude ~2e3 (250 frames × V = 5000) carry ~1e-4 absolute noise per state: compare
    # relative to the largest gradient entry
    assert float((grad.cpu() - ograd).abs().max()) < 1e-3 * max(1.0, float(ograd.abs().max()))
    # bf16 logits / bf16 grad variant (what the training path uses)
    loss16, _, grad16 = ops.ctc_loss(logits.cuda().to(BF16), labels.to(I32).cuda(), ilens.to(I32).cuda(), 0, reduction, False,
                                     want_grad=True, grad_dtype=BF16)
    o2, _, g2 = oc.ctc_loss_and_grad(logits.to(BF16).float(), labels, ilens, 0, reduction, False)
    assert abs(float(loss16) - o2) <= 1e-4 * max(1.0, abs(o2))
    assert float((grad16.float().cpu() - g2).abs().max()) < 5e-3


@pytest.mark.parametrize("b,t,v,smax", [(2, 640, 50, 300), (3, 120, 3, 40)])
def test_ctc_long_label_sequences_and_heavy_label_repeats(b, t, v, smax):
    """2S+1 > 512 extended states (several states per lattice thread: the generic kernel path) and a 2-symbol alphabet (every
    label repeats many times: the linked duplicate chains of the gradient kernel)."""
    from oracle import ctc as oc
    ops = pkg().ops
    g = torch.Generator().manual_seed(b * 7 + t)
    logits = torch.randn(b, t, v, generator=g) * 2.0
    ilens = torch.full((b,), t, dtype=torch.int64)
    ilens[-1] = t - 7
    labels = torch.full((b, smax), -100, dtype=torch.int64)
    for i in range(b):
        s = smax if i == 0 else smax // 2
        lab = torch.randint(1, v, (s,), generator=g)
        if v > 3:
            lab[1::2] = torch.where(lab[1::2] == lab[0::2][: lab[1::2].numel()], (lab[1::2] % (v - 1)) + 1, lab[1::2])   # keep it feasible
        labels[i, :s] = lab
    oloss, onll, ograd = oc.ctc_loss_and_grad(logits, labels, ilens, 0, "sum", True)
    loss, nll, grad = ops.ctc_loss(logits.cuda(), labels.to(I32).cuda(), ilens.to(I32).cuda(), 0, "sum", True, want_grad=True, grad_dtype=F32)
    torch.cuda.synchronize()
    assert math.isfinite(oloss)
    assert abs(float(loss) - oloss) <= 1e-4 * max(1.0, abs(oloss))
    assert rel_err(nll, onll) < 1e-5
    assert float((grad.cpu() - ograd).abs().max()) < 1e-3 * max(1.0, float(ograd.abs().max()))


def test_ctc_golden_cases_infeasible_and_zero_infinity():
    ops = pkg().ops
    gold = np.load(os.path.join(ROOT, "tests", "golden", "golden.npz"))
    for ci in range(3):
        logits = torch.from_numpy(gold[f"ctc{ci}_logits"]).cuda()
        labels = torch.from_numpy(gold[f"ctc{ci}_labels"]).cuda()
        ilens = torch.from_numpy(gold[f"ctc{ci}_input_lengths"]).cuda()
        for red in ("sum", "mean"):
            for zi in (0, 1):
                loss, nll, grad = ops.ctc_loss(logits, labels, ilens, 0, red, bool(zi), want_grad=True, grad_dtype=F32)
                ref_loss = float(gold[f"ctc{ci}_{red}_{zi}_loss"])
                ref_grad = gold[f"ctc{ci}_{red}_{zi}_grad"]
                if np.isinf(ref_loss):
                    assert math.isinf(float(loss))
                    continue
                assert abs(float(loss) - ref_loss) <= 1e-4 * max(1.0, abs(ref_loss)), (ci, red, zi)
                finite = np.isfinite(ref_grad)
                assert np.allclose(grad.cpu().numpy()[finite], ref_grad[finite], atol=5e-5), (ci, red, zi)


# --------------------------------------------------------------------------------------------- greedy
def test_greedy_bit_exact():
    from oracle import ctc as oc
    ops = pkg().ops
    g = torch.Generator().manual_seed(3)
    b, t, v = 6, 250, 5000
    logits = torch.randn(b, t, v, generator=g)
    # make long runs and blanks likely, plus exact ties (first max must win)
    ids = torch.randint(0, 6, (b, t), generator=g)
    logits[torch.arange(b)[:, None], torch.arange(t)[None, :], ids] += 20.0
    logits[0, 5, 17] = logits[0, 5].max()
    logits[0, 5, 9] = logits[0, 5, 17]
    lens = torch.tensor([250, 249, 100, 1, 0, 33])
    ref = oc.greedy_decode(logits, lens, 0)
    out_ids, out_len, frame_ids = ops.ctc_greedy(logits.cuda(), lens.to(I32).cuda(), 0)
    torch.cuda.synchronize()
    out_ids, out_len = out_ids.cpu(), out_len.cpu()
    for i in range(b):
        assert out_ids[i, : int(out_len[i])].tolist() == ref[i], i
        assert (out_ids[i, int(out_len[i]):] == -1).all()
    am = torch.argmax(logits, -1)
    for i in range(b):
        assert torch.equal(frame_ids[i, : int(lens[i])].cpu().long(), am[i, : int(lens[i])])
    # golden strings
    gold = np.load(os.path.join(ROOT, "tests", "golden", "golden.npz"))
    for i in range(5):
        fr = torch.from_numpy(gold[f"greedy{i}_frames"].astype(np.int64))
        lg = torch.full((1, len(fr), 16), -3.0)
        lg[0, torch.arange(len(fr)), fr] = 3.0
        oi, ol, _ = ops.ctc_greedy(lg.cuda().to(BF16), torch.tensor([len(fr)], dtype=I32).cuda(), 0)
        assert oi[0, : int(ol[0])].cpu().tolist() == gold[f"greedy{i}_ids"].tolist()


# --------------------------------------------------------------------------------------------- helpers
def test_im2col_embed_transpose_colsum_cast_add_adamw():
    ops = pkg().ops
    md = pkg().modeling
    g = _g(5)
    # im2col == unfold of the zero-padded input
    b, t, c = 3, 37, 16
    x = torch.randn(b, t, c, device="cuda", generator=g).to(BF16)
    col, t_out = ops.im2col_k5s2(x)
    xp = F.pad(x.float().transpose(1, 2), (2, 2))                       # [b, c, t+4]
    ref = xp.unfold(2, 5, 2)                                            # [b, c, t_out, 5]
    ref = ref.permute(0, 2, 3, 1).reshape(b * t_out, 5 * c)
    assert t_out == (t - 1) // 2 + 1 and torch.equal(col.float(), ref)
    # conv1d + GLU through im2col + GEMM == F.conv1d + F.glu
    conv = torch.nn.Conv1d(c, 24, 5, stride=2, padding=2).cuda()
    with torch.no_grad():
        conv.weight.copy_(conv.weight.to(BF16).float())
    half = 12
    idx = torch.stack([torch.arange(half), torch.arange(half) + half], 1).reshape(-1).cuda()
    wg = conv.weight.detach().permute(0, 2, 1).reshape(24, 5 * c)[idx].to(BF16).contiguous()
    out = ops.gemm(col, wg, bias=conv.bias.detach()[idx].contiguous(), epilogue=pkg()._lib.JL_EPI_GLU, out_dtype=F32)
    refc = F.glu(conv(x.float().transpose(1, 2)), dim=1).transpose(1, 2).reshape(b * t_out, half)
    assert rel_err(out, refc) < 2e-3
    # embed positions
    d, seq = 64, 19
    h = torch.randn(2 * seq, d, device="cuda", generator=g).to(BF16)
    lens = torch.tensor([19, 7], dtype=I32, device="cuda")
    tab = md.sinusoid_table(seq + 2, d).cuda()
    ref = h.float().view(2, seq, d) * 8.0 + tab[2: seq + 2][None]
    ref = ref * (torch.arange(seq, device="cuda")[None, :, None] < lens[:, None, None])
    got = ops.embed_positions_(h.clone(), 8.0, tab, lens, 2, seq)
    assert rel_err(got.float().view(2, seq, d), ref) < 5e-3
    # transpose / colsum / cast / add
    m = torch.randn(70, 45, device="cuda", generator=g).to(BF16)
    assert torch.equal(ops.transpose(m), m.t())
    big = torch.randn(5000, 333, device="cuda", generator=g).to(BF16)
    assert rel_err(ops.colsum(big), big.float().sum(0)) < 1e-4
    f = torch.randn(1001, device="cuda", generator=g)
    assert torch.equal(ops.cast_bf16(f), f.to(BF16))
    a1, a2 = big[:100].contiguous(), big[100:200].contiguous()
    assert torch.equal(ops.add(a1, a2), (a1.float() + a2.float()).to(BF16))
    # AdamW == torch.optim.AdamW
    p = torch.randn(4099, device="cuda", generator=g)
    ref_p = torch.nn.Parameter(p.clone())
    opt = torch.optim.AdamW([ref_p], lr=1e-2, weight_decay=0.05)
    m1, m2 = torch.zeros_like(p), torch.zeros_like(p)
    shadow = torch.empty(4099, dtype=BF16, device="cuda")
    for step in range(1, 4):
        gr = torch.randn(4099, device="cuda", generator=g)
        ref_p.grad = gr.clone() * 0.5
        opt.step()
        ops.adamw_(p, gr, m1, m2, step, 1e-2, weight_decay=0.05, grad_scale=0.5, param_bf16=shadow)
    torch.cuda.synchronize()
    assert rel_err(p, ref_p.detach()) < 1e-5
    assert torch.equal(shadow, p.to(BF16))


# --------------------------------------------------------------------------------------------- fused WFAdapter
@pytest.mark.parametrize("d,b,r,rows,seq,lens", [(768, 256, 32, 1000, 250, [250, 100, 250, 7]), (1024, 256, 32, 300, 150, [150, 149]),
                                                   (128, 64, 16, 77, 77, [50]), (256, 128, 64, 520, 130, [130, 0, 1, 99])])
def test_wfadapter_fused_forward_matches_oracle(d, b, r, rows, seq, lens):
    """One-kernel WFAdapter (LN folded into the first projection) vs the oracle's wf_adapter on the same bf16 factors."""
    from oracle import encoder as oe
    P = pkg()
    ops, md = P.ops, P.modeling
    torch.manual_seed(0)
    ad = md.WFAdapter(d, b, r, num_dialects=2)
    with torch.no_grad():
        ad.norm.weight.copy_(1.0 + 0.1 * torch.randn(d))
        ad.norm.bias.copy_(0.1 * torch.randn(d))
        ad.down_bias.copy_(0.05 * torch.randn(2, b))
        ad.up_bias.copy_(0.05 * torch.randn(2, d))
        for q in ad.parameters():
            q.copy_(q.to(BF16).float())
    ad = ad.cuda()

    class _Enc:                       # minimal stand-in so that JLEngine can be used for its packing helper
        pass
    eng = md.JLEngine.__new__(md.JLEngine)
    eng._shadow = {}
    eng.flat = None
    g = _g(9)
    h = (torch.randn(rows, d, device="cuda", generator=g) * 1.5 + 0.2).to(BF16)
    lengths = torch.tensor(lens, dtype=I32, device="cuda")
    for k in (0, 1):
        pack = eng._wf_pack(ad, k)
        out, mean, rstd = ops.wfadapter_fwd(h, pack, ad.norm.eps, row_lengths=lengths, rows_per_seq=seq, save_stats=True)
        torch.cuda.synchronize()
        w = {"a.norm.weight": ad.norm.weight.detach().cpu(), "a.norm.bias": ad.norm.bias.detach().cpu()}
        for nm in ("down_B", "down_A", "down_bias", "up_B", "up_A", "up_bias"):
            w["a." + nm] = getattr(ad, nm).detach().cpu()
        ref = oe.wf_adapter(w, "a", h.float().cpu(), dialect=k)
        t = torch.arange(rows) % seq
        valid = t < torch.tensor(lens)[torch.arange(rows) // seq]
        ref = ref * valid[:, None]
        assert rel_err(out.float(), ref) < 1e-2, (k, rel_err(out.float(), ref))
        assert float(out.float().cpu()[~valid].abs().max() if (~valid).any() else 0.0) == 0.0
        hm = h.float().cpu()
        assert rel_err(mean, hm.mean(-1)) < 1e-3
        assert rel_err(rstd, 1.0 / torch.sqrt(hm.var(-1, unbiased=False) + ad.norm.eps)) < 1e-3


@pytest.mark.gpu
def test_comm_single_rank_roundtrip():
    """jl_comm_unique_id → jl_comm_init(world 1) → allreduce (identity on one rank) → destroy, all through the C ABI."""
    P = pkg()
    comm = P.JLComm(P.JLComm.unique_id(), 0, 1)
    x = torch.arange(1000, dtype=torch.float32, device="cuda")
    y = comm.allreduce_(x.clone())
    torch.cuda.synchronize()
    assert torch.equal(x, y)
    with pytest.raises(ValueError):
        comm.allreduce_(x.to(torch.bfloat16))
    comm.destroy()
    with pytest.raises(RuntimeError):
        comm.allreduce_(x)


@pytest.mark.parametrize("rows", [8000, 777, 64])
def test_colreduce_multi_matches_torch(rows):
    """Three column reductions in one launch (bias gradients of two projections + LayerNorm dγ / dβ), vs fp64 torch; repeatable
    bit for bit (fixed-order reduction)."""
    P = pkg()
    ops = P.ops
    g = _g(rows)
    dy = (torch.randn(rows, 768, device="cuda", generator=g) * 0.1).to(BF16)
    dq = (torch.randn(rows, 192, device="cuda", generator=g) * 0.1).to(BF16)[:, :]
    dz = (torch.randn(rows, 768, device="cuda", generator=g) * 0.1).to(BF16)
    h = (torch.randn(rows, 768, device="cuda", generator=g) * 2 + 0.3).to(BF16)
    mean = h.float().mean(-1)
    rstd = 1.0 / torch.sqrt(h.float().var(-1, unbiased=False) + 1e-5)
    outs = []
    for _ in range(2):
        o1, o2, o3, o4 = (torch.empty(n, device="cuda") for n in (768, 192, 768, 768))
        ops.colreduce_multi([dict(dy=dy, out_sum=o1), dict(dy=dq, out_sum=o2), dict(dy=dz, x=h, mean=mean, rstd=rstd, out_sum=o3, out_dot=o4)])
        torch.cuda.synchronize()
        outs.append((o1, o2, o3, o4))
    for a, b in zip(outs[0], outs[1]):
        assert torch.equal(a, b)
    o1, o2, o3, o4 = outs[0]
    xhat = (h.double() - mean.double()[:, None]) * rstd.double()[:, None]
    for got, ref in ((o1, dy.double().sum(0)), (o2, dq.double().sum(0)), (o3, dz.double().sum(0)), (o4, (dz.double() * xhat).sum(0))):
        assert float((got.double() - ref).abs().max()) <= 1e-4 * float(ref.abs().max()) + 1e-5


@pytest.mark.parametrize("d,b,r,rows,kk", [(768, 256, 32, 1000, 2), (1024, 256, 32, 300, 1), (128, 64, 16, 77, 3)])
def test_wfadapter_training_forward_saves_intermediates_and_device_pack_matches_host_pack(d, b, r, rows, kk):
    """Training mode of the fused WFAdapter kernel: the same launch writes t1 = LN(h) B_dᵀ, u = relu(t1 A_dᵀ + c_d), t2 = u B_uᵀ and
    the LayerNorm statistics; the LayerNorm-folded operands come from ``jl_wfadapter_pack`` (device) and equal the host pack."""
    P = pkg()
    ops, md = P.ops, P.modeling
    from oracle import encoder as oe
    torch.manual_seed(5)
    ad = P.WFAdapter(d, b, r, num_dialects=kk)
    with torch.no_grad():
        ad.norm.weight.copy_(1.0 + 0.2 * torch.randn(d))
        ad.norm.bias.copy_(0.1 * torch.randn(d))
        ad.down_bias.copy_(0.05 * torch.randn(kk, b))
        ad.up_bias.copy_(0.05 * torch.randn(kk, d))
        for q in (ad.down_B, ad.down_A, ad.up_B, ad.up_A):
            q.mul_(3.0)
        for q in ad.parameters():
            q.copy_(q.to(BF16).float())
    ad = ad.cuda()
    eng = md.JLEngine.__new__(md.JLEngine)
    eng._shadow, eng.flat, eng._wf_bufs = {}, None, {}
    g = _g(3)
    h = (torch.randn(rows, d, device="cuda", generator=g) * 1.5 + 0.2).to(BF16)
    bufs = eng._wf_pack_dev(ad)
    for k in range(kk):
        host = eng._wf_pack(ad, k)
        assert torch.equal(bufs["bd"][k], host["bd"]) and torch.equal(bufs["ad"][k], host["ad"]) and torch.equal(bufs["au"][k], host["au"])
        assert rel_err(bufs["s"][k], host["s"]) < 1e-5 and rel_err(bufs["t"][k], host["t"]) < 1e-4
        pack = {"bd": bufs["bd"][k], "s": bufs["s"][k], "t": bufs["t"][k], "ad": bufs["ad"][k], "au": bufs["au"][k], "bu": eng._bf16(ad.up_B)[k],
                "c_d": ad.down_bias.detach()[k], "c_u": ad.up_bias.detach()[k], "r": r, "b": b}
        t1 = torch.empty((rows, r), dtype=BF16, device="cuda")
        u = torch.empty((rows, b), dtype=BF16, device="cuda")
        t2 = torch.empty((rows, r), dtype=BF16, device="cuda")
        out, mean, rstd = ops.wfadapter_fwd(h, pack, ad.norm.eps, save_stats=True, t1=t1, u=u, t2=t2)
        out2, _, _ = ops.wfadapter_fwd(h, host, ad.norm.eps)
        torch.cuda.synchronize()
        assert torch.equal(out, out2)                              # saving the intermediates does not change the result
        hf = h.float().cpu()
        z = F.layer_norm(hf, (d,), ad.norm.weight.detach().cpu(), ad.norm.bias.detach().cpu(), ad.norm.eps)
        rt1 = z @ ad.down_B.detach().cpu()[k].t()
        ru = torch.relu(rt1 @ ad.down_A.detach().cpu()[k].t() + ad.down_bias.detach().cpu()[k])
        rt2 = ru @ ad.up_B.detach().cpu()[k].t()
        assert rel_err(t1.float(), rt1) < 1e-2 and rel_err(u.float(), ru) < 2e-2 and rel_err(t2.float(), rt2) < 2e-2
        assert rel_err(mean, hf.mean(-1)) < 1e-3 and rel_err(rstd, 1.0 / torch.sqrt(hf.var(-1, unbiased=False) + ad.norm.eps)) < 1e-3


# --------------------------------------------------------------------------------------------- fused AttAdapter
@pytest.mark.parametrize("d,seq,lens,packed", [(768, 250, [250, 100, 250, 7, 129, 128], False), (1024, 256, [256, 1, 200], False),
                                                (128, 77, [50, 77], False), (768, 136, [136, 0, 64], False),
                                                (768, 250, [250, 100, 37, 129, 250], True), (256, 200, [64, 65, 200, 128], True)])
def test_attadapter_fused_forward_matches_oracle(d, seq, lens, packed):
    """One-kernel AttAdapter (LN folded into the q|k|v projection, whole-row softmax, output projection + residual) vs the oracle's
    att_adapter on the same bf16 weights, padded ([B·seq] rows) and packed (cu_seqlens) row layouts, plus the tensors it saves for
    the backward pass.  a7 is definitional: nothing external pins it; the oracle is the definition."""
    from oracle import encoder as oe
    P = pkg()
    ops, md = P.ops, P.modeling
    torch.manual_seed(1)
    ad = md.AttAdapter(d)
    with torch.no_grad():
        ad.norm.weight.copy_(1.0 + 0.1 * torch.randn(d))
        ad.norm.bias.copy_(0.1 * torch.randn(d))
        for lin in (ad.q_proj, ad.k_proj, ad.v_proj, ad.o_proj):
            lin.weight.copy_(torch.randn(lin.weight.shape) * (0.08 if lin is not ad.o_proj else 0.05))
            lin.bias.copy_(0.05 * torch.randn(lin.bias.shape))
        for q in ad.parameters():
            q.copy_(q.to(BF16).float())
    w = {"a." + k: v.detach().clone() for k, v in ad.state_dict().items()}
    ad = ad.cuda()
    b = len(lens)
    g = _g(11)
    hp = (torch.randn(b, seq, d, device="cuda", generator=g) * 1.5 + 0.2).to(BF16)
    lt = torch.tensor(lens)
    valid = torch.arange(seq)[None, :] < lt[:, None]                           # [B, seq]
    hp = hp * valid[:, :, None].cuda()                                         # padded rows are zero, as the engine keeps them
    ref = oe.att_adapter(w, "a", hp.float().cpu(), lt)                         # [B, seq, d]
    wqkv = torch.cat([ad.q_proj.weight, ad.k_proj.weight, ad.v_proj.weight]).detach().to(BF16).contiguous()
    bqkv = torch.cat([ad.q_proj.bias, ad.k_proj.bias, ad.v_proj.bias]).detach().float().contiguous()
    pack = ops.lnfold_pack(wqkv, bqkv, ad.norm.weight.detach(), ad.norm.bias.detach())
    # the fold itself
    ws = (wqkv.float() * ad.norm.weight.detach()[None, :]).to(BF16)
    assert torch.equal(pack["w"], ws)
    assert rel_err(pack["s"], ws.float().sum(-1)) < 1e-5
    assert rel_err(pack["tb"], wqkv.float() @ ad.norm.bias.detach() + bqkv) < 1e-5
    wo = ad.o_proj.weight.detach().to(BF16).contiguous()
    bo = ad.o_proj.bias.detach().float()
    lengths = lt.to(I32).cuda()
    if packed:
        cu = torch.zeros(b + 1, dtype=I32)
        cu[1:] = torch.cumsum(lt, 0)
        rows = int(cu[-1])
        h = hp[valid.cuda()].contiguous()
        ref2 = ref[valid]
        out, sv = ops.attadapter_fwd(h, pack, wo, bo, lengths, b, seq, ad.norm.eps, training=True, cu_seqlens=cu.cuda())
        vrow = torch.ones(rows, dtype=torch.bool)
    else:
        h = hp.reshape(b * seq, d)
        ref2 = ref.reshape(b * seq, d)
        vrow = valid.reshape(-1)
        out, sv = ops.attadapter_fwd(h, pack, wo, bo, lengths, b, seq, ad.norm.eps, zero_padded_rows=True, training=True)
        ref2 = torch.where(vrow[:, None], ref2, torch.zeros(()))      # (an empty utterance is all-NaN in the oracle's softmax)
    torch.cuda.synchronize()
    mean, rstd, qkv, a, lse = sv
    assert rel_err(out.float(), ref2) < 1e-2, rel_err(out.float(), ref2)
    if (~vrow).any():
        assert float(out.float().cpu()[~vrow].abs().max()) == 0.0
    # the adapter's own contribution (out − h), which the residual would otherwise mask
    dlt, dref = (out.float().cpu() - h.float().cpu())[vrow], (ref2 - h.float().cpu())[vrow]
    assert rel_err(dlt, dref) < 4e-2, rel_err(dlt, dref)
    # saved tensors
    hv = h.float().cpu()[vrow]
    assert rel_err(mean.cpu()[vrow], hv.mean(-1)) < 1e-3
    assert rel_err(rstd.cpu()[vrow], 1.0 / torch.sqrt(hv.var(-1, unbiased=False) + ad.norm.eps)) < 1e-3
    z = F.layer_norm(hv, (d,), w["a.norm.weight"], w["a.norm.bias"], ad.norm.eps)
    qkv_ref = z @ wqkv.float().cpu().T + bqkv.cpu()
    assert rel_err(qkv.float().cpu()[vrow], qkv_ref) < 1e-2
    # two clusters per utterance sharing the output columns (col_split = 2): bit-identical outputs and saved tensors
    if d >= 256:
        kw = dict(training=True, col_split=2)
        if packed:
            out_s, sv_s = ops.attadapter_fwd(h, pack, wo, bo, lengths, b, seq, ad.norm.eps, cu_seqlens=cu.cuda(), **kw)
        else:
            out_s, sv_s = ops.attadapter_fwd(h, pack, wo, bo, lengths, b, seq, ad.norm.eps, zero_padded_rows=True, **kw)
        torch.cuda.synchronize()
        assert torch.equal(out_s, out)
        for t_a, t_b in zip(sv_s, sv):
            assert torch.equal(t_a[vrow.cuda()] if t_a.shape[0] == vrow.numel() else t_a, t_b[vrow.cuda()] if t_b.shape[0] == vrow.numel() else t_b)
    # without the saved tensors and without zeroing: padded rows get h + b_o (what the composed path writes)
    if not packed:
        out2, sv2 = ops.attadapter_fwd(h, pack, wo, bo, lengths, b, seq, ad.norm.eps)
        assert sv2 is None
        torch.cuda.synchronize()
        assert torch.equal(out2[vrow.cuda()], out[vrow.cuda()])
        if (~vrow).any():
            exp = (h.float() + bo[None, :]).to(BF16)
            assert torch.equal(out2[~vrow.cuda()], exp[~vrow.cuda()])


@pytest.mark.parametrize("packed", [False, True])
def test_attadapter_fused_matches_composed_path_through_the_engine(packed):
    """Model-level: loss and adapter gradients with the fused AttAdapter forward == the composed LN → GEMM → attention → GEMM path
    on the same weights and batch; bf16 rounding points differ (LN(h) is never rounded in the fused kernel), hence tolerances."""
    from helpers import synth_wave
    P = pkg()
    cfg = P.JLConfig(hidden_size=256, num_hidden_layers=2, num_attention_heads=4, intermediate_size=512, conv_channels=64, vocab_size=48,
                     adapter_attn="att", adapter_ffn="att")
    torch.manual_seed(3)
    model = P.JLForCTC(cfg).cuda()
    model.freeze_base_model()
    waves = [synth_wave(n, i) for i, n in enumerate([32000, 16000, 24000, 8000])]
    fe = P.JLFeatureExtractor(device="cuda")
    feats = fe([w.numpy() for w in waves], sampling_rate=16000)
    labels = torch.full((4, 8), -100, dtype=torch.int64)
    for i, k in enumerate((7, 3, 5, 2)):
        labels[i, :k] = torch.randint(1, 48, (k,))
    res = {}
    for fused in (True, False):
        model.zero_grad(set_to_none=True)
        model.encoder.engine(model.lm_head).fused_att = fused
        loss, _ = model(feats["input_features"], attention_mask=feats["attention_mask"], labels=labels.cuda(), packed=packed)
        loss.backward()
        torch.cuda.synchronize()
        res[fused] = (float(loss), {n: p.grad.detach().float().clone() for n, p in model.named_parameters() if p.grad is not None})
    assert abs(res[True][0] - res[False][0]) <= 2e-3 * abs(res[False][0]), (res[True][0], res[False][0])
    assert res[True][1].keys() == res[False][1].keys() and len(res[True][1]) > 0
    for n, gr in res[False][1].items():
        if float(gr.norm()) == 0.0 or n.endswith("k_proj.bias"):      # the key bias has an analytically zero gradient
            continue
        assert rel_err(res[True][1][n], gr) < 3e-2, (n, rel_err(res[True][1][n], gr))


def test_attadapter_fused_is_bit_reproducible_and_matches_the_composed_path_on_random_lengths():
    """Race check without a sanitizer: 64 utterances of random lengths (0 … 256, both CTAs of a cluster active / one / none), the same
    launch repeated 25 times back to back (PDL on) must give bit-identical outputs and saved tensors, and the result must agree with
    the composed path (LayerNorm → GEMM → attention → GEMM) of the engine on every valid row."""
    P = pkg()
    ops, md = P.ops, P.modeling
    torch.manual_seed(5)
    d, b, seq = 768, 64, 256
    cfg = P.JLConfig(hidden_size=d, num_hidden_layers=1, num_attention_heads=d // 64, intermediate_size=4 * d, adapter_ffn="att")
    model = P.JLForCTC(cfg).cuda().eval()
    eng = model.encoder.engine(model.lm_head)
    ad = model.encoder.layers[0].adapter_ffn
    with torch.no_grad():
        for lin in (ad.q_proj, ad.k_proj, ad.v_proj, ad.o_proj):
            lin.weight.mul_(4.0)
            lin.bias.normal_(0.0, 0.05)
        ad.norm.weight.normal_(1.0, 0.1)
        ad.norm.bias.normal_(0.0, 0.1)
    g = _g(21)
    lens = torch.randint(0, seq + 1, (b,), generator=torch.Generator().manual_seed(3))
    lens[:4] = torch.tensor([0, 1, 128, 129])
    lengths = lens.to(I32).cuda()
    valid = (torch.arange(seq)[None, :] < lens[:, None]).reshape(-1).cuda()
    h = ((torch.randn(b * seq, d, device="cuda", generator=g) * 1.3 + 0.1) * valid[:, None]).to(BF16)
    eng.fused_att = True
    outs = [eng._adapter_fwd(ad, h, lengths, b, seq, True, 0, True) for _ in range(25)]
    torch.cuda.synchronize()
    out0, sv0 = outs[0]
    for out, sv in outs[1:]:
        assert torch.equal(out, out0)
        for t0, t1 in zip(sv0, sv):
            if isinstance(t0, torch.Tensor):
                assert torch.equal(t0[..., :][...] if t0.dim() != 2 or t0.shape[0] != b * seq else t0[valid], t1 if t1.dim() != 2 or t1.shape[0] != b * seq else t1[valid])
    eng.fused_att = False
    ref, _ = eng._adapter_fwd(ad, h, lengths, b, seq, True, 0, True)
    torch.cuda.synchronize()
    assert float(out0[~valid].float().abs().max()) == 0.0 and float(ref[~valid].float().abs().max()) == 0.0
    dl, dr = (out0.float() - h.float())[valid], (ref.float() - h.float())[valid]
    assert rel_err(dl, dr) < 3e-2, rel_err(dl, dr)
    assert rel_err(out0[valid].float(), ref[valid].float()) < 1.5e-2      # two bf16 roundings of h + a (amplified) adapter update


@pytest.mark.parametrize("rows,n,d", [(8000, 192, 768), (1000, 192, 1024), (333, 64, 128), (129, 128, 256), (5, 192, 768), (2500, 32, 1024), (700, 8, 256)])
def test_lnproj_bwd_matches_gemm_plus_layernorm_bwd_and_fp32_autograd(rows, n, d):
    """jl_lnproj_bwd (dy · W + LayerNorm backward in one kernel, row means from dy and the saved projection output) vs (a) plain
    fp32 autograd through LayerNorm → Linear and (b) the two-kernel path it replaces (jl_gemm_bf16 + jl_layernorm_bwd)."""
    P = pkg()
    ops = P.ops
    g = _g(31)
    h = (torch.randn(rows, d, device="cuda", generator=g) * 1.4 + 0.3).to(BF16)
    w = (torch.randn(n, d, device="cuda", generator=g) * 0.06).to(BF16)
    bias = torch.randn(n, device="cuda", generator=g) * 0.1
    gamma = (1.0 + 0.1 * torch.randn(d, device="cuda", generator=g)).requires_grad_(True)
    beta = (0.1 * torch.randn(d, device="cuda", generator=g)).requires_grad_(True)
    dy = (torch.randn(rows, n, device="cuda", generator=g) * 0.5).to(BF16)
    dres = (torch.randn(rows, d, device="cuda", generator=g) * 0.3).to(BF16)
    eps = 1e-5
    # fp32 reference
    hf = h.float().requires_grad_(True)
    z = F.layer_norm(hf, (d,), gamma, beta, eps)
    yref = z @ w.float().T + bias
    (yref * dy.float()).sum().backward()
    dx_ref = hf.grad + dres.float()
    dz_ref = dy.float() @ w.float()
    dgamma_ref, dbeta_ref = gamma.grad.clone(), beta.grad.clone()
    gamma, beta = gamma.detach(), beta.detach()
    # product path: forward pieces the kernel consumes
    zb, mean, rstd = ops.layernorm_fwd(h, gamma, beta, eps, save_stats=True)
    pack = ops.lnfold_pack(w, bias, gamma, beta)
    y = yref.detach().to(BF16)                                   # what the forward pass saved (bf16)
    dx, dz = ops.lnproj_bwd(dy, y, w, pack, gamma, h, mean, rstd, dres, want_dz=True)
    dx2, none = ops.lnproj_bwd(dy, y, w, pack, gamma, h, mean, rstd, dres)
    torch.cuda.synchronize()
    assert none is None and torch.equal(dx, dx2)
    # column sums left for the LayerNorm weight gradients and the bias gradient behind dres
    dx3, none3, cols = ops.lnproj_bwd(dy, y, w, pack, gamma, h, mean, rstd, dres, want_cols=True)
    dg, db, do = (torch.empty(d, device="cuda") for _ in range(3))
    ops.lnproj_bwd_reduce(cols, dg, db, do)
    ops.lnproj_bwd_reduce(cols, None, None, None)
    torch.cuda.synchronize()
    assert none3 is None and torch.equal(dx3, dx) and cols.shape == (3, (rows + 127) // 128, d)
    assert rel_err(db, dbeta_ref) < 5e-3 and rel_err(dg, dgamma_ref) < 5e-3, (rel_err(db, dbeta_ref), rel_err(dg, dgamma_ref))
    assert rel_err(do, dres.float().sum(0)) < 1e-5
    # two CTAs per row tile (column halves): bit-identical dx and column partials
    if d >= 256:
        dx5, _, cols5 = ops.lnproj_bwd(dy, y, w, pack, gamma, h, mean, rstd, dres, want_cols=True, col_split=2)
        dx6, _, cols6 = ops.lnproj_bwd(dy, y, w, pack, gamma, h, mean, rstd, dres, want_cols=True, col_split=1)
        torch.cuda.synchronize()
        assert torch.equal(dx5, dx6) and torch.equal(dx5, dx) and torch.equal(cols5, cols6)
    # accumulate: a second call adds to dgamma / dbeta and overwrites dbias
    ops.lnproj_bwd_reduce(cols, dg, db, do, accumulate=True)
    torch.cuda.synchronize()
    assert rel_err(db, 2 * dbeta_ref) < 5e-3 and rel_err(dg, 2 * dgamma_ref) < 5e-3 and rel_err(do, dres.float().sum(0)) < 1e-5
    # into a caller-provided row slice
    big = torch.zeros((rows + 7, d), dtype=BF16, device="cuda")
    ops.lnproj_bwd(dy, y, w, pack, gamma, h, mean, rstd, dres, out=big[3:3 + rows] if (3 * d * 2) % 16 == 0 else big[8 - 8:rows])
    torch.cuda.synchronize()
    assert torch.equal(big[3:3 + rows] if (3 * d * 2) % 16 == 0 else big[:rows], dx)
    if n % 64 != 0:
        return
    # the projection's own weight / bias gradient without LN(h): (dy ⊙ rstd)ᵀ h finished by lnproj_wgrad
    dx4, none4, dys, wpart = ops.lnproj_bwd(dy, y, w, pack, gamma, h, mean, rstd, dres, want_wgrad_operands=True)
    m0 = torch.empty((n, d), dtype=F32, device="cuda")
    ops.gemm(dys, h, a_layout=1, b_layout=1, out=m0, out_dtype=F32)
    dbq = torch.empty(n, device="cuda")
    ops.lnproj_wgrad(m0, wpart, gamma, beta, dbq)
    torch.cuda.synchronize()
    assert torch.equal(dx4, dx) and none4 is None
    assert rel_err(dys.float(), dy.float() * rstd[:, None]) < 4e-3
    dw_ref = dy.float().T @ z.detach()
    assert rel_err(m0, dw_ref) < 1e-2, rel_err(m0, dw_ref)
    assert rel_err(dbq, dy.float().sum(0)) < 1e-4
    # the same operands from the stand-alone kernel (weight-gradient branch): bit-identical
    dys2, wpart2 = ops.lnproj_wgrad_prep(dy, mean, rstd)
    torch.cuda.synchronize()
    assert torch.equal(dys2, dys) and torch.equal(wpart2, wpart)
    assert rel_err(dz.float(), dz_ref) < 6e-3, rel_err(dz.float(), dz_ref)
    assert rel_err(dx.float() - dres.float(), dx_ref - dres.float()) < 1.5e-2, rel_err(dx.float() - dres.float(), dx_ref - dres.float())
    # the two-kernel path
    dz2 = ops.gemm(dy, w, b_layout=P._lib.JL_LAYOUT_MN) if hasattr(P._lib, "JL_LAYOUT_MN") else ops.gemm(dy, w, b_layout=1)
    dxc, _, _ = ops.layernorm_bwd(dz2, h, gamma, mean, rstd, dres=dres)
    torch.cuda.synchronize()
    e_fused, e_two = rel_err(dx.float() - dres.float(), dx_ref - dres.float()), rel_err(dxc.float() - dres.float(), dx_ref - dres.float())
    assert e_fused <= max(1.5 * e_two, 8e-3), (e_fused, e_two)     # no less accurate than what it replaces (dz is not rounded to bf16 here)


def test_lnfold_pack_multi_equals_single_packs():
    """One launch for many projections (every AttAdapter / WFAdapter of a model at the start of a training step) == one launch each,
    bit for bit, including jobs of different n (192 and 32) in the same launch and more jobs than one launch holds."""
    P = pkg()
    ops = P.ops
    g = _g(41)
    d = 256
    jobs, singles = [], []
    for i in range(53):
        n = 192 if i % 3 else 32
        w = (torch.randn(n, d, device="cuda", generator=g) * 0.1).to(BF16)
        bias = torch.randn(n, device="cuda", generator=g) if i % 2 else None
        gamma = 1.0 + 0.1 * torch.randn(d, device="cuda", generator=g)
        beta = 0.1 * torch.randn(d, device="cuda", generator=g)
        singles.append(ops.lnfold_pack(w, bias, gamma, beta))
        bufs = {k: torch.full_like(v, 7.0) for k, v in singles[-1].items()}
        jobs.append((w, bias, gamma, beta, bufs))
    ops.lnfold_pack_multi(jobs)
    torch.cuda.synchronize()
    for (w, bias, gamma, beta, bufs), ref in zip(jobs, singles):
        for k in ("w", "s", "tb"):
            assert torch.equal(bufs[k], ref[k]), k


def test_lnproj_bwd_row_runs_equal_one_call_per_run():
    """jl_lnproj_bwd with row runs (the dialect runs of a WFAdapter: every run its own factor set of w / s / tb, row tiles never
    straddle a run) == one call per run on the row slices: dx bit for bit, the summed column partials to fp32 accuracy."""
    P = pkg()
    ops = P.ops
    g = _g(51)
    d, n, sets = 768, 32, 4
    runs = [(0, 300, 2), (300, 429, 0), (429, 1500, 3)]                  # set 1 absent; lengths not multiples of 128
    rows = runs[-1][1]
    h = (torch.randn(rows, d, device="cuda", generator=g) * 1.2 + 0.2).to(BF16)
    w = (torch.randn(sets, n, d, device="cuda", generator=g) * 0.07).to(BF16)
    gamma = 1.0 + 0.1 * torch.randn(d, device="cuda", generator=g)
    beta = 0.1 * torch.randn(d, device="cuda", generator=g)
    dy = (torch.randn(rows, n, device="cuda", generator=g) * 0.5).to(BF16)
    y = (torch.randn(rows, n, device="cuda", generator=g)).to(BF16)
    dres = (torch.randn(rows, d, device="cuda", generator=g) * 0.3).to(BF16)
    _, mean, rstd = ops.layernorm_fwd(h, gamma, beta, 1e-5, save_stats=True)
    s_all = torch.empty(sets, n, device="cuda")
    tb_all = torch.empty(sets, n, device="cuda")
    packs = [ops.lnfold_pack(w[k], None, gamma, beta, {"w": torch.empty_like(w[k]), "s": s_all[k], "tb": tb_all[k]}) for k in range(sets)]
    dx, _, cols = ops.lnproj_bwd(dy, y, w.view(sets * n, d), {"s": s_all.view(-1), "tb": tb_all.view(-1)}, gamma, h, mean, rstd, dres,
                                 want_cols=True, runs=runs)
    dg, db = torch.empty(d, device="cuda"), torch.empty(d, device="cuda")
    ops.lnproj_bwd_reduce(cols, dg, db, None)
    ref_dx = torch.empty_like(dx)
    dg_ref, db_ref = torch.zeros(d, device="cuda"), torch.zeros(d, device="cuda")
    off = 0
    for i, (r0, r1, k) in enumerate(runs):
        sl = slice(r0, r1)
        _, _, c_k = ops.lnproj_bwd(dy[sl], y[sl], w[k], packs[k], gamma, h[sl], mean[sl], rstd[sl], dres[sl], want_cols=True, out=ref_dx[sl])
        a, b_, o_ref, o = (torch.empty(d, device="cuda") for _ in range(4))
        ops.lnproj_bwd_reduce(c_k, a, b_, o_ref)
        dg_ref += a
        db_ref += b_
        nt = (r1 - r0 + 127) // 128
        ops.lnproj_bwd_reduce(cols, None, None, o, tile_offset=off, num_tiles=nt)
        off += nt
        torch.cuda.synchronize()
        assert rel_err(o, o_ref) < 1e-6 and rel_err(o, dres[sl].float().sum(0)) < 1e-5, i
    torch.cuda.synchronize()
    assert cols.shape[1] == off
    assert torch.equal(dx, ref_dx)
    assert rel_err(dg, dg_ref) < 1e-5 and rel_err(db, db_ref) < 1e-5
