"""SURVEY §8 f1 (inference): lm_head ⊕ frame argmax ⊕ greedy collapse without a [B·T', V] logits tensor.  The JL_EPI_ARGMAX epilogue
of the tcgen05 GEMM emits per-chunk (max, first argmax) pairs; ``jl_ctc_greedy_from_partials`` reduces them.  Token ids must be
IDENTICAL to decoding the materialised logits of the same GEMM (bit-exact integer work, first maximum wins as torch.argmax)."""
import pytest
import torch

from helpers import pkg, synth_wave

pytestmark = pytest.mark.gpu
BF16, F32, I32 = torch.bfloat16, torch.float32, torch.int32


@pytest.mark.parametrize("m,d,v,batch,seq", [(1000, 256, 5000, 4, 250), (250, 768, 5000, 1, 250), (77, 128, 48, 1, 77), (8000, 768, 5000, 32, 250)])
def test_lm_head_argmax_equals_argmax_of_materialised_logits(m, d, v, batch, seq):
    P = pkg()
    ops = P.ops
    g = torch.Generator(device="cuda").manual_seed(m + v)
    h = torch.randn((m, d), device="cuda", generator=g).to(BF16)
    w = (torch.randn((v, d), device="cuda", generator=g) * 0.05).to(BF16)
    bias = torch.randn((v,), device="cuda", generator=g) * 0.1
    with torch.no_grad():                  # exact ties: duplicated vocabulary rows (identical logits) — the lower index must win
        w[v - 1] = w[3]
        bias[v - 1] = bias[3]
        if v > 40:
            w[37] = w[5]
            bias[37] = bias[5]
    logits = ops.gemm(h, w, bias=bias, out_dtype=F32)
    pmax, pidx = ops.lm_head_argmax(h, w, bias)
    lengths = torch.full((batch,), seq, dtype=I32, device="cuda")
    lengths[-1] = max(1, seq - 13)
    ids, n, frame_ids = ops.ctc_greedy_from_partials(pmax, pidx, lengths, batch, seq, blank=0)
    ids2, n2, frame_ids2 = ops.ctc_greedy(logits.view(batch, seq, v), lengths, blank=0)
    torch.cuda.synchronize()
    ref = logits.argmax(-1).view(batch, seq).to(I32)                      # first maximum wins
    for b in range(batch):
        t = int(lengths[b])
        assert torch.equal(frame_ids[b, :t], ref[b, :t]), f"utterance {b}: fused frame argmax differs from torch.argmax of the logits"
        assert bool((frame_ids[b, t:] == -1).all())
    assert torch.equal(frame_ids, frame_ids2) and torch.equal(ids, ids2) and torch.equal(n, n2)
    # the maxima themselves are the logits' maxima (same accumulators, same bias add)
    assert torch.equal(pmax[:, :m].max(0).values, logits.max(-1).values)


@pytest.mark.parametrize("packed", [False, True])
def test_transcriber_fused_head_ids_equal_unfused(packed):
    P = pkg()
    cfg = P.JLConfig(hidden_size=256, num_hidden_layers=2, num_attention_heads=4, intermediate_size=512, conv_channels=128, vocab_size=5000,
                     adapter_ffn="wf", wf_bottleneck=64, wf_rank=16)
    torch.manual_seed(1)
    model = P.JLForCTC(cfg).cuda().eval()
    ns_list = [48000, 30001, 160000, 9000]
    nmax = max(ns_list)
    wave = torch.zeros((4, nmax))
    for i, n_ in enumerate(ns_list):
        wave[i, :n_] = synth_wave(n_, 60 + i)
    ns = torch.tensor(ns_list, dtype=I32)
    outs = []
    for fused in (True, False):
        tr = P.Transcriber(model, packed=packed, fused_head=fused)
        for _ in range(2):                       # capture, then replay
            ids, nid = tr(wave.pin_memory(), ns)
            torch.cuda.synchronize()
        outs.append([ids[i, : int(nid[i])].cpu().tolist() for i in range(4)])
    assert outs[0] == outs[1]
    assert any(len(x) > 0 for x in outs[0])
    # module-level API
    fe = P.JLFeatureExtractor(device="cuda")
    feats = fe([wave[i, : ns_list[i]].numpy() for i in range(4)], sampling_rate=16000)
    assert model.transcribe(feats["input_features"], feats["attention_mask"]) == model.transcribe(feats["input_features"], feats["attention_mask"], fused_head=False)
    assert model.transcribe(feats["input_features"], feats["attention_mask"]) == outs[0]
