"""World-size-2 gloo test (CPU) of the data-parallel host logic: utterances are sharded across ranks, each rank's
adapter + lm_head gradients (here produced by the oracle) go into the product's flat bucket layout, ONE all-reduce of the
bucket follows, and the result equals the gradients of the un-sharded batch (CTC reduction 'sum' ⇒ gradients add)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from helpers import pkg, synth_wave


def _free_port() -> int:
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _problem():
    P = pkg()
    cfg = P.JLConfig(hidden_size=128, num_hidden_layers=2, num_attention_heads=2, intermediate_size=128, conv_channels=32, vocab_size=24,
                     adapter_attn="att", adapter_ffn="wf", wf_bottleneck=16, wf_rank=8)
    torch.manual_seed(0)
    model = P.JLForCTC(cfg)
    model.freeze_base_model()
    # every shard holds one utterance of the longest length: the reference's conv subsampler does not mask between its two
    # convolutions (modeling_speech_to_text.py:94-100), so results depend on the padded length of the batch
    lens = [6000, 4200, 6000, 3000]
    waves = [synth_wave(n, 10 + i) for i, n in enumerate(lens)]
    g = torch.Generator().manual_seed(3)
    labels = torch.full((4, 3), -100, dtype=torch.int64)
    for i in range(4):
        labels[i, : 2 + (i % 2)] = torch.randint(1, cfg.vocab_size, (2 + (i % 2),), generator=g)
    return P, cfg, model, waves, labels


def _oracle_grads(P, cfg, model, waves, labels):
    from oracle import model as om
    w = om.from_product_state_dict(model.state_dict())
    ocfg = om.OracleConfig(**{k: v for k, v in cfg.to_dict().items() if k in om.OracleConfig.__dataclass_fields__})
    for k, v in w.items():
        v.requires_grad_(om.is_trainable(k))
    loss, _, _ = om.forward_from_waveforms(w, ocfg, waves, labels)
    loss.backward()
    out = {}
    for name, p in model._get_adapters().items():
        key = name[len("encoder."):] if name.startswith("encoder.") else name
        out[name] = w[key].grad.clone()
    return float(loss), out


def _worker(rank: int, world: int, port: int, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    P, cfg, model, waves, labels = _problem()
    frames = [P.feature_extraction.num_frames(len(w)) for w in waves]
    shard = P.shard_utterances(frames, world)[rank]
    loss, grads = _oracle_grads(P, cfg, model, [waves[i] for i in shard], labels[shard])
    plist = P.ordered_trainables(model)
    layout = P.BucketLayout(plist)
    bucket = torch.zeros(layout.total)
    byid = {id(p): n for n, p in model._get_adapters().items()}
    for p in plist:
        layout.view(bucket, p).copy_(grads[byid[id(p)]])
    dist.all_reduce(bucket, op=dist.ReduceOp.SUM)                  # the one collective of the step
    lt = torch.tensor([loss], dtype=torch.float64)
    dist.all_reduce(lt)
    if rank == 0:
        ret["loss"] = float(lt)
        ret["bucket"] = bucket.clone()
        ret["shards"] = P.shard_utterances(frames, world)
    dist.destroy_process_group()


def test_two_rank_bucket_allreduce_equals_single_rank_batch():
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
    P, cfg, model, waves, labels = _problem()
    loss, grads = _oracle_grads(P, cfg, model, waves, labels)
    assert abs(ret["loss"] - loss) <= 1e-4 * abs(loss)
    plist = P.ordered_trainables(model)
    layout = P.BucketLayout(plist)
    byid = {id(p): n for n, p in model._get_adapters().items()}
    bucket = ret["bucket"]
    for p in plist:
        got, ref = layout.view(bucket, p), grads[byid[id(p)]]
        assert float((got - ref).norm()) <= 1e-3 * float(ref.norm()) + 1e-6, byid[id(p)]
    shards = ret["shards"]
    assert sorted(shards[0] + shards[1]) == [0, 1, 2, 3]


def test_shard_utterances_balances_frames():
    P = pkg()
    frames = [2998, 198, 1500, 1400, 700, 650, 300, 2500]
    for world in (1, 2, 4, 8):
        shards = P.shard_utterances(frames, world)
        assert sorted(i for s in shards for i in s) == list(range(len(frames)))
        loads = [sum(frames[i] for i in s) for s in shards]
        assert max(loads) - min(loads) <= max(frames)
    assert P.shard_utterances(frames, 2) == P.shard_utterances(frames, 2)        # deterministic


def test_bucket_layout_views_are_aligned_and_qkv_adjacent():
    P = pkg()
    _, cfg, model, _, _ = _problem()
    plist = P.ordered_trainables(model)
    layout = P.BucketLayout(plist)
    assert all(off % 64 == 0 for off in layout.offset.values())
    ad = model.encoder.layers[0].adapter_attn
    buf = torch.arange(layout.total, dtype=torch.float32)
    cat = layout.cat(buf, [ad.q_proj.weight, ad.k_proj.weight, ad.v_proj.weight])
    assert cat is not None and cat.shape == (192, cfg.hidden_size)
    assert torch.equal(cat[64:128], layout.view(buf, ad.k_proj.weight))
    assert layout.cat(buf, [ad.q_proj.bias, ad.k_proj.bias, ad.v_proj.bias]).shape == (192,)
