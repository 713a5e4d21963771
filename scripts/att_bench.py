"""In-graph timing of the AttAdapter forward: one fused kernel (jl_attadapter_fwd) vs LN → GEMM → jl_attn_fwd → GEMM."""
import importlib, sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
P = importlib.import_module("jiao-liao_speech_recognition_b200")
ops, md = P.ops, P.modeling
BF16 = torch.bfloat16
def timeit(name, fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        for _ in range(n): fn()
    gr.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); gr.replay(); e1.record(); torch.cuda.synchronize()
    print(f"{name:60s} {e0.elapsed_time(e1) * 1e3 / n:8.2f} us", flush=True)
for d in (768, 1024):
    cfg = P.JLConfig(hidden_size=d, num_hidden_layers=1, num_attention_heads=d // 64, intermediate_size=4 * d, adapter_ffn="att")
    model = P.JLForCTC(cfg).cuda().eval()
    eng = model.encoder.engine(model.lm_head)
    ad = model.encoder.layers[0].adapter_ffn
    for B in (1, 4, 32, 128):
        T = 250
        h = torch.randn(B * T, d, device="cuda").to(BF16)
        lengths = torch.full((B,), T, dtype=torch.int32, device="cuda")
        for training in (False, True):
            for fused in (True, False):
                eng.fused_att = fused
                timeit(f"AttAdapter d={d} rows={B*T} {'train' if training else 'infer'} {'fused' if fused else 'composed'}",
                       lambda: eng._adapter_fwd(ad, h, lengths, B, T, training, 0, True))
        out = torch.empty_like(h)
        timeit(f"  (add_bf16, same bytes) rows={B*T}", lambda: ops.add(h, h, out=out))
