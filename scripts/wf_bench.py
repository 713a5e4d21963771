import importlib, sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
P = importlib.import_module("jiao-liao_speech_recognition_b200")
ops, md, L = P.ops, P.modeling, P._lib
BF16 = torch.bfloat16
cfg = P.JLConfig.base(adapter_ffn="wf")
model = P.JLForCTC(cfg).cuda().eval()
eng = model.encoder.engine(model.lm_head)
ad = model.encoder.layers[0].adapter_ffn
def timeit(name, fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        for _ in range(n): fn()
    gr.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); gr.replay(); e1.record(); torch.cuda.synchronize()
    print(f"{name:50s} {e0.elapsed_time(e1) * 1e3 / n:8.2f} us", flush=True)
for B in (1, 32, 512):
    T = 250
    h = torch.randn(B * T, 768, device="cuda").to(BF16)
    lengths = torch.full((B,), T, dtype=torch.int32, device="cuda")
    eng.fused_wf = True
    timeit(f"WFAdapter fused    rows={B*T}", lambda: eng._adapter_fwd(ad, h, lengths, B, T, False, 0, True))
    eng.fused_wf = False
    timeit(f"WFAdapter composed rows={B*T}", lambda: eng._adapter_fwd(ad, h, lengths, B, T, False, 0, True))
    out = torch.empty_like(h)
    timeit(f"  (reference: add_bf16 same bytes) rows={B*T}", lambda: ops.add(h, h, out=out))
