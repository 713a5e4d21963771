import importlib, sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
P = importlib.import_module("jiao-liao_speech_recognition_b200")
ops, L = P.ops, P._lib
BF16 = torch.bfloat16
g = torch.Generator(device="cuda").manual_seed(0)
def bench(m, n, k, bmn=False):
    a = torch.randn(m, k, device="cuda", generator=g).to(BF16)
    b = torch.randn((k, n) if bmn else (n, k), device="cuda", generator=g).to(BF16)
    out = torch.empty(m, n, dtype=BF16, device="cuda")
    res = {}
    for mode in (1, 2, 3):
        L.load().jl_debug_set_gemm_mode(mode)
        for _ in range(3): ops.gemm(a, b, out=out, b_layout=int(bmn))
        gr = torch.cuda.CUDAGraph()
        torch.cuda.synchronize()
        with torch.cuda.graph(gr):
            for _ in range(20): ops.gemm(a, b, out=out, b_layout=int(bmn))
        gr.replay(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); gr.replay(); e1.record(); torch.cuda.synchronize()
        res[mode] = e0.elapsed_time(e1) * 1e3 / 20
    L.load().jl_debug_set_gemm_mode(0)
    fl = 2.0 * m * n * k
    print(f"{m}x{n}x{k} B={'MN' if bmn else 'K'}: " + "  ".join(f"mode{md} {t:6.1f} us {fl / t / 1e6:5.0f} TF" for md, t in res.items()), flush=True)
for shp in [(8000, 2304, 768), (8000, 3072, 768), (8000, 768, 3072), (8000, 768, 768), (8000, 768, 2304), (8192, 8192, 8192), (8000, 5000, 768), (16000, 2304, 768)]:
    bench(*shp)
bench(8000, 768, 3072, True); bench(8000, 3072, 768, True)
