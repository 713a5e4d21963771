"""In-graph timing of jl_lnproj_bwd vs jl_gemm_bf16 (dy · W) + jl_layernorm_bwd."""
import importlib, sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
P = importlib.import_module("jiao-liao_speech_recognition_b200")
ops = P.ops
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
for rows, d in ((8000, 768), (8000, 1024), (32000, 768)):
    n = 192
    h = torch.randn(rows, d, device="cuda").to(BF16)
    w = (torch.randn(n, d, device="cuda") * 0.05).to(BF16)
    bias = torch.zeros(n, device="cuda")
    gamma, beta = torch.ones(d, device="cuda"), torch.zeros(d, device="cuda")
    dy = torch.randn(rows, n, device="cuda").to(BF16)
    y = torch.randn(rows, n, device="cuda").to(BF16)
    dres = torch.randn(rows, d, device="cuda").to(BF16)
    z, mean, rstd = ops.layernorm_fwd(h, gamma, beta, 1e-5, save_stats=True)
    pack = ops.lnfold_pack(w, bias, gamma, beta)
    timeit(f"lnproj_bwd rows={rows} d={d} (+dz)", lambda: ops.lnproj_bwd(dy, y, w, pack, gamma, h, mean, rstd, dres, want_dz=True))
    timeit(f"lnproj_bwd rows={rows} d={d}", lambda: ops.lnproj_bwd(dy, y, w, pack, gamma, h, mean, rstd, dres))
    def two():
        dz = ops.gemm(dy, w, b_layout=1)
        ops.layernorm_bwd(dz, h, gamma, mean, rstd, dres=dres)
    timeit(f"gemm + layernorm_bwd rows={rows} d={d}", two)
