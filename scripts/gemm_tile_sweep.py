"""Tile sweep for the small / odd GEMM shapes of the fine-tune step (tuning aid): for every shape, the time per launch
inside a CUDA graph of the automatic choice and of every forced (kernel, N tile) combination."""
import importlib, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
P = importlib.import_module("jiao-liao_speech_recognition_b200")
ops, L = P.ops, P._lib
lib = L.load()
BF16, F32 = torch.bfloat16, torch.float32
dev = "cuda"
g = torch.Generator(device=dev).manual_seed(0)
MN, KL = L.JL_LAYOUT_MN, L.JL_LAYOUT_K


def rnd(*shape):
    return (torch.randn(*shape, device=dev, generator=g) * 0.1).to(BF16)


def bench(fn, n=40):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        for _ in range(n):
            fn()
    gr.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for _ in range(3):
        e0.record(); gr.replay(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) * 1e3 / n)
    return best


M = 8000
cases = []
# (name, m, n, k, b_layout, bias, residual, epilogue, aux)
cases.append(("att out-proj  8000x768x64  +bias+res", M, 768, 64, KL, True, True, 0))
cases.append(("att qkv       8000x192x768 +bias", M, 192, 768, KL, True, False, 0))
cases.append(("att dz        8000x768x192 B=MN", M, 768, 192, MN, False, False, 0))
cases.append(("att da        8000x64x768  B=MN", M, 64, 768, MN, False, False, 0))
cases.append(("attn out-proj 8000x768x768 +bias+res", M, 768, 768, KL, True, True, 0))
cases.append(("dgrad         8000x768x768", M, 768, 768, KL, False, False, 0))
cases.append(("qkv           8000x2304x768 +bias", M, 2304, 768, KL, True, False, 0))
cases.append(("dqkv->dx      8000x768x2304", M, 768, 2304, KL, False, False, 0))
cases.append(("fc2           8000x768x3072 +bias+res", M, 768, 3072, KL, True, True, 0))
cases.append(("fc1 gelu      8000x3072x768 +bias", M, 3072, 768, KL, True, False, L.JL_EPI_GELU))
cases.append(("wf t1         8000x32x768", M, 32, 768, KL, False, False, 0))
cases.append(("wf u          8000x256x32 +bias relu", M, 256, 32, KL, True, False, L.JL_EPI_RELU))
cases.append(("wf t2         8000x32x256", M, 32, 256, KL, False, False, 0))
cases.append(("wf out        8000x768x32 +bias+res", M, 768, 32, KL, True, True, 0))

configs = [("auto", 0, 0)] + [("1cta bn%d" % bn, 1, bn) for bn in (32, 64, 128, 256)] + [("pair bn%d" % bn, 2, bn) for bn in (128, 192, 256)]
print("| shape | " + " | ".join(c[0] for c in configs) + " |")
print("|---|" + "---:|" * len(configs))
for name, m, n, k, bl, has_bias, has_res, epi in cases:
    a = rnd(m, k)
    b = rnd(n, k) if bl == KL else rnd(k, n)
    bias = torch.zeros(n, device=dev) if has_bias else None
    res = rnd(m, n) if has_res else None
    out = torch.empty(m, n, dtype=BF16, device=dev)
    aux_out = torch.empty(m, n, dtype=BF16, device=dev) if epi == L.JL_EPI_GELU else None
    cells = []
    for cname, mode, bn in configs:
        if mode == 1 and bn == 32 and bl == MN:
            cells.append("-"); continue
        if mode == 2 and (n < 64 or (bn == 192 and bl == MN)):
            cells.append("-"); continue
        lib.jl_debug_set_gemm_mode(mode); lib.jl_debug_set_gemm_bn(bn)
        try:
            t = bench(lambda: ops.gemm(a, b, bias=bias, residual=res, epilogue=epi, aux_out=aux_out, out=out, b_layout=bl))
            cells.append(f"{t:.1f}")
        except Exception as e:  # noqa: BLE001
            cells.append("err")
        finally:
            lib.jl_debug_set_gemm_mode(0); lib.jl_debug_set_gemm_bn(0)
    print(f"| {name} | " + " | ".join(cells) + " |", flush=True)
