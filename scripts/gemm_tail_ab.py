"""A/B of the tail-wave column slices of the single-CTA GEMM kernel (tuning aid): time per launch inside a CUDA graph with
the slices on (jl_debug_set_gemm_tail(2)) and off (0), automatic tile choice and forced single-CTA 256 / 128 tiles."""
import importlib, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
P = importlib.import_module("jiao-liao_speech_recognition_b200")
ops, L = P.ops, P._lib
lib = L.load()
BF16 = torch.bfloat16
dev = "cuda"
g = torch.Generator(device=dev).manual_seed(0)


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


shapes = [(8000, 768, 768), (8000, 768, 3072), (8000, 768, 2304), (8000, 2304, 768), (8000, 3072, 768), (8000, 1024, 1024),
          (8000, 1024, 4096), (8000, 4096, 1024), (8000, 5000, 768), (1000, 768, 768), (1000, 3072, 768), (16000, 768, 768),
          (16000, 768, 3072), (8000, 512, 2560), (8000, 1536, 2560)]
configs = [("auto", 0, 0), ("1cta bn256", 1, 256), ("1cta bn128", 1, 128)]
print("| m x n x k | " + " | ".join(f"{c[0]} whole | {c[0]} sliced" for c in configs) + " |")
print("|---|" + "---:|---:|" * len(configs))
for m, n, k in shapes:
    a, b = rnd(m, k), rnd(n, k)
    bias = torch.zeros(n, device=dev)
    out = torch.empty(m, n, dtype=BF16, device=dev)
    cells = []
    for _, mode, bn in configs:
        for tail in (0, 2):
            lib.jl_debug_set_gemm_mode(mode); lib.jl_debug_set_gemm_bn(bn); lib.jl_debug_set_gemm_tail(tail)
            try:
                cells.append(f"{bench(lambda: ops.gemm(a, b, bias=bias, out=out)):.1f}")
            except Exception:  # noqa: BLE001
                cells.append("err")
            finally:
                lib.jl_debug_set_gemm_mode(0); lib.jl_debug_set_gemm_bn(0); lib.jl_debug_set_gemm_tail(2)
    print(f"| {m} x {n} x {k} | " + " | ".join(cells) + " |", flush=True)
