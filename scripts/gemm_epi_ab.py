"""Where the FFN GEMM epilogues spend their time (tuning aid): 8000 x 3072 x 768 (+ the dgrad shape 8000 x 3072 x 768 with the
activation-gradient epilogue) with bias only / ReLU / GELU without and with the pre-activation copy, per kernel and N tile."""
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


m, n, k = 8000, 3072, 768
a, b = rnd(m, k), rnd(n, k)
bias = torch.zeros(n, device=dev)
out = torch.empty(m, n, dtype=BF16, device=dev)
aux = torch.empty(m, n, dtype=BF16, device=dev)
pre = rnd(m, n)
variants = [
    ("bias", dict(bias=bias)),
    ("bias+relu", dict(bias=bias, epilogue=L.JL_EPI_RELU)),
    ("bias+gelu", dict(bias=bias, epilogue=L.JL_EPI_GELU)),
    ("bias+gelu+pre-activation copy", dict(bias=bias, epilogue=L.JL_EPI_GELU, aux_out=aux)),
    ("bias+gelu, gelu' saved (training forward)", dict(bias=bias, epilogue=L.JL_EPI_GELU_DGELU, aux_out=aux)),
    ("x saved gelu' (training backward)", dict(epilogue=L.JL_EPI_MUL_AUX, aux=pre)),
    ("relu' (reads pre-activation)", dict(epilogue=L.JL_EPI_RELU_BWD, aux=pre)),
    ("gelu' (reads pre-activation)", dict(epilogue=L.JL_EPI_GELU_BWD, aux=pre)),
]
configs = [("auto", 0, 0), ("1cta bn256", 1, 256), ("1cta bn128", 1, 128), ("pair bn192", 2, 192), ("pair bn256", 2, 256), ("pair bn128", 2, 128)]
print(f"{m} x {n} x {k}, us per launch in a CUDA graph")
print("| epilogue | " + " | ".join(c[0] for c in configs) + " |")
print("|---|" + "---:|" * len(configs))
for name, kw in variants:
    cells = []
    for _, mode, bn in configs:
        lib.jl_debug_set_gemm_mode(mode); lib.jl_debug_set_gemm_bn(bn)
        try:
            cells.append(f"{bench(lambda: ops.gemm(a, b, out=out, **kw)):.1f}")
        except Exception as e:  # noqa: BLE001
            cells.append("err")
        finally:
            lib.jl_debug_set_gemm_mode(0); lib.jl_debug_set_gemm_bn(0)
    print(f"| {name} | " + " | ".join(cells) + " |", flush=True)
