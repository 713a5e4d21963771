"""Micro-benchmarks (tuning aid): per-launch cost of small kernels back to back in a stream and inside a CUDA graph."""
import importlib, sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
P = importlib.import_module("jiao-liao_speech_recognition_b200")
ops, L = P.ops, P._lib
BF16, F32 = torch.bfloat16, torch.float32
dev = "cuda"
M, d = 8000, 768
g = torch.Generator(device=dev).manual_seed(0)
h = torch.randn(M, d, device=dev, generator=g).to(BF16)
gamma, beta = torch.ones(d, device=dev), torch.zeros(d, device=dev)
a64 = torch.randn(M, 64, device=dev, generator=g).to(BF16)
wo = torch.randn(d, 64, device=dev, generator=g).to(BF16)
wq = torch.randn(3 * d, d, device=dev, generator=g).to(BF16)
bo = torch.zeros(d, device=dev)
lens = torch.full((32,), 250, dtype=torch.int32, device=dev)
tiny = torch.randn(256, 64, device=dev, generator=g).to(BF16)

def timeit(name, fn, n=50):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    t_stream = e0.elapsed_time(e1) * 1e3 / n
    gr = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        fn()
    torch.cuda.current_stream().wait_stream(s); torch.cuda.synchronize()
    with torch.cuda.graph(gr):
        for _ in range(n): fn()
    gr.replay(); torch.cuda.synchronize()
    e0.record(); gr.replay(); e1.record(); torch.cuda.synchronize()
    t_graph = e0.elapsed_time(e1) * 1e3 / n
    print(f"{name:44s} stream {t_stream:7.2f} us/launch   graph {t_graph:7.2f} us/launch", flush=True)

out = torch.empty(M, d, dtype=BF16, device=dev)
timeit("colsum tiny [256x64]", lambda: ops.colsum(tiny))
timeit("layernorm_fwd [8000x768]", lambda: ops.layernorm_fwd(h, gamma, beta, out=out))
timeit("add_bf16 [8000x768]", lambda: ops.add(h, h, out=out))
timeit("gemm 8000x768x64 +bias+res", lambda: ops.gemm(a64, wo, bias=bo, residual=h, out=out))
timeit("gemm 8000x768x64 plain", lambda: ops.gemm(a64, wo, out=out))
timeit("gemm 256x64x64 tiny", lambda: ops.gemm(tiny, tiny[:64]))
qkv = torch.empty(M, 3 * d, dtype=BF16, device=dev)
timeit("gemm 8000x2304x768", lambda: ops.gemm(h, wq, out=qkv))
L.load().jl_debug_set_gemm_mode(1)
timeit("gemm 8000x2304x768 (1-CTA kernel)", lambda: ops.gemm(h, wq, out=qkv))
timeit("gemm 8000x768x64 +bias+res (1-CTA kernel)", lambda: ops.gemm(a64, wo, bias=bo, residual=h, out=out))
L.load().jl_debug_set_gemm_mode(0)
q1 = qkv[:, :64]; 
timeit("attn_fwd 1 head", lambda: ops.attn_fwd(qkv[:, 0:64], qkv[:, 64:128], qkv[:, 128:192], lens, 32, 250, 1, 0.125))
timeit("attn_fwd 12 heads", lambda: ops.attn_fwd(qkv[:, 0:d], qkv[:, d:2*d], qkv[:, 2*d:], lens, 32, 250, 12, 0.125))
L.load().jl_debug_set_attn_impl(2)
timeit("attn_fwd 12 heads (tc, 3 CTAs/SM build)", lambda: ops.attn_fwd(qkv[:, 0:d], qkv[:, d:2*d], qkv[:, 2*d:], lens, 32, 250, 12, 0.125))
L.load().jl_debug_set_attn_impl(0)
o12, lse12 = ops.attn_fwd(qkv[:, 0:d], qkv[:, d:2*d], qkv[:, 2*d:], lens, 32, 250, 12, 0.125, want_lse=True)
do12 = torch.randn_like(o12)
timeit("attn_bwd 12 heads (tc)", lambda: ops.attn_bwd(qkv[:, 0:d], qkv[:, d:2*d], qkv[:, 2*d:], o12, do12, lse12, lens, 32, 250, 12, 0.125))
o1, lse1 = ops.attn_fwd(qkv[:, 0:64], qkv[:, 64:128], qkv[:, 128:192], lens, 32, 250, 1, 0.125, want_lse=True)
do1 = torch.randn_like(o1)
timeit("attn_bwd 1 head (tc)", lambda: ops.attn_bwd(qkv[:, 0:64], qkv[:, 64:128], qkv[:, 128:192], o1, do1, lse1, lens, 32, 250, 1, 0.125))
L.load().jl_debug_set_attn_impl(3)
timeit("attn_bwd 12 heads (tc, dQ + dKV kernels)", lambda: ops.attn_bwd(qkv[:, 0:d], qkv[:, d:2*d], qkv[:, 2*d:], o12, do12, lse12, lens, 32, 250, 12, 0.125))
timeit("attn_bwd 1 head (tc, dQ + dKV kernels)", lambda: ops.attn_bwd(qkv[:, 0:64], qkv[:, 64:128], qkv[:, 128:192], o1, do1, lse1, lens, 32, 250, 1, 0.125))
L.load().jl_debug_set_attn_impl(1)
timeit("attn_bwd 12 heads (mma.sync)", lambda: ops.attn_bwd(qkv[:, 0:d], qkv[:, d:2*d], qkv[:, 2*d:], o12, do12, lse12, lens, 32, 250, 12, 0.125))
timeit("attn_fwd 12 heads (mma.sync)", lambda: ops.attn_fwd(qkv[:, 0:d], qkv[:, d:2*d], qkv[:, 2*d:], lens, 32, 250, 12, 0.125))
L.load().jl_debug_set_attn_impl(0)
