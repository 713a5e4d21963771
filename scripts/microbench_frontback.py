"""Micro-benchmark of the front end (mel + CMVN) and the CTC stage at the bench workload (32 x 10 s, V = 5000, S = 100):
warm per-call time in a stream and in a CUDA graph, with the number of launches per call."""
import importlib, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
P = importlib.import_module("jiao-liao_speech_recognition_b200")
ops, L = P.ops, P._lib
dev = "cuda"
g = torch.Generator(device=dev).manual_seed(0)
B, N, T, V, S = 32, 160000, 250, 5000, 100
wave = (0.1 * torch.randn(B, N, device=dev, generator=g)).contiguous()
ns = torch.full((B,), N, dtype=torch.int32, device=dev)
fe = P.JLFeatureExtractor(device=dev)
logits = torch.randn(B, T, V, device=dev, generator=g)
labels = torch.randint(1, V, (B, S), device=dev, generator=g, dtype=torch.int32)
lens = torch.full((B,), T - 1, dtype=torch.int32, device=dev)


def timeit(name, fn, n=30):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    L.launch_count_reset(); fn(); nl = L.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    ts = e0.elapsed_time(e1) * 1e3 / n
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        for _ in range(n):
            fn()
    gr.replay(); torch.cuda.synchronize()
    e0.record(); gr.replay(); e1.record(); torch.cuda.synchronize()
    tg = e0.elapsed_time(e1) * 1e3 / n
    print(f"{name:40s} launches {nl:2d}  stream {ts:8.2f} us  graph {tg:8.2f} us", flush=True)


timeit("mel+cmvn 32 x 10 s", lambda: fe.extract_device(wave, ns, return_bf16=True))
timeit("ctc loss+grad [32,250,5000] S=100", lambda: ops.ctc_loss(logits, labels, lens, want_grad=True, grad_dtype=torch.bfloat16))
timeit("ctc loss only", lambda: ops.ctc_loss(logits, labels, lens, want_grad=False))
timeit("ctc greedy", lambda: ops.ctc_greedy(logits, lens))
