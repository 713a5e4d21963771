"""Launches the FFN-shaped GEMM once per epilogue variant (after a warm-up round) for an ncu capture (tuning aid)."""
import importlib, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
P = importlib.import_module("jiao-liao_speech_recognition_b200")
ops, L = P.ops, P._lib
lib = L.load()
BF16 = torch.bfloat16
g = torch.Generator(device="cuda").manual_seed(0)
rnd = lambda *s: (torch.randn(*s, device="cuda", generator=g) * 0.1).to(BF16)
m, n, k = 8000, 3072, 768
a, b = rnd(m, k), rnd(n, k)
bias = torch.zeros(n, device="cuda")
out = torch.empty(m, n, dtype=BF16, device="cuda")
aux = torch.empty(m, n, dtype=BF16, device="cuda")
pre = rnd(m, n)
a2, b2 = rnd(m, 3072), rnd(768, 3072)
out2 = torch.empty(m, 768, dtype=BF16, device="cuda")
res = rnd(m, 768)
bias2 = torch.zeros(768, device="cuda")
lib.jl_debug_set_gemm_mode(1); lib.jl_debug_set_gemm_bn(256)
for rep in range(2):
    ops.gemm(a, b, bias=bias, out=out)                                                   # 0: bias only
    ops.gemm(a, b, bias=bias, epilogue=L.JL_EPI_GELU, out=out)                           # 1: gelu
    ops.gemm(a, b, bias=bias, epilogue=L.JL_EPI_GELU_DGELU, aux_out=aux, out=out)        # 2: gelu + gelu'
    ops.gemm(a, b, epilogue=L.JL_EPI_MUL_AUX, aux=pre, out=out)                          # 3: x aux
    ops.gemm(a2, b2, bias=bias2, residual=res, out=out2)                                 # 4: 8000 x 768 x 3072 + bias + residual
    torch.cuda.synchronize()
