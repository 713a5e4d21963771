import importlib, sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
P = importlib.import_module("jiao-liao_speech_recognition_b200")
ops = P.ops
g = torch.Generator(device="cuda").manual_seed(0)
d = 768
qkv = torch.randn(8000, 3 * d, device="cuda", generator=g).to(torch.bfloat16)
lens = torch.full((32,), 250, dtype=torch.int32, device="cuda")
for _ in range(3):
    o, lse = ops.attn_fwd(qkv[:, 0:d], qkv[:, d:2*d], qkv[:, 2*d:], lens, 32, 250, 12, 0.125, want_lse=True)
    do = torch.randn_like(o)
    dq = ops.attn_bwd(qkv[:, 0:d], qkv[:, d:2*d], qkv[:, 2*d:], o, do, lse, lens, 32, 250, 12, 0.125)
torch.cuda.synchronize()
print("ok")
