"""A few eager launches of the fused AttAdapter forward (ncu target)."""
import importlib, sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
P = importlib.import_module("jiao-liao_speech_recognition_b200")
d, B, T = 768, 32, 250
cfg = P.JLConfig(hidden_size=d, num_hidden_layers=1, num_attention_heads=d // 64, intermediate_size=4 * d, adapter_ffn="att")
model = P.JLForCTC(cfg).cuda().eval()
eng = model.encoder.engine(model.lm_head)
ad = model.encoder.layers[0].adapter_ffn
eng.fused_att = True
h = torch.randn(B * T, d, device="cuda").to(torch.bfloat16)
lengths = torch.full((B,), T, dtype=torch.int32, device="cuda")
for training in (False, True, False, True):
    eng._adapter_fwd(ad, h, lengths, B, T, training, 0, True)
torch.cuda.synchronize()
print("ok")
