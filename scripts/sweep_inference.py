"""BASELINE.json config 5: streaming-free inference sweep, batch 1–512 × 10 s: pinned-host waveforms → mel + encoder (WFAdapter) +
CTC greedy decode → token ids on the host.  Prints one markdown row per batch size (device-resident and end-to-end audio-s/s)."""
import importlib, os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import synth_batch  # noqa: E402
P = importlib.import_module("jiao-liao_speech_recognition_b200")
cfg = P.JLConfig.base(adapter_ffn="wf")
model = P.JLForCTC(cfg).cuda().eval()
print("| batch | ms/step (resident) | audio-s/s (resident) | audio-s/s (e2e: H2D waveforms + D2H ids) | RTF | launches |\n|---:|---:|---:|---:|---:|---:|")
for b in (1, 2, 4, 8, 16, 32, 64, 128, 256, 512):
    tr = P.Transcriber(model, use_cuda_graph=True)
    wave, ns, _, _ = synth_batch(b, 1234, cfg.vocab_size)
    wave = wave.pin_memory()
    for _ in range(3):
        ids, n = tr(wave, ns)
        ids.cpu()
    torch.cuda.synchronize()
    steps = 20 if b <= 128 else 8
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        tr.run_resident()
    e1.record(); torch.cuda.synchronize()
    t_res = e0.elapsed_time(e1) / 1e3 / steps
    t0 = time.perf_counter()
    for _ in range(steps):
        ids, n = tr(wave, ns)
        host = ids.cpu(); n.cpu()
    torch.cuda.synchronize()
    t_e2e = (time.perf_counter() - t0) / steps
    print(f"| {b} | {t_res * 1e3:.2f} | {b * 10 / t_res:,.0f} | {b * 10 / t_e2e:,.0f} | {t_res / (b * 10):.2e} | {tr.launches_per_step} |", flush=True)
    del tr
    torch.cuda.empty_cache()
