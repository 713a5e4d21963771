"""Where the end-to-end step time goes on the host: per-phase wall-clock of the bench's e2e loop (submit / step / loss.item())."""
import importlib, os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
P = importlib.import_module("jiao-liao_speech_recognition_b200")
torch.cuda.set_device(0)
cfg = P.JLConfig.base(**bench.WORKLOADS["base"]["model"])
model = P.JLForCTC(cfg).cuda()
model.freeze_base_model()
trainer = P.AdapterTrainer(model, lr=1e-4)
wave, ns, labels, tp = bench.synth_batch(32, 1234, cfg.vocab_size)
wave_p, labels_p, ns_p = wave.pin_memory(), labels.pin_memory(), ns.pin_memory()
for _ in range(5):
    trainer.step(wave_p, ns_p, labels_p).item()
N = 60
T = {k: [] for k in ("check", "step", "submit", "item", "total")}
torch.cuda.synchronize()
trainer.submit(wave_p, ns_p, labels_p)
t_all = time.perf_counter()
for i in range(N):
    t0 = time.perf_counter()
    loss = trainer.step()
    t1 = time.perf_counter()
    trainer.submit(wave_p, ns_p, labels_p)
    t2 = time.perf_counter()
    v = loss.item()
    t3 = time.perf_counter()
    T["step"].append(t1 - t0); T["submit"].append(t2 - t1); T["item"].append(t3 - t2); T["total"].append(t3 - t0)
t_all = time.perf_counter() - t_all
for _ in range(N):
    t0 = time.perf_counter(); trainer._check_weights(); T["check"].append(time.perf_counter() - t0)
for k, v in T.items():
    v = sorted(v)
    print(f"{k:8s} median {1e3 * v[len(v) // 2]:7.3f} ms   p90 {1e3 * v[int(0.9 * len(v))]:7.3f}   max {1e3 * v[-1]:7.3f}")
print(f"loop: {1e3 * t_all / N:.3f} ms/step")
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(N):
    trainer.step_resident()
e1.record(); torch.cuda.synchronize()
print(f"resident: {e0.elapsed_time(e1) / N:.3f} ms/step")
