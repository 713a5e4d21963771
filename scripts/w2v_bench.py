"""Throughput of the raw-waveform (wav2vec2 / XLS-R) front end and of the whole XLS-R-300M-shaped model (24 layers, d = 1024,
front_end = "wav2vec2") with WFAdapter: inference (waveform → token ids, Transcriber graph) and the adapter fine-tune step."""
import importlib, os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
P = importlib.import_module("jiao-liao_speech_recognition_b200")
ops, L = P.ops, P._lib
from bench import synth_batch  # noqa: E402

SECONDS = 10


def ev_time(fn, n):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / 1e3 / n


def main():
    cfg = P.JLConfig.xlsr(adapter_ffn="wf", vocab_size=5000)
    model = P.JLForCTC(cfg).cuda()
    model.freeze_base_model()
    eng = model.encoder.engine(model.lm_head)
    print("| what | batch | ms | audio-s/s | launches |\n|---|---:|---:|---:|---:|", flush=True)
    for b in (4, 32):
        wave, ns, labels, tp = synth_batch(b, 99, cfg.vocab_size)
        wd, nsd = wave.cuda(), ns.cuda()
        lens = eng.output_lengths(wd, frame_lengths=nsd)
        fz = eng._frozen_pack()
        with torch.no_grad():
            L.launch_count_reset(); eng._wav2vec2_front_end(wd, nsd, lens, fz); nl = L.launch_count()
            t = ev_time(lambda: eng._wav2vec2_front_end(wd, nsd, lens, fz), 10)
        print(f"| front end only (eager) | {b} | {1e3 * t:.2f} | {b * SECONDS / t:.0f} | {nl} |", flush=True)
        tr = P.Transcriber(model)
        wp = wave.pin_memory()
        tr(wp, ns)
        t = ev_time(tr.run_resident, 10)
        print(f"| inference, waveform → ids (graph, resident) | {b} | {1e3 * t:.2f} | {b * SECONDS / t:.0f} | {tr.launches_per_step} |", flush=True)
        t = ev_time(lambda: tr(wp, ns)[0].cpu(), 10)
        print(f"| inference end to end (pinned waveforms in, ids out) | {b} | {1e3 * t:.2f} | {b * SECONDS / t:.0f} | |", flush=True)
    b = 16
    wave, ns, labels, tp = synth_batch(b, 98, cfg.vocab_size)
    labels = labels[:, : int(0.4 * 499)]
    trn = P.AdapterTrainer(model, lr=1e-4, comm=None)
    wp, lp = wave.pin_memory(), labels.pin_memory()
    for _ in range(3):
        loss = trn.step(wp, ns, lp).item()
    t = ev_time(trn.step_resident, 10)
    print(f"| fine-tune step (WFAdapter in 24 layers + lm_head, graph, resident) | {b} | {1e3 * t:.2f} | {b * SECONDS / t:.0f} | {trn.launches_per_step + 1} |", flush=True)
    print(f"\nloss {loss:.2f}; peak memory {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB")


if __name__ == "__main__":
    main()
