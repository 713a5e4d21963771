#!/usr/bin/env python
"""Headline benchmark: audio-seconds/sec of the adapter fine-tune step (BASELINE.json configs[1]: 12-layer d=768
encoder with AttAdapter, CTC loss + adapter-only backward, batch 32 × 10 s of synthetic 16 kHz audio per B200).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path (one process per GPU under torchrun)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU implementation of the same step (oracle port)

One JSON line on stdout (rank 0).  `value` = whole-job audio-s/s with the waveforms already resident in HBM;
`e2e` = the same step through the public API with pinned-host waveforms/labels copied in and the loss read back every
step; `roofline` = the step's tcgen05 GEMM launches re-issued on their own and timed with CUDA events against the
measured bf16 peak; `cpu_baseline` = the oracle timed on this box's host cores on a bounded sample of the workload.
"""
from __future__ import annotations

import argparse
import importlib
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

PKG = "jiao-liao_speech_recognition_b200"
SR = 16000
SECONDS = 10
BATCH = 32            # per GPU (weak scaling)
CPU_SAMPLE_BATCH = 2  # bounded sample of the workload for the CPU arm
METRIC = "audio-seconds/sec (adapter fine-tune step: fwd + CTC loss + adapter-only bwd + allreduce + AdamW)"
UNIT = "audio-s/s"


def synth_batch(batch: int, seed: int, vocab: int):
    """SURVEY §8d synthetic inputs: 0.1·randn + 0.05·Σ_5 sin, clipped; labels S = ⌊0.4·T'⌋ ~ U{1..V-1}."""
    g = torch.Generator().manual_seed(seed)
    n = SR * SECONDS
    t = torch.arange(n, dtype=torch.float32) / SR
    wave = 0.1 * torch.randn(batch, n, generator=g)
    for _ in range(5):
        f = 100.0 + 3900.0 * torch.rand(batch, 1, generator=g)
        ph = 6.283185307179586 * torch.rand(batch, 1, generator=g)
        wave += 0.05 * torch.sin(6.283185307179586 * f * t[None, :] + ph)
    wave.clamp_(-1.0, 1.0)
    frames = 1 + (n - 400) // 160
    tp = ((frames - 1) // 2 + 1 - 1) // 2 + 1
    s = int(0.4 * tp)
    labels = torch.randint(1, vocab, (batch, s), generator=g, dtype=torch.int32)
    ns = torch.full((batch,), n, dtype=torch.int32)
    return wave, ns, labels, tp


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return {"hbm_gbs": p["hbm_gbs"], "bf16_tflops": p["bf16_tflops"], "bf16_tflops_sustained": p["bf16_tflops_sustained"], "src": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "src": "fallback"}


class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index: int):
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            out = ""
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in out.strip().splitlines():
            parts = [x.strip() for x in ln.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx = float(parts[1])
            except ValueError:
                continue
            for nm, v in zip(names, parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------------------------ CPU arm
def cpu_finetune_step_rate(batch: int, repeats: int, warmup: int, threads: int):
    """The oracle's fine-tune step (fp32, autograd, adapter-only grads) on `batch` × 10 s — returns (audio-s/s, seconds/step)."""
    from oracle import model as om
    torch.set_num_threads(threads)
    cfg = om.OracleConfig(adapter_ffn="att")
    w = om.init_weights(cfg, seed=0)
    for k, v in w.items():
        v.requires_grad_(om.is_trainable(k))
    wave, ns, labels, tp = synth_batch(batch, 1234, cfg.vocab_size)
    waves = [wave[i] for i in range(batch)]
    lab = labels.to(torch.int64)
    best = None
    for it in range(warmup + repeats):
        t0 = time.perf_counter()
        loss, _, _ = om.forward_from_waveforms(w, cfg, waves, lab)
        loss.backward()
        for v in w.values():
            v.grad = None
        dt = time.perf_counter() - t0
        if it >= warmup:
            best = dt if best is None else min(best, dt)
    return batch * SECONDS / best, best


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    from oracle import model as om
    cfg = om.OracleConfig(adapter_ffn="att")
    w = om.init_weights(cfg, seed=0)
    for k, v in w.items():
        v.requires_grad_(om.is_trainable(k))
    wave, ns, labels, tp = synth_batch(CPU_SAMPLE_BATCH, 1234, cfg.vocab_size)
    waves = [wave[i] for i in range(CPU_SAMPLE_BATCH)]
    lab = labels.to(torch.int64)

    def step():
        loss, _, _ = om.forward_from_waveforms(w, cfg, waves, lab)
        loss.backward()
        for v in w.values():
            v.grad = None

    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    value = CPU_SAMPLE_BATCH * SECONDS * args.steps / dt
    sample = f"each step = oracle fine-tune step on {CPU_SAMPLE_BATCH} x {SECONDS} s of the {BATCH} x {SECONDS} s batch (fp32, torch CPU ops, {threads} threads)"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": workload_config(args.gpus),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


def workload_config(n_gpus: int, trainable: int = 6234248):
    return {"workload": "BASELINE.json configs[1]: 12-layer d=768 h=12 ffn=3072 encoder (80-mel, 2x conv k5 s2 + GLU) with AttAdapter(b=64) after "
                        "each FFN, V=5000 CTC head; fine-tune step (CTC loss, adapter-only backward, frozen backbone), "
                        f"batch {BATCH} x {SECONDS} s synthetic 16 kHz audio per GPU",
            "per_gpu_batch": BATCH, "global_batch": BATCH * n_gpus, "seconds_per_utterance": SECONDS, "labels_per_utterance": 100,
            "parallelism": f"dp{n_gpus}", "collective": f"one NCCL all-reduce of the {trainable / 1e6:.2f} M-parameter adapter + lm_head fp32 gradient bucket per step",
            "l2_policy": "inputs larger than L2: each step streams ~2 GB of activations (126 MB L2), no explicit flush"}


# ------------------------------------------------------------------------------------------------------------------ GPU arm
def inference_rate(P, steps: int, batch: int = 4):
    """BASELINE.json configs[0] (the reference's own CPU-runnable case): base encoder + WFAdapter, batch 4 x 10 s, waveform →
    greedy token ids through `Transcriber` (one CUDA graph).  Resident = inputs already in HBM; e2e = pinned host waveforms in,
    token ids back on the host, every call."""
    cfg = P.JLConfig.base(adapter_ffn="wf")
    model = P.JLForCTC(cfg).cuda().eval()
    tr = P.Transcriber(model)
    wave, ns, _, _ = synth_batch(batch, 4321, cfg.vocab_size)
    wave_p = wave.pin_memory()
    for _ in range(3):
        ids, n = tr(wave_p, ns)
        ids.cpu()
    e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    torch.cuda.synchronize()
    e[0].record()
    for _ in range(steps):
        tr.run_resident()
    e[1].record()
    torch.cuda.synchronize()
    e[2].record()
    for _ in range(steps):
        ids, n = tr(wave_p, ns)
        ids_h, n_h = ids.cpu(), n.cpu()
    e[3].record()
    torch.cuda.synchronize()
    t_res, t_e2e = e[0].elapsed_time(e[1]) / 1e3 / steps, e[2].elapsed_time(e[3]) / 1e3 / steps
    return {"workload": f"BASELINE.json configs[0]: base encoder + WFAdapter(b=256, r=32), inference, batch {batch} x {SECONDS} s, greedy CTC decode",
            "value": batch * SECONDS / t_res, "unit": UNIT, "ms_per_batch": 1e3 * t_res, "rtf": t_res / (batch * SECONDS),
            "e2e": {"value": batch * SECONDS / t_e2e, "unit": UNIT, "ms_per_batch": 1e3 * t_e2e, "h2d_bytes_per_step": wave_p.numel() * 4 + ns.numel() * 4,
                    "d2h_bytes_per_step": int(ids_h.numel() * 4 + n_h.numel() * 4)},
            "gpu_launches_per_batch": tr.launches_per_step}


def run_ours(args):
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — this path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    P = importlib.import_module(PKG)
    if os.environ.get("JL_PDL") in ("0", "1"):     # tuning aid: programmatic dependent launch off / on (default: on)
        P._lib.load().jl_debug_set_pdl(int(os.environ["JL_PDL"]))
    cfg = P.JLConfig.base(adapter_ffn="att")
    model = P.JLForCTC(cfg).cuda()
    model.freeze_base_model()
    trainer = P.AdapterTrainer(model, lr=1e-4, use_cuda_graph=not args.eager)
    wave, ns, labels, tp = synth_batch(BATCH, 1234 + rank, cfg.vocab_size)
    wave_p, labels_p = wave.pin_memory(), labels.pin_memory()
    audio_s_per_step = BATCH * SECONDS * world

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up through the public path (captures the CUDA graph on the first call)
    for _ in range(max(args.warmup, 3)):
        loss = trainer.step(wave_p, ns, labels_p)
        loss_val = float(loss.item())
    launches_per_step = trainer.launches_per_step + 1   # + fused AdamW (the all-reduce is NCCL's kernel, not counted)

    sampler = ClockSampler(local) if rank == 0 else None
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    # ---- device-resident throughput: graph replay + all-reduce + AdamW, inputs already in HBM
    barrier()
    ev[0].record()
    for _ in range(args.steps):
        trainer.step_resident()
    ev[1].record()
    barrier()
    t_res = ev[0].elapsed_time(ev[1]) / 1e3
    # ---- end to end: pinned host waveforms + labels in, loss out, every step.  The public API is used the way a
    # prefetching data loader drives it: submit(batch i+1) stages the next host → device copy on a copy stream while step i
    # runs; every step's copy (K of them) and every step's loss read-back are inside the timed region.
    ns_p = ns.pin_memory()
    barrier()
    ev[2].record()
    trainer.submit(wave_p, ns_p, labels_p)
    for i in range(args.steps):
        loss = trainer.step()
        if i + 1 < args.steps:
            trainer.submit(wave_p, ns_p, labels_p)
        loss_val = float(loss.item())
    ev[3].record()
    barrier()
    t_e2e = ev[2].elapsed_time(ev[3]) / 1e3
    # the same without overlap (copy, then compute, on one stream) for comparison
    barrier()
    ev[2].record()
    for _ in range(args.steps):
        loss = trainer.step(wave_p, ns_p, labels_p)
        loss_val = float(loss.item())
    ev[3].record()
    barrier()
    t_e2e_serial = ev[2].elapsed_time(ev[3]) / 1e3
    clocks = sampler.stop() if sampler else None
    if world > 1:
        tt = torch.tensor([t_res, t_e2e, t_e2e_serial], device="cuda", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        t_res, t_e2e, t_e2e_serial = float(tt[0]), float(tt[1]), float(tt[2])

    # ---- roofline of the dominant kernel (tcgen05 GEMM): the step's GEMM launches re-issued on their own
    peaks = load_peaks()
    trace = trainer.trace_gemms()
    flops = sum(f for _, _, f in trace)
    torch.cuda.synchronize()
    g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    P.ops.replay_gemm_trace(trace)
    torch.cuda.synchronize()
    reps = 5
    g0.record()
    for _ in range(reps):
        P.ops.replay_gemm_trace(trace)
    g1.record()
    torch.cuda.synchronize()
    t_gemm = g0.elapsed_time(g1) / 1e3 / reps
    achieved = flops / t_gemm / 1e12
    traffic, traffic_note = None, None
    import glob
    tpaths = sorted(glob.glob(os.path.join(ROOT, "profiles", "*_gemm_traffic.json")))     # newest capture wins (r1f < r1i < …)
    tpath = tpaths[-1] if tpaths else ""
    if tpath:
        with open(tpath) as f:
            tj = json.load(f)
        traffic = tj["dram_bytes"]
        traffic_note = (f"ncu dram read+write of one launch of the shape with the largest share of the step {tj['shape_mnk']}: "
                        f"{tj['dram_bytes'] / 1e6:.1f} MB vs {tj['algorithmic_bytes'] / 1e6:.1f} MB algorithmic ({tj['source']})")
    roofline = {"bound": "tensor", "kernel": "gemm_tcgen05_kernel<256> (128 x 256 tiles) / gemm_tcgen05_2cta_kernel<192> (256 x 192 CTA-pair tiles)", "achieved": achieved, "peak": peaks["bf16_tflops_sustained"],
                "unit": "TFLOP/s", "frac": achieved / peaks["bf16_tflops_sustained"], "traffic": traffic, "traffic_note": traffic_note,
                "peak_source": f"{peaks['src']} bf16_tflops_sustained (the {len(trace)} GEMM launches of one step are timed back to back)",
                "launches_per_step": len(trace), "algorithmic_tflop_per_step": flops / 1e12, "gemm_ms_per_step": 1e3 * t_gemm,
                "share_of_step": t_gemm / (t_res / args.steps)}

    # per-shape timing of the step's GEMM launches: the encoder's dense contractions (M = B·T' rows, K and N >= 768) are
    # reported on their own next to the all-launch aggregate (which includes the latency-bound low-rank adapter products)
    if rank == 0:
        import ctypes as C
        lib = P._lib.load()
        groups = {}
        for prm, keep, fl in trace:
            key = (prm.m, prm.n, prm.k, prm.a_layout, prm.b_layout, prm.epilogue, prm.out_dtype)
            groups.setdefault(key, [prm, 0, fl])[1] += 1
        rows = []
        s_ = torch.cuda.current_stream().cuda_stream
        for key, (prm, cnt, fl) in groups.items():
            for _ in range(3):
                lib.jl_gemm_bf16(C.byref(prm), s_)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                lib.jl_gemm_bf16(C.byref(prm), s_)
            e1.record()
            torch.cuda.synchronize()
            us = e0.elapsed_time(e1) * 1e3 / 10
            rows.append((us * cnt, key, cnt, us, fl / us / 1e6, fl))
        rows.sort(reverse=True)
        big = [r for r in rows if min(r[1][0], r[1][1], r[1][2]) >= 768]
        big_fl = sum(r[5] * r[2] for r in big)
        big_us = sum(r[0] for r in big)
        roofline["encoder_gemms"] = {"achieved": big_fl / big_us / 1e6, "frac": big_fl / big_us / 1e6 / peaks["bf16_tflops_sustained"],
                                     "unit": "TFLOP/s", "launches_per_step": sum(r[2] for r in big), "us_per_step": big_us,
                                     "note": "shapes with min(M, N, K) >= 768, each timed warm and back to back"}
        roofline["top_shapes"] = [{"m": r[1][0], "n": r[1][1], "k": r[1][2], "launches": r[2], "us": round(r[3], 1), "tflops": round(r[4])}
                                  for r in rows[:6]]
        if args.gemm_breakdown:
            with open(args.gemm_breakdown, "w") as f:
                f.write("| m | n | k | A | B | epi | out | launches/step | us/launch (warm, back-to-back) | TFLOP/s | us/step |\n|---|---|---|---|---|---|---|---:|---:|---:|---:|\n")
                for tot, key, cnt, us, tf, _ in rows:
                    f.write(f"| {key[0]} | {key[1]} | {key[2]} | {'MN' if key[3] else 'K'} | {'MN' if key[4] else 'K'} | {key[5]} | "
                            f"{'f32' if key[6] else 'bf16'} | {cnt} | {us:.1f} | {tf:.0f} | {tot:.0f} |\n")
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    # ---- secondary figure (not the headline): inference forward waveform → token ids on BASELINE.json configs[0]
    inference = None
    if world == 1 and not args.no_inference:
        inference = inference_rate(P, steps=max(args.steps, 10))
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        v, sec = cpu_finetune_step_rate(CPU_SAMPLE_BATCH, repeats=2, warmup=1, threads=threads)
        cpu = {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
               "sample": f"oracle fine-tune step on {CPU_SAMPLE_BATCH} x {SECONDS} s of the {BATCH} x {SECONDS} s batch, best of 2 after 1 warm-up ({sec:.2f} s/step)"}
    h2d = wave_p.numel() * 4 + labels_p.numel() * 4 + ns.numel() * 4 * 2
    line = {
        "metric": METRIC, "value": audio_s_per_step * args.steps / t_res, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": 1e3 * t_res / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
        "data": "synthetic", "config": workload_config(world, trainer.flat.num_params),
        "e2e": {"value": audio_s_per_step * args.steps / t_e2e, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                "ms_per_step": 1e3 * t_e2e / args.steps, "api": "AdapterTrainer.submit(next batch) + step() + loss.item()",
                "serial_value": audio_s_per_step * args.steps / t_e2e_serial,
                "serial_note": "AdapterTrainer.step(batch) with the copy and the kernels on one stream (no prefetch)"},
        "gpu_launches": launches_per_step * args.steps * 2, "gpu_launches_per_step": launches_per_step,
        "rtf": (t_res / args.steps) / audio_s_per_step, "loss": loss_val, "cuda_graph": not args.eager,
        "roofline": roofline, "cpu_baseline": cpu, "clocks": clocks, "inference": inference,
    }
    emit(line)
    if world > 1:
        dist.destroy_process_group()


_REAL_STDOUT = None


def _quiet_stdout():
    """Libraries (NCCL's version banner, …) write to fd 1; the contract is ONE JSON line on stdout, so everything else is
    sent to stderr and the JSON line goes to the saved descriptor."""
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)


def emit(line: dict) -> None:
    out = _REAL_STDOUT if _REAL_STDOUT is not None else sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    _quiet_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-inference", action="store_true", help="skip the secondary inference (configs[0]) figure")
    ap.add_argument("--gemm-breakdown", default=None, help="write a per-shape GEMM timing table to this file")
    ap.add_argument("--eager", action="store_true", help="no CUDA graph (profiling runs: one kernel launch per API call)")
    args = ap.parse_args()
    if args.impl == "reference":
        if args.steps == 20 and args.warmup == 5:
            args.steps, args.warmup = 3, 1
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
