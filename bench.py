#!/usr/bin/env python
"""Benchmark of the adapter fine-tune step in audio-seconds/sec.

    python bench.py --gpus N --steps K --warmup W                   # BASELINE.json configs[1], the headline (default)
    python bench.py --config large|mixed ...                        # configs[2] / configs[3] (see WORKLOADS)
    python bench.py --impl reference --gpus N --steps K --warmup W  # the reference's CPU implementation of the same step (oracle port)

One JSON line on stdout (rank 0).  `value` = whole-job audio-s/s with the waveforms already resident in HBM (one CUDA-graph replay
per step: forward, CTC loss, adapter-only backward, gradient all-reduce in two halves, fused AdamW); `e2e` = the same step through
the public API with pinned-host waveforms / labels copied in and the loss read back every step; `roofline` = the step's tcgen05
GEMM launches re-issued on their own and timed with CUDA events against the measured bf16 peak (burst AND sustained, with the
clock record that says which applies), plus live HBM rooflines of the mel, LayerNorm, adapter and CTC kernels; `cpu_baseline` = the
oracle's step on the SAME full batch, timed on this box's host cores.
"""
from __future__ import annotations

import argparse
import hashlib
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
BATCH = 32            # utterances per GPU (weak scaling)
UNIT = "audio-s/s"
METRIC = "audio-seconds/sec (adapter fine-tune step: fwd + CTC loss + adapter-only bwd + allreduce + AdamW)"

# name -> (BASELINE.json config index, model kwargs, description)
WORKLOADS = {
    "base": dict(index=1, size="base", model=dict(adapter_ffn="att"),
                 what="12-layer d=768 h=12 ffn=3072 encoder (80-mel, 2x conv k5 s2 + GLU) with AttAdapter(b=64) after each FFN, V=5000 CTC head; "
                      "fine-tune step (CTC loss, adapter-only backward, frozen backbone), batch 32 x 10 s synthetic 16 kHz audio per GPU"),
    "large": dict(index=2, size="large", model=dict(adapter_attn="att", adapter_ffn="wf"),
                  what="24-layer d=1024 h=16 ffn=4096 XLS-R/wav2vec2-large-style stack on the 80-mel front end with AttAdapter(b=64) after the "
                       "attention and WFAdapter(b=256, r=32) after the FFN of every layer, V=5000; fine-tune step, batch 32 x 10 s per GPU"),
    "mixed": dict(index=3, size="base", model=dict(adapter_attn="att", adapter_ffn="wf", num_dialects=4),
                  what="multi-dialect batch: 32 utterances per GPU of U[2, 30] s (seeded), 4 dialects (per-utterance WFAdapter factor sets, "
                       "sorted by dialect) + AttAdapter on the 12-layer d=768 encoder; utterances sharded frame-balanced across ranks; "
                       "fine-tune step in the packed (cu_seqlens) row layout — no GEMM / LayerNorm / attention / CTC work on padding"),
}


def _wave(batch: int, n: int, g: torch.Generator) -> torch.Tensor:
    """SURVEY §8d synthetic audio: 0.1·randn + 0.05·Σ_5 sin(2π f_k t + φ_k), f_k ~ U[100, 4000] Hz, clipped to [-1, 1]."""
    t = torch.arange(n, dtype=torch.float32) / SR
    wave = 0.1 * torch.randn(batch, n, generator=g)
    for _ in range(5):
        f = 100.0 + 3900.0 * torch.rand(batch, 1, generator=g)
        ph = 6.283185307179586 * torch.rand(batch, 1, generator=g)
        wave += 0.05 * torch.sin(6.283185307179586 * f * t[None, :] + ph)
    return wave.clamp_(-1.0, 1.0)


def _tprime(n: int) -> int:
    frames = 0 if n < 400 else 1 + (n - 400) // 160
    return ((frames - 1) // 2 + 1 - 1) // 2 + 1 if frames > 0 else 0


def synth_batch(batch: int, seed: int, vocab: int):
    """Fixed-length batch: `batch` x 10 s; labels S = ⌊0.4·T'⌋ ~ U{1..V-1}.  → (wave, num_samples, labels int32, T')."""
    g = torch.Generator().manual_seed(seed)
    n = SR * SECONDS
    wave = _wave(batch, n, g)
    tp = _tprime(n)
    s = int(0.4 * tp)
    labels = torch.randint(1, vocab, (batch, s), generator=g, dtype=torch.int32)
    ns = torch.full((batch,), n, dtype=torch.int32)
    return wave, ns, labels, tp


def synth_mixed_global(world: int, per_gpu: int, seed: int, vocab: int, num_dialects: int):
    """configs[3]: a GLOBAL batch of per_gpu·world utterances with durations ~ U[2, 30] s and a dialect id each, the same on every
    rank (seeded).  Returns (num_samples list, dialect list, per-utterance seeds)."""
    g = torch.Generator().manual_seed(seed)
    total = per_gpu * world
    dur = 2.0 + 28.0 * torch.rand(total, generator=g)
    ns = [int(SR * float(d)) for d in dur]
    dialects = torch.randint(0, num_dialects, (total,), generator=g).tolist()
    return ns, dialects


def mixed_rank_batch(P, rank: int, world: int, per_gpu: int, seed: int, vocab: int, num_dialects: int):
    """This rank's shard of the global mixed-length batch (frame-balanced, `shard_utterances`), sorted by dialect (utterances of a
    dialect adjacent: one row slice per factor set), padded to the shard's longest utterance."""
    ns_all, dia_all = synth_mixed_global(world, per_gpu, seed, vocab, num_dialects)
    frames = [P.feature_extraction.num_frames(n) for n in ns_all]
    shards = P.shard_utterances(frames, world)
    mine = sorted(shards[rank], key=lambda i: (dia_all[i], i))
    ns = [ns_all[i] for i in mine]
    nmax = (max(ns) + 3) // 4 * 4
    wave = torch.zeros((len(mine), nmax), dtype=torch.float32)
    tps = []
    for r, i in enumerate(mine):
        g = torch.Generator().manual_seed(seed * 1000003 + i)
        wave[r, : ns[r]] = _wave(1, ns[r], g)[0]
        tps.append(_tprime(ns[r]))
    smax = max(1, max(int(0.4 * t) for t in tps))
    labels = torch.full((len(mine), smax), -100, dtype=torch.int32)
    for r, i in enumerate(mine):
        g = torch.Generator().manual_seed(seed * 7919 + i)
        s = int(0.4 * tps[r])
        labels[r, :s] = torch.randint(1, vocab, (s,), generator=g, dtype=torch.int32)
    loads = [sum(frames[i] for i in s) for s in shards]
    info = {"audio_seconds_global": sum(ns_all) / SR, "audio_seconds_rank": sum(ns) / SR, "utterances_rank": len(mine),
            "frames_per_rank": loads, "load_imbalance": (max(loads) / (sum(loads) / len(loads))) if loads else 1.0,
            "padded_seconds_rank": len(mine) * nmax / SR, "tokens_rank": sum(tps), "padded_tokens_rank": len(mine) * max(tps)}
    return wave, torch.tensor(ns, dtype=torch.int32), labels, [dia_all[i] for i in mine], info


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return {"hbm_gbs": p["hbm_gbs"], "bf16_tflops": p["bf16_tflops"], "bf16_tflops_sustained": p["bf16_tflops_sustained"], "src": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "src": "fallback"}


def gemm_source_sha16() -> str:
    """Hash of the sources the GEMM kernels are built from (scripts/ncu_raw_summary.py writes the same hash next to the ncu DRAM
    traffic it extracts, so a committed figure is only reported for the build it was measured on)."""
    h = hashlib.sha256()
    for name in ("gemm_tcgen05.cu", "ptx_sm100.cuh", "common.cuh", "common.cu"):
        with open(os.path.join(ROOT, PKG, "csrc", name), "rb") as f:
            h.update(f.read())
    return h.hexdigest()[:16]


def source_sha16() -> str:
    """Hash of the CUDA sources the library is built from: ties committed ncu figures (profiles/*_traffic.json) to a build."""
    h = hashlib.sha256()
    csrc = os.path.join(ROOT, PKG, "csrc")
    for name in sorted(os.listdir(csrc)):
        if name.endswith((".cu", ".cuh")):
            with open(os.path.join(csrc, name), "rb") as f:
                h.update(f.read())
    return h.hexdigest()[:16]


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
        sm, mx, reasons, power = [], None, set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in out.strip().splitlines():
            parts = [x.strip() for x in ln.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx = float(parts[1])
                power.append(float(parts[2]))
            except ValueError:
                continue
            for nm, v in zip(names, parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm),
                "power_w_max": max(power) if power else None}


# ------------------------------------------------------------------------------------------------------------------ CPU arm
def _oracle_problem(workload: str, rank: int = 0, world: int = 1):
    """The oracle's copy of the workload: (step function running fwd + CTC loss + adapter-only backward, audio seconds per step,
    description).  fp32 torch CPU ops, autograd for the adapter / lm_head gradients (frozen backbone)."""
    from oracle import model as om
    wl = WORKLOADS[workload]
    size = dict(hidden_size=1024, num_hidden_layers=24, num_attention_heads=16, intermediate_size=4096) if wl["size"] == "large" else {}
    cfg = om.OracleConfig(**size, **wl["model"])
    w = om.init_weights(cfg, seed=0)
    for k, v in w.items():
        v.requires_grad_(om.is_trainable(k))
    if workload == "mixed":
        P = importlib.import_module(PKG)
        wave, ns, labels, dialects, info = mixed_rank_batch(P, rank, world, BATCH, 1234, cfg.vocab_size, cfg.num_dialects)
        waves = [wave[i, : int(ns[i])] for i in range(wave.shape[0])]
        audio = info["audio_seconds_rank"]
        desc = f"{len(waves)} utterances of U[2, 30] s ({audio:.0f} audio-s), rank 0's shard"
    else:
        wave, ns, labels, tp = synth_batch(BATCH, 1234, cfg.vocab_size)
        waves = [wave[i] for i in range(BATCH)]
        dialects = 0
        audio = BATCH * SECONDS
        desc = f"the full {BATCH} x {SECONDS} s batch"
    lab = labels.to(torch.int64)

    def step():
        loss, _, _ = om.forward_from_waveforms(w, cfg, waves, lab, dialect=dialects)
        loss.backward()
        for v in w.values():
            v.grad = None

    return step, audio, desc


def cpu_step_rate(workload: str, repeats: int, warmup: int, threads: int):
    """The oracle's fine-tune step on the full per-GPU batch → (audio-s/s, seconds/step, description)."""
    torch.set_num_threads(threads)
    step, audio, desc = _oracle_problem(workload)
    best = None
    for it in range(warmup + repeats):
        t0 = time.perf_counter()
        step()
        dt = time.perf_counter() - t0
        if it >= warmup:
            best = dt if best is None else min(best, dt)
    return audio / best, best, desc


def run_reference(args):
    """The reference arm: the oracle port of the reference's step (the reference publishes no code — /root/reference/README.md:3 —
    so there is nothing to install; `cpu_baseline.kind` = "port") on ALL host cores, on the same workload, for exactly the
    --steps / --warmup the driver asks for.  Rank 0 only; a step costs a few seconds."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    step, audio, desc = _oracle_problem(args.config)
    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    value = audio * args.steps / dt
    sample = f"each step = the oracle's fine-tune step (fp32, torch CPU ops, {threads} threads) on {desc}: the same work per step as one GPU of the product arm"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": workload_config(args.config, args.gpus),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


def workload_config(name: str, n_gpus: int, trainable: int = 0, extra: dict = None):
    wl = WORKLOADS[name]
    cfg = {"workload": f"BASELINE.json configs[{wl['index']}]: {wl['what']}",
           "per_gpu_batch": BATCH, "global_batch": BATCH * n_gpus, "parallelism": f"dp{n_gpus}",
           "collective": "one NCCL all-reduce of the adapter + lm_head fp32 gradient bucket per step, issued in two halves inside the step's "
                         "CUDA graph (the first under the backward pass of the lower layers)"
                         + (f"; {trainable / 1e6:.2f} M trainable parameters" if trainable else ""),
           "l2_policy": "inputs larger than L2: each step streams > 2 GB of activations (126 MB L2), no explicit flush"}
    if name != "mixed":
        cfg.update({"seconds_per_utterance": SECONDS, "labels_per_utterance": int(0.4 * _tprime(SR * SECONDS))})
    if extra:
        cfg.update(extra)
    return cfg


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


def _time_rotating(fn_list, reps: int = 3) -> float:
    """Average seconds per call of the launches in `fn_list` (each on its own buffers: together > L2, so every launch streams its
    operands from HBM).  The sequence is captured into a CUDA graph — as the product runs its kernels — so the figure is device
    time, not Python call overhead; CUDA events on the replaying stream, after a warm-up replay."""
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for fn in fn_list:
            fn()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        keep = [fn() for fn in fn_list]
    graph.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        graph.replay()
    e1.record()
    torch.cuda.synchronize()
    del keep
    return e0.elapsed_time(e1) / 1e3 / (reps * len(fn_list))


def hbm_kernel_rooflines(P, peaks, d: int, m: int, vocab: int, batch: int):
    """Live HBM rooflines of the memory-bound kernels on the path (SURVEY §8d algorithmic bytes), each launched on its own on
    bench-shaped operands, rotating over enough buffer sets that the working set exceeds the 126 MB L2."""
    ops, L = P.ops, P._lib
    dev = torch.device("cuda")
    hbm = peaks["hbm_gbs"]
    out = {}

    def entry(name, nbytes, secs, note, flops=None):
        e = {"algorithmic_bytes": nbytes, "us": 1e6 * secs, "achieved_gbs": nbytes / secs / 1e9, "frac_of_hbm_peak": nbytes / secs / 1e9 / hbm, "note": note}
        if flops is not None:
            e["achieved_fp32_tflops"] = flops / secs / 1e12
        out[name] = e

    # ---- a1 + a2: mel + CMVN, batch x 10 s (waveform read once, fp32 + bf16 features written once)
    fe = P.JLFeatureExtractor(device=dev)
    nsets = 8
    waves = [torch.rand((batch, SR * SECONDS), device=dev) * 0.2 - 0.1 for _ in range(nsets)]
    nsamp = torch.full((batch,), SR * SECONDS, dtype=torch.int32, device=dev)
    frames = 998
    secs = _time_rotating([lambda w=w: fe.extract_device(w, nsamp, frames, return_bf16=True) for w in waves])
    nbytes = batch * (SR * SECONDS * 4 + frames * 80 * 4 + frames * 80 * 2)
    entry("mel_fbank+cmvn", nbytes, secs, f"{batch} x 10 s: 640 000 B read + 998 x 80 x (4 + 2) B written per utterance; fp32 issue rate co-binds "
          "(~13 kFLOP per 961 B frame, SURVEY §8d)", flops=batch * frames * 13.0e3)
    del waves
    # ---- LayerNorm forward / backward on the residual stream [m, d]
    nsets = max(2, int(200e6 // (m * d * 4)) + 1)
    xs = [torch.randn((m, d), device=dev).to(torch.bfloat16) for _ in range(nsets)]
    gamma, beta = torch.ones(d, device=dev), torch.zeros(d, device=dev)
    secs = _time_rotating([lambda x=x: ops.layernorm_fwd(x, gamma, beta, 1e-5, save_stats=True) for x in xs])
    entry("layernorm_fwd", 2 * m * d * 2, secs, f"[{m}, {d}] bf16 read + written once")
    stats = [ops.layernorm_fwd(x, gamma, beta, 1e-5, save_stats=True) for x in xs]
    secs = _time_rotating([lambda x=x, s=s: ops.layernorm_bwd(s[0], x, gamma, s[1], s[2], dres=x) for x, s in zip(xs, stats)])
    entry("layernorm_bwd", 4 * m * d * 2, secs, f"dy, x, residual grad read + dx written, [{m}, {d}] bf16")
    del stats
    # ---- a6: fused WFAdapter forward (read h, write h')
    cfg = P.JLConfig(hidden_size=d, num_attention_heads=d // 64, num_hidden_layers=1, adapter_ffn="wf")
    ad = P.WFAdapter(d, cfg.wf_bottleneck, cfg.wf_rank).cuda()
    eng = P.JLEncoder(cfg).cuda().engine()
    pack = eng._wf_pack(ad, 0)
    secs = _time_rotating([lambda x=x: ops.wfadapter_fwd(x, pack, 1e-5) for x in xs])
    entry("wfadapter_fwd", 2 * m * d * 2, secs, f"one kernel: LN + 4 low-rank GEMMs + residual on [{m}, {d}] bf16 (read h, write h')")
    # ---- a7: AttAdapter forward chain as the engine issues it (LN → q|k|v GEMM → 1-head attention → out-proj + residual)
    att = P.AttAdapter(d).cuda()
    lengths = torch.full((batch,), m // batch, dtype=torch.int32, device=dev)

    def att_fwd(x):
        out_, _ = eng._adapter_fwd(att, x, lengths, batch, m // batch, False, 0, zero_rows=False)
        return out_
    L.launch_count_reset()
    att_fwd(xs[0])
    n_att = L.launch_count()
    secs = _time_rotating([lambda x=x: att_fwd(x) for x in xs])
    entry("attadapter_fwd", 2 * m * d * 2, secs, f"{n_att} launches on [{m}, {d}] bf16 (read h, write h'): the algorithmic bytes of the adapter vs the time of its chain")
    del xs
    # ---- a9: CTC loss + gradient (logits fp32 read by the row-statistics pass and again by the gradient pass, bf16 gradient written)
    t = m // batch
    nsets = 2
    lg = [torch.randn((batch, t, vocab), device=dev) for _ in range(nsets)]
    lab = torch.randint(1, vocab, (batch, int(0.4 * t)), device=dev, dtype=torch.int32)
    secs = _time_rotating([lambda x=x: ops.ctc_loss(x, lab, lengths, want_grad=True) for x in lg])
    entry("ctc_loss+grad", batch * t * vocab * (4 + 4 + 2), secs, f"logits [{batch}, {t}, {vocab}] fp32 read twice (row statistics, gradient) + bf16 gradient written; "
          "includes the serial lattice kernel")
    secs = _time_rotating([lambda x=x: ops.ctc_greedy(x, lengths) for x in lg])
    entry("ctc_greedy", batch * t * vocab * 4, secs, "argmax over V (logits read once) + collapse")
    return out


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
    if os.environ.get("JL_LNPROJ_SPLIT"):          # tuning aid: CTAs per row tile of jl_lnproj_bwd
        P._lib.load().jl_debug_set_lnproj_split(int(os.environ["JL_LNPROJ_SPLIT"]))
    if os.environ.get("JL_GEMM_TAIL"):             # tuning aid: tail-wave policy of the GEMM (jl_debug_set_gemm_tail bit mask)
        P._lib.load().jl_debug_set_gemm_tail(int(os.environ["JL_GEMM_TAIL"]))
    wl = WORKLOADS[args.config]
    cfg = (P.JLConfig.large if wl["size"] == "large" else P.JLConfig.base)(**wl["model"])
    model = P.JLForCTC(cfg).cuda()
    model.freeze_base_model()
    packed = (args.config == "mixed") and not args.padded
    trainer = P.AdapterTrainer(model, lr=1e-4, use_cuda_graph=not args.eager, packed=packed,
                               exchange_in_graph=os.environ.get("JL_EXCHANGE_IN_GRAPH", "1") != "0")
    extra = {}
    if args.config == "mixed":
        wave, ns, labels, dialect, info = mixed_rank_batch(P, rank, world, BATCH, 1234, cfg.vocab_size, cfg.num_dialects)
        audio_s_per_step = info["audio_seconds_global"]
        extra = {"row_layout": "packed (cu_seqlens)" if packed else "padded to the shard's longest utterance", "load_imbalance_max_over_mean": info["load_imbalance"],
                 "rank0": {k: info[k] for k in ("utterances_rank", "audio_seconds_rank", "padded_seconds_rank", "tokens_rank", "padded_tokens_rank")},
                 "audio_seconds_per_step_global": audio_s_per_step}
    else:
        wave, ns, labels, tp = synth_batch(BATCH, 1234 + rank, cfg.vocab_size)
        dialect = 0
        audio_s_per_step = BATCH * SECONDS * world
    wave_p, labels_p, ns_p = wave.pin_memory(), labels.pin_memory(), ns.pin_memory()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up through the public path (captures the CUDA graph on the first call)
    for _ in range(max(args.warmup, 3)):
        loss = trainer.step(wave_p, ns, labels_p, dialect=dialect)
        loss_val = float(loss.item())
    # ... and through the prefetching path the e2e loop uses (its one-time costs — staging buffers, the copy stream, the pinned loss
    # slots — belong to warm-up, not to a timed step)
    for _ in range(2):
        trainer.submit(wave_p, ns_p, labels_p, dialect=dialect)
        loss_val = trainer.step_async().item()
    launches_per_step = trainer.launches_per_step          # our kernels inside the graph, AdamW included (the all-reduce is NCCL's kernel)

    sampler = ClockSampler(local) if rank == 0 else None
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    # ---- device-resident throughput: one graph replay per step, inputs already in HBM
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
    # The loss of step i is read (pinned host memory, its own event) after step i+1 has been launched — what a trainer that logs
    # the loss does — so one step is always queued on the device and the host work between two steps is not exposed; the last
    # loss is read before the closing event.  `sync_value` below is the same loop with loss.item() right after every step.
    barrier()
    ev[2].record()
    trainer.submit(wave_p, ns_p, labels_p, dialect=dialect)
    pending = None
    for i in range(args.steps):
        h = trainer.step_async()
        if i + 1 < args.steps:
            trainer.submit(wave_p, ns_p, labels_p, dialect=dialect)
        if pending is not None:
            loss_val = pending.item()
        pending = h
    loss_val = pending.item()
    ev[3].record()
    barrier()
    t_e2e = ev[2].elapsed_time(ev[3]) / 1e3
    barrier()
    ev[2].record()
    trainer.submit(wave_p, ns_p, labels_p, dialect=dialect)
    for i in range(args.steps):
        loss = trainer.step()
        if i + 1 < args.steps:
            trainer.submit(wave_p, ns_p, labels_p, dialect=dialect)
        loss_val = float(loss.item())
    ev[3].record()
    barrier()
    t_e2e_sync = ev[2].elapsed_time(ev[3]) / 1e3
    # the same without overlap (copy, then compute, on one stream) for comparison
    barrier()
    ev[2].record()
    for _ in range(args.steps):
        loss = trainer.step(wave_p, ns_p, labels_p, dialect=dialect)
        loss_val = float(loss.item())
    ev[3].record()
    barrier()
    t_e2e_serial = ev[2].elapsed_time(ev[3]) / 1e3
    clocks = sampler.stop() if sampler else None
    if world > 1:
        tt = torch.tensor([t_res, t_e2e, t_e2e_serial, t_e2e_sync], device="cuda", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        t_res, t_e2e, t_e2e_serial, t_e2e_sync = float(tt[0]), float(tt[1]), float(tt[2]), float(tt[3])

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
    # which peak applies (B200_PROFILING.md): the burst figure for a kernel timed alone at full clock, the sustained one under the
    # power cap.  Decided from the clock record of THIS run; both fractions are reported.
    capped = bool(clocks and ("sw_power_cap" in clocks.get("reasons", []) or (clocks.get("sm_mhz") and clocks.get("sm_max_mhz") and
                                                                               clocks["sm_mhz"] < 0.95 * clocks["sm_max_mhz"])))
    peak_kind = "bf16_tflops_sustained" if capped else "bf16_tflops"
    traffic, traffic_note = None, "no ncu capture of this build is committed (profiles/*_gemm_traffic.json carries the source hash of the build it was taken from)"
    sha = source_sha16()
    import glob
    for tpath in sorted(glob.glob(os.path.join(ROOT, "profiles", "*_gemm_traffic.json")), reverse=True):
        with open(tpath) as f:
            tj = json.load(f)
        if tj.get("src_sha16") == sha or tj.get("gemm_src_sha16") == gemm_source_sha16():
            traffic = tj["dram_bytes"]
            traffic_note = (f"ncu dram read+write of one launch of the shape with the largest share of the step {tj['shape_mnk']}: "
                            f"{tj['dram_bytes'] / 1e6:.1f} MB vs {tj['algorithmic_bytes'] / 1e6:.1f} MB algorithmic ({tj['source']}; the GEMM sources of this build hash to the same value)")
            break
    roofline = {"bound": "tensor", "kernel": "gemm_tcgen05_kernel<256> (128 x 256 tiles) / gemm_tcgen05_2cta_kernel (256-row CTA-pair tiles)",
                "achieved": achieved, "peak": peaks[peak_kind], "unit": "TFLOP/s", "frac": achieved / peaks[peak_kind], "traffic": traffic, "traffic_note": traffic_note,
                "peak_source": f"{peaks['src']} {peak_kind}: clock record of this run — median SM clock {clocks.get('sm_mhz') if clocks else None} MHz of "
                               f"{clocks.get('sm_max_mhz') if clocks else None}, reasons {clocks.get('reasons') if clocks else None}; the {len(trace)} GEMM launches of one "
                               f"step are timed back to back ({1e3 * t_gemm:.2f} ms)",
                "frac_of_burst_peak": achieved / peaks["bf16_tflops"], "frac_of_sustained_peak": achieved / peaks["bf16_tflops_sustained"],
                "launches_per_step": len(trace), "algorithmic_tflop_per_step": flops / 1e12, "gemm_ms_per_step": 1e3 * t_gemm,
                "share_of_step": t_gemm / (t_res / args.steps), "src_sha16": sha}

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
        if big_us > 0:
            tf = big_fl / big_us / 1e6
            roofline["encoder_gemms"] = {"achieved": tf, "frac": tf / peaks[peak_kind], "frac_of_burst_peak": tf / peaks["bf16_tflops"],
                                         "frac_of_sustained_peak": tf / peaks["bf16_tflops_sustained"],
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
    # ---- HBM rooflines of the memory-bound kernels (mel, LayerNorm, adapters, CTC), measured live
    hbm_kernels = None
    if world == 1 and not args.no_kernel_rooflines:
        tok = BATCH * _tprime(SR * SECONDS)
        hbm_kernels = hbm_kernel_rooflines(P, peaks, cfg.hidden_size, tok, cfg.vocab_size, BATCH)
        roofline["hbm_kernels"] = hbm_kernels
        roofline["hbm_peak_gbs"] = peaks["hbm_gbs"]
    # ---- secondary figure (not the headline): inference forward waveform → token ids on BASELINE.json configs[0]
    inference = None
    if world == 1 and not args.no_inference:
        inference = inference_rate(P, steps=max(args.steps, 10))
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        v, sec, desc = cpu_step_rate(args.config, repeats=3, warmup=1, threads=threads)
        cpu = {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
               "sample": f"oracle fine-tune step on {desc}, best of 3 after 1 warm-up ({sec:.2f} s/step)"}
    h2d = wave_p.numel() * 4 + labels_p.numel() * 4 + ns.numel() * 4 * (3 if packed else 2)
    line = {
        "metric": METRIC, "value": audio_s_per_step * args.steps / t_res, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": 1e3 * t_res / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
        "data": "synthetic", "config": workload_config(args.config, world, trainer.flat.num_params, extra),
        "e2e": {"value": audio_s_per_step * args.steps / t_e2e, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                "ms_per_step": 1e3 * t_e2e / args.steps,
                "api": "AdapterTrainer.submit(next batch) + step_async() + LossHandle.item() of the previous step (every step's loss is read; "
                       "the read of step i overlaps step i+1)",
                "sync_value": audio_s_per_step * args.steps / t_e2e_sync,
                "sync_note": "the same loop with loss.item() immediately after every step (host and device in lockstep)",
                "serial_value": audio_s_per_step * args.steps / t_e2e_serial,
                "serial_note": "AdapterTrainer.step(batch) with the copy and the kernels on one stream (no prefetch)"},
        "gpu_launches": launches_per_step * args.steps * 3, "gpu_launches_per_step": launches_per_step,
        "rtf": (t_res / args.steps) / audio_s_per_step, "loss": loss_val, "cuda_graph": not args.eager,
        "exchange": {"in_graph": trainer.exchange_in_graph, "overlapped_halves": trainer.overlap_exchange, "mode": trainer.flat.comm_mode,
                     "bucket_bytes": trainer.flat.total * 4,
                     "first_part_bytes": (trainer.flat.split_head if trainer.eng.defer_wgrads else trainer.flat.split) * 4,
                     "adapter_wgrads": "deferred to the end of the backward pass (unsplit, concurrent)" if trainer.eng.defer_wgrads else "beside the main chain"},
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
    ap.add_argument("--config", default="base", choices=sorted(WORKLOADS), help="base = BASELINE.json configs[1] (headline), large = configs[2], mixed = configs[3]")
    ap.add_argument("--padded", action="store_true", help="--config mixed: pad to the longest utterance instead of the packed row layout (A/B)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-inference", action="store_true", help="skip the secondary inference (configs[0]) figure")
    ap.add_argument("--no-kernel-rooflines", action="store_true", help="skip the live HBM rooflines of the mel / LayerNorm / adapter / CTC kernels")
    ap.add_argument("--gemm-breakdown", default=None, help="write a per-shape GEMM timing table to this file")
    ap.add_argument("--eager", action="store_true", help="no CUDA graph (profiling runs: one kernel launch per API call)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
