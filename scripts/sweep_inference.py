"""BASELINE.json configs[4]: streaming-free inference sweep, batch 1–512 × 10 s per GPU: pinned-host waveforms → mel + encoder
(WFAdapter) + CTC greedy decode → token ids on the host.  One markdown row per batch size (device-resident and end-to-end
audio-s/s, whole job).  Under torchrun (WORLD_SIZE > 1) every rank runs its own replica on its own batch — inference has no
collective — the ranks start each measurement together and the time is the maximum over the ranks.

    python scripts/sweep_inference.py
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 scripts/sweep_inference.py
"""
import importlib, os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import synth_batch  # noqa: E402
rank, world, local = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist = None
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
P = importlib.import_module("jiao-liao_speech_recognition_b200")
cfg = P.JLConfig.base(adapter_ffn="wf")
model = P.JLForCTC(cfg).cuda().eval()


def barrier():
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()


def max_over_ranks(t: float) -> float:
    if dist is None:
        return t
    tt = torch.tensor([t], device="cuda", dtype=torch.float64)
    dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    return float(tt[0])


if rank == 0:
    print(f"{world} x B200, batch per GPU x 10 s\n")
    print("| batch per GPU | ms/step (resident) | audio-s/s (resident, whole job) | audio-s/s (e2e: H2D waveforms + D2H ids, whole job) | RTF | launches |\n|---:|---:|---:|---:|---:|---:|")
for b in (1, 2, 4, 8, 16, 32, 64, 128, 256, 512):
    tr = P.Transcriber(model, use_cuda_graph=True)
    wave, ns, _, _ = synth_batch(b, 1234 + rank, cfg.vocab_size)
    wave = wave.pin_memory()
    for _ in range(3):
        ids, n = tr(wave, ns)
        ids.cpu()
    steps = 20 if b <= 128 else 8
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(steps):
        tr.run_resident()
    e1.record()
    barrier()
    t_res = max_over_ranks(e0.elapsed_time(e1) / 1e3 / steps)
    barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        ids, n = tr(wave, ns)
        host = ids.cpu(); n.cpu()
    torch.cuda.synchronize()
    t_e2e = max_over_ranks((time.perf_counter() - t0) / steps)
    if rank == 0:
        print(f"| {b} | {t_res * 1e3:.2f} | {world * b * 10 / t_res:,.0f} | {world * b * 10 / t_e2e:,.0f} | {t_res / (world * b * 10):.2e} | {tr.launches_per_step} |", flush=True)
    del tr
    torch.cuda.empty_cache()
if dist is not None:
    dist.destroy_process_group()
