"""Multi-GPU parity (run under torchrun, one process per GPU, NCCL): the all-reduced adapter-gradient bucket of N ranks
(each on its shard of the batch) equals the bucket of one rank on the whole batch; after the fused AdamW step every rank
holds identical parameters.  Prints DDP_CHECK PASS / FAIL on rank 0."""
import importlib
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import synth_wave  # noqa: E402

P = importlib.import_module("jiao-liao_speech_recognition_b200")


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    cfg = P.JLConfig(hidden_size=256, num_hidden_layers=3, num_attention_heads=4, intermediate_size=512, conv_channels=128, vocab_size=96,
                     adapter_attn="att", adapter_ffn="wf", wf_bottleneck=64, wf_rank=16)

    def make():
        torch.manual_seed(0)
        m = P.JLForCTC(cfg).cuda()
        m.freeze_base_model()
        return m

    n = 32000
    per = 2
    total = per * world
    wave = torch.stack([synth_wave(n, 100 + i) for i in range(total)])
    ns = torch.full((total,), n, dtype=torch.int32)
    g = torch.Generator().manual_seed(7)
    labels = torch.randint(1, cfg.vocab_size, (total, 12), generator=g, dtype=torch.int32)
    # N-rank: each rank steps on its shard, one all-reduce of the bucket, fused AdamW with grad_scale 1/world
    tr = P.AdapterTrainer(make(), lr=1e-3, use_cuda_graph=True)
    sl = slice(rank * per, (rank + 1) * per)
    loss = tr.step(wave[sl].pin_memory(), ns[sl], labels[sl]).item()
    torch.cuda.synchronize()
    bucket = tr.flat.grad.clone()
    params = tr.flat.param.clone()
    lt = torch.tensor([loss], device="cuda", dtype=torch.float64)
    dist.all_reduce(lt)
    # reference: one process, whole batch, no collective (world "1": run with the process group hidden)
    ok = True
    if rank == 0:
        ref = P.AdapterTrainer(make(), lr=1e-3, use_cuda_graph=False, comm=None)      # mode "none": no collective
        ref.grad_scale = 1.0 / world                # the whole batch on one rank: sum-reduced CTC gradients, DDP's 1 / world mean
        rloss = ref.step(wave.pin_memory(), ns, labels).item()
        torch.cuda.synchronize()
        gerr = float((bucket - ref.flat.grad).norm() / ref.flat.grad.norm())
        perr = float((params - ref.flat.param).abs().max())
        lerr = abs(float(lt) - rloss) / abs(rloss)
        print(f"collective: {tr.flat.comm_mode} ({'jl_comm_allreduce (C ABI), inside the step graph, two halves' if tr.flat.comm_mode == 'jl' else 'torch.distributed'})")
        print(f"world {world}: loss sum {float(lt):.4f} vs single {rloss:.4f} (rel {lerr:.2e}); bucket rel err {gerr:.2e}; param max diff {perr:.2e}")
        ok = lerr < 1e-3 and gerr < 2e-2 and perr < 1e-3
    # every rank must hold the same parameters after the step
    pmax, pmin = params.clone(), params.clone()
    dist.all_reduce(pmax, op=dist.ReduceOp.MAX)
    dist.all_reduce(pmin, op=dist.ReduceOp.MIN)
    same = bool(torch.equal(pmax, pmin))
    if rank == 0:
        print("DDP_CHECK", "PASS" if (ok and same) else "FAIL", f"(replicas identical: {same})")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
