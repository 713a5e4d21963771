"""Adapter fine-tuning plumbing around the kernels: one flat fp32 bucket for the trainable parameters (adapters +
lm_head), their gradients and AdamW state; one NCCL all-reduce of that bucket per step (the only collective on the
path — the backbone is frozen); a fused AdamW kernel that also refreshes the bf16 shadow the GEMMs read;
and CUDA-graph capture of the whole step (waveform → mel → encoder → CTC → adapter-only backward) so that the
≈ 500 kernel launches of a step cost one graph launch.

Replaces, for this path, DDP's bucketed reducer + ``torch.optim.AdamW`` as a SpeechBrain/HF recipe would use them
(/root/reference/requirements.txt:1,71,75).
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import torch
import torch.distributed as dist

from . import _lib as L
from . import ops
from .comm import JLComm
from .feature_extraction import JLFeatureExtractor, device_tables, num_frames
from .modeling import AttAdapter, GradSink, JLForCTC, subsampled_length, wav2vec2_lengths

BF16, F32, I32 = torch.bfloat16, torch.float32, torch.int32
_ALIGN = 64   # elements; keeps every view 16-byte aligned in both the fp32 and the bf16 buffer


def ordered_trainables(model: JLForCTC) -> List[torch.nn.Parameter]:
    """Trainable parameters, with each AttAdapter's q/k/v weights (and biases) adjacent so that the concatenated
    [192, d] projection the kernels use is a plain view of the bucket."""
    seen, out = set(), []

    def add(p):
        if id(p) not in seen and p.requires_grad:
            seen.add(id(p))
            out.append(p)

    for m in model.modules():
        if isinstance(m, AttAdapter):
            for p in (m.q_proj.weight, m.k_proj.weight, m.v_proj.weight, m.q_proj.bias, m.k_proj.bias, m.v_proj.bias):
                add(p)
    for p in model._get_adapters().values():
        add(p)
    return out


class BucketLayout:
    """Offsets of the trainable parameters inside one flat buffer (each padded to 64 elements so that every view is
    16-byte aligned in fp32 and in bf16).  Device-agnostic: the same layout addresses the fp32 master copy, the gradient
    bucket that is all-reduced, the AdamW moments and the bf16 shadow."""

    def __init__(self, params: Sequence[torch.Tensor]):
        self.offset: Dict[int, int] = {}
        off = 0
        for p in params:
            self.offset[id(p)] = off
            off += (p.numel() + _ALIGN - 1) // _ALIGN * _ALIGN
        self.total = off
        self.num_params = sum(p.numel() for p in params)

    def view(self, buf: torch.Tensor, p: torch.Tensor) -> torch.Tensor:
        o = self.offset[id(p)]
        return buf[o: o + p.numel()].view(p.shape)

    def adjacent(self, ps: Sequence[torch.Tensor]) -> bool:
        for a, b in zip(ps[:-1], ps[1:]):
            if id(a) not in self.offset or id(b) not in self.offset or a.numel() % _ALIGN:
                return False
            if self.offset[id(b)] != self.offset[id(a)] + a.numel() or a.shape[1:] != b.shape[1:]:
                return False
        return id(ps[-1]) in self.offset

    def cat(self, buf: torch.Tensor, ps: Sequence[torch.Tensor]) -> Optional[torch.Tensor]:
        """[Σ rows, ...] view over consecutive parameters (e.g. q/k/v projections), or None if they are not adjacent."""
        if not self.adjacent(ps):
            return None
        o = self.offset[id(ps[0])]
        n = sum(p.numel() for p in ps)
        rows = sum(p.shape[0] for p in ps)
        return buf[o: o + n].view((rows,) + tuple(ps[0].shape[1:]))


class FlatAdapterParams(GradSink):
    """Flat storage for the trainable set.  After construction every trainable ``nn.Parameter``'s ``.data`` is a view
    of ``self.param``; gradients are written by the backward kernels straight into ``self.grad``."""

    def __init__(self, model: JLForCTC):
        super().__init__()
        self.model = model
        self.plist = ordered_trainables(model)
        if not self.plist:
            raise ValueError("no trainable parameters: call model.freeze_base_model() first")
        dev = self.plist[0].device
        if dev.type != "cuda":
            raise RuntimeError("FlatAdapterParams needs the model on a CUDA device")
        self.layout = BucketLayout(self.plist)
        self.offset = self.layout.offset
        self.total, self.num_params = self.layout.total, self.layout.num_params
        off = self.total
        self.param = torch.zeros(off, dtype=F32, device=dev)
        self.grad = torch.zeros(off, dtype=F32, device=dev)
        self.exp_avg = torch.zeros(off, dtype=F32, device=dev)
        self.exp_avg_sq = torch.zeros(off, dtype=F32, device=dev)
        self.bf16 = torch.zeros(off, dtype=BF16, device=dev)
        with torch.no_grad():
            for p in self.plist:
                v = self._view(self.param, p)
                v.copy_(p.data)
                p.data = v
                p.grad = self._view(self.grad, p)
        ops.cast_bf16(self.param, out=self.bf16)
        self.step_count = 0
        self.comm = None          # optional JLComm (C-ABI NCCL communicator); None → torch.distributed, if initialised
        model.encoder.engine(model.lm_head).flat = self

    def _view(self, buf: torch.Tensor, p: torch.Tensor) -> torch.Tensor:
        return self.layout.view(buf, p)

    def _cat(self, buf: torch.Tensor, ps: Sequence[torch.Tensor]) -> Optional[torch.Tensor]:
        return self.layout.cat(buf, ps)

    # ---- views the engine asks for
    def bf16_view(self, p):
        return self._view(self.bf16, p) if id(p) in self.offset else None

    def bf16_cat_view(self, ps):
        return self._cat(self.bf16, ps)

    def f32_cat_view(self, ps):
        return self._cat(self.param, ps)

    # ---- GradSink interface
    def out(self, p, k=None):
        gt = self._view(self.grad, p)
        return gt if k is None else gt[k]

    def out_cat(self, ps):
        v = self._cat(self.grad, ps)
        return v if v is not None else super().out_cat(ps)

    def scatter_cat(self, ps, cat):
        if self._cat(self.grad, ps) is None:
            off = 0
            for p in ps:
                self._view(self.grad, p).copy_(cat[off: off + p.shape[0]])
                off += p.shape[0]

    # ---- collective + optimizer
    def world_size(self) -> int:
        if self.comm is not None:
            return self.comm.world
        return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1

    def allreduce(self) -> None:
        """Sum of the gradient bucket over ranks — the single collective of the fine-tune step (NCCL over NVLink):
        ``jl_comm_allreduce`` when a ``JLComm`` is attached, else ``torch.distributed.all_reduce`` (also the gloo path
        of the CPU tests)."""
        if self.comm is not None:
            self.comm.allreduce_(self.grad)
        elif dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            dist.all_reduce(self.grad, op=dist.ReduceOp.SUM)

    def adamw_step(self, lr: float, beta1: float = 0.9, beta2: float = 0.999, eps: float = 1e-8, weight_decay: float = 0.01) -> None:
        world = self.world_size()
        self.step_count += 1
        ops.adamw_(self.param, self.grad, self.exp_avg, self.exp_avg_sq, self.step_count, lr, beta1, beta2, eps, weight_decay,
                   grad_scale=1.0 / world, param_bf16=self.bf16)


def token_lengths(cfg, num_samples: torch.Tensor) -> torch.Tensor:
    """Encoder frames T' per utterance from its sample count (host tensor), for either front end."""
    ns = num_samples.to(torch.int64)
    if cfg.front_end == "wav2vec2":
        return wav2vec2_lengths(ns, cfg.conv_kernel, cfg.conv_stride).clamp_min(0).to(I32)
    frames = torch.where(ns < 400, torch.zeros_like(ns), (ns - 400) // 160 + 1)
    return subsampled_length(frames).to(I32)


def _dialect_key(dialect):
    """Hashable form of a dialect argument: an int, or a tuple of per-utterance ids."""
    if isinstance(dialect, int):
        return dialect
    return tuple(int(k) for k in (dialect.tolist() if torch.is_tensor(dialect) else dialect))


def shard_utterances(num_frames_per_utt: Sequence[int], world: int) -> List[List[int]]:
    """Length-sorted round-robin assignment of utterance indices to ranks so every rank gets ≈ equal total frames
    (mixed-length config 4).  Deterministic; returns one index list per rank."""
    order = sorted(range(len(num_frames_per_utt)), key=lambda i: (-num_frames_per_utt[i], i))
    shards: List[List[int]] = [[] for _ in range(world)]
    loads = [0] * world
    for i in order:
        r = min(range(world), key=lambda j: (loads[j], j))
        shards[r].append(i)
        loads[r] += num_frames_per_utt[i]
    return shards


class AdapterTrainer:
    """One fine-tune step = H2D(waveforms, labels) → [mel+CMVN → encoder → lm_head → CTC loss+grad → adapter-only
    backward] → all-reduce(adapter grads) → fused AdamW → D2H(loss).  The bracketed part is captured in a CUDA graph
    per (batch, samples, label length) shape."""

    def __init__(self, model: JLForCTC, lr: float = 1e-4, weight_decay: float = 0.01, use_cuda_graph: bool = True, comm="auto"):
        """``comm``: a ``JLComm``; ``"auto"`` (default) builds one from the initialised torch.distributed group when the
        world has more than one rank; ``"torch"`` leaves the all-reduce to ``torch.distributed``; ``None`` = single rank."""
        self.model = model
        self.cfg = model.config
        self.flat = FlatAdapterParams(model)
        if comm == "auto":
            multi = dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
            comm = JLComm.from_torch_distributed() if multi else None
        elif comm == "torch":
            comm = None
        self.flat.comm = comm
        self.eng = model.encoder.engine(model.lm_head)
        self.fe = JLFeatureExtractor(device=self.flat.param.device)
        self.lr, self.weight_decay = lr, weight_decay
        self.use_cuda_graph = use_cuda_graph
        self._graphs: Dict[tuple, dict] = {}
        self._stage: Dict[tuple, dict] = {}
        self._staged = None
        self._copy_stream = None
        self.launches_per_step = 0

    def _body(self, wave, nsamp, lengths, labels, max_frames, dialect=0):
        if self.cfg.front_end == "wav2vec2":
            st = self.eng.forward(wave, lengths, training=True, dialect=dialect, want_logits=True, sample_lengths=nsamp)
        else:
            feats = self.fe.extract_device(wave, nsamp, max_frames, return_bf16=True)
            st = self.eng.forward(feats["input_features_bf16"], lengths, training=True, dialect=dialect, want_logits=True)
        b, t = st.b, st.t
        logits = st.logits.view(b, t, self.cfg.vocab_size)
        loss, nll, grad = ops.ctc_loss(logits, labels, lengths, blank=self.cfg.pad_token_id, reduction=self.cfg.ctc_loss_reduction,
                                       zero_infinity=self.cfg.ctc_zero_infinity, want_grad=True, grad_dtype=BF16)
        self.eng.backward(st, grad.view(b * t, -1), self.flat)
        return loss

    def _static(self, b: int, n: int, s: int, dialect=0) -> dict:
        key = (b, n, s, dialect)                 # the dialect runs are host-side structure baked into the captured graph
        ent = self._graphs.get(key)
        if ent is not None:
            return ent
        dev = self.flat.param.device
        ent = {
            "wave": torch.zeros((b, n), dtype=F32, device=dev),
            "nsamp": torch.full((b,), n, dtype=I32, device=dev),
            "lengths": torch.ones((b,), dtype=I32, device=dev),
            "labels": torch.full((b, s), -100, dtype=I32, device=dev),
            "max_frames": max(num_frames(n), 1),
            "dialect": dialect,
            "graph": None,
            "loss": None,
        }
        self._graphs[key] = ent
        return ent

    def _token_lengths(self, num_samples: torch.Tensor) -> torch.Tensor:
        return token_lengths(self.cfg, num_samples)

    def submit(self, wave: torch.Tensor, num_samples: torch.Tensor, labels: torch.Tensor, dialect=0) -> None:
        """Stage the NEXT batch while the current step is still running: the host → device copies (pinned host memory) go to
        device staging buffers on a dedicated copy stream, so a following ``step()`` without arguments only pays a
        device-to-device copy into the graph's static inputs.  What a prefetching data loader does for the reference's
        trainer; the copy of batch i+1 overlaps the kernels of batch i."""
        b, n = wave.shape
        s = labels.shape[1]
        dev = self.flat.param.device
        key = (b, n, s)
        st = self._stage.get(key)
        if st is None:
            st = {"wave": torch.empty((b, n), dtype=F32, device=dev), "nsamp": torch.empty((b,), dtype=I32, device=dev),
                  "lengths": torch.empty((b,), dtype=I32, device=dev), "labels": torch.empty((b, s), dtype=I32, device=dev),
                  "free": None}
            self._stage[key] = st
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(device=dev)
        cs = self._copy_stream
        if st["free"] is not None:
            cs.wait_event(st["free"])            # the previous step may still be reading the staging buffers
        lengths = self._token_lengths(num_samples)
        with torch.cuda.stream(cs):
            st["wave"].copy_(wave, non_blocking=True)
            st["nsamp"].copy_(num_samples, non_blocking=True)
            st["lengths"].copy_(lengths, non_blocking=True)
            st["labels"].copy_(labels, non_blocking=True)
            ready = torch.cuda.Event()
            ready.record(cs)
        self._staged = (st, ready, b, n, s, _dialect_key(dialect))

    def step(self, wave: Optional[torch.Tensor] = None, num_samples: Optional[torch.Tensor] = None, labels: Optional[torch.Tensor] = None,
             dialect=0) -> torch.Tensor:
        """wave [B, N] fp32 (pinned host or device), num_samples [B] int32 (host), labels [B, S] int32 (host, negative pad),
        ``dialect``: WFAdapter factor set — one id or one id per utterance (same-dialect utterances adjacent).
        Without arguments the batch staged by ``submit()`` is consumed.
        Returns the loss as a 1-element device tensor (call ``.item()`` for the D2H read)."""
        if wave is None:
            if self._staged is None:
                raise RuntimeError("step() without arguments needs a batch staged by submit()")
            st, ready, b, n, s, dialect = self._staged
            self._staged = None
            ent = self._static(b, n, s, dialect)
            cur = torch.cuda.current_stream()
            cur.wait_event(ready)
            for k in ("wave", "nsamp", "lengths", "labels"):
                ent[k].copy_(st[k], non_blocking=True)
            st["free"] = torch.cuda.Event()
            st["free"].record(cur)
        else:
            b, n = wave.shape
            s = labels.shape[1]
            dialect = _dialect_key(dialect)
            ent = self._static(b, n, s, dialect)
            ent["wave"].copy_(wave, non_blocking=True)
            ent["nsamp"].copy_(num_samples, non_blocking=True)
            ent["lengths"].copy_(self._token_lengths(num_samples), non_blocking=True)
            ent["labels"].copy_(labels, non_blocking=True)
        args = (ent["wave"], ent["nsamp"], ent["lengths"], ent["labels"], ent["max_frames"], ent["dialect"])
        if not self.use_cuda_graph:
            L.launch_count_reset()
            loss = self._body(*args)
            self.launches_per_step = L.launch_count()
        else:
            if ent["graph"] is None:
                side = torch.cuda.Stream()
                side.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(side):
                    for _ in range(2):                      # warm-up: lazy packs, cudaFuncSetAttribute, allocator
                        L.launch_count_reset()
                        self._body(*args)
                        self.launches_per_step = L.launch_count()
                torch.cuda.current_stream().wait_stream(side)
                torch.cuda.synchronize()
                graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(graph):
                    ent["loss"] = self._body(*args)
                ent["graph"] = graph
            ent["graph"].replay()
            loss = ent["loss"]
        self.flat.allreduce()
        self.flat.adamw_step(self.lr, weight_decay=self.weight_decay)
        self._last = ent
        return loss

    def step_resident(self) -> torch.Tensor:
        """Repeat the last step on the inputs already resident in HBM (no host↔device copies): graph replay →
        all-reduce → fused AdamW.  Used by bench.py for the device-resident throughput."""
        ent = self._last
        if ent["graph"] is not None:
            ent["graph"].replay()
            loss = ent["loss"]
        else:
            loss = self._body(ent["wave"], ent["nsamp"], ent["lengths"], ent["labels"], ent["max_frames"], ent["dialect"])
        self.flat.allreduce()
        self.flat.adamw_step(self.lr, weight_decay=self.weight_decay)
        return loss

    def trace_gemms(self):
        """One eager pass of the step body with GEMM tracing on → list for ops.replay_gemm_trace()."""
        ent = self._last
        ops.GEMM_TRACE = []
        try:
            self._body(ent["wave"], ent["nsamp"], ent["lengths"], ent["labels"], ent["max_frames"], ent["dialect"])
            trace = ops.GEMM_TRACE
        finally:
            ops.GEMM_TRACE = None
        return trace


class Transcriber:
    """Inference: H2D(waveforms) → [mel+CMVN → encoder → lm_head → greedy collapse] → D2H(token ids), graph-captured
    per (batch, samples) shape."""

    def __init__(self, model: JLForCTC, use_cuda_graph: bool = True):
        self.model = model
        self.cfg = model.config
        self.eng = model.encoder.engine(model.lm_head)
        self.dev = next(model.parameters()).device
        self.fe = JLFeatureExtractor(device=self.dev)
        self.use_cuda_graph = use_cuda_graph
        self._graphs: Dict[tuple, dict] = {}
        self.launches_per_step = 0

    def _body(self, wave, nsamp, lengths, max_frames, dialect=0):
        if self.cfg.front_end == "wav2vec2":
            st = self.eng.forward(wave, lengths, training=False, dialect=dialect, want_logits=True, sample_lengths=nsamp)
        else:
            feats = self.fe.extract_device(wave, nsamp, max_frames, return_bf16=True)
            st = self.eng.forward(feats["input_features_bf16"], lengths, training=False, dialect=dialect, want_logits=True)
        logits = st.logits.view(st.b, st.t, self.cfg.vocab_size)
        ids, n, _ = ops.ctc_greedy(logits, lengths, blank=self.cfg.pad_token_id)
        return ids, n

    @torch.no_grad()
    def __call__(self, wave: torch.Tensor, num_samples: torch.Tensor, dialect=0):
        """wave [B, N] fp32 (pinned host or device), num_samples [B] int32 host → (ids [B, T'] int32 device, lengths [B]).
        ``dialect``: WFAdapter factor set — one id or one id per utterance (same-dialect utterances adjacent)."""
        b, n = wave.shape
        dialect = _dialect_key(dialect)
        key = (b, n, dialect)
        ent = self._graphs.get(key)
        if ent is None:
            ent = {"dialect": dialect, "wave": torch.zeros((b, n), dtype=F32, device=self.dev), "nsamp": torch.full((b,), n, dtype=I32, device=self.dev),
                   "lengths": torch.ones((b,), dtype=I32, device=self.dev), "max_frames": max(num_frames(n), 1), "graph": None, "out": None}
            self._graphs[key] = ent
        ent["wave"].copy_(wave, non_blocking=True)
        ent["nsamp"].copy_(num_samples, non_blocking=True)
        ent["lengths"].copy_(token_lengths(self.cfg, num_samples), non_blocking=True)
        args = (ent["wave"], ent["nsamp"], ent["lengths"], ent["max_frames"], ent["dialect"])
        if not self.use_cuda_graph:
            L.launch_count_reset()
            out = self._body(*args)
            self.launches_per_step = L.launch_count()
            self._last = ent
            return out
        if ent["graph"] is None:
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(2):
                    L.launch_count_reset()
                    self._body(*args)
                    self.launches_per_step = L.launch_count()
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                ent["out"] = self._body(*args)
            ent["graph"] = graph
        ent["graph"].replay()
        self._last = ent
        return ent["out"]

    def run_resident(self):
        ent = self._last
        if ent["graph"] is not None:
            ent["graph"].replay()
            return ent["out"]
        return self._body(ent["wave"], ent["nsamp"], ent["lengths"], ent["max_frames"], ent["dialect"])
